#!/bin/bash
# ncu evidence of round 2, second pass (after the slot-pair FP16 layout, the occupancy-sized grids and the bulk-copy level-1 kernel).
# Run under gpurun on ONE GPU, after the same bench command has exited 0 without ncu.  Plain (non-conditional) graphs: ncu cannot
# profile kernel nodes of graphs that contain conditional nodes.
set -x
out=${1:-gpurun_out}
python bench.py --steps 3 --warmup 12 --no-cpu-baseline --no-roofline --no-e2e --no-verify > $out/r02b_ncu_plain.json 2>/dev/null || exit 1
GLIMS_NO_COND_GRAPH=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv \
    --log-file $out/r02b_launches.csv python bench.py --steps 3 --warmup 12 --no-cpu-baseline --no-roofline --no-e2e --no-verify > $out/r02b_ncu1.log 2>&1
for k in k_spmv32_row_cheb k_split_tma; do
  GLIMS_NO_COND_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 \
      -o $out/r02b_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-verify > $out/r02b_ncu_$k.log 2>&1
done
ls -la $out/r02b_*.ncu-rep
