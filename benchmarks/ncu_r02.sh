#!/bin/bash
# ncu evidence of round 2 (run under gpurun on ONE GPU, after the same commands have exited 0 without ncu).
#   launch list: plain (non-conditional) graphs -- ncu cannot profile kernel nodes of graphs that contain conditional nodes
#   full captures: the dominant kernel, the per-Newton assembly kernels and the fused coarse kernel, with source
set -x
out=${1:-gpurun_out}
python bench.py --steps 3 --warmup 12 --no-cpu-baseline --no-roofline --no-e2e > $out/r02_ncu_plain.json 2>/dev/null || exit 1
GLIMS_NO_COND_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv \
    --log-file $out/r02_launches.csv python bench.py --steps 3 --warmup 12 --no-cpu-baseline --no-roofline --no-e2e > $out/r02_ncu1.log 2>&1
for k in k_cc_rows k_spmv32_row_cheb k_fu k_amg_fused; do
  GLIMS_NO_COND_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 \
      -o $out/r02_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e > $out/r02_ncu_$k.log 2>&1
done
ls -la $out/*.ncu-rep
