set -x
export GLIMS_NO_COND_GRAPH=1
python bench.py --steps 6 --warmup 24 --no-cpu-baseline --no-roofline > gpurun_out/r01f_plain.json 2> gpurun_out/r01f_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 6 --warmup 24 --no-cpu-baseline --no-roofline > gpurun_out/r01f_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmv32_row_cheb -s 30 -c 1 -o gpurun_out/r01f_cheb python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r01f_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tile -s 4 -c 1 -o gpurun_out/r01f_tile python benchmarks/tile_sweep.py --configs 128:12 --out gpurun_out/tile_sweep_final.json > gpurun_out/r01f_ncu3.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:k_spmv_block -s 30 -c 1 -o gpurun_out/r01f_spmv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r01f_ncu4.log 2>&1
ls -la gpurun_out/r01f_*
