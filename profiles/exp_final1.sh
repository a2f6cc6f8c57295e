#!/bin/bash
# round 2 final single-GPU evidence: bench line, launch list of the same command, ncu --set full of the step kernels on both C3 grids
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python profiles/show_bench.py gpurun_out/r2_bench_1gpu.json
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_reference_1gpu.json 2> gpurun_out/r2_bench_reference_1gpu.err
cut -c1-600 gpurun_out/r2_bench_reference_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_1gpu.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r2_launches_bench_1gpu.csv > gpurun_out/r2_launches_bench_1gpu_summary.txt 2>&1
head -12 gpurun_out/r2_launches_bench_1gpu_summary.txt
# steady-state step kernels, published grid (the tile pass B): skip sampling + 5 warm-up steps' launches by name
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_band_tile|k_convect_band|k_ntc<|k_tile_combine|k_props_cached" -s 20 -c 5 \
  -o gpurun_out/r2_full_step_published -f python bench.py --scaling published-dx --no-others --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 5 > gpurun_out/ncu_pub.log 2>&1
python profiles/extract.py gpurun_out/r2_full_step_published.ncu-rep gpurun_out/r2_ncu_full_step_published.csv && cut -c1-400 gpurun_out/r2_ncu_full_step_published.csv
