#!/bin/bash
# round 2 final single-GPU evidence: the GPU suite, smoke(), the bench line, the reference arm, the launch list of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python profiles/show_bench.py gpurun_out/r2_bench_1gpu.json 2>/dev/null
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_reference_1gpu.json 2> gpurun_out/r2_bench_reference_1gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_1gpu.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r2_launches_bench_1gpu.csv > gpurun_out/r2_launches_bench_1gpu_summary.txt 2>&1
