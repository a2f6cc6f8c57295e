#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_core.py tests/test_gpu_statistics.py -q -m gpu -k "ntc or ensemble" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python profiles/show_bench.py gpurun_out/r2_bench_1gpu.json 2>/dev/null
