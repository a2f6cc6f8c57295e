#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python profiles/show_bench.py gpurun_out/r2_bench_1gpu.json 2>/dev/null | cut -c1-200
