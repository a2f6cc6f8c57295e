#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "fp_linear or fp" 2>&1 | tail -1
timeout 300 python bench.py --config c5 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3 > gpurun_out/r3l_c5.json 2> /dev/null
python profiles/show_bench.py gpurun_out/r3l_c5.json 2>/dev/null | sed -n 1,2p
