#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_core.py tests/test_gpu_chunks.py -q -m gpu -k "band or cached or couette_loop or convect_then or chunk" 2>&1 | tail -1
timeout 300 python bench.py --no-others --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 5 > gpurun_out/r3m_samedx.json 2> /dev/null
python profiles/show_bench.py gpurun_out/r3m_samedx.json 2>/dev/null | sed -n 1,2p | cut -c1-260
