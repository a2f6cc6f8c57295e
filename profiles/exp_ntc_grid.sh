#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_core.py tests/test_gpu_statistics.py tests/test_gpu_chunks.py -x -q -m gpu -k "ntc or couette or ensemble or chunk" > gpurun_out/r3d_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r3d_pytest.log
for sc in same-dx published-dx; do
  timeout 300 python bench.py --scaling $sc --no-others --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 5 > gpurun_out/r3d_${sc}.json 2> /dev/null
  echo "$sc"; python profiles/show_bench.py gpurun_out/r3d_${sc}.json 2>/dev/null | sed -n 1,2p
done
timeout 300 python bench.py --config c4 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3 > gpurun_out/r3d_c4.json 2> /dev/null
echo c4; python profiles/show_bench.py gpurun_out/r3d_c4.json 2>/dev/null | sed -n 1,2p
