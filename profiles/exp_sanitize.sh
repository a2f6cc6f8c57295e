#!/bin/bash
# compute-sanitizer over the tile pass B (TMA + mbarrier pipeline): memcheck and racecheck on a few shapes of its parity test, and smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity_core.py -x -q -m gpu \
    -k "test_sort_tile_pass_b and (64-1000-2-2 or 300-250-15-2 or 900-9-2-1 or 7-5000-8-2 or 257-333-1-0)" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" gpurun_out/r2_sanitizer_$tool.log | head -8
done
