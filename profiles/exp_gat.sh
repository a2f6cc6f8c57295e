#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 2 4 6; do
  MB_GATHER_PER_SM=$g timeout 300 python bench.py --config c4 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3 > gpurun_out/r3g_c4_g$g.json 2> /dev/null
  echo "C4 gather CTAs/SM=$g"; python profiles/show_bench.py gpurun_out/r3g_c4_g$g.json 2>/dev/null | sed -n 2,2p
  MB_GATHER_PER_SM=$g timeout 300 python bench.py --scaling same-L --no-others --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 > gpurun_out/r3g_sameL_g$g.json 2> /dev/null
  echo "same-L gather CTAs/SM=$g"; python profiles/show_bench.py gpurun_out/r3g_sameL_g$g.json 2>/dev/null | sed -n 2,2p
done
