#!/bin/bash
# round 2: C2 merge CTA shape (threads per CTA / CTAs per SM of k_merge), launch list of the same-L step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "128 4" "64 8" "32 8" "32 12"; do set -- $cfg
  MB_MERGE_THREADS=$1 MB_MERGE_PER_SM=$2 timeout 300 python bench.py --config c2 --steps 12 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2o_c2_$1_$2.json 2> gpurun_out/r2o_c2_$1_$2.err
  echo "k_merge threads=$1 per_sm=$2"; python profiles/show_bench.py gpurun_out/r2o_c2_$1_$2.json | head -2
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_launches_sameL.csv \
  python bench.py --scaling same-L --no-others --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r2o_launches_sameL.csv | head -14
