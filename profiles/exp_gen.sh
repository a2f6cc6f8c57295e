#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2z}
timeout 900 python -m pytest tests/test_gpu_parity_core.py tests/test_gpu_chunks.py -x -q -m gpu -k "general_sort or many_leavers or sort_general or sort_reference or squashes" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 300 python bench.py --scaling same-L --no-others --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 > gpurun_out/${TAG}_sameL.json 2> gpurun_out/${TAG}_sameL.err
python profiles/show_bench.py gpurun_out/${TAG}_sameL.json | head -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_sameL.csv \
  python bench.py --scaling same-L --no-others --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_launches_sameL.csv 2>/dev/null | head -8
