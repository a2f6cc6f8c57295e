#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r3a}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 300 python bench.py --scaling same-L --no-others --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 > gpurun_out/${TAG}_sameL.json 2> gpurun_out/${TAG}_sameL.err
python profiles/show_bench.py gpurun_out/${TAG}_sameL.json | head -2
timeout 300 python bench.py --config c4 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
python profiles/show_bench.py gpurun_out/${TAG}_c4.json | head -2
