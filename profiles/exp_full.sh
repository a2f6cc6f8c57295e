#!/bin/bash
# the whole GPU suite, the driver's default bench line, and one ncu --set full capture of the tile pass B on the published grid
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2n}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python profiles/show_bench.py gpurun_out/${TAG}_bench.json
SC=published-dx CFG=2 bash profiles/exp_ncu_tile.sh | tail -3
