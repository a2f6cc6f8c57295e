#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r3k64}
echo bulk_min=16
for v in "published-dx 1 2" "same-dx 3 2"; do set -- $v
  MB_SORT_TILE=$2 MB_TILE_CFG=$3 timeout 300 python bench.py --scaling $1 --no-others --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 5 > gpurun_out/${TAG}_$1_$2_$3.json 2> /dev/null
  echo "$1 tile=$2 cfg=$3"; python profiles/show_bench.py gpurun_out/${TAG}_$1_$2_$3.json 2>/dev/null | sed -n 2,2p | grep -oE "'sort.scatter': [0-9.]+"
done
