#!/bin/bash
# round 2, tile pass B: parity first, then the bench shapes in every compiled tile shape (MB_TILE_CFG) against the warp-per-cell scatter (MB_SORT_TILE=0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2m}
timeout 900 python -m pytest tests/test_gpu_parity_core.py -x -q -m gpu -k "tile" > gpurun_out/${TAG}_pytest_tile.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_tile.log
tail -3 gpurun_out/${TAG}_pytest_tile.log
for sc in same-dx published-dx; do
  for v in "1 2"; do
    set -- $v
    MB_TILE_DEBUG=1 MB_SORT_TILE=$1 MB_TILE_CFG=$2 timeout 300 python bench.py --scaling $sc --no-others --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 5 \
      > gpurun_out/${TAG}_${sc}_t$1_c$2.json 2> gpurun_out/${TAG}_${sc}_t$1_c$2.err
    echo "$sc tile=$1 cfg=$2"; python profiles/show_bench.py gpurun_out/${TAG}_${sc}_t$1_c$2.json 2>&1 | head -2; grep "k_band_tile CTA" gpurun_out/${TAG}_${sc}_t$1_c$2.err | tail -2
  done
done
