#!/bin/bash
# one ncu --set full capture of the tile pass B (and of the warp-per-cell scatter for comparison) in a steady-state Couette step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SC=${SC:-same-dx}
CFG=${CFG:-2}
MB_SORT_TILE=1 MB_TILE_CFG=$CFG timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_band_tile -s 6 -c 1 \
  -o gpurun_out/r2_tile_${SC}_c${CFG} -f python bench.py --scaling $SC --no-others --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 5 --particles-per-gpu ${NP:-1.25e8} > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log
python profiles/extract.py gpurun_out/r2_tile_${SC}_c${CFG}.ncu-rep gpurun_out/r2_tile_${SC}_c${CFG}.csv && cat gpurun_out/r2_tile_${SC}_c${CFG}.csv
