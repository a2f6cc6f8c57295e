#!/bin/bash
# SASS evidence of the Blackwell data-movement instructions in the tile pass B (run after `make -C merzbild.jl_b200/csrc`):
# UBLKCP.S.G = cp.async.bulk global -> shared (1-D TMA load), UBLKCP.G.S = cp.async.bulk shared -> global (bulk store),
# SYNCS.* = mbarrier (init / expect_tx arrive / try_wait), FENCE.VIEW.ASYNC.S = fence.proxy.async, UTMACMDFLUSH = bulk commit_group
cd "$(dirname "$0")/.."
OBJ=merzbild.jl_b200/csrc/_build/mb_sort.o
FUN='_ZN2mb11k_band_tileILi4096ELi512ELi2048ELi4ELb1ELi1EEEvNS_8TileArgsE'
{
  echo "# cuobjdump -sass -fun k_band_tile<4096, 512, 2048, 4, true, 1>  ($(nvcc --version | tail -2 | head -1))"
  echo "# mnemonic histogram of the kernel:"
  cuobjdump -sass -fun "$FUN" $OBJ | grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+' | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40
  echo
  echo "# the TMA / mbarrier / async-proxy instructions with two lines of context:"
  cuobjdump -sass -fun "$FUN" $OBJ | grep -E '^\s+/\*[0-9a-f]+\*/' | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//' | grep -nE -B2 -A2 'UBLKCP|UTMA|SYNCS|FENCE.VIEW.ASYNC|BAR.SYNC'
} > profiles/r2_sass_k_band_tile.txt
wc -l profiles/r2_sass_k_band_tile.txt
