#!/bin/bash
# build a variant of libmerzbild_b200.so with extra -D flags for kernel experiments: profiles/build_variant.sh NAME "-DMB_SC_U=6 ..."
# -> merzbild.jl_b200/_variants/libmb_NAME.so (use with MERZBILD_B200_LIB=...)
set -e
cd "$(dirname "$0")/../merzbild.jl_b200/csrc"
name=$1; flags=$2
out=../_variants; mkdir -p $out _build_$name
for f in *.cu; do
  o=_build_$name/${f%.cu}.o
  if [ "$f" = "mb_sort.cu" ] || [ "$f" = "mb_ntc.cu" ] || [ "$f" = "mb_fp.cu" ] || [ "$f" = "mb_props.cu" ] || [ "$f" = "mb_octree.cu" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC $flags -c $f -o $o &
  else
    cp _build/${f%.cu}.o $o
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmb_$name.so _build_$name/*.o -ldl
rm -rf _build_$name
echo built $out/libmb_$name.so
