#!/bin/bash
# Developer helper: build a complete kernel-variant library  variants_<tag>.so  (git-ignored; it travels to the GPU box).
#   bash tools/build_variant.sh <tag> "<extra nvcc -D flags>"
# Select it at run time with LIGHTHAND_B200_LIB=$PWD/variants_<tag>.so (tools/run_variants.sh does that).
set -e
tag=$1; shift
cd "$(dirname "$0")/../lighthand_b200/csrc"
mkdir -p build/var_$tag
for f in capi heatmap metrics warp_normalize warp_tma frei_crop color_jitter pipeline comm jpeg_decode; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $@ -c $f.cu -o build/var_$tag/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants_$tag.so build/var_$tag/*.o -cudart static -ldl
echo built variants_$tag.so
