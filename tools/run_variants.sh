#!/bin/bash
# Developer helper: time kernel-shape variants of the fast warp kernel on the GPU box.
#   build:  nvcc ... -DLH_FAST_THREADS=128 -DLH_FAST_MAXPOS=3200 -DLH_FAST_MINB=8 -c warp_normalize.cu -> variants_<tag>.so
#   run:    bash tools/run_variants.sh <tag> ...        (each .so is selected through LIGHTHAND_B200_LIB)
for v in "$@"; do echo "== $v"; LIGHTHAND_B200_LIB=$PWD/variants_$v.so python tools/microbench.py 2>&1 | tr -d "\n" | sed "s/},/},\n/g" | grep -E "warp"; done
