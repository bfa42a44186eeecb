#!/bin/bash
# Regenerates the committed SASS evidence for the two hot kernels from the built library (no GPU needed).
#   bash profiles/dump_sass.sh r02
set -e
tag=${1:-r02}
lib=lighthand_b200/liblighthand_b200.so
cuobjdump -sass -fun '_ZN2lh15warp_tma_kernelILi1ELi256EEEv14CUtensorMap_stS1_NS_10WarpParamsEiii' $lib | grep -v '^\s*/\* 0x' > profiles/${tag}_sass_warp_tma_kernel_1_256.txt
cuobjdump -sass -fun '_ZN2lh19heatmap_warp_kernelILb1ELi1ELi5EEEvNS_13HeatmapParamsE' $lib | grep -v '^\s*/\* 0x' > profiles/${tag}_sass_heatmap_warp_kernel_train.txt
{
  echo "# SASS mnemonic counts ($tag; cuobjdump -sass $lib)"
  for f in profiles/${tag}_sass_warp_tma_kernel_1_256.txt profiles/${tag}_sass_heatmap_warp_kernel_train.txt; do
    echo "== $f"
    grep -o -E '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9]+\s+)?[A-Z0-9_.]+' $f | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -30
    echo "-- TMA / mbarrier / dot-product evidence:"
    grep -E 'UTMALDG|UBLKCP|SYNCS|IDP\.2A|REDUX|ATOMG|ATOM\.' $f | sed 's/^\s*//' | awk '{ $1=""; print }' | sort | uniq -c | sort -rn | head -12
  done
  echo "== registers / shared memory (cuobjdump -res-usage)"
  cuobjdump -res-usage $lib 2>/dev/null | grep -A1 -E 'warp_tma_kernelILi1ELi256|heatmap_warp_kernelILb1ELi1ELi5E' | grep -E 'Function|REG'
} > profiles/${tag}_sass_summary.txt
