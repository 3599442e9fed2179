#!/bin/bash
# developer tool (GPU box): size sweep for the in-tree library and every _variants/*/lib.so
# usage: tools/ab.sh "32,64,128" "NCC" [variant ...]
sizes=${1:-32,64,128}; types=${2:-NCC}; shift; shift
mkdir -p gpurun_out
echo "== main"; python tools/size_sweep.py $sizes $types 2>&1 | grep pairs_per_s | sed 's/^/main /' | tee gpurun_out/ab_main.txt
for v in "$@"; do
  echo "== $v"; SUBPIXAL_B200_LIB=_variants/$v/lib.so python tools/size_sweep.py $sizes $types 2>&1 | grep pairs_per_s | sed "s/^/$v /" | tee gpurun_out/ab_$v.txt
done
