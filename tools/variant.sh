#!/bin/bash
# developer tool: build a kernel variant into a scratch .so and time it
# usage: tools/variant.sh NAME "-DSPX_CFG32=48,6,8,32,128,6" sizes
set -e
name=$1; defs=$2
out=_variants/$name; mkdir -p $out
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $defs -c subpixal_b200/csrc/spx_cc.cu -o $out/spx_cc.o 2> $out/ptxas.log
grep -E "spill|Used" $out/ptxas.log | paste - - | sed 's/ptxas info    : //g' | cut -c1-160
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/lib.so $out/spx_cc.o subpixal_b200/_native/spx_api.o subpixal_b200/_native/spx_gather.o subpixal_b200/_native/spx_peak.o subpixal_b200/_native/spx_segbox.o subpixal_b200/_native/spx_blot.o subpixal_b200/_native/spx_ingest.o; rm -f $out/spx_cc.o
