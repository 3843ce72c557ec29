#!/bin/bash
# Build the engine library of another git revision into <out.so> for same-box A/B timing:
#   tools/build_variant.sh <git-rev> glom_pytorch_b200/libglom_b200_A.so
#   GLOM_B200_LIB=glom_pytorch_b200/libglom_b200_A.so python tools/diag.py timing
set -e
rev=$1; out=$2; tmp=$(mktemp -d)
git archive "$rev" glom_pytorch_b200/csrc include | tar -x -C "$tmp"
objs=""
for f in glom_api simt_kernels tc_kernels bwd_kernels tc_bwd_kernels; do
  [ -f "$tmp/glom_pytorch_b200/csrc/$f.cu" ] || continue
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static \
       -c "$tmp/glom_pytorch_b200/csrc/$f.cu" -o "$tmp/$f.o" &
  objs="$objs $tmp/$f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC $objs -o "$out"
rm -rf "$tmp"; echo "$out"
