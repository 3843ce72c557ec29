#!/bin/bash
# alternate A (libglom_b200_A.so) and B (current) timings on the same box
for rep in 1 2; do
  for v in A B; do
    if [ $v == A ]; then export GLOM_B200_LIB=$PWD/glom_pytorch_b200/libglom_b200_A.so; else unset GLOM_B200_LIB; fi
    echo "== $v rep $rep"; timeout 120 python tools/diag.py timing 2>&1 | grep -E "forward|gemm2_c|gemm1_g|attention"
  done
done
