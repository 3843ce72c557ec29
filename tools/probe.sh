#!/bin/bash
for d in 0 1 2 3 4 8 12 5 13; do
  echo "== GLOM_B200_DEBUG=$d"; GLOM_B200_DEBUG=$d timeout 120 python tools/diag.py timing 2>&1 | grep -E "gemm|attn|forward" 
done
