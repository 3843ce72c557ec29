#!/bin/bash
for rep in 1 2; do
for lb in 0 1 2 3; do
  echo "== LEVEL_BATCH=$lb rep $rep"; GLOM_B200_LEVEL_BATCH=$lb timeout 120 python tools/diag.py timing 2>&1 | grep -E "forward|gemm2_c|gemm1_g"
done; done
GLOM_B200_LEVEL_BATCH=1 timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 4
GLOM_B200_LEVEL_BATCH=2 timeout 200 python tools/power_probe.py 3 2>&1 | tail -4
GLOM_B200_LEVEL_BATCH=0 timeout 200 python tools/power_probe.py 3 2>&1 | tail -4
