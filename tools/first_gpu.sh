#!/bin/bash
# First GPU bring-up: each stage in its own process (a trap in one does not poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/diag.txt 2>&1
for s in fp32 bf16_stages bf16_full timing; do
  echo "=== $s" >> gpurun_out/diag.txt
  timeout 300 python tools/diag.py $s >> gpurun_out/diag.txt 2>&1
  echo "exit $?" >> gpurun_out/diag.txt
done
tail -n 120 gpurun_out/diag.txt
