#!/bin/bash
mkdir -p gpurun_out
# ncu --set full of one merged-MLP launch (3rd of the process: warm) with source correlation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 2 -c 1 -o gpurun_out/r2_mlp python tools/one_forward.py 4 > gpurun_out/ncu_mlp.log 2>&1
echo "ncu mlp rc $?"; tail -n 3 gpurun_out/ncu_mlp.log
GLOM_B200_SPLIT_MLP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 2 -o gpurun_out/r2_split python tools/one_forward.py 4 > gpurun_out/ncu_split.log 2>&1
echo "ncu split rc $?"; tail -n 3 gpurun_out/ncu_split.log
ls -la gpurun_out/*.ncu-rep
