#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "GLOM_B200_SPLIT_MLP=0 GLOM_B200_L2_PERSIST_MB=64" "GLOM_B200_SPLIT_MLP=0 GLOM_B200_L2_PERSIST_MB=96" "GLOM_B200_SPLIT_MLP=0 GLOM_B200_L2_PERSIST_MB=96 GLOM_B200_MLP_LAG=2 GLOM_B200_MLP_LAG_LO=1"; do
  echo "== $cfg"
  env $cfg timeout 300 ncu --metrics $M --clock-control none -k regex:mlp_kernel -s 2 -c 1 --csv python tools/one_forward.py 4 2>gpurun_out/ncu_err.txt | grep -E "mlp_kernel" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste -sd' '
  grep "persisting" gpurun_out/ncu_err.txt | head -1
done
