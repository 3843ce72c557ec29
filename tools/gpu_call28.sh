#!/bin/bash
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for o in 0 1; do
  echo "PAIR_SYNC=$o"
  GLOM_B200_K2_PAIR_SYNC=$o timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_kernel" -s 4 -c 2 --csv python tools/one_forward.py 4 2>/dev/null | grep "gemm_kernel<1" | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done
