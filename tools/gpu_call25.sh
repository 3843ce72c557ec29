#!/bin/bash
mkdir -p gpurun_out
ab() {
  ( env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_ab.err | tail -n 1 ) > gpurun_out/bench_ab.json
  python - "$*" <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/bench_ab.json'))
    print(sys.argv[1][-40:], 'ms/step', round(j['ms_per_step'],3), {k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in j['roofline']['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for o in 0 1 2 0 1; do
  ab GLOM_B200_K2_HPOL=$o
done
for o in 0 1 2; do
  GLOM_B200_K2_HPOL=$o timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_kernel<1" -s 2 -c 1 --csv python tools/one_forward.py 4 2>/dev/null | grep -E "gemm_kernel" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste -sd' ' | cut -c1-400
done
