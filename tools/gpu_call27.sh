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
for rep in 1 2; do
ab GLOM_B200_K2_PAIR_SYNC=0
ab GLOM_B200_K2_PAIR_SYNC=1
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for o in 0 1; do
  GLOM_B200_K2_PAIR_SYNC=$o timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_kernel<1" -s 2 -c 1 --csv python tools/one_forward.py 4 2>/dev/null | grep "gemm_kernel" | python -c "
import sys,csv
for r in csv.reader(sys.stdin): print('PAIR_SYNC=$o', r[-3], r[-1])"
done
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 )
