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
for o in 0 24 0 24; do ab GLOM_B200_K1_ORDER=$o; done
