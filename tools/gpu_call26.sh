#!/bin/bash
mkdir -p gpurun_out
ab() {
  ( env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_ab.err | tail -n 1 ) > gpurun_out/bench_ab.json
  python - "$*" <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/bench_ab.json'))
    print(sys.argv[1][-40:], 'ms/step', round(j['ms_per_step'],3), "waits", {k:list(v.values()) for k,v in j["clocks"].get("block0_wait_fractions",{}).items()}, {k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in j['roofline']['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
}
ab GLOM_B200_WAIT_COUNTERS=1
ab X=1
( timeout 300 python -m pytest tests -m gpu -x -q -k "in_kernel or golden" 2>&1 | tail -n 3 )
