#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "merged_mlp" 2>&1 | tail -n 30 ) > gpurun_out/pytest_mlp.txt; cat gpurun_out/pytest_mlp.txt
ab() {
  ( env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_ab.err | tail -n 1 ) > gpurun_out/bench_ab.json
  python - "$*" <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/bench_ab.json'))
    print(sys.argv[1], "ikclk", j["clocks"].get("in_kernel_sm_mhz"), 'ms/step', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'W', j['clocks'].get('power_w_median_under_load'), {k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in j['roofline']['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
}
ab GLOM_B200_SPLIT_MLP=0
ab GLOM_B200_SPLIT_MLP=0 GLOM_B200_MLP_LAG=12 GLOM_B200_MLP_LAG_LO=4
( GLOM_B200_MLP_DBG=1 timeout 300 python tools/one_forward.py 2>&1 | grep -E "dbg|ok" | head -40 ) > gpurun_out/mlp_dbg.txt; head -20 gpurun_out/mlp_dbg.txt
