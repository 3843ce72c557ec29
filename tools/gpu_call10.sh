#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 30 ) > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
for hp in 0 1 2 10 12; do
  GLOM_B200_MLP_HPOL=$hp timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mlp_kernel -s 2 -c 2 --csv python tools/one_forward.py 4 2>/dev/null | grep -E "mlp_kernel" | awk -F'","' -v hp=$hp '{print "HPOL=" hp, $(NF-2), $(NF)}' | tr -d '"' | paste -sd' ' 
done
ab() {
  ( env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_ab.err | tail -n 1 ) > gpurun_out/bench_ab.json
  python - "$*" <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/bench_ab.json'))
    print(sys.argv[1], 'ms/step', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'clk', round(j['clocks'].get('device_sm_mhz_after') or 0), {k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in j['roofline']['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
}
ab GLOM_B200_SPLIT_MLP=1
ab GLOM_B200_MLP_DELAY=20 GLOM_B200_MLP_HPOL=0
ab GLOM_B200_MLP_DELAY=20 GLOM_B200_MLP_HPOL=2
ab GLOM_B200_MLP_DELAY=20 GLOM_B200_MLP_HPOL=12
