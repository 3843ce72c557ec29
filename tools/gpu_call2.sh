#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "merged_mlp" 2>&1 | tail -n 30 ) > gpurun_out/pytest_mlp.txt; cat gpurun_out/pytest_mlp.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 30 ) > gpurun_out/pytest_gpu.txt; tail -n 12 gpurun_out/pytest_gpu.txt
for mode in 1 0 1 0; do
  ( GLOM_B200_SPLIT_MLP=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_ab.err | tail -n 1 ) > gpurun_out/bench_split$mode.json
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/bench_split$mode.json'))
    print('split=$mode', 'ms/step', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'instr', round(j['roofline']['instrumented_ms_per_step'],3), 'clk', j['clocks'].get('device_sm_mhz_after'), {k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in j['roofline']['kernels'].items()})
except Exception as e:
    print('split=$mode failed', e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
done
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 100000 python __graft_entry__.py smoke > gpurun_out/racecheck_full.txt 2>&1 ); grep -c "hazard detected" gpurun_out/racecheck_full.txt; tail -n 3 gpurun_out/racecheck_full.txt
python - <<'PY'
import re,collections
t=open('gpurun_out/racecheck_full.txt').read()
blocks=t.split('========= Error:')
c=collections.Counter()
for b in blocks[1:]:
    kind=b.split('\n')[0].strip()[:60]
    w=re.search(r'Write Thread.*? at (.*)',b); r=re.search(r'Read Thread.*? at (.*)',b)
    def loc(m):
        if not m: return '?'
        s=m.group(1)
        mm=re.search(r'(\w+::\w+)\(.*?(in \S+:\d+)?$',s)
        f=re.search(r'in (\S+:\d+)',s)
        k=re.search(r'(glom::\w+)',s)
        return (k.group(1) if k else '?')+' '+(f.group(1) if f else '')
    c[(kind,loc(w),loc(r))]+=1
with open('gpurun_out/racecheck_summary.txt','w') as f:
    for k,v in c.most_common(): f.write(f'{v:6d}  {k}\n')
print(open('gpurun_out/racecheck_summary.txt').read()[:3000])
PY
head -c 40000000 gpurun_out/racecheck_full.txt > gpurun_out/racecheck_full_head.txt; rm -f gpurun_out/racecheck_full.txt
