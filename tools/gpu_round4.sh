#!/bin/bash
# Final evidence run of round 2 (one GPU): tests, smoke, bench lines, launch list, ncu --set full, sanitizers.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 ) > gpurun_out/r2d_pytest_gpu.txt; tail -n 4 gpurun_out/r2d_pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 8 ) > gpurun_out/r2d_smoke.txt; tail -n 3 gpurun_out/r2d_smoke.txt
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -n 1 ) > gpurun_out/r2d_bench.json; cut -c1-300 gpurun_out/r2d_bench.json; tail -n 3 gpurun_out/bench.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -n 1 ) > gpurun_out/r2d_bench_reference_arm.json; cut -c1-300 gpurun_out/r2d_bench_reference_arm.json
for k in 40 200; do ( timeout 600 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -n 1 ) > gpurun_out/r2d_bench_steps$k.json; python -c "import json;j=json.load(open('gpurun_out/r2d_bench_steps$k.json'));print($k, j['ms_per_step'], j['value'], j['clocks'].get('in_kernel_sm_mhz'))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r2d_launches.csv \
   python bench.py --steps 2 --warmup 3 --preheat-s 0 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_kernel" -s 6 -c 3 -f -o gpurun_out/r2d_full \
   python tools/one_forward.py 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc $?"; ls -la gpurun_out/r2d_full.ncu-rep
( timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py bwd 2>&1 | tail -n 6 ) > gpurun_out/r2d_synccheck.txt; tail -n 2 gpurun_out/r2d_synccheck.txt
( timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py bwd 2>&1 | tail -n 6 ) > gpurun_out/r2d_memcheck.txt; tail -n 2 gpurun_out/r2d_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_target.py > gpurun_out/racecheck_full.txt 2>&1
python - <<'PY'
import re, collections
txt = open('gpurun_out/racecheck_full.txt').read()
blocks = txt.split('========= Error: ')[1:]
agg = collections.Counter()
for b in blocks:
    lines = b.split('\n')
    kind = re.sub(r' at __shared__ 0x[0-9a-f]+ in block \(\d+,\d+,\d+\) :', '', lines[0]).strip()
    w = re.sub(r'Thread \(\d+,\d+,\d+\)', 'Thread', lines[1].replace('=========', '').strip())[:150] if len(lines) > 1 else ''
    r = re.sub(r'Thread \(\d+,\d+,\d+\)', 'Thread', lines[2].replace('=========', '').strip())[:150] if len(lines) > 2 else ''
    agg[(kind, w, r)] += 1
out = [f"racecheck on tools/sanitize_target.py: {len(blocks)} reports, grouped by (hazard, writer, reader):"]
for (k, w, r), c in agg.most_common():
    out.append(f"{c:7d}  {k}\n           {w}\n           {r}")
tail = [l for l in txt.split('\n') if 'RACECHECK SUMMARY' in l or 'sanitize target ok' in l]
out += tail
open('gpurun_out/r2d_racecheck_summary.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out)[:3000])
PY
rm -f gpurun_out/racecheck_full.txt
