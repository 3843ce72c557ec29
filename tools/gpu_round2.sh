#!/bin/bash
# Round-2 GPU session: parity tests, smoke, bench (+ reference arm), optional sanitizers / ncu.  Outputs: gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "nobench_tests" ]; then
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 40 ) > gpurun_out/pytest_gpu.txt
echo "--- pytest done"; tail -n 15 gpurun_out/pytest_gpu.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 12 ) > gpurun_out/smoke.txt; cat gpurun_out/smoke.txt
fi
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -n 3 ) > gpurun_out/bench.json; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
for a in "$@"; do
  case $a in
    ref) ( timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -n 2 ) > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json ;;
    steps) for k in 40 200; do ( timeout 600 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -n 1 ) > gpurun_out/bench_steps$k.json; python -c "import json;j=json.load(open('gpurun_out/bench_steps$k.json'));print($k, j['ms_per_step'], j['value'], j['clocks'].get('device_sm_mhz_after'))"; done ;;
    sanitize)
      ( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python __graft_entry__.py smoke 2>&1 | tail -n 30 ) > gpurun_out/racecheck.txt; tail -n 6 gpurun_out/racecheck.txt
      ( timeout 900 compute-sanitizer --tool synccheck python __graft_entry__.py smoke 2>&1 | tail -n 30 ) > gpurun_out/synccheck.txt; tail -n 6 gpurun_out/synccheck.txt
      ( timeout 900 compute-sanitizer --tool memcheck python __graft_entry__.py smoke 2>&1 | tail -n 30 ) > gpurun_out/memcheck.txt; tail -n 6 gpurun_out/memcheck.txt ;;
    ncu)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
         python bench.py --steps 2 --warmup 3 --preheat-s 0 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
      echo "ncu rc $?"; tail -n 3 gpurun_out/launches.csv ;;
  esac
done
