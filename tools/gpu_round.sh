#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 40 ) > gpurun_out/pytest_gpu.txt
echo "--- pytest done"; tail -n 15 gpurun_out/pytest_gpu.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 12 ) > gpurun_out/smoke.txt; cat gpurun_out/smoke.txt
( timeout 600 python bench.py 2>gpurun_out/bench.err | tail -n 3 ) > gpurun_out/bench.json; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 2 ) > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc $?"; tail -n 3 gpurun_out/launches.csv
fi
