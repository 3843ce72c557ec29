#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "resumed or config5 or cached_workspaces or merged" 2>&1 | tail -n 25 ) > gpurun_out/pytest_new.txt; tail -20 gpurun_out/pytest_new.txt
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train 2>gpurun_out/bench.err | tail -n 1 ) > gpurun_out/bench_f3.json
python -c "
import json; j=json.load(open('gpurun_out/bench_f3.json')); print(j['ms_per_step'], {k:(round(v['ms_per_step'],2)) for k,v in j['other_configs'].items()})"; tail -3 gpurun_out/bench.err
