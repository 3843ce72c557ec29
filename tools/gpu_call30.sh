#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 ) > gpurun_out/final_pytest_gpu.txt; tail -n 4 gpurun_out/final_pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 )
( timeout 900 python bench.py 2>gpurun_out/bench.err | tail -n 1 ) > gpurun_out/final_bench.json; python -c "
import json; j=json.load(open('gpurun_out/final_bench.json')); print(j['value'], j['ms_per_step'], j['steps'], 'e2e', j['e2e']['ms_per_step'], j['roofline']['frac'], j['cpu_baseline']['value'], {k:round(v['ms_per_step'],2) for k,v in j['other_configs'].items()}, j['train']['step_ms_max_over_ranks'])"
