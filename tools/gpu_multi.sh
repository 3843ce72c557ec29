#!/bin/bash
# multi-GPU evidence: bench.py under torchrun at N = $1 (default flags: incl. other configs and the training step with the NCCL gradient all-reduce)
N=$1
mkdir -p gpurun_out
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_${N}gpu.err | tail -n 1 ) > gpurun_out/r2c_bench_${N}gpu.json
python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_${N}gpu.json')); print(j['n_gpus'], j['value'], j['ms_per_step'], 'e2e', j['e2e']['ms_per_step']); print('train', j.get('train')); print({k:round(v['ms_per_step'],2) for k,v in j.get('other_configs',{}).items()})"; tail -n 3 gpurun_out/bench_${N}gpu.err
