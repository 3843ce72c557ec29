#!/bin/bash
# multi-GPU evidence: bench.py under torchrun at N = $1 (+ --train)
N=$1
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-other-configs 2>gpurun_out/bench_${N}gpu.err | tail -n 1 ) > gpurun_out/r2b_bench_${N}gpu.json; cut -c1-400 gpurun_out/r2b_bench_${N}gpu.json; tail -n 3 gpurun_out/bench_${N}gpu.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --train 2>gpurun_out/bench_train_${N}gpu.err | tail -n 1 ) > gpurun_out/r2b_bench_train_${N}gpu.json; python -c "
import json; j=json.load(open('gpurun_out/r2b_bench_train_${N}gpu.json')); print('train', j.get('train'))"; tail -n 3 gpurun_out/bench_train_${N}gpu.err
