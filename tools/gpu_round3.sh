#!/bin/bash
# Round-2 evidence run: launch list, ncu --set full of the three hot kernels, sanitizers, full bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -n 1 ) > gpurun_out/r2b_bench.json; cut -c1-600 gpurun_out/r2b_bench.json; tail -n 3 gpurun_out/bench.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -n 1 ) > gpurun_out/r2b_bench_reference_arm.json; cut -c1-400 gpurun_out/r2b_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r2b_launches.csv \
   python bench.py --steps 2 --warmup 3 --preheat-s 0 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc $?"; tail -n 2 gpurun_out/r2b_launches.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_kernel" -s 6 -c 3 -f -o gpurun_out/r2b_full \
   python tools/one_forward.py 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc $?"; ls -la gpurun_out/r2b_full.ncu-rep
( timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py bwd 2>&1 | tail -n 12 ) > gpurun_out/r2b_synccheck.txt; tail -n 4 gpurun_out/r2b_synccheck.txt
( timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py bwd 2>&1 | tail -n 12 ) > gpurun_out/r2b_memcheck.txt; tail -n 4 gpurun_out/r2b_memcheck.txt
( timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_target.py 2>&1 | tail -n 40 ) > gpurun_out/r2b_racecheck.txt; tail -n 6 gpurun_out/r2b_racecheck.txt
