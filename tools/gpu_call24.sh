#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python tools/fuzz_shapes.py 7 40 2>&1 | tail -n 45 ) > gpurun_out/r2d_fuzz_forward.txt; tail -n 12 gpurun_out/r2d_fuzz_forward.txt | cut -c1-260
( timeout 900 python tools/fuzz_grads.py 3 12 2>&1 | tail -n 16 ) > gpurun_out/r2d_fuzz_backward.txt; tail -n 4 gpurun_out/r2d_fuzz_backward.txt | cut -c1-260
