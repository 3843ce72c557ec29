#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "reused or in_kernel or tokenizer_backward or clock" 2>&1 | tail -n 25 ) > gpurun_out/pytest_new.txt; tail -20 gpurun_out/pytest_new.txt
