#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -k "edge or key_passes or large_magnitude" 2>&1 | tail -n 30 ) > gpurun_out/pytest_edge.txt; tail -25 gpurun_out/pytest_edge.txt
