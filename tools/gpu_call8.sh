#!/bin/bash
mkdir -p gpurun_out
( timeout 300 tools/bin/epilogue_probe 2>&1 ) > gpurun_out/epilogue_probe.txt; cat gpurun_out/epilogue_probe.txt
