#!/bin/bash
# Developer tool: proximity tests + timing in one gpurun call.
#   gpurun --timeout 900 -- 'bash tools/gpu_prox.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proximity.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/time_proximity.py crs6 2>&1 | tail -8 | tee gpurun_out/time_proximity.txt
timeout 300 python tools/time_proximity.py crs7 2>&1 | tail -8 | tee -a gpurun_out/time_proximity.txt
