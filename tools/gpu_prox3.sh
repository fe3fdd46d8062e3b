#!/bin/bash
# Developer tool: generated proximity kernels and one-launch checked steering — tests and timings.
out=gpurun_out
mkdir -p $out
bash tools/gpu_health.sh || exit 0
timeout 500 python -m pytest tests/test_proximity.py -m gpu -x -q -o faulthandler_timeout=150 2>&1 | tail -8
timeout 150 python tools/time_proximity.py crs6 $((1<<20)) 6,8,0 2>&1 | tail -12
timeout 200 python tools/time_steer_checked.py crs6 2>&1 | tail -8
timeout 200 python tools/time_steer_checked.py crs7 2>&1 | tail -8
