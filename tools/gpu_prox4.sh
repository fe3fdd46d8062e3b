#!/bin/bash
# Developer tool: one-launch checked steering, variants of the generated kernel (RKB_EXP_CHK), then the proximity tests.
out=gpurun_out
mkdir -p $out
bash tools/gpu_health.sh || exit 0
for v in 0 1 2 3; do echo "== RKB_EXP_CHK=$v"; RKB_EXP_CHK=$v timeout 150 python tools/time_steer_checked.py crs6 2>&1 | tail -3; done
timeout 500 python -m pytest tests/test_proximity.py -m gpu -x -q -o faulthandler_timeout=150 2>&1 | tail -8
