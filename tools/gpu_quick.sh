#!/bin/bash
# Developer tool: fast perf/parity loop on the GPU box:  gpurun --timeout 600 -- 'bash tools/gpu_quick.sh'
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in crs6 crs6_sd crs6_phys; do timeout 300 python tools/time_rollout.py $p $((1<<20)) 100 5 2>&1 | tail -1; done
timeout 300 python tools/time_rollout.py crs7 $((1<<20)) 10 5 2>&1 | tail -1
timeout 300 python tools/time_ops.py crs6 2>&1 | tail -5
