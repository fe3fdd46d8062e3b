#!/bin/bash
# Developer tool (round 2): parity incl. the pair-of-warps kernels, the small-batch A/B, ncu of both mappings at 1024 samples.
#   gpurun --timeout 1800 -- 'bash tools/gpu_r2b.sh'
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== A/B"
timeout 600 python tools/small_batch_ab.py --json $out/small_batch_ab.json 2>&1 | tail -50
echo "== ncu: 1024 samples x 200 steps, thread per sample, then pair of warps"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -c 1 -f -o $out/prof_r2_small_solo \
  python tools/time_rollout.py crs6 1024 200 1 0 > $out/ncu_small_solo.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -c 1 -f -o $out/prof_r2_small_duo \
  python tools/time_rollout.py crs6 1024 200 1 1000000 > $out/ncu_small_duo.log 2>&1; echo rc=$?
