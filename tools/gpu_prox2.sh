#!/bin/bash
# Developer tool: generated proximity kernels — tests, timing over min_blocks, one ncu capture.
#   gpurun --timeout 1200 -- 'bash tools/gpu_prox2.sh <tag> [notests]'
tag=${1:-r2_prox}
out=gpurun_out
mkdir -p $out
bash tools/gpu_health.sh || exit 0
if [ "$2" != "notests" ]; then
timeout 400 python -m pytest tests/test_proximity.py -m gpu -x -q -o faulthandler_timeout=120 2>&1 | tail -8
fi
timeout 150 python tools/time_proximity.py crs6 $((1<<20)) 3,4,5,6,0 2>&1 | tail -16
timeout 150 python tools/time_proximity.py crs7 $((1<<20)) 0 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rkb_prox_spec_d -s 2 -c 1 -f -o $out/prof_prox_$tag \
  python tools/time_proximity.py crs6 > $out/ncu_prox_$tag.log 2>&1
echo "ncu rc=$?"
