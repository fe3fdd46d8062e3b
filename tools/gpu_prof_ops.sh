#!/bin/bash
# Developer tool: ncu --set full of the evaluation kernels (state derivative, generalised forces, M, M + Mdot).
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/run_ops_once.py crs6 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:"serial_(eval|forces|mass)" -s 8 -c 4 -f -o $out/prof_ops_$tag \
  python tools/run_ops_once.py crs6 > $out/ncu_ops_$tag.log 2>&1
echo "ncu rc=$?"
