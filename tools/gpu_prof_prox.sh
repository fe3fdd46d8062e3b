#!/bin/bash
# Developer tool: bench line sanity + one ncu --set full capture of the proximity kernel.
#   gpurun --timeout 900 -- 'bash tools/gpu_prof_prox.sh <tag>'
tag=${1:-r1_prox}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prox_$tag.json 2> gpurun_out/bench_prox_$tag.err || tail -5 gpurun_out/bench_prox_$tag.err
python - <<P
import json
d = json.load(open("gpurun_out/bench_prox_$tag.json"))
print(d["value"], d["ms_per_step"])
for o in d["other_configs"]:
    print(o)
P
timeout 300 python tools/time_proximity.py crs6 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_proximity -s 2 -c 1 -f -o gpurun_out/prof_prox_$tag \
  python tools/time_proximity.py crs6 > gpurun_out/ncu_prox_$tag.log 2>&1
echo "ncu rc=$?"
