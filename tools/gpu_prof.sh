#!/bin/bash
# Developer tool: one ncu --set full capture of the rollout kernel (after a plain run has exited 0).
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $out/bench_prof_$tag.json 2> $out/bench_prof_$tag.err || { echo "plain run failed"; tail -5 $out/bench_prof_$tag.err; exit 1; }
cat $out/bench_prof_$tag.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -s 3 -c 1 -f -o $out/prof_$tag \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"
