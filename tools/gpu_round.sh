#!/bin/bash
# Developer tool: one gpurun call that refreshes everything the round's evidence rests on.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
# Order matters: every profiler pass runs only after the same command has exited 0 without ncu.
tag=${1:-r1_x}
out=gpurun_out
mkdir -p $out
set -o pipefail
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench (ours)"
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; }
cat $out/bench_$tag.json
echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
cat $out/bench_ref_$tag.json
echo "== other entry points"
for p in crs6 crs6_sd crs7; do timeout 300 python tools/time_ops.py $p 2>&1 | tail -8; done
timeout 300 python tools/time_rollout.py crs6_sd $((1<<20)) 100 3 2>&1 | tail -1
timeout 300 python tools/time_rollout.py crs7 $((1<<20)) 10 3 2>&1 | tail -1
timeout 300 python tools/time_rollout.py planar2 1024 1000 3 2>&1 | tail -1
if [ "$2" != "noprof" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_launch_$tag.log 2>&1
echo "rc=$?"
echo "== ncu --set full, rollout kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -s 3 -c 1 -f -o $out/prof_$tag \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_full_$tag.log 2>&1
echo "rc=$?"
fi
