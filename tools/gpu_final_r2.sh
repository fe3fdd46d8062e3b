#!/bin/bash
# Developer tool: the end-of-round evidence of round 2 in one gpurun call (every profiler pass after the same command has
# exited 0 without ncu):   gpurun --timeout 1700 -- 'bash tools/gpu_final_r2.sh'
tag=r2_final
out=gpurun_out
mkdir -p $out
set -o pipefail
if [ "$1" != "notests" ]; then
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
fi
echo "== bench (ours)"
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; }
head -c 1200 $out/bench_$tag.json; echo
echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
head -c 600 $out/bench_ref_$tag.json; echo
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_launch_$tag.log 2>&1
echo "rc=$?"
echo "== ncu --set full, rollout kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -s 3 -c 1 -f -o $out/prof_$tag \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_full_$tag.log 2>&1
echo "rc=$?"
echo "== ncu --set full, nearest-neighbour scan"
timeout 120 python tools/time_nearest.py 12 1048576 4096 1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nearest_scan -s 3 -c 1 -f -o $out/prof_nn_$tag \
  python tools/time_nearest.py 12 1048576 4096 1 > $out/ncu_nn_$tag.log 2>&1
echo "rc=$?"
for a in "12 1048576 256 1" "12 1048576 65536 1" "12 1048576 4096 8" "19 1048576 4096 1" "6 1048576 4096 1" "25 1048576 4096 1"; do timeout 120 python tools/time_nearest.py $a 2>&1 | tail -1; done
echo "== proximity: interpreter and generated kernels, checked steering"
timeout 150 python tools/time_proximity.py crs6 $((1<<20)) 0 2>&1 | tail -6
for n in 1024 16384; do timeout 100 python tools/time_proximity.py crs6 $n 0 2>&1 | grep min_distance; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rkb_prox_spec_d -s 2 -c 1 -f -o $out/prof_prox_$tag \
  python tools/time_proximity.py crs6 > $out/ncu_prox_$tag.log 2>&1
echo "rc=$?"
timeout 200 python tools/time_steer_checked.py crs6 2>&1 | tail -5
timeout 200 python tools/time_steer_checked.py crs7 2>&1 | tail -5
echo "== other entry points"
for p in crs6 crs6_sd crs7; do timeout 300 python tools/time_ops.py $p 2>&1 | tail -8; done
timeout 300 python tools/time_rollout.py free_arm6 $((1<<18)) 10 3 2>&1 | tail -1
