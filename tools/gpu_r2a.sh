#!/bin/bash
# Developer tool (round 2, first GPU pass): parity tests, the small-batch microbenchmarks, a bench line, small-batch timings.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh'
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu"
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== small-batch microbench"
timeout 120 tools/bin/small_batch_microbench 2>&1 | tee $out/small_batch_microbench.txt
echo "== bench (ours)"
timeout 900 python bench.py > $out/bench_r2a.json 2> $out/bench_r2a.err || { echo "bench failed"; tail -20 $out/bench_r2a.err; }
cat $out/bench_r2a.json
echo "== small batches"
timeout 300 python tools/small_batch.py 2>&1 | tail -12
for n in 256 1024 4096 16384; do timeout 120 python tools/time_rollout.py crs6 $n 1000 3 2>&1 | tail -1; done
timeout 120 python tools/time_rollout.py planar2 1024 1000 3 2>&1 | tail -1
