#!/bin/bash
# Developer tool: the end-of-round sequence — tests, bench lines, entry-point timings, launch list, ncu full, smoke.
tag=${1:-r1_final}
bash tools/gpu_round.sh $tag noprof 2>&1 | tail -34
bash tools/gpu_prof.sh $tag | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()"
