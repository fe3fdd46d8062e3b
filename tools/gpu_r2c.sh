#!/bin/bash
# Developer tool (round 2): full parity, A/B on short chains, ncu summaries of the two small-batch mappings (kept small:
# the reports are condensed on the box), one bench line.
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== A/B on short chains"
timeout 600 python tools/small_batch_ab.py --quick 2>&1 | tail -20
echo "== ncu: 1024 samples x 100 steps, thread per sample, then pair of warps"
for v in solo:0 duo:1000000; do
  tag=${v%%:*}; split=${v##*:}
  timeout 300 ncu --set full --clock-control none -k regex:serial_rollout -c 1 -f -o /tmp/prof_small_$tag \
    python tools/time_rollout.py crs6 1024 100 1 $split > $out/ncu_small_$tag.log 2>&1; echo rc=$?
  python tools/ncu_summary.py /tmp/prof_small_$tag.ncu-rep 102400 > $out/r2_small_batch_$tag.md 2>&1
done
echo "== bench"
timeout 900 python bench.py > $out/bench_r2c.json 2> $out/bench_r2c.err || { echo "bench failed"; tail -20 $out/bench_r2c.err; }
python - <<'PY'
import json
l=json.load(open("gpurun_out/bench_r2c.json"))
print({k:l[k] for k in ("value","ms_per_step")}, l["e2e"]["value"], l["roofline"]["frac"])
for o in l["other_configs"]:
    print(o.get("config"), {k:v for k,v in o.items() if k not in ("workload","cpu_baseline")}, (o.get("cpu_baseline") or {}).get("value"))
PY
