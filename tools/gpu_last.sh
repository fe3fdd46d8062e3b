#!/bin/bash
# Developer tool: evidence for a build that only changed the interpreter kernels, in a tight GPU budget: profiler captures of
# the two kernels the constants come from, the launch list, then the GPU tests that drive the interpreter rollouts.
out=gpurun_out
mkdir -p $out
timeout 70 ncu --set full --clock-control none --import-source on -k regex:serial_rollout -s 3 -c 1 -f -o $out/prof_r2_final \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_full_r2_final.log 2>&1
echo "rollout capture rc=$?"
timeout 50 ncu --set full --clock-control none --import-source on -k regex:rkb_prox_spec_d -s 2 -c 1 -f -o $out/prof_prox_r2_final \
  python tools/time_proximity.py crs6 > $out/ncu_prox_r2_final.log 2>&1
echo "proximity capture rc=$?"
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r2_final.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_launch_r2_final.log 2>&1
echo "launch list rc=$?"
timeout 90 python -m pytest tests/test_free_joint.py tests/test_gpu_parity.py -m gpu -x -q -k "free or rk4_one_and_many_steps or rollout_schemes_vs_oracle or long_chains or rk4_with_an_input or golden_reference" 2>&1 | tail -4
timeout 40 python tools/time_rollout.py free_arm6 262144 10 2 2>&1 | tail -1
