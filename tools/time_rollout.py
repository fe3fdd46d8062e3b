#!/usr/bin/env python
"""Developer tool: time rkb_rollout_rk4 (device-resident buffers, CUDA events inside the library).
RKB_LIB_PATH selects the library build.  (Parity is the job of tests/: nothing outside tests/, smoke() and
bench.py's CPU legs touches oracle/.)

    python tools/time_rollout.py [preset] [n_samples] [rk4_steps] [reps] [split_max_samples]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    name = sys.argv[1] if len(sys.argv) > 1 else "crs6"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    p = kte_batch_propagator(presets.make(name))
    if len(sys.argv) > 5:
        p.set_option("split_max_samples", int(sys.argv[5]))
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (n, p.nx))
    u = rng.uniform(-1, 1, (n, p.nu))
    dx, du = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    out = torch.empty_like(dx)
    st = torch.empty((n,), dtype=torch.int32, device="cuda")
    ms = []
    for r in range(reps + 3):
        p.get_next_states(dx, du, 1e-3, steps, out=out, status=st)
        ms.append(p.last_kernel_ms())
    ms = ms[3:]
    best, mean = min(ms), sum(ms) / len(ms)
    print("%s lib=%s n=%d steps=%d serial=%s  kernel ms best %.3f mean %.3f  -> %.4g state-steps/s  status_max %d"
          % (name, os.environ.get("RKB_LIB_PATH", "default"), n, steps, p.is_serial(), best, mean, n * steps / (mean * 1e-3), int(st.max().item())))


if __name__ == "__main__":
    main()
