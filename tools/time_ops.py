#!/usr/bin/env python
"""Developer tool: time the non-rollout entry points (device-resident buffers) and the steer batch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    name = sys.argv[1] if len(sys.argv) > 1 else "crs6_sd"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
    p = kte_batch_propagator(presets.make(name))
    rng = np.random.default_rng(1)
    dx = torch.from_numpy(rng.uniform(-1, 1, (n, p.nx))).cuda()
    du = torch.from_numpy(rng.uniform(-1, 1, (n, p.nu))).cuda()

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        ms = []
        for _ in range(reps):
            fn()
            ms.append(p.last_kernel_ms())
        return min(ms)

    t = timed(lambda: p.get_state_derivatives(dx, du))
    print("%s eval        n=%d  %.3f ms  %.3g evals/s  %.0f GB/s algorithmic" % (name, n, t, n / t * 1e3, n * (p.nx * 2 + p.nu) * 8 / t / 1e6))
    t = timed(lambda: p.get_gen_forces(dx, du))
    print("%s gen_forces  n=%d  %.3f ms  %.3g /s" % (name, n, t, n / t * 1e3))
    t = timed(lambda: p.get_mass_matrices(dx))
    print("%s mass M      n=%d  %.3f ms  %.3g /s" % (name, n, t, n / t * 1e3))
    t = timed(lambda: p.get_mass_matrices(dx, with_derivative=True))
    print("%s mass M+Mdot n=%d  %.3f ms  %.3g /s" % (name, n, t, n / t * 1e3))
    P, R, K = 4096, 256, 100
    x0 = dx[:P].contiguous()
    goal = dx[P:2 * P].contiguous()
    uu = torch.from_numpy(rng.uniform(-5, 5, (P, R, p.nu))).cuda()
    t = timed(lambda: p.steer_batch(x0, goal, uu, 1e-3, K), reps=3)
    print("%s steer_batch P=%d R=%d K=%d  %.3f ms  %.3g state-steps/s" % (name, P, R, K, t, P * R * K / t * 1e3))
    # closed-loop steering: 2^18 tuples, 20 control intervals of 10 RK4 steps (MEAQR's interval = 10 steps)
    m, J = 1 << 18, 20
    g = torch.from_numpy(rng.uniform(-4, 4, (m, p.nu, p.nx))).cuda()
    xs = dx[:m].contiguous() * 0.5
    goal = xs + torch.from_numpy(rng.uniform(-0.3, 0.3, (m, p.nx))).cuda()
    ub, up = du[:m].contiguous(), du[m:2 * m].contiguous() * 0.5
    lo, hi = -2 * np.ones(p.nu), 2 * np.ones(p.nu)
    nd = [None]

    def steer():
        nd[0] = p.steer_feedback(xs, goal, ub, g, up, 1e-2, 1e-3, 10, J, 0.25, bounds=(lo, hi), rate_bounds=(-60 * hi, 60 * hi))[2]

    t = timed(steer, reps=3)
    done = int(nd[0].sum().item())
    print("%s steer_feedback N=%d J<=%d x10 steps  %.3f ms  %.3g state-steps/s (%.1f intervals/sample on average)"
          % (name, m, J, t, done * 10 / t * 1e3, done / m))


if __name__ == "__main__":
    main()
