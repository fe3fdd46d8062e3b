#!/usr/bin/env python
"""Developer tool: launch every non-rollout entry point a few times on 2^21 states of a preset (for ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    name = sys.argv[1] if len(sys.argv) > 1 else "crs6"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
    p = kte_batch_propagator(presets.make(name))
    rng = np.random.default_rng(1)
    dx = torch.from_numpy(rng.uniform(-1, 1, (n, p.nx))).cuda()
    du = torch.from_numpy(rng.uniform(-1, 1, (n, p.nu))).cuda()
    for _ in range(3):
        p.get_state_derivatives(dx, du)
        p.get_gen_forces(dx, du)
        p.get_mass_matrices(dx)
        p.get_mass_matrices(dx, with_derivative=True)
    torch.cuda.synchronize()
    print("done", name, n)


if __name__ == "__main__":
    main()
