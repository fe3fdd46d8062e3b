#!/usr/bin/env python
"""Developer tool: general code vs run-time specialised kernels (rkb_chain_specialize) on arm geometries that are
not among the shipped shapes; also the NVRTC compile time."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    n, steps = 1 << 20, 100
    cases = {"era7": dict(n_revolute=7, axes=presets.ERA_AXES, link_offsets=presets.ERA_LINKS),
             "ssrms7": dict(n_revolute=7, axes=presets.SSRMS_AXES, link_offsets=presets.SSRMS_LINKS),
             "crs6 (shipped shape)": dict(n_revolute=6)}
    rng = np.random.default_rng(1)
    for name, kw in cases.items():
        p = kte_batch_propagator(presets.crs_chain(**kw))
        dx = torch.from_numpy(rng.uniform(-1, 1, (n, p.nx))).cuda()
        du = torch.from_numpy(rng.uniform(-1, 1, (n, p.nu))).cuda()
        out = torch.empty_like(dx)

        def run():
            ms = []
            for _ in range(4):
                p.get_next_states(dx, du, 1e-3, steps, out=out)
                ms.append(p.last_kernel_ms())
            return min(ms[1:])

        t_gen = run()
        t0 = time.time()
        p.specialize()
        t_jit = time.time() - t0
        t_spec = run()
        print("%-22s shipped kernel shape %x: %.2f ms;  specialised in %.1f s: %.2f ms  (%.2fx, %.3g state-steps/s)"
              % (name, 0 if "crs6" not in name else p.kernel_shape(), t_gen, t_jit, t_spec, t_gen / t_spec, n * steps / t_spec * 1e3))


if __name__ == "__main__":
    main()
