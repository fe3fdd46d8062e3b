#!/usr/bin/env python
"""Developer tool: time rkb_min_distance (CRS arm against the MD148 lab, device-resident states).  The reference's
findMinimumDistance on a host core is timed by bench.py's cpu_baseline leg (other_configs / proximity)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    from reak_b200 import proximity as px
    name = sys.argv[1] if len(sys.argv) > 1 else "crs6"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    s = presets.make(name)
    p = kte_batch_propagator(s)
    robot, lab = presets.crs_proxy_models(s, track=(name == "crs7"))
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    rng = np.random.default_rng(1)
    x = rng.uniform(-3, 3, (n, p.nx))
    dx = torch.from_numpy(x).cuda()
    h = p.proxy_handle(pair)
    h.set_option(h.OPT_AUTO_SPECIALIZE, 0)

    def timed(label):
        ms = []
        for k in range(7):
            d, f = p.get_min_distances(pair, dx, with_points=False)
            ms.append(p.last_kernel_ms())
        t = min(ms[2:])
        print("%s min_distance (25 finders) %s n=%d  %.3f ms  %.3g states/s  colliding %.1f %%" % (name, label, n, t, n / t * 1e3, 100.0 * (d < 0).double().mean().item()))
        return d, f

    d0, f0 = timed("interpreter")
    blocks = [int(b) for b in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    for b in blocks:   # generated kernels (rkb_proxy_specialize), compiled for b CTAs per SM (0: the library's default)
        if b:
            h.set_option(h.OPT_MIN_BLOCKS, b)
        t0 = time.perf_counter()
        h.specialize()
        tc = time.perf_counter() - t0
        d, f = timed("generated, min_blocks %s (NVRTC %.1f s)" % (b or "default", tc))
        print("   max |d - d_interpreter| %.2e, finder differs on %d states" % ((d - d0).abs().max().item(), int((f != f0).sum().item())))
    d, f, pts = p.get_min_distances(pair, dx)
    print("%s with points               n=%d  %.3f ms" % (name, n, p.last_kernel_ms()))
    t0 = time.perf_counter()
    d, f, pts = p.get_min_distances(pair, x)
    print("%s host buffers (pageable)   n=%d  %.3f ms end to end" % (name, n, (time.perf_counter() - t0) * 1e3))


if __name__ == "__main__":
    main()
