#!/usr/bin/env python
"""Developer tool: closed-loop steering with the collision test (rkb_steer_feedback_checked) — interval by interval with the
interpreter query, with the generated query, in one launch — next to the unchecked loop on the same tuples."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    from reak_b200 import proximity as px
    name = sys.argv[1] if len(sys.argv) > 1 else "crs6"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    J = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    s = presets.make(name)
    p = kte_batch_propagator(s)
    p.set_option("auto_specialize", 0)
    robot, lab = presets.crs_proxy_models(s, track=(name == "crs7"))
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    g = torch.Generator(device="cuda").manual_seed(1)

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) * (hi - lo) + lo

    x = uni((n, p.nx), -3, 3) * 0.3
    goal = x + uni((n, p.nx), -1, 1)
    ub = uni((n, p.nu), -1, 1)
    gain = uni((n, p.nu, p.nx), -2, 2)

    def run(label, pairs):
        ms, out = [], None
        for k in range(5):
            up = torch.zeros_like(ub)
            out = p.steer_feedback(x, goal, ub, gain, up, 1e-2, 1e-3, 10, J, 0.25, proxy_pairs=pairs)
            ms.append(p.last_kernel_ms())
        t = min(ms[1:])
        steps = int(out[2].sum().item()) * 10
        col = (" collided %.1f %%" % (100.0 * out[4].double().mean().item())) if pairs else ""
        print("%s %-46s n=%d J<=%d  %.3f ms  %.3g state-steps/s (%.2f intervals/tuple)%s" % (name, label, n, J, t, steps / t * 1e3, steps / 10.0 / n, col))
        return t, out

    tu, _ = run("unchecked (one launch)", None)
    t0, a = run("checked, interval by interval, interpreter query", [pair])
    p.proxy_handle(pair).specialize()
    t1, b = run("checked, interval by interval, generated query", [pair])
    p.specialize_checked_steering([pair])
    t2, c = run("checked, one launch", [pair])
    print("   one launch / unchecked = %.2f; n_done equal: %s / %s; max |x - x_interval_by_interval| %.2e" % (
        t2 / tu, bool((a[2] == c[2]).all().item()), bool((b[2] == c[2]).all().item()), (a[0] - c[0]).abs().max().item()))


if __name__ == "__main__":
    main()
