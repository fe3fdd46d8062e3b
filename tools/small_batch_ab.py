#!/usr/bin/env python
"""Developer tool: A/B of the two mappings of a batch to the GPU — one thread per sample against one sample on a
pair of warps (RKB_OPT_SPLIT_MAX_SAMPLES) — over batch sizes 256 ... 2^20, for the RK4 rollout (BASELINE config 1's
planar chain and the 6-DOF arm) and the closed-loop steering loop.  Device time from CUDA events inside the library.

    python tools/small_batch_ab.py [--json out.json]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def best_ms(fn, prop, reps=5):
    fn()
    ms = []
    for _ in range(reps):
        fn()
        ms.append(prop.last_kernel_ms())
    return min(ms)


def main():
    import torch
    from reak_b200 import kte_batch_propagator, presets
    rows = []
    rng = np.random.default_rng(3)
    quick = "--quick" in sys.argv
    chains = [("crs6", 1000), ("planar2", 1000), ("crs7", 200)] if not quick else [("crs3", 1000), ("crs4", 1000), ("crs5", 1000), ("crs6_sd", 500)]
    for name, steps in chains:
        mk = (lambda: presets.crs_chain(n_revolute=int(name[3:]))) if name in ("crs4", "crs5") else (lambda: presets.make(name))
        solo = kte_batch_propagator(mk()).set_option("split_max_samples", 0)
        duo = kte_batch_propagator(mk()).set_option("split_max_samples", 1 << 30)
        for n in ((256, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 1 << 18, 1 << 20) if not quick else (1024, 8192, 12288, 16384)):
            k = steps if n <= 65536 else max(10, steps // 10)
            x = torch.from_numpy(rng.uniform(-1, 1, (n, solo.nx))).cuda()
            u = torch.from_numpy(rng.uniform(-1, 1, (n, solo.nu))).cuda()
            o1, o2 = torch.empty_like(x), torch.empty_like(x)
            s1 = torch.empty((n,), dtype=torch.int32, device="cuda")
            t_solo = best_ms(lambda: solo.get_next_states(x, u if solo.nu else None, 1e-3, k, out=o1, status=s1), solo)
            t_duo = best_ms(lambda: duo.get_next_states(x, u if duo.nu else None, 1e-3, k, out=o2, status=s1), duo)
            assert torch.equal(o1, o2)
            rows.append({"what": "rollout", "chain": name, "samples": n, "rk4_steps": k, "thread_per_sample_ms": t_solo, "pair_of_warps_ms": t_duo,
                         "speedup": t_solo / t_duo, "state_steps_per_s": n * k / (min(t_solo, t_duo) * 1e-3)})
            print("%-8s rollout  n=%7d steps=%4d  thread/sample %8.3f ms   pair of warps %8.3f ms   x%.2f" % (name, n, k, t_solo, t_duo, t_solo / t_duo), flush=True)
    solo = kte_batch_propagator(presets.make("crs6")).set_option("split_max_samples", 0)
    duo = kte_batch_propagator(presets.make("crs6")).set_option("split_max_samples", 1 << 30)
    for n in ((256, 1024, 4096, 16384, 65536) if not quick else ()):
        J = 50
        x = torch.from_numpy(rng.uniform(-.5, .5, (n, 12))).cuda()
        goal = x + 0.3
        ub = torch.from_numpy(rng.uniform(-1, 1, (n, 6))).cuda()
        up = ub * 0.5
        g = torch.from_numpy(rng.uniform(-4, 4, (n, 6, 12))).cuda()
        res = [None, None]

        def run(p, slot):
            res[slot] = p.steer_feedback(x, goal, ub, g, up, 1e-2, 1e-3, 10, J, 1e-9)

        t_solo = best_ms(lambda: run(solo, 0), solo)
        t_duo = best_ms(lambda: run(duo, 1), duo)
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
        rows.append({"what": "steer_feedback", "chain": "crs6", "samples": n, "intervals": J, "substeps": 10, "thread_per_sample_ms": t_solo,
                     "pair_of_warps_ms": t_duo, "speedup": t_solo / t_duo})
        print("crs6     steer    n=%7d J=%d x 10      thread/sample %8.3f ms   pair of warps %8.3f ms   x%.2f" % (n, J, t_solo, t_duo, t_solo / t_duo), flush=True)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
