import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from reak_b200 import kte_batch_propagator, presets
p = kte_batch_propagator(presets.make("crs6"))
rng = np.random.default_rng(0)
for N, J in ((256, 50), (4096, 50), (65536, 50)):
    x = torch.from_numpy(rng.uniform(-.5, .5, (N, 12))).cuda(); goal = x + 0.3
    ub = torch.from_numpy(rng.uniform(-1, 1, (N, 6))).cuda(); up = ub * 0.5
    g = torch.from_numpy(rng.uniform(-4, 4, (N, 6, 12))).cuda()
    useq = torch.from_numpy(rng.uniform(-1, 1, (N, J, 6))).cuda()
    for name, fn in (("steer_feedback", lambda: p.steer_feedback(x, goal, ub, g, up, 1e-2, 1e-3, 10, J, 1e-9)),
                     ("rollout J intervals", lambda: p.rollout(x, useq, 1e-3, 10, scheme="rk4", want_traj=True)),
                     ("rollout 1 interval", lambda: p.get_next_states(x, ub, 1e-3, 10 * J))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print("N=%6d J=%3d %-20s  device %.3f ms  wall %.3f ms per call" % (N, J, name, e0.elapsed_time(e1) / 5, (t1 - t0) * 1e3 / 5))
