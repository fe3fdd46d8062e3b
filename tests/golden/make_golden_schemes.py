"""Regenerates tests/golden/schemes/*.npz from the REAL reference (oracle/_ref/libreak_ref.so): the
state after every control interval when a planner drives the chain with a piecewise-constant input
sequence, one integrate() call per interval (num_int_dtnl_sys::get_next_state,
ctrl/ctrl_sys/num_int_dtnl_system.hpp:166-180), for each fixed-step scheme of
core/integrators/fixed_step_integrators.hpp (euler, midpoint, runge_kutta4, runge_kutta5).
Build container only:

    make -C oracle ref && python tests/golden/make_golden_schemes.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyref  # noqa: E402
from reak_b200 import kte, presets  # noqa: E402

N, J, STEPS, DT = 8, 5, 4, 1e-3
CASES = ["planar2_act", "crs6", "crs6_sd", "crs7", "crs6_twist"]
SCHEMES = {"euler": 1, "midpoint": 2, "rk4": 4, "rk5": 5}


def main():
    if not pyref.have_ref():
        raise SystemExit("oracle/_ref/libreak_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    for idx, name in enumerate(CASES):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = pyref.Reference(c)
        rng = np.random.default_rng(2000 + idx)
        x = rng.uniform(-1.0, 1.0, (N, 2 * c.n_coords))
        u = rng.uniform(-2.0, 2.0, (N, J, c.n_inputs))
        out = dict(x=x, u_seq=u, dt=DT, steps_per_interval=STEPS, n_coords=c.n_coords, n_inputs=c.n_inputs)
        for sname, code in SCHEMES.items():
            xo, traj, st = R.rollout(x, u, code, DT, STEPS)
            assert not st.any()
            out["traj_" + sname] = traj
        np.savez(os.path.join(HERE, "schemes", name + ".npz"), **out)
        print("%-12s n=%d nu=%d" % (name, c.n_coords, c.n_inputs))


if __name__ == "__main__":
    main()
