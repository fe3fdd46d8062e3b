"""Regenerates tests/golden/*.npz from the REAL reference (oracle/_ref/libreak_ref.so = the
unmodified ReaK sources under /root/reference compiled by oracle/Makefile).  Run in the build
container only (the GPU box has no /root/reference and never needs to run this):

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds the inputs (x, u), and the reference's outputs for them: xdot
(kte_nl_system::get_state_derivative), f (gen_coord::f after doMotion/clearForce/doForce), M and
Mdot (mass_matrix_calc::getMassMatrixAndDerivative), and the state after 1 and 25 RK4 steps of
1 ms (runge_kutta4_integrator<double>), plus the flat chain descriptor fields for provenance.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyref  # noqa: E402
from reak_b200 import kte, presets  # noqa: E402

N = 16
CASES = ["pendulum", "planar2", "planar3_sd", "torsion1", "crs3", "crs6", "crs6_phys", "crs6_sd", "crs6_sd_sat",
         "crs6_twist", "crs7", "crs7_phys_sd", "crs6_passive", "planar2_act", "crs6_lin_sd", "planar2_lin_sd", "planar_pr", "crs2d", "crs3_gen", "planar2_gen"]


def main():
    if not pyref.have_ref():
        raise SystemExit("oracle/_ref/libreak_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    only = set(sys.argv[1:])  # optional: regenerate just these fixtures (the seeds depend on the position in CASES only)
    for idx, name in enumerate(CASES):
        if only and name not in only:
            continue
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = pyref.Reference(c)
        rng = np.random.default_rng(1000 + idx)
        x = rng.uniform(-1.0, 1.0, (N, 2 * c.n_coords))
        x[:, 0::2] *= 2.5
        u = rng.uniform(-1.0, 1.0, (N, c.n_inputs))
        xdot, st = R.eval(x, u)
        assert not st.any()
        f = R.gen_forces(x, u)
        M, Md = R.mass(x)
        x1, s1, _ = R.rk4(x, u, 1e-3, 1)
        x25, s25, _ = R.rk4(x, u, 1e-3, 25)
        assert not s1.any() and not s25.any()
        kinds = np.array([e.kind for e in c.elements], dtype=np.int32)
        np.savez(os.path.join(HERE, name + ".npz"), x=x, u=u, xdot=xdot, f=f, M=M, Mdot=Md, x1=x1, x25=x25, kinds=kinds,
                 n_coords=c.n_coords, n_inputs=c.n_inputs)
        print("%-14s n=%d nu=%d elements=%d" % (name, c.n_coords, c.n_inputs, len(kinds)))


if __name__ == "__main__":
    main()
