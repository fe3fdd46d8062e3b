"""Generates tests/golden/steer/steer_loops.npz from the UNMODIFIED reference (oracle/_ref/libreak_ref.so, the unit
oracle/ref_steer_law.cpp): IHAQR_topology::move_position_toward_impl and MEAQR_topology::steer_with_constant_control
(examples/misc/IHAQR_topology.hpp:337-381, MEAQR_topology.hpp:503-561) run over the live kte_nl_system of two preset chains,
IHAQR_topology::get_bounded_input on random triples, and ctrl::detail::runge_kutta4_integrate_impl
(ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97) with an input trajectory.  Run from the repo root:

    python tests/golden/steer/make_golden_steer.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyref  # noqa: E402
from reak_b200 import kte, presets  # noqa: E402
from test_oracle import _steer_case  # noqa: E402

T = 0.05


def main():
    out = {}
    for name in ("crs6", "planar2_act"):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = pyref.Reference(c)
        nu = c.n_inputs
        x0, goal, u_bias, gain, u_prev = _steer_case(c, 10, seed=51, gain_scale=3.0)
        lo, hi, bw = -1.5 * np.ones(nu), 1.2 * np.ones(nu), 25.0 * np.ones(nu)
        u_prev = np.clip(3.0 * u_prev, lo, hi)
        for k, v in (("x0", x0), ("goal", goal), ("u_bias", u_bias), ("gain", gain), ("u_prev", u_prev), ("lo", lo), ("hi", hi), ("bw", bw)):
            out[name + "_" + k] = v
        # IHAQR: 5 intervals of T, steps of T / 100 (100 per interval for T = 0.05), threshold 0.05
        out[name + "_ihaqr_x"] = R.ihaqr_move_toward(x0, goal, u_bias, gain, u_prev, T, 4.5 * T, 0.05, (lo, hi), bw)
        # MEAQR: 8 intervals of T, steps of T / 10, first interval unsaturated
        xm, um, tm = R.meaqr_steer(x0, goal, u_bias, gain, u_prev, T, 7.5 * T, 0.05, (lo, hi), bw)
        out[name + "_meaqr_x"], out[name + "_meaqr_u"], out[name + "_meaqr_t"] = xm, um, tm
        print(name, "IHAQR moved", float(np.abs(out[name + "_ihaqr_x"] - x0).max()), "MEAQR intervals", np.rint(tm / T).astype(int).tolist())
    # get_bounded_input on random triples, three input counts
    ref = C.CDLL(pyref.REF_SO)
    fn = ref.rkref_ihaqr_bounded_input
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(52)
    for nu in (1, 3, 6):
        n = 300
        lo, hi, bw = -rng.uniform(0.5, 2.0, nu), rng.uniform(0.5, 2.0, nu), rng.uniform(0.5, 100.0, nu)
        up, ub = rng.uniform(-2.5, 2.5, (n, nu)), rng.uniform(-3.0, 3.0, (n, nu))
        uc = rng.uniform(-4.0, 4.0, (n, nu)) * rng.choice([0.05, 1.0, 10.0], (n, 1))
        res = np.zeros((n, nu))
        assert fn(nu, p(lo), p(hi), p(bw), 0.02, n, p(up), p(ub), p(uc), p(res)) == 0
        for k, v in (("lo", lo), ("hi", hi), ("bw", bw), ("u_prev", up), ("u_bias", ub), ("u_corr", uc), ("u_out", res)):
            out["law%d_%s" % (nu, k)] = v
    # ctrl::detail::runge_kutta4_integrate_impl itself (runge_kutta4_integrator_sys.hpp:50-97) with an input trajectory
    for name in ("crs6", "crs7_phys_sd"):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = pyref.Reference(c)
        rng = np.random.default_rng(53)
        K = 9
        x = np.zeros((8, 2 * c.n_coords))
        x[:, 0::2], x[:, 1::2] = rng.uniform(-1, 1, (8, c.n_coords)), rng.uniform(-1, 1, (8, c.n_coords))
        nodes = rng.uniform(-2, 2, (8, 2 * K + 1, c.n_inputs))
        xo, st = R.rk4_inputs_concept(x, nodes, 1e-3)
        assert not st.any()
        out["rk4c_%s_x0" % name], out["rk4c_%s_nodes" % name], out["rk4c_%s_xout" % name] = x, nodes, xo
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "steer", "steer_loops.npz"), **out)


if __name__ == "__main__":
    main()
