"""Regenerates tests/golden/free/*.npz — chains with a free_joint_3D (ctrl/mbd_kte/free_joints.cpp:119-208; 13 states and
6 accelerations per free joint, kte_nl_system.hpp:145-147, 205-219, 293-308) — from the REAL reference
(oracle/_ref/libreak_ref.so).  Build container only:

    make -C oracle ref && python tests/golden/make_golden_free.py

Per fixture: inputs (x, u — the quaternion part of x deliberately NOT of unit length: apply_states_and_inputs normalises
it, the integrators advance the raw vector) and the reference's xdot, f (coordinates, then Force / Torque of the joint's
coordinate frame), M, Mdot ((n + 6) x (n + 6)), Tcm / Tcm_dot / Mcm of the first sample, every frame of the first sample
after doMotion / doForce, and the states after 1 and 25 RK4 steps and 10 steps of euler / midpoint / runge_kutta5.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyref  # noqa: E402
from reak_b200 import kte, presets  # noqa: E402

N = 16


def free_batch(compiled, n_samples, seed, q_range=2.5):
    """states of a chain with free joints: (q, qd) pairs, then per free joint position, quaternion (norm 0.8 .. 1.25),
    velocity, angular velocity (free_joint_2D: position, (cos, sin) of norm 0.8 .. 1.25, velocity, angular velocity)"""
    rng = np.random.default_rng(seed)
    n, nu = compiled.n_coords, compiled.n_inputs
    x = rng.uniform(-1.0, 1.0, (n_samples, compiled.nx))
    x[:, 0:2 * n:2] *= q_range
    planar = compiled.desc.dim == 2   # free_joint_2D: 7 states, the rotation as (cos, sin) at 2..3
    for j in range(compiled.n_free):
        o = 2 * n + (7 if planar else 13) * j
        q = rng.normal(size=(n_samples, 2 if planar else 4))
        q *= (rng.uniform(0.8, 1.25, (n_samples, 1)) / np.linalg.norm(q, axis=1, keepdims=True))
        if planar:
            x[:, o + 2:o + 4] = q
        else:
            x[:, o + 3:o + 7] = q
    u = rng.uniform(-1.0, 1.0, (n_samples, nu))
    return x, u


def main():
    if not pyref.have_ref():
        raise SystemExit("oracle/_ref/libreak_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    only = set(sys.argv[1:])  # optional: regenerate just these fixtures
    order = ["free_arm2_twist", "free_arm3", "free_body", "free_planar2", "free_planar_body", "free_arm6"]  # seeds follow this order
    assert sorted(order) == sorted(presets.FREE_PRESETS)
    for idx, name in enumerate(order):
        if only and name not in only:
            continue
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, s.dofs_3D)
        R = pyref.Reference(c)
        x, u = free_batch(c, N, 2000 + idx)
        xdot, st = R.eval(x, u)
        assert not st.any()
        f = R.gen_forces(x, u)
        M, Md = R.mass(x)
        T, Mc, Td = R.tmt(x[0])
        frames = R.frames(x[0], u[0])
        x1, s1, _ = R.rk4(x, u, 1e-3, 1)
        x25, s25, _ = R.rk4(x, u, 1e-3, 25)
        assert not s1.any() and not s25.any()
        other = {("x10_scheme%d" % sch): R.integrate(x, u, sch, 1e-3, 10)[0] for sch in (1, 2, 5)}
        kinds = np.array([e.kind for e in c.elements], dtype=np.int32)
        np.savez(os.path.join(HERE, "free", name + ".npz"), x=x, u=u, xdot=xdot, f=f, M=M, Mdot=Md, Tcm=T, Mcm=Mc, Tcm_dot=Td,
                 frames=frames, x1=x1, x25=x25, kinds=kinds, n_coords=c.n_coords, n_inputs=c.n_inputs, n_free=c.n_free, **other)
        print("%-16s n=%d free=%d nu=%d nx=%d elements=%d" % (name, c.n_coords, c.n_free, c.n_inputs, c.nx, len(kinds)))


if __name__ == "__main__":
    main()
