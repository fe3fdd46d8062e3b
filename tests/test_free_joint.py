"""free_joint_3D (ctrl/mbd_kte/free_joints.cpp:119-208): chains whose state carries the 13 states of a free joint's
coordinate frame after the (q, qd) pairs (kte_nl_system.hpp:145-147, 205-219) and whose mass matrix carries its six
jacobian_3D_3D columns (core/kinetostatics/motion_jacobians.hpp:1077-1203, mass_matrix_calculator.cpp:232-276).

CPU: the oracle restatement against the committed golden vectors of the compiled reference (tests/golden/free/,
tests/golden/make_golden_free.py) and against the live reference where it is built; the host-side lowering.
GPU (-m gpu): the interpreter kernels through the C-ABI against the oracle and the golden vectors."""
import glob
import os

import numpy as np
import pytest

from conftest import rel_err
from reak_b200 import _abi, kte, presets

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(presets.FREE_PRESETS)
TOL_STEP, TOL_LONG = 1e-10, 1e-8


def _compiled(name):
    s = presets.make(name)
    return kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, s.dofs_3D)


def _golden(name):
    return np.load(os.path.join(HERE, "golden", "free", name + ".npz"))


def _batch(compiled, n_samples, seed):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_free", os.path.join(HERE, "golden", "make_golden_free.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.free_batch(compiled, n_samples, seed)


def test_fixtures_present():
    assert sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "free", "*.npz"))) == NAMES


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(name, oracle_built):
    c, g = _compiled(name), _golden(name)
    assert (c.n_coords, c.n_free, c.nx, c.n_acc) == (int(g["n_coords"]), int(g["n_free"]), g["x"].shape[1], g["M"].shape[1])
    O = oracle_built.Oracle(c)
    xd, st = O.eval(g["x"], g["u"])
    assert not st.any() and rel_err(xd, g["xdot"]) < 1e-13
    assert rel_err(O.gen_forces(g["x"], g["u"]), g["f"]) < 1e-13
    M, Md = O.mass(g["x"])
    assert rel_err(M, g["M"]) < 1e-13 and rel_err(Md, g["Mdot"]) < 1e-13
    T, Mc, Td = O.tmt(g["x"][0])
    assert rel_err(T, g["Tcm"]) < 1e-13 and rel_err(Td, g["Tcm_dot"]) < 1e-13 and rel_err(Mc, g["Mcm"]) < 1e-13
    assert rel_err(O.frames(g["x"][0], g["u"][0]), g["frames"]) < 1e-13
    assert rel_err(O.rk4(g["x"], g["u"], 1e-3, 1)[0], g["x1"]) < 1e-13
    assert rel_err(O.rk4(g["x"], g["u"], 1e-3, 25)[0], g["x25"]) < 1e-12
    for sch in (1, 2, 5):
        assert rel_err(O.integrate(g["x"], g["u"], sch, 1e-3, 10)[0], g["x10_scheme%d" % sch]) < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_live_reference(name, oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("compiled reference not built here")
    c = _compiled(name)
    O, R = oracle_built.Oracle(c), oracle_built.Reference(c)
    x, u = _batch(c, 48, 77)
    assert np.array_equal(O.eval(x, u)[0], R.eval(x, u)[0])
    assert np.array_equal(O.gen_forces(x, u), R.gen_forces(x, u))
    (M, Md), (Mr, Mdr) = O.mass(x), R.mass(x)
    assert np.array_equal(M, Mr) and np.array_equal(Md, Mdr)
    assert np.array_equal(O.rk4(x, u, 1e-3, 40)[0], R.rk4(x, u, 1e-3, 40)[0])


def test_state_derivative_structure(oracle_built):
    """kte_nl_system.hpp:293-308: the position part of xdot is the coordinate frame's velocity, the quaternion part is
    getQuaternionDot of the NORMALISED quaternion; scaling the quaternion of the state changes nothing."""
    c = _compiled("free_arm3")
    O = oracle_built.Oracle(c)
    x, u = _batch(c, 8, 5)
    xd = O.eval(x, u)[0]
    o = 2 * c.n_coords
    assert np.array_equal(xd[:, o:o + 3], x[:, o + 7:o + 10])
    q = x[:, o + 3:o + 7] / np.linalg.norm(x[:, o + 3:o + 7], axis=1, keepdims=True)
    w = x[:, o + 10:o + 13]
    qd = 0.5 * np.stack([-(q[:, 1] * w[:, 0] + q[:, 2] * w[:, 1] + q[:, 3] * w[:, 2]),
                         q[:, 0] * w[:, 0] - q[:, 3] * w[:, 1] + q[:, 2] * w[:, 2],
                         q[:, 0] * w[:, 1] + q[:, 3] * w[:, 0] - q[:, 1] * w[:, 2],
                         q[:, 0] * w[:, 2] - q[:, 2] * w[:, 0] + q[:, 1] * w[:, 1]], axis=1)
    assert rel_err(xd[:, o + 3:o + 7], qd) < 1e-15
    x2 = x.copy()
    x2[:, o + 3:o + 7] *= 2.0
    assert rel_err(O.eval(x2, u)[0], xd) < 1e-14


def test_free_body_mass_matrix_is_its_inertia(oracle_built):
    """one rigid body on a free joint: M = diag(m, m, m, I) whatever the state (the two identity blocks of the joint's
    Jacobian seen from the body frame itself)"""
    c = _compiled("free_body")
    O = oracle_built.Oracle(c)
    x, _ = _batch(c, 5, 3)
    M = O.mass(x, with_dot=False)
    want = np.zeros((6, 6))
    want[:3, :3] = 5.0 * np.eye(3)
    want[3:, 3:] = [[0.6, 0.02, -0.01], [0.02, 0.5, 0.03], [-0.01, 0.03, 0.4]]
    assert rel_err(M, np.broadcast_to(want, M.shape)) < 1e-14


def test_lowering_of_free_chains():
    c = _compiled("free_arm3")
    kinds = [e.kind for e in c.elements]
    assert kinds.count(_abi.FREE_3D) == 1 and (c.n_coords, c.n_free, c.nx, c.n_acc) == (3, 1, 19, 9)
    inertias = [e for e in c.elements if e.kind == _abi.INERTIA_3D]
    assert all((e.upstream >> 32) == 1 for e in inertias)             # every body rides on the free joint ...
    assert [e.upstream & 0xffff for e in inertias] == [0, 1, 3, 7]    # ... and on the arm joints below it
    # the reference's mass_matrix_calc dereferences a null Jacobian for a rotor on coordinate i < number of free joints
    # (mass_matrix_calculator.cpp:226-233): rejected, there is no behaviour to match
    s = presets.make("free_arm3")
    q0 = s.dofs_gen[0]
    rotor = kte.inertia_gen("rotor", kte.joint_dependent_gen_coord(q0).add_joint(q0, kte.jacobian_gen_gen()), 0.1)
    s.chain << rotor
    s.mass_calc << rotor
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, s.dofs_3D)
    # a free joint whose coordinate frame is not a system state
    s = presets.make("free_body")
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, [])


def test_library_accepts_free_chain_descriptors():
    """rkb_chain_create validates and lowers on the host (no GPU needed): dimensions of a free chain, rejections"""
    import ctypes as C
    lib = _abi.load_library()
    c = _compiled("free_arm3")
    h = C.c_void_p()
    assert lib.rkb_chain_create(C.byref(c.desc), C.byref(h)) == 0
    assert (lib.rkb_chain_state_dim(h), lib.rkb_chain_dof(h), lib.rkb_chain_input_dim(h), lib.rkb_chain_is_serial(h)) == (19, 3, 3, 0)
    lib.rkb_chain_destroy(h)
    # two free joints: more than the compiled path carries
    c2 = _compiled("free_body")
    arr = (_abi.rkb_element * 3)(c2.elements[0], c2.elements[0], c2.elements[1])
    arr[1].coord, arr[1].frame_a, arr[1].frame_b = 1, 1, 2
    d = _abi.rkb_chain_desc()
    C.memmove(C.byref(d), C.byref(c2.desc), C.sizeof(d))
    d.n_elements, d.n_frames, d.elements = 3, 3, C.cast(arr, C.POINTER(_abi.rkb_element))
    assert lib.rkb_chain_create(C.byref(d), C.byref(h)) == _abi.ERR_UNSUPPORTED


# ---------------------------------------------------------------------------------------------- GPU
def _prop(name):
    from reak_b200 import kte_batch_propagator
    return kte_batch_propagator(presets.make(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_free_chain_against_golden(name):
    g = _golden(name)
    p = _prop(name)
    assert not p.is_serial() and (p.nx, p.na) == (g["x"].shape[1], g["M"].shape[1])
    xd, st = p.get_state_derivatives(g["x"], g["u"])
    assert not st.any() and rel_err(xd, g["xdot"]) < TOL_STEP
    assert rel_err(p.get_gen_forces(g["x"], g["u"]), g["f"]) < TOL_STEP
    M, Md = p.get_mass_matrices(g["x"], with_derivative=True)
    assert rel_err(M, g["M"]) < TOL_STEP and rel_err(Md, g["Mdot"]) < TOL_STEP
    T, Mc, Td = p.get_twist_shaping(g["x"][:1])
    assert rel_err(T[0], g["Tcm"]) < TOL_STEP and rel_err(Td[0], g["Tcm_dot"]) < TOL_STEP and rel_err(Mc, g["Mcm"]) < TOL_STEP
    assert rel_err(p.get_frames(g["x"][:1], g["u"][:1])[0], g["frames"]) < TOL_STEP
    assert rel_err(p.get_next_states(g["x"], g["u"], 1e-3, 1)[0], g["x1"]) < TOL_STEP
    assert rel_err(p.get_next_states(g["x"], g["u"], 1e-3, 25)[0], g["x25"]) < TOL_LONG
    for sch, nm in ((1, "euler"), (2, "midpoint"), (5, "rk5")):
        u_seq = g["u"][:, None, :]
        xo, st = p.rollout(g["x"], u_seq, 1e-3, 10, scheme=nm)
        assert not st.any() and rel_err(xo, g["x10_scheme%d" % sch]) < TOL_LONG, nm


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_free_chain_against_oracle(name, oracle_built):
    p = _prop(name)
    O = oracle_built.Oracle(p.compiled)
    x, u = _batch(p.compiled, 301, 19)
    xd, st = p.get_state_derivatives(x, u)
    assert not st.any() and rel_err(xd, O.eval(x, u)[0]) < TOL_STEP
    M, Md = p.get_mass_matrices(x, with_derivative=True)
    Mo, Mdo = O.mass(x)
    assert rel_err(M, Mo) < TOL_STEP and rel_err(Md, Mdo) < TOL_STEP
    steps = 200 if p.n < 4 else 60   # (the six-joint arm on a free base under random unit torques leaves the finite range after ~150 ms)
    xo, st = p.get_next_states(x, u, 1e-3, steps)
    xr, sr, _ = O.rk4(x, u, 1e-3, steps, n_workers=min(8, os.cpu_count() or 1))
    assert not st.any() and not sr.any() and rel_err(xo, xr) < TOL_LONG
    # SoA buffers and device-resident tensors give the same bits as host AoS
    import torch
    xs, us = np.ascontiguousarray(x.T), np.ascontiguousarray(u.T)
    assert np.array_equal(p.get_next_states(xs, us, 1e-3, 3, soa=True)[0].T, p.get_next_states(x, u, 1e-3, 3)[0])
    xt, ut = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    assert np.array_equal(p.get_state_derivatives(xt, ut)[0].cpu().numpy(), xd)


@pytest.mark.gpu
def test_gpu_free_chain_frame_jacobian_is_a_tmt_row_block():
    """the Jacobian of an inertia's frame w.r.t. (coordinates, free joint) is that inertia's 6 rows of Tcm"""
    p = _prop("free_arm3")
    x, _ = _batch(p.compiled, 9, 4)
    T, _, Td = p.get_twist_shaping(x)
    inertias = [e for e in p.compiled.elements if e.kind == _abi.INERTIA_3D]
    for k, e in enumerate(inertias):
        up = [c for c in range(p.n) if (e.upstream >> c) & 1]
        J, Jd = p.get_frame_jacobian(x, int(e.frame_a), upstream=up, free_joints=[0])
        assert rel_err(J, T[:, 6 * k:6 * k + 6, :]) < 1e-13 and rel_err(Jd, Td[:, 6 * k:6 * k + 6, :]) < 1e-13


@pytest.mark.gpu
def test_gpu_free_chain_rejects_blocked_layout():
    from reak_b200 import kte_batch_propagator
    p = kte_batch_propagator(presets.make("free_arm3"), blocked=True)
    x, u = _batch(p.compiled, 4, 1)
    with pytest.raises(_abi.RkbError):
        p.get_state_derivatives(x, u)


# ---------------------------------------------------------------------------------------------- the reference-side binding
@pytest.mark.parametrize("name", NAMES)
def test_reak_bridge_round_trip_with_free_joint(name, oracle_built):
    """reak_bridge.hpp's compile_kte_system on the LIVE ReaK objects (free_joint_3D, its coordinate frame in
    kte_nl_system::dofs_3D and mass_matrix_calc::Frames3D, jacobian_3D_3D entries of mUpStream3DJoints): descriptor ->
    ReaK objects -> descriptor is the identity"""
    import ctypes as C
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built (needs /root/reference)")
    c = _compiled(name)
    R = oracle_built.Reference(c)
    fn = R.lib.rkref_bridge_desc
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    out = _abi.rkb_chain_desc()
    elems = (_abi.rkb_element * 256)()
    err = C.create_string_buffer(256)
    n = fn(R.h, C.byref(out), elems, 256, err, 256)
    assert n == c.desc.n_elements, err.value
    for field in ("dim", "n_elements", "n_frames", "n_coords", "n_inputs", "base_frame"):
        assert getattr(out, field) == getattr(c.desc, field), field
    for i in range(n):
        a, b = elems[i], c.elements[i]
        assert (a.kind, a.frame_a, a.frame_b, a.coord, a.aux, a.upstream) == (b.kind, b.frame_a, b.frame_b, b.coord, b.aux, b.upstream), i
        assert np.allclose(list(a.p), list(b.p), rtol=0, atol=1e-15), i


@pytest.mark.gpu
def test_gpu_reak_bridge_drop_in_with_free_joint(oracle_built):
    """ReaK::ctrl::kte_batch_system built from a live kte_nl_system WITH a free joint, compared in C++ against that
    kte_nl_system and ReaK's runge_kutta4_integrator (rkref_bridge_gpu_check, oracle/ref_lib.cpp)"""
    import ctypes as C
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    C.CDLL(_abi.LIB_PATH, mode=C.RTLD_GLOBAL)
    c = _compiled("free_arm3")
    R = oracle_built.Reference(c)
    fn = R.lib.rkref_bridge_gpu_check
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
    x, u = _batch(c, 64, 17)
    err = np.zeros(3)
    msg = C.create_string_buffer(512)
    rc = fn(R.h, 64, x.ctypes.data, u.ctypes.data, 1e-3, 20, err.ctypes.data, msg, 512)
    assert rc == 0, msg.value
    assert err[0] < TOL_STEP and err[1] < TOL_STEP and err[2] < TOL_LONG, err


@pytest.mark.gpu
def test_gpu_free_chain_steer_batch_and_linearisation(oracle_built):
    """the entry points built on top of the evaluation / rollout kernels carry the longer state of a free chain through:
    rkb_steer_batch (arg-min over constant-control rollouts) and rkb_linearize (A, B by central differences)"""
    p = _prop("free_arm3")
    O = oracle_built.Oracle(p.compiled)
    rng = np.random.default_rng(5)
    P, R, K = 5, 23, 8
    x0, _ = _batch(p.compiled, P, 41)
    goal, _ = _batch(p.compiled, P, 42)
    u = rng.uniform(-3, 3, (P, R, p.nu))
    idx, bx, bc, st = p.steer_batch(x0, goal, u, 1e-3, K, want_status=True)
    xe, _, _ = O.rk4(np.repeat(x0, R, axis=0), u.reshape(P * R, -1), 1e-3, K)
    cost = np.linalg.norm(xe.reshape(P, R, -1) - goal[:, None, :], axis=2)
    assert np.array_equal(idx, cost.argmin(axis=1)) and not st.any()
    assert rel_err(bx, xe.reshape(P, R, -1)[np.arange(P), idx]) < TOL_STEP and rel_err(bc, cost.min(axis=1)) < TOL_STEP
    # linearisation against the same central differences of the oracle
    n, nx, nu, eps = 11, p.nx, p.nu, 1e-6
    x, uu = _batch(p.compiled, n, 43)
    A, B, st = p.get_linear_blocks(x, uu, eps)
    assert not st.any() and A.shape == (n, nx, nx) and B.shape == (n, nx, nu)
    for d in range(nx + nu):
        xp, xm, up_, um = x.copy(), x.copy(), uu.copy(), uu.copy()
        if d < nx:
            h = eps * np.maximum(1.0, np.abs(x[:, d]))
            xp[:, d] += h; xm[:, d] -= h
            den = xp[:, d] - xm[:, d]
        else:
            h = eps * np.maximum(1.0, np.abs(uu[:, d - nx]))
            up_[:, d - nx] += h; um[:, d - nx] -= h
            den = up_[:, d - nx] - um[:, d - nx]
        col = (O.eval(xp, up_)[0] - O.eval(xm, um)[0]) / den[:, None]
        got = A[:, :, d] if d < nx else B[:, :, d - nx]
        assert rel_err(got, col) < 1e-7, d
    # control sequences with waypoints (one launch per interval on the interpreter)
    J = 4
    useq = rng.uniform(-2, 2, (n, J, nu))
    xo, traj, st = p.rollout(x, useq, 1e-3, 5, scheme="rk4", want_traj=True)
    xr, tr, sr = O.rollout(x, useq, 4, 1e-3, 5)
    assert not st.any() and rel_err(xo, xr) < TOL_LONG and rel_err(traj, tr) < TOL_LONG
