"""GPU parity: libreak_b200.so (through the C-ABI / kte_batch_propagator) against the oracle.

Tolerances are BASELINE.json's: 1e-10 relative per step, 1e-8 after 1000 steps, relative error
taken per component as |a-b| / max(1, |b|).  Evaluation outputs (xdot, f, M, Mdot) are held to
1e-10 as well.  Every chain is run on the kernels the library picks for it and, where that is
the register-resident serial path, once more on the interpreter kernels (RKB_CREATE_INTERPRETER) and
with one thread per sample instead of one sample per pair of warps (RKB_OPT_SPLIT_MAX_SAMPLES = 0).
"""
import os

import numpy as np
import pytest

from conftest import random_batch, rel_err
from reak_b200 import kte, presets

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL_STEP = 1e-10
TOL_LONG = 1e-8
ALL = sorted(presets.PRESETS)


def _make(name, generic=False, split=None):
    """generic: run the chain on the interpreter kernels (rkb_chain_create_ex, RKB_CREATE_INTERPRETER);
    split: RKB_OPT_SPLIT_MAX_SAMPLES (0 = one thread per sample whatever the batch size)."""
    from reak_b200 import kte_batch_propagator
    p = kte_batch_propagator(presets.make(name), interpreter=generic)
    p.set_option("auto_specialize", 0)   # (kernels that change between two calls would disturb the bit-for-bit comparisons)
    if split is not None:
        p.set_option("split_max_samples", split)
    return p


def _variants(name):
    """Every kernel family that can run the chain: what the library picks by itself (small batches of a serial chain:
    one sample on a pair of warps), the one-thread-per-sample serial kernels, and the interpreter."""
    p = _make(name)
    out = [("auto", p)]
    if p.is_serial():
        out.append(("thread-per-sample", _make(name, split=0)))
        out.append(("generic", _make(name, generic=True)))
    return out


@pytest.mark.parametrize("name", ALL)
def test_eval_forces_mass(name, oracle_built):
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        x, u = random_batch(p.compiled, 257, seed=11, q_range=3.0)
        xd, st = p.get_state_derivatives(x, u)
        xd_o, st_o = O.eval(x, u)
        assert not st.any() and not st_o.any()
        assert rel_err(xd, xd_o) < TOL_STEP, (name, label)
        assert rel_err(p.get_gen_forces(x, u), O.gen_forces(x, u)) < TOL_STEP, (name, label)
        M, Md = p.get_mass_matrices(x, with_derivative=True)
        M_o, Md_o = O.mass(x)
        assert rel_err(M, M_o) < TOL_STEP and rel_err(Md, Md_o) < TOL_STEP, (name, label)
        assert rel_err(p.get_mass_matrices(x), M_o) < TOL_STEP, (name, label)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ALL)
def test_against_golden_reference_vectors(name):
    """tests/golden/*.npz hold outputs of the unmodified reference (see tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.isfile(path):
        pytest.skip("no fixture for " + name)
    g = np.load(path)
    for label, p in _variants(name):
        xd, st = p.get_state_derivatives(g["x"], g["u"])
        assert not st.any() and rel_err(xd, g["xdot"]) < TOL_STEP, (name, label)
        assert rel_err(p.get_gen_forces(g["x"], g["u"]), g["f"]) < TOL_STEP, (name, label)
        M, Md = p.get_mass_matrices(g["x"], with_derivative=True)
        assert rel_err(M, g["M"]) < TOL_STEP and rel_err(Md, g["Mdot"]) < TOL_STEP, (name, label)
        assert rel_err(p.get_next_states(g["x"], g["u"], 1e-3, 1)[0], g["x1"]) < TOL_STEP, (name, label)
        assert rel_err(p.get_next_states(g["x"], g["u"], 1e-3, 25)[0], g["x25"]) < TOL_LONG, (name, label)


@pytest.mark.parametrize("name", ALL)
def test_rk4_one_and_many_steps(name, oracle_built):
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        x, u = random_batch(p.compiled, 96, seed=5)
        for steps, tol in ((1, TOL_STEP), (100, TOL_LONG)):
            xo, st = p.get_next_states(x, u, 1e-3, steps)
            xr, sr, _ = O.rk4(x, u, 1e-3, steps)
            assert not st.any() and not sr.any()
            assert rel_err(xo, xr) < tol, (name, label, steps)


def test_cfg1_planar_1024x1000(oracle_built):
    """BASELINE config 1: 2-link planar arm, 1024 random states, 1000 RK4 steps of 1 ms."""
    p = _make("planar2")
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 1024, seed=12345, q_range=np.pi, qd_range=2.0)
    xo, st = p.get_next_states(x, None, 1e-3, 1000)
    xr, sr, _ = O.rk4(x, None, 1e-3, 1000, n_workers=os.cpu_count() or 1)
    assert not st.any() and not sr.any()
    assert rel_err(xo, xr) < TOL_LONG


def test_cfg2_crs6_1000_steps(oracle_built):
    p = _make("crs6")
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 64, seed=12346)
    xo, st = p.get_next_states(x, u, 1e-3, 1000)
    xr, sr, _ = O.rk4(x, u, 1e-3, 1000, n_workers=min(8, os.cpu_count() or 1))
    assert not st.any() and not sr.any()
    assert rel_err(xo, xr) < TOL_LONG


def test_layouts_and_device_buffers(oracle_built):
    import torch
    p = _make("crs6_sd")
    x, u = random_batch(p.compiled, 1000, seed=3)
    ref, _ = p.get_next_states(x, u, 1e-3, 7)
    # SoA host
    xs, st = p.get_next_states(np.ascontiguousarray(x.T), np.ascontiguousarray(u.T), 1e-3, 7, soa=True)
    assert np.array_equal(xs.T, ref)
    # device AoS / SoA (torch tensors are passed zero-copy)
    xt, ut = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    xo, st = p.get_next_states(xt, ut, 1e-3, 7)
    torch.cuda.synchronize()
    assert np.array_equal(xo.cpu().numpy(), ref) and not st.any().item()
    xo, st = p.get_next_states(xt.t().contiguous(), ut.t().contiguous(), 1e-3, 7, soa=True)
    assert np.array_equal(xo.t().cpu().numpy(), ref)
    # zero steps is the identity, n = 0 is a no-op
    same, _ = p.get_next_states(x, u, 1e-3, 0)
    assert np.array_equal(same, x)
    e, _ = p.get_next_states(x[:0], u[:0], 1e-3, 3)
    assert e.shape == (0, p.nx)
    assert p.launch_count() > 0 and p.last_kernel_ms() >= 0.0


def test_single_sample_concept_api(oracle_built):
    from reak_b200.propagator import singularity_error  # noqa: F401
    p = _make("crs6")
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 1, seed=9)
    xd = p.get_state_derivative(None, x[0], u[0], 0.0)
    assert rel_err(xd, O.eval(x, u)[0][0]) < TOL_STEP
    p.set_time_step(2e-3)
    xn = p.get_next_state(None, x[0], u[0], 0.0)
    assert rel_err(xn, O.rk4(x, u, 2e-3, 1)[0][0]) < TOL_STEP
    with pytest.raises(IndexError):
        p.get_state_derivative(None, x[0][:-1], u[0])
    with pytest.raises(IndexError):
        p.get_state_derivative(None, x[0], u[0][:-1])


def test_singular_mass_matrix_sets_status():
    """A massless chain has a singular M: the reference throws singularity_error
    (mat_cholesky.hpp:80-82); the batch path reports it per sample."""
    s = presets.crs_chain(n_revolute=2)
    for k in s.chain.getKTEs():
        if isinstance(k, kte.inertia_gen):
            k.mMass = 0.0
        if isinstance(k, kte.inertia_3D):
            k.mMass, k.mInertiaTensor = 0.0, [0.0] * 6
    from reak_b200 import kte_batch_propagator, _abi
    p = kte_batch_propagator(s)
    x, u = random_batch(p.compiled, 5, seed=1)
    _, st = p.get_state_derivatives(x, u)
    assert (st & _abi.STATUS_SINGULAR).all()
    from reak_b200.propagator import singularity_error
    with pytest.raises(singularity_error):
        p.get_state_derivative(None, x[0], u[0])


def test_steer_batch(oracle_built):
    p = _make("crs6")
    O = oracle_built.Oracle(p.compiled)
    rng = np.random.default_rng(77)
    P, R, K = 6, 37, 10
    x0, _ = random_batch(p.compiled, P, seed=21)
    goal, _ = random_batch(p.compiled, P, seed=22)
    u = rng.uniform(-5, 5, (P, R, p.nu))
    idx, bx, bc, st = p.steer_batch(x0, goal, u, 1e-3, K, want_status=True)
    xe, _, _ = O.rk4(np.repeat(x0, R, axis=0), u.reshape(P * R, -1), 1e-3, K)
    cost = np.linalg.norm(xe.reshape(P, R, -1) - goal[:, None, :], axis=2)
    assert np.array_equal(idx, cost.argmin(axis=1))
    assert rel_err(bx, xe.reshape(P, R, -1)[np.arange(P), idx]) < TOL_STEP
    assert rel_err(bc, cost.min(axis=1)) < TOL_STEP
    assert st.shape == (P, R) and not st.any()


def test_steer_batch_diverged_rollouts_never_win():
    """A rollout whose end state is NaN / inf has no distance to the goal: it must lose against any finite one, also
    when it sits at index 0 (where a plain `<` comparison would keep it), and a pair whose rollouts all diverged
    reports index 0 with an infinite cost."""
    p = _make("crs6")
    rng = np.random.default_rng(78)
    P, R, K = 5, 300, 5          # R > 256: several rollouts per thread of the arg-min kernel
    x0, _ = random_batch(p.compiled, P, seed=24)
    goal, _ = random_batch(p.compiled, P, seed=25)
    u = rng.uniform(-5, 5, (P, R, p.nu))
    clean = p.steer_batch(x0, goal, u, 1e-3, K)
    bad = u.copy()
    bad[0, 0, :] = np.nan                      # pair 0: rollout 0 diverges
    bad[1, int(clean[0][1]), 2] = np.inf       # pair 1: the former winner diverges
    bad[2, :, 0] = np.nan                      # pair 2: every rollout diverges
    bad[3, 0:256, 1] = np.nan                  # pair 3: the first rollout of every thread diverges
    idx, bx, bc, st = p.steer_batch(x0, goal, bad, 1e-3, K, want_status=True)
    from reak_b200 import _abi
    assert st[0, 0] & _abi.STATUS_NONFINITE and st[2].all()
    assert idx[0] != 0 and np.isfinite(bc[0]) and np.isfinite(bx[0]).all()
    assert idx[1] != clean[0][1] and np.isfinite(bc[1]) and bc[1] >= clean[2][1]
    assert idx[2] == 0 and np.isinf(bc[2])
    assert idx[3] >= 256 and np.isfinite(bc[3])
    assert idx[4] == clean[0][4] and bc[4] == clean[2][4]
    if idx[0] == clean[0][0]:
        assert bc[0] == clean[2][0]


def test_steer_feedback_zero_intervals_writes_nothing_into_x_traj():
    """max_intervals = 0 is allowed; x_traj is then a zero-length buffer and the library must not touch it
    (plain C-ABI call with a canary right behind the pointer; host and device buffers)."""
    import ctypes as C
    import torch
    from reak_b200 import _abi
    p = _make("crs6")
    lib = _abi.load_library()
    N = 300
    x0, goal, u_bias, gain, u_prev = _steer_case(p, N, seed=44)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for checked in (False, True):
        canary = np.full(N * p.nx + 64, 1234.5)
        out, nd, col = np.empty_like(x0), np.full(N, -1, dtype=np.int32), np.zeros(N, dtype=np.int32)
        up = u_prev.copy()
        o = _abi.rkb_steer_opts(1e-2, 1e-3, 0.25, 10, 0, 0, 0, None, None, None, None)
        if checked:
            from reak_b200 import kte_batch_propagator, proximity
            system = presets.make("crs6")
            robot, lab = presets.crs_proxy_models(system)
            pp = kte_batch_propagator(system)
            pair = proximity.proxy_query_pair_3D("robot-lab", robot, lab)
            h = proximity.ProxyHandle(lib, pp._h, pair, pp.compiled.frames)
            arr = (C.c_void_p * 1)(h._h)
            rc = lib.rkb_steer_feedback_checked(pp._h, 0, N, vp(x0), vp(goal), vp(u_bias), vp(gain), vp(up), C.byref(o), arr, 1,
                                                vp(out), vp(nd), vp(col), vp(canary), None, 0, None)
        else:
            rc = lib.rkb_steer_feedback(p._h, 0, N, vp(x0), vp(goal), vp(u_bias), vp(gain), vp(up), C.byref(o), vp(out), vp(nd),
                                        vp(canary), None, 0, None)
        assert rc == 0
        assert np.all(canary == 1234.5), "x_traj was written with max_intervals == 0"
        assert np.array_equal(out, x0) and not nd.any() and np.array_equal(up, u_prev)
    # device buffers
    t = lambda a: torch.from_numpy(a).cuda()
    canary = torch.full((N * p.nx + 64,), 1234.5, dtype=torch.float64, device="cuda")
    dx0, dg, db, dgn, dup = t(x0), t(goal), t(u_bias), t(gain), t(u_prev)
    dout, dnd = torch.empty_like(dx0), torch.full((N,), -1, dtype=torch.int32, device="cuda")
    o = _abi.rkb_steer_opts(1e-2, 1e-3, 0.25, 10, 0, 0, 0, None, None, None, None)
    dp = lambda a: C.c_void_p(a.data_ptr())
    rc = lib.rkb_steer_feedback(p._h, 0, N, dp(dx0), dp(dg), dp(db), dp(dgn), dp(dup), C.byref(o), dp(dout), dp(dnd), dp(canary), None,
                                _abi.MEM_DEVICE, None)
    torch.cuda.synchronize()
    assert rc == 0 and bool((canary == 1234.5).all()) and torch.equal(dout, dx0) and not bool(dnd.any())


def test_full_size_properties():
    """At BASELINE config 2's size (2^20 states) the oracle cannot follow; check size-independent
    properties instead: a strided sub-batch reproduces the big batch bit for bit, two half-length
    rollouts compose to the full one bit for bit, and integrating back with -dt returns to the start."""
    p = _make("crs6")
    N = 1 << 20
    x, u = random_batch(p.compiled, N, seed=12346)
    full, st = p.get_next_states(x, u, 1e-3, 20)
    assert not st.any() and np.isfinite(full).all()
    sel = np.arange(0, N, 4099)
    sub, _ = p.get_next_states(x[sel], u[sel], 1e-3, 20)
    assert np.array_equal(sub, full[sel])
    half, _ = p.get_next_states(x, u, 1e-3, 10)
    two, _ = p.get_next_states(half, u, 1e-3, 10)
    assert np.array_equal(two, full)
    back, _ = p.get_next_states(full, u, -1e-3, 20)
    assert rel_err(back, x) < 1e-9


def test_reak_bridge_cpp_drop_in(oracle_built):
    """ReaK::ctrl::kte_batch_system (include/reak_b200/reak_bridge.hpp) built from a LIVE kte_nl_system,
    compared in C++ against that kte_nl_system and ReaK's own runge_kutta4_integrator — the code is
    compiled against the unmodified reference headers inside oracle/_ref/libreak_ref.so."""
    import ctypes as C
    from reak_b200 import _abi
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    C.CDLL(_abi.LIB_PATH, mode=C.RTLD_GLOBAL)
    for name in ("crs6", "crs7_phys_sd", "planar2_lin_sd"):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = oracle_built.Reference(c)
        fn = R.lib.rkref_bridge_gpu_check
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        x, u = random_batch(c, 64, seed=17)
        err = np.zeros(3)
        msg = C.create_string_buffer(512)
        rc = fn(R.h, 64, x.ctypes.data, u.ctypes.data, 1e-3, 20, err.ctypes.data, msg, 512)
        assert rc == 0, msg.value
        assert err[0] < TOL_STEP and err[1] < TOL_STEP and err[2] < TOL_LONG, (name, err)


def test_multi_device_host_rollout():
    """rkb_rollout_rk4_multi: block partition over every visible GPU from one process; bit-identical
    to the single-device result (runs with one device when only one is visible)."""
    import torch
    p = _make("crs6")
    n_dev = torch.cuda.device_count()
    x, u = random_batch(p.compiled, 200003, seed=41)
    ref, st0 = p.get_next_states(x, u, 1e-3, 6)
    for devs in ([0], list(range(n_dev)), [0] * 3):
        got, st = p.get_next_states_multi(x, u, 1e-3, 6, devices=devs)
        assert np.array_equal(got, ref) and np.array_equal(st, st0), devs
    same, st = p.get_next_states_multi(x[:1000], u[:1000], 1e-3, 0, devices=[0])
    assert np.array_equal(same, x[:1000]) and not st.any()


def test_ragged_sizes_and_permuted_dofs(oracle_built):
    """Batch sizes around the 128-thread tile and the 65536-sample pipeline threshold; a chain whose
    state order is not the chain order."""
    p = _make("crs6_sd")
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 66000, seed=19)
    full, _ = p.get_next_states(x, u, 1e-3, 3)
    for n in (1, 2, 127, 128, 129, 65535, 65536, 65537):
        got, st = p.get_next_states(x[:n], u[:n], 1e-3, 3)
        assert np.array_equal(got, full[:n]) and not st.any(), n
    assert rel_err(full[:64], O.rk4(x[:64], u[:64], 1e-3, 3)[0]) < TOL_STEP
    from reak_b200 import kte_batch_propagator
    s = presets.make("crs6_phys")
    perm = [3, 0, 5, 1, 4, 2]
    s.dofs_gen = [s.dofs_gen[i] for i in perm]
    s.mass_calc.mCoords = list(s.dofs_gen)
    s.inputs = [s.inputs[i] for i in (5, 4, 3, 2, 1, 0)]
    q = kte_batch_propagator(s)
    assert q.is_serial()
    Oq = oracle_built.Oracle(q.compiled)
    x, u = random_batch(q.compiled, 77, seed=23)
    assert rel_err(q.get_state_derivatives(x, u)[0], Oq.eval(x, u)[0]) < TOL_STEP
    assert rel_err(q.get_mass_matrices(x), Oq.mass(x, with_dot=False)) < TOL_STEP
    assert rel_err(q.get_gen_forces(x, u), Oq.gen_forces(x, u)) < TOL_STEP
    assert rel_err(q.get_next_states(x, u, 1e-3, 20)[0], Oq.rk4(x, u, 1e-3, 20)[0]) < TOL_LONG


# ---- rkb_rollout: every fixed-step scheme, control sequences, interval trajectories ---------------
SCHEME_CODES = {"euler": 1, "midpoint": 2, "rk4": 4, "rk5": 5}
SCHEME_DIR = os.path.join(GOLDEN_DIR, "schemes")


@pytest.mark.parametrize("name", ["planar2_act", "crs6", "crs6_sd", "crs7", "crs6_twist"])
def test_rollout_schemes_against_golden_reference_trajectories(name):
    """tests/golden/schemes/*.npz: trajectories of the unmodified reference driven one control interval
    at a time (tests/golden/make_golden_schemes.py)."""
    g = np.load(os.path.join(SCHEME_DIR, name + ".npz"))
    for label, p in _variants(name):
        for sname in SCHEME_CODES:
            xo, traj, st = p.rollout(g["x"], g["u_seq"], float(g["dt"]), int(g["steps_per_interval"]), scheme=sname, want_traj=True)
            assert not st.any()
            assert rel_err(traj, g["traj_" + sname]) < TOL_STEP, (name, label, sname)
            assert np.array_equal(xo, traj[:, -1, :])


@pytest.mark.parametrize("name", ["planar2", "planar_pr", "crs3", "crs6_phys", "crs6_sd_sat", "crs7_phys_sd", "crs6_lin_sd"])
def test_rollout_schemes_vs_oracle(name, oracle_built):
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        x, _ = random_batch(p.compiled, 130, seed=21)
        rng = np.random.default_rng(22)
        u_seq = rng.uniform(-1.5, 1.5, (130, 3, p.nu))
        for sname, code in SCHEME_CODES.items():
            xo, traj, st = p.rollout(x, u_seq, 2e-3, 6, scheme=sname, want_traj=True)
            xr, tr, sr = O.rollout(x, u_seq, code, 2e-3, 6)
            assert not st.any() and not sr.any()
            assert rel_err(traj, tr) < TOL_STEP and rel_err(xo, xr) < TOL_STEP, (name, label, sname)
            # without the trajectory the end state is the same, bit for bit
            xo2, st2 = p.rollout(x, u_seq, 2e-3, 6, scheme=sname)
            assert np.array_equal(xo2, xo)


def test_rollout_rk4_single_interval_is_get_next_states():
    import torch
    p = _make("crs6")
    x, u = random_batch(p.compiled, 3000, seed=4)
    ref, _ = p.get_next_states(x, u, 1e-3, 9)
    xo, st = p.rollout(x, u[:, None, :], 1e-3, 9, scheme="rk4")
    assert np.array_equal(xo, ref) and not st.any()
    # a constant sequence equals one long interval; device buffers give the same bits as host buffers
    u_seq = np.ascontiguousarray(np.repeat(u[:, None, :], 3, axis=1))
    xo3, tr3, _ = p.rollout(x, u_seq, 1e-3, 3, scheme="rk4", want_traj=True)
    assert np.array_equal(xo3, ref)
    xt, ut = torch.from_numpy(x).cuda(), torch.from_numpy(u_seq).cuda()
    xd, trd, _ = p.rollout(xt, ut, 1e-3, 3, scheme="rk4", want_traj=True)
    torch.cuda.synchronize()
    assert np.array_equal(xd.cpu().numpy(), ref) and np.array_equal(trd.cpu().numpy(), tr3)


def test_rollout_large_host_batch_is_pipelined_and_consistent():
    """Above 2^16 samples host buffers go through the chunked copy/compute pipeline, trajectory included."""
    p = _make("crs6")
    n = (1 << 16) + 777
    x, _ = random_batch(p.compiled, n, seed=6)
    u_seq = np.random.default_rng(7).uniform(-1, 1, (n, 2, p.nu))
    xo, tr, st = p.rollout(x, u_seq, 1e-3, 2, scheme="rk4", want_traj=True)
    m = 500
    xs, ts, _ = p.rollout(x[-m:], u_seq[-m:], 1e-3, 2, scheme="rk4", want_traj=True)
    assert not st.any() and np.array_equal(xo[-m:], xs) and np.array_equal(tr[-m:], ts)
    xe, te, _ = p.rollout(x, u_seq, 1e-3, 2, scheme="midpoint", want_traj=True)
    xs, ts, _ = p.rollout(x[:m], u_seq[:m], 1e-3, 2, scheme="midpoint", want_traj=True)
    assert np.array_equal(xe[:m], xs) and np.array_equal(te[:m], ts)


def test_rollout_soa_through_the_c_abi():
    """SOA layout of rkb_rollout: u [n_intervals][nu][N], x_traj [n_intervals][2n][N]."""
    import ctypes as C
    from reak_b200 import _abi
    p = _make("crs6_sd")
    n, J = 257, 4
    x, _ = random_batch(p.compiled, n, seed=8)
    u_seq = np.random.default_rng(9).uniform(-1, 1, (n, J, p.nu))
    xo, tr, _ = p.rollout(x, u_seq, 1e-3, 3, scheme="rk5", want_traj=True)
    xs = np.ascontiguousarray(x.T)
    us = np.ascontiguousarray(u_seq.transpose(1, 2, 0))
    xo_s, tr_s, st = np.empty_like(xs), np.empty((J, p.nx, n)), np.zeros(n, dtype=np.int32)
    opts = _abi.rkb_rollout_opts(_abi.SCHEME_RK5, J, 3, 0, 1e-3)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _abi.check(_abi.load_library().rkb_rollout(p._h, 0, n, vp(xs), vp(us), C.byref(opts), vp(xo_s), vp(tr_s), vp(st),
                                               _abi.LAYOUT_SOA, None))
    assert np.array_equal(xo_s.T, xo) and np.array_equal(tr_s.transpose(2, 0, 1), tr) and not st.any()


def test_rollout_argument_errors():
    import ctypes as C
    from reak_b200 import _abi
    p = _make("crs3")
    x, u = random_batch(p.compiled, 4, seed=1)
    lib = _abi.load_library()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.empty_like(x)
    call = lambda o: lib.rkb_rollout(p._h, 0, 4, vp(x), vp(u), C.byref(o) if o is not None else None, vp(out), None, None, 0, None)
    assert call(None) == _abi.ERR_INVALID
    assert call(_abi.rkb_rollout_opts(3, 1, 1, 0, 1e-3)) == _abi.ERR_INVALID        # no such scheme
    assert call(_abi.rkb_rollout_opts(4, 1, 1, 7, 1e-3)) == _abi.ERR_INVALID        # reserved != 0
    assert call(_abi.rkb_rollout_opts(4, 0, 1, 0, 1e-3)) == _abi.ERR_INTEGRATION    # no interval
    assert call(_abi.rkb_rollout_opts(4, 1, -1, 0, 1e-3)) == _abi.ERR_INTEGRATION
    assert call(_abi.rkb_rollout_opts(4, 1, 1, 0, 0.0)) == _abi.ERR_INTEGRATION     # impossible_integration
    assert call(_abi.rkb_rollout_opts(5, 1, 2, 0, -1e-3)) == 0                      # backwards is allowed


# ---- rkb_steer_feedback: the closed-loop steering loop ------------------------------------------------
def _steer_case(p, n, seed, gain_scale=4.0):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-0.5, 0.5, (n, p.nx))
    goal = x0 + rng.uniform(-0.3, 0.3, (n, p.nx))
    u_bias = rng.uniform(-1.0, 1.0, (n, p.nu))
    gain = rng.uniform(-gain_scale, gain_scale, (n, p.nu, p.nx))
    u_prev = rng.uniform(-0.5, 0.5, (n, p.nu))
    return x0, goal, u_bias, gain, u_prev


@pytest.mark.parametrize("name", ["crs6", "crs6_sd", "crs7", "planar2_act"])
@pytest.mark.parametrize("saturate_first", [False, True])
def test_steer_feedback_vs_oracle(name, saturate_first, oracle_built):
    """Loop of MEAQR_topology.hpp:503-561 (saturate_first = False) and IHAQR_topology.hpp:349-378 (True)."""
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        n = 300
        x0, goal, u_bias, gain, u_prev = _steer_case(p, n, seed=41)
        goal[:7] = x0[:7]  # already at the goal: the loop must not start for these
        kw = dict(bounds=(-2 * np.ones(p.nu), 2 * np.ones(p.nu)), rate_bounds=(-60 * np.ones(p.nu), 60 * np.ones(p.nu)))
        args = (x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, 6, 0.25)
        xo, ul, nd, tr, st = p.steer_feedback(*args, saturate_first=saturate_first, want_traj=True, **kw)
        xr, ur, nr, trr, sr = O.steer_feedback(*args, saturate_first=saturate_first, **kw)
        assert not st.any() and not sr.any()
        assert np.array_equal(nd, nr), (name, label)
        assert nd[:7].max() == 0 and 0 < nd.max() <= 6 and len(set(nd.tolist())) > 1  # mixed stopping times
        assert rel_err(xo, xr) < TOL_STEP and rel_err(ul, ur) < TOL_STEP, (name, label)
        for i in range(n):
            assert rel_err(tr[i, :nd[i]], trr[i, :nr[i]]) < TOL_STEP, (name, label, i)
        assert np.array_equal(xo[:7], x0[:7]) and np.array_equal(ul[:7], u_prev[:7])
        # without boxes, without the trajectory
        xo2, ul2, nd2, st2 = p.steer_feedback(*args, saturate_first=saturate_first)
        xr2, ur2, nr2, _, _ = O.steer_feedback(*args, saturate_first=saturate_first)
        assert np.array_equal(nd2, nr2) and rel_err(xo2, xr2) < TOL_STEP and rel_err(ul2, ur2) < TOL_STEP


def test_steer_feedback_device_buffers_and_limits():
    import ctypes as C
    import torch
    from reak_b200 import _abi
    p = _make("crs6")
    x0, goal, u_bias, gain, u_prev = _steer_case(p, 2000, seed=43)
    args = (1e-2, 1e-3, 10, 4, 0.25)
    ref = p.steer_feedback(x0, goal, u_bias, gain, u_prev, *args, want_traj=True)
    t = lambda a: torch.from_numpy(a).cuda()
    dev = p.steer_feedback(t(x0), t(goal), t(u_bias), t(gain), t(u_prev), *args, want_traj=True)
    torch.cuda.synchronize()
    nd = ref[2]
    assert np.array_equal(dev[2].cpu().numpy(), nd)
    assert np.array_equal(dev[0].cpu().numpy(), ref[0]) and np.array_equal(dev[1].cpu().numpy(), ref[1])
    trd = dev[3].cpu().numpy()
    for i in range(0, 2000, 37):
        assert np.array_equal(trd[i, :nd[i]], ref[3][i, :nd[i]])
    # a zero time limit leaves everything as it was
    xo, ul, n0, st = p.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, 0, 0.25)
    assert np.array_equal(xo, x0) and np.array_equal(ul, u_prev) and not n0.any()
    # argument errors
    lib = _abi.load_library()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    out, n_done = np.empty_like(x0), np.zeros(2000, dtype=np.int32)
    up = u_prev.copy()
    bad_lo, bad_hi = np.ones(p.nu), -np.ones(p.nu)

    def call(o, flags=0):
        return lib.rkb_steer_feedback(p._h, 0, 2000, vp(x0), vp(goal), vp(u_bias), vp(gain), vp(up), C.byref(o), vp(out), vp(n_done),
                                      None, None, flags, None)

    assert call(_abi.rkb_steer_opts(1e-2, 0.0, 0.1, 10, 4, 0, 0, None, None, None, None)) == _abi.ERR_INTEGRATION
    assert call(_abi.rkb_steer_opts(1e-2, 1e-3, 0.1, 0, 4, 0, 0, None, None, None, None)) == _abi.ERR_INTEGRATION
    assert call(_abi.rkb_steer_opts(0.0, 1e-3, 0.1, 10, 4, 0, 0, None, None, None, None)) == _abi.ERR_INVALID
    assert call(_abi.rkb_steer_opts(1e-2, 1e-3, 0.1, 10, 4, 0, 0, vp(bad_lo), vp(bad_hi), None, None)) == _abi.ERR_INVALID
    assert call(_abi.rkb_steer_opts(1e-2, 1e-3, 0.1, 10, 4, 0, 0, None, None, None, None), _abi.LAYOUT_SOA) == _abi.ERR_UNSUPPORTED


def test_plain_c_client(tmp_path):
    """examples/c_abi_demo.c: C99 against the C-ABI, closed-form pendulum answers, no Python in the loop."""
    import subprocess
    from test_abi import _build_c_demo
    exe, env = _build_c_demo(tmp_path)
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith(")") and "ok (library version" in r.stdout, (r.stdout, r.stderr)


# ---- maximum sizes -----------------------------------------------------------------------------------
@pytest.mark.parametrize("n_rev,track,springs,serial", [(8, False, True, True), (7, True, True, True), (12, False, True, False),
                                                        (16, False, False, False)])
def test_long_chains(n_rev, track, springs, serial, oracle_built):
    """8 coordinates: the largest register-resident instance; 9..16 (RKB_MAX_COORDS): interpreter kernels
    (at most 96 elements, so 16 stages without the spring / damper pairs)."""
    from reak_b200 import kte_batch_propagator
    s = presets.crs_chain(n_revolute=n_rev, track=track, physical=True, springs=springs)
    p = kte_batch_propagator(s)
    assert p.is_serial() == serial
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 70, seed=15)
    xd, st = p.get_state_derivatives(x, u)
    assert not st.any() and rel_err(xd, O.eval(x, u)[0]) < TOL_STEP
    M, Md = p.get_mass_matrices(x, with_derivative=True)
    Mo, Mdo = O.mass(x)
    assert rel_err(M, Mo) < TOL_STEP and rel_err(Md, Mdo) < TOL_STEP
    assert rel_err(p.get_gen_forces(x, u), O.gen_forces(x, u)) < TOL_STEP
    xo, st = p.get_next_states(x, u, 1e-3, 20)
    assert not st.any() and rel_err(xo, O.rk4(x, u, 1e-3, 20)[0]) < TOL_LONG
    for scheme, code in (("euler", 1), ("rk5", 5)):
        xs, st = p.rollout(x, u[:, None, :], 1e-3, 4, scheme=scheme)
        assert rel_err(xs, O.integrate(x, u, code, 1e-3, 4)[0]) < TOL_STEP


def test_chains_beyond_the_limits_are_rejected():
    from reak_b200 import _abi, kte_batch_propagator
    with pytest.raises(kte.UnsupportedChain):
        kte_batch_propagator(presets.crs_chain(n_revolute=17))           # RKB_MAX_COORDS = 16
    with pytest.raises(_abi.RkbError) as e:
        kte_batch_propagator(presets.crs_chain(n_revolute=16, springs=True))  # 112 elements > 96
    assert e.value.code == _abi.ERR_UNSUPPORTED


# ---- RKB_LAYOUT_BLOCKED: the state order of manipulator_dynamics_model::computeStateRate -----------------
def _to_blocked(x):
    """[..., 2n] interleaved (q0, qd0, q1, qd1, ...) -> (q0 .. qn-1, qd0 .. qdn-1)."""
    return np.ascontiguousarray(np.concatenate([x[..., 0::2], x[..., 1::2]], axis=-1))


@pytest.mark.parametrize("name", ["crs6_sd", "crs7", "planar2_act", "crs6_lin_sd"])
def test_blocked_state_layout(name):
    """Every entry point gives, in blocked order, exactly the bits it gives in interleaved order."""
    import torch
    from reak_b200 import kte_batch_propagator
    s = presets.make(name)
    p, pb = kte_batch_propagator(s), kte_batch_propagator(s, blocked=True)
    n = 777
    x, u = random_batch(p.compiled, n, seed=51)
    xb = _to_blocked(x)
    assert np.array_equal(pb.get_state_derivatives(xb, u)[0], _to_blocked(p.get_state_derivatives(x, u)[0]))
    assert np.array_equal(pb.get_gen_forces(xb, u), p.get_gen_forces(x, u))
    Mb, Mdb = pb.get_mass_matrices(xb, with_derivative=True)
    M, Md = p.get_mass_matrices(x, with_derivative=True)
    assert np.array_equal(Mb, M) and np.array_equal(Mdb, Md)
    assert np.array_equal(pb.get_next_states(xb, u, 1e-3, 7)[0], _to_blocked(p.get_next_states(x, u, 1e-3, 7)[0]))
    # SoA + blocked, device buffers
    xs = pb.get_next_states(torch.from_numpy(np.ascontiguousarray(xb.T)).cuda(), torch.from_numpy(np.ascontiguousarray(u.T)).cuda(),
                            1e-3, 7, soa=True)[0]
    assert np.array_equal(xs.t().cpu().numpy(), _to_blocked(p.get_next_states(x, u, 1e-3, 7)[0]))
    u_seq = np.random.default_rng(52).uniform(-1, 1, (n, 3, p.nu))
    for scheme in ("rk4", "midpoint"):
        xo, tr, _ = p.rollout(x, u_seq, 1e-3, 2, scheme=scheme, want_traj=True)
        xob, trb, _ = pb.rollout(xb, u_seq, 1e-3, 2, scheme=scheme, want_traj=True)
        assert np.array_equal(xob, _to_blocked(xo)) and np.array_equal(trb, _to_blocked(tr))
    if p.nu:
        rng = np.random.default_rng(53)
        P, R = 9, 16
        uu = rng.uniform(-3, 3, (P, R, p.nu))
        i1, b1, c1 = p.steer_batch(x[:P], x[P:2 * P], uu, 1e-3, 5)
        i2, b2, c2 = pb.steer_batch(xb[:P], xb[P:2 * P], uu, 1e-3, 5)
        assert np.array_equal(i1, i2) and np.array_equal(b2, _to_blocked(b1)) and np.allclose(c1, c2, rtol=1e-14)
        goal = x + rng.uniform(-0.3, 0.3, x.shape)
        gain = rng.uniform(-3, 3, (n, p.nu, p.nx))
        gain_b = np.ascontiguousarray(np.concatenate([gain[:, :, 0::2], gain[:, :, 1::2]], axis=2))
        a = p.steer_feedback(x, goal, u, gain, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2, want_traj=True)
        b = pb.steer_feedback(xb, _to_blocked(goal), u, gain_b, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2, want_traj=True)
        assert np.array_equal(a[2], b[2])
        # the law sums gain * (x - goal) in storage order, so the two layouts may differ in the last bits
        assert rel_err(b[0], _to_blocked(a[0])) < 1e-12 and rel_err(b[1], a[1]) < 1e-12


@pytest.mark.parametrize("name", ["crs6", "crs6_sd", "crs7_phys_sd", "planar2_act", "crs6_lin_sd", "crs2d"])
def test_legacy_manipulator_model_state_rate(name, oracle_built):
    """a29: RKB_LAYOUT_BLOCKED serves kte::manipulator_dynamics_model::computeStateRate (ctrl/kte_models/
    manip_dynamics_model.cpp:152-218; ctrl/mbd_kte/manipulator_model.cpp:292-355 is the same code) — checked against
    that class itself, compiled from the reference and assembled over the same chain objects."""
    from reak_b200 import kte_batch_propagator
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    for interp in (False, True):
        pb = kte_batch_propagator(presets.make(name), blocked=True, interpreter=interp)
        R = oracle_built.Reference(pb.compiled)
        x, u = random_batch(pb.compiled, 200, seed=55, q_range=2.0)
        xb = _to_blocked(x)
        want, st_r = R.manip_state_rate(xb, u)
        got, st = pb.get_state_derivatives(xb, u)
        assert not st.any() and not st_r.any() and rel_err(got, want) < TOL_STEP, (name, interp)


def test_reak_steer_space_cpp(oracle_built):
    """ReaK::pp::kte_steer_space (reak_bridge.hpp): steer_position_toward of SteerableSpaceConcept served by the
    batched propagator, checked in C++ against the unmodified reference — every candidate control is
    re-integrated with runge_kutta4_integrator, the winner must be the arg-min, the returned point its end
    state and the steer record its state after every control interval."""
    import ctypes as C
    from reak_b200 import _abi
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    C.CDLL(_abi.LIB_PATH, mode=C.RTLD_GLOBAL)
    for name in ("crs6", "crs7_phys_sd", "planar2_act"):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = oracle_built.Reference(c)
        fn = R.lib.rkref_steer_space_check
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                       C.c_double, C.c_void_p, C.c_char_p, C.c_int]
        P = 5
        a, _ = random_batch(c, P, seed=61)
        b, _ = random_batch(c, P, seed=62)
        lo, hi = -3.0 * np.ones(c.n_inputs), 3.0 * np.ones(c.n_inputs)
        err = np.zeros(3)
        msg = C.create_string_buffer(512)
        rc = fn(R.h, P, a.ctypes.data, b.ctypes.data, 0.3, lo.ctypes.data, hi.ctypes.data, 24, 3, 5, 1e-3, err.ctypes.data, msg, 512)
        assert rc == 0, (name, msg.value)
        assert err[0] < TOL_STEP and err[1] < TOL_STEP and err[2] == 0.0, (name, err)


def test_planner_dispatch_cpp(oracle_built):
    """SURVEY 8(f) rank 1, planner side (oracle/ref_lib.cpp: rkref_planner_dispatch_check): kte_steer_space carries
    the traits ReaK's planners dispatch on, models SteerableSpaceConcept, and the reference's own
    rrg_node_puller::expand_to_nearest (ctrl/graph_alg/node_generators.hpp:59-75) pulls the same vertex, point,
    steer record and edge weight whether the candidates are steered one at a time or all at once by
    batched_steer_visitor; is_free of the space agrees with the reference's findMinimumDistance."""
    import ctypes as C
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    from reak_b200 import proximity as px
    for name, with_env, tol in (("crs6", True, 0.01), ("crs6", False, 0.01), ("crs7", False, 0.01), ("crs6", True, 1e9)):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        R = oracle_built.Reference(c)
        fn = R.lib.rkref_planner_dispatch_check
        fn.restype = C.c_int
        fn.argtypes = ([C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                       C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_char_p, C.c_int])
        K, nx, nu = 7, 2 * c.n_coords, c.n_inputs
        nodes, _ = random_batch(c, K, seed=65, q_range=2.5)
        rng = np.random.default_rng(66)
        target = nodes[3] + rng.uniform(-0.05, 0.05, nx)          # a sample the tree can actually get closer to
        lo, hi = -3.0 * np.ones(nu), 3.0 * np.ones(nu)
        xl, xh = -2.0 * np.ones(nx), 2.0 * np.ones(nx)
        if with_env:
            robot, lab = presets.crs_proxy_models(s)
            pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
            (m1, n1), (m2, n2) = robot.to_c(c.frames), lab.to_c(c.frames)
            a1, a2 = C.cast(m1, C.c_void_p), C.cast(m2, C.c_void_p)
        else:
            a1 = a2 = None
            n1 = n2 = 0
        out, flags = np.zeros(5, dtype=np.int32), np.zeros(2 * K, dtype=np.int32)
        p_new, sample, err = np.zeros(nx), np.zeros(nx), np.zeros(3)
        msg = C.create_string_buffer(512)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = fn(R.h, K, vp(nodes), vp(target), vp(lo), vp(hi), vp(xl), vp(xh), 32, 2, 5, 1e-3, tol, a1, n1, a2, n2,
                vp(out), vp(flags), vp(p_new), vp(sample), vp(err), msg, 512)
        assert rc == 0, (name, msg.value)
        assert out[0] == out[1] and out[3] == 0 and out[4] == K, (name, out)   # same vertex; never the move_position_toward branch
        assert out[2] >= 1 and (out[2] == K if out[0] < 0 else out[2] == out[0] + 1), (name, out)  # stops at the first success
        assert np.all(err == 0.0), (name, err)                                  # bit-identical point, record, weight
        if tol > 1.0:
            assert out[0] == -1                                                 # nothing progresses 1e9 x the distance
        assert np.all((sample >= xl) & (sample <= xh))
        assert np.array_equal(flags[:K], flags[K:])
        if with_env:
            d, _, _ = R.min_distance(pair, nodes)
            assert np.array_equal(flags[:K] != 0, d >= 0.0), (name, flags, d)
            assert (d < 0).any() and (d >= 0).any()
        else:
            assert flags.all()


# ---- the trigonometric paths of the serial kernels -------------------------------------------------------
@pytest.mark.parametrize("name", ["crs6", "crs7"])
def test_large_angles_and_fast_joints(name, oracle_built):
    """sincos_reduced serves |q| < 1e9 (Cody-Waite by pi/2), the library sincos anything beyond; RK4 stages 2-4
    use the small-angle shift only while |dt q_dot| < 2^-5 and evaluate in full otherwise.  All of them against
    the oracle, on the rollout and the evaluation kernels."""
    p = _make(name)
    O = oracle_built.Oracle(p.compiled)
    rng = np.random.default_rng(71)
    n = 96
    x, u = random_batch(p.compiled, n, seed=72)
    rev = [2 * k for k in range(p.n)] if name == "crs6" else [2 * k for k in range(1, p.n)]  # crs7: coordinate 0 is the track
    cases = {}
    big = x.copy()
    big[:, rev] = rng.uniform(-1.0, 1.0, (n, len(rev))) * 10.0 ** rng.integers(0, 9, (n, len(rev)))  # up to 1e8 rad
    cases["angles up to 1e8"] = (big, 1e-3, 1e-10)
    huge = x.copy()
    huge[:, rev[0]] = rng.uniform(2e9, 1e12, n)          # beyond the reduction's range: library path
    cases["angles beyond 1e9"] = (huge, 1e-3, 1e-10)
    fast = x.copy()
    fast[:, [k + 1 for k in rev]] = rng.uniform(-80.0, 80.0, (n, len(rev)))  # |dt q_dot| up to 0.08 > 2^-5
    cases["fast joints"] = (fast, 1e-3, 1e-10)
    mixed = x.copy()
    mixed[::2, rev[1] + 1] = 40.0                        # every other sample leaves the small-angle path: divergent warps
    cases["mixed"] = (mixed, 1e-3, 1e-10)
    for label, (xx, dt, tol) in cases.items():
        xd, st = p.get_state_derivatives(xx, u)
        assert not st.any() and rel_err(xd, O.eval(xx, u)[0]) < tol, (name, label)
        xo, st = p.get_next_states(xx, u, dt, 3)
        xr, sr, _ = O.rk4(xx, u, dt, 3)
        # a state component of 1e8..1e12 carries an absolute rounding error of its own ulp: compare relative to its size
        err = float(np.max(np.abs(xo - xr) / np.maximum(1.0, np.abs(xr))))
        assert not st.any() and not sr.any() and err < tol, (name, label, err)


# ---- torsion springs where the kernels differ from the reference by construction -----------------------------
# The reference takes axis_angle(conj(Q1) Q2): angle 2 acos(w) with a dead zone |sin(q/2)| <= 1e-7, the axis flipped
# for w < 0 (rotations_3D.hpp:1985-2010, torsion_spring.cpp:106-129).  The serial kernels use the wrapped joint
# angle, the interpreter 2 atan2(|v|, |w|).  acos is ill-conditioned at w -> 1: the reference's OWN angle carries an
# absolute error of about eps / |sin(q/2)| there (measured on the oracle: 8e-11 rad at q = 2e-7, 4e-11 at 1e-6),
# which the kernels do not reproduce; the tolerance below allows exactly that, per spring, and is 1e-10 elsewhere.
EDGE_ANGLES = [0.0, 1e-8, -1e-8, 2e-7 - 1e-12, -(2e-7 - 1e-12), 2e-7 + 1e-12, -(2e-7 + 1e-12), 1e-6, -1e-5,
               np.pi - 1e-9, np.pi + 1e-9, -np.pi + 1e-9, -np.pi - 1e-9, 2 * np.pi + 1e-8, 2 * np.pi + 1e-6,
               -2 * np.pi - 3e-7, 40.0, -40.0, 12 * np.pi, 12 * np.pi + 1e-8, 39.0 * np.pi - 1e-9]


def _spring_tolerance(p, x):
    """per sample: 1e-10 + sum over torsion_spring_3D of stiffness * 4.5e-16 / |sin(q/2)| outside the dead zone"""
    from reak_b200 import _abi
    d = p.compiled.desc
    tol = np.full(x.shape[0], TOL_STEP)
    joint_of_end = {}
    for e in range(d.n_elements):
        E = d.elements[e]
        if E.kind in (_abi.REVOLUTE_3D,):
            joint_of_end[E.frame_b] = E.coord
    for e in range(d.n_elements):
        E = d.elements[e]
        if E.kind == _abi.TORSION_SPRING_3D and E.frame_b in joint_of_end:
            s = np.abs(np.sin(0.5 * x[:, 2 * joint_of_end[E.frame_b]]))
            tol += np.where(s > 1e-7, abs(E.p[0]) * 4.5e-16 / np.maximum(s, 1e-7), 0.0)
    return tol


@pytest.mark.parametrize("name", ["crs6_sd", "crs6_sd_sat", "crs7_phys_sd", "planar3_sd", "torsion1"])
def test_torsion_spring_edges(name, oracle_built):
    """Dead zone, its two edges, q = +-pi from both sides, whole turns plus a hair, |q| up to 40 rad, with and
    without saturation, 3D and 2D springs — serial kernels and interpreter against the oracle (= the reference)."""
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        n = p.n
        first = 1 if name.startswith("crs7") else 0  # crs7: coordinate 0 is the prismatic track
        rng = np.random.default_rng(81)
        rows = []
        for a in EDGE_ANGLES:            # every revolute joint at the edge angle, velocities random
            r = rng.uniform(-1.0, 1.0, 2 * n)
            r[2 * first::2] = a
            rows.append(r)
        for a in EDGE_ANGLES:            # one joint at the edge, the others anywhere in +-40 rad
            for k in range(first, n):
                r = rng.uniform(-1.0, 1.0, 2 * n)
                r[2 * first::2] = rng.uniform(-40.0, 40.0, n - first)
                r[2 * k] = a
                rows.append(r)
        x = np.array(rows)
        u = rng.uniform(-1.0, 1.0, (x.shape[0], p.nu))
        tol = _spring_tolerance(p, x)
        f, f_o = p.get_gen_forces(x, u), O.gen_forces(x, u)
        err = np.max(np.abs(f - f_o) / np.maximum(1.0, np.abs(f_o)), axis=1)
        assert np.all(err < tol), (name, label, "f", float(np.max(err / tol)), x[np.argmax(err / tol)])
        xd, st = p.get_state_derivatives(x, u)
        xd_o, st_o = O.eval(x, u)
        err = np.max(np.abs(xd - xd_o) / np.maximum(1.0, np.abs(xd_o)), axis=1)
        assert not st.any() and not st_o.any() and np.all(err < tol), (name, label, "xdot", float(np.max(err / tol)))
        # rollouts that start far from the discontinuities: |q| up to 40 rad, saturation active where configured
        xr = rng.uniform(-1.0, 1.0, (64, 2 * n))
        xr[:, 2 * first::2] = rng.uniform(-40.0, 40.0, (64, n - first))
        w = xr[:, 2 * first::2] - 2 * np.pi * np.rint(xr[:, 2 * first::2] / (2 * np.pi))
        keep = np.all((np.abs(w) > 0.05) & (np.abs(np.abs(w) - np.pi) > 0.05), axis=1)
        xr, ur = xr[keep], rng.uniform(-1.0, 1.0, (int(keep.sum()), p.nu))
        assert xr.shape[0] >= 16
        xo, st = p.get_next_states(xr, ur, 1e-3, 20)
        xo_o, st_o, _ = O.rk4(xr, ur, 1e-3, 20)
        assert not st.any() and not st_o.any() and rel_err(xo, xo_o) < TOL_LONG, (name, label, "rk4")


def test_non_finite_states_raise_the_status_bit():
    p = _make("crs6")
    x, u = random_batch(p.compiled, 64, seed=73)
    x[5, 0] = np.nan
    x[9, 3] = np.inf
    xd, st = p.get_state_derivatives(x, u)
    xo, st2 = p.get_next_states(x, u, 1e-3, 2)
    from reak_b200 import _abi
    for s in (st, st2):
        assert s[5] & _abi.STATUS_NONFINITE and s[9] & _abi.STATUS_NONFINITE
        assert not np.delete(s, [5, 9]).any()
    assert np.isfinite(np.delete(xo, [5, 9], axis=0)).all()


@pytest.mark.parametrize("name", ALL)
def test_twist_shaping_matrices(name, oracle_built):
    """rkb_twist_shaping = mass_matrix_calc::get_TMT_TdMT: Tcm, Mcm, Tcm_dot against the oracle, and
    M = Tcm^T Mcm Tcm, Mdot = Tcm_dot^T Mcm Tcm + transpose against rkb_mass_matrix (serial kernels where they apply)."""
    p = _make(name)
    O = oracle_built.Oracle(p.compiled)
    x, _ = random_batch(p.compiled, 40, seed=23, q_range=2.0)
    T, Mc, Td = p.get_twist_shaping(x)
    for i in (0, 7, 39):
        To, Mo, Tdo = O.tmt(x[i:i + 1])
        assert np.array_equal(Mc, Mo)
        assert rel_err(T[i], To) < TOL_STEP and rel_err(Td[i], Tdo) < TOL_STEP, (name, i)
    M, Md = p.get_mass_matrices(x, with_derivative=True)
    MT = np.einsum("rs,nsc->nrc", Mc, T)
    M2 = np.einsum("nra,nrb->nab", T, MT)
    S = np.einsum("nra,nrb->nab", Td, MT)
    assert rel_err(M, 0.5 * (M2 + M2.transpose(0, 2, 1))) < TOL_STEP and rel_err(Md, S + S.transpose(0, 2, 1)) < TOL_STEP
    only_T, _ = p.get_twist_shaping(x, with_derivative=False)
    assert np.array_equal(only_T, T)


@pytest.mark.parametrize("name", ALL)
def test_frames_after_motion_and_force(name, oracle_built):
    """rkb_frames: kinematics and wrenches of every frame after doMotion / clearForce / doForce, against the oracle
    (which matches the live reference frame by frame, tests/test_oracle.py)."""
    p = _make(name)
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 33, seed=27, q_range=2.0)
    fr = p.get_frames(x, u)
    assert fr.shape == (33, p.compiled.desc.n_frames, 25)
    for i in (0, 16, 32):
        assert rel_err(fr[i], O.frames(x[i:i + 1], u[i:i + 1])) < TOL_STEP, (name, i)


def test_cpp_host_class_builds_the_same_chain(tmp_path):
    """examples/cpp_host_demo.cpp assembles the 6-DOF CRS arm with reak_b200::chain_builder and runs it through
    kte_batch_propagator (C++11, no ReaK); the Python mirror of the same model must give the same bits."""
    import subprocess
    from test_abi import _build_cpp_demo
    exe, env = _build_cpp_demo(tmp_path)
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    lines = r.stdout.split("\n")
    vals = np.array([[float(t) for t in l.split()] for l in lines if l and not l.startswith("M") and not l.startswith("D")])
    prox = np.array([[float(t) for t in l.split()[1:]] for l in lines if l.startswith("D")])
    Mrow = np.array([float(l.split()[1]) for l in lines if l.startswith("M")])
    N = 64
    s = 88172645463325252
    raw = []
    for _ in range(N * 18):
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        raw.append((s >> 11) / 9007199254740992.0 * 2.0 - 1.0)
    x, u = np.array(raw[:N * 12]).reshape(N, 12), np.array(raw[N * 12:]).reshape(N, 6)
    p = _make("crs6")
    xd, _ = p.get_state_derivatives(x, u)
    xo, _ = p.get_next_states(x, u, 1e-3, 25)
    assert np.array_equal(vals[:, 0].reshape(N, 12), xd) and np.array_equal(vals[:, 1].reshape(N, 12), xo)
    assert np.array_equal(Mrow.reshape(6, 6), p.get_mass_matrices(x[:1])[0])
    # the proximity models the demo writes down as rkb_shape literals are those of presets.crs_proxy_models
    from reak_b200 import proximity as px
    sys6 = presets.make("crs6")
    from reak_b200.propagator import kte_batch_propagator
    p6 = kte_batch_propagator(sys6)
    robot, lab = presets.crs_proxy_models(sys6)
    d, f = p6.get_min_distances(px.proxy_query_pair_3D("robot-lab", robot, lab), 3.0 * x, with_points=False)
    # the quaternion literals of the demo differ from axis_angle's in the last bit, and the arm's first capsule sits
    # exactly between the two track capsules (lab shapes 3 and 4): which of the two wins that tie is rounding
    pf = prox[:, 1].astype(np.int32)
    assert prox.shape == (N, 2) and np.max(np.abs(prox[:, 0] - d)) < 1e-12
    assert np.all((pf == f) | ((pf // 5 == f // 5) & (pf % 5 >= 3) & (f % 5 >= 3))) and np.mean(pf == f) > 0.8


def test_streams_and_threads():
    """Device buffers run on the caller's CUDA stream; one handle may be shared by threads (its calls are
    serialised), two handles run concurrently.  Everything must reproduce the default-stream result."""
    import threading
    import torch
    p, p2 = _make("crs6"), _make("crs6")
    x, u = random_batch(p.compiled, 20000, seed=81)
    ref, _ = p.get_next_states(x, u, 1e-3, 8)
    xt, ut = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        a, _ = p.get_next_states(xt, ut, 1e-3, 4)
        b, _ = p.get_next_states(a, ut, 1e-3, 4)      # ordered after `a` on the same stream
    side.synchronize()
    assert np.array_equal(b.cpu().numpy(), ref)
    results, errors = {}, []

    def work(tag, prop, lo, hi):
        try:
            for _ in range(5):
                results[tag] = prop.get_next_states(x[lo:hi], u[lo:hi], 1e-3, 8)[0]
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=("s0", p, 0, 10000)), threading.Thread(target=work, args=("s1", p, 10000, 20000)),
               threading.Thread(target=work, args=("o0", p2, 0, 20000))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    assert np.array_equal(np.concatenate([results["s0"], results["s1"]]), ref) and np.array_equal(results["o0"], ref)


def test_steer_feedback_fused_kernel_equals_the_launch_per_interval_path():
    """Serial chains run the whole steering loop in one launch; RKB_OPT_FUSED_STEER = 0 selects the law-kernel +
    rollout-per-interval path the interpreter chains use.  Same decisions, results within rounding."""
    p = _make("crs6_sd")
    x0, goal, u_bias, gain, u_prev = _steer_case(p, 500, seed=91)
    goal[:5] = x0[:5]                                         # these never start
    goal[5:60] = x0[5:60] + 0.62 / np.sqrt(p.nx)               # these start just outside the proximity ball
    kw = dict(bounds=(-2 * np.ones(p.nu), 2 * np.ones(p.nu)), rate_bounds=(-60 * np.ones(p.nu), 60 * np.ones(p.nu)), want_traj=True)
    a = p.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, 7, 0.6, **kw)
    p.set_option("fused_steer", 0)
    assert p.get_option("fused_steer") == 0
    b = p.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, 7, 0.6, **kw)
    p.set_option("fused_steer", 1)
    assert np.array_equal(a[2], b[2]) and len(set(a[2].tolist())) > 1
    assert rel_err(a[0], b[0]) < 1e-12 and rel_err(a[1], b[1]) < 1e-12
    for i in range(500):
        assert rel_err(a[3][i, :a[2][i]], b[3][i, :b[2][i]]) < 1e-12


def test_rollout_sequence_fused_kernel_equals_the_launch_per_interval_path():
    """RK4 control sequences on the serial kernels run in one launch; RKB_OPT_FUSED_SEQUENCE = 0 selects one launch
    per interval (what the other schemes and the interpreter chains use).  Bit-identical."""
    p = _make("crs7")
    x, _ = random_batch(p.compiled, 3000, seed=93)
    u_seq = np.random.default_rng(94).uniform(-2, 2, (3000, 6, p.nu))
    a = p.rollout(x, u_seq, 1e-3, 5, scheme="rk4", want_traj=True)
    p.set_option("fused_sequence", 0)
    b = p.rollout(x, u_seq, 1e-3, 5, scheme="rk4", want_traj=True)
    p.set_option("fused_sequence", 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not a[2].any() and not b[2].any()


# ---- direct-kinematics Jacobian of a frame -----------------------------------------------
@pytest.mark.parametrize("name", ["crs6", "crs7_phys_sd", "crs6_twist", "planar3_sd", "crs2d"])
def test_frame_jacobian(name, oracle_built):
    """rkb_frame_jacobian = jacobian_gen_3D / _2D::get_jac_relative_to(frame) of every upstream joint, the rows
    manip_kin_mdl_jac_calculator stacks for an end-effector frame.  For the frame of an inertia with that inertia's
    upstream set these are that inertia's rows of Tcm / Tcm_dot (checked against the oracle = the reference's
    get_TMT_TdMT); for the last frame of the chain with every coordinate upstream, J q_dot must be the frame's twist
    as doMotion leaves it (rkb_frames) and Jdot q_dot its acceleration with q_ddot = 0."""
    from reak_b200 import _abi
    p = _make(name)
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 50, seed=151, q_range=2.0)
    d = p.compiled.desc
    dim3 = d.dim == 3
    rows_per = 6 if dim3 else 3
    row = sum(1 for e in range(d.n_elements) if d.elements[e].kind == _abi.INERTIA_GEN)
    T, Mc, Td = p.get_twist_shaping(x)
    checked = 0
    for e in range(d.n_elements):
        E = d.elements[e]
        if E.kind not in (_abi.INERTIA_3D, _abi.INERTIA_2D):
            continue
        up = [c for c in range(p.n) if (E.upstream >> c) & 1]
        J, Jd = p.get_frame_jacobian(x, int(E.frame_a), upstream=up)
        assert np.array_equal(J, T[:, row:row + rows_per, :]) and np.array_equal(Jd, Td[:, row:row + rows_per, :]), (name, e)
        To, _, Tdo = O.tmt(x[3:4])
        assert rel_err(J[3], To[row:row + rows_per]) < TOL_STEP and rel_err(Jd[3], Tdo[row:row + rows_per]) < TOL_STEP
        row += rows_per
        checked += 1
    assert checked >= 2
    # the end effector: last frame written by the chain, every coordinate upstream
    last = [int(d.elements[e].frame_b) for e in range(d.n_elements)
            if d.elements[e].kind in (_abi.REVOLUTE_3D, _abi.PRISMATIC_3D, _abi.RIGID_LINK_3D, _abi.REVOLUTE_2D, _abi.PRISMATIC_2D, _abi.RIGID_LINK_2D)][-1]
    J, Jd = p.get_frame_jacobian(x, last)
    fr = p.get_frames(x, u)[:, last, :]
    qd = x[:, 1::2]
    twist = np.einsum("nrc,nc->nr", J, qd)
    if dim3:
        # frame_3D: Velocity is in parent (world) coordinates, AngVelocity local; the Jacobian rows are both local
        from scipy.spatial.transform import Rotation
        Rw = Rotation.from_quat(fr[:, [4, 5, 6, 3]]).as_matrix()
        v_local = np.einsum("nji,nj->ni", Rw, fr[:, 7:10])
        assert rel_err(twist[:, :3], v_local) < 1e-9 and rel_err(twist[:, 3:], fr[:, 10:13]) < 1e-9
    else:
        c, s_ = fr[:, 3], fr[:, 4]
        v_local = np.stack([c * fr[:, 7] + s_ * fr[:, 8], -s_ * fr[:, 7] + c * fr[:, 8]], axis=1)
        assert rel_err(twist[:, :2], v_local) < 1e-9 and rel_err(twist[:, 2], fr[:, 10]) < 1e-9
    J_only = p.get_frame_jacobian(x, last, with_derivative=False)
    assert np.array_equal(J_only, J)
    lib = _abi.load_library()
    import ctypes as C
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.rkb_frame_jacobian(p._h, 0, 50, vp(x), 999, 1, vp(J), None, 0, None) == _abi.ERR_INVALID
    assert lib.rkb_frame_jacobian(p._h, 0, 50, vp(x), last, 1 << 20, vp(J), None, 0, None) == _abi.ERR_INVALID


# ---- RK4 with an input trajectory (a27) ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["crs6", "crs6_sd", "crs7_phys_sd", "planar2_act", "crs6_lin_sd", "planar2"])
def test_rk4_with_an_input_trajectory(name, oracle_built):
    """rkb_rollout_rk4_inputs = ctrl::detail::runge_kutta4_integrate_impl (runge_kutta4_integrator_sys.hpp:50-97): the
    input is read at t, t + dt/2 (twice) and t + dt.  Against the oracle's restatement (pinned on the CPU against the
    reference's runge_kutta4_integrator with a rate function that reads the node of the time it is asked at), on the
    serial kernels and the interpreter, AoS and SoA, host and device buffers; equal nodes reproduce rkb_rollout_rk4."""
    import torch
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        n, K = 301, 9
        x, u = random_batch(p.compiled, n, seed=121)
        nodes = np.random.default_rng(122).uniform(-2.0, 2.0, (n, 2 * K + 1, p.nu))
        got, st = p.get_next_states_input_trajectory(x, nodes, 1e-3)
        want, st_o = O.rk4_inputs(x, nodes, 1e-3)
        assert not st.any() and not st_o.any() and rel_err(got, want) < TOL_STEP, (name, label)
        if p.nu:
            held, _ = p.get_next_states(x, nodes[:, 0, :].copy(), 1e-3, K)
            assert rel_err(got, held) > 1e-6                         # the trajectory matters ...
            same, _ = p.get_next_states_input_trajectory(x, np.repeat(nodes[:, :1, :], 2 * K + 1, axis=1), 1e-3)
            assert rel_err(same, held) < 1e-13, (name, label)        # ... and equal nodes are a held input
        dev, _ = p.get_next_states_input_trajectory(torch.from_numpy(x).cuda(), torch.from_numpy(nodes).cuda(), 1e-3)
        assert np.array_equal(dev.cpu().numpy(), got), (name, label)
    # SoA through the C-ABI: [2K+1][nu][N]
    import ctypes as C
    from reak_b200 import _abi
    p = _make("crs6")
    x, _ = random_batch(p.compiled, 257, seed=123)
    nodes = np.random.default_rng(124).uniform(-1, 1, (257, 7, 6))
    ref, _ = p.get_next_states_input_trajectory(x, nodes, 1e-3)
    xs, us = np.ascontiguousarray(x.T), np.ascontiguousarray(nodes.transpose(1, 2, 0))
    out = np.empty_like(xs)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib = _abi.load_library()
    assert lib.rkb_rollout_rk4_inputs(p._h, 0, 257, vp(xs), vp(us), 1e-3, 3, vp(out), None, _abi.LAYOUT_SOA, None) == 0
    assert np.array_equal(out.T, ref)
    assert lib.rkb_rollout_rk4_inputs(p._h, 0, 257, vp(xs), vp(us), 0.0, 3, vp(out), None, 0, None) == _abi.ERR_INTEGRATION
    assert lib.rkb_rollout_rk4_inputs(p._h, 0, 257, vp(xs), None, 1e-3, 3, vp(out), None, 0, None) == _abi.ERR_INVALID


# ---- linearisation ---------------------------------------------------------------------
@pytest.mark.parametrize("name", ["crs6", "crs6_sd", "crs7", "planar2_act", "crs6_lin_sd", "pendulum"])
def test_linear_blocks_by_central_differences(name, oracle_built):
    """rkb_linearize: A = d xdot / d x, B = d xdot / d u about (x, u) by central differences with the step
    h = eps max(1, |component|).  Checked against the same differences of the ORACLE's get_state_derivative
    (a difference quotient amplifies the 1e-15 agreement of the evaluations by 1 / (2 h): tolerance 1e-7 at
    eps = 1e-6), against the structure every mechanical system has (d q_dot / d q = 0, d q_dot / d q_dot = 1,
    d q_dot / d u = 0), and against M^-1 for d q_ddot / d u of an actuated joint."""
    for label, p in _variants(name):
        O = oracle_built.Oracle(p.compiled)
        n, nx, nu = 37, p.nx, p.nu
        x, u = random_batch(p.compiled, n, seed=111, q_range=1.5)
        eps = 1e-6
        A, B, st = p.get_linear_blocks(x, u if nu else None, eps)
        assert not st.any()
        Ao, Bo = np.empty((n, nx, nx)), np.empty((n, nx, nu))
        for d in range(nx + nu):
            xp, xm, up_, um = x.copy(), x.copy(), u.copy(), u.copy()
            if d < nx:
                h = eps * np.maximum(1.0, np.abs(x[:, d]))
                xp[:, d] += h; xm[:, d] -= h
                den = xp[:, d] - xm[:, d]
            else:
                h = eps * np.maximum(1.0, np.abs(u[:, d - nx]))
                up_[:, d - nx] += h; um[:, d - nx] -= h
                den = up_[:, d - nx] - um[:, d - nx]
            col = (O.eval(xp, up_)[0] - O.eval(xm, um)[0]) / den[:, None]
            if d < nx:
                Ao[:, :, d] = col
            else:
                Bo[:, :, d - nx] = col
        # torsion_spring_3D: the reference's own 2 acos(w) carries ~1e-16 / |sin(q/2)| of noise (see the spring edge test),
        # which the difference quotient amplifies as well
        tol = 2e-6 if "_sd" in name else 1e-7
        assert rel_err(A, Ao) < tol, (name, label)
        # kinematic rows: q_dot does not depend on q or u, and is the identity in q_dot
        assert np.abs(A[:, 0::2, 0::2]).max() < 1e-9 and np.abs(A[:, 0::2, 1::2] - np.eye(p.n)[None]).max() < 1e-9
        if nu:
            assert rel_err(B, Bo) < tol, (name, label)
            assert np.abs(B[:, 0::2, :]).max() == 0.0
        if name == "crs7":   # the track is prismatic and first in the chain: nothing below it removes the axial part, so
            Minv = np.linalg.inv(p.get_mass_matrices(x))                  # d q_ddot / d u_0 is column 0 of M^-1
            assert rel_err(B[:, 1::2, 0], Minv[:, :, 0]) < 1e-6
    # argument checks
    import ctypes as C
    from reak_b200 import _abi
    p = _make("crs6")
    lib = _abi.load_library()
    x, u = random_batch(p.compiled, 4, seed=1)
    A = np.empty((4, 12, 12))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.rkb_linearize(p._h, 0, 4, vp(x), vp(u), 1e-6, None, None, None, 0, None) == _abi.ERR_INVALID
    assert lib.rkb_linearize(p._h, 0, 4, vp(x), vp(u), 1e-6, vp(A), None, None, _abi.LAYOUT_SOA, None) == _abi.ERR_UNSUPPORTED
    assert lib.rkb_linearize(p._h, 0, 4, vp(x), vp(u), 0.0, vp(A), None, None, 0, None) == 0   # eps <= 0: the default step
    big_x, big_u = random_batch(p.compiled, 40000, seed=2)                                  # more than one slice
    A2, B2, st2 = p.get_linear_blocks(big_x, big_u)
    A3, B3, _ = p.get_linear_blocks(big_x[16000:16400], big_u[16000:16400])
    assert np.array_equal(A2[16000:16400], A3) and np.array_equal(B2[16000:16400], B3) and not st2.any()


# ---- small batches: one sample on a pair of warps ----------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_pair_of_warps_kernels_equal_thread_per_sample(name):
    """RKB_OPT_SPLIT_MAX_SAMPLES: rollouts, control sequences and steering loops of small batches run with the forces
    on one warp and the mass matrix + solve on its partner.  The same instruction sequences run, so the results are
    those of the one-thread-per-sample kernels BIT FOR BIT — for every serial preset, ragged batch sizes (lanes and
    whole pairs beyond the batch), shared start states (steer batch), both state orders, and loops whose samples stop
    after different numbers of intervals."""
    from reak_b200 import kte_batch_propagator
    duo = _make(name, split=1 << 20)
    if not duo.is_serial():
        pytest.skip("interpreter chain")
    assert duo.get_option("split_max_samples") == 1 << 20 and _make(name).get_option("split_max_samples") == (8192 if duo.n >= 4 else 0)
    solo = _make(name, split=0)
    rng = np.random.default_rng(101)
    for n in (1, 31, 33, 64, 65, 1000):
        x, u = random_batch(duo.compiled, n, seed=102 + n, q_range=2.0)
        a, sa = duo.get_next_states(x, u, 1e-3, 9)
        b, sb = solo.get_next_states(x, u, 1e-3, 9)
        assert np.array_equal(a, b) and np.array_equal(sa, sb), (name, n)
    n = 777
    x, u = random_batch(duo.compiled, n, seed=103)
    useq = rng.uniform(-2, 2, (n, 4, duo.nu))
    a = duo.rollout(x, useq, 1e-3, 3, scheme="rk4", want_traj=True)
    b = solo.rollout(x, useq, 1e-3, 3, scheme="rk4", want_traj=True)
    assert all(np.array_equal(p, q) for p, q in zip(a, b)), name
    if duo.nu:
        P, R = 9, 37
        uu = rng.uniform(-4, 4, (P, R, duo.nu))
        a = duo.steer_batch(x[:P], x[P:2 * P], uu, 1e-3, 6, want_status=True)
        b = solo.steer_batch(x[:P], x[P:2 * P], uu, 1e-3, 6, want_status=True)
        assert all(np.array_equal(p, q) for p, q in zip(a, b)), name
        x0, goal, u_bias, gain, u_prev = _steer_case(duo, 300, seed=104)
        goal[:5] = x0[:5]
        goal[5:40] = x0[5:40] + 0.62 / np.sqrt(duo.nx)
        kw = dict(bounds=(-2 * np.ones(duo.nu), 2 * np.ones(duo.nu)), rate_bounds=(-60 * np.ones(duo.nu), 60 * np.ones(duo.nu)), want_traj=True)
        for sat in (False, True):
            a = duo.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 5, 6, 0.6, saturate_first=sat, **kw)
            b = solo.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 5, 6, 0.6, saturate_first=sat, **kw)
            nd = a[2]
            assert np.array_equal(nd, b[2]) and len(set(nd.tolist())) > 1, name
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[4], b[4]), name
            for i in range(300):
                assert np.array_equal(a[3][i, :nd[i]], b[3][i, :nd[i]]), (name, i)
    # the (q..., qd...) state order goes through the same kernels
    s = presets.make(name)
    db, sb_ = kte_batch_propagator(s, blocked=True), kte_batch_propagator(presets.make(name), blocked=True)
    db.set_option("split_max_samples", 1 << 20)
    sb_.set_option("split_max_samples", 0)
    assert np.array_equal(db.get_next_states(x[:100], u[:100], 1e-3, 4)[0], sb_.get_next_states(x[:100], u[:100], 1e-3, 4)[0])


def test_pair_of_warps_is_only_used_below_the_threshold():
    """above RKB_OPT_SPLIT_MAX_SAMPLES the one-thread-per-sample kernel runs: same launch count, same results; the
    threshold only moves work between two bit-identical kernels"""
    p = _make("crs6", split=100)
    x, u = random_batch(p.compiled, 101, seed=105)
    big, _ = p.get_next_states(x, u, 1e-3, 3)        # 101 samples: thread per sample
    small, _ = p.get_next_states(x[:100], u[:100], 1e-3, 3)   # 100 samples: pairs of warps
    assert np.array_equal(big[:100], small)
    from reak_b200 import _abi
    assert _abi.load_library().rkb_chain_set_option(p._h, 99, 1) == _abi.ERR_INVALID
    assert _abi.load_library().rkb_chain_set_option(p._h, _abi.OPT_SPLIT_MAX_SAMPLES, -2) == _abi.ERR_INVALID


# ---- run-time specialisation (NVRTC) ---------------------------------------------------------------------
@pytest.mark.parametrize("which", ["era", "ssrms", "crs6_sd"])
def test_runtime_specialised_kernels(which, oracle_built):
    """rkb_chain_specialize compiles the serial kernels for the chain's own structure.  The ERA / SSRMS joint
    layouts are not among the shipped shapes (general code until specialised); every entry point must agree
    with the oracle before and after, and the two sets of kernels with each other."""
    from reak_b200 import kte_batch_propagator
    if which == "era":
        s = presets.crs_chain(n_revolute=7, axes=presets.ERA_AXES, link_offsets=presets.ERA_LINKS, physical=True, springs=True)
    elif which == "ssrms":
        s = presets.crs_chain(n_revolute=7, axes=presets.SSRMS_AXES, link_offsets=presets.SSRMS_LINKS)
    else:
        s = presets.make("crs6_sd")
    p = kte_batch_propagator(s)
    assert p.is_serial() and not p.is_specialized()
    if which != "crs6_sd":
        assert p.kernel_shape() == 0  # general code
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 300, seed=97)
    u_seq = np.random.default_rng(98).uniform(-1, 1, (300, 3, p.nu))

    def everything():
        xd, st = p.get_state_derivatives(x, u)
        f = p.get_gen_forces(x, u)
        M, Md = p.get_mass_matrices(x, with_derivative=True)
        M1 = p.get_mass_matrices(x)
        xo, st2 = p.get_next_states(x, u, 1e-3, 12)
        xs, tr, _ = p.rollout(x, u_seq, 1e-3, 4, scheme="rk4", want_traj=True)
        xe, _ = p.rollout(x, u_seq, 1e-3, 4, scheme="rk5")
        sf = p.steer_feedback(x, x + 0.2, u, np.ones((300, p.nu, p.nx)) * 0.3, 0.5 * u, 1e-2, 1e-3, 5, 3, 0.1,
                              bounds=(-2 * np.ones(p.nu), 2 * np.ones(p.nu)))
        assert not st.any() and not st2.any()
        return [xd, f, M, Md, M1, xo, xs, tr, xe, sf[0], sf[1]]

    before = everything()
    p.specialize()
    assert p.is_specialized() and p.kernel_shape() != 0
    after = everything()
    for a, b in zip(before, after):
        assert rel_err(a, b) < 1e-11
    assert rel_err(after[0], O.eval(x, u)[0]) < TOL_STEP and rel_err(after[1], O.gen_forces(x, u)) < TOL_STEP
    Mo, Mdo = O.mass(x)
    assert rel_err(after[2], Mo) < TOL_STEP and rel_err(after[3], Mdo) < TOL_STEP
    assert rel_err(after[5], O.rk4(x, u, 1e-3, 12)[0]) < TOL_LONG
    xr, trr, _ = O.rollout(x, u_seq, 4, 1e-3, 4)
    assert rel_err(after[6], xr) < TOL_STEP and rel_err(after[7], trr) < TOL_STEP
    # a second handle of the same structure reuses the compiled kernels
    p2 = kte_batch_propagator(s).specialize()
    assert np.array_equal(p2.get_next_states(x, u, 1e-3, 12)[0], after[5])


def test_automatic_specialisation_and_its_disk_cache(tmp_path, oracle_built):
    """RKB_OPT_AUTO_SPECIALIZE (default on): a serial chain whose structure the shipped kernels do not match asks for
    its own kernels on the first call of >= 4096 samples — NVRTC on a background thread, the calls made meanwhile run
    on the shipped kernels — and leaves the cubin in the disk cache, from which a later PROCESS is served on its
    first call.  Small calls never trigger it; results stay within rounding of the oracle throughout."""
    import subprocess
    import sys
    import time
    script = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
from reak_b200 import kte_batch_propagator, presets
s = presets.crs_chain(n_revolute=7, axes=presets.ERA_AXES, link_offsets=presets.ERA_LINKS, physical=True)
p = kte_batch_propagator(s)
rng = np.random.default_rng(5)
x, u = rng.uniform(-1, 1, (8192, p.nx)), rng.uniform(-1, 1, (8192, p.nu))
small, _ = p.get_next_states(x[:100], u[:100], 1e-3, 3)
assert not p.is_specialized()
first, _ = p.get_next_states(x, u, 1e-3, 3)
at_first = p.is_specialized()
t0 = time.time()
calls = 1
while not p.is_specialized() and time.time() - t0 < 120:
    time.sleep(0.2)
    p.get_next_states(x, u, 1e-3, 3)
    calls += 1
last, _ = p.get_next_states(x, u, 1e-3, 3)
off = kte_batch_propagator(s).set_option("auto_specialize", 0)
plain, _ = off.get_next_states(x, u, 1e-3, 3)
print(json.dumps({"at_first": bool(at_first), "specialised": bool(p.is_specialized()), "calls": calls, "wait_s": time.time() - t0,
                  "drift": float(np.abs(first - last).max()), "vs_plain": float(np.abs(plain - last).max()),
                  "off_specialised": bool(off.is_specialized()), "shape": p.kernel_shape()}))
""" % ROOT
    env = dict(os.environ, RKB_CACHE_DIR=str(tmp_path / "cache"))

    def run():
        r = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        import json
        return json.loads(r.stdout.strip().splitlines()[-1])

    a = run()
    assert a["specialised"] and not a["at_first"] and a["calls"] > 1, a        # compiled in the background, several calls on the shipped kernels
    assert a["shape"] != 0 and a["drift"] < 1e-11 and a["vs_plain"] < 1e-11 and not a["off_specialised"], a
    files = list((tmp_path / "cache").glob("*.cubin"))
    assert len(files) == 1 and files[0].stat().st_size > 100000
    b = run()
    assert b["at_first"] and b["calls"] == 1 and b["wait_s"] < 1.0, b          # served from the disk cache on the first large call
    assert b["drift"] == 0.0
    # a damaged cache file is recompiled, not trusted
    files[0].write_bytes(b"RKBCUBIN 12\n" + b"x\n" * 8 + b"not a cubin!")
    c = run()
    assert c["specialised"] and not c["at_first"], c


def test_specialised_kernels_on_every_device():
    """An 8-joint chain needs more than 48 KB of dynamic shared memory per CTA: the opt-in is per device, so the
    run-time specialised kernels must be prepared on every device the handle touches (all visible GPUs)."""
    import torch
    from reak_b200 import kte_batch_propagator
    s = presets.crs_chain(n_revolute=8, axes=presets.ERA_AXES, link_offsets=presets.ERA_LINKS)
    p = kte_batch_propagator(s)
    x, u = random_batch(p.compiled, 70001, seed=99)
    ref, _ = p.get_next_states(x, u, 1e-3, 3)
    Mref = p.get_mass_matrices(x[:500])
    p.specialize()
    n_dev = torch.cuda.device_count()
    got, st = p.get_next_states_multi(x, u, 1e-3, 3, devices=list(range(n_dev)))
    assert not st.any() and rel_err(got, ref) < 1e-11
    assert rel_err(p.get_mass_matrices(x[:500]), Mref) < 1e-11


@pytest.mark.parametrize("name", ALL)
def test_every_serial_preset_specialises(name, oracle_built):
    """Run-time specialisation over every preset the serial kernels take (1 to 7 coordinates, planar chains,
    prismatic joints, springs, full tensors, twisted links): same results as the shipped kernels, parity with the oracle."""
    p = _make(name)
    if not p.is_serial():
        from reak_b200 import _abi
        with pytest.raises(_abi.RkbError) as e:
            p.specialize()
        assert e.value.code == _abi.ERR_UNSUPPORTED
        return
    O = oracle_built.Oracle(p.compiled)
    x, u = random_batch(p.compiled, 200, seed=101)
    before = (p.get_state_derivatives(x, u)[0], p.get_next_states(x, u, 1e-3, 10)[0], p.get_mass_matrices(x, with_derivative=True)[1])
    p.specialize()
    after = (p.get_state_derivatives(x, u)[0], p.get_next_states(x, u, 1e-3, 10)[0], p.get_mass_matrices(x, with_derivative=True)[1])
    for a, b in zip(before, after):
        assert rel_err(a, b) < 1e-11, name
    assert rel_err(after[0], O.eval(x, u)[0]) < TOL_STEP and rel_err(after[1], O.rk4(x, u, 1e-3, 10)[0]) < TOL_STEP
