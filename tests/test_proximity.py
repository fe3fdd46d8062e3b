"""Proximity queries (rkb_min_distance) against the compiled reference.

CPU (`-m "not gpu"`): the DEVICE source of the finders (reak_b200/csrc/kte_proximity.cuh) is compiled for the
host by tests/host_build/prox_host.cpp and fed the program rkb_proxy_create lowered (host-only code of the
product library, no GPU needed); its answers are held against
  (a) the live reference, proxy_query_pair_3D::findMinimumDistance of oracle/_ref/libreak_ref.so, and
  (b) tests/golden/proximity/proximity.npz, outputs of that library committed with their generator.
GPU (`-m gpu`): the same comparisons through the C-ABI on the device.

Tolerance: distances and points are O(1) lengths in metres; 1e-10 absolute (the finders are a few dozen
double operations after the chain's forward kinematics).  ccylinder-box runs a golden-section search whose
comparisons can flip on rounding differences: its distance is only defined to the search tolerance, and
gets 1e-6.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import random_batch
from reak_b200 import _abi, kte, presets
from reak_b200 import proximity as px

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "proximity", "proximity.npz")
KINDS = {"plane": px.plane, "sphere": px.sphere, "ccylinder": px.capped_cylinder, "cylinder": px.cylinder, "box": px.box}
TOL = 1e-10
TOL_SEARCH = 1e-6


def random_pose(rng, spread=1.2):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return px.pose_3D(rng.uniform(-spread, spread, size=3), q)


def random_shape(rng, kind, anchor=None, spread=1.2):
    pose = random_pose(rng, spread)
    if kind == "plane":
        return px.plane("pl", anchor, pose, rng.uniform(0.5, 3.0, size=2))
    if kind == "sphere":
        return px.sphere("sp", anchor, pose, rng.uniform(0.05, 0.6))
    if kind == "ccylinder":
        return px.capped_cylinder("cc", anchor, pose, rng.uniform(0.1, 1.5), rng.uniform(0.03, 0.4))
    if kind == "cylinder":
        return px.cylinder("cy", anchor, pose, rng.uniform(0.1, 1.5), rng.uniform(0.03, 0.4))
    return px.box("bx", anchor, pose, rng.uniform(0.1, 1.5, size=3))


def has_search(pair):
    kinds = {s.kind for s in pair.model1.mShapeList} | {s.kind for s in pair.model2.mShapeList}
    return _abi.SHAPE_CCYLINDER in kinds and _abi.SHAPE_BOX in kinds


def agree(got, want, tol):
    """(distance, finder, points) triples; a different finder is accepted only on a tie."""
    d, f, p = got
    dr, fr, pr = want
    fin = np.isfinite(dr)
    assert np.array_equal(np.isfinite(d), fin)
    assert np.all(d[~fin] == dr[~fin])
    assert np.max(np.abs(d[fin] - dr[fin]), initial=0.0) < tol
    same = f == fr
    assert np.all(np.abs(d[~same] - dr[~same]) < tol), "different finder without a tie"
    ok = fin & same
    assert np.max(np.abs(p[ok] - pr[ok]), initial=0.0) < max(tol, 1e-9) * (1e3 if tol > TOL else 1.0)


# ---- the host build of the device source --------------------------------------------------------
@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("prox_host") / "libprox_host.so")
    src = os.path.join(HERE, "host_build", "prox_host.cpp")
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.prox_host_min_distance.restype = C.c_int
    lib.prox_host_min_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


class HostProximity(object):
    """chain + proxy lowered by the product library (host code), finders run by the host build."""

    def __init__(self, host_lib, system, pair):
        self.compiled = kte.compile_chain(system.chain, system.mass_calc, system.dofs_gen, system.inputs)
        self.lib = _abi.load_library()
        self.h = C.c_void_p()
        _abi.check(self.lib.rkb_chain_create(C.byref(self.compiled.desc), C.byref(self.h)), "rkb_chain_create")
        self.proxy = px.ProxyHandle(self.lib, self.h, pair, self.compiled.frames)
        size = host_lib.prox_host_program_size()
        self.blob = C.create_string_buffer(size)
        assert self.lib.rkb_proxy_program(self.proxy._h, self.blob, size) == size
        self.host = host_lib

    def min_distance(self, frames):
        """frames [N][n_frames][>=7] world poses -> (distance, finder, points)."""
        N = frames.shape[0]
        d, f, p = np.zeros(N), np.zeros(N, dtype=np.int32), np.zeros((N, 6))
        fn = self.host.prox2d_host_min_distance if self.compiled.dim == 2 else self.host.prox_host_min_distance
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for i in range(N):
            fr = np.ascontiguousarray(frames[i][:, :7])
            di = C.c_double()
            f[i] = fn(self.blob, fr.ctypes.data_as(C.c_void_p), fr.shape[0], C.byref(di), p[i].ctypes.data_as(C.c_void_p))
            d[i] = di.value
        return d, f, p

    def gather(self, frames, max_records):
        """frames [N][n_frames][>=7] -> (count [N], finder [N][M], records [N][M][7]) of gatherCollisionPoints"""
        N = frames.shape[0]
        cnt = np.zeros(N, dtype=np.int32)
        fnd = np.full((N, max_records), -1, dtype=np.int32)
        rec = np.zeros((N, max_records, 7))
        rec[:, :, 0] = np.inf
        fn = self.host.prox2d_host_gather if self.compiled.dim == 2 else self.host.prox_host_gather
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for i in range(N):
            fr = np.ascontiguousarray(frames[i][:, :7])
            cnt[i] = fn(self.blob, fr.ctypes.data_as(C.c_void_p), fr.shape[0], max_records, rec[i].ctypes.data_as(C.c_void_p),
                        fnd[i].ctypes.data_as(C.c_void_p))
        return cnt, fnd, rec

    def close(self):
        self.proxy.close()
        self.lib.rkb_chain_destroy(self.h)


def ref_frames(R, x):
    return np.stack([R.frames(x[i:i + 1]) for i in range(x.shape[0])])


def need_ref(oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("k2", sorted(KINDS))
@pytest.mark.parametrize("k1", sorted(KINDS))
def test_single_finder_host_vs_reference(k1, k2, host_lib, oracle_built):
    """every (kind, kind) combination, both model orders: one shape fixed in the world, one on the arm"""
    need_ref(oracle_built)
    rng = np.random.default_rng(sum(ord(c) for c in k1) * 131 + sum(ord(c) for c in k2))
    s = presets.make("crs6")
    for trial in range(24):
        a = random_shape(rng, k1, s.joint_end_frames[trial % 6] if trial % 2 else None, spread=0.8 if trial % 2 else 1.2)
        b = random_shape(rng, k2, None)
        if trial % 2 == 0:
            b.pose.position = tuple(np.array(b.pose.position) + np.array([0.0, -3.3, 0.8]))  # near the arm
        pair = px.proxy_query_pair_3D("t", px.proxy_query_model_3D("a").addShape(a), px.proxy_query_model_3D("b").addShape(b))
        H = HostProximity(host_lib, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 4, seed=trial, q_range=3.0)
        want = R.min_distance(pair, x)
        got = H.min_distance(ref_frames(R, x))
        if not pair.finder_pairs():
            assert np.all(np.isinf(got[0])) and np.all(got[1] == -1) and np.all(np.isinf(want[0]))
        else:
            agree(got, want, TOL_SEARCH if has_search(pair) else TOL)
        H.close()


def special_pairs():
    """configurations that land in the finders' special branches"""
    I = px.pose_3D()
    out = []
    # parallel capsules (the always-true overlap branch), overlapping and far apart along the axis
    out.append(("cc_parallel", px.capped_cylinder("a", None, px.pose_3D((0, 0, 0)), 1.0, 0.1), px.capped_cylinder("b", None, px.pose_3D((0.5, 0.2, 0.3)), 0.6, 0.2)))
    out.append(("cc_parallel_far", px.capped_cylinder("a", None, px.pose_3D((0, 0, 0)), 1.0, 0.1), px.capped_cylinder("b", None, px.pose_3D((0.5, 0.2, 3.0)), 0.6, 0.2)))
    # capsule and cylinder lying flat on a plane, cylinder standing on its end
    flat = px.pose_3D.axis_angle(np.pi / 2, (1, 0, 0), (0.2, 0.1, 0.5))
    out.append(("plane_cc_flat", px.plane("p", None, I, (2, 2)), px.capped_cylinder("b", None, flat, 0.8, 0.1)))
    out.append(("plane_cy_flat", px.plane("p", None, I, (2, 2)), px.cylinder("b", None, flat, 0.8, 0.1)))
    out.append(("plane_cy_end", px.plane("p", None, I, (2, 2)), px.cylinder("b", None, px.pose_3D((0.2, 0.1, 0.9)), 0.8, 0.1)))
    out.append(("plane_cc_below", px.plane("p", None, I, (2, 2)), px.capped_cylinder("b", None, px.pose_3D.axis_angle(0.4, (0, 1, 0), (0.2, 0.1, -0.2)), 0.8, 0.1)))
    # sphere inside a box (each nearest face), outside a corner, above a cylinder's end, beside its rim
    for k, c in enumerate([(0.4, 0.0, 0.0), (0.0, -0.4, 0.1), (0.1, 0.0, 0.2), (1.0, 1.0, 1.0)]):
        out.append(("sphere_box_%d" % k, px.sphere("s", None, px.pose_3D(c), 0.05), px.box("b", None, I, (1.0, 1.0, 0.5))))
    out.append(("sphere_cy_top", px.sphere("s", None, px.pose_3D((0.05, 0.02, 0.9)), 0.1), px.cylinder("c", None, I, 1.0, 0.3)))
    out.append(("sphere_cy_rim", px.sphere("s", None, px.pose_3D((0.6, 0.2, 0.9)), 0.1), px.cylinder("c", None, I, 1.0, 0.3)))
    out.append(("sphere_cy_side", px.sphere("s", None, px.pose_3D((0.6, 0.2, 0.1)), 0.1), px.cylinder("c", None, I, 1.0, 0.3)))
    out.append(("sphere_cc_cap", px.sphere("s", None, px.pose_3D((0.2, 0.1, -1.2)), 0.1), px.capped_cylinder("c", None, I, 1.0, 0.3)))
    # penetrations
    out.append(("sphere_sphere_in", px.sphere("s", None, px.pose_3D((0.1, 0, 0)), 0.3), px.sphere("t", None, px.pose_3D((0.3, 0.1, 0)), 0.2)))
    out.append(("cc_box_in", px.capped_cylinder("c", None, px.pose_3D.axis_angle(0.3, (1, 1, 0), (0.1, 0.1, 0.1)), 0.5, 0.05), px.box("b", None, I, (1.0, 1.0, 1.0))))
    out.append(("plane_box", px.plane("p", None, I, (2, 2)), px.box("b", None, px.pose_3D.axis_angle(0.7, (1, 2, 3), (0.1, 0.2, 0.9)), (0.3, 0.5, 0.7))))
    out.append(("plane_plane", px.plane("p", None, I, (2, 2)), px.plane("q", None, px.pose_3D.axis_angle(0.7, (1, 2, 3), (0.1, 0.2, 0.9)), (1.0, 3.0))))
    out.append(("plane_plane_side", px.plane("p", None, I, (2, 2)), px.plane("q", None, px.pose_3D.axis_angle(1.2, (1, 0, 0), (3.0, 0.2, 0.4)), (1.0, 1.0))))
    return out


@pytest.mark.parametrize("case", special_pairs(), ids=[c[0] for c in special_pairs()])
def test_special_branches_host_vs_reference(case, host_lib, oracle_built):
    need_ref(oracle_built)
    _, a, b = case
    s = presets.make("crs6")
    for first, second in ((a, b), (b, a)):
        pair = px.proxy_query_pair_3D("t", px.proxy_query_model_3D("a").addShape(first), px.proxy_query_model_3D("b").addShape(second))
        H = HostProximity(host_lib, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 1, seed=1)
        agree(H.min_distance(ref_frames(R, x)), R.min_distance(pair, x), TOL_SEARCH if has_search(pair) else TOL)
        H.close()


def mixed_models(s, rng, n1=6, n2=7):
    kinds = sorted(KINDS)
    m1, m2 = px.proxy_query_model_3D("robot"), px.proxy_query_model_3D("world")
    for k in range(n1):
        m1.addShape(random_shape(rng, kinds[k % 5], s.joint_end_frames[k % len(s.joint_end_frames)], spread=0.3))
    for k in range(n2):
        sh = random_shape(rng, kinds[(k + 2) % 5], None, spread=1.0)
        sh.pose.position = tuple(np.array(sh.pose.position) + np.array([0.0, -3.3, 0.8]))
        m2.addShape(sh)
    return px.proxy_query_pair_3D("mixed", m1, m2)


def test_crs_lab_host_vs_reference(host_lib, oracle_built):
    """the CRS arm's proximity model against the MD148 lab, shapes riding on the joint end frames"""
    need_ref(oracle_built)
    s = presets.make("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    assert len(pair.finder_pairs()) == 25
    H = HostProximity(host_lib, s, pair)
    assert H.proxy.n_finders == 25 and H.proxy.finder(7) == pair.finder_pairs()[7]
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 256, seed=11, q_range=3.1)
    want = R.min_distance(pair, x)
    agree(H.min_distance(ref_frames(R, x)), want, TOL)
    assert (want[0] < 0).any() and (want[0] > 0).any()  # both colliding and free states in the sample
    assert len(set(want[1].tolist())) >= 3                # and more than one finder wins
    H.close()


def test_mixed_models_host_vs_reference(host_lib, oracle_built):
    """every shape kind on both sides, including pairs without a finder and the culling test"""
    need_ref(oracle_built)
    s = presets.make("crs7")
    for seed in range(4):
        pair = mixed_models(s, np.random.default_rng(100 + seed))
        H = HostProximity(host_lib, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 64, seed=seed, q_range=2.0)
        agree(H.min_distance(ref_frames(R, x)), R.min_distance(pair, x), TOL_SEARCH)
        H.close()


def golden_pair(s, g):
    m = [px.proxy_query_model_3D("m1"), px.proxy_query_model_3D("m2")]
    ctor = {1: "plane", 2: "sphere", 3: "ccylinder", 4: "cylinder", 5: "box"}
    for which, key in enumerate(("shapes1", "shapes2")):
        for row in g[key]:
            kind, anchor = int(row[0]), int(row[1])
            sh = px.shape_3D("g", None if anchor < 0 else anchor, px.pose_3D(row[2:5], row[5:9]), row[9:12])
            sh.kind = kind
            assert ctor[kind]
            m[which].addShape(sh)
    return px.proxy_query_pair_3D("golden", m[0], m[1])


def test_host_vs_golden(host_lib):
    """committed outputs of the reference (tests/golden/proximity/make_golden_proximity.py); runs without oracle/_ref"""
    g = np.load(GOLDEN)
    for tag, preset in (("crs_lab", "crs6"), ("mixed", "crs7")):
        s = presets.make(preset)
        sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
        pair = golden_pair(s, sub)
        H = HostProximity(host_lib, s, pair)
        got = H.min_distance(sub["frames"])
        agree(got, (sub["distance"], sub["finder"], sub["points"]), TOL_SEARCH if tag == "mixed" else TOL)
        H.close()


def test_reference_side_binding_round_trip(oracle_built):
    """reak_bridge.hpp's compile_proxy_model, run on live geom:: shapes riding on live ReaK frames, hands back
    the shape list those shapes were built from"""
    need_ref(oracle_built)
    s = presets.make("crs7")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    R = oracle_built.Reference(c)
    robot, lab = presets.crs_proxy_models(s, track=True)
    for model in (robot, lab, mixed_models(s, np.random.default_rng(5)).model1):
        src, out, anchors = R.bridge_proxy(model)
        for k in range(len(model.mShapeList)):
            assert out[k].kind == src[k].kind and anchors[k] == src[k].anchor
            assert np.allclose(list(out[k].position), list(src[k].position), rtol=0, atol=0)
            assert np.allclose(list(out[k].quat), list(src[k].quat), rtol=0, atol=1e-15)
            assert np.allclose(list(out[k].dims), list(src[k].dims), rtol=0, atol=0)


def test_reference_side_binding_round_trip_planar(oracle_built):
    """the planar overload of reak_bridge.hpp's compile_proxy_model on live geom::circle / capped_rectangle / rectangle
    riding on live frame_2D's"""
    need_ref(oracle_built)
    s = presets.make("crs2d")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    R = oracle_built.Reference(c)
    pair = mixed_models2(c, np.random.default_rng(5))
    for model in (pair.model1, pair.model2):
        src, out, anchors = R.bridge_proxy(model)
        for k in range(len(model.mShapeList)):
            assert out[k].kind == src[k].kind and anchors[k] == src[k].anchor
            assert np.allclose(list(out[k].position)[:2], list(src[k].position)[:2], rtol=0, atol=0)
            assert np.allclose(list(out[k].quat)[:2], list(src[k].quat)[:2], rtol=0, atol=1e-15)
            assert np.allclose(list(out[k].dims), list(src[k].dims), rtol=0, atol=0)


@pytest.mark.parametrize("preset,track", [("crs6", False), ("crs7", True)])
def test_checked_steering_loop_matches_meaqr_steer_with_constant_control_itself(preset, track, oracle_built):
    """The checked steering loop the GPU path is held against (rkref_steer_feedback_checked: restated loop, the reference's
    proximity query) against MEAQR_topology::steer_with_constant_control ITSELF with with_collision_check = true — the
    unmodified class over the live kte_nl_system (oracle/ref_steer_law.cpp), its virtual is_free_impl answered by the same
    proxy pairs (findMinimumDistance >= 0 after doMotion, MEAQR_topology.hpp:921-940).  End state, last accepted input,
    number of accepted intervals and the collision flag."""
    need_ref(oracle_built)
    s = presets.make(preset)
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    R = oracle_built.Reference(c)
    robot, lab = presets.crs_proxy_models(s, track=track)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    nu, nx = c.n_inputs, 2 * c.n_coords
    n, J, T = 40, 6, 0.05
    rng = np.random.default_rng(21)
    x0 = np.zeros((n, nx))
    x0[:, 0::2] = rng.uniform(-2.5, 2.5, (n, c.n_coords))
    x0[:, 1::2] = rng.uniform(-2.0, 2.0, (n, c.n_coords))
    x0 = x0[R.min_distance(pair, x0)[0] >= 0.0][:16]            # free start states
    n = x0.shape[0]
    goal = x0 + rng.uniform(-1.0, 1.0, (n, nx))
    u_bias = rng.uniform(-1.0, 1.0, (n, nu))
    gain = rng.uniform(-2.0, 2.0, (n, nu, nx))
    u_prev = rng.uniform(-0.5, 0.5, (n, nu))
    lo, hi, bw = -2.0 * np.ones(nu), 2.0 * np.ones(nu), 60.0 * np.ones(nu)
    try:
        xw, uw, tw, cw = R.meaqr_steer(x0, goal, u_bias, gain, u_prev, T, (J - 0.5) * T, 0.25, (lo, hi), bw,
                                       is_free=lambda x: R.min_distance(pair, x[None, :])[0][0] >= 0.0)
    except NotImplementedError:
        pytest.skip("prebuilt libreak_ref.so without the MEAQR unit")
    xg, ug, nd, _, st, cg = R.steer_feedback(x0, goal, u_bias, gain, u_prev, T, T * 1e-1, 10, J, 0.25, saturate_first=False,
                                             bounds=(lo, hi), rate_bounds=(-bw, bw), proxy_pairs=[pair])
    assert np.array_equal(cg, cw) and (cw == 1).any() and (cw == 0).any()
    assert np.array_equal(nd, np.rint(tw / T).astype(np.int32)) and ((cw == 1) & (nd > 0)).any()
    assert np.max(np.abs(xg - xw)) < 1e-12 and np.max(np.abs(ug - uw)) < 1e-12


def test_proxy_create_rejects():
    lib = _abi.load_library()
    s = presets.make("crs6")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    h = C.c_void_p()
    _abi.check(lib.rkb_chain_create(C.byref(c.desc), C.byref(h)), "rkb_chain_create")

    def create(shape, n=1):
        arr = (_abi.rkb_shape * max(n, 1))()
        for k in range(n):
            arr[k] = shape.to_c(c.frames)
        out = C.c_void_p()
        rc = lib.rkb_proxy_create(h, arr, n, arr, n, C.byref(out))
        if rc == 0:
            lib.rkb_proxy_destroy(out)
        return rc

    assert create(px.sphere("s", None, None, 0.1)) == 0
    assert create(px.sphere("s", None, None, -0.1)) == _abi.ERR_INVALID             # non-positive size
    assert create(px.sphere("s", 999, None, 0.1)) == _abi.ERR_INVALID               # anchor out of range
    assert create(px.sphere("s", None, px.pose_3D(quat=(2, 0, 0, 0)), 0.1)) == _abi.ERR_INVALID  # not a rotation
    bad = px.sphere("s", None, None, 0.1)
    bad.kind = 9
    assert create(bad) == _abi.ERR_INVALID
    assert create(px.sphere("s", None, None, 0.1), n=_abi.PROXY_MAX_SHAPES + 1) == _abi.ERR_UNSUPPORTED
    lib.rkb_chain_destroy(h)
    # planar chains have no 3D frames to anchor to
    s2 = presets.make("planar2")
    c2 = kte.compile_chain(s2.chain, s2.mass_calc, s2.dofs_gen, s2.inputs)
    h2 = C.c_void_p()
    _abi.check(lib.rkb_chain_create(C.byref(c2.desc), C.byref(h2)), "rkb_chain_create")
    arr = (_abi.rkb_shape * 1)()
    arr[0] = px.sphere("s", None, None, 0.1).to_c()
    out = C.c_void_p()
    assert lib.rkb_proxy_create(h2, arr, 1, arr, 1, C.byref(out)) == _abi.ERR_UNSUPPORTED
    lib.rkb_chain_destroy(h2)


# ---- planar models (proxy_query_pair_2D): circle, capped rectangle, rectangle on planar chains -------------
KINDS2 = {"circle": px.circle, "crect": px.capped_rectangle, "rectangle": px.rectangle}


def agree2(got, want, tol, pair=None):
    """as agree(); crossing centre lines make the reference divide 0 by 0 for the two points (prox_crect_crect.cpp:121-124,
    prox_crect_rectangle.cpp:198-203): NaN there is the expected answer, in the same places"""
    (d, f, p), (dr, fr, pr) = got, want
    p, pr = p.copy(), pr.copy()
    if pair is not None:
        # ... and centre lines that cross up to rounding (distance between them ~1e-17) give two points in a direction made
        # of rounding noise, in the reference as here: only the distance is defined for those
        n2 = len(pair.model2.mShapeList)
        for i in range(len(dr)):
            if fr[i] < 0 or f[i] != fr[i]:
                continue
            sa, sb = pair.model1.mShapeList[fr[i] // n2], pair.model2.mShapeList[fr[i] % n2]
            if sa.kind == _abi.SHAPE_CRECT and sb.kind == _abi.SHAPE_CRECT and abs(dr[i] + 0.5 * sa.dims[1] + 0.5 * sb.dims[1]) < 1e-9:
                p[i] = pr[i] = 0.0
    assert np.array_equal(np.isnan(p), np.isnan(pr)) and np.array_equal(np.isnan(d), np.isnan(dr))
    agree((d, f, np.nan_to_num(p, nan=0.0)), (dr, fr, np.nan_to_num(pr, nan=0.0)), tol)


def random_shape2(rng, kind, anchor=None, spread=1.0, angle=None):
    pose = px.pose_2D(rng.uniform(-spread, spread, size=2), rng.uniform(-np.pi, np.pi) if angle is None else angle)
    if kind == "circle":
        return px.circle("ci", anchor, pose, rng.uniform(0.05, 0.5))
    if kind == "crect":
        return px.capped_rectangle("cr", anchor, pose, (rng.uniform(0.2, 1.2), rng.uniform(0.05, 0.4)))
    return px.rectangle("re", anchor, pose, rng.uniform(0.2, 1.2, size=2))


def mixed_models2(compiled, rng, n1=5, n2=6):
    kinds = sorted(KINDS2)
    m1, m2 = px.proxy_query_model_2D("robot"), px.proxy_query_model_2D("world")
    nf = compiled.desc.n_frames
    for k in range(n1):
        m1.addShape(random_shape2(rng, kinds[k % 3], int(rng.integers(1, nf)), spread=0.3))
    for k in range(n2):
        m2.addShape(random_shape2(rng, kinds[(k + 1) % 3], None, spread=1.5))
    return px.proxy_query_pair_2D("mixed2", m1, m2)


@pytest.mark.parametrize("k2", sorted(KINDS2))
@pytest.mark.parametrize("k1", sorted(KINDS2))
def test_planar_single_finder_host_vs_reference(k1, k2, host_lib, oracle_built):
    """every (kind, kind) combination of planar shapes, both model orders: one fixed in the world, one on the arm"""
    need_ref(oracle_built)
    rng = np.random.default_rng(sum(ord(c) for c in k1) * 17 + sum(ord(c) for c in k2))
    s = presets.make("planar3_sd")
    for trial in range(24):
        probe = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        a = random_shape2(rng, k1, (1 + trial % (probe.desc.n_frames - 1)) if trial % 2 else None, spread=0.6 if trial % 2 else 1.2)
        b = random_shape2(rng, k2, None)
        pair = px.proxy_query_pair_2D("t", px.proxy_query_model_2D("a").addShape(a), px.proxy_query_model_2D("b").addShape(b))
        H = HostProximity(host_lib, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 4, seed=trial, q_range=3.0)
        agree2(H.min_distance(ref_frames(R, x)), R.min_distance(pair, x), TOL)
        H.close()


def planar_special_pairs():
    """axis-parallel configurations (the 1e-5 tests on the tangent) and containment"""
    out = []
    P = px.pose_2D
    out.append(("crect_parallel", px.capped_rectangle("a", None, P((0, 0), 0.0), (1.0, 0.2)), px.capped_rectangle("b", None, P((0.3, 0.5), 0.0), (0.6, 0.3))))
    out.append(("crect_parallel_far", px.capped_rectangle("a", None, P((0, 0), 0.0), (1.0, 0.2)), px.capped_rectangle("b", None, P((3.0, 0.5), np.pi), (0.6, 0.3))))
    out.append(("crect_rect_vertical", px.capped_rectangle("a", None, P((0.9, 0.1), np.pi / 2), (1.0, 0.2)), px.rectangle("b", None, P((0, 0), 0.0), (1.0, 0.8))))
    out.append(("crect_rect_horizontal", px.capped_rectangle("a", None, P((0.1, -0.9), 0.0), (1.0, 0.2)), px.rectangle("b", None, P((0, 0), 0.0), (1.0, 0.8))))
    out.append(("crect_rect_end_in_slab", px.capped_rectangle("a", None, P((1.2, 0.1), 0.3), (0.8, 0.1)), px.rectangle("b", None, P((0, 0), 0.0), (1.0, 0.8))))
    out.append(("crect_rect_crossing", px.capped_rectangle("a", None, P((0.1, 0.1), 0.7), (2.0, 0.1)), px.rectangle("b", None, P((0, 0), 0.2), (1.0, 0.8))))
    out.append(("circle_in_rect", px.circle("a", None, P((0.1, 0.25)), 0.05), px.rectangle("b", None, P((0, 0), 0.4), (1.0, 0.8))))
    out.append(("circle_in_crect", px.circle("a", None, P((0.1, 0.02)), 0.05), px.capped_rectangle("b", None, P((0, 0), 0.4), (1.0, 0.3))))
    out.append(("circle_crect_cap", px.circle("a", None, P((0.9, 0.2)), 0.1), px.capped_rectangle("b", None, P((0, 0), 0.0), (1.0, 0.3))))
    out.append(("rect_rect_overlap", px.rectangle("a", None, P((0.2, 0.1), 0.3), (1.0, 0.6)), px.rectangle("b", None, P((0, 0), 0.0), (1.0, 0.8))))
    out.append(("circle_circle_in", px.circle("a", None, P((0.1, 0.0)), 0.3), px.circle("b", None, P((0.3, 0.1)), 0.2)))
    return out


@pytest.mark.parametrize("case", planar_special_pairs(), ids=[c[0] for c in planar_special_pairs()])
def test_planar_special_branches_host_vs_reference(case, host_lib, oracle_built):
    need_ref(oracle_built)
    _, a, b = case
    s = presets.make("planar2")
    for first, second in ((a, b), (b, a)):
        pair = px.proxy_query_pair_2D("t", px.proxy_query_model_2D("a").addShape(first), px.proxy_query_model_2D("b").addShape(second))
        H = HostProximity(host_lib, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 1, seed=1)
        agree2(H.min_distance(ref_frames(R, x)), R.min_distance(pair, x), TOL)
        H.close()


@pytest.mark.parametrize("preset", ["planar3_sd", "crs2d", "planar_pr"])
def test_planar_mixed_models_host_vs_reference(preset, host_lib, oracle_built):
    """several shapes per model on revolute / prismatic planar chains: the culling test, gatherCollisionPoints"""
    need_ref(oracle_built)
    s = presets.make(preset)
    for seed in range(3):
        probe = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        pair = mixed_models2(probe, np.random.default_rng(40 + seed))
        H = HostProximity(host_lib, s, pair)
        assert H.proxy.n_finders == 30
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 64, seed=seed, q_range=2.5)
        fr = ref_frames(R, x)
        want = R.min_distance(pair, x)
        agree2(H.min_distance(fr), want, TOL, pair)
        assert (want[0] < 0).any() and (want[0] > 0).any()
        cnt, fnd, rec = H.gather(fr, 30)
        cr, recr = R.collision_points(pair, x, 30)
        assert np.array_equal(cnt, cr) and cnt.max() > 1
        # (distances of the records; their points are the finders' own, checked above, and undefined for crossing centre lines)
        dg, dw = rec[:, :, 0], recr[:, :, 0]
        assert np.array_equal(np.isfinite(dg), np.isfinite(dw)) and np.max(np.abs(dg[np.isfinite(dw)] - dw[np.isfinite(dw)])) < 1e-9
        H.close()


GOLDEN2 = os.path.join(HERE, "golden", "proximity", "proximity_2d.npz")


def golden_pair2(g):
    m = [px.proxy_query_model_2D("m1"), px.proxy_query_model_2D("m2")]
    for which, key in enumerate(("shapes1", "shapes2")):
        for row in g[key]:
            kind, anchor = int(row[0]), int(row[1])
            sh = px.shape_2D("g", None if anchor < 0 else anchor, px.pose_2D(row[2:4], np.arctan2(row[6], row[5])), row[9:12])
            sh.kind = kind
            sh.pose = px.pose_3D(tuple(row[2:5]), tuple(row[5:9]))   # the stored (cos, sin), bit for bit
            m[which].addShape(sh)
    return px.proxy_query_pair_2D("golden2", m[0], m[1])


def test_planar_host_vs_golden(host_lib):
    """committed outputs of the reference (tests/golden/proximity/make_golden_proximity_2d.py); runs without oracle/_ref"""
    g = np.load(GOLDEN2)
    for tag, preset in (("arm3", "planar3_sd"), ("crs2d", "crs2d")):
        s = presets.make(preset)
        sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
        pair = golden_pair2(sub)
        H = HostProximity(host_lib, s, pair)
        agree2(H.min_distance(sub["frames"]), (sub["distance"], sub["finder"], sub["points"]), TOL, pair)
        H.close()


def test_planar_proxy_create_rejects():
    """planar shapes need a planar chain and the other way round; a planar rotation must be a unit (cos, sin)"""
    lib = _abi.load_library()
    for preset, shape, want in (("planar2", px.circle("c", None, None, 0.1), 0), ("planar2", px.sphere("s", None, None, 0.1), _abi.ERR_UNSUPPORTED),
                                ("crs6", px.circle("c", None, None, 0.1), _abi.ERR_UNSUPPORTED), ("planar2", px.rectangle("r", None, None, (0.1, -1.0)), _abi.ERR_INVALID)):
        s = presets.make(preset)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        h = C.c_void_p()
        _abi.check(lib.rkb_chain_create(C.byref(c.desc), C.byref(h)), "rkb_chain_create")
        arr = (_abi.rkb_shape * 1)()
        arr[0] = shape.to_c(c.frames)
        out = C.c_void_p()
        rc = lib.rkb_proxy_create(h, arr, 1, arr, 1, C.byref(out))
        assert rc == want, (preset, shape.kind, rc)
        if rc == 0:
            lib.rkb_proxy_destroy(out)
        lib.rkb_chain_destroy(h)


# ---- the generated (run-time specialised) source, compiled for the host -------------------------------
class SpecHost(object):
    """rkb_proxy_source() of a chain + pair — the text rkb_proxy_specialize gives NVRTC — compiled by g++ with
    tests/host_build/prox_spec_host.h and run on the coordinates themselves (its own forward kinematics)."""

    def __init__(self, tmp, system, pair, min_blocks=None):
        self.compiled = kte.compile_chain(system.chain, system.mass_calc, system.dofs_gen, system.inputs, getattr(system, "dofs_3D", ()))
        lib = _abi.load_library()
        h = C.c_void_p()
        _abi.check(lib.rkb_chain_create(C.byref(self.compiled.desc), C.byref(h)), "rkb_chain_create")
        proxy = px.ProxyHandle(lib, h, pair, self.compiled.frames)
        if min_blocks:
            proxy.set_option(px.ProxyHandle.OPT_MIN_BLOCKS, min_blocks)
        self.source = proxy.source()
        proxy.close()
        lib.rkb_chain_destroy(h)
        SpecHost.count = getattr(SpecHost, "count", 0) + 1
        cu, so = os.path.join(tmp, "spec%d.cu" % SpecHost.count), os.path.join(tmp, "libspec%d.so" % SpecHost.count)
        with open(cu, "w") as f:
            f.write(self.source)
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-include", os.path.join(HERE, "host_build", "prox_spec_host.h"),
                        "-I", os.path.join(HERE, "..", "reak_b200", "csrc"), "-x", "c++", cu, "-o", so], check=True)
        self.lib = C.CDLL(so)
        self.lib.prox_spec_host.restype = C.c_int
        self.lib.prox_spec_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]

    def min_distance(self, x, with_points=True):
        n = self.compiled.n_coords
        N = x.shape[0]
        d, f, p = np.zeros(N), np.zeros(N, dtype=np.int32), np.zeros((N, 6))
        for i in range(N):
            q = np.ascontiguousarray(x[i, 0:2 * n:2])
            fc = np.ascontiguousarray(x[i, 2 * n:2 * n + 7]) if x.shape[1] > 2 * n else np.zeros(7)
            di = C.c_double()
            f[i] = self.lib.prox_spec_host(q.ctypes.data_as(C.c_void_p), fc.ctypes.data_as(C.c_void_p), int(with_points), C.byref(di),
                                           p[i].ctypes.data_as(C.c_void_p))
            d[i] = di.value
        return d, f, p


@pytest.fixture(scope="module")
def spec_tmp(tmp_path_factory):
    return str(tmp_path_factory.mktemp("prox_spec"))


@pytest.mark.parametrize("preset,track", [("crs6", False), ("crs7", True), ("crs6_twist", False), ("crs7_phys_sd", True), ("crs6_sd_sat", False)])
def test_generated_source_crs_lab_vs_reference(preset, track, spec_tmp, oracle_built):
    """the straight-line kernel source of the CRS arm against the MD148 lab: forward kinematics with axis-aligned joints
    and links, world-fixed shapes as literals; crs6_twist has rotated links (the general quaternion product)"""
    need_ref(oracle_built)
    s = presets.make(preset)
    robot, lab = presets.crs_proxy_models(s, track=track)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    H = SpecHost(spec_tmp, s, pair)
    assert "rkb_prox_spec" not in H.source.split("RKB_PROX_SPEC_KERNELS")[0] and "RKB_PROX_SPEC_KERNELS(%d, 0, " % H.compiled.n_coords in H.source
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 256, seed=11, q_range=3.1)
    want = R.min_distance(pair, x)
    agree(H.min_distance(x), want, TOL)
    d0, f0, _ = H.min_distance(x, with_points=False)   # the distance-only kernel's path
    assert np.max(np.abs(d0 - want[0])) < TOL and np.all((f0 == want[1]) | (np.abs(d0 - want[0]) < TOL))
    assert (want[0] < 0).any() and (want[0] > 0).any()


def test_generated_source_mixed_models_vs_reference(spec_tmp, oracle_built):
    """every shape kind on both sides, rotated local poses, pairs without a finder, the culling test"""
    need_ref(oracle_built)
    s = presets.make("crs7")
    for seed in range(4):
        pair = mixed_models(s, np.random.default_rng(100 + seed))
        H = SpecHost(spec_tmp, s, pair)
        R = oracle_built.Reference(H.compiled)
        x, _ = random_batch(H.compiled, 64, seed=seed, q_range=2.0)
        agree(H.min_distance(x), R.min_distance(pair, x), TOL_SEARCH)


def test_generated_source_general_axes_and_moving_second_model(spec_tmp, oracle_built):
    """joint axes that are not unit vectors of the frame (physical CRS), shapes of BOTH models riding on the chain"""
    need_ref(oracle_built)
    s = presets.make("crs6_phys")
    rng = np.random.default_rng(7)
    m1, m2 = px.proxy_query_model_3D("upper"), px.proxy_query_model_3D("lower")
    kinds = sorted(KINDS)
    for k in range(4):
        m1.addShape(random_shape(rng, kinds[(k + 1) % 5], s.joint_end_frames[3 + k % 3], spread=0.3))
        m2.addShape(random_shape(rng, kinds[(k + 3) % 5], s.joint_end_frames[k % 2] if k < 3 else None, spread=0.4))
    pair = px.proxy_query_pair_3D("self", m1, m2)
    H = SpecHost(spec_tmp, s, pair)
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 96, seed=3, q_range=3.0)
    agree(H.min_distance(x), R.min_distance(pair, x), TOL_SEARCH)


def _spatial_presets():
    out = []
    for n in sorted(presets.PRESETS):
        s = presets.make(n)
        if kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs).dim == 3:
            out.append(n)
    return out


ALL_3D = _spatial_presets()


@pytest.mark.parametrize("preset", ALL_3D)
def test_generated_source_every_spatial_preset_vs_reference(preset, spec_tmp, oracle_built):
    """the generator on every 3D preset chain (springs, dampers, rotors, generalized-coordinate elements, rotated links,
    prismatic tracks, physical axes): shapes of every kind on random frames of the chain against world-fixed ones — the
    text compiles, and its forward kinematics and finder sequence give the reference's distances"""
    need_ref(oracle_built)
    s = presets.make(preset)
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    rng = np.random.default_rng(sum(ord(ch) for ch in preset))
    kinds = sorted(KINDS)
    m1, m2 = px.proxy_query_model_3D("robot"), px.proxy_query_model_3D("world")
    for k in range(5):
        m1.addShape(random_shape(rng, kinds[k % 5], int(rng.integers(0, c.desc.n_frames)), spread=0.3))
        m2.addShape(random_shape(rng, kinds[(k + 2) % 5], None, spread=1.0))
    pair = px.proxy_query_pair_3D("any", m1, m2)
    H = SpecHost(spec_tmp, s, pair)
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 48, seed=5, q_range=2.5)
    agree(H.min_distance(x), R.min_distance(pair, x), TOL_SEARCH)


@pytest.mark.parametrize("preset", ["free_arm3", "free_arm2_twist", "free_body"])
def test_generated_source_free_base_vs_reference(preset, spec_tmp, oracle_built):
    """a free_joint_3D in front of the arm: the generated kinematics read the joint's pose states (position; quaternion,
    normalised as apply_states_and_inputs does) — against the reference's doMotion + findMinimumDistance"""
    need_ref(oracle_built)
    s = presets.make(preset)
    probe = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, s.dofs_3D)
    rng = np.random.default_rng(11)
    kinds = sorted(KINDS)
    m1, m2 = px.proxy_query_model_3D("robot"), px.proxy_query_model_3D("world")
    for k in range(4):
        m1.addShape(random_shape(rng, kinds[k % 5], int(rng.integers(1, probe.desc.n_frames)), spread=0.3))
        m2.addShape(random_shape(rng, kinds[(k + 2) % 5], None, spread=1.0))
    pair = px.proxy_query_pair_3D("free", m1, m2)
    H = SpecHost(spec_tmp, s, pair)
    assert "RKB_PROX_SPEC_KERNELS(%d, 1, " % probe.n_coords in H.source and "freec.q" in H.source
    R = oracle_built.Reference(H.compiled)
    n, nq = 64, probe.n_coords
    x = np.zeros((n, 2 * nq + 13))
    x[:, :2 * nq] = rng.uniform(-2.0, 2.0, (n, 2 * nq))
    x[:, 2 * nq:2 * nq + 3] = rng.uniform(-0.6, 0.6, (n, 3))
    quat = rng.normal(size=(n, 4))
    x[:, 2 * nq + 3:2 * nq + 7] = quat / np.linalg.norm(quat, axis=1, keepdims=True) * rng.uniform(0.7, 1.3, (n, 1))   # not normalised
    x[:, 2 * nq + 7:] = rng.uniform(-1.0, 1.0, (n, 6))
    want = R.min_distance(pair, x)
    agree(H.min_distance(x), want, TOL_SEARCH)
    assert np.ptp(want[0]) > 0.1 and len(set(want[1].tolist())) > 1   # the states matter, and more than one finder wins


def test_generated_source_vs_golden(spec_tmp):
    """committed outputs of the reference; runs without oracle/_ref (the golden file stores the states)"""
    g = np.load(GOLDEN)
    for tag, preset in (("crs_lab", "crs6"), ("mixed", "crs7")):
        if tag + "_x" not in g.files:
            pytest.skip("golden file without states")
        s = presets.make(preset)
        sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
        H = SpecHost(spec_tmp, s, golden_pair(s, sub))
        agree(H.min_distance(sub["x"]), (sub["distance"], sub["finder"], sub["points"]), TOL_SEARCH if tag == "mixed" else TOL)


def test_generated_source_is_keyed_by_its_constants(spec_tmp):
    """two pairs that differ in one dimension give different texts (the cubin cache is keyed by a hash of the text);
    the same pair gives the same text twice; min_blocks lands in the kernel's launch bounds"""
    s = presets.make("crs6")
    robot, lab = presets.crs_proxy_models(s)
    a = SpecHost(spec_tmp, s, px.proxy_query_pair_3D("a", robot, lab))
    b = SpecHost(spec_tmp, s, px.proxy_query_pair_3D("b", robot, lab))
    assert a.source == b.source
    lab.mShapeList[3].dims = tuple(np.array(lab.mShapeList[3].dims) * [1.0, 1.5, 1.0][:len(lab.mShapeList[3].dims)])
    c = SpecHost(spec_tmp, s, px.proxy_query_pair_3D("c", robot, lab), min_blocks=5)
    assert c.source != a.source and "RKB_PROX_SPEC_KERNELS(6, 0, 5)" in c.source


def test_checked_steering_source_builds_for_sm_100a(spec_tmp):
    """the source rkb_steer_checked_specialize hands to NVRTC — kte_serial.cuh's steering kernel with the generated collision
    test of two pairs as its CHK parameter — compiles with nvcc for sm_100a (cross-compilation, no GPU needed)"""
    import shutil
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    from reak_b200.propagator import kte_batch_propagator
    s = presets.make("crs7")
    P = kte_batch_propagator(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    robot, lab = presets.crs_proxy_models(s, track=True)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    extra = px.proxy_query_pair_3D("tool-obstacle", px.proxy_query_model_3D("tool").addShape(px.sphere("tool", s.joint_end_frames[-1], None, 0.12)),
                                   px.proxy_query_model_3D("obstacle").addShape(px.box("crate", None, px.pose_3D((0.3, -3.0, 0.9)), (0.5, 0.5, 0.5))))
    src = P.checked_steering_source([pair, extra])
    expr = src.strip().splitlines()[-1].split("// kernel: ")[1]
    assert expr.startswith("rkb::serial_steer_kernel<7, ") and expr.endswith("RkbSteerCheck>")
    assert "pair0::prox_spec<false>" in src and "pair1::prox_spec<false>" in src
    cu = os.path.join(spec_tmp, "steerchk.cu")
    with open(cu, "w") as f:
        f.write(src + "template __global__ void %s(const __grid_constant__ SerialParams, const __grid_constant__ SteerArgs);\n" % expr)
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-cubin", "-I", os.path.join(HERE, "..", "reak_b200", "csrc"),
                        "-o", os.path.join(spec_tmp, "steerchk.cubin"), cu], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    # an interpreter chain has no steering kernel to build the test into
    s2 = presets.make("crs6_lin_sd")
    P2 = kte_batch_propagator(s2.chain, s2.mass_calc, s2.dofs_gen, s2.inputs)
    if not P2.is_serial():
        robot2, lab2 = presets.crs_proxy_models(s2)
        with pytest.raises(Exception):
            P2.checked_steering_source([px.proxy_query_pair_3D("robot-lab", robot2, lab2)])


# ---- GPU ------------------------------------------------------------------------------------------
def _gpu_prop(preset, auto=False):
    """auto = False: no kernels compiled in the background (they would take over between two calls of a test that compares
    results bit for bit); the tests of the generated kernels ask for them explicitly"""
    from reak_b200.propagator import kte_batch_propagator
    s = presets.make(preset)
    P = kte_batch_propagator(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    if not auto:
        P.set_option("auto_specialize", 0)
    return s, P


@pytest.mark.gpu
def test_gpu_crs_lab_vs_reference(oracle_built):
    need_ref(oracle_built)
    s, P = _gpu_prop("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 4096, seed=5, q_range=3.1)
    want = R.min_distance(pair, x)
    got = P.get_min_distances(pair, x)
    agree(got, want, TOL)
    free = P.is_free([pair], x)
    assert np.array_equal(free, ~(want[0] < 0.0)) and free.any() and (~free).any()


@pytest.mark.gpu
def test_gpu_track_arm_vs_reference(oracle_built):
    need_ref(oracle_built)
    s, P = _gpu_prop("crs7")
    robot, lab = presets.crs_proxy_models(s, track=True)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 2048, seed=6, q_range=2.5)
    agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_mixed_models_vs_reference(seed, oracle_built):
    need_ref(oracle_built)
    s, P = _gpu_prop("crs7")
    pair = mixed_models(s, np.random.default_rng(100 + seed))
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 1024, seed=seed, q_range=2.0)
    agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL_SEARCH)


@pytest.mark.gpu
@pytest.mark.parametrize("case", special_pairs(), ids=[c[0] for c in special_pairs()])
def test_gpu_special_branches_vs_reference(case, oracle_built):
    need_ref(oracle_built)
    _, a, b = case
    s, P = _gpu_prop("crs6")
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 3, seed=1)
    for first, second in ((a, b), (b, a)):
        pair = px.proxy_query_pair_3D("t", px.proxy_query_model_3D("a").addShape(first), px.proxy_query_model_3D("b").addShape(second))
        agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL_SEARCH if has_search(pair) else TOL)


@pytest.mark.gpu
def test_gpu_vs_golden():
    g = np.load(GOLDEN)
    for tag, preset in (("crs_lab", "crs6"), ("mixed", "crs7")):
        s, P = _gpu_prop(preset)
        sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
        pair = golden_pair(s, sub)
        agree(P.get_min_distances(pair, sub["x"]), (sub["distance"], sub["finder"], sub["points"]), TOL_SEARCH if tag == "mixed" else TOL)


def _same(got, want, tol):
    """interpreter kernel against generated kernel: same finder unless a tie, distances and points to rounding"""
    agree(tuple(np.asarray(a) for a in got), tuple(np.asarray(a) for a in want), tol)


@pytest.mark.gpu
@pytest.mark.parametrize("preset,track", [("crs6", False), ("crs7", True), ("crs6_twist", False), ("crs6_phys", False)])
def test_gpu_specialized_crs_lab(preset, track, oracle_built):
    """rkb_proxy_specialize: the query of this chain and pair as generated straight-line CUDA (NVRTC) — against the
    compiled reference and against the interpreter kernel; is_free served by it"""
    need_ref(oracle_built)
    s, P = _gpu_prop(preset)
    robot, lab = presets.crs_proxy_models(s, track=track)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    h = P.proxy_handle(pair)
    h.set_option(h.OPT_AUTO_SPECIALIZE, 0)
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 4096 + 77, seed=5, q_range=3.1)
    want = R.min_distance(pair, x)
    before = P.get_min_distances(pair, x)
    free_before = P.is_free([pair], x)
    assert not h.is_specialized()
    h.specialize()
    assert h.is_specialized()
    d_only, f_only = P.get_min_distances(pair, x, with_points=False)   # the generated kernel
    fin = np.isfinite(want[0])
    assert np.max(np.abs(d_only[fin] - want[0][fin])) < TOL and np.all((f_only == want[1]) | (np.abs(d_only - want[0]) < TOL))
    assert np.max(np.abs(d_only - before[0])) < 1e-12 and np.all((f_only == before[1]) | (np.abs(d_only - before[0]) < 1e-12))
    got = P.get_min_distances(pair, x)                                 # with the two points: the interpreter kernel
    agree(got, want, TOL)
    assert np.array_equal(P.is_free([pair], x), free_before) and np.array_equal(free_before, ~(want[0] < 0.0))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_specialized_mixed_models(seed, oracle_built):
    need_ref(oracle_built)
    s, P = _gpu_prop("crs7")
    pair = mixed_models(s, np.random.default_rng(100 + seed))
    P.proxy_handle(pair).specialize()
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 1024, seed=seed, q_range=2.0)
    agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL_SEARCH)


@pytest.mark.gpu
def test_gpu_specialized_free_base_and_layouts():
    """a free-floating base in front of the arm (the pose states of free_joint_3D feed the kinematics), device-resident
    and structure-of-arrays buffers: generated kernel against the interpreter kernel"""
    import torch
    from reak_b200.propagator import kte_batch_propagator
    s = presets.make("free_arm6")
    P = kte_batch_propagator(s).set_option("auto_specialize", 0)
    rng = np.random.default_rng(9)
    m1, m2 = px.proxy_query_model_3D("arm"), px.proxy_query_model_3D("world")
    kinds = sorted(KINDS)
    frames = [f for f in P.compiled.frames][-4:]
    for k in range(4):
        m1.addShape(random_shape(rng, kinds[k % 5], frames[k % len(frames)], spread=0.3))
        m2.addShape(random_shape(rng, kinds[(k + 2) % 5], None, spread=1.0))
    pair = px.proxy_query_pair_3D("free", m1, m2)
    h = P.proxy_handle(pair)
    h.set_option(h.OPT_AUTO_SPECIALIZE, 0)
    x = rng.uniform(-1.0, 1.0, (3000, P.nx))
    d0, f0 = P.get_min_distances(pair, x, with_points=False)
    dx = torch.from_numpy(x).cuda()
    h.specialize()
    for d, f in (P.get_min_distances(pair, x, with_points=False), tuple(a.cpu().numpy() for a in P.get_min_distances(pair, dx, with_points=False))):
        assert np.max(np.abs(d - d0)) < 1e-6 and np.all((f == f0) | (np.abs(d - d0) < 1e-6))
    assert (d0 < 0).any() and (d0 > 0).any()


@pytest.mark.gpu
def test_gpu_proxy_auto_specialize():
    """default behaviour: once 4096 states have been queried the compilation starts in the background, queries made
    meanwhile run on the interpreter, later ones on the generated kernel — same answers throughout"""
    import time
    s, P = _gpu_prop("crs6", auto=True)
    robot, lab = presets.crs_proxy_models(s)
    lab.mShapeList[0].pose.position = (0.013, 0.0, 0.0)   # a pair no other test has put into the cubin cache
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    h = P.proxy_handle(pair)
    x, _ = random_batch(P.compiled, 8192, seed=2, q_range=3.0)
    small = P.get_min_distances(pair, x[:100], with_points=False)
    assert not h.is_specialized()
    first = P.get_min_distances(pair, x, with_points=False)
    t0 = time.time()
    while not h.is_specialized() and time.time() - t0 < 180:
        time.sleep(0.2)
        P.get_min_distances(pair, x[:500], with_points=False)   # (small queries count towards the 4096 states as well)
    assert h.is_specialized(), "no generated kernel after 180 s"
    for got, was in ((P.get_min_distances(pair, x, with_points=False), first), (P.get_min_distances(pair, x[:100], with_points=False), small)):
        assert np.max(np.abs(got[0] - was[0])) < 1e-12 and np.all((got[1] == was[1]) | (np.abs(got[0] - was[0]) < 1e-12))


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["planar3_sd", "crs2d", "planar_pr", "free_planar2"])
def test_gpu_planar_models_vs_reference(preset, oracle_built):
    """proxy_query_pair_2D on planar chains (revolute, prismatic, free_joint_2D): findMinimumDistance, is_free and
    gatherCollisionPoints through the C-ABI against the live reference"""
    need_ref(oracle_built)
    from reak_b200.propagator import kte_batch_propagator
    s = presets.make(preset)
    P = kte_batch_propagator(s).set_option("auto_specialize", 0)
    R = oracle_built.Reference(P.compiled)
    for seed in range(2):
        pair = mixed_models2(P.compiled, np.random.default_rng(60 + seed))
        x, _ = random_batch(P.compiled, 1500, seed=seed, q_range=2.5)
        if P.nx > x.shape[1]:  # free_joint_2D: position, (cos, sin) not normalised, velocity, angular velocity
            rng = np.random.default_rng(seed)
            ang, scale = rng.uniform(-np.pi, np.pi, 1500), rng.uniform(0.8, 1.2, 1500)
            x = np.hstack([x, rng.uniform(-0.5, 0.5, (1500, 2)), (scale * np.cos(ang))[:, None], (scale * np.sin(ang))[:, None],
                           rng.uniform(-1, 1, (1500, 3))])
        want = R.min_distance(pair, x)
        got = P.get_min_distances(pair, x)
        agree2(got, want, TOL, pair)
        assert (want[0] < 0).any() and (want[0] > 0).any()
        assert np.array_equal(P.is_free([pair], x), ~(want[0] < 0.0))
        cnt, fnd, rec = P.gather_collision_points(pair, x[:300])
        cr, recr = R.collision_points(pair, x[:300], rec.shape[1])
        dg, dw = rec[:, :, 0], recr[:, :, 0]
        assert np.array_equal(cnt, cr) and np.array_equal(np.isfinite(dg), np.isfinite(dw))
        assert np.max(np.abs(dg[np.isfinite(dw)] - dw[np.isfinite(dw)]), initial=0.0) < 1e-9


@pytest.mark.gpu
def test_gpu_planar_vs_golden_and_special_branches(oracle_built):
    g = np.load(GOLDEN2)
    for tag, preset in (("arm3", "planar3_sd"), ("crs2d", "crs2d")):
        s, P = _gpu_prop(preset)
        sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
        pair = golden_pair2(sub)
        agree2(P.get_min_distances(pair, sub["x"]), (sub["distance"], sub["finder"], sub["points"]), TOL, pair)
    if not oracle_built.have_ref():
        return
    s, P = _gpu_prop("planar2")
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 2, seed=1)
    for _, a, b in planar_special_pairs():
        for first, second in ((a, b), (b, a)):
            pair = px.proxy_query_pair_2D("t", px.proxy_query_model_2D("a").addShape(first), px.proxy_query_model_2D("b").addShape(second))
            agree2(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL)


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """2^20 device-resident states: the answer of a sample does not depend on its place in the batch, ragged
    tails are handled, a pair without finders reports +inf / -1, and no distance is non-finite."""
    import torch
    s, P = _gpu_prop("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    N = (1 << 20) + 37
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand((N, P.nx), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1) * 3.0
    d, f, pts = P.get_min_distances(pair, x)
    assert torch.isfinite(d).all() and ((f >= 0) & (f < 25)).all()
    perm = torch.randperm(N, device="cuda", generator=gen)
    d2, f2, pts2 = P.get_min_distances(pair, x[perm].contiguous())
    assert torch.equal(d2, d[perm]) and torch.equal(f2, f[perm]) and torch.equal(pts2, pts[perm])
    # outside contact the segment between the two points has the reported length (every finder of this pair)
    sel = d > 1e-3
    seg = (pts[sel, 3:] - pts[sel, :3]).norm(dim=1)
    assert sel.any() and (seg - d[sel]).abs().max().item() < 1e-9
    none = px.proxy_query_pair_3D("none", px.proxy_query_model_3D("a").addShape(px.box("b", None, None, (1, 1, 1))),
                                  px.proxy_query_model_3D("b").addShape(px.box("c", None, None, (1, 1, 1))))
    dn, fn = P.get_min_distances(none, x[:100].contiguous(), with_points=False)
    assert torch.isinf(dn).all() and (fn == -1).all()


# ---- the steering loop with its collision test (rkb_steer_feedback_checked) ---------------------------
def _steer_case(P, n, seed):
    rng = np.random.default_rng(seed)
    x0 = np.concatenate([rng.uniform(-3.0, 3.0, (n, P.n)), rng.uniform(-4.0, 4.0, (n, P.n))], axis=1)
    if not P.blocked:  # interleaved (q0, qd0, q1, qd1, ...)
        x0 = np.stack([x0[:, :P.n], x0[:, P.n:]], axis=2).reshape(n, P.nx)
    goal = x0 + rng.uniform(-1.0, 1.0, (n, P.nx))
    u_bias = rng.uniform(-1.0, 1.0, (n, P.nu))
    gain = rng.uniform(-2.0, 2.0, (n, P.nu, P.nx))
    u_prev = rng.uniform(-0.5, 0.5, (n, P.nu))
    return x0, goal, u_bias, gain, u_prev


def _same_loop(got, want, J):
    xo, ul, nd, tr, st, col = got
    xr, ur, nr, trr, sr, colr = want
    assert np.array_equal(nd, nr) and np.array_equal(col, colr)
    assert np.max(np.abs(xo - xr)) < 1e-9 and np.max(np.abs(ul - ur)) < 1e-9
    for i in range(len(nd)):
        assert np.max(np.abs(tr[i, :nd[i]] - trr[i, :nr[i]]), initial=0.0) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("preset,track", [("crs6", False), ("crs7", True)])
def test_gpu_steer_feedback_checked_vs_reference(preset, track, oracle_built):
    """loops of MEAQR_topology.hpp:503-561 / IHAQR_topology.hpp:349-378 with with_collision_check = true: the live
    reference integrates, queries its proxy pairs at x_next and accepts or stops"""
    need_ref(oracle_built)
    s, P = _gpu_prop(preset)
    robot, lab = presets.crs_proxy_models(s, track=track)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    je = s.joint_end_frames
    extra = px.proxy_query_pair_3D("tool-obstacle", px.proxy_query_model_3D("tool").addShape(px.sphere("tool", je[-1], None, 0.12)),
                                   px.proxy_query_model_3D("obstacle").addShape(px.box("crate", None, px.pose_3D((0.3, -3.0, 0.9)), (0.5, 0.5, 0.5))))
    R = oracle_built.Reference(P.compiled)
    n, J = 500, 8
    x0, goal, u_bias, gain, u_prev = _steer_case(P, n, seed=9)
    kw = dict(bounds=(-2 * np.ones(P.nu), 2 * np.ones(P.nu)), rate_bounds=(-60 * np.ones(P.nu), 60 * np.ones(P.nu)))
    args = (x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, J, 0.25)
    for pairs in ([pair], [pair, extra]):
        got = P.steer_feedback(*args, want_traj=True, proxy_pairs=pairs, **kw)
        want = R.steer_feedback(*args, proxy_pairs=pairs, **kw)
        _same_loop(got, want, J)
        nd, col = got[2], got[5]
        assert (col == 1).any() and (col == 0).any()
        assert ((col == 1) & (nd > 0)).any(), "no loop stopped on a collision after its first interval"
        assert np.all(nd[col == 1] < J)
    # a pair without finders never rejects: the checked loop equals the unchecked (fused) one
    none = px.proxy_query_pair_3D("none", px.proxy_query_model_3D("a").addShape(px.box("b", None, None, (1, 1, 1))),
                                  px.proxy_query_model_3D("b").addShape(px.box("c", None, None, (1, 1, 1))))
    a = P.steer_feedback(*args, want_traj=True, proxy_pairs=[none], **kw)
    b = P.steer_feedback(*args, want_traj=True, **kw)
    assert not a[5].any() and np.array_equal(a[2], b[2])
    assert np.max(np.abs(a[0] - b[0])) < 1e-10 and np.max(np.abs(a[1] - b[1])) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("preset,track", [("crs6", False), ("crs7", True), ("crs6_sd", False)])
def test_gpu_steer_checked_one_launch_vs_reference(preset, track, oracle_built):
    """rkb_steer_checked_specialize: the steering kernel compiled with the collision test of the pairs built in — against
    the live reference's checked loop and against the interval-by-interval path"""
    need_ref(oracle_built)
    s, P = _gpu_prop(preset)
    robot, lab = presets.crs_proxy_models(s, track=track)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    je = s.joint_end_frames
    extra = px.proxy_query_pair_3D("tool-obstacle", px.proxy_query_model_3D("tool").addShape(px.sphere("tool", je[-1], None, 0.12)),
                                   px.proxy_query_model_3D("obstacle").addShape(px.box("crate", None, px.pose_3D((0.3, -3.0, 0.9)), (0.5, 0.5, 0.5))))
    R = oracle_built.Reference(P.compiled)
    n, J = 500, 8
    x0, goal, u_bias, gain, u_prev = _steer_case(P, n, seed=9)
    kw = dict(bounds=(-2 * np.ones(P.nu), 2 * np.ones(P.nu)), rate_bounds=(-60 * np.ones(P.nu), 60 * np.ones(P.nu)))
    args = (x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, J, 0.25)
    for pairs in ([pair], [pair, extra]):
        before = P.steer_feedback(*args, want_traj=True, proxy_pairs=pairs, **kw)
        n_launch = P.launch_count()
        assert not P.checked_steering_is_specialized(pairs)
        P.specialize_checked_steering(pairs)
        assert P.checked_steering_is_specialized(pairs)
        l0 = P.launch_count()
        got = P.steer_feedback(*args, want_traj=True, proxy_pairs=pairs, **kw)
        assert P.launch_count() - l0 == 1 and n_launch > 1, "checked steering is not one launch"
        _same_loop(got, R.steer_feedback(*args, proxy_pairs=pairs, **kw), J)
        _same_loop(got, before, J)
        assert np.array_equal(got[4], before[4])   # status words
        nd, col = got[2], got[5]
        assert (col == 1).any() and (col == 0).any() and ((col == 1) & (nd > 0)).any()
    # without a steer record, and J = 0
    a = P.steer_feedback(*args, proxy_pairs=[pair], **kw)
    b = P.steer_feedback(*args, want_traj=True, proxy_pairs=[pair], **kw)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[5])
    z = P.steer_feedback(x0, goal, u_bias, gain, u_prev, 1e-2, 1e-3, 10, 0, 0.25, proxy_pairs=[pair], **kw)
    assert np.array_equal(z[0], x0) and not z[2].any() and not z[4].any()


@pytest.mark.gpu
def test_gpu_steer_checked_auto_specialize():
    """default behaviour: the first checked call of >= 4096 tuples starts the compilation, calls made meanwhile run
    interval by interval, later ones in one launch — same loops throughout"""
    import time
    s, P = _gpu_prop("crs6", auto=True)
    robot, lab = presets.crs_proxy_models(s)
    lab.mShapeList[0].pose.position = (0.017, 0.0, 0.0)   # a pair no other test has put into the cubin cache
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    n, J = 4200, 4
    args = _steer_case(P, n, seed=4) + (1e-2, 1e-3, 10, J, 0.25)
    first = P.steer_feedback(*args, want_traj=True, proxy_pairs=[pair])
    t0 = time.time()
    while not P.checked_steering_is_specialized([pair]) and time.time() - t0 < 180:
        time.sleep(0.2)
        P.steer_feedback(*args, want_traj=True, proxy_pairs=[pair])
    assert P.checked_steering_is_specialized([pair]), "no generated kernel after 180 s"
    l0 = P.launch_count()
    later = P.steer_feedback(*args, want_traj=True, proxy_pairs=[pair])
    assert P.launch_count() - l0 == 1
    _same_loop(later, first, J)


@pytest.mark.gpu
def test_gpu_steer_feedback_checked_device_buffers():
    import torch
    s, P = _gpu_prop("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    n, J = 4099, 5
    case = _steer_case(P, n, seed=2)
    args_h = case + (1e-2, 1e-3, 10, J, 0.25)
    args_d = tuple(torch.from_numpy(a).cuda() for a in case) + (1e-2, 1e-3, 10, J, 0.25)
    h = P.steer_feedback(*args_h, want_traj=True, proxy_pairs=[pair])
    d = P.steer_feedback(*args_d, want_traj=True, proxy_pairs=[pair])
    for a, b in zip(h, d):
        b = b.cpu().numpy()
        if a.ndim == 3:  # the steer record is only defined below n_done
            for i in range(n):
                assert np.array_equal(a[i, :h[2][i]], b[i, :h[2][i]])
        else:
            assert np.array_equal(a, b)
    # the end state is free whenever an interval was accepted, and a rejected sample sits on its last accepted state
    dist, _ = P.get_min_distances(pair, h[0], with_points=False)
    assert np.all(dist[h[2] > 0] >= 0.0)
    stuck = (h[5] == 1) & (h[2] == 0)
    assert stuck.any() and np.array_equal(h[0][stuck], case[0][stuck])


@pytest.mark.gpu
def test_gpu_many_frames_vs_reference(oracle_built):
    """a 9-joint arm has 19 frames: the kernel instance for more than 16 frames"""
    need_ref(oracle_built)
    from reak_b200.propagator import kte_batch_propagator
    s = presets.crs_chain(n_revolute=9)
    P = kte_batch_propagator(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    assert P.compiled.desc.n_frames > 16
    pair = mixed_models(s, np.random.default_rng(77), n1=9, n2=6)
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 512, seed=4, q_range=2.5)
    agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL_SEARCH)


def branched_system():
    """a two-joint arm with a side branch: the rigid link of the branch starts from a frame that is not the one
    written last (the kernel keeps such frames in its slot array instead of in registers)"""
    from reak_b200.presets import kte_system
    s = kte_system("branched")
    s.joint_end_frames = []
    base = kte.frame_3D()
    base.Acceleration = [0.0, 0.0, 9.81]
    base.Position = [0.1, -0.2, 0.3]
    base.Quat = kte.axis_angle_quat(0.4, (0.0, 0.0, 1.0))
    ident = (1.0, 0.0, 0.0, 0.0)
    tensor = (0.3, 0.0, 0.0, 0.2, 0.0, 0.1)
    ups, rotors = [], []

    def joint(idx, axis, frm):
        coord, jac, end = kte.gen_coord(), kte.jacobian_gen_3D(), kte.frame_3D()
        j = kte.revolute_joint_3D("joint_%d" % idx, coord, axis, frm, end, jac)
        dep = kte.joint_dependent_gen_coord(coord)
        dep.add_joint(coord, kte.jacobian_gen_gen(1.0, 0.0))
        rotor = kte.inertia_gen("rotor_%d" % idx, dep, 1.0)
        act = kte.driving_actuator_gen("act_%d" % idx, coord, j)
        s.chain << act << rotor << j
        s.inputs.append(act)
        s.dofs_gen.append(coord)
        ups.append((coord, jac))
        rotors.append(rotor)
        s.joint_end_frames.append(end)
        return end

    def link(name, frm, offset, n_up, mass):
        nxt = kte.frame_3D()
        depf = kte.joint_dependent_frame_3D(nxt)
        for c, j in ups[:n_up]:
            depf.add_joint(c, j)
        inertia = kte.inertia_3D(name + "_inertia", depf, mass, tensor)
        s.chain << kte.rigid_link_3D(name, frm, nxt, kte.pose_3D(offset, ident)) << inertia
        s.mass_calc << inertia
        return nxt

    e1 = joint(0, (0.0, 0.0, 1.0), base)
    n1 = link("link_0", e1, (0.0, 0.0, 0.4), 1, 2.0)
    e2 = joint(1, (0.0, -1.0, 0.0), n1)
    n2 = link("link_1", e2, (0.0, 0.0, 0.3), 2, 1.0)
    n3 = link("branch", n1, (0.25, 0.0, 0.1), 1, 0.5)  # starts from n1 although n2 was written last
    for r in rotors:
        s.mass_calc << r
    for c in s.dofs_gen:
        s.mass_calc << c
    s.tips = (n2, n3)
    return s


def _branched_pair(s):
    m1 = px.proxy_query_model_3D("arm").addShape(px.capped_cylinder("upper", s.tips[0], px.pose_3D((0, 0, -0.15)), 0.3, 0.04)) \
        .addShape(px.sphere("side", s.tips[1], None, 0.07)).addShape(px.sphere("elbow", s.joint_end_frames[1], None, 0.06))
    m2 = px.proxy_query_model_3D("world").addShape(px.plane("floor", None, px.pose_3D((0, 0, 0.45)), (3, 3))) \
        .addShape(px.box("crate", None, px.pose_3D((0.45, -0.1, 0.85)), (0.3, 0.3, 0.3))) \
        .addShape(px.capped_cylinder("post", None, px.pose_3D((-0.2, 0.1, 0.8)), 1.0, 0.05))
    return px.proxy_query_pair_3D("branched", m1, m2)


def test_branched_chain_host_vs_reference(host_lib, oracle_built):
    need_ref(oracle_built)
    s = branched_system()
    pair = _branched_pair(s)
    H = HostProximity(host_lib, s, pair)
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 64, seed=3, q_range=3.0)
    want = R.min_distance(pair, x)
    agree(H.min_distance(ref_frames(R, x)), want, TOL_SEARCH)
    assert (want[0] < 0).any() and (want[0] > 0).any() and len(set(want[1].tolist())) >= 3
    H.close()


@pytest.mark.gpu
def test_gpu_branched_chain_vs_reference(oracle_built):
    need_ref(oracle_built)
    from reak_b200.propagator import kte_batch_propagator
    s = branched_system()
    P = kte_batch_propagator(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    pair = _branched_pair(s)
    R = oracle_built.Reference(P.compiled)
    x, u = random_batch(P.compiled, 777, seed=3, q_range=3.0)
    agree(P.get_min_distances(pair, x), R.min_distance(pair, x), TOL_SEARCH)
    # and the dynamics of the same branched chain, for good measure
    xd, st = P.get_state_derivatives(x, u)
    assert not st.any() and np.max(np.abs(xd - R.eval(x, u)[0]) / np.maximum(1.0, np.abs(xd))) < 1e-10


@pytest.mark.gpu
def test_gpu_is_free_two_pairs_vs_reference(oracle_built):
    """rkb_is_free = manip_dk_proxy_env_impl::is_free: no pair with a negative findMinimumDistance"""
    need_ref(oracle_built)
    import torch
    s, P = _gpu_prop("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    extra = px.proxy_query_pair_3D("tool-obstacle", px.proxy_query_model_3D("tool").addShape(px.sphere("tool", s.joint_end_frames[-1], None, 0.12)),
                                   px.proxy_query_model_3D("obstacle").addShape(px.box("crate", None, px.pose_3D((0.3, -3.0, 0.9)), (0.5, 0.5, 0.5))))
    none = px.proxy_query_pair_3D("none", px.proxy_query_model_3D("a").addShape(px.box("b", None, None, (1, 1, 1))),
                                  px.proxy_query_model_3D("b").addShape(px.box("c", None, None, (1, 1, 1))))
    R = oracle_built.Reference(P.compiled)
    x, _ = random_batch(P.compiled, 3001, seed=8, q_range=3.1)
    d1, d2 = R.min_distance(pair, x)[0], R.min_distance(extra, x)[0]
    want = ~((d1 < 0) | (d2 < 0))
    got = P.is_free([pair, extra, none], x)
    assert np.array_equal(got, want) and want.any() and (~want).any()
    assert ((d1 >= 0) & (d2 < 0)).any(), "the second pair never decides"
    assert np.array_equal(P.is_free([pair], x), ~(d1 < 0)) and P.is_free([none], x).all()
    got_d = P.is_free([pair, extra, none], torch.from_numpy(x).cuda())
    assert np.array_equal(got_d.cpu().numpy(), want)


# ---- gatherCollisionPoints ------------------------------------------------------------------------
def _gather_agree(cnt, rec, want_cnt, want_rec, tol):
    assert np.array_equal(cnt, want_cnt)
    live = np.isfinite(want_rec[:, :, 0])
    assert np.array_equal(live, np.isfinite(rec[:, :, 0]))
    assert np.max(np.abs(rec[live] - want_rec[live]), initial=0.0) < tol


def test_gather_collision_points_host_vs_reference(host_lib, oracle_built):
    """proxy_query_pair_3D::gatherCollisionPoints (proxy_query_model.cpp:402-421): the device source compiled for the host
    against the live reference — which finders collide, in which order, distance and both points of each"""
    need_ref(oracle_built)
    s = presets.make("crs6")
    robot, lab = presets.crs_proxy_models(s)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    H = HostProximity(host_lib, s, pair)
    R = oracle_built.Reference(H.compiled)
    x, _ = random_batch(H.compiled, 400, seed=31, q_range=3.1)
    want_cnt, want_rec = R.collision_points(pair, x, 25)
    cnt, fnd, rec = H.gather(ref_frames(R, x), 25)
    _gather_agree(cnt, rec, want_cnt, want_rec, 1e-12)
    assert want_cnt.max() >= 2 and (want_cnt == 0).any()
    assert all((np.diff(fnd[i, :cnt[i]]) > 0).all() for i in range(len(cnt)))   # createProxFinderList order
    # a record buffer shorter than the number of collisions: all are counted, the first are kept
    cnt2, fnd2, rec2 = H.gather(ref_frames(R, x), 1)
    assert np.array_equal(cnt2, want_cnt) and np.array_equal(rec2[:, 0], rec[:, 0])
    H.close()


@pytest.mark.gpu
def test_gpu_gather_collision_points_vs_reference(oracle_built):
    need_ref(oracle_built)
    for preset, track in (("crs6", False), ("crs7", True)):
        s, P = _gpu_prop(preset)
        robot, lab = presets.crs_proxy_models(s, track=track)
        pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
        R = oracle_built.Reference(P.compiled)
        x, _ = random_batch(P.compiled, 3000, seed=32, q_range=3.1)
        cnt, fnd, rec = P.gather_collision_points(pair, x)       # one slot per finder of the pair: nothing is dropped
        want_cnt, want_rec = R.collision_points(pair, x, rec.shape[1])
        _gather_agree(cnt, rec, want_cnt, want_rec, TOL)
        # (no simple relation to findMinimumDistance holds: that query evaluates its first finder unconditionally and culls
        # the others against the running minimum, this one culls every finder against 0 — and a plane's "bounding radius"
        # does not bound what its finders measure, DESIGN.md 4.4.  Both are held to the reference separately.)
        assert (cnt > 0).any() and (cnt == 0).any()
        # device buffers, short record buffer
        import torch
        c2, f2, r2 = P.gather_collision_points(pair, torch.from_numpy(x).cuda(), max_records=2)
        assert np.array_equal(c2.cpu().numpy(), cnt) and np.array_equal(r2.cpu().numpy(), rec[:, :2])
