"""SURVEY f4, nearest neighbours: rkb_nearest — the search ReaK's planners run before every steer
(ReaK::pp::linear_neighbor_search / dvp_tree: ctrl/path_planning/topological_search.hpp:91-112, 238-270, 586-596;
metric_space_search.hpp) — for a batch of query points at once.

Oracle: kto_nearest (oracle/kte_oracle.c), pinned bit for bit against ReaK::pp::min_dist_linear_search compiled from the
reference (rkref_nearest) and against tests/golden/nearest/nearest.npz generated from it.  GPU: indices AND distances must be
bit-identical to the oracle (index work: exact)."""
import os

import numpy as np
import pytest

from reak_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    g = np.load(os.path.join(HERE, "golden", "nearest", "nearest.npz"))
    for n in range(int(g["n_cases"])):
        dim, k, radius = g["par%d" % n]
        yield g["v%d" % n], g["q%d" % n], int(k), float(radius), g["idx%d" % n], g["dist%d" % n], g["cnt%d" % n]


def test_oracle_matches_golden(oracle_built):
    for v, q, k, radius, idx, dist, cnt in _golden():
        a = oracle_built.nearest("oracle", v, q, k, radius)
        assert np.array_equal(a[0], idx) and np.array_equal(a[1], dist) and np.array_equal(a[2], cnt)


def test_oracle_matches_live_reference(oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("compiled reference not built here")
    rng = np.random.default_rng(9)
    for dim, k, radius in ((1, 2, np.inf), (3, 1, np.inf), (12, 7, 0.9), (24, 16, np.inf), (45, 3, 3.5)):
        v = rng.uniform(-1.0, 1.0, (1500, dim))
        q = rng.uniform(-1.0, 1.0, (30, dim))
        a, b = oracle_built.nearest("oracle", v, q, k, radius), oracle_built.nearest("ref", v, q, k, radius)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (dim, k, radius)
    # the single-neighbour form keeps the FIRST of equal minima (topological_search.hpp:102-110)
    v = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])
    for which in ("oracle", "ref"):
        assert oracle_built.nearest(which, v, np.zeros((1, 2)), 1)[0][0, 0] == 0


def test_search_semantics(oracle_built):
    """the radius is exclusive (compare(d, radius) must hold), fewer than k -> -1 / +inf, empty vertex set"""
    v = np.array([[0.0], [1.0], [2.0], [3.0]])
    q = np.array([[0.0]])
    idx, dist, cnt = oracle_built.nearest("oracle", v, q, 3, 2.0)
    assert idx.tolist() == [[0, 1, -1]] and dist[0, :2].tolist() == [0.0, 1.0] and np.isinf(dist[0, 2]) and cnt[0] == 2
    idx, dist, cnt = oracle_built.nearest("oracle", v, q, 4, np.nextafter(2.0, 3.0))
    assert idx.tolist() == [[0, 1, 2, -1]]
    idx, dist, cnt = oracle_built.nearest("oracle", np.zeros((0, 1)), q, 2)
    assert idx.tolist() == [[-1, -1]] and cnt[0] == 0


def test_argument_checks():
    lib = _abi.load_library()
    import ctypes as C
    v = np.zeros((4, 3)); q = np.zeros((2, 3)); idx = np.zeros((2, 1), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    inf = float("inf")
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 0, 1, inf, p(idx), None, None, 0, None) == _abi.ERR_INVALID           # dim
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 3, 0, inf, p(idx), None, None, 0, None) == _abi.ERR_INVALID           # k
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 3, 17, inf, p(idx), None, None, 0, None) == _abi.ERR_INVALID          # k > RKB_NEAREST_MAX_K
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 49, 1, inf, p(idx), None, None, 0, None) == _abi.ERR_INVALID          # dim > RKB_NEAREST_MAX_DIM
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 3, 1, -1.0, p(idx), None, None, 0, None) == _abi.ERR_INVALID          # radius
    assert lib.rkb_nearest(0, 4, p(v), 2, p(q), 3, 1, inf, None, None, None, 0, None) == _abi.ERR_INVALID             # no output
    assert lib.rkb_nearest(0, 4, p(v), 0, None, 3, 1, inf, None, None, None, 0, None) == 0                            # nothing asked


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_against_golden():
    from reak_b200.nearest import nearest_neighbors
    for v, q, k, radius, idx, dist, cnt in _golden():
        a = nearest_neighbors(v, q, k, radius)
        assert np.array_equal(a[0], idx) and np.array_equal(a[1], dist) and np.array_equal(a[2], cnt)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3, 6, 12, 13, 19, 32, 45, 48])
def test_gpu_bit_identical_to_oracle(dim, oracle_built):
    from reak_b200.nearest import nearest_neighbors
    rng = np.random.default_rng(100 + dim)
    for V, Q, k, radius in ((1, 3, 1, np.inf), (63, 129, 4, np.inf), (64, 1, 16, np.inf), (65, 127, 2, 0.8 * np.sqrt(dim / 6.0)),
                            (1000, 300, 16, np.inf), (20011, 97, 5, np.inf), (5000, 2000, 1, 0.5 * np.sqrt(dim / 6.0))):
        v = rng.uniform(-1.0, 1.0, (V, dim))
        q = rng.uniform(-1.0, 1.0, (Q, dim))
        a = nearest_neighbors(v, q, k, radius)
        b = oracle_built.nearest("oracle", v, q, k, radius)
        assert np.array_equal(a[0], b[0]), (dim, V, Q, k, radius)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (dim, V, Q, k, radius)


@pytest.mark.gpu
def test_gpu_ties_empty_and_device_buffers(oracle_built):
    import torch
    from reak_b200.nearest import nearest_neighbors
    rng = np.random.default_rng(7)
    # duplicated vertices: equal distances -> ascending vertex index, across tile and chunk boundaries
    base = rng.uniform(-1.0, 1.0, (900, 6))
    v = np.concatenate([base, base[::-1], base])
    q = base[:50] + 1e-3
    idx, dist, cnt = nearest_neighbors(v, q, 6)
    ref = oracle_built.nearest("oracle", v, q, 6)
    assert np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1])
    for r in range(50):
        assert sorted(idx[r, :3].tolist()) == idx[r, :3].tolist() and dist[r, 0] == dist[r, 1] == dist[r, 2]
    # empty vertex set, query on a vertex
    idx, dist, cnt = nearest_neighbors(np.zeros((0, 6)), q, 2)
    assert (idx == -1).all() and np.isinf(dist).all() and not cnt.any()
    idx, dist, cnt = nearest_neighbors(base, base[:10], 1)
    assert idx[:, 0].tolist() == list(range(10)) and not dist.any()
    # device-resident tensors on the current stream
    big_v, big_q = rng.uniform(-1.0, 1.0, (100003, 12)), rng.uniform(-1.0, 1.0, (513, 12))
    a = nearest_neighbors(torch.from_numpy(big_v).cuda(), torch.from_numpy(big_q).cuda(), 3)
    b = nearest_neighbors(big_v, big_q, 3)
    c = oracle_built.nearest("oracle", big_v, big_q, 3)
    assert np.array_equal(a[0].cpu().numpy(), b[0]) and np.array_equal(a[1].cpu().numpy(), b[1])
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and np.array_equal(b[2], c[2])


@pytest.mark.gpu
def test_gpu_reak_side_batched_neighbor_search(oracle_built):
    """ReaK::pp::batched_neighbor_search (reak_bridge.hpp) — the batched form of linear_neighbor_search's iterator calls
    (topological_search.hpp:647-692) — compared in C++ with ReaK::pp::min_dist_linear_search, point by point"""
    import ctypes as C
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref/libreak_ref.so not built")
    C.CDLL(_abi.LIB_PATH, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(oracle_built.REF_SO)
    fn = lib.rkref_nn_bridge_check
    fn.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_char_p, C.c_int]
    rng = np.random.default_rng(21)
    for dim, V, Q, k, radius in ((12, 4000, 200, 5, np.inf), (19, 1500, 64, 16, 2.2), (2, 300, 33, 3, 0.2), (6, 0, 5, 2, np.inf)):
        v = rng.uniform(-1.0, 1.0, (V, dim))
        q = rng.uniform(-1.0, 1.0, (Q, dim))
        msg = C.create_string_buffer(256)
        assert fn(V, v.ctypes.data, Q, q.ctypes.data, dim, k, radius, msg, 256) == 0, (dim, V, msg.value)
