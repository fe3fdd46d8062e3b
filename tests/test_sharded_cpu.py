"""world_size-2 gloo test of the N > 1 path: block partition by sample index, no data-path
collective, one all-gather of the end states.  The per-rank compute is injected (the oracle
stands in for the GPU rollout here: this test covers the host-side sharding logic only)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _test_pair(system, px):
    """a capsule and a sphere on the last two joints of the arm against a floor and a crate"""
    je = system.joint_end_frames
    robot = px.proxy_query_model_3D("robot").addShape(px.capped_cylinder("l", je[-2], px.pose_3D((0, 0, 0.1)), 0.2, 0.05)) \
        .addShape(px.sphere("t", je[-1], None, 0.08))
    world = px.proxy_query_model_3D("world").addShape(px.plane("floor", None, px.pose_3D((0, -3.3, 0.45)), (4, 4))) \
        .addShape(px.box("crate", None, px.pose_3D((0.2, -3.2, 0.9)), (0.3, 0.3, 0.3)))
    return px.proxy_query_pair_3D("robot-world", robot, world)


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import pyref
    from reak_b200 import kte, presets
    from reak_b200.sharded import shard_bounds, sharded_propagator
    from conftest import random_batch

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    s = presets.make("crs3")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    O = pyref.Oracle(c)
    x, u = random_batch(c, n_total, seed=31)

    def compute(xb, ub, dt, k):
        xo, st, _ = O.rk4(xb, ub, dt, k)
        return xo, st

    sp = sharded_propagator(None, comm_device=None)
    full, st = sp.get_next_states(x, u, 1e-3, 5, compute=compute)
    lo, hi = shard_bounds(n_total, rank, world)
    loc, st2 = sp.get_next_states(x[lo:hi], u[lo:hi], 1e-3, 5, local_input=True, n_total=n_total, compute=compute)
    assert np.array_equal(full.numpy(), loc.numpy()) and np.array_equal(st.numpy(), st2.numpy())
    # the pipelined variant (piece c is gathered while piece c + 1 integrates) lands every piece in block-partition order;
    # with a size that does not divide it falls back to the single gather
    import torch
    pre, pre_st = torch.full((n_total, x.shape[1]), np.nan, dtype=torch.float64), torch.full((n_total,), -1, dtype=torch.int32)
    ch, st3 = sp.get_next_states(x, u, 1e-3, 5, compute=compute, chunks=4, out=pre, status=pre_st)
    assert ch is pre and st3 is pre_st
    assert np.array_equal(full.numpy(), ch.numpy()) and np.array_equal(st.numpy(), st3.numpy())
    ch2, st4 = sp.get_next_states(x, u, 1e-3, 5, compute=compute, chunk_samples=5)   # pieces of 5, the last one shorter
    assert np.array_equal(full.numpy(), ch2.numpy()) and np.array_equal(st.numpy(), st4.numpy())

    # steer: pairs are sharded, rollouts of a pair stay together
    P, R = 5, 7
    rng = np.random.default_rng(5)
    x0, goal, uu = x[:P], x[P:2 * P], rng.uniform(-2, 2, (P, R, c.n_inputs))

    def steer(x0b, gb, ub, dt, k):
        p = x0b.shape[0]
        xe, _, _ = O.rk4(np.repeat(x0b, R, axis=0), ub.reshape(p * R, -1), dt, k)
        cost = np.linalg.norm(xe.reshape(p, R, -1) - gb[:, None, :], axis=2)
        idx = cost.argmin(axis=1).astype(np.int32)
        return idx, xe.reshape(p, R, -1)[np.arange(p), idx], cost.min(axis=1)

    idx, bx, bc = sp.steer_batch(x0, goal, uu, 1e-3, 3, compute=steer)
    # control sequences and closed-loop steering shard the same way
    u_seq = rng.uniform(-1, 1, (n_total, 3, c.n_inputs))
    rx, rtr, rst = sp.rollout(x, u_seq, 1e-3, 2, scheme="rk5", compute=lambda xb, ub, d, k, sc: O.rollout(xb, ub, 5, d, k))
    goal_all = x + rng.uniform(-0.3, 0.3, x.shape)
    gain = rng.uniform(-3, 3, (n_total, c.n_inputs, 2 * c.n_coords))

    def feedback(x0b, gb, ubb, gnb, upb, *a, **k):
        xo, ul, nd, _, s2 = O.steer_feedback(x0b, gb, ubb, gnb, upb, *a, **k)
        return xo, ul, nd, s2

    fx, fu, fn, fs = sp.steer_feedback(x, goal_all, u, gain, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2, compute=feedback)
    extra = {}
    if pyref.have_ref():
        # proximity queries and the checked steering loop shard by state / tuple too (live reference as the compute)
        from reak_b200 import proximity as px
        R = pyref.Reference(c)
        pair = _test_pair(s, px)
        pd, pf = sp.get_min_distances(pair, x, compute=lambda pr, xb: R.min_distance(pr, xb)[:2])

        def checked(x0b, gb, ubb, gnb, upb, *a, **k):
            xo, ul, nd, _, s2, col = R.steer_feedback(x0b, gb, ubb, gnb, upb, *a, proxy_pairs=[pair], **k)
            return xo, ul, nd, s2, col

        cx, cu, cn, cs, cc = sp.steer_feedback(x, goal_all, u, gain, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2, compute=checked)
        extra = dict(pd=pd.numpy(), pf=pf.numpy(), cx=cx.numpy(), cn=cn.numpy(), cc=cc.numpy())
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), full=full.numpy(), st=st.numpy(), idx=idx.numpy(), bx=bx.numpy(), bc=bc.numpy(),
             rx=rx.numpy(), rtr=rtr.numpy(), fx=fx.numpy(), fu=fu.numpy(), fn=fn.numpy(), **extra)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])
def test_two_rank_gloo_gather(n_total, tmp_path, oracle_built):
    import torch.multiprocessing as mp
    from oracle import pyref
    from reak_b200 import kte, presets
    from conftest import random_batch

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    s = presets.make("crs3")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    x, u = random_batch(c, n_total, seed=31)
    want, st, _ = pyref.Oracle(c).rk4(x, u, 1e-3, 5)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for g in got:
        assert np.array_equal(g["full"], want) and not g["st"].any()
    assert np.array_equal(got[0]["idx"], got[1]["idx"]) and np.array_equal(got[0]["bx"], got[1]["bx"])
    assert got[0]["idx"].shape == (5,) and got[0]["bx"].shape == (5, 6)
    # sharded control-sequence rollout and closed-loop steering equal the unsharded oracle run
    O = pyref.Oracle(c)
    rng = np.random.default_rng(5)
    rng.uniform(-2, 2, (5, 7, c.n_inputs))  # same stream as the workers
    u_seq = rng.uniform(-1, 1, (n_total, 3, c.n_inputs))
    wx, wtr, _ = O.rollout(x, u_seq, 5, 1e-3, 2)
    goal_all = x + rng.uniform(-0.3, 0.3, x.shape)
    gain = rng.uniform(-3, 3, (n_total, c.n_inputs, 2 * c.n_coords))
    fx, fu, fn, _, _ = O.steer_feedback(x, goal_all, u, gain, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2)
    for g in got:
        assert np.array_equal(g["rx"], wx) and np.array_equal(g["rtr"], wtr)
        assert np.array_equal(g["fx"], fx) and np.array_equal(g["fu"], fu) and np.array_equal(g["fn"], fn)
    if pyref.have_ref():
        from reak_b200 import proximity as px
        R = pyref.Reference(c)
        pair = _test_pair(s, px)
        wd, wf, _ = R.min_distance(pair, x)
        cx, _, cn, _, _, cc = R.steer_feedback(x, goal_all, u, gain, 0.5 * u, 1e-2, 1e-3, 10, 3, 0.2, proxy_pairs=[pair])
        assert (wd < 0).any() and (wd > 0).any()
        for g in got:
            assert np.array_equal(g["pd"], wd) and np.array_equal(g["pf"], wf)
            assert np.array_equal(g["cx"], cx) and np.array_equal(g["cn"], cn) and np.array_equal(g["cc"], cc)


def test_shard_bounds_cover_everything():
    from reak_b200.sharded import shard_bounds
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
