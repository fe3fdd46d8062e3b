"""N > 1 on real GPUs: reak_b200.sharded over NCCL, one process per GPU (needs >= 2 devices; skipped otherwise —
run it with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`).  The sharded results must be
bit-identical to one GPU integrating the whole batch (samples are independent), with the end states all-gathered in
one piece and in pipelined pieces (piece c travels while piece c + 1 integrates)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from reak_b200 import kte_batch_propagator, presets
    from reak_b200.sharded import shard_bounds, sharded_propagator
    from conftest import random_batch

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    prop = kte_batch_propagator(presets.make("crs6_sd"), device=rank)
    x, u = random_batch(prop.compiled, n_total, seed=91)
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    sp = sharded_propagator(prop, comm_device=dev)
    # whole batch handed to every rank (device tensors), one gather
    full, st = sp.get_next_states(xd, ud, 1e-3, 20)
    # only the rank's block, pipelined gather into preallocated tensors
    lo, hi = shard_bounds(n_total, rank, world)
    pre = torch.full((n_total, prop.nx), float("nan"), dtype=torch.float64, device=dev)
    pre_st = torch.full((n_total,), -1, dtype=torch.int32, device=dev)
    ch, st2 = sp.get_next_states(xd[lo:hi].contiguous(), ud[lo:hi].contiguous(), 1e-3, 20, local_input=True, n_total=n_total,
                                 chunks=4, out=pre, status=pre_st)
    # no collective at all: the rollout kernel stores into every rank's symmetric buffer over NVLink
    ps = sp.get_next_states(xd[lo:hi].contiguous(), ud[lo:hi].contiguous(), 1e-3, 20, local_input=True, n_total=n_total, peer_stores=True)
    peer_used = getattr(sp, "_peer_error", None) is None and ps[0].data_ptr() != pre.data_ptr() and (n_total % world == 0)
    ps2 = sp.get_next_states(xd[lo:hi].contiguous(), ud[lo:hi].contiguous(), 1e-3, 20, local_input=True, n_total=n_total, peer_stores=True)
    ps_x, ps_st = ps2[0].clone(), ps2[1].clone()
    # host (numpy) inputs through the staged path of the library
    hfull, hst = sp.get_next_states(x, u, 1e-3, 20)
    # steer batch: pairs sharded
    P, R = 6, 16
    rng = np.random.default_rng(92)
    uu = rng.uniform(-3, 3, (P, R, prop.nu))
    idx, bx, bc = sp.steer_batch(x[:P], x[P:2 * P], uu, 1e-3, 10)
    one = None
    if rank == 0:  # the same batch on one GPU, unsharded
        o, s1 = prop.get_next_states(xd, ud, 1e-3, 20)
        i1, b1, c1 = prop.steer_batch(x[:P], x[P:2 * P], uu, 1e-3, 10)
        one = dict(one=o.cpu().numpy(), one_st=s1.cpu().numpy(), one_idx=i1, one_bx=b1, one_bc=c1)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), full=full.cpu().numpy(), st=st.cpu().numpy(), ch=ch.cpu().numpy(),
             st2=st2.cpu().numpy(), hfull=hfull.cpu().numpy(), idx=idx.cpu().numpy(), bx=bx.cpu().numpy(), bc=bc.cpu().numpy(),
             is_pre=np.array(ch.data_ptr() == pre.data_ptr()), ps=ps_x.cpu().numpy(), ps_st=ps_st.cpu().numpy(),
             peer_used=np.array(bool(peer_used)), peer_error=np.array(str(getattr(sp, "_peer_error", None))), **(one or {}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [4096, 1000])
def test_two_rank_nccl_sharded_propagator(n_total, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    want, want_st = got[0]["one"], got[0]["one_st"]
    assert not want_st.any()
    for g in got:
        assert np.array_equal(g["full"], want) and np.array_equal(g["st"], want_st)
        assert np.array_equal(g["ch"], want) and np.array_equal(g["st2"], want_st) and bool(g["is_pre"])
        assert np.array_equal(g["hfull"], want)
        assert np.array_equal(g["ps"], want) and np.array_equal(g["ps_st"], want_st)     # whichever path served it
        assert bool(g["peer_used"]), "peer-store gather fell back to NCCL: %s" % g["peer_error"]
        assert np.array_equal(g["idx"], got[0]["one_idx"]) and np.array_equal(g["bx"], got[0]["one_bx"])
        assert np.array_equal(g["bc"], got[0]["one_bc"])
