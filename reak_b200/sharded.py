"""Sample-sharded propagation over the GPUs of one box: one process per GPU (torch.distributed),
contiguous block partition by sample index (pairs, for the steer batch), NO communication while
integrating, and one all-gather of the end states at the end — NCCL over NVLink/NVSwitch on GPUs,
gloo in the CPU tests.  The reference has no counterpart (it is single-threaded); the semantics
per sample are those of kte_batch_propagator.get_next_states / steer_batch.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """[lo, hi) of the contiguous block of `n` samples owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def _dist():
    import torch.distributed as dist
    return dist


def _all_gather_rows(local, n_total, group=None):
    """All-gather row blocks of unequal length (block partition of n_total rows) into the full array.
    `local` is a torch tensor [rows_of_this_rank, ...] on the device the backend communicates from."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    cap = max(sizes) if sizes else 0
    tail = tuple(local.shape[1:])
    if min(sizes) == cap:
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((cap,) + tail, dtype=local.dtype, device=local.device)
    padded[: sizes[rank]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


class sharded_propagator(object):
    """Wraps one kte_batch_propagator per rank.  Every rank passes the same full-batch arrays (or
    only its own block with `local_input=True`); results come back as the full batch on every rank."""

    def __init__(self, propagator, group=None, comm_device=None):
        self.prop, self.group = propagator, group
        dist = _dist()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.comm_device = comm_device  # torch device the collective runs from (cuda:k for NCCL, cpu for gloo)

    # ---- all-gather by the rollout kernel's own stores (peer memory over NVLink) -----------------------------------
    def _peer_buffers(self, n_total, nx, device):
        """Symmetric result buffers [n_total][nx] float64 + [n_total] int32 on every rank, each rank's mapped into all the
        others (torch.distributed's symmetric memory: CUDA virtual-memory handles exchanged at rendezvous).  Allocated
        once per shape.  Returns None where peer mapping is not available (then the NCCL path is used)."""
        cache = self.__dict__.setdefault("_peer_cache", {})
        key = (int(n_total), int(nx), str(device))
        if key in cache:
            return cache[key]
        entry = None
        try:
            import torch
            import torch.distributed._symmetric_memory as symm_mem
            dist = _dist()
            group = self.group if self.group is not None else dist.group.WORLD
            x = symm_mem.empty((n_total, nx), dtype=torch.float64, device=device)
            st = symm_mem.empty((n_total,), dtype=torch.int32, device=device)
            hx, hs = symm_mem.rendezvous(x, group), symm_mem.rendezvous(st, group)
            entry = {"x": x, "st": st, "hx": hx, "hs": hs, "px": [int(p) for p in hx.buffer_ptrs], "ps": [int(p) for p in hs.buffer_ptrs]}
        except Exception as e:  # no peer access / no symmetric-memory support in this build or on this box
            self.__dict__["_peer_error"] = "%s: %s" % (type(e).__name__, e)
            entry = None
        # every rank must take the same path: agree on the outcome
        try:
            import torch
            dist = _dist()
            ok = torch.tensor([1 if entry is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                entry = None
        except Exception:
            entry = None
        cache[key] = entry
        return entry

    def release_peer_buffers(self):
        """Drop the symmetric buffers (collectively: call it on every rank, before the process group is destroyed)."""
        cache = self.__dict__.get("_peer_cache", {})
        for entry in list(cache.values()):
            if entry is not None:
                entry.clear()
        cache.clear()

    def get_next_states_peer_stores(self, xb, ub, dt, n_steps, n_total):
        """This rank's block integrated by rkb_rollout_rk4_scatter: the kernel stores every end state into ALL ranks'
        copies of the gathered batch (its own and, over NVLink, the peers'), so that no collective follows — only a
        barrier.  xb, ub: this rank's block, CUDA tensors.  Returns (x_out[n_total][nx], status[n_total]) — the rank's
        symmetric buffers, valid until the next call of the same shape — or None where peer mapping is unavailable."""
        import ctypes as C
        import torch
        from . import _abi
        prop = self.prop
        if n_total % self.world or xb.shape[0] != n_total // self.world or not prop.is_serial():
            return None
        buf = self._peer_buffers(n_total, prop.nx, xb.device)
        if buf is None:
            return None
        n_loc = n_total // self.world
        px = (C.c_void_p * self.world)(*buf["px"])
        ps = (C.c_void_p * self.world)(*buf["ps"])
        flags = _abi.MEM_DEVICE | _abi.LAYOUT_AOS | (_abi.LAYOUT_BLOCKED if prop.blocked else 0)
        stream = C.c_void_p(torch.cuda.current_stream(xb.device).cuda_stream)
        buf["hx"].barrier(channel=0)   # nobody is still reading the previous call's results
        _abi.check(prop._lib.rkb_rollout_rk4_scatter(prop._h, prop.device, n_loc, C.c_void_p(xb.data_ptr()),
                                                     C.c_void_p(ub.data_ptr()) if prop.nu else None, float(dt), int(n_steps), self.world,
                                                     px, ps, self.rank * n_loc, flags, stream), "rkb_rollout_rk4_scatter")
        buf["hx"].barrier(channel=1)   # every rank's stores have landed in every copy
        return buf["x"], buf["st"]

    def _to_comm(self, a, dtype):
        import torch
        if type(a).__module__.startswith("torch"):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(dtype)
        return t.to(self.comm_device) if self.comm_device is not None else t

    def get_next_states(self, x, u, dt, n_steps, local_input=False, n_total=None, compute=None, chunks=1, out=None, status=None,
                        chunk_samples=None, peer_stores=False):
        """Returns (x_out[n_total][nx], status[n_total]) gathered on every rank (torch tensors on the
        communication device).  `compute(x_block, u_block, dt, n_steps) -> (x_out, status)` defaults to
        the wrapped propagator's GPU rollout.

        chunks > 1 (and n_total divisible by world * chunks): the rank's block is integrated in `chunks` pieces and
        the all-gather of piece c is issued asynchronously right behind its rollout, so that it travels while piece
        c + 1 integrates — only the last piece's gather is exposed.  chunk_samples: the piece size instead of
        the piece count (the last piece takes the remainder) — use a multiple of the propagator's wave_samples(), every
        launch ends with a partial wave.  out / status: preallocated result tensors on the communication device
        ([n_total][nx] float64, [n_total] int32).

        peer_stores=True (CUDA tensors, serial chain, equal blocks): no collective at all — the rollout kernel itself
        stores every end state into all ranks' result buffers over NVLink (get_next_states_peer_stores); the returned
        tensors are then the rank's symmetric buffers (out / status are not used).  Falls back to the NCCL paths where
        peer mapping is unavailable."""
        import torch
        compute = compute or self.prop.get_next_states
        if local_input:
            if n_total is None:
                raise ValueError("n_total is required with local_input")
            xb, ub = x, u
        else:
            n_total = x.shape[0]
            lo, hi = shard_bounds(n_total, self.rank, self.world)
            xb, ub = x[lo:hi], (u[lo:hi] if u is not None else None)
        if peer_stores and n_total > 0 and compute == self.prop.get_next_states and type(xb).__module__.startswith("torch") and xb.is_cuda:
            res = self.get_next_states_peer_stores(xb.contiguous(), ub.contiguous() if ub is not None else None, dt, n_steps, n_total)
            if res is not None:
                return res
        if n_total > 0 and n_total % self.world == 0:
            n_loc = n_total // self.world
            if chunk_samples and 0 < chunk_samples < n_loc:
                bounds = list(range(0, n_loc, int(chunk_samples))) + [n_loc]
                return self._get_next_states_chunked(xb, ub, dt, n_steps, n_total, compute, bounds, out, status)
            if chunks > 1 and n_loc % chunks == 0:
                bounds = [c * (n_loc // chunks) for c in range(chunks + 1)]
                return self._get_next_states_chunked(xb, ub, dt, n_steps, n_total, compute, bounds, out, status)
        xo, st = compute(xb, ub, dt, n_steps)
        xo_t, st_t = self._to_comm(xo, torch.float64), self._to_comm(st, torch.int32)
        full, full_st = _all_gather_rows(xo_t, n_total, self.group), _all_gather_rows(st_t, n_total, self.group)
        if out is not None:
            out.copy_(full); full = out
        if status is not None:
            status.copy_(full_st); full_st = status
        return full, full_st

    def _get_next_states_chunked(self, xb, ub, dt, n_steps, n_total, compute, bounds, out, status):
        """bounds: piece c covers rows bounds[c] .. bounds[c + 1] of every rank's block (all ranks cut alike)"""
        import torch
        dist = _dist()
        n_loc = n_total // self.world
        assert xb.shape[0] == n_loc and bounds[0] == 0 and bounds[-1] == n_loc
        full, st_loc, works = out, [], []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            xo, st = compute(xb[lo:hi], ub[lo:hi] if ub is not None else None, dt, n_steps)
            xo_t = self._to_comm(xo, torch.float64).contiguous()
            st_loc.append(self._to_comm(st, torch.int32))
            if full is None:
                full = torch.empty((n_total,) + tuple(xo_t.shape[1:]), dtype=torch.float64, device=xo_t.device)
            # rank r's piece lands where the block partition puts it: rows r n_loc + lo ...
            dst = [full[r * n_loc + lo: r * n_loc + hi] for r in range(self.world)]
            works.append(dist.all_gather(dst, xo_t, group=self.group, async_op=True))
        st_t = torch.cat(st_loc, dim=0).contiguous()
        full_st = status if status is not None else torch.empty((n_total,), dtype=torch.int32, device=st_t.device)
        works.append(dist.all_gather_into_tensor(full_st, st_t, group=self.group, async_op=True))
        for w in works:
            w.wait()
        return full, full_st

    def rollout(self, x, u_seq, dt, steps_per_interval, scheme="rk4", compute=None):
        """kte_batch_propagator.rollout sharded by sample: returns (x_out[N][nx], x_traj[N][J][nx], status[N])
        gathered on every rank.  `compute(x, u_seq, dt, steps, scheme) -> (x_out, x_traj, status)`."""
        import torch
        compute = compute or (lambda xb, ub, d, k, sc: self.prop.rollout(xb, ub, d, k, scheme=sc, want_traj=True))
        n_total = x.shape[0]
        lo, hi = shard_bounds(n_total, self.rank, self.world)
        xo, tr, st = compute(x[lo:hi], u_seq[lo:hi], dt, steps_per_interval, scheme)
        return (_all_gather_rows(self._to_comm(xo, torch.float64), n_total, self.group),
                _all_gather_rows(self._to_comm(tr, torch.float64), n_total, self.group),
                _all_gather_rows(self._to_comm(st, torch.int32), n_total, self.group))

    def steer_feedback(self, x0, goal, u_bias, gain, u_prev, *args, compute=None, **kw):
        """kte_batch_propagator.steer_feedback sharded by tuple: returns (x_out, u_last, n_done, status[, collided])
        gathered on every rank; positional and keyword arguments after u_prev (proxy_pairs=... included) are passed through."""
        import torch
        compute = compute or self.prop.steer_feedback
        n_total = x0.shape[0]
        lo, hi = shard_bounds(n_total, self.rank, self.world)
        outs = compute(x0[lo:hi], goal[lo:hi], u_bias[lo:hi], gain[lo:hi], u_prev[lo:hi], *args, **kw)
        g = lambda a, dt: _all_gather_rows(self._to_comm(a, dt), n_total, self.group)
        # (x_out, u_last, n_done, status) and, with proxy_pairs=..., collided: states and inputs are doubles, the rest int32
        return tuple(g(a, torch.float64 if k < 2 else torch.int32) for k, a in enumerate(outs))

    def get_min_distances(self, pair, x, compute=None):
        """kte_batch_propagator.get_min_distances sharded by state: (distance[N], finder[N]) gathered on every rank.
        `compute(pair, x_block) -> (distance, finder)`."""
        import torch
        compute = compute or (lambda pr, xb: self.prop.get_min_distances(pr, xb, with_points=False))
        n_total = x.shape[0]
        lo, hi = shard_bounds(n_total, self.rank, self.world)
        d, f = compute(pair, x[lo:hi])[:2]
        return (_all_gather_rows(self._to_comm(d, torch.float64), n_total, self.group),
                _all_gather_rows(self._to_comm(f, torch.int32), n_total, self.group))

    def steer_batch(self, x0, goal, u, dt, n_steps, compute=None):
        """Pairs are never split across ranks, so the per-pair arg-min stays on one device; only
        (best_idx, best_x, best_cost) per pair travel."""
        import torch
        compute = compute or self.prop.steer_batch
        P = x0.shape[0]
        lo, hi = shard_bounds(P, self.rank, self.world)
        idx, bx, bc = compute(x0[lo:hi], goal[lo:hi], u[lo:hi], dt, n_steps)[:3]
        return (_all_gather_rows(self._to_comm(idx, torch.int32), P, self.group),
                _all_gather_rows(self._to_comm(bx, torch.float64), P, self.group),
                _all_gather_rows(self._to_comm(bc, torch.float64), P, self.group))
