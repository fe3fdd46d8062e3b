"""Batched nearest-neighbour queries (rkb_nearest): what ReaK::pp::linear_neighbor_search / dvp_tree answer for one
sample at a time (ctrl/path_planning/topological_search.hpp:586-596, 619-637; metric_space_search.hpp), for a whole
batch of samples against the vertices of the motion graph.  numpy arrays (host) or torch CUDA tensors (device)."""
import ctypes as C

import numpy as np

from . import _abi


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def nearest_neighbors(vertices, queries, k=1, radius=float("inf"), device=0):
    """vertices [V][dim], queries [Q][dim] -> (index [Q][k] int32, distance [Q][k], count [Q] int32).
    index is -1 / distance +inf where fewer than k vertices lie within `radius` (strictly closer than it)."""
    lib = _abi.load_library()
    on_dev = _is_torch(queries)
    if on_dev != _is_torch(vertices):
        raise TypeError("vertices and queries must live in the same memory space")
    if on_dev:
        import torch
        if not (queries.is_cuda and vertices.is_cuda) or queries.dtype != torch.float64 or vertices.dtype != torch.float64:
            raise TypeError("device buffers must be float64 CUDA tensors")
        q, v = queries.contiguous(), vertices.contiguous()
    else:
        q = np.ascontiguousarray(queries, dtype=np.float64)
        v = np.ascontiguousarray(vertices, dtype=np.float64)
    if len(q.shape) != 2 or q.shape[1] < 1:
        raise IndexError("queries must be [Q][dim]")
    dim, Q = int(q.shape[1]), int(q.shape[0])
    V = int(v.shape[0]) if len(v.shape) == 2 else 0
    if V and int(v.shape[1]) != dim:
        raise IndexError("Point dimension mismatch!")
    if on_dev:
        import torch
        device = q.device.index
        idx = torch.empty((Q, k), dtype=torch.int32, device=q.device)
        dist = torch.empty((Q, k), dtype=torch.float64, device=q.device)
        cnt = torch.empty((Q,), dtype=torch.int32, device=q.device)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t.numel() else None
        stream = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
        flags = _abi.MEM_DEVICE
    else:
        idx = np.empty((Q, k), dtype=np.int32)
        dist = np.empty((Q, k))
        cnt = np.empty((Q,), dtype=np.int32)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        stream, flags = None, _abi.MEM_HOST
    rc = lib.rkb_nearest(int(device), V, ptr(v), Q, ptr(q), dim, int(k), float(radius), ptr(idx), ptr(dist), ptr(cnt), flags, stream)
    if rc != 0:
        if rc == _abi.ERR_CUDA:
            raise RuntimeError("rkb_nearest: " + lib.rkb_nearest_last_error().decode())
        _abi.check(rc, "rkb_nearest")
    return idx, dist, cnt
