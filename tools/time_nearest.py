"""Developer tool: rkb_nearest throughput on device-resident buffers (CUDA events on torch's current stream).
    python tools/time_nearest.py [dim] [V] [Q] [k]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reak_b200.nearest import nearest_neighbors  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 12
V = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g = torch.Generator(device="cuda").manual_seed(1)
v = torch.rand((V, dim), dtype=torch.float64, device="cuda", generator=g)
q = torch.rand((Q, dim), dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    nearest_neighbors(v, q, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    nearest_neighbors(v, q, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pairs = float(V) * Q
dimp = (dim + 3) // 4 * 4
print("nearest dim=%d V=%d Q=%d k=%d  %.3f ms  %.3e pairs/s  %.2f TFLOP/s (3 flop per coordinate and pair, padded dim %d: %.2f)"
      % (dim, V, Q, k, ms, pairs / ms * 1e3, pairs * 3 * dim / ms * 1e-9, dimp, pairs * 3 * dimp / ms * 1e-9))
