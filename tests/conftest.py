import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C++ drop-in checks inside oracle/_ref call into the product library through weak references: its
    # symbols must be visible before libreak_ref.so is mapped (oracle/pyref.py: preload_product)
    from oracle import pyref
    pyref.preload_product()


def rel_err(a, b):
    """max |a-b| / max(1, |b|) — the per-component relative error BASELINE.json's tolerance is stated in."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def random_batch(compiled, n_samples, seed, q_range=1.0, qd_range=1.0, u_range=1.0):
    rng = np.random.default_rng(seed)
    n, nu = compiled.n_coords, compiled.n_inputs
    x = np.empty((n_samples, 2 * n))
    x[:, 0::2] = rng.uniform(-q_range, q_range, (n_samples, n))
    x[:, 1::2] = rng.uniform(-qd_range, qd_range, (n_samples, n))
    u = rng.uniform(-u_range, u_range, (n_samples, nu))
    return x, u


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import pyref
    if not os.path.isfile(pyref.ORACLE_SO):
        pyref.build(("oracle",))
    return pyref
