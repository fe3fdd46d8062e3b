"""The interpreter kernels' DEVICE source (reak_b200/csrc/kte_generic.cu) compiled for the host by
tests/host_build/generic_host.cpp (test infrastructure; the product never loads it) and run on the program
rkb_chain_create lowered: chain walk, generalised forces, M / Mdot, twist-shaping matrices, frames and the integrators of
every preset — free joints and the _gen elements included — against the oracle, without a GPU.  (Arithmetic differs
from the device build only by the contraction of multiply-adds, hence the same tolerances as the GPU tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from reak_b200 import _abi, kte, presets

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
NAMES = sorted(presets.PRESETS) + sorted(presets.FREE_PRESETS)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    if not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found (types only are needed)")
    out = str(tmp_path_factory.mktemp("gen_host") / "libgen_host.so")
    src = os.path.join(HERE, "host_build", "generic_host.cpp")
    subprocess.run(["g++", "-std=c++14", "-O1", "-w", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-I" + CUDA_INC, "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.gen_host_eval.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.gen_host_rollout.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


class HostChain(object):
    def __init__(self, host, name):
        s = presets.make(name)
        self.c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, getattr(s, "dofs_3D", ()))
        lib = _abi.load_library()
        h = C.c_void_p()
        _abi.check(lib.rkb_chain_create_ex(C.byref(self.c.desc), _abi.CREATE_INTERPRETER, C.byref(h)), "rkb_chain_create_ex")
        size = host.gen_host_program_size()
        self.blob = C.create_string_buffer(size)
        assert lib.rkb_chain_program(h, self.blob, size) == size
        self.rows = lib.rkb_twist_shaping_rows(h)
        lib.rkb_chain_destroy(h)
        self.host = host

    def run(self, op, x, u, out_dim, second=False):
        N = x.shape[0]
        out = np.zeros((N, out_dim))
        out2 = np.zeros((N, out_dim)) if second else None
        st = np.zeros(N, dtype=np.int32)
        xc, uc = np.ascontiguousarray(x), np.ascontiguousarray(u)
        self.host.gen_host_eval(self.blob, op, N, xc.ctypes.data, self.c.nx, uc.ctypes.data if uc.size else None, self.c.n_inputs,
                                out.ctypes.data, out2.ctypes.data if second else None, out_dim, st.ctypes.data)
        return out, out2, st

    def rollout(self, x, u, dt, steps):
        N = x.shape[0]
        out = np.zeros((N, self.c.nx))
        st = np.zeros(N, dtype=np.int32)
        xc, uc = np.ascontiguousarray(x), np.ascontiguousarray(u)
        self.host.gen_host_rollout(self.blob, N, xc.ctypes.data, self.c.nx, uc.ctypes.data if uc.size else None, self.c.n_inputs, dt, steps, None,
                                   out.ctypes.data, st.ctypes.data)
        return out, st


def _states(c, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, (n, c.nx))
    x[:, 0:2 * c.n_coords:2] *= 2.0
    u = rng.uniform(-1.0, 1.0, (n, c.n_inputs))
    return x, u


@pytest.mark.parametrize("name", NAMES)
def test_interpreter_source_on_the_host_matches_the_oracle(name, host, oracle_built):
    H = HostChain(host, name)
    c = H.c
    O = oracle_built.Oracle(c)
    x, u = _states(c, 24, 5)
    xd, _, st = H.run(0, x, u, c.nx)
    xo, so = O.eval(x, u)
    assert not st.any() and not so.any() and rel_err(xd, xo) < 1e-10
    assert rel_err(H.run(1, x, u, c.n_acc)[0], O.gen_forces(x, u)) < 1e-10
    M, Md, _ = H.run(2, x, u, c.n_acc * c.n_acc, second=True)
    Mo, Mdo = O.mass(x)
    assert rel_err(M.reshape(Mo.shape), Mo) < 1e-10 and rel_err(Md.reshape(Mo.shape), Mdo) < 1e-10
    if H.rows:
        T, Td, _ = H.run(3, x[:1], u[:1], H.rows * c.n_acc, second=True)
        To, _, Tdo = O.tmt(x[0])
        assert rel_err(T.reshape(To.shape), To) < 1e-10 and rel_err(Td.reshape(To.shape), Tdo) < 1e-10
    fr, _, _ = H.run(4, x[:1], u[:1], 25 * c.desc.n_frames)
    assert rel_err(fr.reshape(-1, 25), O.frames(x[0], u[0])) < 1e-10
    xe, st = H.rollout(x, u, 1e-3, 20)
    xr, sr, _ = O.rk4(x, u, 1e-3, 20)
    assert not st.any() and not sr.any() and rel_err(xe, xr) < 1e-9
