"""CPU tests of the drop-in boundary: libreak_b200.so loads, exports every symbol include/reak_b200.h
declares, validates/lower descriptors on the host, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from reak_b200 import _abi, kte, presets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "reak_b200.h")).read()
    return sorted(set(re.findall(r"RKB_API[^;(]*?\b(rkb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
        assert n in _abi.SYMBOLS, "ctypes table misses %s" % n
    assert lib.rkb_version() == 120
    assert lib.rkb_strerror(0) == b"ok" and b"CPU fallback" in lib.rkb_strerror(_abi.ERR_CUDA)


def test_struct_layout_matches_header():
    assert C.sizeof(_abi.rkb_element) == 128
    assert C.sizeof(_abi.rkb_base_frame) == 19 * 8
    assert _abi.rkb_chain_desc.elements.offset == 24 + 19 * 8
    assert C.sizeof(_abi.rkb_rollout_opts) == 24 and _abi.rkb_rollout_opts.dt.offset == 16
    assert C.sizeof(_abi.rkb_steer_opts) == 72 and _abi.rkb_steer_opts.u_lower.offset == 40


def _create(desc):
    lib = _abi.load_library()
    h = C.c_void_p()
    rc = lib.rkb_chain_create(C.byref(desc), C.byref(h))
    return rc, h


# planar chains are embedded in the x-y plane and run on the same kernels; two-anchor springs / dampers need the interpreter
SERIAL = {"crs6", "crs6_phys", "crs6_sd", "crs6_sd_sat", "crs6_twist", "crs7", "crs7_phys_sd", "torsion1", "crs3", "crs6_passive",
          "pendulum", "planar2", "planar2_act", "planar3_sd", "planar_pr", "crs2d"}


@pytest.mark.parametrize("name", sorted(presets.PRESETS))
def test_chain_create_and_lowering(name):
    lib = _abi.load_library()
    s = presets.make(name)
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    rc, h = _create(c.desc)
    assert rc == 0
    assert lib.rkb_chain_dof(h) == c.n_coords
    assert lib.rkb_chain_state_dim(h) == 2 * c.n_coords
    assert lib.rkb_chain_input_dim(h) == c.n_inputs
    assert lib.rkb_chain_is_serial(h) == (1 if name in SERIAL else 0)
    lib.rkb_chain_destroy(h)


def test_malformed_descriptors_are_rejected():
    lib = _abi.load_library()
    s = presets.make("crs3")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    assert lib.rkb_chain_create(None, C.byref(C.c_void_p())) == _abi.ERR_INVALID
    saved = c.desc.n_frames
    c.desc.n_frames = 2  # frame ids out of range
    assert _create(c.desc)[0] == _abi.ERR_INVALID
    c.desc.n_frames = saved
    kind = c.elements[3].kind
    c.elements[3].kind = 99
    assert _create(c.desc)[0] == _abi.ERR_INVALID
    c.elements[3].kind = _abi.FREE_3D  # free joints are reserved, not compiled
    assert _create(c.desc)[0] == _abi.ERR_UNSUPPORTED
    c.elements[3].kind = kind
    c.elements[2].p[0] = float("nan")
    assert _create(c.desc)[0] == _abi.ERR_INVALID
    c.elements[2].p[0] = 0.0
    c.desc.n_inputs = c.desc.n_inputs + 1  # an input nobody drives
    assert _create(c.desc)[0] == _abi.ERR_INVALID
    c.desc.n_inputs -= 1
    assert _create(c.desc)[0] == 0


def test_argument_errors_without_touching_the_gpu():
    lib = _abi.load_library()
    s = presets.make("crs3")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    rc, h = _create(c.desc)
    x = np.zeros((4, 6)); u = np.zeros((4, 3)); o = np.zeros((4, 6))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.rkb_rollout_rk4(h, 0, 4, p(x), p(u), 0.0, 3, p(o), None, 0, None) == _abi.ERR_INTEGRATION
    assert lib.rkb_rollout_rk4(h, 0, 4, p(x), p(u), 1e-3, -1, p(o), None, 0, None) == _abi.ERR_INTEGRATION
    assert lib.rkb_rollout_rk4(h, 0, 4, None, p(u), 1e-3, 1, p(o), None, 0, None) == _abi.ERR_INVALID
    assert lib.rkb_rollout_rk4(h, 0, 0, None, None, 1e-3, 1, None, None, 0, None) == 0  # empty batch is a no-op
    assert lib.rkb_eval(None, 0, 4, p(x), p(u), p(o), None, 0, None) == _abi.ERR_INVALID
    assert lib.rkb_launch_count(h) == 0 and lib.rkb_last_kernel_ms(h) < 0
    lib.rkb_chain_destroy(h)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from reak_b200 import kte_batch_propagator
    p = kte_batch_propagator(presets.make("crs6"))
    with pytest.raises(_abi.RkbError) as e:
        p.get_state_derivatives(np.zeros((2, 12)), np.zeros((2, 6)))
    assert e.value.code == _abi.ERR_CUDA


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "reak_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) :
                text = open(os.path.join(dirpath, f)).read()
                for line in text.splitlines():
                    code = line.split("#")[0].split("//")[0]
                    assert "oracle" not in code.lower() or "import" not in code and "include" not in code, (f, line)


def _build_c_demo(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + os.path.dirname(_abi.LIB_PATH), "-lreak_b200", "-lm",
                           "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_abi.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_plain_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/c_abi_demo.c is C99 against include/reak_b200.h only.  Without a CUDA device the first
    compute call must come back as RKB_ERR_CUDA: there is no CPU fallback behind the C-ABI."""
    import subprocess
    import torch
    exe, env = _build_c_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the demo is run by tests/test_gpu_parity.py")
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "-> -4" in r.stderr, (r.returncode, r.stdout, r.stderr)


def test_proximity_demo_is_c99_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/proximity_demo.c (planar and spatial proximity models, rkb_proxy_specialize, rkb_is_free) is C99 against
    include/reak_b200.h only, -Wall -Wextra -Werror; without a CUDA device the first query returns RKB_ERR_CUDA"""
    import subprocess
    import torch
    exe = os.path.join(str(tmp_path), "proximity_demo")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "proximity_demo.c"), "-L" + os.path.dirname(_abi.LIB_PATH), "-lreak_b200", "-lm",
                           "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_abi.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 1 and "rkb_min_distance" in r.stderr and "-> -4" in r.stderr, (r.returncode, r.stdout, r.stderr)


def _build_cpp_demo(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "cpp_host_demo")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cpp_host_demo.cpp"), "-L" + os.path.dirname(_abi.LIB_PATH), "-lreak_b200",
                           "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_abi.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_cpp_host_class_is_standalone_cxx11(tmp_path):
    """include/reak_b200/kte_batch_propagator.hpp needs neither ReaK nor Boost nor CUDA headers: C++11, -pedantic -Werror.
    Without a GPU the propagator raises propagator_error (RKB_ERR_CUDA)."""
    import subprocess
    import torch
    exe, env = _build_cpp_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the demo is run by tests/test_gpu_parity.py")
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def test_specialize_needs_a_gpu_and_a_serial_chain():
    lib = _abi.load_library()
    for name, want in (("crs6", _abi.ERR_CUDA), ("planar2_lin_sd", _abi.ERR_UNSUPPORTED)):
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        rc, h = _create(c.desc)
        assert rc == 0 and lib.rkb_chain_is_specialized(h) == 0
        import torch
        if not torch.cuda.is_available() or want == _abi.ERR_UNSUPPORTED:
            assert lib.rkb_chain_specialize(h, 0) == want
        lib.rkb_chain_destroy(h)


def test_proxy_specialize_error_paths():
    """rkb_proxy_specialize / rkb_steer_checked_specialize fail loudly, never silently: no GPU -> RKB_ERR_CUDA (there is no CPU
    path), an interpreter chain has no steering kernel to build a collision test into -> RKB_ERR_UNSUPPORTED, options are
    range-checked, NULL handles are refused"""
    import ctypes as C
    import torch
    from reak_b200 import proximity as px
    lib = _abi.load_library()
    s = presets.make("crs6")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    rc, h = _create(c.desc)
    assert rc == 0
    robot, lab = presets.crs_proxy_models(s)
    pair = px.ProxyHandle(lib, h, px.proxy_query_pair_3D("robot-lab", robot, lab), c.frames)
    assert lib.rkb_proxy_is_specialized(pair._h) == 0
    assert lib.rkb_proxy_set_option(pair._h, px.ProxyHandle.OPT_MIN_BLOCKS, 0) == _abi.ERR_INVALID
    assert lib.rkb_proxy_set_option(pair._h, px.ProxyHandle.OPT_MIN_BLOCKS, 9) == _abi.ERR_INVALID
    assert lib.rkb_proxy_set_option(pair._h, px.ProxyHandle.OPT_MIN_BLOCKS, 5) == 0
    assert lib.rkb_proxy_set_option(pair._h, 77, 1) == _abi.ERR_INVALID
    assert lib.rkb_proxy_set_option(None, px.ProxyHandle.OPT_AUTO_SPECIALIZE, 0) == _abi.ERR_INVALID
    assert lib.rkb_proxy_specialize(None, 0) == _abi.ERR_INVALID
    arr = (C.c_void_p * 1)(pair._h)
    assert lib.rkb_steer_checked_specialize(h, 0, None, 1) == _abi.ERR_INVALID
    assert lib.rkb_steer_checked_specialize(h, 0, arr, 0) == _abi.ERR_INVALID
    assert lib.rkb_steer_checked_is_specialized(h, arr, 1) == 0
    if not torch.cuda.is_available():
        assert lib.rkb_proxy_specialize(pair._h, 0) == _abi.ERR_CUDA
        assert lib.rkb_steer_checked_specialize(h, 0, arr, 1) == _abi.ERR_CUDA
        assert lib.rkb_proxy_is_specialized(pair._h) == 0
    n = lib.rkb_proxy_source(pair._h, None, 0)
    assert n > 1000
    small = C.create_string_buffer(16)
    assert lib.rkb_proxy_source(pair._h, small, 16) == _abi.ERR_INVALID       # too small a buffer is refused, not overrun
    pair.close()
    lib.rkb_chain_destroy(h)
    # an interpreter chain (two-anchor linear springs): no serial steering kernel
    s2 = presets.make("crs6_lin_sd")
    c2 = kte.compile_chain(s2.chain, s2.mass_calc, s2.dofs_gen, s2.inputs)
    rc, h2 = _create(c2.desc)
    assert rc == 0 and lib.rkb_chain_is_serial(h2) == 0
    robot2, lab2 = presets.crs_proxy_models(s2)
    pair2 = px.ProxyHandle(lib, h2, px.proxy_query_pair_3D("robot-lab", robot2, lab2), c2.frames)
    arr2 = (C.c_void_p * 1)(pair2._h)
    assert lib.rkb_steer_checked_source(h2, arr2, 1, None, 0) == _abi.ERR_UNSUPPORTED
    assert lib.rkb_proxy_source(pair2._h, None, 0) > 1000                      # (the query itself can still be generated)
    pair2.close()
    lib.rkb_chain_destroy(h2)
