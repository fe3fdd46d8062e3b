"""SURVEY f4, model I/O: ReaK XML archives (`.rkx`, core/serialization/xml_archiver.cpp; the model files of
examples/robot_airship/build_P3R3R_model.cpp:78) read by the library's own reader (rkb_rkx_read / rkb_rkx_load,
reak_b200/csrc/rkb_rkx.cu — no ReaK code).

The fixtures under tests/golden/rkx/ were WRITTEN by the reference (xml_oarchive over the live kte_nl_system) and each
comes with the descriptor the reference side derives from it (ReaK's xml_iarchive + reak_bridge.hpp); see
tests/golden/make_golden_rkx.py.  CPU: the reader reproduces those descriptors byte for byte, and — where the compiled
reference is present — does so for a freshly written file of every preset.  GPU: a chain loaded from a file evaluates
like the chain it was written from."""
import ctypes as C
import glob
import importlib.util
import os

import numpy as np
import pytest

from conftest import random_batch, rel_err
from reak_b200 import _abi, kte, presets

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "rkx", "*.rkx")))


def _tools():
    spec = importlib.util.spec_from_file_location("make_golden_rkx", os.path.join(HERE, "golden", "make_golden_rkx.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _read(path):
    lib = _abi.load_library()
    d, elems, err = _abi.rkb_chain_desc(), (_abi.rkb_element * 256)(), C.create_string_buffer(256)
    n = lib.rkb_rkx_read(os.fsencode(path), C.byref(d), elems, 256, err, 256)
    assert n >= 0, err.value
    return d, elems, n


def test_fixtures_present():
    assert len(FILES) >= 9 and all(os.path.isfile(f[:-4] + ".npz") for f in FILES)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_reader_reproduces_the_reference_side_descriptor(path):
    g = np.load(path[:-4] + ".npz")
    d, elems, n = _read(path)
    header, base, raw = _tools().descriptor_arrays(d, elems, n)
    assert np.array_equal(header, g["header"])
    assert np.array_equal(base, g["base"])           # exact: both sides parse the same decimal strings
    assert np.array_equal(raw, g["elements"])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_file_descriptor_is_the_preset_to_six_digits(path):
    """the archiver prints 6 significant digits (xml_archiver.cpp): structure identical, parameters to 1e-5 relative"""
    name = os.path.basename(path)[:-4]
    s = presets.make(name)
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, getattr(s, "dofs_3D", ()))
    f = kte.read_rkx(path)
    assert (f.desc.dim, f.desc.n_elements, f.desc.n_frames, f.desc.n_coords, f.desc.n_inputs, f.desc.base_frame, f.nx) == \
           (c.desc.dim, c.desc.n_elements, c.desc.n_frames, c.desc.n_coords, c.desc.n_inputs, c.desc.base_frame, c.nx)
    for a, b in zip(f.elements, c.elements):
        assert (a.kind, a.frame_a, a.frame_b, a.coord, a.aux, a.upstream) == (b.kind, b.frame_a, b.frame_b, b.coord, b.aux, b.upstream)
        assert np.allclose(list(a.p), list(b.p), rtol=1e-5, atol=1e-12)


def test_every_preset_round_trips_through_a_file_the_reference_writes(oracle_built, tmp_path):
    if not oracle_built.have_ref():
        pytest.skip("compiled reference not built here")
    T = _tools()
    for name in sorted(presets.PRESETS) + sorted(presets.FREE_PRESETS):
        path = str(tmp_path / (name + ".rkx"))
        R = T.save_rkx(name, path)
        header, base, raw = T.reference_descriptor(R, path)
        d, elems, n = _read(path)
        h2, b2, r2 = T.descriptor_arrays(d, elems, n)
        assert np.array_equal(header, h2) and np.array_equal(base, b2) and np.array_equal(raw, r2), name


def test_reader_errors(tmp_path):
    lib = _abi.load_library()
    d, err = _abi.rkb_chain_desc(), C.create_string_buffer(256)
    assert lib.rkb_rkx_read(b"/nonexistent/model.rkx", C.byref(d), None, 0, err, 256) == _abi.ERR_INVALID
    bad = tmp_path / "bad.rkx"
    bad.write_text("<?xml version=\"1.0\"?>\n<something_else></something_else>\n")
    assert lib.rkb_rkx_read(os.fsencode(str(bad)), C.byref(d), None, 0, err, 256) == _abi.ERR_UNSUPPORTED and b"reak_serialization" in err.value
    # an archive whose first object is not a kte_nl_system
    src = open(FILES[0]).read().replace('type_ID="3257925634.0"', 'type_ID="12345.0"', 1)
    other = tmp_path / "other.rkx"
    other.write_text(src)
    assert lib.rkb_rkx_read(os.fsencode(str(other)), C.byref(d), None, 0, err, 256) == _abi.ERR_UNSUPPORTED and b"kte_nl_system" in err.value
    # truncated file
    cut = tmp_path / "cut.rkx"
    cut.write_text(open(FILES[0]).read()[:5000])
    assert lib.rkb_rkx_read(os.fsencode(str(cut)), C.byref(d), None, 0, err, 256) == _abi.ERR_UNSUPPORTED
    # element count only / buffer too small
    n = lib.rkb_rkx_read(os.fsencode(FILES[0]), C.byref(d), None, 0, err, 256)
    assert n == d.n_elements > 0
    small = (_abi.rkb_element * 2)()
    assert lib.rkb_rkx_read(os.fsencode(FILES[0]), C.byref(d), small, 2, err, 256) == _abi.ERR_NOMEM
    with pytest.raises(kte.UnsupportedChain):
        kte.read_rkx(str(bad))


def test_oracle_on_file_descriptor_matches_oracle_on_preset(oracle_built):
    """dynamics of the chain in the file = dynamics of the chain it was written from, to the file's 6 digits"""
    for name in ("crs6", "free_arm3", "planar3_sd"):
        path = os.path.join(HERE, "golden", "rkx", name + ".rkx")
        s = presets.make(name)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, getattr(s, "dofs_3D", ()))
        f = kte.read_rkx(path)
        rng = np.random.default_rng(3)
        x = rng.uniform(-1.0, 1.0, (32, c.nx))
        u = rng.uniform(-1.0, 1.0, (32, c.n_inputs))
        a = oracle_built.Oracle(f).eval(x, u)[0]
        b = oracle_built.Oracle(c).eval(x, u)[0]
        assert rel_err(a, b) < 1e-4, name


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_chain_loaded_from_file(path, oracle_built):
    """rkb_rkx_load -> handle -> the kernels; checked against the oracle on the SAME descriptor (1e-10) and against the
    preset the file was written from (6 digits)"""
    from reak_b200 import kte_batch_propagator
    name = os.path.basename(path)[:-4]
    p = kte_batch_propagator.from_rkx(path)
    q = kte_batch_propagator(presets.make(name))
    assert (p.nx, p.nu, p.na, p.is_serial()) == (q.nx, q.nu, q.na, q.is_serial())
    rng = np.random.default_rng(8)
    x = rng.uniform(-1.0, 1.0, (200, p.nx))
    u = rng.uniform(-1.0, 1.0, (200, p.nu))
    O = oracle_built.Oracle(p.compiled)
    xd, st = p.get_state_derivatives(x, u)
    assert not st.any() and rel_err(xd, O.eval(x, u)[0]) < 1e-10
    xo, st = p.get_next_states(x, u, 1e-3, 50)
    assert not st.any() and rel_err(xo, O.rk4(x, u, 1e-3, 50)[0]) < 1e-8
    assert rel_err(xo, q.get_next_states(x, u, 1e-3, 50)[0]) < 1e-4
    # the one-call form of the C-ABI
    lib = _abi.load_library()
    h, err = C.c_void_p(), C.create_string_buffer(256)
    assert lib.rkb_rkx_load(os.fsencode(path), 0, C.byref(h), err, 256) == 0, err.value
    assert lib.rkb_chain_state_dim(h) == p.nx
    lib.rkb_chain_destroy(h)
