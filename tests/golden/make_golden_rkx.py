"""Regenerates tests/golden/rkx/: `.rkx` model files WRITTEN BY THE REFERENCE (ReaK::serialization::xml_oarchive over the
live kte_nl_system, oracle/ref_lib.cpp: rkref_save_rkx) and, next to each, the flat descriptor the reference side derives
from it — the file loaded back by ReaK's own xml_iarchive and flattened by include/reak_b200/reak_bridge.hpp
(rkref_load_rkx_desc).  The product's reader (rkb_rkx_read, no ReaK code) must reproduce those descriptors byte for byte.
Build container only:

    make -C oracle ref && python tests/golden/make_golden_rkx.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyref  # noqa: E402
from reak_b200 import _abi, kte, presets  # noqa: E402

CASES = ["crs6", "crs7_phys_sd", "planar3_sd", "planar2_lin_sd", "crs6_lin_sd", "crs3_gen", "free_arm3", "crs6_twist", "planar_pr"]


def save_rkx(name, path):
    s = presets.make(name)
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs, getattr(s, "dofs_3D", ()))
    R = pyref.Reference(c)
    fn = R.lib.rkref_save_rkx
    fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    if fn(R.h, path.encode(), 0) != 0:
        raise RuntimeError("rkref_save_rkx failed for " + name)
    return R


def reference_descriptor(R, path):
    """(header ints, base doubles, raw element bytes) of the descriptor ReaK + bridge derive from the file"""
    fn = R.lib.rkref_load_rkx_desc
    fn.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    out, elems, err = _abi.rkb_chain_desc(), (_abi.rkb_element * 256)(), C.create_string_buffer(256)
    n = fn(path.encode(), C.byref(out), elems, 256, err, 256)
    if n < 0:
        raise RuntimeError("rkref_load_rkx_desc: " + err.value.decode())
    return descriptor_arrays(out, elems, n)


def descriptor_arrays(d, elems, n):
    header = np.array([d.dim, d.n_elements, d.n_frames, d.n_coords, d.n_inputs, d.base_frame], dtype=np.int32)
    base = np.array(list(d.base.position) + list(d.base.quat) + list(d.base.velocity) + list(d.base.ang_velocity)
                    + list(d.base.acceleration) + list(d.base.ang_acceleration))
    raw = np.frombuffer(bytes(memoryview(elems))[:n * C.sizeof(_abi.rkb_element)], dtype=np.uint8).copy()
    return header, base, raw


def main():
    if not pyref.have_ref():
        raise SystemExit("oracle/_ref/libreak_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name in CASES:
        path = os.path.join(HERE, "rkx", name + ".rkx")
        R = save_rkx(name, path)
        header, base, raw = reference_descriptor(R, path)
        np.savez(os.path.join(HERE, "rkx", name + ".npz"), header=header, base=base, elements=raw)
        print("%-16s %6d bytes, %d elements" % (name, os.path.getsize(path), header[1]))


if __name__ == "__main__":
    main()
