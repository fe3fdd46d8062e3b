"""Generates tests/golden/proximity/proximity_2d.npz from the UNMODIFIED reference (oracle/_ref/libreak_ref.so):
proxy_query_pair_2D::findMinimumDistance (geometry/proximity/proxy_query_model.cpp:163-190) for two random planar models
holding every planar shape kind (circle, capped_rectangle, rectangle), riding on a 3-link revolute arm and on the
prismatic + revolute planar CRS analog, at seeded random states.  Run from the repo root:

    python tests/golden/proximity/make_golden_proximity_2d.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import random_batch  # noqa: E402
from oracle import pyref  # noqa: E402
from reak_b200 import kte, presets  # noqa: E402
from test_proximity import mixed_models2  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_proximity import shape_rows  # noqa: E402


def main():
    out = {}
    for tag, preset, n in (("arm3", "planar3_sd", 128), ("crs2d", "crs2d", 96)):
        s = presets.make(preset)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        pair = mixed_models2(c, np.random.default_rng(7))
        R = pyref.Reference(c)
        x, _ = random_batch(c, n, seed=78, q_range=2.5)
        d, f, p = R.min_distance(pair, x)
        out[tag + "_x"] = x
        out[tag + "_frames"] = np.stack([R.frames(x[i:i + 1])[:, :7] for i in range(n)])
        out[tag + "_shapes1"] = shape_rows(pair.model1, c.frames)
        out[tag + "_shapes2"] = shape_rows(pair.model2, c.frames)
        out[tag + "_distance"], out[tag + "_finder"], out[tag + "_points"] = d, f, p
        print(tag, "finders used:", sorted(set(f.tolist())), "colliding:", int((d < 0).sum()), "of", n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "proximity", "proximity_2d.npz"), **out)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
