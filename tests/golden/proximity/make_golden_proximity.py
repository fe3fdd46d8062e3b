"""Generates tests/golden/proximity/proximity.npz from the UNMODIFIED reference (oracle/_ref/libreak_ref.so):
proxy_query_pair_3D::findMinimumDistance for (a) the CRS arm's proximity model against the MD148 lab and
(b) two random models holding every shape kind, at seeded random states.  Run from the repo root:

    python tests/golden/proximity/make_golden_proximity.py
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
from reak_b200 import proximity as px  # noqa: E402
from test_proximity import mixed_models  # noqa: E402


def shape_rows(model, frames):
    rows = []
    for s in model.mShapeList:
        c = s.to_c(frames)
        rows.append([c.kind, c.anchor] + list(c.position) + list(c.quat) + list(c.dims))
    return np.array(rows, dtype=np.float64)


def main():
    out = {}
    for tag, preset, n, q_range in (("crs_lab", "crs6", 192, 3.1), ("mixed", "crs7", 96, 2.0)):
        s = presets.make(preset)
        c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
        if tag == "crs_lab":
            robot, lab = presets.crs_proxy_models(s)
            pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
        else:
            pair = mixed_models(s, np.random.default_rng(2024))
        R = pyref.Reference(c)
        x, _ = random_batch(c, n, seed=77, q_range=q_range)
        d, f, p = R.min_distance(pair, x)
        frames = np.stack([R.frames(x[i:i + 1])[:, :7] for i in range(n)])
        out[tag + "_x"] = x
        out[tag + "_frames"] = frames
        out[tag + "_shapes1"] = shape_rows(pair.model1, c.frames)
        out[tag + "_shapes2"] = shape_rows(pair.model2, c.frames)
        out[tag + "_distance"], out[tag + "_finder"], out[tag + "_points"] = d, f, p
        print(tag, "finders used:", sorted(set(f.tolist())), "colliding:", int((d < 0).sum()), "of", n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "proximity", "proximity.npz"), **out)


if __name__ == "__main__":
    main()
