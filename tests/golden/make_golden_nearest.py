"""Regenerates tests/golden/nearest/nearest.npz from the REAL reference: ReaK::pp::min_dist_linear_search
(ctrl/path_planning/topological_search.hpp:91-112, 238-270) compiled into oracle/_ref/libreak_ref.so (rkref_nearest).
Build container only:   make -C oracle ref && python tests/golden/make_golden_nearest.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyref  # noqa: E402

CASES = [(12, 1, np.inf), (12, 5, np.inf), (6, 3, 0.7), (19, 16, np.inf), (2, 4, 0.05), (13, 8, 1.7)]


def main():
    if not pyref.have_ref():
        raise SystemExit("oracle/_ref/libreak_ref.so missing")
    out = {}
    for n, (dim, k, radius) in enumerate(CASES):
        rng = np.random.default_rng(500 + n)
        v = rng.uniform(-1.0, 1.0, (700, dim))
        q = rng.uniform(-1.0, 1.0, (24, dim))
        idx, dist, cnt = pyref.nearest("ref", v, q, k, radius)
        out.update({"v%d" % n: v, "q%d" % n: q, "idx%d" % n: idx, "dist%d" % n: dist, "cnt%d" % n: cnt,
                    "par%d" % n: np.array([dim, k, radius])})
        print(dim, k, radius, "neighbours found:", int(cnt.sum()))
    np.savez(os.path.join(HERE, "nearest", "nearest.npz"), n_cases=len(CASES), **out)


if __name__ == "__main__":
    main()
