#!/usr/bin/env python
"""Golden vectors for the ground-truth generator, produced by RUNNING THE REAL REFERENCE
``simulations.ground_truths.gaussian_random_field`` (ground_truths.py:14-33) in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_grf.py

For each case the white-noise draw (``np.random.normal(size=(y_dim, x_dim))`` under a fixed seed) and the field
the reference makes from it are stored; the host generator of the product (ipp_rl_b200/simulations/ground_truths.py)
is checked against them while generating."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402,F401  (reference import path + stub modules)

import numpy as np  # noqa: E402
from simulations import ground_truths as ref_gt  # noqa: E402  (reference)

from ipp_rl_b200.simulations import ground_truths as our_gt  # noqa: E402

CASES = [("s50", 50, 50, 5.0, 11), ("s200", 200, 200, 5.0, 12), ("odd", 31, 33, 3.0, 13), ("rect", 48, 20, 4.0, 14)]


def main():
    out = {}
    worst = 0.0
    for name, X, Y, r, seed in CASES:
        np.random.seed(seed)
        white = np.random.normal(size=(Y, X))
        np.random.seed(seed)
        field = ref_gt.gaussian_random_field(lambda k: k ** (-r), X, Y)
        np.random.seed(seed)
        ours = our_gt.gaussian_random_field(lambda k: k ** (-r), X, Y)
        worst = max(worst, float(np.max(np.abs(ours - field))))
        out[f"{name}_white"] = white.astype(np.float32)
        out[f"{name}_field"] = field
        out[f"{name}_dims"] = np.array([X, Y])
        out[f"{name}_radius"] = np.array(r)
        out[f"{name}_seed"] = np.array(seed)
    out["names"] = np.array([c[0] for c in CASES])
    np.savez_compressed(os.path.join(HERE, "golden_grf.npz"), **out)
    print("host generator vs reference: worst |err|", worst)


if __name__ == "__main__":
    main()
