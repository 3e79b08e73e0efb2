"""Shared test helpers: load golden fixtures, rebuild the same space with the oracle."""
import os

import numpy as np

from oracle import thb_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPACES = ["readme2d", "block2d", "twoblock2d", "graded2d", "block3d", "readme3d"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"space_{name}.npz")))


def oracle_space_from_golden(g):
    d = int(g["ndim"])
    nlev = int(g["num_levels"])
    sp = O.Space([g[f"knots0_{k}"] for k in range(d)], tuple(int(p) for p in g["degrees"]), nlev)
    sp.build_hierarchy_from_domain_sequence()
    for row in g["boxes"]:
        lev = int(row[0])
        lo, hi = row[1:1 + d], row[1 + d:1 + 2 * d]
        sp.refine_box([int(x) for x in lo], [int(x) for x in hi], lev)
    sp.build_hierarchy_from_domain_sequence()
    return sp


def rel_err(a, b, scale=1.0):
    """|a-b| / max(|b|, scale) -- the tolerance metric of SURVEY.md hard part 6."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), scale)
