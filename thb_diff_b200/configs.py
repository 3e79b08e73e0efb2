"""Synthetic spaces for the BASELINE.json configurations."""
import numpy as np

from .multilevel_spline_space import BSpline, Space, TensorProduct

README_U0 = np.array([0, 0, 0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1, 1, 1], dtype=np.float64)   # p=2
README_U1 = np.array([0, 0, 0, 0, 0.2, 0.4, 0.5, 0.6, 0.8, 0.9, 1, 1, 1, 1], dtype=np.float64)           # p=3
README_U2 = np.array([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1, 1, 1], dtype=np.float64)                           # p=2


def config1_space():
    """C1: README 2-D space (degrees 2x3), 3 levels, level-0 cells (2,2) and (3,1) refined."""
    sp = Space(TensorProduct([BSpline(README_U0, 2), BSpline(README_U1, 3)]), 3)
    sp._refine_cell((2, 2), 0)
    sp._refine_cell((3, 1), 0)
    sp.build_hierarchy_from_domain_sequence()
    return sp


def config2_space():
    """C2: README 3-D space (degrees 2/3/2), 3 levels, level-0 cells (2,2,2) and (3,1,0) refined."""
    sp = Space(TensorProduct([BSpline(README_U0, 2), BSpline(README_U1, 3), BSpline(README_U2, 2)]), 3)
    sp._refine_cell((2, 2, 2), 0)
    sp._refine_cell((3, 1, 0), 0)
    sp.build_hierarchy_from_domain_sequence()
    return sp


def uniformish_knots(ncells, p, rng, jitter=0.1):
    """open knot vector with `ncells` spans whose interior breakpoints are jittered by +-jitter*h"""
    h = 1.0 / ncells
    brk = np.arange(ncells + 1) * h
    brk[1:-1] += rng.uniform(-jitter, jitter, ncells - 1) * h
    return np.concatenate([np.zeros(p), brk, np.ones(p)])


# nested, centred, graded refinement boxes (in cells of the level that is refined); ~25 % of the cells of
# each region are refined and every box keeps >= 4 cells of margin inside its parent.
_BOXES = [(0, 3, 13), (1, 10, 22), (2, 24, 40), (3, 54, 74)]


def cubic3d_space(active_levels=4, base_cells=16, ndim=3, degree=3, seed=0):
    """C3 (active_levels=4) / C5 (active_levels=5): degree-3, 'uniform-ish' knots, nested centred boxes."""
    rng = np.random.default_rng(seed)
    bs = [BSpline(uniformish_knots(base_cells, degree, rng), degree) for _ in range(ndim)]
    sp = Space(TensorProduct(bs), active_levels)
    scale = base_cells / 16.0
    for lev, lo, hi in _BOXES[: active_levels - 1]:
        a, b = int(round(lo * scale)), int(round(hi * scale))
        sp.refine_box((a,) * ndim, (b,) * ndim, lev)
    sp.build_hierarchy_from_domain_sequence()
    return sp


def greville_ctrl_pts(sp, bump=0.1):
    """Level-major control points [nCP, 3]: Greville abscissae (so the map reproduces the parameters)
    plus a smooth bump on the last coordinate -- the synthetic geometry of SURVEY 8d."""
    from .bspline_funcs import grevilleAbscissae

    rows = []
    for lev in range(sp.num_levels):
        g = grevilleAbscissae(sp.sh_fns[lev], sp.degrees, sp.knotvectors[lev]).reshape(-1, sp.ndim)
        if sp.ndim == 2:
            g = np.concatenate([g, np.zeros((len(g), 1))], axis=1)
        if bump:
            g = g.copy()
            g[:, 2] += bump * np.sin(2 * np.pi * g[:, 0]) * np.sin(2 * np.pi * g[:, 1])
        rows.append(g)
    return np.concatenate(rows).astype(np.float32)
