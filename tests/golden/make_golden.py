"""Regenerates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (see _refshim.py for how it is
loaded without jax) plus the reference's compiled compute_pts.cpp (oracle/_ref).

    python tests/golden/make_golden.py        # only works in the build container (/root/reference)

What is recorded, per space configuration (all float64, fp16 table quantisation disabled unless the
file name says fp16):
  knot vectors, degrees, refinement recipe  -> cells / fns activity masks, Coeff matrices, sh_fns
  params -> (level, cell) per point, num_supp, Jm (flat support ids), PHI of compute_THB_fns_tp
  a checksum of every dense fn_coeffs table (and the full tables for the small 2-D space)
Derivative goldens (not produced by reference code, whose jacfwd path needs JAX) come from
scipy.interpolate.BSpline and are stored in kat.npz together with the span table of SURVEY Q10.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

U_A = np.array([0, 0, 0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1, 1, 1])       # p=2 (README.md:67)
U_B = np.array([0, 0, 0, 0, 0.2, 0.4, 0.5, 0.6, 0.8, 0.9, 1, 1, 1, 1])               # p=3 (README.md:68)
U_C = np.array([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1, 1, 1])                               # p=2 (README.md:81)
U_D = np.array([0, 0, 0, 0, 0.25, 0.5, 0.75, 1, 1, 1, 1])                            # p=3

CONFIGS = {
    # name: (knots, degrees, num_levels, [(level, lo, hi) boxes refined in order], grid shape, extra points)
    "readme2d": ((U_A, U_B), (2, 3), 3, [(0, (2, 2), (3, 3)), (0, (3, 1), (4, 2))], (20, 20), []),
    "block2d": ((U_A, U_B), (2, 3), 3, [(0, (2, 1), (6, 5))], (17, 13),
                [(0.25, 0.35), (0.2, 0.4), (0.55, 0.75), (0.05, 0.05), (0.0, 0.0), (0.999, 0.001)]),
    "twoblock2d": ((U_A, U_B), (2, 3), 3, [(0, (2, 1), (6, 5)), (1, (5, 3), (10, 8))], (15, 11), []),
    "graded2d": ((U_A, U_B), (2, 3), 4, [(0, (1, 0), (9, 7)), (1, (6, 4), (14, 10))], (19, 14), []),
    "block3d": ((U_C, U_D, U_C), (2, 3, 2), 3, [(0, (0, 0, 1), (4, 4, 5))], (7, 6, 5), [(0.3, 0.3, 0.5)]),
    "readme3d": ((U_A, U_B, U_C), (2, 3, 2), 2, [(0, (2, 2, 2), (3, 3, 3)), (0, (3, 1, 0), (4, 2, 1))], (5, 4, 3), []),
}


def run_config(R, name, spec):
    from itertools import product

    knots, degrees, nlev, boxes, grid, extra = spec
    S, C, B = R.space, R.core, R.bspline_funcs
    tp = S.TensorProduct([S.BSpline(np.array(U, dtype=np.float64), p) for U, p in zip(knots, degrees)])
    h = S.Space(tp, nlev)
    h.build_hierarchy_from_domain_sequence()
    for lev, lo, hi in boxes:
        for c in product(*[range(a, b) for a, b in zip(lo, hi)]):
            h._refine_cell(tuple(c), lev)
    h.build_hierarchy_from_domain_sequence()

    params = B.generate_parametric_coordinates(grid)
    if extra:
        params = np.vstack([params, np.array(extra, dtype=np.float64)])
    # the kernel's canonical inputs are float32: make params f32-representable
    params = params.astype(np.float32).astype(np.float64)

    ac = C.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    fc = C.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    spans, num_supp, _ = C.compute_active_span_v2(params, h.knotvectors, h.cells, h.degrees, ac, fc)
    supp_list, num_supp2 = C.compute_active_span(params, h.knotvectors, h.cells, h.degrees, ac)
    assert list(num_supp) == list(num_supp2)
    PHI, ns3 = C.compute_THB_fns_tp(params, spans, ac, fc, h.sh_fns, h.knotvectors, h.degrees)
    assert list(ns3) == list(num_supp)

    off = np.zeros(nlev + 1, dtype=np.int64)
    for lev in range(nlev):
        off[lev + 1] = off[lev] + int(np.prod(h.sh_fns[lev]))
    Jm = np.array([off[fl] + np.ravel_multi_index(fn, h.sh_fns[fl]) for s in supp_list for fl, fn in s], dtype=np.int64)

    d = len(degrees)
    out = dict(
        degrees=np.array(degrees), num_levels=np.array(nlev), ndim=np.array(d),
        params=params, span_level=np.array([s[0] for s in spans]),
        span_cell=np.array([[int(x) for x in s[1]] for s in spans]),
        num_supp=np.array(num_supp, dtype=np.int64), Jm=Jm, PHI=np.asarray(PHI, dtype=np.float64),
        boxes=np.array([[lev, *lo, *hi] for lev, lo, hi in boxes]),
    )
    for k in range(d):
        out[f"knots0_{k}"] = np.asarray(knots[k], dtype=np.float64)
    for lev in range(nlev):
        out[f"cells_{lev}"] = h.cells[lev].astype(np.uint8)
        out[f"fns_{lev}"] = h.fns[lev].astype(np.uint8)
        out[f"sh_fns_{lev}"] = np.array(h.sh_fns[lev])
        for k in range(d):
            out[f"knots_{lev}_{k}"] = np.asarray(h.knotvectors[lev][k], dtype=np.float64)
        if lev < nlev - 1:
            for k in range(d):
                out[f"coeff_{lev}_{k}"] = np.asarray(h.Coeff[lev][k], dtype=np.float64)
        t = np.asarray(fc[lev], dtype=np.float64)
        w = np.cos(np.arange(t.size, dtype=np.float64) * 0.37).reshape(t.shape)
        out[f"fc_checksum_{lev}"] = np.array([t.sum(), (t * w).sum(), float(np.count_nonzero(t))])
        if d == 2 and name in ("block2d",):
            out[f"fc_{lev}"] = t.astype(np.float32)
    pou = np.add.reduceat(out["PHI"], np.concatenate([[0], np.cumsum(out["num_supp"])[:-1]]))
    print(f"{name}: N={len(params)} nnz={len(Jm)} S in [{min(num_supp)},{max(num_supp)}] "
          f"fns={ {l: int(h.fns[l].sum()) for l in h.fns} } cells={ {l: int(h.cells[l].sum()) for l in h.cells} } "
          f"PoU dev={np.abs(pou - 1).max():.2e}")
    return out


def kat(R):
    """Span table (reference find_span_array_jax), basisFun KAT, scipy derivative KATs."""
    from scipy.interpolate import BSpline

    B = R.bspline_funcs
    u = np.array([-0.1, 0, 1e-5, 0.2, 0.39999, 0.4, 0.9, 0.99999, 1, 1.1])
    spans = np.asarray(B.find_span_array_jax(u, U_B, 3))
    N = B.basisFun(5, 0.45, 3, U_B)
    # derivatives of the 4 functions non-zero at u=0.45 (span 5 -> functions 2..5)
    n = len(U_B) - 4
    dN = []
    for j in range(2, 6):
        c = np.zeros(n)
        c[j] = 1
        dN.append(float(BSpline(U_B, c, 3).derivative()(0.45)))
    # scalar spans from the reference's findSpan (called the way utils.py does, n = #fns - 1)
    us = np.array([0.0, 0.05, 0.2, 0.45, 0.5, 0.77, 0.9, 0.95])
    fs = np.array([B.findSpan(n - 1, 3, float(x), U_B) for x in us])
    # a wider random derivative table over both README knot vectors
    rng = np.random.default_rng(7)
    tab = {}
    for tag, U, p in (("A", U_A, 2), ("B", U_B, 3)):
        pts = np.sort(rng.random(23)).astype(np.float32).astype(np.float64)
        nf = len(U) - p - 1
        sp = np.asarray(B.find_span_array_jax(pts, U, p))
        vals = np.zeros((len(pts), p + 1))
        ders = np.zeros((len(pts), p + 1))
        for a, x in enumerate(pts):
            vals[a] = B.basisFun(int(sp[a]), float(x), p, U)
            for r in range(p + 1):
                c = np.zeros(nf)
                c[sp[a] - p + r] = 1
                ders[a, r] = float(BSpline(U, c, p).derivative()(x))
        tab[f"pts_{tag}"], tab[f"span_{tag}"], tab[f"N_{tag}"], tab[f"dN_{tag}"] = pts, sp, vals, ders
    return dict(span_u=u, span_idx=spans, U_B=U_B, U_A=U_A, N_045=N, dN_045=np.array(dN), fs_u=us, fs_idx=fs, **tab)


def contraction(R):
    """compute_pts.cpp (the reference's CPU extension, compiled unmodified) on seeded random CSR data."""
    import torch

    from oracle import build_ref

    build_ref.build()
    ext = build_ref.load()
    rng = np.random.default_rng(3)
    out = {}
    for tag, (N, nCP, smin, smax) in {"a": (37, 50, 1, 9), "b": (64, 200, 12, 30), "c": (5, 7, 0, 4)}.items():
        ns = rng.integers(smin, smax + 1, size=N)
        if tag == "c":
            ns[1] = 0  # empty segment
        nnz = int(ns.sum())
        cp = rng.standard_normal((nCP, 3)).astype(np.float32)
        Jm = rng.integers(0, nCP, size=nnz).astype(np.int64)
        phi = rng.random(nnz).astype(np.float32)
        g = rng.standard_normal((N, 3)).astype(np.float32)
        cs = np.cumsum(ns).astype(np.int64)
        fwd = ext.cpp_forward(torch.from_numpy(cp), torch.from_numpy(Jm), torch.from_numpy(phi), torch.from_numpy(cs)).numpy()
        bwd = ext.cpp_backward(torch.from_numpy(g), torch.from_numpy(cp), torch.from_numpy(Jm), torch.from_numpy(phi), torch.from_numpy(cs)).numpy()
        for k, v in dict(cp=cp, Jm=Jm, phi=phi, cumsum=cs, grad_out=g, fwd=fwd, bwd=bwd).items():
            out[f"{tag}_{k}"] = v
    return out


def bezier_and_vtu(R):
    """Bezier extraction (bspline_funcs.py:242-284), per-point multilevel extraction operators
    (core.py:295-379) and the hexahedral adaptive grid of plot3DAdaptiveGrid (utils.py:296-399; pyvista is replaced
    by a stand-in that records what would have been written to the .vtu file)."""
    import types
    from itertools import product

    S, C, B, U = R.space, R.core, R.bspline_funcs, R.utils
    out = {}
    for tag, knots, p in (("A", U_A, 2), ("B", U_B, 3), ("D", U_D, 3)):
        out[f"bez_{tag}"] = np.stack(B.bezier_extraction(np.asarray(knots, dtype=np.float64), p))
    for name in ("block2d", "block3d"):
        knots, degrees, nlev, boxes, grid, extra = CONFIGS[name]
        tp = S.TensorProduct([S.BSpline(np.array(Uk, dtype=np.float64), p) for Uk, p in zip(knots, degrees)])
        h = S.Space(tp, nlev)
        h.build_hierarchy_from_domain_sequence()
        for lev, lo, hi in boxes:
            for c in product(*[range(a, b) for a, b in zip(lo, hi)]):
                h._refine_cell(tuple(c), lev)
        h.build_hierarchy_from_domain_sequence()
        ac = C.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
        fc = C.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
        rng = np.random.default_rng(11)
        prm = rng.random((23, len(degrees))).astype(np.float32).astype(np.float64) * 0.998 + 0.001
        spans, num_supp, _ = C.compute_active_span_v2(prm, h.knotvectors, h.cells, h.degrees, ac, fc)
        ops = C.compute_multilevel_bezier_extraction_operators(prm, spans, ac, fc, h.sh_cells, h.sh_fns, h.knotvectors, h.degrees)
        out[f"{name}_params"] = prm
        out[f"{name}_num_supp"] = np.array(num_supp, dtype=np.int64)
        out[f"{name}_ops"] = np.concatenate([np.asarray(o, dtype=np.float64).reshape(len(o), -1) for o in ops])
        if name == "block3d":
            rec = {}

            class Grid:
                def __init__(self, cells, cell_types, points):
                    rec["cells"], rec["types"], rec["points"] = np.asarray(cells), np.asarray(cell_types), np.asarray(points)

                def save(self, filename):
                    rec["filename"] = filename

            U.pv.CellType = types.SimpleNamespace(HEXAHEDRON=12)
            U.pv.UnstructuredGrid = Grid
            off = np.cumsum([0] + [int(np.prod(h.sh_fns[l])) for l in range(nlev)])
            from thb_diff_b200 import configs as my_configs  # Greville control points (same formula as bench.py)
            import thb_diff_b200.multilevel_spline_space as mine

            hs = mine.Space(mine.TensorProduct([mine.BSpline(np.array(Uk, dtype=np.float64), p) for Uk, p in zip(knots, degrees)]), nlev)
            cp = my_configs.greville_ctrl_pts(hs).astype(np.float64)
            cps = {l: cp[off[l]:off[l + 1]].reshape(*h.sh_fns[l], 3) for l in range(nlev)}
            U.plot3DAdaptiveGrid("golden_grid", ac, cps, h.knotvectors, fc, h.sh_fns, h.degrees)
            out["vtu_cells"], out["vtu_types"], out["vtu_points"] = rec["cells"], rec["types"], rec["points"]
            out["vtu_ctrl_pts"] = cp
    return out


def main():
    R = _refshim.load_reference(fp16_tables=False)
    np.savez_compressed(os.path.join(HERE, "bezier_vtu.npz"), **bezier_and_vtu(R))
    for name, spec in CONFIGS.items():
        np.savez_compressed(os.path.join(HERE, f"space_{name}.npz"), **run_config(R, name, spec))
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat(R))
    np.savez_compressed(os.path.join(HERE, "contraction.npz"), **contraction(R))
    R16 = _refshim.load_reference(fp16_tables=True)
    np.savez_compressed(os.path.join(HERE, "space_block2d_fp16.npz"), **run_config(R16, "block2d_fp16", CONFIGS["block2d"]))


if __name__ == "__main__":
    main()
