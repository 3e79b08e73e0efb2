"""Bezier extraction of the hierarchy -- SURVEY 8f rank 4 (the "wire format" side of the hot path: per-element
extraction operators for IGA solvers).

Mirrors  bezier_extraction  [reference: THB/bspline_funcs.py:242-284]  and
compute_multilevel_bezier_extraction_operators  [reference: THB/core.py:295-379]  (same names, argument order and
results).  The reference slices its dense [nfns(lev) x nfns(finest)] coefficient tables per (point, support); here
the packed coefficient tiles are used instead: a THB function restricted to an active cell is  <tile, tp>  in the
cell's own-level tensor-product basis, so its Bernstein coefficients on the finest-level cell `cf` below it are
    C_s = tile_s  x_0 M_0[cf_0]  x_1 M_1[cf_1] ( x_2 M_2[cf_2] ),
    M_k[cf] = R_k(level -> finest)[window of the cell, window of cf] @ Cbez_k[cf]         ((p+1) x (p+1) per dimension)
with R_k the product of the 1-D subdivision matrices and Cbez_k the Bezier extraction operators of the finest knot
vector.  Host-side post-processing (numpy / torch float64); nothing here is on the timed path."""
import numpy as np
import torch

from .multilevel_spline_space import assemble_Tmatrix


def bezier_extraction(knot, p):
    """Bezier extraction operators of a knot vector: list (one per non-empty span) of [p+1, p+1] matrices C with
    N_a = sum_j C[a, j] B_j  (N: the p+1 B-splines of the span, B: the Bernstein polynomials of the span).
    [reference: THB/bspline_funcs.py:242-284].  Computed from the knot-insertion matrix that raises every interior
    breakpoint to multiplicity p (Boehm), restricted to the functions of each span."""
    U = np.asarray(knot, dtype=np.float64)
    p = int(p)
    brk, mult = np.unique(U, return_counts=True)
    if len(brk) < 2:
        return [np.eye(p + 1)]
    fine = np.concatenate([np.repeat(brk[0], mult[0]), np.repeat(brk[1:-1], np.maximum(mult[1:-1], p)), np.repeat(brk[-1], mult[-1])])
    T = assemble_Tmatrix(U, fine, p=p)            # [n_fine, n]:  N_old_j = sum_i T[i, j] N_fine_i
    out = []
    for e in range(len(brk) - 1):
        i = int(np.searchsorted(U, brk[e], side="right")) - 1       # span of the coarse knot vector
        f = int(np.searchsorted(fine, brk[e], side="right")) - 1    # span of the Bezier knot vector
        out.append(np.ascontiguousarray(T[f - p:f + 1, i - p:i + 1].T))
    return out


class ExtractionTables:
    """M[lev][k]: [ncells_finest_k, p_k+1, p_k+1] -- level-`lev` window of the ancestor cell -> Bernstein basis of the
    finest cell (see module docstring)."""

    def __init__(self, knotvectors, degrees):
        L = len(knotvectors)
        d = len(degrees)
        self.L, self.d, self.degrees = L, d, tuple(int(p) for p in degrees)
        kv = {l: [np.asarray(U, dtype=np.float64) for U in knotvectors[l]] for l in range(L)}
        self.fine_knots = kv[L - 1]
        self.M = {}
        for k in range(d):
            p = self.degrees[k]
            Cb = np.stack(bezier_extraction(kv[L - 1][k], p))                    # [nc_fine, p+1, p+1]
            nc = Cb.shape[0]
            R = np.eye(len(kv[L - 1][k]) - p - 1)                                # level -> finest, built downwards
            for lev in range(L - 1, -1, -1):
                if lev < L - 1:
                    R = assemble_Tmatrix(kv[lev][k], kv[lev + 1][k], p=p).T @ R   # [nf(lev), nf(L-1)]
                cf = np.arange(nc)
                c = cf >> (L - 1 - lev)
                rows = c[:, None] + np.arange(p + 1)[None, :]
                cols = cf[:, None] + np.arange(p + 1)[None, :]
                sub = R[rows[:, :, None], cols[:, None, :]]                       # [nc, p+1, p+1]
                self.M[(lev, k)] = np.einsum("cij,cjb->cib", sub, Cb)

    def finest_cells(self, params):
        """finest-level cell index per point and dimension (span rule of find_span_array_jax)"""
        prm = np.asarray(params, dtype=np.float64)
        out = np.empty(prm.shape, dtype=np.int64)
        for k in range(self.d):
            U, p = self.fine_knots[k], self.degrees[k]
            n = len(U) - p - 1
            s = np.searchsorted(U, prm[:, k], side="right") - 1
            s = np.where(prm[:, k] == U[-1], n - 1, np.minimum(s, n))
            out[:, k] = s - p
        return out


def _pair_tiles(P, slots):
    """[nnz, T] float64 coefficient rows (cell-level basis) of every (point, support) pair, reference pair order."""
    info = P.cellinfo_host
    S = info[slots, 5].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(S)])
    nnz = int(off[-1])
    rows = np.zeros((nnz, P.T), dtype=np.float64)
    tiles = P.tiles.cpu().numpy()
    dm = P.dmask.cpu().numpy().view(np.uint64)
    for i, s in enumerate(slots):
        ndir, toff = int(info[s, 7]), int(info[s, 6])
        b = int(off[i])
        if ndir:
            pos = [t for t in range(P.T) if (int(dm[s]) >> t) & 1]
            rows[b + np.arange(ndir), pos] = 1.0
        nt = int(S[i]) - ndir
        if nt:
            rows[b + ndir:b + ndir + nt] = tiles[toff:toff + nt, :P.T]
    return rows, off


def compute_multilevel_bezier_extraction_operators(params, ac_spans, ac_cells_supp, fn_coeffs, cell_shapes, fn_shapes,
                                                   knotvectors, degrees):
    """[reference: THB/core.py:295-379] -> list (one entry per point) of arrays [S_i, p_0+1, p_1+1(, p_2+1)]: the
    Bernstein coefficients, on the finest-level cell that contains the point, of every THB function supported on the
    point's active cell (support order of compute_active_span).

    `ac_spans` or `ac_cells_supp` must be the PointSupports object returned by compute_active_span(_v2) (it carries
    the packed hierarchy and the cell of every point); the other arguments keep the reference's positions."""
    from .core import PointSupports, RefinementOperators

    supp = ac_spans if isinstance(ac_spans, PointSupports) else ac_cells_supp
    if not isinstance(supp, PointSupports):
        raise RuntimeError("pass the supports object returned by compute_active_span / compute_active_span_v2")
    P = supp.P
    if not P.has_tiles:
        if not isinstance(fn_coeffs, RefinementOperators):
            raise RuntimeError("the hierarchy has no coefficient tiles yet and fn_coeffs is not a RefinementOperators object")
        P.attach_tiles(fn_coeffs.coeffs)
    slots = supp.slot.cpu().numpy().astype(np.int64)
    if (slots < 0).any():
        raise ValueError("a point lies outside the domain")
    tabs = ExtractionTables(knotvectors, degrees)
    prm = np.asarray(params.cpu() if isinstance(params, torch.Tensor) else params, dtype=np.float64)
    cf = tabs.finest_cells(prm)
    rows, off = _pair_tiles(P, slots)
    d, deg = P.ndim, P.degrees
    n = [p + 1 for p in deg]
    lev = P.cellinfo_host[slots, 0].astype(np.int64)
    S = np.diff(off)
    pt = np.repeat(np.arange(len(slots)), S)
    t = torch.from_numpy(rows.reshape(-1, *n))
    Ms = [torch.from_numpy(np.stack([tabs.M[(int(l), k)][c] for l, c in zip(lev, cf[:, k])]))[pt] for k in range(d)]
    if d == 2:
        ops = torch.einsum("sab,sax,sby->sxy", t, Ms[0], Ms[1])
    else:
        ops = torch.einsum("sabc,sax,sby,scz->sxyz", t, Ms[0], Ms[1], Ms[2])
    ops = ops.numpy()
    return [ops[off[i]:off[i + 1]] for i in range(len(slots))]
