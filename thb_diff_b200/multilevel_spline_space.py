"""Hierarchical spline space -- host-side set-up, vectorised for spaces with millions of functions.

Mirrors the reference's public classes (same names, same attribute layout) so that user scripts written
against ``THB.multilevel_spline_space`` keep working:  BSpline, TensorProduct, Space
[reference: THB/multilevel_spline_space.py:13-151].  This is the producer of the hot path's inputs; the
hot path itself lives in csrc/ and is reached through core.py / torch_funcs.py.

Differences from the reference, all deliberate:
  * masks are uint8 (``cells[lev][idx] == 1`` still works), no Python loops over functions
  * ``refine_box`` refines a whole index box at once (the reference only has the per-cell call)
  * knot insertion matrices come from Boehm's algorithm instead of the Oslo recursion
"""
from itertools import product as _iproduct

import numpy as np


def refine_knotvector(knotvector, p):
    """Insert the midpoint of every non-empty knot span; end knots keep multiplicity p+1, interior
    multiplicities collapse to 1.  [reference: THB/bspline_funcs.py:7-17]"""
    kv = np.asarray(knotvector, dtype=np.float64)
    breaks = np.unique(kv)
    out = np.empty(2 * len(breaks) - 1)
    out[0::2] = breaks
    out[1::2] = 0.5 * (breaks[1:] + breaks[:-1])
    return np.concatenate([kv[:p], out, kv[len(kv) - p:]])


def assemble_Tmatrix(knotVec, newKnotVec, knotVec_len=None, newKnotVec_len=None, p=None):
    """Knot-insertion matrix T[len(new)-p-1, len(old)-p-1] with  B^old_j = sum_i T[i,j] B^new_i.
    Same signature and result as the reference [THB/bspline_funcs.py:197-239]; computed by inserting the
    new knots one at a time (Boehm), which is O(#new * p) row updates instead of an O(n^2 p) recursion."""
    U = np.asarray(knotVec, dtype=np.float64)
    V = np.asarray(newKnotVec, dtype=np.float64)
    if p is None:
        raise TypeError("degree p is required")
    n = len(U) - p - 1
    A = np.eye(n)                      # rows: current basis, cols: original basis
    cur = U.copy()
    # multiset difference V \ U (both sorted)
    iu = 0
    new = []
    for v in V:
        if iu < len(U) and U[iu] == v:
            iu += 1
        else:
            new.append(v)
    if iu != len(U):
        raise ValueError("new knot vector does not contain the old one")
    for x in new:
        k = int(np.searchsorted(cur, x, side="right")) - 1        # cur[k] <= x < cur[k+1]
        m = A.shape[0]
        B = np.zeros((m + 1, n))
        B[: k - p + 1] = A[: k - p + 1]
        for i in range(k - p + 1, k + 1):
            a = (x - cur[i]) / (cur[i + p] - cur[i])
            B[i] = a * A[i] + (1.0 - a) * A[i - 1]
        B[k + 1:] = A[k:]
        A = B
        cur = np.insert(cur, k + 1, x)
    return A


class BSpline:
    def __init__(self, knotvector, degree):
        self.knotvector = np.asarray(knotvector, dtype=np.float64)
        self.degree = int(degree)

    def refine_bspline(self):
        self.next_level = BSpline(refine_knotvector(self.knotvector, self.degree), self.degree)
        return self.next_level


class TensorProduct:
    def __init__(self, bsplines):
        self.bsplines = list(bsplines)

    def refine_tensorproduct(self):
        self.next_level = TensorProduct([b.refine_bspline() for b in self.bsplines])
        return self.next_level


def _window_any(mask, p):
    """out[i] = any(mask[max(i-p,0) : min(i+1,n)]) along every axis; shape grows by p per axis.
    (support cells of function i are cells i-p..i clipped to the grid, reference core.py:452-481)"""
    out = mask.astype(bool)
    for ax, pk in enumerate(p):
        pad = [(0, 0)] * out.ndim
        pad[ax] = (pk, pk)
        padded = np.pad(out, pad)
        n = out.shape[ax] + pk
        acc = np.zeros_like(np.take(padded, range(n), axis=ax))
        for s in range(pk + 1):
            acc |= np.take(padded, range(s, s + n), axis=ax)
        out = acc
    return out


class Space:
    """Same attributes as the reference's Space: H, degrees, num_levels, ndim, knotvectors, sh_knots,
    sh_cells, sh_fns, cells, Coeff, fns, fn_states."""

    def __init__(self, tensor_product, num_levels):
        self.H = {0: tensor_product}
        self.degrees = tuple(b.degree for b in tensor_product.bsplines)
        for lev in range(1, num_levels):
            self.H[lev] = self.H[lev - 1].refine_tensorproduct()
        self.num_levels = int(num_levels)
        self.ndim = len(tensor_product.bsplines)
        self.fn_states = []
        self.initialize_datastructures()
        self.compute_coefficients()
        self.fns = None

    def initialize_datastructures(self):
        self.knotvectors = {lev: tuple(b.knotvector for b in tp.bsplines) for lev, tp in self.H.items()}
        self.sh_knots = {lev: tuple(len(np.unique(U)) for U in self.knotvectors[lev]) for lev in self.knotvectors}
        self.sh_cells = {lev: tuple(k - 1 for k in self.sh_knots[lev]) for lev in self.knotvectors}
        self.sh_fns = {lev: tuple(c + p for c, p in zip(self.sh_cells[lev], self.degrees)) for lev in self.knotvectors}
        self.cells = {lev: np.zeros(self.sh_cells[lev], dtype=np.uint8) for lev in self.knotvectors}
        self.cells[0][...] = 1

    def compute_coefficients(self):
        """Coeff[lev][dim][i, j] = weight of level-(lev+1) function j in level-lev function i."""
        self.Coeff = {}
        for lev in range(self.num_levels - 1):
            self.Coeff[lev] = {
                k: assemble_Tmatrix(self.knotvectors[lev][k], self.knotvectors[lev + 1][k], p=self.degrees[k]).T.copy()
                for k in range(self.ndim)
            }

    def build_hierarchy_from_domain_sequence(self):
        """Function activity masks from the cell masks.  [reference: multilevel_spline_space.py:94-124]"""
        fns = {lev: np.zeros(self.sh_fns[lev], dtype=np.uint8) for lev in range(self.num_levels)}
        fns[0][...] = 1
        for lev in range(self.num_levels - 1):
            touches_active_cell = _window_any(self.cells[lev] == 1, self.degrees)
            dropped = (fns[lev] == 1) & ~touches_active_cell
            if not dropped.any():
                break
            fns[lev][dropped] = 0
            # children of dropped functions: separable boolean contraction with the subdivision sparsity
            kids = dropped.astype(np.float32)
            for k in range(self.ndim):
                nz = (self.Coeff[lev][k] != 0).astype(np.float32)       # [nf(lev), nf(lev+1)]
                kids = np.moveaxis(np.tensordot(kids, nz, axes=([k], [0])), -1, k)
                kids = (kids > 0).astype(np.float32)
            fns[lev + 1][kids > 0] = 1
        self.fns = fns
        self.fn_states.append(fns)
        return fns

    def _refine_cell(self, cellIdx, level):
        self.refine_box(tuple(cellIdx), tuple(c + 1 for c in cellIdx), level)

    def refine_box(self, lo, hi, level):
        """Deactivate the cells of the index box [lo, hi) at `level` and activate all their children."""
        assert level < self.num_levels - 1
        self.cells[level][tuple(slice(a, b) for a, b in zip(lo, hi))] = 0
        self.cells[level + 1][tuple(slice(2 * a, 2 * b) for a, b in zip(lo, hi))] = 1

    def _refine_basis_fn(self, fnIdx, level):
        lo = tuple(max(i - p, 0) for i, p in zip(fnIdx, self.degrees))
        hi = tuple(min(i + 1, n) for i, n in zip(fnIdx, self.sh_cells[level]))
        self.refine_box(lo, hi, level)

    def _collapse_cell(self, cellIdx, level):
        # same (quirky) effect as reference multilevel_spline_space.py:142-151
        parent = tuple(int(c) // 2 for c in cellIdx)
        self.cells[level - 1][tuple(cellIdx)] = 1
        for off in _iproduct(range(2), repeat=self.ndim):
            self.cells[level][tuple(2 * c + o for c, o in zip(parent, off))] = 0


class ControlPoints:
    """Control-point bookkeeping across refinements.  [reference: THB/multilevel_spline_space.py:154-188]

    update_CP(CP_arr, fns, Coeffs): control points of functions that became active since the last call are
    obtained by prolongation of the next coarser level (knot-insertion weights, separable per dimension);
    all other rows are kept.  Returns {lev: ndarray [*sh_fns[lev], ndim]} like the reference.  Vectorised:
    the prolongation of a whole level is one tensor contraction per dimension."""

    def __init__(self, H_space):
        self.h_space = H_space
        self.max_lev = H_space.num_levels - 1
        self.ndim = H_space.ndim
        self.CP_status = {lev: np.zeros_like(H_space.fns[lev]) for lev in range(self.max_lev + 1)}
        self.CP_status[0] = np.ones_like(self.CP_status[0])

    def update_CP(self, CP_arr, fns, Coeffs):
        CP_arr = np.asarray(CP_arr, dtype=np.float64)
        sh = self.h_space.sh_fns
        off = np.cumsum([0] + [int(np.prod(sh[lev])) for lev in range(self.max_lev + 1)])
        cpd = CP_arr.shape[1]
        ctrl = {lev: CP_arr[off[lev]:off[lev + 1]].reshape(*sh[lev], cpd).copy() for lev in range(self.max_lev + 1)}
        for lev in range(1, self.max_lev + 1):
            new = (np.asarray(self.CP_status[lev]) == 0) & (np.asarray(fns[lev]) == 1)
            if not new.any():
                continue
            fine = ctrl[lev - 1]
            for k in range(self.ndim):  # P[j] = sum_i Coeff[i, j] * P_coarse[i] along every dimension
                fine = np.moveaxis(np.tensordot(fine, np.asarray(Coeffs[lev - 1][k], dtype=np.float64), axes=([k], [0])), -1, k)
            ctrl[lev][new] = fine[new]
        self.CP_status = {lev: np.array(fns[lev], copy=True) for lev in fns}
        return ctrl
