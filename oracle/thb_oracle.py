"""CPU ORACLE (test infrastructure -- NOT part of the product).

A float64 NumPy restatement of the THB-Diff evaluation hot path and of the host set-up code that
produces its inputs.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product (``thb_diff_b200``) never does.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by executing the reference's own
``Space`` / ``compute_active_cells_active_supp`` / ``compute_refinement_operators`` /
``compute_active_span_v2`` / ``compute_THB_fns_tp`` code (see tests/golden/make_golden.py) and by the
reference's compiled ``compute_pts.cpp`` (oracle/_ref); ``tests/test_oracle_golden.py`` checks every
function below against them.

Every function cites the reference lines (relative to /root/reference) whose behaviour it restates.
"""
from itertools import product as _iproduct

import numpy as np

# --------------------------------------------------------------------------------------------------
# 1-D primitives
# --------------------------------------------------------------------------------------------------


def refine_knotvector(U, p):
    """Dyadic refinement: every distinct knot plus every midpoint, end knots repeated p+1 times.
    Interior multiplicities are dropped.  [THB/bspline_funcs.py:7-17]"""
    U = np.asarray(U, dtype=np.float64)
    brk = np.unique(U)
    fine = np.unique(np.concatenate([brk, 0.5 * (brk[:-1] + brk[1:])]))
    return np.concatenate([U[:p], fine, U[len(U) - p:]])


def generate_parametric_coordinates(shape):
    """Synthetic parametric grid, numpy 'xy' meshgrid order (2-D: x fastest; 3-D: z fastest, then x,
    y slowest); values linspace(1e-5, 1, n, endpoint=False).  [THB/bspline_funcs.py:20-35, SURVEY Q14]"""
    axes = [np.linspace(1e-5, 1.0, n, endpoint=False) for n in shape]
    grids = np.meshgrid(*axes)
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def greville_abscissae(fn_sh, degrees, knotvectors):
    """Greville points as control net, shape [*fn_sh, d].  [THB/bspline_funcs.py:38-51]"""
    d = len(fn_sh)
    per_dim = []
    for k in range(d):
        U = np.asarray(knotvectors[k], dtype=np.float64)
        per_dim.append(np.array([U[i + 1:i + degrees[k] + 1].sum() / degrees[k] for i in range(fn_sh[k])]))
    out = np.zeros((*fn_sh, d))
    for k in range(d):
        shp = [1] * d
        shp[k] = fn_sh[k]
        out[..., k] = per_dim[k].reshape(shp)
    return out


def find_span_array(params, U, p):
    """Vector span rule of the reference: searchsorted(U,u,'right')-1; >n -> n; u==U[-1] -> n-1;
    u<U[0] -> -1  (n = len(U)-p-1 = number of functions).  [THB/bspline_funcs.py:77-83, SURVEY Q10]"""
    U = np.asarray(U)
    u = np.asarray(params)
    n = len(U) - p - 1
    idx = np.searchsorted(U, u, side="right") - 1
    idx = np.where(idx > n, n, idx)
    idx = np.where(u == U[-1], n - 1, idx)
    return idx.astype(np.int64)


def find_span_scalar(n_minus_1, p, u, U):
    """NURBS-book A2.1 binary search with n = (#functions - 1).  [THB/bspline_funcs.py:62-74].
    In-domain only (the reference loops forever out of domain)."""
    n = n_minus_1
    if not (U[0] <= u <= U[-1]):
        raise ValueError("find_span_scalar is only defined in-domain")
    if u == U[n + 1]:
        return n
    lo, hi = p, n + 1
    mid = (lo + hi) // 2
    while u < U[mid] or u >= U[mid + 1]:
        if u < U[mid]:
            hi = mid
        else:
            lo = mid
        mid = (lo + hi) // 2
    return mid


def basis_funs(i, u, p, U):
    """The p+1 non-vanishing B-spline values at span i (NURBS-book A2.2).  [THB/bspline_funcs.py:86-101]"""
    N = np.zeros(p + 1)
    left = np.zeros(p + 1)
    right = np.zeros(p + 1)
    N[0] = 1.0
    for j in range(1, p + 1):
        left[j] = u - U[i + 1 - j]
        right[j] = U[i + j] - u
        carry = 0.0
        for r in range(j):
            t = N[r] / (right[r + 1] + left[j - r])
            N[r] = carry + right[r + 1] * t
            carry = left[j - r] * t
        N[j] = carry
    return N


def basis_funs_and_ders(i, u, p, U):
    """Values and first derivatives of the p+1 non-vanishing functions at span i.
    The reference obtains the derivative with jax.jacfwd of the dense recursion
    [THB/bspline_funcs.py:149-194]; in-domain that equals the classical formula
    N'_{j,p} = p N_{j,p-1}/(U[j+p]-U[j]) - p N_{j+1,p-1}/(U[j+p+1]-U[j+1])  used here."""
    N = basis_funs(i, u, p, U)
    dN = np.zeros(p + 1)
    if p == 0:
        return N, dN
    low = basis_funs(i, u, p - 1, U)  # functions i-p+1 .. i of degree p-1
    a = np.zeros(p + 2)
    for r in range(1, p + 1):
        den = U[i + r] - U[i - p + r]
        a[r] = low[r - 1] / den if den != 0 else 0.0
    for r in range(p + 1):
        dN[r] = p * (a[r] - a[r + 1])
    return N, dN


def basis_all_dense(u, U, p):
    """Dense Cox-de Boor over all knot intervals, last knot nudged by eps so u==U[-1] lands in the last
    interval, 0/0 -> 0.  Restates basisFun_jax [THB/bspline_funcs.py:104-109,149-182]; returns the p+1
    window [span-p, span]."""
    U = np.asarray(U, dtype=np.float64)
    K = np.where(U == U[-1], U[-1] + np.finfo(np.float64).eps, U)
    N = ((u >= K[:-1]) & (u < K[1:])).astype(np.float64)

    def safe_div(a, b):
        z = (a == 0) & (b == 0)
        return np.where(z, 0.0, a) / np.where(z, 1.0, b)

    for q in range(1, p + 1):
        t1 = safe_div(N[:-1] * (u - K[:-q - 1]), K[q:-1] - K[:-q - 1])
        t2 = safe_div(N[1:] * (K[q + 1:] - u), K[q + 1:] - K[1:-q])
        N = t1 + t2
    s = int(find_span_array(np.array([u]), U, p)[0]) - p
    return N[s:s + p + 1]


def knot_insertion_matrix(U, V, p):
    """T with  B^U_j = sum_i T[i, j] B^V_i  for nested knot vectors U (coarse) in V (fine) -- the Oslo
    recursion of assemble_Tmatrix.  Shape [len(V)-p-1, len(U)-p-1].  [THB/bspline_funcs.py:197-239]"""
    U = np.asarray(U, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    nu, nv = len(U), len(V)
    T = ((V[:nv - 1, None] >= U[None, :nu - 1]) & (V[:nv - 1, None] < U[None, 1:nu])).astype(np.float64)
    for q in range(1, p + 1):
        rows, cols = nv - q - 1, nu - q - 1
        Tn = np.zeros((rows, cols))
        vq = V[q:q + rows][:, None]            # V[i+q]
        uj = U[:cols][None, :]                 # U[j]
        ujq = U[q:q + cols][None, :]           # U[j+q]
        uj1 = U[1:1 + cols][None, :]           # U[j+1]
        ujq1 = U[q + 1:q + 1 + cols][None, :]  # U[j+q+1]
        d1 = ujq - uj
        d2 = ujq1 - uj1
        with np.errstate(divide="ignore", invalid="ignore"):
            a = np.where(d1 != 0, (vq - uj) / np.where(d1 != 0, d1, 1.0), 0.0)
            b = np.where(d2 != 0, (ujq1 - vq) / np.where(d2 != 0, d2, 1.0), 0.0)
        Tn = a * T[:rows, :cols] + b * T[:rows, 1:cols + 1]
        T = Tn
    return T


# --------------------------------------------------------------------------------------------------
# Hierarchical space (host set-up; restated only to synthesise inputs)
# --------------------------------------------------------------------------------------------------


class Space:
    """Multi-level tensor-product spline space.  [THB/multilevel_spline_space.py:34-151]

    knotvectors[lev] = tuple of 1-D knot vectors; sh_cells / sh_fns per level; cells[lev] 0/1 float
    arrays (level 0 all active); Coeff[lev][dim] = knot_insertion_matrix(...).T with shape
    [nfns(lev), nfns(lev+1)]; fns[lev] 0/1 activity after build_hierarchy_from_domain_sequence()."""

    def __init__(self, knotvectors0, degrees, num_levels):
        self.degrees = tuple(int(p) for p in degrees)
        self.ndim = len(self.degrees)
        self.num_levels = num_levels
        kv = {0: tuple(np.asarray(U, dtype=np.float64) for U in knotvectors0)}
        for lev in range(1, num_levels):
            kv[lev] = tuple(refine_knotvector(kv[lev - 1][k], self.degrees[k]) for k in range(self.ndim))
        self.knotvectors = kv
        self.sh_cells = {lev: tuple(len(np.unique(U)) - 1 for U in kv[lev]) for lev in kv}
        self.sh_fns = {lev: tuple(c + p for c, p in zip(self.sh_cells[lev], self.degrees)) for lev in kv}
        self.cells = {lev: np.zeros(self.sh_cells[lev]) for lev in kv}
        self.cells[0][...] = 1.0
        self.Coeff = {
            lev: {k: knot_insertion_matrix(kv[lev][k], kv[lev + 1][k], self.degrees[k]).T for k in range(self.ndim)}
            for lev in range(num_levels - 1)
        }
        self.fns = None

    def _refine_cell(self, cellIdx, level):
        """Parent cell -> 0, its 2^d children -> 1.  [multilevel_spline_space.py:132-140]"""
        self.cells[level][tuple(cellIdx)] = 0
        for off in _iproduct(range(2), repeat=self.ndim):
            self.cells[level + 1][tuple(2 * c + o for c, o in zip(cellIdx, off))] = 1

    def refine_box(self, lo, hi, level):
        """Refine every cell of the half-open index box [lo, hi) at `level` (convenience for tests)."""
        for c in _iproduct(*[range(a, b) for a, b in zip(lo, hi)]):
            self._refine_cell(c, level)

    def build_hierarchy_from_domain_sequence(self):
        """Function activity: a level-l function is dropped when none of its support cells is active at
        level l; the children (non-zeros of the subdivision rows) of dropped functions become active at
        l+1; stops at the first level where nothing is dropped.  [multilevel_spline_space.py:94-124,
        core.py:382-406 (children), core.py:452-481 (support cells)]"""
        H = {lev: np.zeros(self.sh_fns[lev]) for lev in range(self.num_levels)}
        H[0][...] = 1.0
        for lev in range(self.num_levels - 1):
            dropped = []
            for fn in zip(*np.nonzero(H[lev])):
                rngs = [range(max(i - p, 0), min(i + 1, nc)) for i, p, nc in zip(fn, self.degrees, self.sh_cells[lev])]
                if not any(self.cells[lev][c] == 1.0 for c in _iproduct(*rngs)):
                    H[lev][fn] = 0
                    dropped.append(fn)
            if not dropped:
                break
            for fn in dropped:
                kids = [np.nonzero(self.Coeff[lev][k][fn[k]])[0] for k in range(self.ndim)]
                for kid in _iproduct(*kids):
                    H[lev + 1][kid] = 1
        self.fns = H
        return H


def compute_active_cells_active_supp(cells, fns, degrees):
    """{lev: {cell: [(fn_lev, fnIdx), ...]}} -- for each active cell the active functions of its own
    level's (p+1)^d window (C order), then of each ancestor cell (cell//2) down to level 0.
    [THB/core.py:14-40, 409-449]"""
    out = {}
    for lev in sorted(cells.keys()):
        cur = {}
        for cell in zip(*np.nonzero(cells[lev])):
            supp = []
            c = tuple(int(x) for x in cell)
            for l in range(lev, -1, -1):
                for fn in _iproduct(*[range(ci, ci + p + 1) for ci, p in zip(c, degrees)]):
                    if fns[l][fn] == 1:
                        supp.append((l, fn))
                c = tuple(ci // 2 for ci in c)
            cur[tuple(int(x) for x in cell)] = supp
        out[lev] = cur
    return out


def compute_refinement_operators(fns, Coeff, degrees, quantize_fp16=False, max_level="ones"):
    """Dense per-level tables fn_coeffs[lev][*fnIdx, *finestIdx]: the level-lev function written in the
    finest-level basis after SINGLE-LEVEL truncation (children active at lev+1 are zeroed, the rest is
    projected untruncated).  [THB/core.py:43-114]

    quantize_fp16 reproduces core.py:111; max_level='ones' reproduces core.py:113 (SURVEY
    Q6), 'identity' is the mathematically correct table.  Only feasible for small spaces (O(nfns^2))."""
    L = max(fns.keys())
    d = len(degrees)
    if L < 1:
        raise ValueError("the reference needs at least two levels")
    let = "abcdefghijkl"

    def outer(mats):  # [i,j],[k,l],.. -> [i,k,..,j,l,..]
        ins = ",".join(let[2 * k] + let[2 * k + 1] for k in range(d))
        outs = "".join(let[2 * k] for k in range(d)) + "".join(let[2 * k + 1] for k in range(d))
        return np.einsum(ins + "->" + outs, *mats)

    def compose(A, B):  # A[fn_l, fn_{l+1}] , B[fn_{l+1}, fn_L]
        return np.tensordot(A, B, axes=(list(range(d, 2 * d)), list(range(d))))

    step = {lev: outer([Coeff[lev][k] for k in range(d)]) for lev in range(L)}
    to_finest = {L - 1: None}  # projection level lev+1 -> L
    proj = {}
    if L >= 2:
        proj[L - 2] = step[L - 1]
        for lev in range(L - 3, -1, -1):
            proj[lev] = compose(step[lev + 1], proj[lev + 1])
    out = {}
    for lev in range(L - 1, -1, -1):
        cur = step[lev].copy()
        act = np.nonzero(fns[lev + 1] == 1)
        cur[(Ellipsis,) + act] = 0.0
        if lev < L - 1:
            cur = compose(cur, proj[lev])
        out[lev] = cur.astype(np.float16).astype(np.float64) if quantize_fp16 else cur
    shp = fns[L].shape
    if max_level == "ones":
        out[L] = np.ones((*shp, *shp))
    else:
        n = int(np.prod(shp))
        out[L] = np.eye(n).reshape(*shp, *shp)
    return out


# --------------------------------------------------------------------------------------------------
# Hot path
# --------------------------------------------------------------------------------------------------


def _level_cells(params, knotvectors, degrees, lev):
    d = len(degrees)
    return np.stack([find_span_array(params[:, k], knotvectors[lev][k], degrees[k]) - degrees[k] for k in range(d)], axis=1)


def compute_active_span(params, knotvectors, cells, degrees, ac_cells_ac_supp, return_spans=False):
    """Per point: the finest level whose cell (span - p per dim) is active; returns that cell's support
    list and its length.  [THB/core.py:157-189]; with return_spans also the (lev, cell) pairs of
    compute_active_span_v2 [THB/core.py:117-154].

    Spans are decided on float32-rounded params and knots -- callers pass f32-representable
    values; comparisons are then exact in any precision."""
    params = np.asarray(params)
    L = max(cells.keys())
    per_lev = {lev: _level_cells(params, knotvectors, degrees, lev) for lev in range(L + 1)}
    supp, nsupp, spans = [], [], []
    for i in range(len(params)):
        for lev in range(L, -1, -1):
            c = tuple(int(x) for x in per_lev[lev][i])
            if cells[lev][c] == 1:
                s = ac_cells_ac_supp[lev][c]
                supp.append(s)
                nsupp.append(len(s))
                spans.append((lev, c))
                break
    if return_spans:
        return supp, nsupp, spans
    return supp, nsupp


def flat_fn_offsets(fn_shapes):
    """Level-major flat numbering offsets of functions / control points.
    [THB/core.py:632-634, THB/torch_funcs.py:53-59]"""
    L = max(fn_shapes.keys())
    off = np.zeros(L + 2, dtype=np.int64)
    for lev in range(L + 1):
        off[lev + 1] = off[lev] + int(np.prod(fn_shapes[lev]))
    return off


def flat_support_indices(ac_cells_supp, fn_shapes):
    """Jm: offset[fn_lev] + C-order ravel of fnIdx for every (point, support) pair, point-major.
    [THB/core.py:644-650, THB/torch_funcs.py:72-76]"""
    off = flat_fn_offsets(fn_shapes)
    return np.array(
        [off[fl] + np.ravel_multi_index(fn, fn_shapes[fl]) for supp in ac_cells_supp for fl, fn in supp], dtype=np.int64
    )


def thb_basis_fns(params, ac_cells_supp, num_supp, fn_coeffs, fn_shapes, knotvectors, degrees, mode="value"):
    """Flat THB basis values for every (point, support) pair, evaluated the way the reference's float64
    loop does: finest-level span + Cox-de Boor, (p+1)^d tensor product, dot with the finest-level window
    of the support function's dense coefficient table.  [THB/core.py:232-268 (authoritative loop),
    core.py:629-701 (vectorised form with SURVEY Q2/Q3 repaired)]

    mode: 'value'; 'mixed' = the reference's return_mode=[False, True] (product of 1-D first
    derivatives in EVERY dimension); 'grad' = true gradient, returns [nnz, d]."""
    L = max(knotvectors.keys())
    d = len(degrees)
    params = np.asarray(params, dtype=np.float64)
    out = []
    for i, g in enumerate(params):
        span = [int(find_span_array(np.array([g[k]]), knotvectors[L][k], degrees[k])[0]) for k in range(d)]
        Nd = [basis_funs_and_ders(span[k], g[k], degrees[k], np.asarray(knotvectors[L][k], dtype=np.float64)) for k in range(d)]
        win = tuple(slice(span[k] - degrees[k], span[k] + 1) for k in range(d))

        def tp(sel):
            t = Nd[0][sel[0]]
            for k in range(1, d):
                t = np.multiply.outer(t, Nd[k][sel[k]])
            return t

        if mode == "value":
            tps = [tp([0] * d)]
        elif mode == "mixed":
            tps = [tp([1] * d)]
        else:
            tps = [tp([1 if k == j else 0 for k in range(d)]) for j in range(d)]
        for fl, fn in ac_cells_supp[i]:
            sub = fn_coeffs[fl][tuple(fn)][win]
            vals = [float(np.sum(sub * t)) for t in tps]
            out.append(vals[0] if mode != "grad" else vals)
    return np.array(out, dtype=np.float64)


def eval_forward(ctrl_pts, Jm, PHI, cumsum):
    """out[i] = sum_{j in [cumsum[i-1], cumsum[i])} PHI[j] * ctrl_pts[Jm[j]]  (inclusive cumsum).
    [THB/THB_extensions/compute_pts.cpp:4-29, compute_pts_cuda_kernels.cu:3-26]"""
    ctrl_pts = np.asarray(ctrl_pts, dtype=np.float64)
    cumsum = np.asarray(cumsum, dtype=np.int64)
    n = len(cumsum)
    out = np.zeros((n, ctrl_pts.shape[1]))
    if n == 0:
        return out
    seg = np.repeat(np.arange(n), np.diff(np.concatenate([[0], cumsum])))
    np.add.at(out, seg, np.asarray(PHI, dtype=np.float64)[:, None] * ctrl_pts[np.asarray(Jm)])
    return out


def eval_backward(grad_out, ctrl_pts_shape, Jm, PHI, cumsum):
    """grad_ctrl_pts[Jm[j]] += PHI[j] * grad_out[i].  [compute_pts.cpp:31-54, compute_pts_cuda_kernels.cu:28-47]"""
    grad_out = np.asarray(grad_out, dtype=np.float64)
    cumsum = np.asarray(cumsum, dtype=np.int64)
    g = np.zeros(ctrl_pts_shape)
    n = len(cumsum)
    if n == 0:
        return g
    seg = np.repeat(np.arange(n), np.diff(np.concatenate([[0], cumsum])))
    np.add.at(g, np.asarray(Jm), np.asarray(PHI, dtype=np.float64)[:, None] * grad_out[seg])
    return g


# --------------------------------------------------------------------------------------------------
# Table-free evaluation (any size): the same truncated functions evaluated from their definition
# --------------------------------------------------------------------------------------------------


class DirectEvaluator:
    """Evaluates the reference's (single-level-)truncated functions from the definition

        trunc(B^l_i)(x) = sum_{j child of i, fns[l+1][j] != 1}  prod_k Coeff[l][k][i_k, j_k] * B^{l+1}_j(x)

    (l < L; for l == L the function itself, i.e. the 'identity' reading of SURVEY Q6) with plain
    Cox-de Boor at level l+1 -- no dense tables, so it scales to the BASELINE configs.  In exact
    arithmetic this equals compute_refinement_operators + finest-level evaluation [THB/core.py:43-114,
    232-268]; tests check agreement to 1e-13 on the small golden spaces."""

    def __init__(self, knotvectors, degrees, cells, fns, Coeff):
        self.kv = {l: tuple(np.asarray(U, dtype=np.float64) for U in knotvectors[l]) for l in knotvectors}
        self.deg = tuple(degrees)
        self.d = len(degrees)
        self.cells, self.fns, self.Coeff = cells, fns, Coeff
        self.L = max(knotvectors.keys())
        self.off = flat_fn_offsets({l: fns[l].shape for l in fns})

    def _locate(self, x):
        for lev in range(self.L, -1, -1):
            c = tuple(int(find_span_array(np.array([x[k]]), self.kv[lev][k], self.deg[k])[0]) - self.deg[k] for k in range(self.d))
            if all(0 <= ci < n for ci, n in zip(c, self.cells[lev].shape)) and self.cells[lev][c] == 1:
                return lev, c
        return None

    def _supports(self, lev, c):
        out = []
        for l in range(lev, -1, -1):
            for fn in _iproduct(*[range(ci, ci + p + 1) for ci, p in zip(c, self.deg)]):
                if self.fns[l][fn] == 1:
                    out.append((l, fn))
            c = tuple(ci // 2 for ci in c)
        return out

    def _level_basis(self, x, lev):
        """span-window start and (N, dN) per dim at level lev"""
        res = []
        for k in range(self.d):
            U = self.kv[lev][k]
            s = int(find_span_array(np.array([x[k]]), U, self.deg[k])[0])
            N, dN = basis_funs_and_ders(s, x[k], self.deg[k], U)
            res.append((s - self.deg[k], N, dN))
        return res

    def point(self, x, want_grad=False):
        """-> (lev, cell, Jm[S], PHI[S] (, dPHI[S, d]))"""
        x = np.asarray(x, dtype=np.float64)
        loc = self._locate(x)
        if loc is None:
            return None
        lev, c = loc
        supp = self._supports(lev, c)
        basis = {}
        Jm, PHI, dPHI = [], [], []
        for fl, fn in supp:
            Jm.append(int(self.off[fl] + np.ravel_multi_index(fn, self.fns[fl].shape)))
            el = fl if fl == self.L else fl + 1
            if el not in basis:
                basis[el] = self._level_basis(x, el)
            B = basis[el]
            if fl == self.L:
                idx = [fn[k] - B[k][0] for k in range(self.d)]
                ok = all(0 <= idx[k] <= self.deg[k] for k in range(self.d))
                v = np.prod([B[k][1][idx[k]] for k in range(self.d)]) if ok else 0.0
                g = [np.prod([B[k][2 if k == j else 1][idx[k]] for k in range(self.d)]) if ok else 0.0 for j in range(self.d)]
            else:
                # children of fn inside the level-(fl+1) window of x, skipping active (truncated) ones
                w = []
                for k in range(self.d):
                    s0 = B[k][0]
                    w.append(self.Coeff[fl][k][fn[k], s0:s0 + self.deg[k] + 1])
                v = 0.0
                g = [0.0] * self.d
                for t in _iproduct(*[range(self.deg[k] + 1) for k in range(self.d)]):
                    child = tuple(B[k][0] + t[k] for k in range(self.d))
                    if self.fns[el][child] == 1:
                        continue
                    cw = np.prod([w[k][t[k]] for k in range(self.d)])
                    if cw == 0.0:
                        continue
                    v += cw * np.prod([B[k][1][t[k]] for k in range(self.d)])
                    for j in range(self.d):
                        g[j] += cw * np.prod([B[k][2 if k == j else 1][t[k]] for k in range(self.d)])
            PHI.append(v)
            dPHI.append(g)
        if want_grad:
            return lev, c, np.array(Jm), np.array(PHI), np.array(dPHI)
        return lev, c, np.array(Jm), np.array(PHI)
