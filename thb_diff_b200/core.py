"""THB evaluation core with the reference's call signatures [THB/core.py].

    ac_cells  = compute_active_cells_active_supp(cells, fns, degrees)
    fn_coeffs = compute_refinement_operators(fns, Coeff, degrees)
    ac_cell_supp, num_supp = compute_active_span(params, knotvectors, cells, degrees, ac_cells)
    PHI = THB_basis_fns(params, ac_cell_supp, num_supp, fn_coeffs, sh_fns, knotvectors, degrees)

The two set-up calls return light objects that remember their inputs (and still answer dict-style
queries); the two hot-path calls run on the GPU and return device tensors wrapped in list-like objects,
because Python lists of 10^9 tuples are not an option at the BASELINE sizes.  Dicts produced by the
reference's own set-up functions are accepted as well (legacy path, small spaces only).
"""
from collections.abc import Mapping, Sequence

import numpy as np
import torch

from . import engine
from .packed import PackedHierarchy

# --------------------------------------------------------------------------------------------------------
# set-up objects
# --------------------------------------------------------------------------------------------------------


class ActiveSupports(Mapping):
    """Result of compute_active_cells_active_supp.  ac[lev][cell] -> [(fn_lev, fnIdx), ...] is answered
    on demand from the masks (same order as the reference, core.py:424-449)."""

    def __init__(self, cells, fns, degrees):
        self.cells, self.fns, self.degrees = cells, fns, tuple(int(p) for p in degrees)
        self._packed = {}

    def __getitem__(self, lev):
        return _LevelSupports(self, lev)

    def __iter__(self):
        return iter(sorted(self.cells.keys()))

    def __len__(self):
        return len(self.cells)

    def cell_supports(self, lev, cell):
        from itertools import product

        out = []
        c = tuple(int(x) for x in cell)
        for l in range(lev, -1, -1):
            for fn in product(*[range(ci, ci + p + 1) for ci, p in zip(c, self.degrees)]):
                if self.fns[l][fn] == 1:
                    out.append((l, fn))
            c = tuple(ci // 2 for ci in c)
        return out

    def packed(self, knotvectors, device="cuda"):
        key = (id(knotvectors), str(device))
        if key not in self._packed:
            self._packed[key] = PackedHierarchy.build(knotvectors, self.degrees, self.cells, self.fns, None, device=device)
        return self._packed[key]


class _LevelSupports(Mapping):
    def __init__(self, parent, lev):
        self.parent, self.lev = parent, lev

    def __getitem__(self, cell):
        if self.parent.cells[self.lev][tuple(cell)] != 1:
            raise KeyError(cell)
        return self.parent.cell_supports(self.lev, cell)

    def __iter__(self):
        return (tuple(int(x) for x in c) for c in zip(*np.nonzero(self.parent.cells[self.lev])))

    def __len__(self):
        return int(np.count_nonzero(self.parent.cells[self.lev]))


def compute_active_cells_active_supp(cells, fns, degrees):
    """[THB/core.py:14-40]"""
    return ActiveSupports(cells, fns, degrees)


class RefinementOperators(Mapping):
    """Result of compute_refinement_operators: the subdivision matrices and activity masks from which the
    kernels' coefficient tiles are built.  op[lev] materialises the reference's dense table (small spaces
    only; float32, no fp16 quantisation, identity at the finest level -- SURVEY Q5/Q6)."""

    def __init__(self, fns, coeffs, degrees):
        self.fns, self.coeffs, self.degrees = fns, coeffs, tuple(int(p) for p in degrees)

    def __iter__(self):
        return iter(sorted(self.fns.keys()))

    def __len__(self):
        return len(self.fns)

    def __getitem__(self, lev):
        return self.dense(lev)

    def dense(self, lev):
        L = max(self.fns.keys())
        d = len(self.degrees)
        shp_l, shp_L = self.fns[lev].shape, self.fns[L].shape
        if float(np.prod(shp_l)) * float(np.prod(shp_L)) > 5e8:
            raise MemoryError("dense refinement table is too large; use the packed path")
        if lev == L:
            return np.eye(int(np.prod(shp_L)), dtype=np.float32).reshape(*shp_L, *shp_L)
        # single-level truncation, then untruncated projection to the finest level, one dimension at a time
        let = "abcdefgh"
        step = np.einsum(",".join(let[2 * k] + let[2 * k + 1] for k in range(d)) + "->" + "".join(let[2 * k] for k in range(d)) + "".join(let[2 * k + 1] for k in range(d)),
                         *[np.asarray(self.coeffs[lev][k], dtype=np.float64) for k in range(d)])
        step[(Ellipsis,) + np.nonzero(np.asarray(self.fns[lev + 1]) == 1)] = 0.0
        for m in range(lev + 1, L):
            for k in range(d):
                step = np.moveaxis(np.tensordot(step, np.asarray(self.coeffs[m][k], dtype=np.float64), axes=([d + k], [0])), -1, d + k)
        return step.astype(np.float32)


def compute_refinement_operators(fns, coeffs, degrees):
    """[THB/core.py:43-114]"""
    return RefinementOperators(fns, coeffs, degrees)


# --------------------------------------------------------------------------------------------------------
# hot path
# --------------------------------------------------------------------------------------------------------


class PointSupports(Sequence):
    """ac_cell_supp of compute_active_span: per point the support list of its active cell.  Device side:
    .slot [N] (active-cell slot id), .num_supp [N], .P (packed hierarchy), .plan.  Indexing returns the
    reference's list of (fn_level, fnIdx) tuples for that point (host conversion, on request only)."""

    def __init__(self, P, plan, num_supp):
        self.P, self.plan, self.num_supp = P, plan, num_supp
        self.slot = plan.slot[: plan.n]
        self._host = None

    def __len__(self):
        return self.plan.n

    def _tables(self):
        if self._host is None:
            self._host = (self.slot.cpu().numpy(), self.P.cellinfo_host, self.P.supp_jm.cpu().numpy())
        return self._host

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        slot, info, jm = self._tables()
        s = int(slot[i])
        if s < 0:
            return []
        ids = jm[info[s, 4]: info[s, 4] + info[s, 5]]
        out = []
        for f in ids:
            lev = int(np.searchsorted(self.P.fn_off, f, side="right")) - 1
            out.append((lev, tuple(int(x) for x in np.unravel_index(int(f - self.P.fn_off[lev]), self.P.sh_fns[lev]))))
        return out

    def active_cells(self):
        """[(level, cellIdx)] per point -- the first return value of compute_active_span_v2 (core.py:117-154)."""
        slot, info, _ = self._tables()
        d = self.P.ndim
        return [(int(info[s, 0]), tuple(int(x) for x in info[s, 1:1 + d])) if s >= 0 else None for s in slot]

    def offsets(self):
        """exclusive prefix sum of num_supp (int64, device) and nnz (python int; synchronises once)"""
        if getattr(self, "_off", None) is None:
            inc = torch.cumsum(self.num_supp.long(), dim=0)
            self._nnz = int(inc[-1]) if len(inc) else 0
            self._off = (inc - self.num_supp.long()).contiguous()
            self._cumsum = inc
        return self._off, self._nnz

    def Jm(self, dtype=torch.int64):
        """flat support ids for every (point, support) pair [THB/core.py:644-650]"""
        off, nnz = self.offsets()
        return engine.expand_supports(self.P, self.slot, off, nnz, dtype)


class NumSupp(Sequence):
    """num_supp of compute_active_span: list-like view of an int32 device tensor (.tensor)."""

    def __init__(self, t):
        self.tensor = t
        self._host = None

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, i):
        if self._host is None:
            self._host = self.tensor.cpu().numpy()
        v = self._host[i]
        return int(v) if np.ndim(v) == 0 else v.tolist()

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


def _packed_from_dict(knotvectors, cells, degrees, ac_dict, device):
    """Legacy: the caller holds the reference's dict {lev: {cell: [(fn_lev, fnIdx)]}}.  Only the support
    structure is packed (function masks are recovered from the lists); small spaces only."""
    L = max(cells.keys()) + 1
    d = len(degrees)
    fns = {}
    for lev in range(L):
        sh = tuple(len(np.unique(knotvectors[lev][k])) - 1 + degrees[k] for k in range(d))
        fns[lev] = np.zeros(sh, dtype=np.uint8)
    for lev in ac_dict:
        for cell, supp in ac_dict[lev].items():
            for fl, fn in supp:
                fns[fl][tuple(fn)] = 1
    P = PackedHierarchy.build(knotvectors, degrees, cells, fns, None, device=device)
    # the dict is authoritative: verify that the rebuilt lists agree
    info, jm = P.cellinfo_host, P.supp_jm.cpu().numpy()
    for s in range(min(P.nslots, 64)):
        lev, c = int(info[s, 0]), tuple(int(x) for x in info[s, 1:1 + d])
        want = [int(P.fn_off[fl] + np.ravel_multi_index(tuple(fn), P.sh_fns[fl])) for fl, fn in ac_dict[lev][c]]
        if want != jm[info[s, 4]: info[s, 4] + info[s, 5]].tolist():
            raise RuntimeError("ac_cells_ac_supp is not in the reference's order; cannot pack it")
    return P


def compute_active_span(params, knotvectors, cells, degrees, ac_cells_ac_supp, device="cuda"):
    """[THB/core.py:157-189] -> (ac_cell_supp, num_supp).  GPU: finest-level span search per point, dyadic
    walk to the active cell, grouping of the points by cell (the plan reused by THB_basis_fns)."""
    if isinstance(ac_cells_ac_supp, ActiveSupports):
        P = ac_cells_ac_supp.packed(knotvectors, device)
    elif isinstance(ac_cells_ac_supp, PackedHierarchy):
        P = ac_cells_ac_supp
    else:
        P = _packed_from_dict(knotvectors, cells, degrees, ac_cells_ac_supp, device)
    # this plan feeds THB_basis_fns (k_basis_rows), which is fastest with 256 points per work item (1.72 against 2.13 ms at
    # 512 on config 3); the fused path plans its own batches with engine.POINTS_PER_ITEM
    plan = engine.Plan(P, params, points_per_item=min(256, engine.POINTS_PER_ITEM or 256))
    ns = plan.num_supp()
    supp = PointSupports(P, plan, ns)
    return supp, NumSupp(ns)


def compute_active_span_v2(params, knotvectors, cells, degrees, ac_cells_ac_supp, fn_coeffs=None, device="cuda"):
    """[THB/core.py:117-154] -> ([(lev, cellIdx)], num_supp, supports object)"""
    supp, ns = compute_active_span(params, knotvectors, cells, degrees, ac_cells_ac_supp, device)
    return supp.active_cells(), ns, supp


def _channels(return_mode, ndim):
    v, dmode = bool(return_mode[0]), return_mode[1]
    ch = []
    if v:
        ch.append(engine.CH_VALUE)
    if dmode == "grad":
        ch += engine.grad_channels(ndim)
    elif dmode:
        ch.append(engine.mixed_channel(ndim))
    if not ch:
        raise ValueError("return_mode selects nothing")
    return ch


def THB_basis_fns(params, ac_cells_supp, num_supp, fn_coeffs, fn_shapes, knotvectors, degrees, return_mode=[True, False]):
    """[THB/core.py:604-701] flat THB basis values for every (point, support) pair (float32 CUDA tensor).

    return_mode: [True, False] values; [False, True] the reference's "derivative" (product of the 1-D first
    derivatives in every dimension, i.e. the mixed partial -- SURVEY Q12); [True, True] both.  Extension:
    return_mode[1] == "grad" returns the true gradient as an [nnz, d] tensor."""
    ndim = len(degrees)
    ch = _channels(return_mode, ndim)
    if isinstance(fn_coeffs, (RefinementOperators, PackedHierarchy)):
        if not isinstance(ac_cells_supp, PointSupports):
            raise RuntimeError("packed coefficient operators need the ac_cell_supp object returned by compute_active_span")
        P = ac_cells_supp.P
        if isinstance(fn_coeffs, RefinementOperators) and not P.has_tiles:
            P.attach_tiles(fn_coeffs.coeffs)
        plan = ac_cells_supp.plan
        off, nnz = ac_cells_supp.offsets()
        outs = engine.basis_tiled(P, plan, off, nnz, ch)
    else:
        outs = _basis_dense_legacy(params, ac_cells_supp, num_supp, fn_coeffs, fn_shapes, knotvectors, degrees, ch)
    res = []
    k = 0
    if return_mode[0]:
        res.append(outs[k]); k += 1
    if return_mode[1] == "grad":
        res.append(torch.stack(outs[k:k + ndim], dim=1))
    elif return_mode[1]:
        res.append(outs[k])
    return res[0] if len(res) == 1 else tuple(res)


def _basis_dense_legacy(params, ac_cells_supp, num_supp, fn_coeffs, fn_shapes, knotvectors, degrees, channels):
    """fn_coeffs is the reference's dict of dense tables [*sh_fns[lev], *sh_fns[max_lev]]."""
    import ctypes as C

    from ._lib import check, lib, ptr, stream_ptr

    dev = torch.device("cuda")
    L = max(knotvectors.keys())
    ndim = len(degrees)
    n_fns = np.zeros(L + 2, dtype=np.int64)
    for lev in range(L + 1):
        n_fns[lev + 1] = n_fns[lev] + int(np.prod(fn_shapes[lev]))
    row = int(np.prod(fn_shapes[L]))
    flat = torch.cat([torch.as_tensor(np.asarray(fn_coeffs[lev], dtype=np.float32)).reshape(-1, row) for lev in range(L + 1)]).to(dev)
    if isinstance(ac_cells_supp, PointSupports):
        jm = ac_cells_supp.Jm()
        ns = ac_cells_supp.num_supp.long()
    else:
        jm = torch.tensor([int(n_fns[fl] + np.ravel_multi_index(tuple(fn), fn_shapes[fl])) for supp in ac_cells_supp for fl, fn in supp],
                          dtype=torch.int64, device=dev)
        ns = torch.as_tensor(np.asarray(num_supp, dtype=np.int64), device=dev)
    cumsum = torch.cumsum(ns, 0).contiguous()
    nnz = jm.shape[0]
    prm = engine._as_params(params, ndim, dev)
    kn = [np.asarray(knotvectors[L][k], dtype=np.float32) for k in range(ndim)]
    koff = np.cumsum([0] + [len(k) for k in kn[:-1]]).astype(np.int32)
    knots = torch.from_numpy(np.concatenate(kn)).to(dev)
    I3 = C.c_int32 * 3
    pad = lambda v: I3(*[int(x) for x in v] + [0] * (3 - len(v)))
    outs = []
    for ch in channels:
        phi = torch.zeros(nnz, dtype=torch.float32, device=dev)
        check(lib.thb_basis_dense(ptr(prm), prm.shape[0], ndim, pad(degrees), ptr(knots), pad(koff), pad([len(k) for k in kn]),
                                  pad(fn_shapes[L]), ptr(flat), ptr(jm), 1, ptr(cumsum), int(ch), ptr(phi), stream_ptr(dev)), "thb_basis_dense")
        outs.append(phi)
    return outs


def compute_THB_fns_tp(params, ac_spans, ac_cells_supp, fn_coeffs, fn_shapes, knotvectors, degrees):
    """[THB/core.py:203-268] -> (PHI, num_supp); same numbers as THB_basis_fns, kept for API parity.
    `ac_cells_supp` must be an ActiveSupports / RefinementOperators pair produced by this package."""
    cells = ac_cells_supp.cells
    supp, ns = compute_active_span(params, knotvectors, cells, degrees, ac_cells_supp)
    return THB_basis_fns(params, supp, ns, fn_coeffs, fn_shapes, knotvectors, degrees), ns


def THB_basis_fns_tp_parallel(params, ac_cells_supp, fn_coeffs, fn_shapes, knotvectors, degrees):
    """[THB/core.py:573-601] the reference's process-pool variant of the basis loop; here the same GPU kernel
    as THB_basis_fns (`ac_cells_supp` is the per-point support object of compute_active_span)."""
    return THB_basis_fns(params, ac_cells_supp, getattr(ac_cells_supp, "num_supp", None), fn_coeffs, fn_shapes, knotvectors, degrees)
