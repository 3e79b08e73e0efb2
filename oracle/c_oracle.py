"""ctypes wrapper of the plain-C oracle (oracle/thb_oracle_c.c).  Test infrastructure: used by tests/, by
__graft_entry__.smoke() and by bench.py's CPU-baseline / --impl reference legs only."""
import ctypes as C

import numpy as np

from . import build_c

MAXL = 8


class OracleSpace(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nlevels", C.c_int32), ("degree", C.c_int32 * 3), ("tile_stride", C.c_int32), ("info_w", C.c_int32),
        ("knot_off", (C.c_int32 * 3) * MAXL), ("knot_len", (C.c_int32 * 3) * MAXL), ("ncells", (C.c_int32 * 3) * MAXL),
        ("cell_slot_off", C.c_int64 * MAXL),
        ("knots", C.c_void_p), ("cell_slot", C.c_void_p), ("cellinfo", C.c_void_p), ("dmask", C.c_void_p),
        ("supp_jm", C.c_void_p), ("tiles", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.oracle_thb_eval.restype = C.c_int64
        _lib.oracle_thb_eval.argtypes = [C.POINTER(OracleSpace), C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_find_spans.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_basis_funs.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_eval_forward.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int]
        _lib.oracle_eval_backward.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_int]
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def use_all_host_threads():
    """OpenMP threads = the CPUs this process may run on, whatever OMP_NUM_THREADS a launcher exported (torchrun sets 1)."""
    import os

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_threads(C.c_int(n))
    return num_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostTables:
    """Flat hierarchy tables as host numpy arrays (layout documented in include/thb_b200.h)."""

    def __init__(self, ndim, nlevels, degrees, tile_stride, knot_off, knot_len, ncells, cell_slot_off, knots, cell_slot, cellinfo, dmask,
                 supp_jm, tiles):
        self.keep = [np.ascontiguousarray(a) for a in (knots, cell_slot, cellinfo, dmask, supp_jm, tiles)]
        s = OracleSpace()
        s.ndim, s.nlevels, s.tile_stride = ndim, nlevels, tile_stride
        s.info_w = int(self.keep[2].shape[1])
        for k in range(3):
            s.degree[k] = int(degrees[k]) if k < ndim else 0
        for l in range(nlevels):
            for k in range(ndim):
                s.knot_off[l][k], s.knot_len[l][k], s.ncells[l][k] = int(knot_off[l][k]), int(knot_len[l][k]), int(ncells[l][k])
            if ndim == 2:
                s.ncells[l][2] = 1
            s.cell_slot_off[l] = int(cell_slot_off[l])
        s.knots, s.cell_slot, s.cellinfo, s.dmask, s.supp_jm, s.tiles = [a.ctypes.data for a in self.keep]
        self.desc = s
        self.ndim = ndim

    def eval(self, params, ctrl_pts=None, mode=0, want_phi=False, offsets=None, nnz=0):
        """-> dict(out [n,cpd] f64, slot [n] i32, phi [nnz] f64, inside)"""
        params = np.ascontiguousarray(params, dtype=np.float32)
        n = params.shape[0]
        res = {}
        out = slot = phi = None
        cpd = 0
        cp = None
        if ctrl_pts is not None:
            cp = np.ascontiguousarray(ctrl_pts, dtype=np.float32)
            cpd = cp.shape[1]
            out = np.zeros((n, cpd), dtype=np.float64)
        slot = np.zeros(n, dtype=np.int32)
        if want_phi:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            phi = np.zeros(nnz, dtype=np.float64)
        res["inside"] = int(lib().oracle_thb_eval(C.byref(self.desc), _p(params), n, int(mode), _p(cp) if cp is not None else None, cpd,
                                                  _p(out) if out is not None else None, _p(slot), _p(offsets) if want_phi else None,
                                                  _p(phi) if want_phi else None))
        res["out"], res["slot"], res["phi"] = out, slot, phi
        return res


def find_spans(params, knots, p):
    params = np.ascontiguousarray(params, dtype=np.float32)
    knots = np.ascontiguousarray(knots, dtype=np.float32)
    out = np.zeros(len(params), dtype=np.int32)
    lib().oracle_find_spans(_p(params), 1, len(params), _p(knots), len(knots), int(p), _p(out))
    return out


def basis_funs(span, u, p, U):
    U = np.ascontiguousarray(U, dtype=np.float64)
    N, dN = np.zeros(p + 1), np.zeros(p + 1)
    lib().oracle_basis_funs(int(span), float(u), int(p), _p(U), _p(N), _p(dN))
    return N, dN


def eval_forward(cp, Jm, phi, cumsum):
    cp, Jm, phi, cumsum = (np.ascontiguousarray(cp, dtype=np.float32), np.ascontiguousarray(Jm, dtype=np.int64),
                           np.ascontiguousarray(phi, dtype=np.float32), np.ascontiguousarray(cumsum, dtype=np.int64))
    out = np.zeros((len(cumsum), cp.shape[1]), dtype=np.float32)
    lib().oracle_eval_forward(_p(cp), _p(Jm), _p(phi), _p(cumsum), _p(out), len(cumsum), cp.shape[1])
    return out


def eval_backward(go, cp_shape, Jm, phi, cumsum):
    go, Jm, phi, cumsum = (np.ascontiguousarray(go, dtype=np.float32), np.ascontiguousarray(Jm, dtype=np.int64),
                           np.ascontiguousarray(phi, dtype=np.float32), np.ascontiguousarray(cumsum, dtype=np.int64))
    g = np.zeros(cp_shape, dtype=np.float32)
    lib().oracle_eval_backward(_p(go), _p(Jm), _p(phi), _p(cumsum), _p(g), len(cumsum), cp_shape[0], cp_shape[1])
    return g


class OracleDirect(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nlevels", C.c_int32), ("degree", C.c_int32 * 3),
        ("nfn", (C.c_int32 * 3) * MAXL), ("ncell", (C.c_int32 * 3) * MAXL), ("nk", (C.c_int32 * 3) * MAXL),
        ("knots", (C.c_void_p * 3) * MAXL), ("cells", C.c_void_p * MAXL), ("fns", C.c_void_p * MAXL),
        ("coeff", (C.c_void_p * 3) * MAXL), ("fn_off", C.c_int64 * MAXL),
    ]


class DirectTables:
    """Table-free evaluator in C (oracle_direct_eval): same inputs and results as thb_oracle.DirectEvaluator.point,
    for 1e4..1e6 points.  knotvectors / cells / fns: per-level dicts as in the reference's Space; Coeff[l][k] is the
    level l -> l+1 matrix [nfn_l, nfn_{l+1}] of dimension k (thb_oracle.knot_insertion_matrix(...).T)."""

    def __init__(self, knotvectors, degrees, cells, fns, Coeff, knots_as_f32=True):
        L = max(knotvectors.keys())
        d = len(degrees)
        s = OracleDirect()
        s.ndim, s.nlevels = d, L + 1
        self.keep = []
        off = 0
        for k in range(3):
            s.degree[k] = int(degrees[k]) if k < d else 0
        for l in range(L + 1):
            s.fn_off[l] = off
            off += int(np.prod(fns[l].shape))
            c8 = np.ascontiguousarray((np.asarray(cells[l]) == 1).astype(np.uint8))
            f8 = np.ascontiguousarray((np.asarray(fns[l]) == 1).astype(np.uint8))
            self.keep += [c8, f8]
            s.cells[l], s.fns[l] = c8.ctypes.data, f8.ctypes.data
            for k in range(d):
                U = np.asarray(knotvectors[l][k])
                # the product's canonical inputs are float32: spans are decided on the f32 knot values
                U = np.ascontiguousarray(U.astype(np.float32).astype(np.float64) if knots_as_f32 else U.astype(np.float64))
                self.keep.append(U)
                s.knots[l][k], s.nk[l][k] = U.ctypes.data, len(U)
                s.nfn[l][k], s.ncell[l][k] = fns[l].shape[k], cells[l].shape[k]
                if l < L:
                    M = np.ascontiguousarray(np.asarray(Coeff[l][k], dtype=np.float64))
                    assert M.shape == (fns[l].shape[k], fns[l + 1].shape[k])
                    self.keep.append(M)
                    s.coeff[l][k] = M.ctypes.data
        self.s, self.d = s, d
        fn = lib().oracle_direct_eval
        fn.restype = C.c_int64
        fn.argtypes = [C.POINTER(OracleDirect), C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 6

    def eval(self, x, max_supp=256, want_grad=False):
        """-> dict(lev [n], cell [n,3], nsupp [n], jm [n,max_supp], phi [n,max_supp](, dphi [n,max_supp,3]))"""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        n = len(x)
        lev = np.empty(n, np.int32)
        cell = np.empty((n, 3), np.int32)
        ns = np.empty(n, np.int32)
        jm = np.zeros((n, max_supp), np.int64)
        phi = np.zeros((n, max_supp), np.float64)
        dphi = np.zeros((n, max_supp, 3), np.float64) if want_grad else None
        worst = lib().oracle_direct_eval(C.byref(self.s), _p(x), n, max_supp, _p(lev), _p(cell), _p(ns), _p(jm), _p(phi),
                                         _p(dphi) if dphi is not None else None)
        if worst > max_supp:
            raise RuntimeError(f"max_supp too small: {worst}")
        r = dict(lev=lev, cell=cell, nsupp=ns, jm=jm, phi=phi)
        if want_grad:
            r["dphi"] = dphi
        return r
