"""Thin Python layer over the C ABI: allocates outputs with torch (so the caching allocator owns the
memory), passes raw device pointers and the current CUDA stream, raises RuntimeError on library errors.
No arithmetic happens here."""
import ctypes as C

import torch

from . import _lib
from ._lib import OrderedDesc, PlanDesc, check, lib, ptr, require_cuda, stream_ptr

import os as _os

# Points per work item (multiple of 64).  Measured on config 3 (16.8 M points), 128 / 256 / 512 / 1024: evaluation kernel 0.188 /
# 0.176 / 0.171 / 0.174 ms, fused backward 0.556 / 0.541 / 0.528 / 0.529 ms, pass 0.454 / 0.442 / 0.436 / 0.438 ms (an item
# switch costs a net load and a pipeline refill).  Small batches want more, smaller items for the ~2400 resident warps of the
# persistent kernels: with 512 the strong-scaling step of a 1/8 slab went from 0.117 to 0.132 ms and the 8-GPU fit from 5 045 to
# 4 564 steps/s.  THB_POINTS_PER_ITEM overrides the choice.
POINTS_PER_ITEM = int(_os.environ.get("THB_POINTS_PER_ITEM", "0"))   # 0: by batch size
LARGE_BATCH_ITEMS = 8_000_000


def points_per_item_for(n):
    if POINTS_PER_ITEM:
        return POINTS_PER_ITEM
    return 512 if n >= LARGE_BATCH_ITEMS else 256

CH_VALUE = 0
PLAN_HISTOGRAM, PLAN_ITEMS, PLAN_SCATTER = 1, 2, 4   # thb_plan_build_phase


def grad_channels(ndim):
    return [1 << k for k in range(ndim)]


def mixed_channel(ndim):
    return (1 << ndim) - 1


def _on(device):
    """Every library call runs with the tensors' device current: the library launches on the CURRENT device, so a
    hierarchy on cuda:1 used from a process whose current device is cuda:0 must switch first (stream handles, cached
    grids and kernel attributes are per device)."""
    return torch.cuda.device(device)


def _as_params(params, ndim, device):
    """[N, ndim] float32 contiguous CUDA tensor (float64 / numpy inputs are rounded to f32 -- SURVEY Q11)."""
    if not isinstance(params, torch.Tensor):
        import numpy as np

        params = torch.from_numpy(np.ascontiguousarray(np.asarray(params)))
    if params.dim() != 2 or params.shape[1] != ndim:
        raise RuntimeError(f"params must have shape [N, {ndim}]")
    return params.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


class Plan:
    """The points of one batch grouped by active cell (device resident)."""

    def __init__(self, P, params, points_per_item=None, capacity=None, light=None):
        """light (default off; THB_PLAN_LIGHT=1): thb_plan_build writes only the permutation of the points (4 bytes per
        point instead of a 16-byte record, and it does not re-read the parameters); eval_nets gathers the parameters
        through it.  Every other consumer (PHI rows, the per-point formulations, the backward, caller_index) materialises
        the records on first use (ensure_records).  Measured on config 3: plan 0.207 -> 0.154 ms, evaluation kernel 0.176
        -> 0.207 ms (three 4-byte gathers per point instead of one 16-byte record), pass 0.467 -> 0.443 ms."""
        self.P = P
        self.params = _as_params(params, P.ndim, P.device)
        if points_per_item is None:
            points_per_item = points_per_item_for(max(int(self.params.shape[0]), int(capacity or 0)))
        if light is None:
            light = _os.environ.get("THB_PLAN_LIGHT", "0") not in ("0", "")
        if light and points_per_item > 256 and getattr(P, "poly", None) is not None:
            points_per_item = 256   # a light plan's items are gathered through index buffers of 256 entries
        self.light = bool(light) and points_per_item <= 256 and getattr(P, "poly", None) is not None
        n = self.params.shape[0]
        dev = P.device
        i32 = dict(dtype=torch.int32, device=dev)
        self.n = n
        self.capacity = cap = max(n, int(capacity or 0), 1)   # buffers hold up to `capacity` points (see rebind)
        self.slot = torch.empty(cap, **i32)
        self.scratch = torch.empty(cap, **i32)
        self.sorted_params = torch.empty((cap, 4), dtype=torch.float32, device=dev)  # (u0, u1, u2|0, caller index bits)
        self.cell_count = torch.empty(P.nslots + 1, **i32)
        self.cell_start = torch.empty(P.nslots + 1, **i32)
        self.cell_cursor = torch.empty(P.nslots + 1, **i32)
        self.max_items = cap // points_per_item + P.nslots + 1
        self.items = torch.empty((self.max_items, 16), **i32)
        self.counters = torch.zeros(4, **i32)
        d = PlanDesc()
        d.npoints, d.points_per_item, d.max_items = n, points_per_item, self.max_items
        d.slot, d.scratch, d.sorted_params = self.slot.data_ptr(), self.scratch.data_ptr(), self.sorted_params.data_ptr()
        d.cell_count, d.cell_start, d.cell_cursor = self.cell_count.data_ptr(), self.cell_start.data_ptr(), self.cell_cursor.data_ptr()
        d.items, d.counters = self.items.data_ptr(), self.counters.data_ptr()
        self.desc = d
        self.build()

    def build(self, phase=None):
        """phase None: the whole plan; else a combination of PLAN_HISTOGRAM, PLAN_ITEMS, PLAN_SCATTER (thb_plan_build_phase),
        issued in that order over the calls."""
        self.version = self.P.version
        self.desc.light = 1 if self.light else 0
        self.desc.params = self.params.data_ptr()
        with _on(self.P.device):
            if phase is None:
                check(lib.thb_plan_build(C.byref(self.P.desc()), ptr(self.params), C.byref(self.desc), stream_ptr(self.P.device)), "thb_plan_build")
            else:
                check(lib.thb_plan_build_phase(C.byref(self.P.desc()), ptr(self.params), C.byref(self.desc), int(phase),
                                               stream_ptr(self.P.device)), "thb_plan_build_phase")

    def ensure_records(self):
        """The 16-byte point records of a light plan (thb_plan_records), built once per plan build."""
        if self.desc.light == 1:
            with _on(self.P.device):
                check(lib.thb_plan_records(C.byref(self.P.desc()), ptr(self.params), C.byref(self.desc), stream_ptr(self.P.device)), "thb_plan_records")
            self.desc.light = 2
        return self

    def rebind(self, params):
        """Plans another batch (at most `capacity` points) in the same buffers."""
        params = _as_params(params, self.P.ndim, self.P.device)
        if params.shape[0] > self.capacity:
            raise RuntimeError("batch larger than the plan's capacity")
        self.params, self.n = params, params.shape[0]
        self.desc.npoints = self.n
        self.build()
        return self

    def set_rows_in_plan_order(self, flag=True):
        """Rows of `out` / `grad_out` of the nets kernels follow the plan's (cell) order instead of the caller's: row j
        belongs to point caller_index()[j].  Coalesced rows for callers that keep per-point data (targets) in plan
        order; with scattered points it roughly halves the evaluation and backward times."""
        self.desc.rows_in_plan_order = 1 if flag else 0
        return self

    def caller_index(self):
        """int64 [n_inside]: caller-order index of the point in row j of the plan order (synchronises once)."""
        self.ensure_records()
        n_in = int(self.counters[2])
        return self.sorted_params[:n_in, 3].contiguous().view(torch.int32).long()

    def canonicalize(self):
        """Puts the points of every cell in a CANONICAL order (increasing caller index).  thb_plan_build places the points
        of a cell in the order its scatter atomics happen to arrive, which differs from run to run; sums over the points
        of a work item (the fused backward) inherit that order.  After this call two plans of the same batch are
        bit-identical, and with the ordered backward (deterministic=True) so are the gradients.  One device sort of the
        batch (torch.sort, stable radix sort); call it once for a fit, whose plan is built once."""
        self.ensure_records()
        n_in = int(self.counters[2])
        if n_in == 0:
            return self
        rec = self.sorted_params[:n_in]
        idx = rec[:, 3].contiguous().view(torch.int32).long()
        key = (self.slot[: self.n].long()[idx] << 32) | idx          # records are already grouped by cell: the cell order is kept
        order = torch.argsort(key, stable=True)
        self.sorted_params[:n_in] = rec[order]
        if self.light:   # keep the permutation of a light plan in step with its records
            self.scratch[:n_in] = self.sorted_params[:n_in, 3].contiguous().view(torch.int32)
        return self

    def refresh(self):
        """Work items embed per-cell table offsets: rebuild if the hierarchy's tables changed since build()."""
        if self.version != self.P.version:
            self.build()
        return self

    # derived quantities (device tensors; no sync)
    def num_supp(self):
        S = self.P.cellinfo[:, 5]
        s = self.slot[: self.n].long()
        return torch.where(s >= 0, S[s.clamp(min=0)], torch.zeros_like(s, dtype=torch.int32))

    def stats(self, with_sep=False):
        """(#in-domain points, nnz, nnz of tiled supports[, nnz of rank-one tiled supports]) -- synchronises."""
        cnt = self.cell_count[: self.P.nslots].long()
        S = self.P.cellinfo[:, 5].long()
        nd = self.P.cellinfo[:, 7].long()
        res = (int(cnt.sum()), int((cnt * S).sum()), int((cnt * (S - nd)).sum()))
        if with_sep:
            res = res + (int((cnt * self.P.cellinfo[:, 9].long()).sum()),)
        return res


def active_span(P, params, want_num_supp=True):
    params = _as_params(params, P.ndim, P.device)
    n = params.shape[0]
    slot = torch.empty(n, dtype=torch.int32, device=P.device)
    ns = torch.empty(n, dtype=torch.int32, device=P.device) if want_num_supp else None
    with _on(P.device):
        check(lib.thb_active_span(C.byref(P.desc()), ptr(params), n, ptr(slot), ptr(ns), stream_ptr(P.device)), "thb_active_span")
    return slot, ns


def expand_supports(P, slot, offsets, nnz, dtype=torch.int64):
    jm = torch.empty(nnz, dtype=dtype, device=P.device)
    with _on(P.device):
        check(lib.thb_expand_supports(C.byref(P.desc()), ptr(slot), ptr(offsets), slot.shape[0], ptr(jm), int(dtype == torch.int64),
                                      stream_ptr(P.device)), "thb_expand_supports")
    return jm


def _chan_array(channels):
    return (C.c_int32 * len(channels))(*[int(c) for c in channels])


def basis_tiled(P, plan, offsets, nnz, channels=(CH_VALUE,)):
    """PHI for each channel, flat [nnz] in the reference's (point-major) pair order."""
    if not P.has_tiles:
        raise RuntimeError("the hierarchy has no coefficient tiles yet (PackedHierarchy.attach_tiles / from_space)")
    plan.refresh().ensure_records()
    outs = [torch.empty(nnz, dtype=torch.float32, device=P.device) for _ in channels]  # every entry is written
    arr = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    with _on(P.device):
        check(lib.thb_basis_tiled(C.byref(P.desc()), C.byref(plan.desc), ptr(plan.params), ptr(offsets), _chan_array(channels),
                                  len(channels), arr, stream_ptr(P.device)), "thb_basis_tiled")
    return outs


FORMULATIONS = {"local_net": 0, "tile_net": 1, "per_point": 2}


def _same_device(a, b):
    a, b = torch.device(a), torch.device(b)
    if a.type != b.type:
        return False
    ia = a.index if a.index is not None else torch.cuda.current_device()
    ib = b.index if b.index is not None else torch.cuda.current_device()
    return ia == ib


def _check_rows(tensors, plan, cpd, name):
    """Raw pointers go to the library: every per-point buffer must be a contiguous f32 [plan.n, cp_dim] tensor on the
    hierarchy's device."""
    for t in tensors:
        require_cuda(t, name, torch.float32)
        if not _same_device(t.device, plan.P.device) or t.dim() != 2 or t.shape[0] < plan.n or t.shape[1] != cpd:
            raise RuntimeError(f"{name} must be [{plan.n}, {cpd}] float32 on {plan.P.device}, got {tuple(t.shape)} on {t.device}")


NETS_DERIVATIVES, NETS_MONOMIAL = 1, 2


def _net_flags(P, with_derivatives, monomial=None):
    """flags of thb_cell_nets / thb_eval_nets: monomial nets by default whenever the hierarchy carries the span
    polynomials (THB_NET_MONO=0 switches back to control-point nets evaluated with Cox-de Boor)."""
    if monomial is None:
        monomial = getattr(P, "poly", None) is not None and _os.environ.get("THB_NET_MONO", "1") != "0"
    return (NETS_DERIVATIVES if with_derivatives else 0) | (NETS_MONOMIAL if monomial else 0)


def net_workspace(P, with_derivatives=False):
    """Device buffer for the local control nets of P ([nslots][4 or 8][TS] f32), allocated once per hierarchy."""
    key = "_net_ws8" if with_derivatives else "_net_ws4"
    ws = getattr(P, key, None)
    n = lib.thb_net_workspace_floats(C.byref(P.desc()), NETS_DERIVATIVES if with_derivatives else 0)   # control-point form: the larger
    if ws is None or ws.numel() != n:
        ws = torch.empty(n, dtype=torch.float32, device=P.device)
        setattr(P, key, ws)
    return ws


def cell_nets(P, ctrl_pts, with_derivatives=False, plan=None, out=None, monomial=None, ctas_per_sm=0):
    """Local control nets W of the active cells (thb_cell_nets): every cell, or only those holding points of `plan`.
    They stay valid for as long as ctrl_pts (and the hierarchy) do not change.  monomial (default: yes when available):
    the nets are stored as the coefficients of the cell's polynomial in cell-local coordinates, evaluated by Horner."""
    require_cuda(ctrl_pts, "ctrl_pts", torch.float32)
    if ctrl_pts.shape[1] not in (3, 4) or ctrl_pts.shape[0] != P.nCP:
        raise RuntimeError(f"ctrl_pts must be [{P.nCP}, 3 or 4]")
    if not P.has_tiles:
        raise RuntimeError("the hierarchy has no coefficient tiles yet (PackedHierarchy.attach_tiles / from_space)")
    nets = out if out is not None else net_workspace(P, with_derivatives)
    flags = _net_flags(P, with_derivatives, monomial)
    if plan is not None:
        plan.refresh()
    with _on(P.device):
        check(lib.thb_cell_nets(C.byref(P.desc()), C.byref(plan.desc) if plan is not None else None, ptr(ctrl_pts), ctrl_pts.shape[1],
                                flags | ((int(ctas_per_sm) & 0xff) << 8), ptr(nets), stream_ptr(P.device)), "thb_cell_nets")
    nets._thb_flags = flags
    return nets


SMALL_BATCH = int(_os.environ.get("THB_SMALL_BATCH", "12000000"))   # plan_and_nets(overlap="auto"): below this many points the
                                                                    # nets go beside the plan's later phases


def plan_and_nets(P, plan, ctrl_pts, with_derivatives=False, overlap="auto"):
    """plan.build() and the local control nets of ctrl_pts, enqueued so that they run CONCURRENTLY on two streams; the current
    stream then waits for the nets.  Returns the nets.  overlap:
      True               the nets of EVERY cell beside the whole plan (they depend on the hierarchy and the control points only).
                         Best for large batches: 0.443 against 0.477 ms serial on config 3 (16.8 M points); both kernels fill the
                         SMs, so in effect the nets run first and their tail shares the machine with the histogram pass.
      "after_histogram"  histogram pass first, then the nets of the cells that HOLD POINTS beside scan + work items + scatter.
                         Best below the full batch (a rank's slab, a chunk), whose kernels run at their latency and leave the SMs
                         half empty -- measured on slabs of the config-3 grid, serial / this / True: 1/8 (1.4 M points, 14 k of
                         52 k cells) 0.114 / 0.103 / 0.160 ms, 1/4 0.172 / 0.163 / 0.192, 1/2 0.284 / 0.279 / 0.297; for the full
                         batch it equals the serial order (0.485 against 0.450 ms).
      "scatter"          the nets beside the scatter pass alone (measured: never better than the other two).
      "auto" (default)   True from SMALL_BATCH (12 M) points up, "after_histogram" below.
      False              plan first, then the nets of the cells that hold points, on one stream."""
    if overlap == "auto":
        overlap = True if plan.n >= SMALL_BATCH else "after_histogram"
    if not overlap:
        plan.build()
        return cell_nets(P, ctrl_pts, with_derivatives, plan=plan)
    cur = torch.cuda.current_stream(P.device)
    side = getattr(P, "_side_stream", None)
    if side is None:
        side = P._side_stream = torch.cuda.Stream(device=P.device)
    if overlap in ("after_histogram", "scatter"):
        first = PLAN_HISTOGRAM if overlap == "after_histogram" else PLAN_HISTOGRAM | PLAN_ITEMS
        plan.build(phase=first)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            nets = cell_nets(P, ctrl_pts, with_derivatives, plan=plan, ctas_per_sm=int(_os.environ.get("THB_OVERLAP_NETS_CTAS", "0")))
        plan.build(phase=(PLAN_HISTOGRAM | PLAN_ITEMS | PLAN_SCATTER) & ~first)
        cur.wait_stream(side)
        return nets
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        # (capping the resident CTAs of the nets kernel -- THB_NETS_CTAS -- so that the plan kernels find registers beside
        # it was measured on config 3: 0 = uncapped 0.467 ms per pass, 3: 0.461, 4-6: 0.468, 2: 0.559)
        nets = cell_nets(P, ctrl_pts, with_derivatives, ctas_per_sm=int(_os.environ.get("THB_OVERLAP_NETS_CTAS", "0")))
    plan.build()
    cur.wait_stream(side)
    return nets


def eval_nets(P, plan, nets, cp_dim=3, channels=(CH_VALUE,), out=None, with_derivatives=None, monomial=None):
    """Evaluates the plan's points against prebuilt nets (thb_eval_nets).  The layout of `nets` (derivative planes,
    monomial form) is the one cell_nets recorded on the tensor; pass with_derivatives / monomial for foreign buffers."""
    flags = getattr(nets, "_thb_flags", None)
    if flags is None or with_derivatives is not None or monomial is not None:
        if with_derivatives is None:
            with_derivatives = bool(flags & NETS_DERIVATIVES) if flags is not None else \
                nets.numel() == lib.thb_net_workspace_floats(C.byref(P.desc()), NETS_DERIVATIVES)
        if monomial is None:
            monomial = bool(flags & NETS_MONOMIAL) if flags is not None else False
        flags = (NETS_DERIVATIVES if with_derivatives else 0) | (NETS_MONOMIAL if monomial else 0)
    plan.refresh()
    if not flags & NETS_MONOMIAL:
        plan.ensure_records()   # a light plan is gathered through its permutation by the monomial kernels only
    if out is None:
        out = [torch.zeros((plan.n, cp_dim), dtype=torch.float32, device=P.device) for _ in channels]
    if len(out) != len(channels):
        raise RuntimeError("one output buffer per channel")
    _check_rows(out, plan, cp_dim, "out")
    arr = (C.c_void_p * len(out))(*[o.data_ptr() for o in out])
    with _on(P.device):
        check(lib.thb_eval_nets(C.byref(P.desc()), C.byref(plan.desc), ptr(nets), flags, cp_dim, _chan_array(channels),
                                len(channels), arr, stream_ptr(P.device)), "thb_eval_nets")
    return out


def eval_fused(P, plan, ctrl_pts, channels=(CH_VALUE,), cellnet=True, out=None, formulation=None):
    """Fused THB_basis_fns + THBEval.forward.  formulation: "local_net" (default; cellnet=True), "per_point"
    (cellnet=False: PHI of every support is formed per point, the reference's formulation) or "tile_net"."""
    form = FORMULATIONS[formulation] if formulation is not None else (0 if cellnet else 2)
    require_cuda(ctrl_pts, "ctrl_pts", torch.float32)
    cpd = ctrl_pts.shape[1]
    if cpd not in (3, 4):
        raise RuntimeError("fused evaluation supports control points of width 3 or 4")
    if ctrl_pts.shape[0] != P.nCP:
        raise RuntimeError(f"ctrl_pts must have {P.nCP} rows")
    if not P.has_tiles:
        raise RuntimeError("the hierarchy has no coefficient tiles yet (PackedHierarchy.attach_tiles / from_space)")
    plan.refresh()
    if form != 0 or not (_net_flags(P, False) & NETS_MONOMIAL):
        plan.ensure_records()
    if out is None:
        out = [torch.zeros((plan.n, cpd), dtype=torch.float32, device=P.device) for _ in channels]
    if len(out) != len(channels):
        raise RuntimeError("one output buffer per channel")
    _check_rows(out, plan, cpd, "out")
    arr = (C.c_void_p * len(out))(*[o.data_ptr() for o in out])
    ws = net_workspace(P, any(int(c) != 0 for c in channels)) if form == 0 else None
    with _on(P.device):
        check(lib.thb_eval_fused(C.byref(P.desc()), C.byref(plan.desc), ptr(plan.params), ptr(ctrl_pts), cpd, _chan_array(channels),
                                 len(channels), arr, form, ptr(ws) if ws is not None else None, stream_ptr(P.device)), "thb_eval_fused")
    return out


def eval_fused_backward(P, plan, grad_out, cp_dim=None, formulation="local_net", deterministic=None):
    """grad_ctrl_pts = A^T grad_out for the plan's points.  "local_net" (default): per-item reduction into per-cell net
    gradients + one transposed fold per cell; "per_point": the per-item tile kernel.  deterministic=True (default from
    THB_DETERMINISTIC): the ordered form of "local_net" -- no atomics, every sum in a fixed order; with a canonical plan
    (Plan.canonicalize) the result is bit-identical from run to run."""
    require_cuda(grad_out, "grad_out", torch.float32)
    cpd = grad_out.shape[1]
    if cpd not in (3, 4):
        raise RuntimeError("fused backward supports control points of width 3 or 4")
    if not P.has_tiles:
        raise RuntimeError("the hierarchy has no coefficient tiles yet (PackedHierarchy.attach_tiles / from_space)")
    plan.refresh().ensure_records()
    _check_rows([grad_out], plan, cpd, "grad_out")
    if deterministic is None:
        deterministic = deterministic_default()
    if deterministic:
        if formulation != "local_net":
            raise RuntimeError("the ordered (deterministic) backward exists for the local-net formulation only")
        return _ordered_backward(P, plan, grad_out, cpd, dense=True)
    g = torch.empty((P.nCP, cpd), dtype=torch.float32, device=P.device)
    ws = _gnet_workspace(P) if formulation == "local_net" else None
    with _on(P.device):
        check(lib.thb_eval_fused_backward(C.byref(P.desc()), C.byref(plan.desc), ptr(plan.params), ptr(grad_out), cpd, ptr(g), P.nCP,
                                          ptr(ws) if ws is not None else None, stream_ptr(P.device)), "thb_eval_fused_backward")
    return g


def deterministic_default():
    """THB_DETERMINISTIC=1 makes the ordered backward the default of eval_fused_backward / eval_fused_backward_rows."""
    return _os.environ.get("THB_DETERMINISTIC", "0") not in ("0", "")


def _ordered_backward(P, plan, grad_out, cpd, dense):
    """thb_eval_fused_backward_ordered: every sum in a fixed order (no atomics).  dense: [nCP, cpd], else compact rows."""
    tr_ptr, tr_idx = P.ordered_scatter_tables()
    ids, _ = P.active_rows()
    ws = getattr(P, "_ordered_ws", None)
    key = (plan.max_items, cpd)
    if ws is None or ws[0] != key:
        ws = P._ordered_ws = (key,
                              torch.empty((plan.max_items, 4, P.TS), dtype=torch.float32, device=P.device),
                              torch.zeros(P.nslots, dtype=torch.int32, device=P.device),
                              torch.empty((P.supp_jm.shape[0], cpd), dtype=torch.float32, device=P.device))
    od = OrderedDesc()
    od.item_ws, od.item0, od.contrib = ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr()
    od.tr_ptr, od.tr_idx = tr_ptr.data_ptr(), tr_idx.data_ptr()
    od.active_ids = ids.data_ptr() if dense else None
    od.nsupp, od.nrows = int(P.supp_jm.shape[0]), int(ids.shape[0])
    g = torch.empty((P.nCP if dense else ids.shape[0], cpd), dtype=torch.float32, device=P.device)
    with _on(P.device):
        check(lib.thb_eval_fused_backward_ordered(C.byref(P.desc()), C.byref(plan.desc), ptr(plan.params), ptr(grad_out), cpd, C.byref(od),
                                                  ptr(g), g.shape[0], ptr(_gnet_workspace(P)), stream_ptr(P.device)),
              "thb_eval_fused_backward_ordered")
    return g


def _gnet_workspace(P):
    ws = getattr(P, "_gnet_ws", None)
    n = lib.thb_net_workspace_floats(C.byref(P.desc()), 0)   # net GRADIENTS: always the control-point form, four planes
    if ws is None or ws.numel() != n:
        ws = P._gnet_ws = torch.zeros(n, dtype=torch.float32, device=P.device)   # zero before the first call; self-cleaning after
    return ws


def eval_fused_backward_rows(P, plan, grad_out, out=None, deterministic=None):
    """The same gradient in COMPACT rows: [P.n_active, cp_dim], row i = flat function id P.active_rows()[0][i]
    (thb_eval_fused_backward_rows).  Only active functions receive a gradient, so this is everything there is."""
    require_cuda(grad_out, "grad_out", torch.float32)
    cpd = grad_out.shape[1]
    if cpd not in (3, 4):
        raise RuntimeError("fused backward supports control points of width 3 or 4")
    if not P.has_tiles:
        raise RuntimeError("the hierarchy has no coefficient tiles yet (PackedHierarchy.attach_tiles / from_space)")
    plan.refresh().ensure_records()
    _check_rows([grad_out], plan, cpd, "grad_out")
    ids, row_map = P.active_rows()
    if deterministic is None:
        deterministic = deterministic_default()
    if deterministic:
        g = _ordered_backward(P, plan, grad_out, cpd, dense=False)
        if out is not None:
            out.copy_(g)
            return out
        return g
    g = out if out is not None else torch.empty((ids.shape[0], cpd), dtype=torch.float32, device=P.device)
    if g.shape != (ids.shape[0], cpd) or not g.is_contiguous() or not _same_device(g.device, P.device) or g.dtype != torch.float32:
        raise RuntimeError(f"out must be a contiguous float32 [{ids.shape[0]}, {cpd}] tensor on {P.device}")
    with _on(P.device):
        check(lib.thb_eval_fused_backward_rows(C.byref(P.desc()), C.byref(plan.desc), ptr(plan.params), ptr(grad_out), cpd, ptr(row_map),
                                               ptr(g), ids.shape[0], ptr(_gnet_workspace(P)), stream_ptr(P.device)), "thb_eval_fused_backward_rows")
    return g


def mse_workspace(device):
    return torch.zeros(int(lib.thb_mse_workspace_bytes()), dtype=torch.uint8, device=device)


def mse_grad(out, target, scale, grad_out, loss, workspace):
    """loss[0] = scale * sum (out - target)^2, grad_out = 2 scale (out - target) in one pass (thb_mse_grad)."""
    for t, name in ((out, "out"), (target, "target"), (grad_out, "grad_out")):
        require_cuda(t, name, torch.float32)
    if out.shape != target.shape or out.shape != grad_out.shape:
        raise RuntimeError("out, target and grad_out must have the same shape")
    with _on(out.device):
        check(lib.thb_mse_grad(ptr(out), ptr(target), out.numel(), float(scale), ptr(grad_out), ptr(loss), ptr(workspace),
                               stream_ptr(out.device)), "thb_mse_grad")
    return loss


def adam_rows(params, row_ids, grad_rows, exp_avg, exp_avg_sq, lr, betas, eps, step=0, step_dev=None, grad_scale=1.0):
    """torch.optim.Adam's update on the rows `row_ids` of the dense parameter `params` (thb_adam_rows)."""
    require_cuda(params, "params", torch.float32)
    require_cuda(row_ids, "row_ids", torch.int32)
    for t, name in ((grad_rows, "grad_rows"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        require_cuda(t, name, torch.float32)
        if t.shape != (row_ids.shape[0], params.shape[1]):
            raise RuntimeError(f"{name} must be [{row_ids.shape[0]}, {params.shape[1]}]")
    with _on(params.device):
        check(lib.thb_adam_rows(ptr(params), ptr(row_ids), ptr(grad_rows), ptr(exp_avg), ptr(exp_avg_sq), row_ids.shape[0], params.shape[1],
                                float(lr), float(betas[0]), float(betas[1]), float(eps), float(grad_scale), int(step), ptr(step_dev),
                                stream_ptr(params.device)), "thb_adam_rows")
    return params


def adam_rows_peer(params, row_ids, peer_grads, peer_flags, rank, exp_avg, exp_avg_sq, lr, betas, eps, step_dev, err, grad_scale=1.0):
    """All-reduce of the gradient rows over NVLink peer memory fused with the Adam step (thb_adam_rows_peer).
    peer_grads / peer_flags: lists of device pointers (ints), one per rank, as mapped into this process."""
    require_cuda(params, "params", torch.float32)
    require_cuda(row_ids, "row_ids", torch.int32)
    world = len(peer_grads)
    ga = (C.c_void_p * world)(*[int(p) for p in peer_grads])
    fa = (C.c_void_p * world)(*[int(p) for p in peer_flags])
    with _on(params.device):
        check(lib.thb_adam_rows_peer(ptr(params), ptr(row_ids), ga, fa, world, int(rank), ptr(exp_avg), ptr(exp_avg_sq), row_ids.shape[0],
                                     params.shape[1], float(lr), float(betas[0]), float(betas[1]), float(eps), float(grad_scale),
                                     ptr(step_dev), ptr(err), stream_ptr(params.device)), "thb_adam_rows_peer")
    return params


def _idx(t, name):
    if t.dtype not in (torch.int64, torch.int32):
        raise RuntimeError(f"{name} must be int64 or int32")
    return int(t.dtype == torch.int64)


def eval_forward(ctrl_pts, Jm, PHI, cumsum):
    """THB_eval.forward: out[N, cp_dim] (freshly allocated, like the reference)."""
    require_cuda(ctrl_pts, "ctrl_pts", torch.float32)
    require_cuda(Jm, "Jm_array")
    require_cuda(PHI, "tensor_prod", torch.float32)
    require_cuda(cumsum, "num_supp_bs_cumsum")
    if ctrl_pts.dim() != 2 or not 1 <= ctrl_pts.shape[1] <= 4:
        raise RuntimeError("ctrl_pts must be [nCP, 1..4]")
    if Jm.shape[0] != PHI.shape[0]:
        raise RuntimeError("Jm_array and tensor_prod must have the same length")
    n = cumsum.shape[0]
    out = torch.zeros((n, ctrl_pts.shape[1]), dtype=torch.float32, device=ctrl_pts.device)
    with _on(ctrl_pts.device):
        check(lib.thb_eval_forward(ptr(ctrl_pts), ptr(Jm), _idx(Jm, "Jm_array"), ptr(PHI), ptr(cumsum), _idx(cumsum, "cumsum"), ptr(out), n,
                                   ctrl_pts.shape[1], stream_ptr(ctrl_pts.device)), "thb_eval_forward")
    return out


def eval_backward(grad_output, ctrl_pts, Jm, PHI, cumsum, deterministic=None):
    """THB_eval.backward: grad_ctrl_pts[nCP, cp_dim].  deterministic=True (default from THB_DETERMINISTIC): instead of the
    warp-aggregated atomics, the (point, support) products are sorted by support id (stable radix sort) and summed per
    control point in that order -- bit-identical from run to run, at the price of materialising the nnz x cp_dim products
    (the reference itself is order-dependent: one atomicAdd per pair, compute_pts_cuda_kernels.cu:43)."""
    require_cuda(ctrl_pts, "ctrl_pts", torch.float32)
    grad_output = grad_output.contiguous()
    require_cuda(grad_output, "grad_output", torch.float32)
    require_cuda(Jm, "Jm_array")
    require_cuda(PHI, "tensor_prod", torch.float32)
    require_cuda(cumsum, "num_supp_bs_cumsum")
    n = cumsum.shape[0]
    if grad_output.shape != (n, ctrl_pts.shape[1]):
        raise RuntimeError("grad_output must be [N, cp_dim]")
    if deterministic is None:
        deterministic = deterministic_default()
    if deterministic:
        lengths = torch.diff(cumsum.long(), prepend=cumsum.new_zeros(1).long())
        seg = torch.repeat_interleave(torch.arange(n, device=cumsum.device), lengths)
        order = torch.argsort(Jm.long(), stable=True)
        vals = (PHI[order].unsqueeze(1) * grad_output[seg[order]]).contiguous()
        counts = torch.bincount(Jm.long(), minlength=ctrl_pts.shape[0])
        return torch.segment_reduce(vals, "sum", lengths=counts, axis=0, unsafe=True).to(torch.float32).contiguous()
    g = torch.empty_like(ctrl_pts)
    with _on(ctrl_pts.device):
        check(lib.thb_eval_backward(ptr(grad_output), ptr(Jm), _idx(Jm, "Jm_array"), ptr(PHI), ptr(cumsum), _idx(cumsum, "cumsum"), ptr(g), n,
                                    ctrl_pts.shape[0], ctrl_pts.shape[1], stream_ptr(ctrl_pts.device)), "thb_eval_backward")
    return g


def find_spans(params1d, knots, degree):
    require_cuda(params1d, "params", torch.float32)
    require_cuda(knots, "knots", torch.float32)
    n = params1d.shape[0]
    out = torch.empty(n, dtype=torch.int32, device=params1d.device)
    with _on(params1d.device):
        check(lib.thb_find_spans(ptr(params1d), 1, n, ptr(knots), knots.shape[0], int(degree), ptr(out), stream_ptr(params1d.device)), "thb_find_spans")
    return out


def basis_1d(params1d, knots, degree, derivs=False):
    require_cuda(params1d, "params", torch.float32)
    require_cuda(knots, "knots", torch.float32)
    n = params1d.shape[0]
    v = torch.empty((n, degree + 1), dtype=torch.float32, device=params1d.device)
    dv = torch.empty_like(v) if derivs else None
    with _on(params1d.device):
        check(lib.thb_basis_1d(ptr(params1d), 1, n, ptr(knots), knots.shape[0], int(degree), ptr(v), ptr(dv), stream_ptr(params1d.device)), "thb_basis_1d")
    return (v, dv) if derivs else v


def tensor_product(vectors):
    """Outer product of 2 or 3 vectors, or of batches of them ([B, n_k] each): thb_tensor_product."""
    v = [require_cuda(x.contiguous(), "factor", torch.float32) for x in vectors]
    if len(v) not in (2, 3) or any(x.dim() != v[0].dim() or x.dim() not in (1, 2) for x in v):
        raise RuntimeError("compute_tensor_product takes 2 or 3 vectors (or [B, n] batches of vectors)")
    batched = v[0].dim() == 2
    B = v[0].shape[0] if batched else 1
    if batched and any(x.shape[0] != B for x in v):
        raise RuntimeError("batch sizes differ")
    n = [x.shape[-1] for x in v]
    out = torch.empty(([B] if batched else []) + n, dtype=torch.float32, device=v[0].device)
    with _on(v[0].device):
        check(lib.thb_tensor_product(ptr(v[0]), ptr(v[1]), ptr(v[2]) if len(v) == 3 else None, B, n[0], n[1], n[2] if len(v) == 3 else 1,
                                     ptr(out), stream_ptr(v[0].device)), "thb_tensor_product")
    return out


class GridPlan:
    """Work items of a TENSOR-PRODUCT GRID of parameters (axes[k]: the non-decreasing values of parameter k) for
    thb_eval_grid: per active cell the index box of the grid points it holds, cut into boxes of at most 16 values per axis.
    Built on the host with numpy from the hierarchy's knots, once per grid -- there is no per-point sort.

    order: row of point (i0, i1[, i2]) in the output.  "xy" (default) is the reference's generate_parametric_coordinates
    order (numpy meshgrid 'xy': i1 slowest, then i0, i2 fastest; 2-D: i1 slowest); "ij" is C order."""

    BOX = 16

    def __init__(self, P, axes, order="xy"):
        import numpy as np

        d = P.ndim
        if len(axes) != d:
            raise RuntimeError(f"GridPlan needs {d} axes")
        ax = [np.ascontiguousarray((a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)), dtype=np.float32).reshape(-1) for a in axes]
        if any(len(a) == 0 or np.any(np.diff(a) < 0) for a in ax):
            raise RuntimeError("GridPlan: every axis must be a non-empty, non-decreasing sequence")
        n = [len(a) for a in ax]
        info = P.cellinfo_host
        knots = P.knots.detach().cpu().numpy()
        lo = np.zeros((P.nslots, 3), dtype=np.int64)
        cnt = np.ones((P.nslots, 3), dtype=np.int64)
        inside = 1
        for k in range(d):
            p = P.degrees[k]
            for l in range(P.nlevels):
                U = knots[P.knot_off[l, k]: P.knot_off[l, k] + P.knot_len[l, k]]
                nc = int(P.sh_cells[l][k])
                B = U[p: p + nc + 1]
                first = np.searchsorted(ax[k], B, side="left").astype(np.int64)    # first grid value >= breakpoint
                first[-1] = np.searchsorted(ax[k], B[-1], side="right")             # u == U[-1] belongs to the last cell
                sel = info[:, 0] == l
                c = info[sel, 1 + k]
                lo[sel, k] = first[c]
                cnt[sel, k] = first[c + 1] - first[c]
                if l == 0:
                    inside *= int(first[-1] - first[0])
        npts = cnt.prod(axis=1)
        if int(npts.sum()) != inside:
            raise RuntimeError("GridPlan: the active cells do not partition the domain (overlapping active cells of different "
                               "levels); use the general path (engine.Plan)")
        slots = np.nonzero(npts > 0)[0]
        m = (cnt[slots] + self.BOX - 1) // self.BOX                                # boxes per axis
        per = m.prod(axis=1)
        rep = np.repeat(np.arange(len(slots)), per)
        t = np.arange(int(per.sum())) - np.repeat(np.cumsum(per) - per, per)
        mm = m[rep]
        b2 = t % mm[:, 2]
        b1 = (t // mm[:, 2]) % mm[:, 1]
        b0 = t // (mm[:, 2] * mm[:, 1])
        bidx = np.stack([b0, b1, b2], axis=1)
        start = lo[slots][rep] + bidx * self.BOX
        size = np.minimum(self.BOX, cnt[slots][rep] - bidx * self.BOX)
        items = np.zeros((len(rep), 8), dtype=np.int32)
        items[:, 0] = slots[rep]
        items[:, 1], items[:, 2], items[:, 3], items[:, 4], items[:, 5], items[:, 6] = start[:, 0], size[:, 0], start[:, 1], size[:, 1], start[:, 2], size[:, 2]
        items[:, 7] = info[slots[rep], 0]
        cells = np.zeros((len(rep), 4), dtype=np.int32)
        cells[:, :d] = info[slots[rep], 1:1 + d]
        big_first = np.argsort(-(size.prod(axis=1)), kind="stable")               # dynamic scheduler: large boxes first
        self.P, self.n, self.order = P, n, order
        self.n_points, self.n_inside, self.nitems = int(np.prod(n)), inside, len(rep)
        self.items = torch.from_numpy(items[big_first]).to(P.device)
        self.cells = torch.from_numpy(cells[big_first]).to(P.device)
        self.axes = [torch.from_numpy(a).to(P.device) for a in ax]
        if order == "xy":
            st = [n[2], n[0] * n[2], 1] if d == 3 else [1, n[0], 0]
        elif order == "ij":
            st = [n[1] * n[2], n[2], 1] if d == 3 else [n[1], 1, 0]
        else:
            raise RuntimeError("order must be 'xy' or 'ij'")
        self.strides = (C.c_int64 * 3)(*st)
        self.cursor = torch.zeros(1, dtype=torch.int32, device=P.device)
        self.version = P.version


def eval_grid(P, gplan, ctrl_pts=None, nets=None, out=None):
    """Geometry on the tensor-product grid of `gplan` (thb_eval_grid): [prod(n), cp_dim] in the plan's row order.  Pass
    ctrl_pts (the monomial nets are built here, for every cell) or prebuilt monomial nets."""
    if gplan.version != P.version:
        raise RuntimeError("the hierarchy changed since the GridPlan was built")
    if nets is None:
        nets = cell_nets(P, ctrl_pts, monomial=True)
    if not getattr(nets, "_thb_flags", NETS_MONOMIAL) & NETS_MONOMIAL:
        raise RuntimeError("eval_grid needs monomial nets (cell_nets(..., monomial=True))")
    cpd = ctrl_pts.shape[1] if ctrl_pts is not None else 3
    if out is None:
        alloc = torch.empty if gplan.n_inside == gplan.n_points else torch.zeros
        out = alloc((gplan.n_points, cpd), dtype=torch.float32, device=P.device)
    require_cuda(out, "out", torch.float32)
    if out.shape != (gplan.n_points, cpd):
        raise RuntimeError(f"out must be [{gplan.n_points}, {cpd}]")
    with _on(P.device):
        check(lib.thb_eval_grid(C.byref(P.desc()), ptr(nets), ptr(gplan.items), ptr(gplan.cells), gplan.nitems, ptr(gplan.axes[0]),
                                ptr(gplan.axes[1]), ptr(gplan.axes[2]) if P.ndim == 3 else None, gplan.strides, cpd, ptr(out),
                                ptr(gplan.cursor), stream_ptr(P.device)), "thb_eval_grid")
    return out


def fma_peak(variant, iters=4096, repeats=5):
    """Measured FP32 FMA throughput (TFLOP/s) of the running GPU -- the FP32 roofline denominator."""
    sink = torch.zeros(4, dtype=torch.float32, device="cuda")
    flops = C.c_double(0.0)
    st = stream_ptr()
    check(lib.thb_fma_peak(variant, iters, ptr(sink), C.byref(flops), st), "thb_fma_peak")
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.thb_fma_peak(variant, iters, ptr(sink), C.byref(flops), st), "thb_fma_peak")
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


class HostPipeline:
    """Evaluation from HOST buffers: params (pinned, [N, d] f32) -> geometry (pinned, [N, cp_dim] f32).

    The batch is cut into contiguous chunks; three CUDA streams overlap the H2D copy of chunk c+1, the
    plan + fused kernel of chunk c and the D2H copy of chunk c-1 (PCIe is full duplex), so a step costs about
    max(H2D, D2H, compute) instead of their sum.  Device buffers are allocated once.

    use_graph: the whole step (copies, ~10 launches per chunk, the cross-stream dependencies) is captured into ONE CUDA
    graph the first time it runs with a given set of buffers and replayed afterwards, so the host issues a single
    launch per step instead of a few hundred calls; a step with other buffers is captured again (eager fallback if
    capture is not possible)."""

    def __init__(self, P, n_points, cp_dim=3, n_chunks=8, cellnet=True, formulation=None, use_graph=True, ramp=0):
        self.P, self.n, self.cpd = P, int(n_points), cp_dim
        self.use_graph, self._graph, self._graph_key = bool(use_graph), None, None
        self.formulation = formulation if formulation is not None else ("local_net" if cellnet else "per_point")
        n_chunks = max(1, min(n_chunks, (self.n + 65535) // 65536))
        # chunk sizes ramp up (1, 2, 4, ...) at the start and down at the end: the pipeline's fill time is the H2D copy of
        # the FIRST chunk and its drain time the D2H copy of the LAST one, so those two are kept small
        r = min(ramp, max(0, (n_chunks - 1) // 2))
        w = [1 << i for i in range(r)] + [1 << r] * (n_chunks - 2 * r) + [1 << i for i in reversed(range(r))]
        edges = [0]
        acc = 0
        for x in w:
            acc += x
            edges.append((self.n * acc) // sum(w))
        self.bounds = [(edges[i], edges[i + 1]) for i in range(n_chunks) if edges[i + 1] > edges[i]]
        m = max(hi - lo for lo, hi in self.bounds)
        dev = P.device
        self.pbuf = [torch.empty((m, P.ndim), dtype=torch.float32, device=dev) for _ in range(2)]
        self.obuf = [torch.empty((m, cp_dim), dtype=torch.float32, device=dev) for _ in range(2)]
        self.plans = [None, None]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))

    def run(self, params_host, ctrl_pts, out_host):
        """Enqueues one full evaluation on the current stream; returns after everything is enqueued (synchronise the
        stream / device before reading out_host)."""
        if not self.use_graph:
            return self._enqueue(params_host, ctrl_pts, out_host)
        key = (params_host.data_ptr(), ctrl_pts.data_ptr(), out_host.data_ptr(), self.P.version, self.formulation)
        if self._graph_key != key:
            self._enqueue(params_host, ctrl_pts, out_host)   # eager once: allocates plans / workspaces, sizes the grids
            torch.cuda.synchronize(self.P.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(params_host, ctrl_pts, out_host)
                self._graph, self._graph_key = g, key
                self._keep = (params_host, ctrl_pts, out_host)  # the graph holds raw pointers into these
            except Exception:
                self.use_graph, self._graph, self._graph_key = False, None, None
                torch.cuda.synchronize(self.P.device)
                return self._enqueue(params_host, ctrl_pts, out_host)
        self._graph.replay()
        return None

    def _enqueue(self, params_host, ctrl_pts, out_host):
        cur = torch.cuda.current_stream(self.P.device)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        done, outd = {}, {}
        nets = None
        if self.formulation == "local_net":  # the nets depend on ctrl_pts only: built once, shared by all chunks
            with torch.cuda.stream(self.s_cmp):
                nets = cell_nets(self.P, ctrl_pts)
        for c, (lo, hi) in enumerate(self.bounds):
            b, m = c % 2, hi - lo
            with torch.cuda.stream(self.s_in):
                if c >= 2:
                    self.s_in.wait_event(done[c - 2])  # the kernel that read this params buffer has finished
                self.pbuf[b][:m].copy_(params_host[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in)
                if c >= 2:
                    self.s_cmp.wait_event(outd[c - 2])  # the D2H copy that read this output buffer has finished
                pl = self.plans[b]
                if pl is None:
                    pl = self.plans[b] = Plan(self.P, self.pbuf[b][:m], capacity=self.pbuf[b].shape[0])
                else:
                    pl.rebind(self.pbuf[b][:m])
                self.obuf[b][:m].zero_()  # rows of points outside the domain are not written by the kernels
                if nets is not None:
                    eval_nets(self.P, pl, nets, self.cpd, out=[self.obuf[b][:m]], with_derivatives=False)
                else:
                    eval_fused(self.P, pl, ctrl_pts, formulation=self.formulation, out=[self.obuf[b][:m]])
                done[c] = torch.cuda.Event()
                done[c].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done[c])
                out_host[lo:hi].copy_(self.obuf[b][:m], non_blocking=True)
                outd[c] = torch.cuda.Event()
                outd[c].record(self.s_out)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_in)
        return outd[len(self.bounds) - 1]
