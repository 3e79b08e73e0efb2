"""Multi-GPU use of the hot path: one process per GPU (torchrun), parametric points sharded across ranks.

Evaluation (spans, basis, geometry, derivatives) needs NO communication: every rank evaluates its own
points against replicated tables.  Differentiable fitting reduces exactly one tensor per step -- the
control-point gradient -- with an NCCL all-reduce (SUM); parameters stay replicated, so the optimiser step is
identical on every rank.  Only ACTIVE functions ever receive a gradient, so the tensor that is reduced is the compact
[n_active, cp_dim] form (0.5 MB instead of 31 MB on BASELINE config 3), not the dense [nCP, cp_dim] one.
The reference has no distributed code at all; its fit loop is THB/torch_funcs.py:8-47 on one device.

Sharding a FIT (fixed points, many steps) by cells: the per-cell work of a step (local control nets, the transposed
fold of the net gradients) is as large as the per-point work, so ranks must own disjoint sets of CELLS, not arbitrary
subsets of the points -- `partition_by_cell` moves every point to the rank that owns its cell once, at set-up
(all-to-all), balanced by points + per-cell cost ("sort by active cell then split evenly")."""
import os as _os

import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n points: rank r owns [lo, hi)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_points(params, rank, world, mode="contiguous"):
    """Rows of `params` owned by `rank`: contiguous slabs (grid order keeps cells together) or strided
    (every rank sees the same mix of cells -- better balance when refinement is localised)."""
    if mode == "strided":
        return params[rank::world]
    lo, hi = shard_bounds(params.shape[0], rank, world)
    return params[lo:hi]


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_grad_(grad, group=None):
    """SUM all-reduce of the control-point gradient in place (NCCL on GPUs, any backend for tests)."""
    if _world(group) > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


def allreduce_active_rows_(grad, active_ids, group=None):
    """The same reduction for a DENSE [nCP, cp_dim] gradient (a generic torch optimiser wants one) moving only the rows
    that can be non-zero: gather the active rows, all-reduce them, put them back."""
    if _world(group) > 1:
        idx = active_ids.long()
        rows = grad.index_select(0, idx)
        dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)
        grad.index_copy_(0, idx, rows)
    return grad


def balanced_bounds(cost, world):
    """Splits a sequence of non-negative costs into `world` contiguous ranges of (nearly) equal total cost.
    Returns int64 [world + 1] bounds (non-decreasing, first 0, last len(cost))."""
    cum = torch.cumsum(cost.to(torch.float64), 0)
    n = cum.numel()
    total = float(cum[-1]) if n else 0.0
    targets = torch.arange(1, world, dtype=torch.float64, device=cum.device) * (total / world)
    inner = torch.searchsorted(cum, targets, right=False) + 1   # first entry NOT owned by rank r-1
    b = torch.cat([torch.zeros(1, dtype=torch.int64, device=cum.device), inner.to(torch.int64),
                   torch.tensor([n], dtype=torch.int64, device=cum.device)])
    return torch.cummax(b.clamp(max=n), 0).values


def cell_owner_bounds(cell_points, world, cell_cost=100.0):
    """Splits the active cells (slot order = level-major, C order: spatially coherent) into `world` contiguous ranges of
    equal cost; cost of a cell = its points + cell_cost if it holds any (the nets / unfold work of a cell is worth about
    a hundred point evaluations).  cell_points: [nslots] global point counts.  Returns int64 [world + 1] slot bounds."""
    c = cell_points.to(torch.float64)
    return balanced_bounds(c + (c > 0).to(torch.float64) * float(cell_cost), world)


def grid_slab_bounds(P, axes, world, cell_cost=100.0, axis=0):
    """Strong scaling of a FIXED tensor-product grid of points (BASELINE config 3 / 5: 'param grid sharded over 1/2/4/8
    GPUs'): contiguous slabs of grid planes along `axis`, balanced by points + per-cell cost, so that every rank builds
    local control nets only for the cells its slab touches.  axes: per-dimension 1-D parameter values (device tensors).
    The cost of a plane = its points + cell_cost * (cells it touches, each cell shared among the planes that cross it).
    Returns int64 [world + 1] plane bounds along `axis`."""
    from . import engine

    n_ax = [int(a.shape[0]) for a in axes]
    if world == 1:
        return torch.tensor([0, n_ax[axis]], dtype=torch.int64, device=P.device)
    per_plane = 1
    for k, n in enumerate(n_ax):
        per_plane *= n if k != axis else 1
    chunk = max(1, (1 << 23) // per_plane)   # planes per pass (about 8M points at a time)

    def slots(lo, hi):
        sub = [a[lo:hi] if k == axis else a for k, a in enumerate(axes)]
        pts = torch.stack(torch.meshgrid(*sub, indexing="ij"), dim=-1)
        pts = pts.movedim(axis, 0).reshape(-1, len(axes)).contiguous()      # plane-major
        return engine.active_span(P, pts, want_num_supp=False)[0].long().view(hi - lo, per_plane)

    cnt = torch.zeros(P.nslots, dtype=torch.int64, device=P.device)
    for lo in range(0, n_ax[axis], chunk):                                   # pass 1: points per cell over the whole grid
        s = slots(lo, min(lo + chunk, n_ax[axis]))
        cnt += torch.bincount(s[s >= 0], minlength=P.nslots)
    plane = torch.empty(n_ax[axis], dtype=torch.float64, device=P.device)
    share = float(cell_cost) / cnt.clamp(min=1).to(torch.float64)
    for lo in range(0, n_ax[axis], chunk):                                   # pass 2: cost of every plane
        s = slots(lo, min(lo + chunk, n_ax[axis]))
        w = torch.where(s >= 0, 1.0 + share[s.clamp(min=0)], torch.ones((), dtype=torch.float64, device=P.device))
        plane[lo: lo + s.shape[0]] = w.sum(1)
    return balanced_bounds(plane, world)


def estimate_cell_cost(points, cells, times):
    """Cost of a cell in point-equivalents from measured per-rank step times: least squares of t_r = a * points_r + b * cells_r
    over the ranks, returns b / a (None when the fit is degenerate).  points / cells / times: sequences, one entry per rank.
    A fit measures its first steps with a guessed cost, calls this and re-partitions once (cells of a finer level cost more
    than the guess, coarse cells with thousands of points less)."""
    import numpy as np

    A = np.stack([np.asarray(points, dtype=np.float64), np.asarray(cells, dtype=np.float64)], axis=1)
    t = np.asarray(times, dtype=np.float64)
    if len(t) < 2 or np.linalg.matrix_rank(A) < 2:
        return None
    (a, b), *_ = np.linalg.lstsq(A, t, rcond=None)
    if not (a > 0 and b > 0):
        return None
    return float(b / a)


def partition_by_cell(slot, nslots, payload, group=None, cell_cost=100.0):
    """Moves every point to the rank that owns its active cell.

    slot: int32/int64 [n] active-cell slot of this rank's points (engine.active_span; -1 = outside the domain: such a
    point goes to rank 0); payload: float tensor [n, w] travelling with the points (parameters, targets, ...).
    Returns (payload rows now owned by this rank [m, w], slot bounds [world + 1]).  Collectives: one all-reduce of the
    cell histogram, one all-to-all of the counts, one all-to-all of the rows -- once per fit, not per step."""
    world = _world(group)
    slot = slot.long()
    inside = slot >= 0
    hist = torch.bincount(slot[inside], minlength=nslots)
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    bounds = cell_owner_bounds(hist, world, cell_cost)
    if world == 1:
        return payload, bounds
    dest = torch.searchsorted(bounds[1:-1].contiguous(), slot.clamp(min=0), right=True)
    dest = torch.where(inside, dest, torch.zeros_like(dest))
    order = torch.argsort(dest, stable=True)
    send = payload.index_select(0, order).contiguous()
    n_send = torch.bincount(dest, minlength=world)
    n_recv = torch.empty_like(n_send)
    dist.all_to_all_single(n_recv, n_send, group=group)
    recv = payload.new_empty((int(n_recv.sum()), payload.shape[1]))
    dist.all_to_all_single(recv, send, output_split_sizes=n_recv.tolist(), input_split_sizes=n_send.tolist(), group=group)
    return recv, bounds


class _ScaledSumSquaredError(torch.autograd.Function):
    """scale * sum((out - target)^2) in three passes over the data (difference, dot product, scaled difference as
    the gradient) instead of the ~10 of the composed torch expression; the fit loop's loss is bandwidth-bound."""

    @staticmethod
    def forward(ctx, out, target, scale):
        diff = out - target
        ctx.save_for_backward(diff)
        ctx.scale = scale
        return torch.dot(diff.reshape(-1), diff.reshape(-1)) * scale

    @staticmethod
    def backward(ctx, g):
        (diff,) = ctx.saved_tensors
        return diff * (g * (2.0 * ctx.scale)), None, None


class ShardedFit:
    """MSE fitting of the control points to targets at this rank's points, through autograd and ANY torch optimiser.

        fit = ShardedFit(packed, params_of_this_rank, targets_of_this_rank, n_total)
        loss = fit.step(ctrl_pts, optimizer)      # forward, backward, grad all-reduce, optimiser step

    The gradient handed to the optimiser is the dense [nCP, cp_dim] tensor; only its active rows cross the wire.
    FitEngine below is the same loop without autograd, with the gradient kept compact end to end."""

    def __init__(self, packed, params, targets, n_total, cellnet=True, group=None):
        from .torch_funcs import THB_fused_module

        # the points are fixed for the whole fit: keep outputs, targets and gradients in the plan's cell order (coalesced
        # rows) when the local-net kernels are used and every point lies inside the domain
        self.module = THB_fused_module(packed, params, cellnet=cellnet)
        if cellnet and self.module.points_inside() == self.module.plan.n:
            self.module.use_plan_order_rows()
        self.targets = targets if self.module.perm is None else targets[self.module.perm].contiguous()
        self.n_total = int(n_total)
        self.group = group
        self.active_ids = packed.active_rows()[0]

    def step(self, ctrl_pts, optimizer):
        optimizer.zero_grad(set_to_none=True)
        out = self.module(ctrl_pts)
        # sum of squared errors over this shard, normalised by the GLOBAL point count, so that the
        # all-reduced gradient is the gradient of the global mean squared error
        loss = _ScaledSumSquaredError.apply(out, self.targets, 1.0 / (self.n_total * out.shape[1]))
        loss.backward()
        allreduce_active_rows_(ctrl_pts.grad, self.active_ids, self.group)
        optimizer.step()
        return loss.detach()


class PeerRows:
    """The compact gradient rows of every rank in memory all ranks of the node can read (torch symmetric memory: the same
    allocation on every GPU, mapped into every process over NVLink), twice -- steps alternate between the two buffers --
    plus one flag array per rank.  With it the gradient all-reduce is part of the optimiser kernel (thb_adam_rows_peer)
    instead of a collective-library call."""

    def __init__(self, n_rows, cp_dim, device, group=None):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("PeerRows: at most 16 ranks")
        self.grads, self.grad_ptrs, self._handles = [], [], []
        for _ in range(2):
            t = symm.empty((n_rows, cp_dim), dtype=torch.float32, device=device)
            t.zero_()
            h = symm.rendezvous(t, group=group)
            self.grads.append(t)
            self.grad_ptrs.append([int(p) for p in h.buffer_ptrs])
            self._handles.append(h)
        self.flags = symm.empty(64, dtype=torch.int32, device=device)
        self.flags.zero_()
        hf = symm.rendezvous(self.flags, group=group)
        self._handles.append(hf)
        self.flag_ptrs = [int(p) for p in hf.buffer_ptrs]
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)   # every rank's buffers and flags are zero before anyone signals


class FitEngine:
    """The fit step of BASELINE config 4 with everything compact and nothing through autograd:

        nets(ctrl_pts) -> evaluation -> one-pass MSE loss + output gradient -> backward into COMPACT active rows
        -> all-reduce of [n_active, cp_dim] (NCCL) -> Adam on those rows of the dense parameter

        eng = FitEngine(packed, params_of_this_rank, targets_of_this_rank, n_total, lr=1e-2)
        loss = eng.step(ctrl_pts)        # ctrl_pts: dense [nCP, 3|4] float32 CUDA tensor, updated in place

    Adam never moves a row whose gradient has always been zero, so updating the active rows only is the dense
    torch.optim.Adam step (tests compare the two).  `loss` is this rank's share of the global mean squared error
    (device tensor, no synchronisation); sum it over ranks for the global value.
    use_graph: the whole step, all-reduce included, is captured into one CUDA graph after the first eager step.
    deterministic (default: THB_DETERMINISTIC): canonical point order inside the plan and the ordered backward -- the loss
    (thb_mse_grad is ordered already) and the control points are then bit-identical from run to run on one rank."""

    def __init__(self, packed, params, targets, n_total, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, group=None, use_graph=False,
                 deterministic=None, peer=None):
        from . import engine

        self.E, self.P, self.group = engine, packed, group
        self.plan = engine.Plan(packed, params)
        n_in = int(self.plan.counters[2])
        if n_in != self.plan.n:
            raise RuntimeError("FitEngine needs every point inside the domain")
        self.deterministic = engine.deterministic_default() if deterministic is None else bool(deterministic)
        if self.deterministic:
            self.plan.canonicalize()   # the points of every cell in caller order: sums over them no longer depend on the run
        perm = self.plan.caller_index()
        self.plan.set_rows_in_plan_order(True)
        dev = packed.device
        cpd = targets.shape[1]
        if cpd not in (3, 4):
            raise RuntimeError("targets must be [N, 3] or [N, 4]")
        self.cpd, self.n, self.n_total = cpd, self.plan.n, int(n_total)
        self.targets = targets.to(device=dev, dtype=torch.float32)[perm].contiguous()
        self.out = torch.empty((self.n, cpd), dtype=torch.float32, device=dev)
        self.gout = torch.empty_like(self.out)
        self.ids, self.row_map = packed.active_rows()
        na = self.ids.shape[0]
        self.grad = torch.zeros((na, cpd), dtype=torch.float32, device=dev)
        self.m, self.v = torch.zeros_like(self.grad), torch.zeros_like(self.grad)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.mse_ws = engine.mse_workspace(dev)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.use_graph, self._graph, self._graph_x = bool(use_graph), {}, None
        self.timers = None
        # gradient exchange: NVLink peer memory fused into the optimiser kernel when the ranks can map each other's memory
        # (peer=None: try it, fall back to the NCCL all-reduce; peer=True: required; peer=False: NCCL)
        self.peer, self._parity, self._nsteps = None, 0, 0
        if _world(group) > 1 and peer is not False and _os.environ.get("THB_PEER_ROWS", "1") != "0":
            try:
                self.peer = PeerRows(na, cpd, dev, group)
            except Exception as e:
                if peer:
                    raise
                self.peer_error = repr(e)

    def _enqueue(self, x, parity=0):
        E, P = self.E, self.P
        t = self.timers
        grad = self.peer.grads[parity] if self.peer is not None else self.grad
        if t:
            t[0].record()
        nets = E.cell_nets(P, x, plan=self.plan)
        E.eval_nets(P, self.plan, nets, self.cpd, out=[self.out], with_derivatives=False)
        E.mse_grad(self.out, self.targets, 1.0 / (self.n_total * self.cpd), self.gout, self.loss, self.mse_ws)
        E.eval_fused_backward_rows(P, self.plan, self.gout, out=grad, deterministic=self.deterministic)
        if t:
            t[1].record()
        if self.peer is not None:
            # arrive + wait on the peers' flags, sum of the peers' rows in rank order, Adam: one enqueue, no collective call
            if t:
                t[2].record()
            E.adam_rows_peer(x, self.ids, self.peer.grad_ptrs[parity], self.peer.flag_ptrs, self.peer.rank, self.m, self.v, self.lr,
                             self.betas, self.eps, self.step_dev, self.peer.err)
        else:
            allreduce_grad_(grad, self.group)
            if t:
                t[2].record()
            E.adam_rows(x, self.ids, grad, self.m, self.v, self.lr, self.betas, self.eps, step_dev=self.step_dev)
        if t:
            t[3].record()

    def step(self, ctrl_pts):
        x = ctrl_pts
        if x.requires_grad:
            x = x.detach()
        parity = self._parity if self.peer is not None else 0
        self._parity ^= 1
        self._nsteps += 1
        if not self.use_graph:
            self._enqueue(x, parity)
            return self.loss
        if self._graph_x != x.data_ptr():
            self._graph, self._graph_x, self._eager_left = {}, x.data_ptr(), 2   # two eager steps (one per buffer), then capture
        if self._eager_left > 0:
            self._enqueue(x, parity)   # eager: allocations, NCCL warm-up
            self._eager_left -= 1
            if self._eager_left == 0:
                torch.cuda.synchronize(self.P.device)
                try:
                    for par in ((0, 1) if self.peer is not None else (0,)):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._enqueue(x, par)   # the capture itself does not run the step
                        self._graph[par] = g
                except Exception:
                    self.use_graph, self._graph = False, {}
                    torch.cuda.synchronize(self.P.device)
            return self.loss
        self._graph[parity].replay()
        return self.loss

    def time_phases(self, ctrl_pts, iters=10):
        """(compute ms, all-reduce ms, optimiser ms) per step from CUDA events around the three phases (eager launches)."""
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
        keep = self.use_graph
        self.use_graph = False
        for k in range(iters):
            self.timers = ev[k]
            self.step(ctrl_pts)
        self.timers = None
        self.use_graph = keep
        torch.cuda.synchronize(self.P.device)
        f = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in ev) / iters
        # NCCL: compute | all-reduce | Adam.  Peer memory: compute | 0 | fused (flags + peer reads + Adam)
        return f(0, 1), f(1, 2), f(2, 3)
