"""Multi-GPU use of the hot path: one process per GPU (torchrun), parametric points sharded across ranks.

Evaluation (spans, basis, geometry, derivatives) needs NO communication: every rank evaluates its own
points against replicated tables.  Differentiable fitting reduces exactly one tensor per step -- the
control-point gradient [nCP, cp_dim] -- with an NCCL all-reduce (SUM); parameters stay replicated, so the
optimiser step is identical on every rank.  The reference has no distributed code at all."""
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


def allreduce_grad_(grad, group=None):
    """SUM all-reduce of the control-point gradient in place (NCCL on GPUs, any backend for tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


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
    """MSE fitting of the control points to targets at this rank's points.

        fit = ShardedFit(packed, params_of_this_rank, targets_of_this_rank, n_total)
        loss = fit.step(ctrl_pts, optimizer)      # forward, backward, grad all-reduce, optimiser step
    """

    def __init__(self, packed, params, targets, n_total, cellnet=True, group=None):
        from .torch_funcs import THB_fused_module

        # the points are fixed for the whole fit: keep outputs, targets and gradients in the plan's cell order (coalesced
        # rows) when the local-net kernels are used and every point lies inside the domain
        try:
            self.module = THB_fused_module(packed, params, cellnet=cellnet, plan_order_rows=bool(cellnet))
        except RuntimeError:
            self.module = THB_fused_module(packed, params, cellnet=cellnet)
        self.targets = targets if self.module.perm is None else targets[self.module.perm].contiguous()
        self.n_total = int(n_total)
        self.group = group

    def step(self, ctrl_pts, optimizer):
        optimizer.zero_grad(set_to_none=True)
        out = self.module(ctrl_pts)
        # sum of squared errors over this shard, normalised by the GLOBAL point count, so that the
        # all-reduced gradient is the gradient of the global mean squared error
        loss = _ScaledSumSquaredError.apply(out, self.targets, 1.0 / (self.n_total * out.shape[1]))
        loss.backward()
        allreduce_grad_(ctrl_pts.grad, self.group)
        optimizer.step()
        return loss.detach()
