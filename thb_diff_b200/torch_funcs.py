"""Torch adapter / autograd boundary with the reference's names [THB/torch_funcs.py]:
THBEval, THB_nn_module, prepare_data_for_CUDA_evaluation, prepare_data_for_evaluation, Evaluate --
plus the fused path (THBFusedEval / THB_fused_module) that never materialises PHI."""
import numpy as np
import torch

from . import THB_eval, engine
from .core import PointSupports
from .packed import PackedHierarchy


class THBEval(torch.autograd.Function):
    """[THB/torch_funcs.py:22-47] out = A @ ctrl_pts with A = CSR(PHI, Jm, cumsum); gradient only w.r.t.
    ctrl_pts.  device == "cuda" uses device tensors as they are; any other value stages CPU
    tensors through the GPU (the reference would call its CPU build)."""

    @staticmethod
    def forward(ctx, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum, device):
        ctx.save_for_backward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)
        ctx.device = device
        if _is_cuda(device):
            return THB_eval.forward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)
        return THB_eval.cpp_forward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)

    @staticmethod
    def backward(ctx, grad_output):
        ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum = ctx.saved_tensors
        device = ctx.device
        g = None
        if ctx.needs_input_grad[0]:
            if _is_cuda(device):
                g = THB_eval.backward(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)
            else:
                g = THB_eval.cpp_backward(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)
        g_phi = None
        if ctx.needs_input_grad[2]:
            # the reference's THBEval has no such gradient; its pure-torch Evaluate does (autograd through
            # ctrl_pts[Jm] * PHI): d out / d PHI_j = <grad_out[point of j], ctrl_pts[Jm_j]>
            n = num_supp_bs_cumsum.shape[0]
            lengths = torch.diff(num_supp_bs_cumsum, prepend=num_supp_bs_cumsum.new_zeros(1))
            seg = torch.repeat_interleave(torch.arange(n, device=lengths.device), lengths)
            g_phi = (grad_output[seg.to(grad_output.device)] * ctrl_pts[Jm_array.long()]).sum(1).to(tensor_prod.dtype)
        return g, None, g_phi, None, None


def _is_cuda(device):
    return device == "cuda" or (isinstance(device, str) and device.startswith("cuda")) or (isinstance(device, torch.device) and device.type == "cuda")


class THB_nn_module(torch.nn.Module):
    """[THB/torch_funcs.py:8-19]"""

    def __init__(self, Jm, PHI, num_supp_cumsum, device):
        super().__init__()
        self.Jm, self.PHI, self.num_supp_cumsum, self.device = Jm, PHI, num_supp_cumsum, device

    def forward(self, ctrl_pts):
        return THBEval.apply(ctrl_pts, self.Jm, self.PHI, self.num_supp_cumsum, self.device)


def _stack_ctrl_pts(ctrl_pts, device):
    if isinstance(ctrl_pts, dict):
        L = max(ctrl_pts.keys())
        cp_dim = np.asarray(ctrl_pts[0]).shape[-1]
        return torch.cat([torch.as_tensor(np.asarray(ctrl_pts[lev])).reshape(-1, cp_dim).float() for lev in range(L + 1)]).to(device)
    return torch.as_tensor(ctrl_pts).float().to(device)


def prepare_data_for_CUDA_evaluation(PHI, ac_spans, num_supp, ctrl_pts, ac_cells_ac_supp, fn_sh, device):
    """[THB/torch_funcs.py:50-80] -> (ctrl_pts [nCP, cp_dim] f32, Jm int64 [nnz], PHI f32 [nnz],
    inclusive cumsum int64 [N], device).  `ac_spans` may be the PointSupports object returned by
    compute_active_span (fast: Jm is expanded on the GPU) or the reference's list of (lev, cellIdx)."""
    cp = _stack_ctrl_pts(ctrl_pts, device)
    phi = PHI.float().to(device) if isinstance(PHI, torch.Tensor) else torch.from_numpy(np.asarray(PHI)).float().to(device)
    if isinstance(ac_spans, PointSupports):
        Jm = ac_spans.Jm(torch.int64).to(device)
        ac_spans.offsets()
        cumsum = ac_spans._cumsum.to(device)
    else:
        L = max(fn_sh.keys())
        nCP = np.zeros(L + 2, dtype=np.int64)
        for lev in range(L + 1):
            nCP[lev + 1] = nCP[lev] + int(np.prod(fn_sh[lev]))
        Jm = torch.tensor([int(nCP[fl] + np.ravel_multi_index(tuple(fn), fn_sh[fl])) for lev, cell in ac_spans
                           for fl, fn in ac_cells_ac_supp[lev][tuple(cell)]], dtype=torch.int64).to(device)
        cumsum = torch.from_numpy(np.asarray(num_supp, dtype=np.int64)).cumsum(0).to(device)
    return cp, Jm, phi, cumsum, device


def prepare_data_for_evaluation(PHI, ac_spans, num_supp, ctrl_pts, ac_cells_ac_supp, fn_sh, device):
    """[THB/torch_funcs.py:83-119] same packing with segment ids instead of the cumsum."""
    cp, Jm, phi, cumsum, device = prepare_data_for_CUDA_evaluation(PHI, ac_spans, num_supp, ctrl_pts, ac_cells_ac_supp, fn_sh, device)
    n = cumsum.shape[0]
    lengths = torch.diff(cumsum, prepend=cumsum.new_zeros(1))
    seg = torch.repeat_interleave(torch.arange(n, device=cumsum.device), lengths).unsqueeze(1).expand(-1, cp.shape[1])
    return cp, Jm, phi.unsqueeze(1), seg, n


def Evaluate(ctrl_pts, Jm, PHI, segment_ids, num_pts, device):
    """[THB/torch_funcs.py:122-128] out = zeros(num_pts, cp_dim).scatter_add_(0, segment_ids, ctrl_pts[Jm] * PHI): the same
    contraction, here through the CSR kernel.  segment_ids are the non-decreasing point ids prepare_data_for_evaluation
    produces.  Differentiable with respect to ctrl_pts and PHI like the reference's torch expression; tensors that live on
    the CPU (device not "cuda") are staged through the GPU -- there is no CPU arithmetic."""
    seg = segment_ids[:, 0].contiguous()
    counts = torch.bincount(seg, minlength=num_pts)
    return THBEval.apply(ctrl_pts, Jm, PHI.reshape(-1).contiguous(), torch.cumsum(counts, 0), "cuda" if _is_cuda(device) else "cpu")


# ---- fused path -----------------------------------------------------------------------------------------


class THBFusedEval(torch.autograd.Function):
    """basis + evaluation in one kernel (PHI stays on the SM); backward = fused scatter into grad ctrl_pts."""

    @staticmethod
    def forward(ctx, ctrl_pts, module):
        ctx.module = module
        return engine.eval_fused(module.P, module.plan, ctrl_pts.contiguous(), (engine.CH_VALUE,), cellnet=module.cellnet)[0]

    @staticmethod
    def backward(ctx, grad_output):
        m = ctx.module
        return engine.eval_fused_backward(m.P, m.plan, grad_output.contiguous(),
                                          formulation="local_net" if m.cellnet else "per_point"), None


class THB_fused_module(torch.nn.Module):
    """Evaluates the THB geometry at fixed parametric points for changing control points.

        m = THB_fused_module(space_or_packed, params); pts = m(ctrl_pts); pts.sum().backward()
    """

    def __init__(self, space, params, device="cuda", cellnet=True, plan_order_rows=False):
        """plan_order_rows=True: the rows of the output (and of the incoming gradient) follow the plan's cell order, row j
        = point self.perm[j] of `params` -- coalesced rows; permute per-point data (targets) with self.perm once."""
        super().__init__()
        self.P = space if isinstance(space, PackedHierarchy) else PackedHierarchy.from_space(space, device=device)
        self.plan = engine.Plan(self.P, params)
        self.cellnet = cellnet
        self.perm = None
        if plan_order_rows:
            self.use_plan_order_rows()

    def points_inside(self):
        """number of the plan's points that lie inside the domain (synchronises once)"""
        return int(self.plan.counters[2])

    def use_plan_order_rows(self):
        """Switches the rows of the output / incoming gradient to the plan's cell order (see __init__)."""
        if not self.cellnet:
            raise RuntimeError("plan_order_rows needs the local-net formulation (cellnet=True)")
        perm = self.plan.caller_index()
        if perm.numel() != self.plan.n:
            raise RuntimeError("plan_order_rows needs every point inside the domain")
        self.perm = perm
        self.plan.set_rows_in_plan_order(True)
        return self

    def forward(self, ctrl_pts):
        return THBFusedEval.apply(ctrl_pts, self)

    def derivatives(self, ctrl_pts):
        """d(geometry)/d(u_k) for every parametric direction: [N, ndim, cp_dim]"""
        outs = engine.eval_fused(self.P, self.plan, ctrl_pts.contiguous(), engine.grad_channels(self.P.ndim))
        return torch.stack(outs, dim=1)
