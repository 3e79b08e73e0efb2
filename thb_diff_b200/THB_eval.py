"""Drop-in for the reference's torch extension module ``THB_eval``
[THB/THB_extensions/compute_pts_cuda.cpp:19-41 (forward/backward), compute_pts.cpp:4-59 (cpp_forward/
cpp_backward)].  All four names are exported (the reference's CUDA and CPU builds export disjoint pairs,
SURVEY Q15).  The cpp_* names accept CPU tensors like the reference's CPU build but the arithmetic still
runs in the CUDA kernels: inputs are staged to the GPU and the result is copied back -- there is no CPU
implementation in this package."""
import torch

from . import engine


def forward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum):
    return engine.eval_forward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)


def backward(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum):
    return engine.eval_backward(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)


def _stage(*ts):
    if not torch.cuda.is_available():
        raise RuntimeError("THB_eval: no CUDA device; thb_diff_b200 has no CPU implementation")
    return [t.contiguous().to("cuda", non_blocking=True) for t in ts]


def cpp_forward(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum):
    dev = ctrl_pts.device
    return forward(*_stage(ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)).to(dev)


def cpp_backward(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum):
    dev = ctrl_pts.device
    return backward(*_stage(grad_output, ctrl_pts, Jm_array, tensor_prod, num_supp_bs_cumsum)).to(dev)
