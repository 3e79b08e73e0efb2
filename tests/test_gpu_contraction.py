"""THB_eval.forward / backward (CSR contraction kernels) vs the reference's compiled compute_pts.cpp
(golden fixtures) and vs the oracle on seeded random CSR data."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5  # |a-b| <= TOL * max(|b|, scale), scale = max|ref| of the tensor (SURVEY hard part 6)


def _close(a, b):
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 1.0, 1e-30)
    return np.abs(np.asarray(a, dtype=np.float64) - b).max(initial=0.0) <= TOL * scale


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_golden_compute_pts(tag, idx_dtype):
    from thb_diff_b200 import THB_eval

    c = dict(np.load(f"{GOLDEN}/contraction.npz"))
    cp = torch.from_numpy(c[f"{tag}_cp"]).cuda()
    Jm = torch.from_numpy(c[f"{tag}_Jm"]).to(idx_dtype).cuda()
    phi = torch.from_numpy(c[f"{tag}_phi"]).cuda()
    cs = torch.from_numpy(c[f"{tag}_cumsum"]).to(idx_dtype).cuda()
    g = torch.from_numpy(c[f"{tag}_grad_out"]).cuda()
    out = THB_eval.forward(cp, Jm, phi, cs)
    assert out.shape == (len(cs), 3) and out.dtype == torch.float32
    assert _close(out.cpu().numpy(), c[f"{tag}_fwd"])
    gcp = THB_eval.backward(g, cp, Jm, phi, cs)
    assert gcp.shape == cp.shape
    assert _close(gcp.cpu().numpy(), c[f"{tag}_bwd"])


@pytest.mark.parametrize("cp_dim", [1, 2, 3, 4])
def test_random_csr_vs_oracle(cp_dim):
    from thb_diff_b200 import THB_eval

    rng = np.random.default_rng(11 + cp_dim)
    N, nCP = 5000, 700
    # runs of identical support lists (as produced by points sharing a cell) + ragged lengths incl. 0 and > 192
    lists = [rng.integers(0, nCP, size=s) for s in (0, 1, 12, 36, 64, 150, 200, 33)]
    which = np.repeat(rng.integers(0, len(lists), size=N // 10 + 1), 10)[:N]
    ns = np.array([len(lists[w]) for w in which])
    Jm = np.concatenate([lists[w] for w in which]).astype(np.int64)
    phi = rng.random(len(Jm)).astype(np.float32)
    cs = np.cumsum(ns).astype(np.int64)
    cp = rng.standard_normal((nCP, cp_dim)).astype(np.float32)
    g = rng.standard_normal((N, cp_dim)).astype(np.float32)
    out = THB_eval.forward(torch.from_numpy(cp).cuda(), torch.from_numpy(Jm).cuda(), torch.from_numpy(phi).cuda(), torch.from_numpy(cs).cuda())
    assert _close(out.cpu().numpy(), O.eval_forward(cp, Jm, phi, cs))
    gcp = THB_eval.backward(torch.from_numpy(g).cuda(), torch.from_numpy(cp).cuda(), torch.from_numpy(Jm).cuda(), torch.from_numpy(phi).cuda(),
                            torch.from_numpy(cs).cuda())
    assert _close(gcp.cpu().numpy(), O.eval_backward(g, cp.shape, Jm, phi, cs))


def test_empty_and_errors():
    from thb_diff_b200 import THB_eval

    cp = torch.zeros((4, 3), device="cuda")
    out = THB_eval.forward(cp, torch.zeros(0, dtype=torch.int64, device="cuda"), torch.zeros(0, device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert out.shape == (0, 3)
    with pytest.raises(RuntimeError):
        THB_eval.forward(cp.double(), torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, device="cuda"), torch.ones(1, dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError):
        THB_eval.forward(cp.cpu(), torch.zeros(1, dtype=torch.int64), torch.zeros(1), torch.ones(1, dtype=torch.int64))


def test_autograd_module_and_cpp_names():
    from thb_diff_b200 import THB_eval
    from thb_diff_b200.torch_funcs import THB_nn_module

    c = dict(np.load(f"{GOLDEN}/contraction.npz"))
    cp = torch.from_numpy(c["b_cp"]).cuda().requires_grad_(True)
    m = THB_nn_module(torch.from_numpy(c["b_Jm"]).cuda(), torch.from_numpy(c["b_phi"]).cuda(), torch.from_numpy(c["b_cumsum"]).cuda(), "cuda")
    out = m(cp)
    g = torch.from_numpy(c["b_grad_out"]).cuda()
    (out * g).sum().backward()
    assert _close(cp.grad.cpu().numpy(), c["b_bwd"])
    # CPU tensors through the cpp_* names (staged through the GPU)
    o2 = THB_eval.cpp_forward(torch.from_numpy(c["a_cp"]), torch.from_numpy(c["a_Jm"]), torch.from_numpy(c["a_phi"]), torch.from_numpy(c["a_cumsum"]))
    assert not o2.is_cuda and _close(o2.numpy(), c["a_fwd"])
    b2 = THB_eval.cpp_backward(torch.from_numpy(c["a_grad_out"]), torch.from_numpy(c["a_cp"]), torch.from_numpy(c["a_Jm"]), torch.from_numpy(c["a_phi"]),
                               torch.from_numpy(c["a_cumsum"]))
    assert _close(b2.numpy(), c["a_bwd"])
