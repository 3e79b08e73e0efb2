"""THB_eval.forward / backward (CSR contraction kernels) vs the reference's compiled compute_pts.cpp
(golden fixtures) and vs the oracle on seeded random CSR data."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err
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


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("cp_dim", [1, 2, 3, 4])
def test_random_csr_vs_oracle(cp_dim, idx_dtype):
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
    Jd, cd = torch.from_numpy(Jm).to(idx_dtype).cuda(), torch.from_numpy(cs).to(idx_dtype).cuda()
    out = THB_eval.forward(torch.from_numpy(cp).cuda(), Jd, torch.from_numpy(phi).cuda(), cd)
    assert _close(out.cpu().numpy(), O.eval_forward(cp, Jm, phi, cs))
    gcp = THB_eval.backward(torch.from_numpy(g).cuda(), torch.from_numpy(cp).cuda(), Jd, torch.from_numpy(phi).cuda(), cd)
    assert _close(gcp.cpu().numpy(), O.eval_backward(g, cp.shape, Jm, phi, cs))


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_long_segments_and_misaligned_views(idx_dtype):
    """Segments longer than one shared-memory stage of the bulk-staged forward (walked from global memory), chunks that end
    at the end of the arrays, and PHI / Jm views that are not 16-byte aligned (register-path kernels)."""
    from thb_diff_b200 import THB_eval

    rng = np.random.default_rng(5)
    nCP = 300
    ns = np.concatenate([rng.integers(0, 90, size=70), [5000, 0, 2304, 2305, 17, 768, 769, 767], rng.integers(60, 130, size=100), [3]])
    Jm = rng.integers(0, nCP, size=int(ns.sum())).astype(np.int64)
    phi = rng.random(len(Jm)).astype(np.float32)
    cs = np.cumsum(ns).astype(np.int64)
    cp = rng.standard_normal((nCP, 3)).astype(np.float32)
    ref = O.eval_forward(cp, Jm, phi, cs)
    cpd = torch.from_numpy(cp).cuda()
    Jd, pd, cd = torch.from_numpy(Jm).to(idx_dtype).cuda(), torch.from_numpy(phi).cuda(), torch.from_numpy(cs).to(idx_dtype).cuda()
    assert _close(THB_eval.forward(cpd, Jd, pd, cd).cpu().numpy(), ref)
    # the same data one element into a larger buffer: base pointers 4 bytes off a 16-byte boundary
    Jo = torch.zeros(len(Jm) + 1, dtype=idx_dtype, device="cuda"); Jo[1:] = Jd
    po = torch.zeros(len(Jm) + 1, device="cuda"); po[1:] = pd
    assert _close(THB_eval.forward(cpd, Jo[1:], po[1:], cd).cpu().numpy(), ref)


def test_staged_forward_with_int64_indices():
    """THB_CSR_BULK=2 routes 64-bit indices through the bulk-staged kernel too (read once per process: own process)."""
    import os, subprocess, sys

    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "from oracle import thb_oracle as O\n"
        "from thb_diff_b200 import THB_eval\n"
        "rng = np.random.default_rng(3)\n"
        "ns = np.concatenate([rng.integers(0, 140, size=400), [3000, 0, 1152, 1153, 768, 769], rng.integers(60, 70, size=300)])\n"
        "Jm = rng.integers(0, 500, size=int(ns.sum())).astype(np.int64); phi = rng.random(len(Jm)).astype(np.float32)\n"
        "cs = np.cumsum(ns).astype(np.int64); cp = rng.standard_normal((500, 3)).astype(np.float32)\n"
        "out = THB_eval.forward(torch.from_numpy(cp).cuda(), torch.from_numpy(Jm).cuda(), torch.from_numpy(phi).cuda(), torch.from_numpy(cs).cuda())\n"
        "ref = O.eval_forward(cp, Jm, phi, cs)\n"
        "err = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()\n"
        "assert err <= 1e-5, err\n"
        "print('ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, THB_CSR_BULK="2"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


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


def test_evaluate_and_prepare_data_for_evaluation_match_the_golden_contraction():
    """a16: prepare_data_for_evaluation + Evaluate [THB/torch_funcs.py:83-128] on the golden block3d space against the
    oracle contraction (values, gradient w.r.t. ctrl_pts and -- like the reference's pure-torch Evaluate -- w.r.t. PHI);
    CUDA tensors directly and CPU tensors staged through the GPU."""
    from helpers import load_golden, product_space_from_golden
    from oracle import thb_oracle as O
    from thb_diff_b200 import core
    from thb_diff_b200.torch_funcs import Evaluate, prepare_data_for_evaluation

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, ac)
    rng = np.random.default_rng(3)
    L = max(h.sh_fns.keys())
    cp_dict = {lev: rng.standard_normal(tuple(h.sh_fns[lev]) + (3,)).astype(np.float32) for lev in range(L + 1)}
    cp_flat = np.concatenate([cp_dict[lev].reshape(-1, 3) for lev in range(L + 1)])
    cs = np.cumsum(g["num_supp"])
    ref = O.eval_forward(cp_flat, g["Jm"], g["PHI"], cs)
    go = rng.standard_normal(ref.shape).astype(np.float32)
    refg = O.eval_backward(go, cp_flat.shape, g["Jm"], g["PHI"], cs)
    seg_np = np.repeat(np.arange(len(cs)), g["num_supp"])
    ref_gphi = (go[seg_np] * cp_flat[g["Jm"]]).sum(1)
    # the reference's list-of-spans form and the PointSupports form give the same packing
    spans_list = [(int(l), tuple(int(v) for v in c)) for l, c in zip(g["span_level"], g["span_cell"])]
    for spans in (supp, spans_list):
        for device in ("cuda", "cpu"):
            cp, Jm, PHI, seg, n = prepare_data_for_evaluation(g["PHI"], spans, g["num_supp"], cp_dict, ac, h.sh_fns, device)
            assert n == len(cs) and PHI.shape == (len(g["PHI"]), 1) and seg.shape == (len(g["PHI"]), 3)
            assert np.array_equal(Jm.cpu().numpy(), g["Jm"]) and np.array_equal(seg[:, 0].cpu().numpy(), seg_np)
            assert cp.device.type == device and Jm.device.type == device
            a = cp.clone().requires_grad_(True)
            phi = PHI.clone().requires_grad_(True)
            out = Evaluate(a, Jm, phi, seg, n, device)
            assert out.device.type == device
            assert rel_err(out.detach().cpu().numpy(), ref, scale=np.abs(ref).max()).max() <= TOL
            (out * torch.from_numpy(go).to(device)).sum().backward()
            assert rel_err(a.grad.cpu().numpy(), refg, scale=np.abs(refg).max()).max() <= TOL
            assert rel_err(phi.grad.cpu().numpy()[:, 0], ref_gphi, scale=np.abs(ref_gphi).max()).max() <= TOL


def test_compute_tensor_product_kernel():
    """a7: compute_tensor_product [THB/bspline_funcs.py:54-59] = einsum('i,j->ij') / einsum('i,j,k->ijk'), single and batched."""
    from thb_diff_b200 import bspline_funcs as B

    rng = np.random.default_rng(0)
    a, b, c = rng.standard_normal(3).astype(np.float32), rng.standard_normal(4).astype(np.float32), rng.standard_normal(5).astype(np.float32)
    r2 = B.compute_tensor_product([a, b])
    r3 = B.compute_tensor_product([torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda()])
    assert r2.device.type == "cpu" and r3.is_cuda
    assert np.array_equal(r2.numpy(), np.einsum("i,j->ij", a, b))
    assert np.allclose(r3.cpu().numpy(), np.einsum("i,j,k->ijk", a, b, c), rtol=1e-6, atol=0)
    A, Bm, Cm = (torch.from_numpy(rng.standard_normal((1000, n)).astype(np.float32)).cuda() for n in (4, 3, 4))
    rb = B.compute_tensor_product([A, Bm, Cm])
    assert rb.shape == (1000, 4, 3, 4)
    assert torch.allclose(rb, torch.einsum("qi,qj,qk->qijk", A, Bm, Cm), rtol=1e-6, atol=0)
    # the per-point tensor product of the path: basis values of a point in each direction (THB/core.py:652-677)
    U = np.array([0, 0, 0, 0, 0.25, 0.5, 0.75, 1, 1, 1, 1.0], dtype=np.float32)
    u = np.array([0.3, 0.6, 0.9], dtype=np.float32)
    N = [B.basis_fns_vmap(u[k:k + 1], U, 3)[0] for k in range(3)]
    tp = B.compute_tensor_product(N)
    assert abs(float(tp.sum()) - 1.0) < 1e-6


def test_deterministic_csr_backward():
    """THB_eval.backward with deterministic=True: same values as the atomic kernel, bit-identical between calls."""
    from thb_diff_b200 import engine

    c = dict(np.load(f"{GOLDEN}/contraction.npz"))
    cp = torch.from_numpy(c["c_cp"]).cuda()
    Jm = torch.from_numpy(c["c_Jm"]).cuda()
    phi = torch.from_numpy(c["c_phi"]).cuda()
    cs = torch.from_numpy(c["c_cumsum"]).cuda()
    g = torch.from_numpy(c["c_grad_out"]).cuda()
    a = engine.eval_backward(g, cp, Jm, phi, cs, deterministic=True)
    b = engine.eval_backward(g, cp, Jm, phi, cs, deterministic=True)
    assert torch.equal(a, b) and a.shape == cp.shape and a.dtype == torch.float32
    assert _close(a.cpu().numpy(), c["c_bwd"])
