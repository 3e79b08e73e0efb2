"""BASELINE config 4 (differentiable fitting): compact-row backward, one-pass MSE loss + gradient, row-restricted Adam and
the FitEngine loop -- against the dense fused backward (itself pinned to the oracle in test_gpu_fused.py), torch
expressions in float64 and torch.optim.Adam.  The reference's loop is torch ops around THBEval (THB/torch_funcs.py:8-47)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, product_space_from_golden, rel_err
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(name="block3d", cp_dim=3, seed=0):
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    plan = engine.Plan(P, g["params"])
    cp = np.random.default_rng(seed).standard_normal((P.nCP, cp_dim)).astype(np.float32)
    return g, h, P, plan, cp


def test_active_rows_are_exactly_the_supports_of_active_cells():
    g, h, P, plan, cp = _setup()
    ids, row_map = P.active_rows()
    want = np.unique(P.supp_jm.cpu().numpy())
    assert np.array_equal(ids.cpu().numpy(), want)
    rm = row_map.cpu().numpy()
    assert np.array_equal(rm[want], np.arange(len(want))) and (np.delete(rm, want) == -1).all()
    assert P.n_active == len(want)


@pytest.mark.parametrize("cp_dim", [3, 4])
@pytest.mark.parametrize("name", ["block3d", "readme3d", "graded2d"])
def test_backward_rows_equals_dense_backward_and_oracle(name, cp_dim):
    from thb_diff_b200 import engine

    g, h, P, plan, cp = _setup(name, cp_dim)
    cs = np.cumsum(g["num_supp"])
    go = np.random.default_rng(1).standard_normal((len(cs), cp_dim)).astype(np.float32)
    ref = O.eval_backward(go, cp.shape, g["Jm"], g["PHI"], cs)
    rows = engine.eval_fused_backward_rows(P, plan, torch.from_numpy(go).cuda())
    ids = P.active_rows()[0].long()
    dense = torch.zeros((P.nCP, cp_dim), device="cuda")
    dense[ids] = rows
    assert rel_err(dense.cpu().numpy(), ref, scale=np.abs(ref).max()).max() <= TOL
    # everything outside the active rows is zero in the reference gradient too
    mask = np.ones(P.nCP, bool)
    mask[ids.cpu().numpy()] = False
    assert np.abs(ref[mask]).max(initial=0.0) == 0.0


@pytest.mark.parametrize("n", [1, 3, 4, 1001, 3 * 100003])
def test_mse_grad_one_pass(n):
    from thb_diff_b200 import engine

    gen = torch.Generator(device="cuda").manual_seed(n)
    out = torch.randn(n, device="cuda", generator=gen)
    tgt = torch.randn(n, device="cuda", generator=gen)
    scale = 1.0 / n
    gout = torch.empty_like(out)
    loss = torch.zeros(1, device="cuda")
    ws = engine.mse_workspace(out.device)
    for _ in range(2):  # the workspace is left ready for the next call
        engine.mse_grad(out.view(-1, 1), tgt.view(-1, 1), scale, gout.view(-1, 1), loss, ws)
        want = float(((out.double() - tgt.double()) ** 2).sum() * scale)
        assert abs(float(loss) - want) <= 1e-6 * max(abs(want), 1e-30)
        assert torch.equal(gout, (2.0 * scale) * (out - tgt)) or float((gout - 2.0 * scale * (out - tgt)).abs().max()) <= 1e-7 * scale * 10


def test_mse_grad_is_deterministic():
    from thb_diff_b200 import engine

    out = torch.randn(1_000_003, device="cuda")
    tgt = torch.randn(1_000_003, device="cuda")
    ws = engine.mse_workspace(out.device)
    vals = []
    for _ in range(4):
        loss = torch.zeros(1, device="cuda")
        engine.mse_grad(out.view(-1, 1), tgt.view(-1, 1), 0.5, torch.empty_like(out).view(-1, 1), loss, ws)
        vals.append(float(loss))
    assert len(set(vals)) == 1


@pytest.mark.parametrize("cp_dim", [3, 4])
def test_adam_rows_is_torch_adam_on_the_rows_it_can_change(cp_dim):
    from thb_diff_b200 import engine

    torch.manual_seed(0)
    nrow, na = 5000, 777
    x = torch.randn(nrow, cp_dim, device="cuda")
    ids = torch.randperm(nrow, device="cuda")[:na].sort().values.to(torch.int32)
    ref = x.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2)
    m, v = torch.zeros(na, cp_dim, device="cuda"), torch.zeros(na, cp_dim, device="cuda")
    step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(6):
        g = torch.randn(na, cp_dim, device="cuda")
        dense = torch.zeros_like(x)
        dense[ids.long()] = g
        ref.grad = dense
        opt.step()
        if it % 2 == 0:
            engine.adam_rows(x, ids, g, m, v, 1e-2, (0.9, 0.999), 1e-8, step_dev=step_dev)
        else:  # host-side step number, device counter kept in sync by hand
            engine.adam_rows(x, ids, g, m, v, 1e-2, (0.9, 0.999), 1e-8, step=it + 1)
            step_dev += 1
        assert float((x - ref.detach()).abs().max()) <= 2e-6
    assert int(step_dev) == 6


@pytest.mark.parametrize("use_graph", [False, True])
def test_fit_engine_follows_autograd_adam(use_graph):
    """FitEngine (compact rows, no autograd) against THB_fused_module + composed MSE + torch.optim.Adam (dense)."""
    from thb_diff_b200 import configs
    from thb_diff_b200.distributed import FitEngine
    from thb_diff_b200.torch_funcs import THB_fused_module

    g, h, P, plan, _ = _setup("block3d")
    rng = np.random.default_rng(5)
    prm = torch.from_numpy(rng.random((30011, 3)).astype(np.float32)).cuda()
    tgt = torch.stack([prm[:, 0], prm[:, 1], prm[:, 2] + 0.1 * torch.sin(6.28 * prm[:, 0]) * torch.sin(6.28 * prm[:, 1])], 1).contiguous()
    x0 = torch.from_numpy(configs.greville_ctrl_pts(h, bump=0.0)).cuda()
    a = x0.clone().requires_grad_(True)
    opt = torch.optim.Adam([a], lr=1e-2)
    mod = THB_fused_module(P, prm)
    b = x0.clone()
    eng = FitEngine(P, prm, tgt, prm.shape[0], lr=1e-2, use_graph=use_graph)
    la, lb = [], []
    for _ in range(8):
        opt.zero_grad()
        loss = ((mod(a) - tgt) ** 2).mean()
        loss.backward()
        opt.step()
        la.append(float(loss))
        lb.append(float(eng.step(b)))
    assert lb[-1] < lb[0]
    assert np.allclose(la, lb, rtol=2e-4, atol=1e-9)
    assert float((a.detach() - b).abs().max()) <= 5e-5   # 8 Adam steps of size 1e-2: sign-stable, f32 noise only


def test_sharded_fit_runs_with_plan_order_rows_and_active_row_allreduce():
    from thb_diff_b200 import configs
    from thb_diff_b200.distributed import ShardedFit

    g, h, P, plan, _ = _setup("block3d")
    prm = torch.rand((20000, 3), device="cuda")
    tgt = prm.clone()
    x = (torch.from_numpy(configs.greville_ctrl_pts(h, bump=0.0)).cuda() + 0.01).requires_grad_(True)
    opt = torch.optim.Adam([x], lr=1e-2)
    fit = ShardedFit(P, prm, tgt, prm.shape[0])
    assert fit.module.perm is not None
    l0 = float(fit.step(x, opt))
    for _ in range(10):
        l1 = float(fit.step(x, opt))
    assert l1 < l0


def test_tensors_on_a_non_current_device():
    """the library launches on the current device -- every wrapper switches to the tensors' device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    torch.cuda.set_device(0)
    P = PackedHierarchy.from_space(h, device="cuda:1")
    plan = engine.Plan(P, g["params"])
    cp = np.random.default_rng(0).standard_normal((P.nCP, 3)).astype(np.float32)
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], np.cumsum(g["num_supp"]))
    out = engine.eval_fused(P, plan, torch.from_numpy(cp).to("cuda:1"))[0]
    assert out.device.index == 1
    assert rel_err(out.cpu().numpy(), ref, scale=np.abs(ref).max()).max() <= TOL


@pytest.mark.parametrize("cp_dim", [3, 4])
@pytest.mark.parametrize("name", ["block3d", "readme3d", "graded2d"])
def test_ordered_backward_matches_the_oracle_and_is_bit_reproducible(name, cp_dim):
    """deterministic=True: no atomics, every sum in a fixed order.  Two plans built independently (their scatter order
    differs from run to run) and made canonical give BIT-IDENTICAL gradients; the values match the oracle."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    d = h.ndim
    rng = np.random.default_rng(9)
    prm = np.concatenate([g["params"], rng.random((150_000, d)).astype(np.float32)])   # cells with several work items
    go = torch.from_numpy(rng.standard_normal((len(prm), cp_dim)).astype(np.float32)).cuda()
    grads, rows = [], []
    for _ in range(3):
        plan = engine.Plan(P, prm).canonicalize()
        grads.append(engine.eval_fused_backward(P, plan, go, deterministic=True))
        rows.append(engine.eval_fused_backward_rows(P, plan, go, deterministic=True))
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    assert torch.equal(rows[0], rows[1])
    ids = P.active_rows()[0].long()
    assert torch.equal(grads[0][ids], rows[0])
    fast = engine.eval_fused_backward(P, engine.Plan(P, prm), go, deterministic=False)
    scale = float(fast.abs().max())
    assert float((fast - grads[0]).abs().max()) <= 1e-5 * scale
    # against the oracle on the golden points alone
    cs = np.cumsum(g["num_supp"])
    cp_shape = (P.nCP, cp_dim)
    go_s = go[: len(cs)].cpu().numpy()
    ref = O.eval_backward(go_s, cp_shape, g["Jm"], g["PHI"], cs)
    got = engine.eval_fused_backward(P, engine.Plan(P, g["params"]).canonicalize(), torch.from_numpy(go_s).cuda(), deterministic=True).cpu().numpy()
    assert rel_err(got, ref, scale=np.abs(ref).max()).max() <= TOL


def test_deterministic_fit_is_bit_reproducible():
    from thb_diff_b200 import configs
    from thb_diff_b200.distributed import FitEngine

    g, h, P, plan, _ = _setup("block3d")
    prm = torch.rand((100_003, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    tgt = torch.stack([prm[:, 0], prm[:, 1], prm[:, 2] + 0.05 * torch.sin(9.0 * prm[:, 0])], 1).contiguous()
    x0 = torch.from_numpy(configs.greville_ctrl_pts(h, bump=0.0)).cuda()
    runs = []
    for _ in range(2):
        x = x0.clone()
        eng = FitEngine(P, prm, tgt, prm.shape[0], lr=1e-2, deterministic=True)
        losses = [float(eng.step(x)) for _ in range(5)]
        runs.append((x.clone(), losses))
    assert runs[0][1] == runs[1][1] and torch.equal(runs[0][0], runs[1][0])
    assert runs[0][1][-1] < runs[0][1][0]
