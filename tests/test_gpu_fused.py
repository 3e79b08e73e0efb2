"""Fused basis+evaluation kernels (forward, derivative channels, backward) vs the oracle."""
import numpy as np
import pytest
import torch

from helpers import SPACES, load_golden, product_space_from_golden, rel_err
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
IDENTITY_OK = [s for s in SPACES if s != "twoblock2d"]


def _setup(name, onehot=True, cp_dim=3, seed=0):
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda", onehot_direct=onehot)
    plan = engine.Plan(P, g["params"])
    cp = np.random.default_rng(seed).standard_normal((P.nCP, cp_dim)).astype(np.float32)
    cs = np.cumsum(g["num_supp"])
    return g, h, P, plan, cp, cs


@pytest.mark.parametrize("formulation", ["local_net", "per_point", "tile_net"])
@pytest.mark.parametrize("onehot", [True, False])
@pytest.mark.parametrize("name", IDENTITY_OK)
def test_fused_forward(name, onehot, formulation):
    from thb_diff_b200 import engine

    g, h, P, plan, cp, cs = _setup(name, onehot)
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], cs)
    out = engine.eval_fused(P, plan, torch.from_numpy(cp).cuda(), formulation=formulation)[0].cpu().numpy()
    assert rel_err(out, ref, scale=np.abs(ref).max()).max() <= TOL


@pytest.mark.parametrize("ppt", ["1", "2"])
def test_fused_forward_points_per_thread_variants(ppt, monkeypatch):
    from thb_diff_b200 import engine

    monkeypatch.setenv("THB_FUSED_PPT", ppt)
    g, h, P, plan, cp, cs = _setup("block3d")
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], cs)
    out = engine.eval_fused(P, plan, torch.from_numpy(cp).cuda())[0].cpu().numpy()
    assert rel_err(out, ref, scale=np.abs(ref).max()).max() <= TOL


def test_fused_cp_dim4():
    from thb_diff_b200 import engine

    g, h, P, plan, cp, cs = _setup("block3d", cp_dim=4)
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], cs)
    out = engine.eval_fused(P, plan, torch.from_numpy(cp).cuda())[0].cpu().numpy()
    assert out.shape == ref.shape and rel_err(out, ref, scale=np.abs(ref).max()).max() <= TOL


@pytest.mark.parametrize("formulation", ["local_net", "per_point"])
@pytest.mark.parametrize("name", ["graded2d", "block3d"])
def test_greville_control_points_reproduce_the_parameters(name, formulation):
    """Linear reproduction: evaluating with Greville abscissae as control points returns the params, and
    the first derivatives are the identity."""
    from thb_diff_b200 import configs, engine

    g, h, P, plan, _, _ = _setup(name)
    cp = torch.from_numpy(configs.greville_ctrl_pts(h, bump=0.0)).cuda()
    out = engine.eval_fused(P, plan, cp, formulation=formulation)[0].cpu().numpy()
    d = h.ndim
    assert np.abs(out[:, :d] - g["params"]).max() < 1e-5
    ders = engine.eval_fused(P, plan, cp, engine.grad_channels(d), formulation=formulation)
    for k in range(d):
        J = ders[k].cpu().numpy()[:, :d]
        e = np.zeros(d)
        e[k] = 1
        assert np.abs(J - e).max() < 2e-4


@pytest.mark.parametrize("cp_dim", [3, 4])
@pytest.mark.parametrize("formulation", ["local_net", "per_point"])
@pytest.mark.parametrize("name", IDENTITY_OK)
def test_fused_backward(name, formulation, cp_dim):
    from thb_diff_b200 import engine

    g, h, P, plan, cp, cs = _setup(name, cp_dim=cp_dim)
    go = np.random.default_rng(1).standard_normal((len(cs), cp_dim)).astype(np.float32)
    ref = O.eval_backward(go, cp.shape, g["Jm"], g["PHI"], cs)
    got = engine.eval_fused_backward(P, plan, torch.from_numpy(go).cuda(), formulation=formulation).cpu().numpy()
    assert rel_err(got, ref, scale=np.abs(ref).max()).max() <= TOL


def test_fused_module_autograd_matches_csr_path():
    from thb_diff_b200 import core
    from thb_diff_b200.torch_funcs import THB_fused_module, THB_nn_module, prepare_data_for_CUDA_evaluation

    g, h, P, plan, cp, cs = _setup("block3d")
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, P)
    PHI = core.THB_basis_fns(g["params"], supp, ns, P, h.sh_fns, h.knotvectors, h.degrees)
    cpt, Jm, phi, cumsum, dev = prepare_data_for_CUDA_evaluation(PHI, supp, ns, cp, None, h.sh_fns, "cuda")
    a = cpt.clone().requires_grad_(True)
    b = cpt.clone().requires_grad_(True)
    out_a = THB_nn_module(Jm, phi, cumsum, "cuda")(a)
    out_b = THB_fused_module(P, g["params"])(b)
    w = torch.randn_like(out_a)
    (out_a * w).sum().backward()
    (out_b * w).sum().backward()
    assert rel_err(out_b.detach().cpu().numpy(), out_a.detach().cpu().numpy(), scale=float(out_a.detach().abs().max())).max() <= TOL
    assert rel_err(b.grad.cpu().numpy(), a.grad.cpu().numpy(), scale=float(a.grad.abs().max())).max() <= TOL


def test_many_points_per_cell_and_ragged_tail():
    """More points per cell than one work item holds, item tails, repeated evaluation with one plan."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    rng = np.random.default_rng(2)
    prm = rng.random((20011, 2)).astype(np.float32)
    prm[:5000] = prm[:5000] * 0.05 + np.array([0.31, 0.41], dtype=np.float32)  # ~5000 points in very few cells
    plan = engine.Plan(P, prm)
    cp = rng.standard_normal((P.nCP, 3)).astype(np.float32)
    cpt = torch.from_numpy(cp).cuda()
    out = engine.eval_fused(P, plan, cpt, formulation="per_point")[0]
    out2 = engine.eval_fused(P, plan, cpt, formulation="local_net")[0]
    sp = __import__("helpers").oracle_space_from_golden(g)
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, sp.Coeff)
    idx = rng.choice(len(prm), 300, replace=False)
    ref = np.stack([(lambda r: (r[3][:, None] * cp[r[2]].astype(np.float64)).sum(0))(de.point(prm[i])) for i in idx])
    scale = np.abs(ref).max()
    assert rel_err(out.cpu().numpy()[idx], ref, scale=scale).max() <= TOL
    assert rel_err(out2.cpu().numpy()[idx], ref, scale=scale).max() <= TOL
    assert torch.equal(engine.eval_fused(P, plan, cpt, formulation="per_point")[0], out)  # deterministic forward
    assert torch.equal(engine.eval_fused(P, plan, cpt)[0], out2)


def test_host_pipeline_matches_device_path():
    """engine.HostPipeline (chunked H2D / compute / D2H on three streams) == one-shot device evaluation."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    rng = np.random.default_rng(5)
    n = 300001
    prm_np = rng.random((n, 3)).astype(np.float32)
    prm_np[7::1001] += 1.5                                   # a few points outside the domain: their rows must come back zero
    prm = torch.from_numpy(prm_np).pin_memory()
    cp = torch.from_numpy(rng.standard_normal((P.nCP, 3)).astype(np.float32)).cuda()
    ref = engine.eval_fused(P, engine.Plan(P, prm.cuda()), cp)[0]
    assert float(ref[7::1001].abs().max()) == 0.0
    out = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    pipe = engine.HostPipeline(P, n, n_chunks=4)
    for _ in range(2):  # second run reuses plans and buffers
        out.fill_(123.0)
        pipe.run(prm, cp, out)
        torch.cuda.synchronize()
        # same points, possibly a different sub-batch path (summation order) than in the one-shot plan
        assert float((out.cuda() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("n", [0, 1, 33])
def test_tiny_batches(n):
    """empty / single-point / ragged batches through the whole pipeline (plan, PHI, fused forward + backward)"""
    from thb_diff_b200 import core, engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    prm = g["params"][:n].astype(np.float32).reshape(n, 3)
    supp, ns = core.compute_active_span(prm, h.knotvectors, h.cells, h.degrees, P)
    PHI = core.THB_basis_fns(prm, supp, ns, P, h.sh_fns, h.knotvectors, h.degrees)
    nnz = int(g["num_supp"][:n].sum())
    assert len(ns) == n and PHI.shape == (nnz,)
    assert rel_err(PHI.cpu().numpy(), g["PHI"][:nnz]).max(initial=0.0) <= TOL
    cp = torch.randn((P.nCP, 3), device="cuda")
    out = engine.eval_fused(P, supp.plan, cp)[0]
    assert out.shape == (n, 3)
    grad = engine.eval_fused_backward(P, supp.plan, torch.ones((n, 3), device="cuda"))
    assert grad.shape == (P.nCP, 3)
    if n == 0:
        assert float(grad.abs().max()) == 0.0
    else:
        ref = O.eval_forward(cp.cpu().numpy(), g["Jm"][:nnz], g["PHI"][:nnz], np.cumsum(g["num_supp"][:n]))
        assert rel_err(out.cpu().numpy(), ref, scale=np.abs(ref).max()).max() <= TOL


def test_domain_boundary_and_outside_points_fused():
    """u == last knot evaluates in the last cell (reference span rule); outside / NaN points give zeros"""
    from thb_diff_b200 import configs, engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    prm = np.array([[1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.0, 0.3, 0.0], [1.5, 0.5, 0.5], [np.nan, 0.1, 0.1], [0.5, -1e-6, 0.5]], dtype=np.float32)
    plan = engine.Plan(P, prm)
    cp = torch.from_numpy(configs.greville_ctrl_pts(h, bump=0.0)).cuda()
    out = engine.eval_fused(P, plan, cp)[0].cpu().numpy()
    assert np.abs(out[:3] - prm[:3]).max() < 1e-6      # linear reproduction right at the boundary
    assert np.all(out[3:] == 0.0)                       # out-of-domain points: flagged (slot -1), zero output
    assert plan.slot[:6].cpu().numpy()[3:].tolist() == [-1, -1, -1]


def test_local_net_needs_a_workspace_and_nets_are_reusable():
    """C-ABI contract of the nets entry points: thb_eval_fused(local_net) without a workspace is an error, not a
    fallback; nets built once (all cells) serve several batches until the control points change."""
    import ctypes as C

    from thb_diff_b200 import _lib, engine

    g, h, P, plan, cp, cs = _setup("block3d")
    cpt = torch.from_numpy(cp).cuda()
    out = torch.zeros((plan.n, 3), device="cuda")
    arr = (C.c_void_p * 1)(out.data_ptr())
    ch = (C.c_int32 * 1)(0)
    rc = _lib.lib.thb_eval_fused(C.byref(P.desc()), C.byref(plan.desc), _lib.ptr(plan.params), _lib.ptr(cpt), 3, ch, 1, arr, 0, None,
                                 _lib.stream_ptr(P.device))
    assert rc != 0 and b"workspace" in _lib.lib.thb_last_error()
    assert _lib.lib.thb_net_workspace_floats(C.byref(P.desc()), 0) == P.nslots * 4 * P.TS
    assert _lib.lib.thb_net_workspace_floats(C.byref(P.desc()), 1) == P.nslots * 8 * P.TS
    nets = engine.cell_nets(P, cpt)                                  # every active cell
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], cs)
    a = engine.eval_nets(P, plan, nets, 3)[0].cpu().numpy()
    assert rel_err(a, ref, scale=np.abs(ref).max()).max() <= TOL
    half = engine.Plan(P, g["params"][: len(cs) // 2])               # another batch, same nets
    b = engine.eval_nets(P, half, nets, 3)[0].cpu().numpy()
    assert rel_err(b, ref[: len(cs) // 2], scale=np.abs(ref).max()).max() <= TOL
    # control-point nets: derivative channels need the relative planes; monomial nets serve them from the same four planes
    assert _lib.lib.thb_net_workspace_floats(C.byref(P.desc()), 3) == P.nslots * 4 * P.TS
    nets_cp = engine.cell_nets(P, cpt, monomial=False)
    with pytest.raises(RuntimeError):
        engine.eval_nets(P, plan, nets_cp, 3, channels=engine.grad_channels(3), with_derivatives=False)
    if nets._thb_flags & engine.NETS_MONOMIAL:
        d_mono = engine.eval_nets(P, plan, engine.cell_nets(P, cpt), 3, channels=engine.grad_channels(3))
        d_cp = engine.eval_nets(P, plan, engine.cell_nets(P, cpt, with_derivatives=True, monomial=False).clone(), 3, channels=engine.grad_channels(3))
        for x, y in zip(d_mono, d_cp):
            assert float((x - y).abs().max()) <= 2e-5 * max(float(y.abs().max()), 1.0)


def test_rows_in_plan_order():
    """thb_plan_desc::rows_in_plan_order: the same numbers, rows permuted by Plan.caller_index() -- forward and backward."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    rng = np.random.default_rng(9)
    prm = rng.random((50021, 3)).astype(np.float32)
    cp = torch.from_numpy(rng.standard_normal((P.nCP, 3)).astype(np.float32)).cuda()
    go = torch.from_numpy(rng.standard_normal((len(prm), 3)).astype(np.float32)).cuda()
    plan = engine.Plan(P, prm)
    ref = engine.eval_fused(P, plan, cp)[0]
    refg = engine.eval_fused_backward(P, plan, go)
    perm = plan.caller_index()
    assert torch.equal(torch.sort(perm).values, torch.arange(len(prm), device="cuda"))
    plan.set_rows_in_plan_order(True)
    out = engine.eval_fused(P, plan, cp)[0]
    assert torch.equal(out, ref[perm])
    gs = engine.eval_fused_backward(P, plan, go[perm].contiguous())
    assert float((gs - refg).abs().max()) <= 1e-5 * float(refg.abs().max())
    with pytest.raises(RuntimeError):
        engine.eval_fused(P, plan, cp, formulation="per_point")      # only the nets path knows the row order
    plan.set_rows_in_plan_order(False)
    assert torch.equal(engine.eval_fused(P, plan, cp)[0], ref)


@pytest.mark.parametrize("cp_dim", [3, 4])
@pytest.mark.parametrize("name", IDENTITY_OK)
def test_monomial_nets_and_control_point_nets_against_the_oracle(name, cp_dim):
    """Both layouts of the local control nets -- control points evaluated with Cox-de Boor, and the monomial form evaluated
    by nested Horner (the default) -- value and every first-derivative channel against the oracle contraction."""
    from thb_diff_b200 import engine

    g, h, P, plan, cp, cs = _setup(name, cp_dim=cp_dim)
    d = h.ndim
    cpt = torch.from_numpy(cp).cuda()
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], cs)
    chans = [engine.CH_VALUE] + engine.grad_channels(d)
    outs = {}
    for mono in (False, True):
        nets = engine.cell_nets(P, cpt, with_derivatives=True, monomial=mono)
        assert bool(nets._thb_flags & engine.NETS_MONOMIAL) == mono
        outs[mono] = [o.cpu().numpy() for o in engine.eval_nets(P, plan, nets, cp_dim, chans)]
        assert rel_err(outs[mono][0], ref, scale=np.abs(ref).max()).max() <= TOL
        nets4 = engine.cell_nets(P, cpt, with_derivatives=False, monomial=mono).clone()
        v = engine.eval_nets(P, plan, nets4, cp_dim, monomial=mono)[0].cpu().numpy()
        assert rel_err(v, ref, scale=np.abs(ref).max()).max() <= TOL
    # derivative channels: the two layouts against each other and against the oracle's derivative PHI (golden spaces carry
    # the mixed partial only, so the per-direction reference comes from the C oracle on the packed tables)
    from helpers import host_tables_from_packed

    ht = host_tables_from_packed(P)
    for k in range(d):
        r = ht.eval(g["params"], cp, mode=1 << k)["out"]
        for mono in (False, True):
            assert rel_err(outs[mono][1 + k], r, scale=max(np.abs(r).max(), 1.0)).max() <= 2 * TOL, (k, mono)


@pytest.mark.parametrize("name", ["block3d", "graded2d"])
def test_light_plan_and_record_plan_give_the_same_results(name):
    """Plan(light=True): only the permutation is built and k_net_eval gathers the parameters through it; Plan(light=False):
    16-byte records.  Same geometry bit for bit (the arithmetic per point is identical), work items of many sizes."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    rng = np.random.default_rng(4)
    prm = np.concatenate([g["params"], rng.random((70_001, h.ndim)).astype(np.float32), np.full((3, h.ndim), 2.0, np.float32)])
    cp = torch.from_numpy(rng.standard_normal((P.nCP, 3)).astype(np.float32)).cuda()
    a_plan, b_plan = engine.Plan(P, prm, light=True), engine.Plan(P, prm, light=False)
    assert a_plan.light and not b_plan.light
    chans = [engine.CH_VALUE] + engine.grad_channels(h.ndim)
    for ch in ([engine.CH_VALUE], chans):
        a = engine.eval_fused(P, a_plan, cp, ch)
        assert a_plan.desc.light == 1            # still light: no records were needed
        b = engine.eval_fused(P, b_plan, cp, ch)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # consumers that need records materialise them
    go = torch.randn((len(prm), 3), device="cuda")
    ga = engine.eval_fused_backward(P, a_plan, go)
    assert a_plan.desc.light == 2
    gb = engine.eval_fused_backward(P, b_plan, go)
    assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max())
    assert torch.equal(engine.eval_fused(P, a_plan, cp)[0], engine.eval_fused(P, b_plan, cp)[0])
