"""BASELINE config 5 space (3-D cubic, 5 active levels) and config 3 (4 levels) through the CUDA path against
(a) the table-free C evaluator (truncated functions from their definition, oracle_direct_eval) at 2e4 points concentrated
on the level interfaces and (b) the float64 C oracle on every point: cells / support ids bit-exact, PHI, geometry and the
three first-derivative channels (the reference's `derivative` outputs, THB/core.py:686-701, plus the true gradient)."""
import numpy as np
import pytest
import torch

from helpers import compare_with_direct, direct_tables, host_tables_from_packed, interface_points, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module", params=[4, 5])
def space(request):
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=request.param)
    P = PackedHierarchy.from_space(sp, device="cuda")
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
    return request.param, sp, P, cp


def test_cuda_phi_against_the_table_free_evaluator(space):
    from thb_diff_b200 import core, engine

    levels, sp, P, cp = space
    prm = interface_points(20000, 3, seed=10 + levels)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    slot = supp.slot.cpu().numpy()
    assert (slot >= 0).all()
    info = P.cellinfo_host[slot]
    off, nnz = supp.offsets()
    chans = [engine.CH_VALUE] + engine.grad_channels(3)
    phis = [p.cpu().numpy() for p in engine.basis_tiled(P, supp.plan, off, nnz, chans)]
    r = compare_with_direct(direct_tables(sp), prm, info, supp.Jm().cpu().numpy(), phis[0], off.cpu().numpy().astype(np.int64), TOL, dPHI=phis[1:])
    assert set(r["lev"].tolist()) == set(range(levels))
    assert np.abs(np.add.reduceat(phis[0], off.cpu().numpy()) - 1).max() < 1e-5      # partition of unity


def test_cuda_geometry_and_derivatives_against_the_c_oracle(space):
    from oracle import thb_oracle as O
    from thb_diff_b200 import core, engine

    levels, sp, P, cp = space
    prm = np.concatenate([interface_points(60000, 3, seed=20 + levels), np.random.default_rng(levels).random((60000, 3)).astype(np.float32)])
    ht = host_tables_from_packed(P)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    S = np.asarray(ns).astype(np.int64)
    off = np.cumsum(S) - S
    cph = cp.cpu().numpy()
    ref = ht.eval(prm, cph)
    assert np.array_equal(supp.slot.cpu().numpy(), ref["slot"])
    scale = np.abs(ref["out"]).max()
    # at most four channels per call: value + the three first derivatives (config 5's outputs), then the mixed partial
    outs = engine.eval_fused(P, supp.plan, cp, [engine.CH_VALUE] + engine.grad_channels(3))
    outs = outs + engine.eval_fused(P, supp.plan, cp, [engine.mixed_channel(3)])
    assert rel_err(outs[0].cpu().numpy(), ref["out"], scale=scale).max() <= TOL
    Jm = supp.Jm().cpu().numpy()
    cs = np.cumsum(S)
    for j, mode in enumerate([1, 2, 4, 7]):
        r = ht.eval(prm, cph, mode=mode, want_phi=True, offsets=off, nnz=int(S.sum()))
        o = outs[1 + j].cpu().numpy()
        cond = O.eval_forward(np.abs(cph), Jm, np.abs(r["phi"]), cs)
        assert (np.abs(o - r["out"]) <= TOL * np.maximum(cond, 1.0)).all(), mode
        if mode != 7:   # first derivatives of an O(1) map: relative to their own size too
            assert rel_err(o, r["out"], scale=np.abs(r["out"]).max()).max() <= TOL, mode
    # backward (value channel) against the oracle's transposed contraction
    go = np.random.default_rng(1).standard_normal((len(prm), 3)).astype(np.float32)
    phi = ht.eval(prm, None, want_phi=True, offsets=off, nnz=int(S.sum()))["phi"]
    refg = O.eval_backward(go, cph.shape, Jm, phi, cs)
    g = engine.eval_fused_backward(P, supp.plan, torch.from_numpy(go).cuda()).cpu().numpy()
    assert rel_err(g, refg, scale=np.abs(refg).max()).max() <= TOL
    rows = engine.eval_fused_backward_rows(P, supp.plan, torch.from_numpy(go).cuda()).cpu().numpy()
    ids = P.active_rows()[0].cpu().numpy()
    assert rel_err(rows, refg[ids], scale=np.abs(refg).max()).max() <= TOL


def test_c5_grid_slab_properties():
    """A 64-plane slab of the 512^3 grid of config 5: partition of unity, Greville reproduction, identity Jacobian."""
    from thb_diff_b200 import configs, engine
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=5)
    P = PackedHierarchy.from_space(sp, device="cuda")
    ax = torch.from_numpy(np.linspace(1e-5, 1.0, 512, endpoint=False)).float().cuda()
    prm = torch.stack(torch.meshgrid(ax[224:288], ax, ax, indexing="ij"), -1).reshape(-1, 3).contiguous()
    plan = engine.Plan(P, prm)
    assert plan.stats()[0] == prm.shape[0]
    grev = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    outs = engine.eval_fused(P, plan, grev, [engine.CH_VALUE] + engine.grad_channels(3))
    assert float((outs[0] - prm).abs().max()) < 1e-5
    eye = torch.eye(3, device="cuda")
    for k in range(3):
        assert float((outs[1 + k] - eye[k]).abs().max()) < 1e-3
    pou = engine.eval_fused(P, plan, torch.ones((P.nCP, 3), device="cuda"))[0]
    assert float((pou - 1).abs().max()) < 1e-5
    lev = P.cellinfo[:, 0][plan.slot[: plan.n].long()]
    assert set(torch.unique(lev).tolist()) == {0, 1, 2, 3, 4}
