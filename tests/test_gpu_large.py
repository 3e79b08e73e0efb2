"""BASELINE config 3 space (3-D cubic, 4 active levels) at reduced point counts plus size-independent
properties at the full 256^3 grid: partition of unity, linear reproduction, adjointness of the backward."""
import numpy as np
import pytest
import torch

from helpers import host_tables_from_packed, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def c3():
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=4)
    P = PackedHierarchy.from_space(sp, device="cuda")
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
    return sp, P, cp


def test_c3_against_c_oracle_everywhere(c3):
    """grid + random points (unsorted), every point compared with the float64 C oracle"""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import core, engine

    sp, P, cp = c3
    rng = np.random.default_rng(3)
    prm = np.concatenate([B.generate_parametric_coordinates((48, 40, 56)), rng.random((150000, 3))]).astype(np.float32)
    ht = host_tables_from_packed(P)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    S = np.asarray(ns).astype(np.int64)
    off = np.cumsum(S) - S
    ref = ht.eval(prm, cp.cpu().numpy(), want_phi=True, offsets=off, nnz=int(S.sum()))
    assert np.array_equal(supp.slot.cpu().numpy(), ref["slot"])                  # bit-exact cell lookup
    assert np.array_equal(S, P.cellinfo_host[ref["slot"], 5])
    PHI = core.THB_basis_fns(prm, supp, ns, P, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert rel_err(PHI.cpu().numpy(), ref["phi"]).max() <= TOL
    scale = np.abs(ref["out"]).max()
    for formulation in ("local_net", "per_point", "tile_net"):
        out = engine.eval_fused(P, supp.plan, cp, formulation=formulation)[0].cpu().numpy()
        assert rel_err(out, ref["out"], scale=scale).max() <= TOL
    # true gradient channels against the oracle's derivative modes.  The derivative basis values are
    # O(p * #cells) ~ 400 here while the geometry derivative they sum to is O(1): float32 cancellation bounds
    # the error by 1e-5 of the magnitude of the SUMMED TERMS, cond = sum_s |dPHI_s| |CP_s| per point.
    from oracle import thb_oracle as O

    Jm = supp.Jm().cpu().numpy()
    cs = np.cumsum(S)
    cph = cp.cpu().numpy()
    ders = {f: engine.eval_fused(P, supp.plan, cp, engine.grad_channels(3), formulation=f) for f in ("local_net", "per_point")}
    for k in range(3):
        r = ht.eval(prm, cph, mode=1 << k, want_phi=True, offsets=off, nnz=int(S.sum()))
        cond = O.eval_forward(np.abs(cph), Jm, np.abs(r["phi"]), cs)
        for f, dd in ders.items():
            o = dd[k].cpu().numpy()
            assert (np.abs(o - r["out"]) <= TOL * np.maximum(cond, 1.0)).all(), f
            assert rel_err(o, r["out"], scale=np.abs(r["out"]).max()).max() <= TOL, f  # thanks to the relative-CP contraction


def test_c3_sampled_against_table_free_oracle(c3):
    """independent of the tile tables: the reference's truncated functions evaluated from their definition"""
    from oracle import thb_oracle as O
    from thb_diff_b200 import core

    sp, P, cp = c3
    Coeff = {l: {k: O.knot_insertion_matrix(sp.knotvectors[l][k], sp.knotvectors[l + 1][k], sp.degrees[k]).T for k in range(3)} for l in range(3)}
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, Coeff)
    rng = np.random.default_rng(4)
    # points concentrated where levels meet (the truncation / tile code paths)
    prm = (0.5 + (rng.random((120, 3)) - 0.5) * np.array([0.7, 0.7, 0.7])).astype(np.float32)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    PHI = core.THB_basis_fns(prm, supp, ns, P, sp.sh_fns, sp.knotvectors, sp.degrees).cpu().numpy()
    Jm = supp.Jm().cpu().numpy()
    pos = 0
    levels = set()
    for i, x in enumerate(prm):
        lev, c, jm, phi = de.point(x.astype(np.float64))
        n = len(jm)
        levels.add(lev)
        assert np.array_equal(Jm[pos:pos + n], jm)
        assert rel_err(PHI[pos:pos + n], phi).max() <= TOL
        pos += n
    assert len(levels) >= 3


def test_c3_full_size_properties(c3):
    """256^3 grid: partition of unity, linear reproduction (Greville control points), first derivatives of
    the identity map, <A x, g> == <x, A^T g>."""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import configs, engine

    sp, P, cp = c3
    prm = B.grid_parametric_coordinates_device((256, 256, 256))
    plan = engine.Plan(P, prm)
    n_in, nnz, nnz_t = plan.stats()
    assert n_in == 256 ** 3
    ones = torch.ones((P.nCP, 3), device="cuda")
    pou = engine.eval_fused(P, plan, ones)[0]
    assert float((pou - 1).abs().max()) < 1e-5
    grev = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    out = engine.eval_fused(P, plan, grev)[0]
    assert float((out - prm).abs().max()) < 1e-5
    out_c = engine.eval_fused(P, plan, grev, formulation="per_point")[0]
    assert float((out_c - prm).abs().max()) < 1e-5
    ders = engine.eval_fused(P, plan, grev, engine.grad_channels(3))
    eye = torch.eye(3, device="cuda")
    for k in range(3):
        assert float((ders[k] - eye[k]).abs().max()) < 5e-4
    del pou, out_c, ders
    g = torch.randn((prm.shape[0], 3), device="cuda")
    x = torch.randn((P.nCP, 3), device="cuda")
    Ax = engine.eval_fused(P, plan, x)[0]
    ATg = engine.eval_fused_backward(P, plan, g)
    ATg2 = engine.eval_fused_backward(P, plan, g, formulation="per_point")
    assert float((ATg - ATg2).abs().max()) <= 1e-4 * float(ATg2.abs().max())
    lhs = float((Ax.double() * g.double()).sum())
    rhs = float((x.double() * ATg.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), float(Ax.double().norm() * g.double().norm()))


def test_plan_slots_at_full_size_match_the_span_kernel(c3):
    """Batches of >= 4 M points take the vector histogram kernel (four points per thread, proven span table): its slots
    must equal those of thb_active_span (binary search + level walk, an independent kernel) bit for bit -- grid points,
    random points, points on breakpoints and the domain ends, outside points, NaN, and a tail that is not a multiple of 4."""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import engine

    sp, P, cp = c3
    g = torch.Generator(device="cuda").manual_seed(7)
    grid = B.grid_parametric_coordinates_device((256, 256, 256))
    rnd = torch.rand((4_200_003, 3), device="cuda", generator=g)
    kn = torch.from_numpy(np.unique(np.asarray(sp.knotvectors[len(sp.knotvectors) - 1][0], dtype=np.float32))).cuda()
    idx = torch.randint(0, len(kn), (300_000, 3), device="cuda", generator=g)
    onk = kn[idx]
    onk[::3] = torch.nextafter(onk[::3], torch.zeros_like(onk[::3]))
    onk[1::3] = torch.nextafter(onk[1::3], torch.ones_like(onk[1::3]) * 2)
    special = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [1.0, 0.5, 0.0], [-1e-7, 0.5, 0.5], [0.5, 1.0000001, 0.5],
                            [float("nan"), 0.5, 0.5], [0.5, 0.5, float("inf")]], device="cuda")
    for prm in (grid, torch.cat([rnd, onk, special]).contiguous()):
        plan = engine.Plan(P, prm)
        ref, _ = engine.active_span(P, prm)
        assert torch.equal(plan.slot[: prm.shape[0]], ref)
        n_in = int((ref >= 0).sum())
        assert int(plan.counters[2]) == n_in
        cnt = torch.bincount(ref[ref >= 0].long(), minlength=P.nslots)
        assert torch.equal(plan.cell_count[: P.nslots].long(), cnt)
        del plan


def test_sharded_fit_reduces_loss(c3):
    from thb_diff_b200 import configs
    from thb_diff_b200.distributed import ShardedFit

    sp, P, cp = c3
    rng = np.random.default_rng(0)
    prm = torch.from_numpy(rng.random((200000, 3)).astype(np.float32)).cuda()
    tgt = torch.stack([prm[:, 0], prm[:, 1], prm[:, 2] + 0.1 * torch.sin(6.2831853 * prm[:, 0]) * torch.sin(6.2831853 * prm[:, 1])], dim=1)
    x = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=1e-2, fused=True)
    fit = ShardedFit(P, prm, tgt, prm.shape[0])
    losses = [float(fit.step(x, opt)) for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0]


def test_adaptive_fit_example_improves_with_refinement():
    """examples/adaptive_fit.py: fit -> refine -> prolong control points (ControlPoints.update_CP) -> refit."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "adaptive_fit.py")
    spec = importlib.util.spec_from_file_location("adaptive_fit", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    losses = mod.main(steps=150, n=50_000, verbose=False)
    assert losses[1] < losses[0] and losses[2] < losses[1]
