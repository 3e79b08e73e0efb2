"""Every tensor-product shape the library has kernels for (p in {1,2,3} per dimension as instantiated in
csrc/thb_tile.cu), on a small three-level hierarchy with non-uniform knots: cell lookup bit-exact, PHI, the fused
evaluation in all three formulations, first-derivative channels and both fused backward formulations against the
float64 C oracle (which follows compute_THB_fns_tp, THB/core.py:232-268, on the same packed tables), plus the two
properties that pin the TABLES themselves for shapes without a reference golden: partition of unity and Greville
linear reproduction."""
import numpy as np
import pytest
import torch

from helpers import host_tables_from_packed, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
SHAPES = [(1, 1), (2, 2), (3, 3), (2, 3), (3, 2), (1, 1, 1), (2, 2, 2), (3, 3, 3), (2, 3, 2), (2, 2, 3), (2, 3, 3), (3, 2, 2),
          (3, 2, 3), (3, 3, 2)]


def _space(degrees, seed):
    from thb_diff_b200 import configs
    from thb_diff_b200.multilevel_spline_space import BSpline, Space, TensorProduct

    rng = np.random.default_rng(seed)
    d = len(degrees)
    nc = 8 if d == 3 else 12
    sp = Space(TensorProduct([BSpline(configs.uniformish_knots(nc, p, rng), p) for p in degrees]), 3)
    sp.build_hierarchy_from_domain_sequence()
    lo0, hi0 = nc // 4, nc - nc // 4                      # graded nested boxes with >= p cells of margin
    sp.refine_box((lo0,) * d, (hi0,) * d, 0)
    a, b = 2 * lo0 + 3, 2 * hi0 - 3
    sp.refine_box((a,) * d, (b,) * d, 1)
    sp.build_hierarchy_from_domain_sequence()
    return sp


@pytest.mark.parametrize("degrees", SHAPES, ids=lambda s: "p" + "".join(map(str, s)))
def test_shape_against_c_oracle(degrees):
    from thb_diff_b200 import configs, core, engine
    from thb_diff_b200.packed import PackedHierarchy
    from oracle import thb_oracle as O

    d = len(degrees)
    sp = _space(degrees, seed=sum(degrees) * 7 + d)
    P = PackedHierarchy.from_space(sp, device="cuda")
    rng = np.random.default_rng(5)
    prm = rng.random((6000, d)).astype(np.float32)
    prm[:400] = prm[:400] * 0.02 + 0.49                  # a crowd in a few finest cells: multi-sub-batch items
    prm[400] = 0.0
    prm[401] = 1.0                                       # domain corners
    cp = rng.standard_normal((P.nCP, 3)).astype(np.float32)
    cpt = torch.from_numpy(cp).cuda()
    ht = host_tables_from_packed(P)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    S = np.asarray(ns).astype(np.int64)
    off = np.cumsum(S) - S
    ref = ht.eval(prm, cp, want_phi=True, offsets=off, nnz=int(S.sum()))
    assert np.array_equal(supp.slot.cpu().numpy(), ref["slot"]) and (ref["slot"] >= 0).all()
    PHI = core.THB_basis_fns(prm, supp, ns, P, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert rel_err(PHI.cpu().numpy(), ref["phi"]).max() <= TOL
    scale = np.abs(ref["out"]).max()
    for form in ("local_net", "per_point", "tile_net"):
        out = engine.eval_fused(P, supp.plan, cpt, formulation=form)[0].cpu().numpy()
        assert rel_err(out, ref["out"], scale=scale).max() <= TOL, form
    # first derivatives (true gradient components)
    Jm = supp.Jm().cpu().numpy()
    cs = np.cumsum(S)
    for form in ("local_net", "per_point"):
        ders = engine.eval_fused(P, supp.plan, cpt, engine.grad_channels(d), formulation=form)
        for k in range(d):
            r = ht.eval(prm, cp, mode=1 << k, want_phi=True, offsets=off, nnz=int(S.sum()))
            cond = O.eval_forward(np.abs(cp), Jm, np.abs(r["phi"]), cs)
            assert (np.abs(ders[k].cpu().numpy() - r["out"]) <= TOL * np.maximum(cond, 1.0)).all(), (form, k)
    # backward: A^T g with A from the oracle's PHI
    g = rng.standard_normal((len(prm), 3)).astype(np.float32)
    refg = O.eval_backward(g, cp.shape, Jm, ref["phi"], cs)
    for form in ("local_net", "per_point"):
        got = engine.eval_fused_backward(P, supp.plan, torch.from_numpy(g).cuda(), formulation=form).cpu().numpy()
        assert rel_err(got, refg, scale=np.abs(refg).max()).max() <= TOL, form
    # the tables themselves: partition of unity and linear reproduction with Greville control points
    ones = torch.ones((P.nCP, 3), device="cuda")
    assert float((engine.eval_fused(P, supp.plan, ones)[0] - 1).abs().max()) <= TOL
    grev = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    rep = engine.eval_fused(P, supp.plan, grev)[0][:, :d].cpu().numpy()
    assert np.abs(rep - prm).max() <= TOL


@pytest.mark.parametrize("degrees", [(2, 3), (3, 3), (3, 2, 3)], ids=lambda s: "p" + "".join(map(str, s)))
def test_large_batches_take_the_vector_histogram_kernel(degrees):
    """From 4 M points up thb_plan_build runs k_plan_count_vec (2-D and 3-D instantiations, four points per thread): slots
    and histogram bit-identical to thb_active_span on a non-uniform three-level hierarchy, incl. breakpoints, the domain
    ends, outside points, NaN and a tail that is not a multiple of four."""
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    d = len(degrees)
    sp = _space(degrees, seed=3)
    P = PackedHierarchy.from_space(sp, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = torch.rand((4_100_001, d), device="cuda", generator=g)
    kn = [torch.from_numpy(np.unique(np.asarray(sp.knotvectors[len(sp.knotvectors) - 1][k], dtype=np.float32))).cuda() for k in range(d)]
    onk = torch.stack([kn[k][torch.randint(0, len(kn[k]), (60_000,), device="cuda", generator=g)] for k in range(d)], dim=1)
    onk[::3] = torch.nextafter(onk[::3], torch.zeros_like(onk[::3]))
    onk[1::3] = torch.nextafter(onk[1::3], torch.full_like(onk[1::3], 2.0))
    special = torch.tensor([[0.0] * d, [1.0] * d, [-1e-7] + [0.5] * (d - 1), [0.5] * (d - 1) + [1.0000001],
                            [float("nan")] + [0.5] * (d - 1), [0.5] * (d - 1) + [float("inf")]], device="cuda")
    prm = torch.cat([rnd, onk, special]).contiguous()
    plan = engine.Plan(P, prm)
    ref, _ = engine.active_span(P, prm)
    assert torch.equal(plan.slot[: prm.shape[0]], ref)
    assert torch.equal(plan.cell_count[: P.nslots].long(), torch.bincount(ref[ref >= 0].long(), minlength=P.nslots))
    assert int(plan.counters[2]) == int((ref >= 0).sum())
