"""BASELINE.json configs 1 and 2 at their full sizes (SURVEY 8d expected shapes) through the reference-named API."""
import numpy as np
import pytest
import torch

from helpers import host_tables_from_packed, rel_err
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_space(sp):
    """oracle Space with the same knots / refinement (built independently of the product's masks)"""
    o = O.Space([sp.knotvectors[0][k] for k in range(sp.ndim)], sp.degrees, sp.num_levels)
    o.build_hierarchy_from_domain_sequence()
    return o


def test_config1_readme_2d_full():
    """C1: degrees (2,3), 3 levels, level-0 cells (2,2),(3,1) refined, 100x100 grid: N=1e4, S=12, nnz=120000, nCP=1796;
    compared with the oracle's reference-faithful dense-table path on every point."""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import configs, core
    from thb_diff_b200.torch_funcs import THB_nn_module, prepare_data_for_CUDA_evaluation

    sp = configs.config1_space()
    params = B.generate_parametric_coordinates((100, 100)).astype(np.float32)
    ac = core.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = core.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees)
    supp, ns = core.compute_active_span(params, sp.knotvectors, sp.cells, sp.degrees, ac)
    PHI = core.THB_basis_fns(params, supp, ns, fc, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert len(ns) == 10_000 and set(np.asarray(ns).tolist()) == {12} and PHI.shape == (120_000,) and supp.P.nCP == 1796
    assert {l: int(sp.fns[l].sum()) for l in sp.fns} == {0: 120, 1: 0, 2: 0}
    assert {l: int(sp.cells[l].sum()) for l in sp.cells} == {0: 68, 1: 8, 2: 0}

    o = _oracle_space(sp)
    o._refine_cell((2, 2), 0)
    o._refine_cell((3, 1), 0)
    o.build_hierarchy_from_domain_sequence()
    oac = O.compute_active_cells_active_supp(o.cells, o.fns, o.degrees)
    ofc = O.compute_refinement_operators(o.fns, o.Coeff, o.degrees)
    osupp, ons, ospans = O.compute_active_span(params.astype(np.float64), o.knotvectors, o.cells, o.degrees, oac, return_spans=True)
    assert np.array_equal(np.asarray(ns), np.array(ons))
    assert supp.active_cells() == ospans                                     # bit-exact (level, cell)
    Jm_ref = O.flat_support_indices(osupp, o.sh_fns)
    assert np.array_equal(supp.Jm().cpu().numpy(), Jm_ref)
    PHI_ref = O.thb_basis_fns(params.astype(np.float64), osupp, ons, ofc, o.sh_fns, o.knotvectors, o.degrees)
    assert rel_err(PHI.cpu().numpy(), PHI_ref).max() <= TOL
    # evaluation + gradient through the reference-named torch layer
    cp = configs.greville_ctrl_pts(sp)
    cpt, Jm, phi, cumsum, dev = prepare_data_for_CUDA_evaluation(PHI, supp, ns, cp, None, sp.sh_fns, "cuda")
    x = cpt.clone().requires_grad_(True)
    out = THB_nn_module(Jm, phi, cumsum, "cuda")(x)
    ref = O.eval_forward(cp, Jm_ref, PHI_ref, np.cumsum(ons))
    assert rel_err(out.detach().cpu().numpy(), ref, scale=np.abs(ref).max()).max() <= TOL
    w = torch.randn_like(out)
    (out * w).sum().backward()
    gref = O.eval_backward(w.cpu().numpy(), cp.shape, Jm_ref, PHI_ref, np.cumsum(ons))
    assert rel_err(x.grad.cpu().numpy(), gref, scale=np.abs(gref).max()).max() <= TOL


def test_config2_readme_3d_full():
    """C2: degrees (2,3,2), 3 levels, cells (2,2,2),(3,1,0) refined, 50^3 grid: N=125000, S=36, nnz=4.5M, nCP=33972."""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import configs, core, engine

    sp = configs.config2_space()
    params = B.generate_parametric_coordinates((50, 50, 50)).astype(np.float32)
    ac = core.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = core.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees)
    supp, ns = core.compute_active_span(params, sp.knotvectors, sp.cells, sp.degrees, ac)
    PHI = core.THB_basis_fns(params, supp, ns, fc, sp.sh_fns, sp.knotvectors, sp.degrees)
    P = supp.P
    assert len(ns) == 125_000 and set(np.asarray(ns).tolist()) == {36} and PHI.shape == (4_500_000,) and P.nCP == 33_972
    assert sp.sh_fns == {0: (12, 10, 7), 1: (22, 17, 12), 2: (42, 31, 22)}
    # every point against the float64 C oracle
    ht = host_tables_from_packed(P)
    S = np.asarray(ns).astype(np.int64)
    off = np.cumsum(S) - S
    cp = configs.greville_ctrl_pts(sp)
    r = ht.eval(params, cp, want_phi=True, offsets=off, nnz=int(S.sum()))
    assert np.array_equal(supp.slot.cpu().numpy(), r["slot"])
    assert rel_err(PHI.cpu().numpy(), r["phi"]).max() <= TOL
    out = engine.eval_fused(P, supp.plan, torch.from_numpy(cp).cuda())[0].cpu().numpy()
    assert rel_err(out, r["out"], scale=np.abs(r["out"]).max()).max() <= TOL
    assert np.abs(out - params).max() < 0.11  # Greville net + 0.1 bump on z: the map is the identity up to the bump
    # sampled points against the table-free oracle built from an independently constructed space
    o = _oracle_space(sp)
    o._refine_cell((2, 2, 2), 0)
    o._refine_cell((3, 1, 0), 0)
    o.build_hierarchy_from_domain_sequence()
    de = O.DirectEvaluator(o.knotvectors, o.degrees, o.cells, o.fns, o.Coeff)
    Jm = supp.Jm().cpu().numpy()
    phi = PHI.cpu().numpy()
    for i in np.random.default_rng(0).choice(len(params), 60, replace=False):
        lev, c, jm, ph = de.point(params[i].astype(np.float64))
        assert np.array_equal(Jm[off[i]:off[i] + 36], jm)
        assert rel_err(phi[off[i]:off[i] + 36], ph).max() <= TOL
