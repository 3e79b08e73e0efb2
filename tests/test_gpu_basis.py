"""THB_basis_fns on the GPU vs golden PHI of the reference's float64 loop (compute_THB_fns_tp).
Tolerance: |a-b| <= 1e-5 * max(|b|, scale) with scale = 1 (PHI sums to 1 per point) for values and
scale = max|ref| of the point set for derivatives -- BASELINE.json north_star, SURVEY hard part 6."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, SPACES, load_golden, oracle_space_from_golden, product_space_from_golden, rel_err
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
IDENTITY_OK = [s for s in SPACES if s != "twoblock2d"]  # twoblock2d has active finest-level functions


def _run(h, params, return_mode=[True, False], onehot=True):
    from thb_diff_b200 import core
    from thb_diff_b200.packed import PackedHierarchy

    P = PackedHierarchy.from_space(h, device="cuda", onehot_direct=onehot)
    supp, ns = core.compute_active_span(params, h.knotvectors, h.cells, h.degrees, P)
    return supp, ns, core.THB_basis_fns(params, supp, ns, P, h.sh_fns, h.knotvectors, h.degrees, return_mode)


@pytest.mark.parametrize("onehot", [True, False])
@pytest.mark.parametrize("name", IDENTITY_OK)
def test_phi_matches_reference_loop(name, onehot):
    g = load_golden(name)
    h = product_space_from_golden(g)
    supp, ns, PHI = _run(h, g["params"], onehot=onehot)
    assert PHI.dtype == torch.float32 and PHI.shape == (len(g["PHI"]),)
    assert rel_err(PHI.cpu().numpy(), g["PHI"]).max() <= TOL
    # partition of unity
    pou = np.add.reduceat(PHI.cpu().numpy().astype(np.float64), np.concatenate([[0], np.cumsum(g["num_supp"])[:-1]]))
    assert np.abs(pou - 1).max() < 1e-5


def test_reference_api_sequence_with_setup_objects():
    """README call order with this package's set-up objects."""
    from thb_diff_b200 import core

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    fc = core.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, ac)
    PHI = core.THB_basis_fns(g["params"], supp, ns, fc, h.sh_fns, h.knotvectors, h.degrees)
    assert rel_err(PHI.cpu().numpy(), g["PHI"]).max() <= TOL


def test_twoblock_identity_semantics_against_direct_oracle():
    """Space with active finest-level functions: the packed path uses the identity there; compare with the
    oracle's table-free evaluator (same reading)."""
    g = load_golden("twoblock2d")
    h = product_space_from_golden(g)
    supp, ns, PHI = _run(h, g["params"])
    sp = oracle_space_from_golden(g)
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, sp.Coeff)
    ref = np.concatenate([de.point(x)[3] for x in g["params"]])
    assert rel_err(PHI.cpu().numpy(), ref).max() <= TOL


@pytest.mark.parametrize("name", ["block2d", "graded2d", "block3d"])
def test_derivative_modes(name):
    g = load_golden(name)
    h = product_space_from_golden(g)
    sp = oracle_space_from_golden(g)
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, sp.Coeff)
    ref_g = np.concatenate([de.point(x, want_grad=True)[4] for x in g["params"]])
    supp, ns, (PHI, dPHI) = _run(h, g["params"], return_mode=[True, "grad"])
    assert dPHI.shape == ref_g.shape
    scale = np.abs(ref_g).max()
    assert rel_err(dPHI.cpu().numpy(), ref_g, scale=scale).max() <= TOL
    # sum over supports of each gradient component vanishes (derivative of the partition of unity)
    starts = np.concatenate([[0], np.cumsum(g["num_supp"])[:-1]])
    s = np.add.reduceat(dPHI.cpu().numpy().astype(np.float64), starts, axis=0)
    assert np.abs(s).max() < 1e-4 * scale
    # the reference's "derivative" mode = mixed partial
    oa = O.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = O.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees, max_level="identity")
    osupp, ons = O.compute_active_span(g["params"], sp.knotvectors, sp.cells, sp.degrees, oa)
    ref_m = O.thb_basis_fns(g["params"], osupp, ons, fc, sp.sh_fns, sp.knotvectors, sp.degrees, mode="mixed")
    _, _, mixed = _run(h, g["params"], return_mode=[False, True])
    assert rel_err(mixed.cpu().numpy(), ref_m, scale=np.abs(ref_m).max()).max() <= TOL


def test_appendix_a3_derivative_kat():
    """SURVEY Appendix A.3, point (0.25, 0.35) of the block2d space."""
    g = load_golden("block2d")
    h = product_space_from_golden(g)
    prm = np.array([[0.25, 0.35]], dtype=np.float32)
    _, _, (PHI, dPHI) = _run(h, prm, return_mode=[True, "grad"])
    du = [11, 0.8333333333, 0, 0, -0.01953125, -1.02734375, -3.25, -0.703125, 0, -0.55, -6.215625, -2.109375, 0.01953125, 0.8440104167, 1.178125]
    dv = [3, 1.25, 0, 0, -0.029296875, -0.416015625, 0.09375, 0.3515625, -0.17578125, -2.64609375, -1.4203125, 1.0546875, -0.029296875, -0.466015625, -0.5671875]
    got = dPHI.cpu().numpy()
    assert np.abs(got[:, 0] - du).max() < 2e-4 and np.abs(got[:, 1] - dv).max() < 2e-4


@pytest.mark.parametrize("name,kw", [("block2d", {}), ("twoblock2d", {}), ("block3d", {}), ("block2d", {"quantize_fp16": True})])
def test_legacy_dense_tables(name, kw):
    """THB_basis_fns fed with the reference's dense dict tables (incl. its 'ones' finest level and fp16
    quantisation): must reproduce the golden PHI of the reference loop, whatever the tables say."""
    from thb_diff_b200 import core

    g = load_golden(name) if not kw else dict(np.load(f"{GOLDEN}/space_block2d_fp16.npz"))
    sp = oracle_space_from_golden(g)
    h = product_space_from_golden(g)
    oa = O.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = O.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees, **kw)  # reference-faithful tables
    # (a) fully legacy inputs: python lists of tuples
    osupp, ons = O.compute_active_span(g["params"], sp.knotvectors, sp.cells, sp.degrees, oa)
    PHI = core.THB_basis_fns(g["params"], osupp, ons, fc, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert rel_err(PHI.cpu().numpy(), g["PHI"]).max() <= TOL
    # (b) GPU span lookup + dense tables
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, ac)
    PHI2 = core.THB_basis_fns(g["params"], supp, ns, fc, h.sh_fns, h.knotvectors, h.degrees)
    assert rel_err(PHI2.cpu().numpy(), g["PHI"]).max() <= TOL


def test_dense_materialisation_of_refinement_operators():
    from thb_diff_b200 import core

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    fc = core.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    for lev in (0, 1):
        assert np.abs(fc[lev] - g[f"fc_{lev}"]).max() < 1e-6


def test_unsupported_degree_is_an_error():
    from thb_diff_b200 import core, multilevel_spline_space as M
    from thb_diff_b200.packed import PackedHierarchy

    U = np.array([0, 0, 0, 0, 0, 0.5, 1, 1, 1, 1, 1.0])  # degree 4
    sp = M.Space(M.TensorProduct([M.BSpline(U, 4), M.BSpline(U, 4)]), 2)
    sp.build_hierarchy_from_domain_sequence()
    P = PackedHierarchy.from_space(sp, device="cuda")
    prm = np.array([[0.3, 0.3]], dtype=np.float32)
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    with pytest.raises(RuntimeError):
        core.THB_basis_fns(prm, supp, ns, P, sp.sh_fns, sp.knotvectors, sp.degrees)
