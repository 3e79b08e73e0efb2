"""Span search, 1-D bases and compute_active_span on the GPU: integer results bit-exact."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, SPACES, load_golden, product_space_from_golden
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu


def test_span_table_q10_bit_exact():
    from thb_diff_b200 import bspline_funcs as B

    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    got = B.find_span_array_jax(k["span_u"], k["U_B"], 3).cpu().numpy()
    assert got.tolist() == [-1, 3, 3, 4, 4, 5, 9, 9, 9, 10]
    rng = np.random.default_rng(5)
    for U, p in ((k["U_A"], 2), (k["U_B"], 3)):
        u = np.concatenate([rng.random(20000), U, U - 1e-7, U + 1e-7, [np.nan]]).astype(np.float32)
        ref = O.find_span_array(u, U.astype(np.float32), p)
        ref[-1] = -1  # NaN: defined as out of domain here
        got = B.find_span_array_jax(u, U, p).cpu().numpy()
        assert np.array_equal(got, ref)


def test_basis_1d_values_and_derivatives():
    from thb_diff_b200 import bspline_funcs as B

    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    N = B.basis_fns_vmap(np.array([0.45]), k["U_B"], 3).cpu().numpy()
    assert N.shape == (1, 4)  # never squeezed
    assert np.allclose(N[0], [0.0083333333, 0.371875, 0.6041666667, 0.015625], atol=2e-6)
    dN = B.der_basis_fns_vmap(np.array([0.45]), k["U_B"], 3).cpu().numpy()
    assert np.allclose(dN[0], [-0.5, -5.4375, 5.0, 0.9375], atol=5e-5)
    for tag, p in (("A", 2), ("B", 3)):
        pts, U = k[f"pts_{tag}"], k[f"U_{tag}"]
        N = B.basis_fns_vmap(pts, U, p).cpu().numpy()
        dN = B.der_basis_fns_vmap(pts, U, p).cpu().numpy()
        assert np.abs(N - k[f"N_{tag}"]).max() < 1e-5
        assert np.abs(dN - k[f"dN_{tag}"]).max() < 1e-5 * np.abs(k[f"dN_{tag}"]).max()
        assert np.abs(N.sum(1) - 1).max() < 1e-5 and np.abs(dN.sum(1)).max() < 1e-4


@pytest.mark.parametrize("name", SPACES)
def test_compute_active_span_bit_exact(name):
    from thb_diff_b200 import core

    g = load_golden(name)
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, ac)
    assert np.array_equal(np.asarray(ns), g["num_supp"])
    cells = supp.active_cells()
    assert [c[0] for c in cells] == g["span_level"].tolist()
    assert [list(c[1]) for c in cells] == g["span_cell"].tolist()
    assert np.array_equal(supp.Jm().cpu().numpy(), g["Jm"])
    assert np.array_equal(supp.Jm(torch.int32).cpu().numpy(), g["Jm"])
    # list view: same tuples as the reference dict
    oa = O.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    for i in (0, len(supp) // 2, len(supp) - 1):
        lev, c = cells[i]
        assert supp[i] == [(fl, tuple(fn)) for fl, fn in oa[lev][c]]
    # plan is a permutation grouped by cell (a light plan holds the permutation only until its records are asked for)
    if supp.plan.light:
        perm_light = supp.plan.scratch[: supp.plan.n].cpu().numpy().copy()
    supp.plan.ensure_records()
    rec = supp.plan.sorted_params[: supp.plan.n].cpu().numpy()
    if supp.plan.light:
        assert np.array_equal(perm_light, np.ascontiguousarray(rec[:, 3]).view(np.int32))
    perm = np.ascontiguousarray(rec[:, 3]).view(np.int32)
    assert np.array_equal(np.sort(perm), np.arange(len(perm)))
    slot_sorted = supp.slot.cpu().numpy()[perm]
    assert np.all(np.diff(slot_sorted) >= 0)
    assert np.array_equal(rec[:, : h.ndim], np.asarray(g["params"], dtype=np.float32)[perm])


def test_plan_build_in_two_phases_equals_the_whole(tmp_path):
    """thb_plan_build_phase: histogram, then scan + work items, then scatter = thb_plan_build; every overlap mode of
    plan_and_nets (the nets beside different parts of the plan, on a second stream) evaluates to the same bits."""
    from thb_diff_b200 import configs, engine
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=3)
    P = PackedHierarchy.from_space(sp, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    prm = torch.rand((200_000, 3), device="cuda", generator=g)
    whole = engine.Plan(P, prm).canonicalize()
    split = engine.Plan(P, prm)
    split.sorted_params.zero_(); split.items.zero_(); split.cell_count.zero_(); split.slot.fill_(-7)
    split.build(phase=engine.PLAN_HISTOGRAM)
    assert torch.equal(split.slot, whole.slot) and torch.equal(split.cell_count[: P.nslots], whole.cell_count[: P.nslots])
    split.build(phase=engine.PLAN_ITEMS)
    assert torch.equal(split.cell_start[: P.nslots + 1], whole.cell_start[: P.nslots + 1])
    nit = int(whole.counters[0])
    assert int(split.counters[0]) == nit and torch.equal(split.items[:nit], whole.items[:nit])
    split.build(phase=engine.PLAN_SCATTER)
    split.canonicalize()
    n_in = int(whole.counters[2])
    assert int(split.counters[2]) == n_in and torch.equal(split.sorted_params[:n_in], whole.sorted_params[:n_in])
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
    outs = []
    for mode in (False, True, "scatter", "after_histogram", "auto"):
        nets = engine.plan_and_nets(P, split, cp, overlap=mode)
        outs.append(engine.eval_nets(P, split, nets, 3, with_derivatives=False)[0].clone())
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    with pytest.raises(RuntimeError):
        split.build(phase=8)


def test_out_of_domain_points_are_flagged():
    from thb_diff_b200 import core

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    prm = np.array([[0.5, 0.5], [-0.1, 0.5], [0.5, 1.5], [1.0, 1.0], [np.nan, 0.2]], dtype=np.float32)
    supp, ns = core.compute_active_span(prm, h.knotvectors, h.cells, h.degrees, ac)
    s = supp.slot.cpu().numpy()
    assert s[0] >= 0 and s[3] >= 0          # u == U[-1] belongs to the last cell (span rule n-1)
    assert s[1] == -1 and s[2] == -1 and s[4] == -1
    assert np.asarray(ns).tolist()[1:3] == [0, 0]


def test_legacy_dict_input():
    """compute_active_span fed with the reference-style dict of support lists."""
    from thb_diff_b200 import core

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    d = O.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    supp, ns = core.compute_active_span(g["params"], h.knotvectors, h.cells, h.degrees, d)
    assert np.array_equal(np.asarray(ns), g["num_supp"])
    assert np.array_equal(supp.Jm().cpu().numpy(), g["Jm"])


def test_reference_named_aliases():
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import core

    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    assert B.findSpan(9, 3, 0.45, k["U_B"]) == 5
    assert np.allclose(B.basisFun(5, 0.45, 3, k["U_B"]), k["N_045"], atol=2e-6)
    assert np.allclose(B.basisFun_jax(0.45, k["U_B"], 3).cpu().numpy()[0], k["N_045"], atol=2e-6)
    g = load_golden("block2d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    fc = core.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    spans, ns, supp = core.compute_active_span_v2(g["params"], h.knotvectors, h.cells, h.degrees, ac, fc)
    assert [s[0] for s in spans] == g["span_level"].tolist()
    PHI = core.THB_basis_fns_tp_parallel(g["params"], supp, fc, h.sh_fns, h.knotvectors, h.degrees)
    PHI2, ns2 = core.compute_THB_fns_tp(g["params"], spans, ac, fc, h.sh_fns, h.knotvectors, h.degrees)
    assert np.abs(PHI.cpu().numpy() - g["PHI"]).max() < 1e-5 and torch.equal(PHI, PHI2)
