"""The plain-C oracle vs the NumPy oracle / golden fixtures (which are pinned to the reference), and the
host-side packed-table builder (thb_diff_b200.packed) through it -- all on CPU."""
import numpy as np
import pytest

from helpers import GOLDEN, SPACES, emulate_packed_point, host_tables_from_packed, load_golden, product_space_from_golden
from oracle import c_oracle
from oracle import thb_oracle as O

IDENTITY_OK = [s for s in SPACES if s != "twoblock2d"]


def test_c_spans_and_basis():
    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    assert c_oracle.find_spans(k["span_u"], k["U_B"], 3).tolist() == [-1, 3, 3, 4, 4, 5, 9, 9, 9, 10]
    N, dN = c_oracle.basis_funs(5, 0.45, 3, k["U_B"])
    assert np.allclose(N, k["N_045"], atol=1e-15) and np.allclose(dN, k["dN_045"], atol=1e-12)


def test_c_contraction_matches_compiled_reference():
    c = dict(np.load(f"{GOLDEN}/contraction.npz"))
    for tag in "abc":
        fwd = c_oracle.eval_forward(c[f"{tag}_cp"], c[f"{tag}_Jm"], c[f"{tag}_phi"], c[f"{tag}_cumsum"])
        bwd = c_oracle.eval_backward(c[f"{tag}_grad_out"], c[f"{tag}_cp"].shape, c[f"{tag}_Jm"], c[f"{tag}_phi"], c[f"{tag}_cumsum"])
        assert np.array_equal(fwd, c[f"{tag}_fwd"])   # same float accumulation order as compute_pts.cpp
        assert np.array_equal(bwd, c[f"{tag}_bwd"])


@pytest.mark.parametrize("onehot", [True, False])
@pytest.mark.parametrize("name", IDENTITY_OK)
def test_packed_tables_reproduce_reference_phi(name, onehot):
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cpu", onehot_direct=onehot)
    ht = host_tables_from_packed(P)
    off = np.concatenate([[0], np.cumsum(g["num_supp"])[:-1]])
    cp = np.random.default_rng(0).standard_normal((P.nCP, 3)).astype(np.float32)
    r = ht.eval(g["params"], cp, want_phi=True, offsets=off, nnz=len(g["PHI"]))
    info = P.cellinfo_host
    assert np.array_equal(info[r["slot"], 0], g["span_level"])
    assert np.array_equal(info[r["slot"], 1:1 + P.ndim], g["span_cell"])
    assert np.array_equal(info[r["slot"], 5], g["num_supp"])
    assert np.abs(r["phi"] - g["PHI"]).max() < 5e-7          # tables and knots are float32
    ref = O.eval_forward(cp, g["Jm"], g["PHI"], np.cumsum(g["num_supp"]))
    assert np.abs(r["out"] - ref).max() < 5e-6
    # the python emulation used elsewhere agrees too
    e = emulate_packed_point(P, g["params"][3])
    assert np.abs(e[3] - g["PHI"][off[3]:off[3] + g["num_supp"][3]]).max() < 5e-7


def test_packed_builder_rejects_bad_input():
    from thb_diff_b200 import multilevel_spline_space as M
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    bad = {l: h.fns[l].copy() for l in h.fns}
    bad[1][0, 0] = 1  # an active level-1 function over an active level-0 cell
    with pytest.raises(ValueError):
        PackedHierarchy.build(h.knotvectors, h.degrees, h.cells, bad, h.Coeff, device="cpu")
    U = np.array([0, 0, 0, 0.5, 0.5, 1, 1, 1.0])  # repeated interior knot
    with pytest.raises(ValueError):  # interior multiplicities are not supported (refinement would drop them)
        sp = M.Space(M.TensorProduct([M.BSpline(U, 2), M.BSpline(U, 2)]), 2)
        sp.build_hierarchy_from_domain_sequence()
        PackedHierarchy.from_space(sp, device="cpu")


def test_bench_space_partition_of_unity_and_linear_reproduction():
    """The C3 space generator (graded nested boxes): PoU and Greville reproduction on random points."""
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=4)  # BASELINE config 3 space
    P = PackedHierarchy.from_space(sp, device="cpu")
    ht = host_tables_from_packed(P)
    prm = np.random.default_rng(1).random((20000, 3)).astype(np.float32)
    cp = configs.greville_ctrl_pts(sp, bump=0.0)
    ones = np.ones((P.nCP, 1), dtype=np.float32)
    r = ht.eval(prm, cp)
    assert r["inside"] == len(prm)
    assert np.abs(r["out"] - prm).max() < 2e-6
    assert np.abs(ht.eval(prm, ones)["out"] - 1).max() < 2e-6
    levels = P.cellinfo_host[r["slot"], 0]
    assert set(levels.tolist()) == {0, 1, 2, 3}


def test_recursive_truncation_option():
    """truncation='recursive' (proper THB) == the reference's single-level truncation on graded spaces, and
    restores partition of unity / linear reproduction where single-level truncation loses them."""
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    rng = np.random.default_rng(0)
    for name, graded in (("graded2d", True), ("block3d", True), ("twoblock2d", False)):
        g = load_golden(name)
        h = product_space_from_golden(g)
        prm = rng.random((3000, h.ndim)).astype(np.float32)
        res = {}
        for tr in ("single", "recursive"):
            P = PackedHierarchy.from_space(h, device="cpu", truncation=tr)
            ht = host_tables_from_packed(P)
            pou = np.abs(ht.eval(prm, np.ones((P.nCP, 3), dtype=np.float32))["out"][:, 0] - 1).max()
            lin = np.abs(ht.eval(prm, configs.greville_ctrl_pts(h, bump=0.0))["out"][:, :h.ndim] - prm).max()
            res[tr] = (P.tiles.numpy(), pou, lin)
        assert res["recursive"][1] < 1e-6 and res["recursive"][2] < 1e-6
        if graded:
            assert np.array_equal(res["single"][0], res["recursive"][0])
        else:
            assert res["single"][1] > 0.3   # the reference's behaviour on this non-graded space
    with pytest.raises(ValueError):
        PackedHierarchy.from_space(h, device="cpu", truncation="both")


@pytest.mark.parametrize("name", IDENTITY_OK)
def test_direct_c_evaluator_reproduces_reference_goldens(name):
    """oracle_direct_eval (table-free, C) against the outputs of the reference itself: levels, cells, Jm bit-exact, PHI."""
    from helpers import direct_tables

    g = load_golden(name)
    sp = oracle_space_from_golden_(g)
    dt = direct_tables(sp)
    dt_exact = c_oracle.DirectTables(sp.knotvectors, sp.degrees, sp.cells, sp.fns, dt_coeff(sp), knots_as_f32=False)
    r = dt_exact.eval(g["params"].astype(np.float64), max_supp=int(g["num_supp"].max()))
    assert np.array_equal(r["lev"], g["span_level"]) and np.array_equal(r["cell"][:, :sp.ndim], g["span_cell"])
    assert np.array_equal(r["nsupp"], g["num_supp"])
    valid = np.arange(r["jm"].shape[1])[None, :] < r["nsupp"][:, None]
    assert np.array_equal(r["jm"][valid], g["Jm"])
    assert np.abs(r["phi"][valid] - g["PHI"]).max() < 1e-13
    # f32-rounded knots (what the product sees) move PHI by f32 rounding only
    r32 = dt.eval(g["params"].astype(np.float64), max_supp=int(g["num_supp"].max()))
    assert np.array_equal(r32["jm"][valid], g["Jm"]) and np.abs(r32["phi"][valid] - g["PHI"]).max() < 5e-6


def oracle_space_from_golden_(g):
    from helpers import oracle_space_from_golden

    return oracle_space_from_golden(g)


def dt_coeff(sp):
    L = max(sp.knotvectors.keys())
    return {l: {k: O.knot_insertion_matrix(sp.knotvectors[l][k], sp.knotvectors[l + 1][k], sp.degrees[k]).T for k in range(sp.ndim)}
            for l in range(L)}


def test_direct_c_evaluator_equals_the_python_one_with_gradients():
    g = load_golden("block3d")
    sp = oracle_space_from_golden_(g)
    Coeff = dt_coeff(sp)
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, Coeff)
    dt = c_oracle.DirectTables(sp.knotvectors, sp.degrees, sp.cells, sp.fns, Coeff, knots_as_f32=False)
    x = np.random.default_rng(7).random((40, 3))
    r = dt.eval(x, max_supp=256, want_grad=True)
    for i in range(len(x)):
        lev, c, jm, phi, dphi = de.point(x[i], want_grad=True)
        n = len(jm)
        assert r["lev"][i] == lev and tuple(r["cell"][i]) == tuple(c) and r["nsupp"][i] == n
        assert np.array_equal(r["jm"][i, :n], jm)
        assert np.abs(r["phi"][i, :n] - phi).max() < 1e-14 and np.abs(r["dphi"][i, :n] - dphi).max() < 1e-11


@pytest.mark.parametrize("levels", [4, 5])
def test_packed_tables_of_the_bench_spaces_against_the_table_free_evaluator(levels):
    """BASELINE config 3 (4 levels) and config 5 (5 levels): the product's packed tables (host builder, evaluated by the C
    oracle) against the definition of the truncated functions at 2e4 points concentrated on the level interfaces --
    cells / support ids bit-exact, PHI and the three gradient channels within 1e-5."""
    from helpers import compare_with_direct, direct_tables, interface_points
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=levels)
    P = PackedHierarchy.from_space(sp, device="cpu")
    ht = host_tables_from_packed(P)
    prm = interface_points(20000, 3, seed=levels)
    r0 = ht.eval(prm, None)
    info = P.cellinfo_host[r0["slot"]]
    S = info[:, 5].astype(np.int64)
    off = np.cumsum(S) - S
    nnz = int(S.sum())
    phi = ht.eval(prm, None, want_phi=True, offsets=off, nnz=nnz)["phi"]
    dphi = [ht.eval(prm, None, mode=1 << k, want_phi=True, offsets=off, nnz=nnz)["phi"] for k in range(3)]
    Jm = np.concatenate([P.supp_jm.numpy()[info[i, 4]: info[i, 4] + info[i, 5]] for i in range(len(prm))]).astype(np.int64)
    r = compare_with_direct(direct_tables(sp), prm, info, Jm, phi, off, 1e-5, dPHI=dphi)
    assert set(r["lev"].tolist()) == set(range(levels))
