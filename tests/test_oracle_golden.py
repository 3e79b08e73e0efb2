"""Pins the oracle (oracle/thb_oracle.py) to outputs of the reference's own code (tests/golden)."""
import numpy as np
import pytest

from helpers import GOLDEN, SPACES, load_golden, oracle_space_from_golden
from oracle import thb_oracle as O


def test_span_table_q10():
    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    assert list(k["span_idx"]) == [-1, 3, 3, 4, 4, 5, 9, 9, 9, 10]
    got = O.find_span_array(k["span_u"], k["U_B"], 3)
    assert np.array_equal(got, k["span_idx"])
    got32 = O.find_span_array(k["span_u"].astype(np.float32), k["U_B"].astype(np.float32), 3)
    assert np.array_equal(got32, k["span_idx"])
    n = len(k["U_B"]) - 4
    assert [O.find_span_scalar(n - 1, 3, float(u), k["U_B"]) for u in k["fs_u"]] == list(k["fs_idx"])


def test_basis_kat_and_derivatives():
    k = dict(np.load(f"{GOLDEN}/kat.npz"))
    N, dN = O.basis_funs_and_ders(5, 0.45, 3, k["U_B"])
    assert np.allclose(N, k["N_045"], rtol=0, atol=1e-15)
    assert np.allclose(N, [0.0083333333, 0.371875, 0.6041666667, 0.015625], atol=1e-9)
    assert np.allclose(dN, [-0.5, -5.4375, 5.0, 0.9375], atol=1e-12)
    assert np.allclose(dN, k["dN_045"], atol=1e-12)
    for tag, p in (("A", 2), ("B", 3)):
        U = k[f"U_{tag}"]
        for x, s, Nref, dref in zip(k[f"pts_{tag}"], k[f"span_{tag}"], k[f"N_{tag}"], k[f"dN_{tag}"]):
            N, dN = O.basis_funs_and_ders(int(s), float(x), p, U)
            assert np.allclose(N, Nref, atol=1e-15)
            assert np.allclose(dN, dref, atol=1e-11)
            assert np.allclose(O.basis_all_dense(float(x), U, p), Nref, atol=1e-14)  # basisFun_jax == basisFun in-domain


@pytest.mark.parametrize("name", SPACES)
def test_space_setup_matches_reference(name):
    g = load_golden(name)
    sp = oracle_space_from_golden(g)
    d = sp.ndim
    for lev in range(sp.num_levels):
        assert tuple(g[f"sh_fns_{lev}"]) == sp.sh_fns[lev]
        assert np.array_equal(g[f"cells_{lev}"], sp.cells[lev].astype(np.uint8))
        assert np.array_equal(g[f"fns_{lev}"], sp.fns[lev].astype(np.uint8))
        for k in range(d):
            assert np.array_equal(g[f"knots_{lev}_{k}"], sp.knotvectors[lev][k])
            if lev < sp.num_levels - 1:
                assert np.allclose(g[f"coeff_{lev}_{k}"], sp.Coeff[lev][k], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", SPACES)
def test_hot_path_matches_reference(name):
    g = load_golden(name)
    sp = oracle_space_from_golden(g)
    ac = O.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = O.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees)  # faithful: 'ones' at max level
    for lev in fc:
        t = fc[lev]
        w = np.cos(np.arange(t.size, dtype=np.float64) * 0.37).reshape(t.shape)
        chk = np.array([t.sum(), (t * w).sum(), float(np.count_nonzero(t))])
        assert np.allclose(chk, g[f"fc_checksum_{lev}"], rtol=1e-12, atol=1e-9)
        if f"fc_{lev}" in g:
            assert np.allclose(t, g[f"fc_{lev}"], rtol=0, atol=1e-7)
    supp, ns, spans = O.compute_active_span(g["params"], sp.knotvectors, sp.cells, sp.degrees, ac, return_spans=True)
    assert np.array_equal(np.array([s[0] for s in spans]), g["span_level"])          # bit-exact
    assert np.array_equal(np.array([s[1] for s in spans]), g["span_cell"])
    assert np.array_equal(np.array(ns), g["num_supp"])
    assert np.array_equal(O.flat_support_indices(supp, sp.sh_fns), g["Jm"])
    PHI = O.thb_basis_fns(g["params"], supp, ns, fc, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert np.abs(PHI - g["PHI"]).max() < 1e-14


def test_fp16_tables_reproduce_reference_quantisation():
    g = dict(np.load(f"{GOLDEN}/space_block2d_fp16.npz"))
    sp = oracle_space_from_golden(g)
    ac = O.compute_active_cells_active_supp(sp.cells, sp.fns, sp.degrees)
    fc = O.compute_refinement_operators(sp.fns, sp.Coeff, sp.degrees, quantize_fp16=True)
    supp, ns = O.compute_active_span(g["params"], sp.knotvectors, sp.cells, sp.degrees, ac)
    PHI = O.thb_basis_fns(g["params"], supp, ns, fc, sp.sh_fns, sp.knotvectors, sp.degrees)
    assert np.abs(PHI - g["PHI"]).max() < 1e-14


@pytest.mark.parametrize("name", ["block2d", "graded2d", "block3d", "readme3d"])
def test_direct_evaluator_equals_table_path(name):
    """Table-free evaluation == dense-table evaluation on spaces whose finest level is inactive."""
    g = load_golden(name)
    sp = oracle_space_from_golden(g)
    de = O.DirectEvaluator(sp.knotvectors, sp.degrees, sp.cells, sp.fns, sp.Coeff)
    pos = 0
    for i, x in enumerate(g["params"]):
        lev, c, Jm, PHI = de.point(x)
        n = int(g["num_supp"][i])
        assert lev == g["span_level"][i] and tuple(c) == tuple(g["span_cell"][i])
        assert np.array_equal(Jm, g["Jm"][pos:pos + n])
        assert np.abs(PHI - g["PHI"][pos:pos + n]).max() < 1e-13
        pos += n


def test_contraction_matches_compiled_reference():
    c = dict(np.load(f"{GOLDEN}/contraction.npz"))
    for tag in "abc":
        fwd = O.eval_forward(c[f"{tag}_cp"], c[f"{tag}_Jm"], c[f"{tag}_phi"], c[f"{tag}_cumsum"])
        bwd = O.eval_backward(c[f"{tag}_grad_out"], c[f"{tag}_cp"].shape, c[f"{tag}_Jm"], c[f"{tag}_phi"], c[f"{tag}_cumsum"])
        assert np.allclose(fwd, c[f"{tag}_fwd"], rtol=1e-5, atol=1e-5)
        assert np.allclose(bwd, c[f"{tag}_bwd"], rtol=1e-5, atol=1e-5)


def test_grid_order_q14():
    p = O.generate_parametric_coordinates((2, 3, 4))
    # z fastest, then x, y slowest
    assert p.shape == (24, 3)
    assert p[1, 2] > p[0, 2] and p[1, 0] == p[0, 0] and p[1, 1] == p[0, 1]
    assert p[4, 0] > p[0, 0] and p[4, 1] == p[0, 1]
    assert p[8, 1] > p[0, 1]
    q = O.generate_parametric_coordinates((3, 2))
    assert q[1, 0] > q[0, 0] and q[1, 1] == q[0, 1] and q[3, 1] > q[0, 1]
    assert q[0, 0] == 1e-5
