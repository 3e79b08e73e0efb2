"""Bezier extraction and the adaptive-grid export, against outputs of the reference itself
(tests/golden/bezier_vtu.npz, produced by make_golden.bezier_and_vtu from THB/bspline_funcs.py:242-284,
THB/core.py:295-379 and THB/utils.py:296-399 with a recording stand-in for pyvista)."""
import base64
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, load_golden, product_space_from_golden

G = dict(np.load(os.path.join(GOLDEN, "bezier_vtu.npz")))
U_A = np.array([0, 0, 0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1, 1, 1])
U_B = np.array([0, 0, 0, 0, 0.2, 0.4, 0.5, 0.6, 0.8, 0.9, 1, 1, 1, 1])
U_D = np.array([0, 0, 0, 0, 0.25, 0.5, 0.75, 1, 1, 1, 1])


@pytest.mark.parametrize("tag,U,p", [("A", U_A, 2), ("B", U_B, 3), ("D", U_D, 3)])
def test_bezier_extraction_matches_the_reference(tag, U, p):
    from thb_diff_b200.bezier import bezier_extraction

    C = np.stack(bezier_extraction(U, p))
    assert C.shape == G[f"bez_{tag}"].shape
    assert np.abs(C - G[f"bez_{tag}"]).max() <= 1e-14
    assert np.abs(C.sum(axis=1) - 1).max() <= 1e-14  # each Bernstein polynomial is a partition of the span's B-splines


def test_bezier_extraction_reproduces_the_basis():
    """N_a(u) = sum_j C[a, j] B_j(t) on every span (scipy B-splines as the independent check)."""
    from math import comb

    from scipy.interpolate import BSpline

    from thb_diff_b200.bezier import bezier_extraction

    p = 3
    C = bezier_extraction(U_B, p)
    brk = np.unique(U_B)
    for e in range(len(brk) - 1):
        u = np.linspace(brk[e], brk[e + 1], 7)[1:-1]
        t = (u - brk[e]) / (brk[e + 1] - brk[e])
        bern = np.stack([comb(p, j) * t**j * (1 - t) ** (p - j) for j in range(p + 1)])
        i = int(np.searchsorted(U_B, brk[e], side="right")) - 1
        for a in range(p + 1):
            c = np.zeros(len(U_B) - p - 1)
            c[i - p + a] = 1
            assert np.abs(BSpline(U_B, c, p)(u) - C[e][a] @ bern).max() <= 1e-13


def _parse_vtu(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*>([^<]*)</DataArray>', txt):
        raw = base64.b64decode(m.group(3))
        n = int(np.frombuffer(raw[:8], dtype="<u8")[0])
        dt = {"Float32": "<f4", "Int64": "<i8", "UInt8": "u1", "Int32": "<i4"}[m.group(1)]
        out[m.group(2)] = np.frombuffer(raw[8:8 + n], dtype=dt)
    return out


def test_vtu_writer_round_trip(tmp_path):
    from thb_diff_b200.utils import VTK_HEXAHEDRON, write_vtu

    rng = np.random.default_rng(0)
    pts = rng.random((20, 3)).astype(np.float32)
    conn = rng.integers(0, 20, size=(5, 8))
    path = write_vtu(str(tmp_path / "g"), pts, conn, {"level": np.arange(5, dtype=np.int32)})
    assert path.endswith(".vtu")
    a = _parse_vtu(path)
    assert np.array_equal(a["Points"].reshape(-1, 3), pts)
    assert np.array_equal(a["connectivity"].reshape(-1, 8), conn)
    assert np.array_equal(a["offsets"], np.arange(1, 6) * 8)
    assert np.array_equal(a["types"], np.full(5, VTK_HEXAHEDRON, dtype=np.uint8))
    assert np.array_equal(a["level"], np.arange(5))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["block2d", "block3d"])
def test_multilevel_bezier_operators_match_the_reference(name):
    from thb_diff_b200 import core
    from thb_diff_b200.bezier import compute_multilevel_bezier_extraction_operators

    g = load_golden(name)
    h = product_space_from_golden(g)
    prm = G[f"{name}_params"]
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    fc = core.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    spans, ns, supp = core.compute_active_span_v2(prm, h.knotvectors, h.cells, h.degrees, ac, fc)
    assert list(ns) == G[f"{name}_num_supp"].tolist()
    ops = compute_multilevel_bezier_extraction_operators(prm, supp, ac, fc, h.sh_cells, h.sh_fns, h.knotvectors, h.degrees)
    assert [len(o) for o in ops] == G[f"{name}_num_supp"].tolist()
    assert ops[0].shape[1:] == tuple(p + 1 for p in h.degrees)
    got = np.concatenate([o.reshape(len(o), -1) for o in ops])
    ref = G[f"{name}_ops"]
    assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())  # tiles are stored in float32


@pytest.mark.gpu
def test_adaptive_grid_export_matches_the_reference(tmp_path):
    from thb_diff_b200 import core
    from thb_diff_b200.utils import plot3DAdaptiveGrid

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    fc = core.compute_refinement_operators(h.fns, h.Coeff, h.degrees)
    path, pts, conn = plot3DAdaptiveGrid(str(tmp_path / "grid"), ac, G["vtu_ctrl_pts"], h.knotvectors, fc, h.sh_fns, h.degrees)
    ref_cells = G["vtu_cells"].reshape(-1, 9)
    assert (ref_cells[:, 0] == 8).all() and (G["vtu_types"] == 12).all()
    ref_xyz = G["vtu_points"][ref_cells[:, 1:]]            # [ncells, 8, 3]: what the reference wrote, cell by cell
    assert conn.shape == (len(ref_cells), 8)
    assert np.abs(pts[conn] - ref_xyz).max() <= 1e-5 * max(1.0, np.abs(ref_xyz).max())
    assert len(pts) <= len(G["vtu_points"])                # merged on parameters: never more nodes than the reference
    a = _parse_vtu(path)
    assert np.array_equal(a["connectivity"].reshape(-1, 8), conn)
    assert np.array_equal(a["Points"].reshape(-1, 3), pts)
    assert set(np.unique(a["level"])) <= set(range(h.num_levels))


def test_extraction_tables_reproduce_coarse_functions_on_fine_cells():
    """ExtractionTables.M[(lev, k)][cf]: Bernstein coefficients, on the finest cell cf, of the level-`lev` B-splines of
    the ancestor cell -- checked against scipy B-splines of that level (host-only arithmetic, no GPU)."""
    from math import comb

    from scipy.interpolate import BSpline

    from thb_diff_b200.bezier import ExtractionTables
    from thb_diff_b200.multilevel_spline_space import refine_knotvector

    p = 3
    kv = {0: [U_B]}
    for lev in (1, 2):
        kv[lev] = [refine_knotvector(kv[lev - 1][0], p)]
    tabs = ExtractionTables(kv, (p,))
    fine = np.unique(kv[2][0])
    rng = np.random.default_rng(0)
    for lev in (0, 1, 2):
        U = kv[lev][0]
        M = tabs.M[(lev, 0)]
        assert M.shape == (len(fine) - 1, p + 1, p + 1)
        for cf in rng.choice(len(fine) - 1, 6, replace=False):
            c = cf >> (2 - lev)
            u = np.linspace(fine[cf], fine[cf + 1], 5)[1:-1]
            t = (u - fine[cf]) / (fine[cf + 1] - fine[cf])
            bern = np.stack([comb(p, j) * t**j * (1 - t) ** (p - j) for j in range(p + 1)])
            for i in range(p + 1):
                coef = np.zeros(len(U) - p - 1)
                coef[c + i] = 1
                assert np.abs(BSpline(U, coef, p)(u) - M[cf, i] @ bern).max() <= 1e-12
    cells = tabs.finest_cells(np.array([[0.0], [0.31], [1.0]]))
    assert cells[0, 0] == 0 and cells[2, 0] == len(fine) - 2 and fine[cells[1, 0]] <= 0.31 < fine[cells[1, 0] + 1]
