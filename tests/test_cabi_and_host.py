"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the
header declares (no compute calls), descriptor structs match the header layout, host set-up code."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from thb_diff_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "thb_b200.h")).read()
    declared = set(re.findall(r"THB_API\s+[\w\s\*]+?\b(thb_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in thb_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), "ctypes binding and header disagree"
    assert _lib.lib.thb_version() == 100
    assert _lib.launch_count() == 0  # nothing was launched by loading


def test_descriptor_layout_matches_header(tmp_path):
    """sizeof / offsetof of every field as the C compiler sees include/thb_b200.h == the ctypes mirror."""
    import subprocess

    from thb_diff_b200 import _lib

    structs = {"thb_space_desc": _lib.SpaceDesc, "thb_plan_desc": _lib.PlanDesc}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "thb_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_plan_phase_constants_and_item_size_policy():
    """The Python mirror of the THB_PLAN_* phase bits equals the header's; work items hold 512 points in large batches and
    256 in small ones unless THB_POINTS_PER_ITEM fixes the size."""
    from thb_diff_b200 import engine

    hdr = open(os.path.join(ROOT, "include", "thb_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(THB_PLAN_\w+)\s+(\d+)", hdr))
    assert (int(consts["THB_PLAN_HISTOGRAM"]), int(consts["THB_PLAN_ITEMS"]), int(consts["THB_PLAN_SCATTER"])) == (
        engine.PLAN_HISTOGRAM, engine.PLAN_ITEMS, engine.PLAN_SCATTER) == (1, 2, 4)
    if engine.POINTS_PER_ITEM == 0:
        assert engine.points_per_item_for(16_777_216) == 512 and engine.points_per_item_for(2_000_000) == 256
        assert engine.points_per_item_for(engine.LARGE_BATCH_ITEMS) == 512
    else:
        assert engine.points_per_item_for(1) == engine.POINTS_PER_ITEM


def test_no_cpu_fallback():
    import torch

    from thb_diff_b200 import THB_eval, engine

    cp = torch.zeros((4, 3))
    with pytest.raises(RuntimeError):
        engine.eval_forward(cp, torch.zeros(1, dtype=torch.int64), torch.zeros(1), torch.ones(1, dtype=torch.int64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            THB_eval.cpp_forward(cp, torch.zeros(1, dtype=torch.int64), torch.zeros(1), torch.ones(1, dtype=torch.int64))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "thb_diff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("CPU oracle", "").lower() or f == "distributed.py", f


def test_host_setup_helpers():
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import multilevel_spline_space as M
    from oracle import thb_oracle as O

    U = np.array([0, 0, 0, 0, 0.2, 0.4, 0.5, 0.6, 0.8, 0.9, 1, 1, 1, 1.0])
    V = M.refine_knotvector(U, 3)
    assert np.array_equal(V, O.refine_knotvector(U, 3))
    assert np.allclose(M.assemble_Tmatrix(U, V, len(U), len(V), 3), O.knot_insertion_matrix(U, V, 3), atol=1e-15)
    assert np.array_equal(B.generate_parametric_coordinates((3, 4, 5)), O.generate_parametric_coordinates((3, 4, 5)))
    g = B.grevilleAbscissae((10, 10), (3, 3), (U, U))
    assert np.allclose(g, O.greville_abscissae((10, 10), (3, 3), (U, U)))
    # device grid generator == numpy generator rounded to f32 (run on CPU here)
    dev = B.grid_parametric_coordinates_device((3, 4, 5), device="cpu")
    assert np.array_equal(dev.numpy(), O.generate_parametric_coordinates((3, 4, 5)).astype(np.float32))
    a = B.grid_parametric_coordinates_device((3, 8, 5), device="cpu", y_first=1, y_step=2)
    full = O.generate_parametric_coordinates((3, 8, 5)).astype(np.float32).reshape(8, 3, 5, 3)
    assert np.array_equal(a.numpy(), full[1::2].reshape(-1, 3))


def test_active_supports_dict_view_matches_reference_order():
    from helpers import load_golden, product_space_from_golden
    from oracle import thb_oracle as O
    from thb_diff_b200 import core

    g = load_golden("block2d")
    h = product_space_from_golden(g)
    ac = core.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    ref = O.compute_active_cells_active_supp(h.cells, h.fns, h.degrees)
    assert sorted(ac.keys()) == sorted(ref.keys())
    for lev in ref:
        assert len(ac[lev]) == len(ref[lev])
        for cell in list(ref[lev])[:20]:
            assert ac[lev][cell] == ref[lev][cell]


def test_update_cp_preserves_the_geometry():
    """ControlPoints.update_CP (prolongation of control points on refinement).  The reference's own method
    cannot run (its level offsets are not cumulative, multilevel_spline_space.py:166-176 -> reshape error), so
    the intended semantics are pinned by their defining property: refining cells and prolonging the control
    points leaves the geometry unchanged.  Evaluated on CPU with the C oracle on the packed tables."""
    from helpers import host_tables_from_packed
    from thb_diff_b200 import configs
    from thb_diff_b200 import multilevel_spline_space as M
    from thb_diff_b200.packed import PackedHierarchy

    h = M.Space(M.TensorProduct([M.BSpline(configs.README_U0, 2), M.BSpline(configs.README_U1, 3)]), 4)
    h.build_hierarchy_from_domain_sequence()
    cps = M.ControlPoints(h)
    rng = np.random.default_rng(9)
    n = sum(int(np.prod(h.sh_fns[l])) for l in range(4))
    CP = np.zeros((n, 2))
    CP[: int(np.prod(h.sh_fns[0]))] = rng.standard_normal((int(np.prod(h.sh_fns[0])), 2))
    prm = rng.random((4000, 2)).astype(np.float32)

    def evaluate(cp):
        P = PackedHierarchy.from_space(h, device="cpu")
        cp3 = np.concatenate([cp, np.zeros((len(cp), 1))], axis=1).astype(np.float32)
        return host_tables_from_packed(P).eval(prm, cp3)["out"][:, :2]

    ref = evaluate(CP)
    for lev, lo, hi in ((0, (1, 0), (9, 7)), (1, (6, 4), (14, 10))):   # graded nested boxes
        h.refine_box(lo, hi, lev)
        h.build_hierarchy_from_domain_sequence()
        c = cps.update_CP(CP, h.fns, h.Coeff)
        CP = np.concatenate([c[l].reshape(-1, 2) for l in range(4)])
        assert np.abs(evaluate(CP) - ref).max() < 5e-6
    assert {l: int(h.fns[l].sum()) for l in h.fns} == {0: 60, 1: 220, 2: 126, 3: 0}


def test_exact_span_lut_decides_every_float_with_one_comparison():
    """packed.py proves the span lookup table: with the device's float32 bin formula, the finest cell of EVERY float32 u
    of the domain is lut[bin(u)] or the next one, decided by one comparison with the breakpoint in between
    (k_plan_count_fast<.., EXACT>).  Emulated here in numpy float32 on the adversarial inputs: every breakpoint, its
    float neighbours on both sides, and the first / last float of every bin."""
    import numpy as np

    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    for sp in (configs.cubic3d_space(4), configs.config2_space(), configs.cubic3d_space(5)):
        P = PackedHierarchy.from_space(sp, device="cpu")
        assert P.lut_exact
        lut, off, nbs, scs = P._lut
        L = P.nlevels - 1
        for k in range(P.ndim):
            p = P.degrees[k]
            U = sp.knotvectors[L][k].astype(np.float32)
            B = U[p:len(U) - p]
            nc = len(B) - 1
            nb, sc = nbs[k], np.float32(scs[k])
            # adversarial inputs
            xs = [B, np.nextafter(B, np.float32(-np.inf)), np.nextafter(B, np.float32(np.inf))]
            edges = (B[0].astype(np.float64) + np.arange(nb + 1) / float(sc)).astype(np.float32)
            for j in range(-3, 4):
                e = edges.copy()
                for _ in range(abs(j)):
                    e = np.nextafter(e, np.float32(np.inf if j > 0 else -np.inf))
                xs.append(e)
            xs.append(np.random.default_rng(k).uniform(B[0], B[-1], 200000).astype(np.float32))
            x = np.concatenate(xs)
            x = x[(x >= B[0]) & (x <= B[-1])]
            bins = np.minimum(((x - B[0]) * sc).astype(np.int32), nb - 1)
            c0 = lut[off[k] + bins]
            cell = np.minimum(c0 + (x >= B[np.minimum(c0 + 1, nc)]).astype(np.int32), nc - 1)
            want = np.clip(np.searchsorted(B, x, side="right") - 1, 0, nc - 1)   # find_span_array_jax rule, u == U[-1] -> last cell
            assert np.array_equal(cell, want)
