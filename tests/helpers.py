"""Shared test helpers: load golden fixtures, rebuild the same space with the oracle."""
import os

import numpy as np

from oracle import thb_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPACES = ["readme2d", "block2d", "twoblock2d", "graded2d", "block3d", "readme3d"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"space_{name}.npz")))


def oracle_space_from_golden(g):
    d = int(g["ndim"])
    nlev = int(g["num_levels"])
    sp = O.Space([g[f"knots0_{k}"] for k in range(d)], tuple(int(p) for p in g["degrees"]), nlev)
    sp.build_hierarchy_from_domain_sequence()
    for row in g["boxes"]:
        lev = int(row[0])
        lo, hi = row[1:1 + d], row[1 + d:1 + 2 * d]
        sp.refine_box([int(x) for x in lo], [int(x) for x in hi], lev)
    sp.build_hierarchy_from_domain_sequence()
    return sp


def rel_err(a, b, scale=1.0):
    """|a-b| / max(|b|, scale) -- the tolerance metric of SURVEY.md hard part 6."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), scale)


def product_space_from_golden(g):
    from thb_diff_b200 import multilevel_spline_space as M

    d = int(g["ndim"])
    nlev = int(g["num_levels"])
    tp = M.TensorProduct([M.BSpline(g[f"knots0_{k}"], int(g["degrees"][k])) for k in range(d)])
    h = M.Space(tp, nlev)
    h.build_hierarchy_from_domain_sequence()
    for row in g["boxes"]:
        h.refine_box(tuple(int(x) for x in row[1:1 + d]), tuple(int(x) for x in row[1 + d:1 + 2 * d]), int(row[0]))
    h.build_hierarchy_from_domain_sequence()
    return h


def emulate_packed_point(P, x, want_grad=False):
    """NumPy emulation of the tile kernels on the packed tables (tests the host-side builder on CPU):
    finest-level span -> dyadic walk to the active cell -> cell-level Cox-de Boor -> direct + tiled PHI."""
    d, L, p, T = P.ndim, P.nlevels, P.degrees, P.T
    knots = P.knots.cpu().numpy().astype(np.float64)
    cs = P.cell_slot.cpu().numpy()
    info = P.cellinfo.cpu().numpy()
    dmask = P.dmask.cpu().numpy().view(np.uint64)
    jm = P.supp_jm.cpu().numpy()
    tiles = P.tiles.cpu().numpy().astype(np.float64)

    def U(l, k):
        return knots[P.knot_off[l, k]:P.knot_off[l, k] + P.knot_len[l, k]]

    cf = [int(O.find_span_array(np.array([x[k]]), U(L - 1, k), p[k])[0]) - p[k] for k in range(d)]
    if any(c < 0 or c >= P.sh_cells[L - 1][k] for k, c in enumerate(cf)):
        return None
    slot = -1
    for l in range(L - 1, -1, -1):
        c = [v >> (L - 1 - l) for v in cf]
        slot = cs[P.cell_slot_off[l] + np.ravel_multi_index(c, P.sh_cells[l])]
        if slot >= 0:
            break
    if slot < 0:
        return None
    lev, c, soff, S, toff, ndir = info[slot, 0], info[slot, 1:1 + d], info[slot, 4], info[slot, 5], info[slot, 6], info[slot, 7]
    Nd = [O.basis_funs_and_ders(int(c[k]) + p[k], x[k], p[k], U(lev, k)) for k in range(d)]

    def tp(sel):
        t = Nd[0][sel[0]]
        for k in range(1, d):
            t = np.multiply.outer(t, Nd[k][sel[k]])
        return t.reshape(-1)

    chans = [tp([0] * d)] + ([tp([1 if k == j else 0 for k in range(d)]) for j in range(d)] if want_grad else [])
    res = []
    for t in chans:
        phi = [t[b] for b in range(T) if (int(dmask[slot]) >> b) & 1] if ndir else []
        assert len(phi) == ndir
        for s in range(S - ndir):
            phi.append(float(tiles[toff + s, :T] @ t))
        res.append(np.array(phi))
    out = (int(lev), tuple(int(v) for v in c), jm[soff:soff + S].astype(np.int64), res[0])
    if want_grad:
        out = out + (np.stack(res[1:], axis=1),)
    return out


def host_tables_from_packed(P):
    """The packed tables of a (CPU- or GPU-resident) PackedHierarchy as input of the C oracle."""
    from oracle import c_oracle

    return c_oracle.HostTables(P.ndim, P.nlevels, P.degrees, P.TS, P.knot_off, P.knot_len,
                               [list(P.sh_cells[l]) for l in range(P.nlevels)], P.cell_slot_off,
                               P.knots.cpu().numpy(), P.cell_slot.cpu().numpy(), P.cellinfo.cpu().numpy(),
                               P.dmask.cpu().numpy().view(np.uint64), P.supp_jm.cpu().numpy(), P.tiles.cpu().numpy())


def direct_tables(sp):
    """The table-free C evaluator (oracle_direct_eval) of a product / oracle Space: truncated functions from their
    definition, refinement matrices by knot insertion -- nothing of the product's packed tables is used."""
    from oracle import c_oracle

    L = max(sp.knotvectors.keys())
    d = len(sp.degrees)
    Coeff = {l: {k: O.knot_insertion_matrix(sp.knotvectors[l][k], sp.knotvectors[l + 1][k], sp.degrees[k]).T for k in range(d)} for l in range(L)}
    return c_oracle.DirectTables(sp.knotvectors, sp.degrees, sp.cells, sp.fns, Coeff)


def interface_points(n, ndim, seed, lo=0.12, hi=0.88):
    """Random float32 points, half of them within a few finest cells of the faces of the nested refinement boxes of
    configs.cubic3d_space (where truncation and full tiles live), the rest uniform in [lo, hi]^d."""
    rng = np.random.default_rng(seed)
    pts = lo + (hi - lo) * rng.random((n, ndim))
    faces = np.array([3, 13]) / 16.0, np.array([10, 22]) / 32.0, np.array([24, 40]) / 64.0, np.array([54, 74]) / 128.0
    m = n // 2
    for i in range(m):
        f = faces[rng.integers(len(faces))]
        k = rng.integers(ndim)
        pts[i, k] = f[rng.integers(2)] + rng.normal() * 0.01
    return np.clip(pts, 1e-4, 1 - 1e-4).astype(np.float32)


def compare_with_direct(dt, prm, slot_info, Jm, PHI, offsets, tol, dPHI=None):
    """Levels / cells / support ids bit-exact and PHI (and gradient channels) within tol against the table-free evaluator.
    slot_info: cellinfo rows of the points [n, >=6]; Jm/PHI flat at offsets."""
    r = dt.eval(prm.astype(np.float64), max_supp=int(slot_info[:, 5].max()), want_grad=dPHI is not None)
    d = dt.d
    assert np.array_equal(r["lev"], slot_info[:, 0])
    assert np.array_equal(r["cell"][:, :d], slot_info[:, 1:1 + d])
    assert np.array_equal(r["nsupp"], slot_info[:, 5])
    n, ms = r["jm"].shape
    col = np.arange(ms)[None, :]
    valid = col < r["nsupp"][:, None]
    flat = (offsets[:, None] + col)[valid]
    assert np.array_equal(Jm[flat], r["jm"][valid])
    err = rel_err(PHI[flat], r["phi"][valid]).max()
    assert err <= tol, err
    if dPHI is not None:
        for k in range(d):
            ref = r["dphi"][..., k][valid]
            e = np.abs(dPHI[k][flat] - ref).max() / np.abs(ref).max()
            assert e <= tol, (k, e)
    return r
