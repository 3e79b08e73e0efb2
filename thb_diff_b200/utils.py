"""Adaptive-grid export -- SURVEY 8f rank 4.

plot3DAdaptiveGrid mirrors the reference's VTU export [reference: THB/utils.py:296-399]: every active cell becomes one
hexahedron whose eight corners are the geometry evaluated at the corners of the cell in parameter space.  The
reference evaluates each corner with a Python loop over the cell's supports and writes through pyvista; here all
corners go through the fused GPU evaluation (thb_eval_fused) in one call and the file is written directly as VTK XML
(no pyvista).  Corners shared by neighbouring cells are merged on their PARAMETRIC coordinates (exact), so the mesh
is conforming wherever the hierarchy is, independent of floating-point noise in the evaluated geometry."""
import base64
import struct

import numpy as np
import torch

VTK_QUAD, VTK_HEXAHEDRON = 9, 12


def adaptive_grid_cells(P, ctrl_pts):
    """(points [npts, 3] f32, connectivity [ncells, 2^d] int64, level [ncells] int32, params [npts, d] f64) of the
    active cells of the packed hierarchy P, cells in slot order (level ascending, C order -- the reference's dict
    order), corners in VTK order: (x1,y1,z1) (x2,y1,z1) (x2,y2,z1) (x1,y2,z1) then the same at z2."""
    from . import engine

    d = P.ndim
    info = P.cellinfo_host
    knots = P.knots.cpu().numpy().astype(np.float64)
    lo = np.empty((P.nslots, d))
    hi = np.empty((P.nslots, d))
    for k in range(d):
        base = P.knot_off[info[:, 0], k] + P.degrees[k] + info[:, 1 + k]
        lo[:, k], hi[:, k] = knots[base], knots[base + 1]
    pick = [(0, 0), (1, 0), (1, 1), (0, 1)]
    corners = [(a, b, c) for c in (0, 1) for a, b in pick] if d == 3 else pick
    prm = np.stack([np.stack([np.where(sel[k], hi[:, k], lo[:, k]) for k in range(d)], axis=1) for sel in corners], axis=1)
    flat = prm.reshape(-1, d)
    uniq, inverse = np.unique(flat, axis=0, return_inverse=True)
    cp = ctrl_pts
    if isinstance(cp, dict):  # reference layout: {level: [*sh_fns, 3]}
        cp = np.concatenate([np.asarray(cp[l], dtype=np.float32).reshape(-1, np.asarray(cp[l]).shape[-1]) for l in sorted(cp)])
    cp = torch.as_tensor(cp, dtype=torch.float32, device=P.device).contiguous()
    if cp.shape[1] == 2:
        cp = torch.cat([cp, torch.zeros_like(cp[:, :1])], dim=1).contiguous()
    plan = engine.Plan(P, uniq.astype(np.float32))
    pts = engine.eval_fused(P, plan, cp)[0][:, :3].cpu().numpy()
    return pts, inverse.reshape(P.nslots, len(corners)).astype(np.int64), info[:, 0].astype(np.int32), uniq


def write_vtu(filename, points, connectivity, cell_data=None):
    """VTK XML UnstructuredGrid (binary, base64 inline) with one VTK_HEXAHEDRON (8 nodes) or VTK_QUAD (4) per row."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    conn = np.ascontiguousarray(connectivity, dtype=np.int64)
    ncell, nv = conn.shape
    ctype = VTK_HEXAHEDRON if nv == 8 else VTK_QUAD

    def enc(a):
        raw = a.tobytes()
        return base64.b64encode(struct.pack("<Q", len(raw)) + raw).decode()

    def arr(name, a, typ, ncomp=None):
        nc = f' NumberOfComponents="{ncomp}"' if ncomp else ""
        return f'<DataArray type="{typ}" Name="{name}"{nc} format="binary">{enc(a)}</DataArray>\n'

    out = ['<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n',
           f'<UnstructuredGrid>\n<Piece NumberOfPoints="{len(pts)}" NumberOfCells="{ncell}">\n',
           "<Points>\n", arr("Points", pts, "Float32", 3), "</Points>\n<Cells>\n",
           arr("connectivity", conn.reshape(-1), "Int64"),
           arr("offsets", np.arange(1, ncell + 1, dtype=np.int64) * nv, "Int64"),
           arr("types", np.full(ncell, ctype, dtype=np.uint8), "UInt8"), "</Cells>\n"]
    if cell_data:
        out.append("<CellData>\n")
        for name, a in cell_data.items():
            a = np.ascontiguousarray(a)
            typ = {"int32": "Int32", "int64": "Int64", "float32": "Float32", "float64": "Float64", "uint8": "UInt8"}[str(a.dtype)]
            out.append(arr(name, a, typ))
        out.append("</CellData>\n")
    out.append("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    path = filename if str(filename).endswith(".vtu") else str(filename) + ".vtu"
    with open(path, "w") as f:
        f.write("".join(out))
    return path


def plot3DAdaptiveGrid(filename, ac_cells, ctrl_pts, knot_vectors, fn_coeffs, fn_shapes, degrees, cells=None, device="cuda"):
    """[reference: THB/utils.py:296-399] writes `filename`.vtu with one hexahedron per active cell of the hierarchy.

    ac_cells: the ActiveSupports object of compute_active_cells_active_supp (or a PackedHierarchy); fn_coeffs: the
    RefinementOperators object of compute_refinement_operators; ctrl_pts: {level: [*sh_fns, 3]} as in the reference or
    a flat level-major [nCP, 3] array.  Returns (path, points, connectivity)."""
    from .core import ActiveSupports, RefinementOperators
    from .packed import PackedHierarchy

    if isinstance(ac_cells, PackedHierarchy):
        P = ac_cells
    elif isinstance(ac_cells, ActiveSupports):
        P = ac_cells.packed(knot_vectors, device)
    else:
        raise RuntimeError("ac_cells must come from compute_active_cells_active_supp of this package (or be a PackedHierarchy)")
    if not P.has_tiles:
        if not isinstance(fn_coeffs, RefinementOperators):
            raise RuntimeError("the hierarchy has no coefficient tiles yet and fn_coeffs is not a RefinementOperators object")
        P.attach_tiles(fn_coeffs.coeffs)
    pts, conn, level, _ = adaptive_grid_cells(P, ctrl_pts)
    path = write_vtu(filename, pts, conn, {"level": level})
    return path, pts, conn


def plot_active_3D_cells(ac_cells, knotvectors, wd, filename, device="cuda"):
    """[reference: THB/utils.py:464-511] the active cells in PARAMETER space as hexahedra (geometry = identity)."""
    import os

    from .core import ActiveSupports
    from .packed import PackedHierarchy

    P = ac_cells if isinstance(ac_cells, PackedHierarchy) else ac_cells.packed(knotvectors, device) if isinstance(ac_cells, ActiveSupports) else None
    if P is None:
        raise RuntimeError("ac_cells must come from compute_active_cells_active_supp of this package (or be a PackedHierarchy)")
    d = P.ndim
    info = P.cellinfo_host
    knots = P.knots.cpu().numpy().astype(np.float64)
    lo, hi = np.empty((P.nslots, d)), np.empty((P.nslots, d))
    for k in range(d):
        base = P.knot_off[info[:, 0], k] + P.degrees[k] + info[:, 1 + k]
        lo[:, k], hi[:, k] = knots[base], knots[base + 1]
    pick = [(0, 0), (1, 0), (1, 1), (0, 1)]
    corners = [(a, b, c) for c in (0, 1) for a, b in pick] if d == 3 else pick
    prm = np.stack([np.stack([np.where(sel[k], hi[:, k], lo[:, k]) for k in range(d)], axis=1) for sel in corners], axis=1)
    uniq, inverse = np.unique(prm.reshape(-1, d), axis=0, return_inverse=True)
    pts = np.zeros((len(uniq), 3), dtype=np.float32)
    pts[:, :d] = uniq
    return write_vtu(os.path.join(wd, filename), pts, inverse.reshape(P.nslots, len(corners)), {"level": info[:, 0].astype(np.int32)})
