"""THB_basis_fns (PHI rows, k_basis_rows) on config 3: timing only.  Env: THB_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs, core, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

def t(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
off, nnz = supp.offsets()
out = engine.basis_tiled(P, supp.plan, off, nnz)
print({"lib": os.environ.get("THB_LIB", "-"), "basis_rows_ms": round(t(lambda: engine.basis_tiled(P, supp.plan, off, nnz)), 4), "nnz": nnz})
