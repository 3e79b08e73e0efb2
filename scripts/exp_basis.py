import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, core, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
prm = B.grid_parametric_coordinates_device((256, 256, 256))
supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
off, nnz = supp.offsets()
for _ in range(3):
    out = engine.basis_tiled(P, supp.plan, off, nnz)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = engine.basis_tiled(P, supp.plan, off, nnz)
b.record(); torch.cuda.synchronize()
print("basis_tiled ms", a.elapsed_time(b) / 5, "nnz", nnz)
