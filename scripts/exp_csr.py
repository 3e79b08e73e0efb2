"""THB_basis_fns (PHI materialised) + THB_eval.forward / backward on the CSR of config 3, for timing / ncu."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, core, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

def t(fn, it=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
off, nnz = supp.offsets()
print(f"basis_tiled {t(lambda: engine.basis_tiled(P, supp.plan, off, nnz), 3):.3f} ms  nnz {nnz}")
PHI = engine.basis_tiled(P, supp.plan, off, nnz)[0]
g = torch.randn((prm.shape[0], 3), device="cuda")
for dt in (torch.int64, torch.int32):
    Jm = supp.Jm(dt); cs = supp._cumsum.to(dt)
    print(dt, f"forward {t(lambda: engine.eval_forward(cp, Jm, PHI, cs)):.3f} ms  backward {t(lambda: engine.eval_backward(g, cp, Jm, PHI, cs)):.3f} ms")
    del Jm
