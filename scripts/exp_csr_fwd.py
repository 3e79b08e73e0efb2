"""THB_eval.forward on the CSR of config 3 (int64 / int32 indices): timing only.  Env: THB_LIB, THB_CSR_BULK."""
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
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
off, nnz = supp.offsets()
PHI = engine.basis_tiled(P, supp.plan, off, nnz)[0]
res = {"lib": os.environ.get("THB_LIB", "-"), "bulk": os.environ.get("THB_CSR_BULK", "-")}
for dt in (torch.int64, torch.int32):
    Jm = supp.Jm(dt); cs = supp._cumsum.to(dt)
    res[str(dt)] = round(t(lambda: engine.eval_forward(cp, Jm, PHI, cs)), 4)
    del Jm
print(res)
