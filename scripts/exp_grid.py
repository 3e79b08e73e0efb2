import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs, engine
from thb_diff_b200.packed import PackedHierarchy
def t(fn, it=50, warm=5):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
ax = [np.linspace(1e-5, 1.0, 256, endpoint=False).astype(np.float32)] * 3
gp = engine.GridPlan(P, ax)
out = torch.empty((256 ** 3, 3), device="cuda")
nets = engine.cell_nets(P, cp, monomial=True)
res = {"items": gp.nitems, "nets": t(lambda: engine.cell_nets(P, cp, monomial=True)), "grid_eval": t(lambda: engine.eval_grid(P, gp, cp, nets=nets, out=out)),
       "step": t(lambda: engine.eval_grid(P, gp, cp, out=out))}
print(json.dumps(res))
