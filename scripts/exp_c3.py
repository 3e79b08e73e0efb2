"""C3 fused evaluation only (for ncu captures of the light / tile kernels)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
plan = engine.Plan(P, prm)
out = [torch.empty((prm.shape[0], 3), device="cuda")]
for _ in range(3): engine.eval_fused(P, plan, cp, out=out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): engine.eval_fused(P, plan, cp, out=out)
b.record(); torch.cuda.synchronize()
print(f"c3 kernel {a.elapsed_time(b) / 10:.3f} ms; items {plan.counters.tolist()[0]}")
