"""Per-point floor of the fused kernel: a space with NO refinement (every support is own-level, no tile is read)."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy
from thb_diff_b200.multilevel_spline_space import BSpline, Space, TensorProduct

prm = B.grid_parametric_coordinates_device((256, 256, 256))
rng = np.random.default_rng(0)
sp0 = Space(TensorProduct([BSpline(configs.uniformish_knots(16, 3, rng), 3) for _ in range(3)]), 2)
sp0.build_hierarchy_from_domain_sequence()
P = PackedHierarchy.from_space(sp0, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp0)).cuda()
plan = engine.Plan(P, prm)
out = [torch.empty((prm.shape[0], 3), device="cuda")]
for _ in range(3): engine.eval_fused(P, plan, cp, out=out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): engine.eval_fused(P, plan, cp, out=out)
b.record(); torch.cuda.synchronize()
print(f"unrefined kernel {a.elapsed_time(b) / 10:.3f} ms")
