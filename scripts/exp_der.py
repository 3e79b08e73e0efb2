"""Value + three first derivatives on config 3 (the per-GPU share of BASELINE config 5 uses the same kernels)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
plan = engine.Plan(P, prm)
ch = [0] + engine.grad_channels(3)
outs = [torch.zeros((prm.shape[0], 3), device="cuda") for _ in ch]
def run(): engine.eval_fused(P, plan, cp, ch, out=outs)
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print(f"nets(8 planes) + eval(4 channels) {a.elapsed_time(b) / 10:.3f} ms")
