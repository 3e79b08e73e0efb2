"""Config 4 fit step: GPU time per step vs host enqueue time per step."""
import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs
from thb_diff_b200.packed import PackedHierarchy
from thb_diff_b200.distributed import ShardedFit

sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(0)
prm = torch.rand((10_000_000, 3), device="cuda", generator=gen)
tgt = torch.stack([prm[:, 0], prm[:, 1], prm[:, 2] + 0.1 * torch.sin(6.2831853 * prm[:, 0]) * torch.sin(6.2831853 * prm[:, 1])], dim=1).contiguous()
x = torch.nn.Parameter(torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda())
import os
kw = dict(capturable=True) if os.environ.get("FIT_GRAPH") else {}
opt = torch.optim.Adam([x], lr=1e-2, fused=True, **kw)
fit = ShardedFit(P, prm, tgt, prm.shape[0], **({"graph": True} if os.environ.get("FIT_GRAPH") else {}))
for _ in range(5): l = fit.step(x, opt)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(20): l = fit.step(x, opt)
b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
if os.environ.get("FIT_PROFILE"):
    torch.cuda.profiler.start()
    for _ in range(2): fit.step(x, opt)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
print(f"GPU {a.elapsed_time(b) / 20:.3f} ms/step, host enqueue {(t1 - t0) / 20 * 1e3:.3f} ms/step, loss {float(l):.3e}")
