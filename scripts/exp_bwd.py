"""Fused backward on C3 grid points and on 10 M random points (config 4)."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

def t(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
for name, prm in (("grid 256^3", B.grid_parametric_coordinates_device((256, 256, 256))),
                  ("random 10M", torch.from_numpy(np.random.default_rng(0).random((10_000_000, 3), dtype=np.float32)).cuda())):
    plan = engine.Plan(P, prm)
    g = torch.randn((prm.shape[0], 3), device="cuda")
    a = engine.eval_fused_backward(P, plan, g, formulation="local_net")
    b = engine.eval_fused_backward(P, plan, g, formulation="per_point")
    print(name, "max rel diff", float((a - b).abs().max() / b.abs().max()),
          f"local_net {t(lambda: engine.eval_fused_backward(P, plan, g, formulation='local_net')):.3f} ms",
          f"per_point {t(lambda: engine.eval_fused_backward(P, plan, g, formulation='per_point')):.3f} ms")
