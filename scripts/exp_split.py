import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy
from thb_diff_b200.multilevel_spline_space import BSpline, Space, TensorProduct

import os
FORM = os.environ.get("FORM", "local_net")
def timeit(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

prm = B.grid_parametric_coordinates_device((256, 256, 256))
# (a) unrefined: every point is direct-only
rng = np.random.default_rng(0)
sp0 = Space(TensorProduct([BSpline(configs.uniformish_knots(16, 3, rng), 3) for _ in range(3)]), 2)
sp0.build_hierarchy_from_domain_sequence()
for name, sp in (("unrefined", sp0), ("c3", configs.cubic3d_space(4))):
    P = PackedHierarchy.from_space(sp, device="cuda")
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
    plan = engine.Plan(P, prm)
    out = [torch.empty((prm.shape[0], 3), device="cuda")]
    t = timeit(lambda: engine.eval_fused(P, plan, cp, out=out, formulation=FORM))
    tp = timeit(lambda: plan.build())
    n_in, nnz, nt = plan.stats()
    print(f"{name}: kernel {t:.3f} ms  plan {tp:.3f} ms  nnz_tiled {nt/1e6:.1f}M  -> {prm.shape[0]/t/1e6:.2f} Gpts/s kernel-only")
