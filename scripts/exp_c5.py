"""Config 5 per-GPU share (5 levels, value + 3 first derivatives, 256^3 sub-grid of the 512^3 grid): timing of the step.
Env THB_NET_PPT (2 / 4), THB_NET_MONO."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy
def t(fn, it=10, warm=3):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
sp = configs.cubic3d_space(5)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
ax = torch.linspace(1e-5, 1.0, 513)[:512].cuda()
prm = torch.stack(torch.meshgrid(ax[::2], ax[::2], ax[::2], indexing="ij"), -1).reshape(-1, 3).contiguous()
plan = engine.Plan(P, prm)
ch = [engine.CH_VALUE] + engine.grad_channels(3)
out = [torch.empty((prm.shape[0], 3), device="cuda") for _ in ch]
res = {"ppt": os.environ.get("THB_NET_PPT"), "mono": os.environ.get("THB_NET_MONO", "1")}
res["plan"] = t(plan.build)
res["nets8"] = t(lambda: engine.cell_nets(P, cp, with_derivatives=True, plan=plan))
nets = engine.cell_nets(P, cp, with_derivatives=True, plan=plan)
res["eval4ch"] = t(lambda: engine.eval_nets(P, plan, nets, 3, ch, out=out))
res["step"] = t(lambda: (plan.build(), engine.eval_fused(P, plan, cp, ch, out=out)))
res["jac_err"] = max(float((out[1 + k][:, j] - (1.0 if j == k else 0.0)).abs().max()) for k in range(3) for j in range(2))
print(json.dumps(res))
