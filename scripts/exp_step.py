"""One step of the headline (plan + nets + evaluation) on config 3, 256^3 points: serial vs nets-beside-plan overlap
(nets beside the whole plan, or beside the scatter pass only).
Env: THB_PLAN_CTAS_PER_SM, THB_SCATTER_CTAS_PER_SM, THB_OVERLAP_NETS_CTAS, THB_COUNT_VEC.  Prints JSON."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

def t(fn, it=200, warm=20):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / it, 4)

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
plan = engine.Plan(P, prm)
out = [torch.zeros((prm.shape[0], 3), device="cuda")]
def step(mode):
    def f():
        nets = engine.plan_and_nets(P, plan, cp, overlap=mode)
        engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
    return f
res = {k: os.environ.get(k, "-") for k in ("THB_PLAN_CTAS_PER_SM", "THB_SCATTER_CTAS_PER_SM", "THB_OVERLAP_NETS_CTAS", "THB_COUNT_VEC")}
res.update({"plan": t(plan.build), "plan_phase1": t(lambda: plan.build(phase=1)), "plan_phase2": t(lambda: plan.build(phase=2)),
            "serial": t(step(False)), "overlap": t(step(True)), "overlap_scatter": t(step("scatter"))})
step(False)(); torch.cuda.synchronize(); ref = out[0].clone()
for m in (True, "scatter"):
    out[0].zero_(); step(m)(); torch.cuda.synchronize()
    res[f"same_{m}"] = bool(torch.equal(ref, out[0]))
print(json.dumps(res))
