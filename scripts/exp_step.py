"""One step of the headline (plan + nets + evaluation) on config 3, 256^3 points: serial vs nets-beside-plan overlap.
Env: THB_PLAN_CTAS_PER_SM.  Prints JSON."""
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
    return a.elapsed_time(b) / it

prm = B.grid_parametric_coordinates_device((256, 256, 256))
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
plan = engine.Plan(P, prm)
out = [torch.zeros((prm.shape[0], 3), device="cuda")]
def serial():
    nets = engine.plan_and_nets(P, plan, cp, overlap=False)
    engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
def overlap():
    nets = engine.plan_and_nets(P, plan, cp, overlap=True)
    engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
res = {"ctas_per_sm": os.environ.get("THB_PLAN_CTAS_PER_SM", "16"), "lut_exact": bool(P.lut_exact), "plan": t(plan.build), "serial": t(serial), "overlap": t(overlap)}
ref = out[0].clone(); serial(); torch.cuda.synchronize()
res["same"] = bool(torch.equal(ref, out[0]))
print(json.dumps(res))
