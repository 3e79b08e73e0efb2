"""One rank's share of the strong-scaling step on ONE GPU: slab `rank` of `world` of the fixed 256^3 grid of config 3
(plan + nets of the touched cells + evaluation).  argv: world rank.  Prints graph / eager step times and phases."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from thb_diff_b200 import configs, engine
from thb_diff_b200.distributed import grid_slab_bounds
from thb_diff_b200.packed import PackedHierarchy

world, rank = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
sp = configs.cubic3d_space(4)
P = PackedHierarchy.from_space(sp, device="cuda")
cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
axes = bench.grid_axes(256, dev)
bnd = grid_slab_bounds(P, axes, world, cell_cost=100.0).tolist()
pts = bench.grid_slab(axes, bnd[rank], bnd[rank + 1])
plan = engine.Plan(P, pts)
out = [torch.zeros((pts.shape[0], 3), device=dev)]
MODE = {"serial": False, "all": True}.get(os.environ.get("SLAB_OVERLAP", "serial"), os.environ.get("SLAB_OVERLAP", "serial"))
def step():
    nets = engine.plan_and_nets(P, plan, cp, overlap=MODE)
    engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
def t(fn, it=200, warm=20):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / it, 4)
res = {"mode": str(MODE), "world": world, "rank": rank, "planes": bnd[rank + 1] - bnd[rank], "points": pts.shape[0], "eager": t(step)}
run, isg = bench.graphed(step, dev)
res["graph"] = t(run)
ph = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(50)]
for e in ph:
    e[0].record(); plan.build(); e[1].record()
    nets = engine.cell_nets(P, cp, plan=plan); e[2].record()
    engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False); e[3].record()
torch.cuda.synchronize()
res["phases"] = [round(sum(e[k].elapsed_time(e[k + 1]) for e in ph) / len(ph), 4) for k in range(3)]
res["cells_with_points"] = int((plan.cell_count[: P.nslots] > 0).sum())
res["items"] = int(plan.counters[0])
print(json.dumps(res))
