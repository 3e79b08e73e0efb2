"""Every kernel of the path once on BASELINE config 3 (256^3 points), inside a cudaProfilerStart/Stop range, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o out python scripts/prof_all.py
and, without ncu, CUDA-event timings of the same calls (printed as JSON).  Args: [what ...] from
plan nets eval bwd basis csr32 csr64 fit (default: all)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs, core, engine, bspline_funcs as B
from thb_diff_b200.packed import PackedHierarchy

what = set(sys.argv[1:]) or {"plan", "nets", "eval", "bwd", "basis", "csr32", "csr64", "grid", "fit"}
under_ncu = "NV_NSIGHT_INJECTION_PORT_BASE" in os.environ or os.environ.get("THB_PROFILING") == "1"

def t(fn, it=10, warm=3):
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
g = torch.randn((prm.shape[0], 3), device="cuda")
calls = {}
if "plan" in what: calls["plan"] = plan.build
if "nets" in what: calls["nets"] = lambda: engine.cell_nets(P, cp, plan=plan)
if "eval" in what:
    nets = engine.cell_nets(P, cp, plan=plan)
    calls["eval"] = lambda: engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
if "bwd" in what:
    calls["bwd"] = lambda: engine.eval_fused_backward(P, plan, g)
    calls["bwd_rows"] = lambda: engine.eval_fused_backward_rows(P, plan, g)
if what & {"basis", "csr32", "csr64"}:
    supp, ns = core.compute_active_span(prm, sp.knotvectors, sp.cells, sp.degrees, P)
    off, nnz = supp.offsets()
    if "basis" in what: calls["basis"] = lambda: engine.basis_tiled(P, supp.plan, off, nnz)
    if what & {"csr32", "csr64"}:
        PHI = engine.basis_tiled(P, supp.plan, off, nnz)[0]
        for name, dt in (("csr64", torch.int64), ("csr32", torch.int32)):
            if name in what:
                Jm = supp.Jm(dt); cs = supp._cumsum.to(dt)
                calls[name + "_fwd"] = (lambda Jm=Jm, cs=cs: engine.eval_forward(cp, Jm, PHI, cs))
                calls[name + "_bwd"] = (lambda Jm=Jm, cs=cs: engine.eval_backward(g, cp, Jm, PHI, cs))
if "grid" in what:
    import numpy as np
    ax = [np.linspace(1e-5, 1.0, 256, endpoint=False).astype(np.float32)] * 3
    gp = engine.GridPlan(P, ax)
    gnets = engine.cell_nets(P, cp, monomial=True).clone()
    gnets._thb_flags = engine.NETS_MONOMIAL
    gout = torch.empty((256 ** 3, 3), device="cuda")
    calls["grid"] = lambda: engine.eval_grid(P, gp, cp, nets=gnets, out=gout)
if "fit" in what:
    from thb_diff_b200.distributed import FitEngine
    gen = torch.Generator(device="cuda").manual_seed(0)
    fp = torch.rand((10_000_000, 3), device="cuda", generator=gen)
    ft = fp.clone()
    fx = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    feng = FitEngine(P, fp, ft, fp.shape[0])
    calls["fit_step"] = lambda: feng.step(fx)
    feng_d = FitEngine(P, fp, ft, fp.shape[0], deterministic=True)
    calls["fit_step_ordered"] = lambda: feng_d.step(fx)
res = {}
if not under_ncu:
    for k, fn in calls.items(): res[k] = t(fn)
    print(json.dumps(res))
else:
    for fn in calls.values(): fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for fn in calls.values(): fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
