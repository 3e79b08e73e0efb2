#!/usr/bin/env python
"""Benchmark of the THB basis+evaluation hot path (BASELINE.json metric: points/sec, fraction of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 3 -- 3-D, degree 3, 4 active levels (nested graded boxes, ~25 % of
each region refined, 'uniform-ish' knots), a 256^3 parametric grid PER GPU (weak scaling: the y axis of a
256 x 256N x 256 grid is dealt round-robin to the N ranks, so every rank sees the same mix of cells).
One step = `--inner` passes of: params (resident in HBM) -> points grouped by active cell (thb_plan_build) -> local
control nets -> fused basis+evaluation kernel -> geometry [N,3] in HBM (the passes are identical; they make the timed
region last about a second so that clocks are sampled inside it).  No collective is on this path (points are independent).

Every run, at every N, also measures (flat keys in `config` / `roofline`, details under `blocks`):
  strong : the FIXED 256^3 grid of config 3 cut into N cost-balanced slabs (strong scaling of the same step)
  fit_c4 : BASELINE config 4 -- 10M target points, MSE, backward into compact active-function gradient rows, NCCL
           all-reduce of those rows, Adam (steps/s; all-reduce timed apart)
  c5     : BASELINE config 5 -- 5 active levels, value + three first derivatives on the 512^3 grid cut into N slabs
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 256
SAMPLE_REF = 2048


def cdb_flops(p):
    return sum(2 + 5 * j for j in range(1, p + 1))


def flop_model(degrees, cpd, n_in, nnz, nnz_tiled, nnz_sep=0):
    """(executed FLOPs of the kernel's algorithm, FLOPs of SURVEY 8d's dense gather-multiply model).
    Executed: per point Cox-de Boor + tensor product + cp_dim window dot products (own-level supports);
    per (point, rank-one tiled support) one (p_k+1)-dot per dimension, two multiplications, the contraction;
    per (point, full-tile support) a (p+1)^d dot and the contraction."""
    d = len(degrees)
    T = int(np.prod([p + 1 for p in degrees]))
    tp = (degrees[0] + 1) * (degrees[1] + 1) + (T if d == 3 else 0)
    per_pt = sum(cdb_flops(p) for p in degrees) + tp
    sep = 2 * sum(p + 1 for p in degrees) + (d - 1) + 2 * cpd
    executed = n_in * (per_pt + cpd * 2 * T) + nnz_sep * sep + (nnz_tiled - nnz_sep) * (2 * T + 2 * cpd)
    dense = n_in * per_pt + nnz * (2 * T + 2 * cpd)
    return float(executed), float(dense)


def flop_model_local_net(P, plan, cpd, n_in, mono=False):
    """(FLOPs of k_net_eval, FLOPs of k_cell_nets), thb_net_impl.cuh.
    Evaluation, control-point nets: per point Cox-de Boor and the sum-factorised contraction with the cell's net
    (T + N0*N2 + N0 FMAs per output component, 3-D).  Monomial nets: per point three (subtract, FMA) pairs for the local
    coordinates and a nested Horner scheme, (N1-1)*N0*N2 + (N0-1)*N2 + (N2-1) FMAs per output component.
    Nets: per cell that holds points, per rank-one support 1.5 multiplications + cp_dim FMAs for each of the TS window
    entries, per full tile cp_dim FMAs per entry; monomial form: + three passes of N_k FMAs per entry and component."""

    deg = P.degrees
    d = len(deg)
    N = [p + 1 for p in deg] + ([1] if d == 2 else [])
    T = N[0] * N[1] * N[2]
    if mono:
        per_pt = 3 * d + cpd * 2 * ((N[1] - 1) * N[0] * N[2] + (N[0] - 1) * N[2] + (N[2] - 1))
    else:
        per_pt = sum(cdb_flops(p) for p in deg) + cpd * 2 * (T + N[0] * N[2] + N[0])
    used = (plan.cell_count[: P.nslots] > 0).long()
    info = P.cellinfo.long()
    nt, nsep = info[:, 5] - info[:, 7], info[:, 9]
    fold = (used * nsep).sum() * P.TS * (1.5 + 2 * cpd) + (used * (nt - nsep)).sum() * P.TS * 2 * cpd
    if mono:
        fold = fold + used.sum() * T * cpd * 2 * sum(N[:d])
    return float(n_in * per_pt), float(fold)


def bind_host_to_gpu(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (the e2e path moves 400 MB per step through
    pinned host memory; with one rank per GPU the ranks otherwise share whatever NUMA node the launcher started them
    on).  Returns the number of CPUs bound to, or None when NVML / affinity is not available."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return len(cpus) if cpus else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_workload(rank, world, device, dense_tables=False):
    import torch

    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import configs
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=4)
    P = PackedHierarchy.from_space(sp, device=device, onehot_direct=not dense_tables)
    params = B.grid_parametric_coordinates_device((GRID, GRID * world, GRID), device=device, y_first=rank, y_step=world)
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).to(device)
    return sp, P, params, cp


def host_tables(P):
    from oracle import c_oracle

    return c_oracle.HostTables(P.ndim, P.nlevels, P.degrees, P.TS, P.knot_off, P.knot_len, [list(P.sh_cells[l]) for l in range(P.nlevels)],
                               P.cell_slot_off, P.knots.cpu().numpy(), P.cell_slot.cpu().numpy(), P.cellinfo.cpu().numpy(),
                               P.dmask.cpu().numpy().view(np.uint64), P.supp_jm.cpu().numpy(), P.tiles.cpu().numpy())


def cpu_port_rate(P, params_host, cp_host, target_s=12.0):
    """points/s of the plain-C oracle port (OpenMP, all host threads) on a strided sample of the workload."""
    from oracle import c_oracle

    ht = host_tables(P)
    c_oracle.use_all_host_threads()
    n = len(params_host)
    probe = params_host[:: max(n // 100_000, 1)]
    t0 = time.perf_counter()
    ht.eval(probe, cp_host)
    dt = time.perf_counter() - t0
    m = int(min(n, max(len(probe), len(probe) * target_s / max(dt, 1e-3))))
    sample = params_host[:: max(n // m, 1)]
    t0 = time.perf_counter()
    ht.eval(sample, cp_host)
    dt = time.perf_counter() - t0
    return len(sample) / dt, c_oracle.num_threads(), len(sample)


def time_cuda(fn, iters=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def extras(sp, P, params, cp, hbm_peak):
    """Secondary measurements on the same workload (N=1 only): the API-faithful path with PHI materialised
    (compute_active_span + THB_basis_fns, then THB_eval.forward / backward on the CSR) next to the REFERENCE'S
    OWN CUDA KERNELS (oracle/_ref/cuda = compute_pts_cuda_kernels.cu compiled unmodified for sm_100a) on the
    same tensors, and the fused backward."""
    import torch

    from oracle import build_ref
    from thb_diff_b200 import core, engine

    n = params.shape[0]
    out = {}
    supp, ns = core.compute_active_span(params, sp.knotvectors, sp.cells, sp.degrees, P)
    off, nnz = supp.offsets()
    t_basis = time_cuda(lambda: engine.basis_tiled(P, supp.plan, off, nnz), iters=3, warm=1)
    PHI = engine.basis_tiled(P, supp.plan, off, nnz)[0]
    out["THB_basis_fns"] = {"ms": t_basis, "points_per_s": n / (t_basis * 1e-3), "nnz": nnz,
                            "hbm_frac": ((nnz * 4 + n * 24) / (t_basis * 1e-3) / 1e9) / hbm_peak}
    cumsum = supp._cumsum
    g = torch.randn((n, 3), device=params.device)
    for name, dt in (("int64", torch.int64), ("int32", torch.int32)):
        Jm = supp.Jm(dt)
        cs = cumsum.to(dt) if dt == torch.int64 or nnz < 2 ** 31 else cumsum
        isz = 8 if dt == torch.int64 else 4
        tf = time_cuda(lambda: engine.eval_forward(cp, Jm, PHI, cs))
        tb = time_cuda(lambda: engine.eval_backward(g, cp, Jm, PHI, cs))
        out[f"THB_eval_{name}"] = {
            "forward_ms": tf, "backward_ms": tb, "forward_points_per_s": n / (tf * 1e-3), "backward_points_per_s": n / (tb * 1e-3),
            "forward_hbm_frac": ((nnz * (4 + isz) + n * (isz + 12)) / (tf * 1e-3) / 1e9) / hbm_peak,
            "backward_hbm_frac": ((nnz * (4 + isz) + n * (isz + 12)) / (tb * 1e-3) / 1e9) / hbm_peak}
        if dt == torch.int64:
            ref = build_ref.load_cuda()
            if ref is not None:
                o_ref = ref.forward(cp, Jm, PHI, cs)
                o_our = engine.eval_forward(cp, Jm, PHI, cs)
                rf = time_cuda(lambda: ref.forward(cp, Jm, PHI, cs), iters=3, warm=1)
                rb = time_cuda(lambda: ref.backward(g, cp, Jm, PHI, cs), iters=3, warm=1)
                out["reference_cuda_kernels_sm100a"] = {
                    "forward_ms": rf, "backward_ms": rb, "max_abs_diff_forward": float((o_ref - o_our).abs().max()),
                    "speedup_forward": rf / tf, "speedup_backward": rb / tb,
                    "what": "THB/THB_extensions/compute_pts_cuda_kernels.cu compiled unmodified (oracle/_ref/cuda), same tensors"}
                del o_ref, o_our
        elif "reference_cuda_kernels_sm100a" in out:   # the same work with the 32-bit indices this library also accepts
            rc = out["reference_cuda_kernels_sm100a"]
            rc["speedup_forward_int32_indices"], rc["speedup_backward_int32_indices"] = rc["forward_ms"] / tf, rc["backward_ms"] / tb
        del Jm
    tfb = time_cuda(lambda: engine.eval_fused_backward(P, supp.plan, g))
    tfb_pp = time_cuda(lambda: engine.eval_fused_backward(P, supp.plan, g, formulation="per_point"))
    out["fused_backward"] = {"ms": tfb, "points_per_s": n / (tfb * 1e-3), "per_point_formulation_ms": tfb_pp,
                             "what": "k_net_backward (per-item reduction into per-cell net gradients) + k_cell_unfold (transposed fold per cell)"}
    del PHI, g, supp
    torch.cuda.empty_cache()

    return out


def grid_axes(n, device):
    import torch

    return [torch.from_numpy(np.linspace(1e-5, 1.0, int(n), endpoint=False)).to(device=device, dtype=torch.float32) for _ in range(3)]


def grid_slab(axes, lo, hi):
    """Points of the planes [lo, hi) of axis 0 of the tensor-product grid, axis 0 slowest in memory."""
    import torch

    return torch.stack(torch.meshgrid(axes[0][lo:hi], axes[1], axes[2], indexing="ij"), dim=-1).reshape(-1, 3).contiguous()


def timed_max_over_ranks(fn, iters, warm, world, device):
    """ms per call of fn (CUDA events on the current stream, barrier on both sides, max over ranks)."""
    import torch
    import torch.distributed as dist

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def graphed(fn, device):
    """fn captured into one CUDA graph (falls back to eager launches when capture is not possible)."""
    import torch

    fn()
    torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay, True
    except Exception:
        torch.cuda.synchronize()
        return fn, False


def strong_block(P, cp, rank, world, device, cell_cost):
    """The fixed 256^3 grid of config 3 cut into `world` slabs of x planes balanced by points + per-cell cost; one step
    = plan + nets of the cells the slab touches + evaluation, replayed as one CUDA graph."""
    import torch
    import torch.distributed as dist

    from thb_diff_b200 import engine
    from thb_diff_b200.distributed import grid_slab_bounds

    axes = grid_axes(GRID, device)
    bnd = grid_slab_bounds(P, axes, world, cell_cost=cell_cost).tolist()
    pts = grid_slab(axes, bnd[rank], bnd[rank + 1])
    plan = engine.Plan(P, pts)
    out = [torch.zeros((pts.shape[0], 3), dtype=torch.float32, device=device)]

    def step():
        # a slab is a small batch: the histogram pass first, then the nets of the touched cells on a second stream beside the
        # rest of the plan (plan_and_nets picks that below 12 M points; the full batch keeps the nets beside the whole plan)
        nets = engine.plan_and_nets(P, plan, cp, overlap="auto")
        engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)

    run, is_graph = graphed(step, device)
    ms = timed_max_over_ranks(run, 200, 20, world, device)
    mine = torch.tensor([timed_max_over_ranks(run, 200, 5, 1, device)], device=device)   # this rank alone, for the balance
    all_ms = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_ms, mine)
    else:
        all_ms = [mine]
    # where the step goes on this rank (eager launches, CUDA events): the parts that do not shrink with the slab explain
    # the distance from linear scaling
    ph = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(20)]
    for e in ph:
        e[0].record(); plan.build(); e[1].record()
        nets = engine.cell_nets(P, cp, plan=plan); e[2].record()
        engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False); e[3].record()
    torch.cuda.synchronize()
    phases = torch.tensor([sum(e[k].elapsed_time(e[k + 1]) for e in ph) / len(ph) for k in range(3)], device=device)
    if world > 1:
        dist.all_reduce(phases, op=dist.ReduceOp.MAX)
    # Greville control points reproduce the parameters: the slab's geometry must equal its own points
    err = float((out[0][:, :2] - pts[:, :2]).abs().max())
    return {"points_per_s": GRID ** 3 / (ms * 1e-3), "ms_per_step": ms, "planes": [bnd[r + 1] - bnd[r] for r in range(world)],
            "rank_ms": [float(x) for x in all_ms], "cuda_graph": is_graph, "greville_max_abs_err_xy": err,
            "eager_phase_ms_max_over_ranks": {"plan": float(phases[0]), "nets": float(phases[1]), "eval": float(phases[2])},
            "cells_with_points": int((plan.cell_count[: P.nslots] > 0).sum()),
            "what": "fixed 256^3 grid, x-plane slabs balanced by points + cell_cost per touched cell; plan + nets + evaluation per step"}


def grid_block(P, cp, rank, world, device):
    """The tensor-product-grid path (thb_eval_grid): same 256^3 grid per GPU as the headline, but evaluated as a GRID -- no
    plan (the points of a cell are an index box; engine.GridPlan lists the boxes on the host once per grid) and sum
    factorisation at the cell level.  Step = local control nets (monomial) + grid evaluation.  Reported next to the
    headline, which treats the same points as an arbitrary batch."""
    import time as _t

    import torch

    from thb_diff_b200 import engine

    ax = [np.linspace(1e-5, 1.0, GRID, endpoint=False).astype(np.float32), np.linspace(1e-5, 1.0, GRID * world, endpoint=False).astype(np.float32)[rank::world],
          np.linspace(1e-5, 1.0, GRID, endpoint=False).astype(np.float32)]
    t0 = _t.perf_counter()
    gp = engine.GridPlan(P, ax)
    t_plan = _t.perf_counter() - t0
    out = torch.empty((gp.n_points, 3), dtype=torch.float32, device=device)

    def step():
        engine.eval_grid(P, gp, cp, out=out)

    ms = timed_max_over_ranks(step, 200, 20, world, device)
    nets = engine.cell_nets(P, cp, monomial=True)
    ms_eval = timed_max_over_ranks(lambda: engine.eval_grid(P, gp, cp, nets=nets, out=out), 200, 10, world, device)
    return {"points_per_s": gp.n_points * world / (ms * 1e-3), "ms_per_step": ms, "grid_kernel_ms": ms_eval, "boxes": gp.nitems,
            "grid_plan_host_ms_once": t_plan * 1e3,
            "hbm_frac_of_output_write": (gp.n_points * 12 / (ms_eval * 1e-3) / 1e9) / 6538.9,
            "what": "nets (monomial, every cell) + thb_eval_grid on the 256^3 grid of this rank; GridPlan built once per grid on the host"}


def c5_block(rank, world, device, cell_cost):
    """BASELINE config 5: 5 active levels, value + the three first-derivative vectors on the 512^3 grid cut into slabs."""
    import torch

    from thb_diff_b200 import configs, engine
    from thb_diff_b200.distributed import grid_slab_bounds
    from thb_diff_b200.packed import PackedHierarchy

    sp5 = configs.cubic3d_space(active_levels=5)
    P5 = PackedHierarchy.from_space(sp5, device=device)
    cp5 = torch.from_numpy(configs.greville_ctrl_pts(sp5)).to(device)
    axes = grid_axes(512, device)
    bnd = grid_slab_bounds(P5, axes, world, cell_cost=cell_cost).tolist()
    pts = grid_slab(axes, bnd[rank], bnd[rank + 1])
    plan = engine.Plan(P5, pts)
    ch = [engine.CH_VALUE] + engine.grad_channels(3)
    o5 = [torch.empty((pts.shape[0], 3), device=device) for _ in ch]

    def step():
        plan.build()
        engine.eval_fused(P5, plan, cp5, ch, out=o5)

    ms = timed_max_over_ranks(step, 10, 3, world, device)
    # the x / y rows of the control points are Greville abscissae: d x_k / d u_k = 1 and the cross terms vanish
    jac = max(float((o5[1 + k][:, j] - (1.0 if j == k else 0.0)).abs().max()) for k in range(3) for j in range(2))
    pos = float((o5[0][:, :2] - pts[:, :2]).abs().max())
    n_tot = 512 ** 3
    return {"points_per_s": n_tot / (ms * 1e-3), "ms_per_step": ms, "points": n_tot, "points_this_rank": int(pts.shape[0]),
            "levels": 5, "nCP": P5.nCP, "table_bytes": P5.nbytes(), "jacobian_max_abs_err": jac, "position_max_abs_err": pos,
            "planes": [bnd[r + 1] - bnd[r] for r in range(world)],
            "what": "512^3 grid in x-plane slabs over the ranks; plan + nets + evaluation, 4 channels (value, d/du, d/dv, d/dw) [N,3] each"}


def fit_block(sp, P, rank, world, device, cell_cost, n_total=10_000_000):
    """BASELINE config 4: 10M synthetic targets, MSE, backward into the compact active rows, all-reduce (NCCL), Adam."""
    import torch
    import torch.distributed as dist

    from thb_diff_b200 import configs, engine
    from thb_diff_b200.distributed import FitEngine, estimate_cell_cost, partition_by_cell, shard_bounds

    gen = torch.Generator(device=device).manual_seed(0)
    prm = torch.rand((n_total, 3), device=device, generator=gen)          # the same 10M points whatever the world size
    lo, hi = shard_bounds(n_total, rank, world)
    prm = prm[lo:hi].contiguous()                                         # this rank's arbitrary share ...
    tgt = torch.stack([prm[:, 0], prm[:, 1], prm[:, 2] + 0.1 * torch.sin(6.2831853 * prm[:, 0]) * torch.sin(6.2831853 * prm[:, 1])
                       * torch.sin(6.2831853 * prm[:, 2])], dim=1)
    payload = torch.cat([prm, tgt], dim=1)
    slot, _ = engine.active_span(P, prm, want_num_supp=False)
    x = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).to(device)
    balance = {"cell_cost_first": cell_cost}
    for attempt in range(2):
        rows, bounds = partition_by_cell(slot, P.nslots, payload, cell_cost=cell_cost)   # ... moved to the owner of its cell
        prm, tgt = rows[:, :3].contiguous(), rows[:, 3:].contiguous()
        if world == 1 or attempt == 1:
            break
        # measured re-balancing: a few steps with the guessed cell cost, per-rank compute time, least squares t = a points + b cells
        eng = FitEngine(P, prm, tgt, n_total, lr=1e-2)
        xx = x.clone()
        for _ in range(3):
            eng.step(xx)
        t_c = eng.time_phases(xx, iters=10)[0]
        ncell = int((eng.plan.cell_count[: P.nslots] > 0).sum())
        mine = torch.tensor([float(prm.shape[0]), float(ncell), t_c], device=device)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        pts_r, cells_r, t_r = ([float(v[k]) for v in allv] for k in range(3))
        est = estimate_cell_cost(pts_r, cells_r, t_r)
        balance.update({"first_points": pts_r, "first_cells": cells_r, "first_compute_ms": t_r, "cell_cost_estimated": est})
        del eng, xx
        if est is None or max(t_r) <= 1.1 * (sum(t_r) / world):
            break
        cell_cost = est
    balance["cell_cost_used"] = cell_cost
    del rows, slot, payload
    res = {"balance": balance}
    variants = [("eager", False, None), ("graph", True, None)] + ([("graph_nccl", True, False)] if world > 1 else [])
    for name, use_graph, peer in variants:
        xx = x.clone()
        eng = FitEngine(P, prm, tgt, n_total, lr=1e-2, use_graph=use_graph, peer=peer)
        losses = []
        for _ in range(3):
            l = eng.step(xx).clone()
            if world > 1:
                dist.all_reduce(l)
            losses.append(float(l))
        ms = timed_max_over_ranks(lambda: eng.step(xx), 50, 5, world, device)
        l = eng.step(xx).clone()
        if world > 1:
            dist.all_reduce(l)
        res[name] = {"ms_per_step": ms, "adam_steps_per_s": 1e3 / ms, "loss_first": losses[0], "loss_third": losses[2], "loss_last": float(l),
                     "cuda_graph": bool(eng._graph),
                     "exchange": "none (1 rank)" if world == 1 else ("NVLink peer memory, fused into the optimiser kernel (thb_adam_rows_peer)"
                                                                     if eng.peer is not None else "NCCL all-reduce of the compact rows")}
        if eng.peer is not None:
            res[name]["peer_wait_timeouts"] = int(eng.peer.err.item())
        elif world > 1 and peer is None:
            res[name]["peer_unavailable"] = getattr(eng, "peer_error", None)
        if not use_graph:
            c, a, o = eng.time_phases(xx, iters=20)
            t = torch.tensor([c, a, o], device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tmin = torch.tensor([c, a, o], device=device)
            if world > 1:
                dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            # the rank that arrives last does not wait for anyone: its exchange phase is the cost of the exchange itself
            res["exchange_ms_min_over_ranks"] = float(tmin[2]) if eng.peer is not None else float(tmin[1])
            res["phases_ms"] = {"compute": float(t[0]), "allreduce": float(t[1]), "adam": float(t[2]),
                                "note": "eager launches; with peer memory the exchange is inside the `adam` kernel pair (flag wait + peer reads + Adam); "
                                        "times include waiting for the slowest rank"}
            cnt = torch.tensor([float(prm.shape[0])], device=device)
            allc = [torch.zeros_like(cnt) for _ in range(world)]
            if world > 1:
                dist.all_gather(allc, cnt)
            else:
                allc = [cnt]
            res["points_per_rank"] = [int(v) for v in allc]
        if name == "graph_nccl":
            c, a, o = eng.time_phases(xx, iters=20)
            tmin = torch.tensor([a], device=device)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            res["nccl_allreduce_ms_min_over_ranks"] = float(tmin[0])
        del eng
    best = min(res[k]["ms_per_step"] for k, _, _ in variants)
    res.update({"adam_steps_per_s": 1e3 / best, "ms_per_step": best, "points": n_total, "n_active_rows": P.n_active,
                "allreduce_bytes": P.n_active * 3 * 4, "dense_gradient_bytes": P.nCP * 3 * 4,
                "what": "10M random points moved once to the rank owning their cell (all-to-all), plan built once; step = nets + evaluation + "
                        "one-pass MSE + backward into compact active rows + all-reduce of those rows + Adam on them"})
    return res


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path on the host cores, on a bounded sample of the same workload.
    The reference's Python basis code cannot hold this space (its dense tables would need ~1e11 entries), so
    PHI comes from the plain-C oracle port (all host threads) and the PHI.ctrl_pts contraction runs through
    the reference's own compiled CPU extension (oracle/_ref = THB_extensions/compute_pts.cpp, unmodified)
    when it is present, else through the C restatement of it."""
    if rank != 0:
        return
    import torch

    from oracle import build_ref, c_oracle

    sp, P, params, cp = build_workload(0, max(world, 1), "cpu")
    n = params.shape[0]
    sample = params[:: n // SAMPLE_REF][:SAMPLE_REF].contiguous().numpy()
    cph = cp.numpy()
    ht = host_tables(P)
    c_oracle.use_all_host_threads()
    torch.set_num_threads(c_oracle.num_threads())
    ext = build_ref.load()
    kind = "reference" if ext is not None else "port"

    def step():
        r = ht.eval(sample, None)
        S = P.cellinfo_host[np.maximum(r["slot"], 0), 5] * (r["slot"] >= 0)
        cs = np.cumsum(S).astype(np.int64)
        off = cs - S
        r = ht.eval(sample, None, want_phi=True, offsets=off, nnz=int(cs[-1]))
        info = P.cellinfo_host
        jm_all = P.supp_jm.numpy()
        Jm = np.concatenate([jm_all[info[s, 4]: info[s, 4] + info[s, 5]] for s in r["slot"] if s >= 0]).astype(np.int64)
        phi = r["phi"].astype(np.float32)
        if ext is not None:
            out = ext.cpp_forward(torch.from_numpy(cph), torch.from_numpy(Jm), torch.from_numpy(phi), torch.from_numpy(cs))
        else:
            out = c_oracle.eval_forward(cph, Jm, phi, cs)
        return out

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = len(sample) / dt
    port_rate, threads, _ = cpu_port_rate(P, params.numpy(), cph, target_s=5.0)
    line = {
        "impl": "reference", "metric": "THB basis+eval points/sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: 3-D degree-3 THB, 4 levels, 256^3 grid per GPU", "sample": len(sample),
                   "same_config": False, "same_config_note": "same space and grid, but a 2048-point strided sample per step and PHI by the C port: "
                                                             "the reference's own Python basis code cannot hold this space (dense tables ~1e11 entries)"},
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": threads, "kind": kind,
                         "sample": f"{len(sample)} strided points of the 256^3 grid per step; PHI by the C port on {threads} threads, "
                                   f"contraction by {'the reference compute_pts.cpp (single thread)' if ext is not None else 'the C restatement'}; "
                                   f"C port alone (basis+eval, all threads): {port_rate:.3e} points/s"},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dense", action="store_true", help="store one-hot tiles explicitly (SURVEY 8d's dense operation count)")
    ap.add_argument("--formulation", default="local_net", choices=["local_net", "per_point", "tile_net"],
                    help="local_net (default): supports folded into a per-cell control net; per_point: PHI of every support per point")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--inner", type=int, default=100, help="identical passes of the hot path per timed step (timed region of about a second)")
    ap.add_argument("--no-overlap", action="store_true", help="plan, nets and evaluation one after the other on one stream")
    ap.add_argument("--no-blocks", action="store_true", help="skip the strong-scaling / config 4 / config 5 blocks")
    ap.add_argument("--cell-cost", type=float, default=100.0, help="cost of a touched cell in point evaluations (slab / cell balancing)")
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--e2e-ramp", type=int, default=0, help="number of doubling chunk sizes at each end of the e2e pipeline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from thb_diff_b200 import _lib, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_host_to_gpu(local)   # before any pinned allocation: host buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    sp, P, params, cp = build_workload(rank, world, device, dense_tables=args.dense)
    n = params.shape[0]
    plan = engine.Plan(P, params)
    n_in, nnz, nnz_tiled, nnz_sep = plan.stats(with_sep=True)
    if args.dense:
        nnz_sep = 0
        args.formulation = "per_point"
    out = [torch.zeros((n, 3), dtype=torch.float32, device=device)]
    st = torch.cuda.current_stream()

    mono = engine._net_flags(P, False) & engine.NETS_MONOMIAL != 0

    def step(timers=None):
        """One pass.  Timed region: the plan on the current stream and, beside it on a second stream, the local control nets
        (they depend on the control points only), then the evaluation kernel -- engine.plan_and_nets(overlap=True).
        With timers: the same three pieces one after the other on one stream, each between two events (breakdown pass)."""
        if args.formulation != "local_net":
            if timers is not None:
                timers[0].record()
            plan.build()
            if timers is not None:
                timers[1].record()
                timers[3].record()
            engine.eval_fused(P, plan, cp, formulation=args.formulation, out=out)
            if timers is not None:
                timers[2].record()
            return
        if timers is None:
            nets = engine.plan_and_nets(P, plan, cp, overlap=not args.no_overlap)
            engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
            return
        timers[0].record()
        plan.build()
        timers[1].record()
        nets = engine.cell_nets(P, cp, plan=plan)
        timers[3].record()
        engine.eval_nets(P, plan, nets, 3, out=out, with_derivatives=False)
        timers[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    fp32_peak = max(engine.fma_peak(0), engine.fma_peak(1))  # TFLOP/s, measured on this GPU now

    inner = max(1, args.inner)
    for _ in range(max(args.warmup, 3) * min(inner, 10)):
        step()
    clk = ClockSampler(local)
    clk.__enter__()  # samples during the timed region only (steps x inner passes, about a second)
    barrier()
    launches0 = _lib.launch_count()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for k in range(args.steps * inner):
        step()
    t_end.record()
    barrier()
    clk.__exit__()
    launches = _lib.launch_count() - launches0
    total_ms = t_beg.elapsed_time(t_end)
    # breakdown pass (not the headline): the three pieces serialised on one stream, CUDA events around each
    nb = max(20, min(100, args.steps * inner // 10))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nb)]
    for k in range(nb):
        step(ev[k])
    torch.cuda.synchronize()
    plan_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / nb
    nets_ms = sum(e[1].elapsed_time(e[3]) for e in ev) / nb   # local control nets of the cells that hold points
    kern_ms = sum(e[3].elapsed_time(e[2]) for e in ev) / nb   # the evaluation kernel
    del ev
    tmax = torch.tensor([total_ms], device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    ms_per_pass = ms_per_step / inner
    value = n * world * inner / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): H2D params, plan, fused eval, D2H geometry
    params_host = params.cpu().pin_memory()
    out_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    pipe = engine.HostPipeline(P, n, cp_dim=3, n_chunks=args.e2e_chunks, formulation=args.formulation, ramp=args.e2e_ramp)

    def e2e_step():
        # public host-buffer API: chunked H2D -> plan + fused kernel -> D2H, overlapped on three streams
        pipe.run(params_host, cp, out_host)

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ksteps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(ksteps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / ksteps], device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n * world / (float(e2e_ms.item()) * 1e-3)

    def copy_floor():
        """The two host<->device copies of one e2e step alone, on two streams (PCIe is full duplex): the floor of e2e."""
        d_in = torch.empty_like(params)
        s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        res = {}
        for name, both in (("h2d", False), ("d2h", False), ("duplex", True)):
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                s1.wait_stream(torch.cuda.current_stream())
                s2.wait_stream(torch.cuda.current_stream())
                if both or name == "h2d":
                    with torch.cuda.stream(s1):
                        d_in.copy_(params_host, non_blocking=True)
                if both or name == "d2h":
                    with torch.cuda.stream(s2):
                        out_host.copy_(out[0], non_blocking=True)
                torch.cuda.current_stream().wait_stream(s1)
                torch.cuda.current_stream().wait_stream(s2)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            res[name + "_ms"] = best
        res["h2d_gbs"] = n * P.ndim * 4 / res["h2d_ms"] / 1e6
        res["d2h_gbs"] = n * 3 * 4 / res["d2h_ms"] / 1e6
        return res

    copies = copy_floor()
    step()
    torch.cuda.synchronize()
    e2e_diff = float((out_host.to(device) - out[0]).abs().max())  # the pipelined host path must reproduce the device path

    # ---- roofline of the dominant kernel (fused tile kernel), from the live CUDA-event timing above
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    f_exec, f_dense = flop_model(P.degrees, 3, n_in, nnz, nnz_tiled, nnz_sep)
    f_nets = 0.0
    if args.formulation == "local_net":
        f_exec, f_nets = flop_model_local_net(P, plan, 3, n_in, mono=mono)
    elif args.formulation == "tile_net":  # per point the tensor product and the window dot products; all tiles folded per item
        cnt = plan.cell_count[: P.nslots].long()
        nt = (P.cellinfo[:, 5] - P.cellinfo[:, 7]).long()
        ppi = int(plan.desc.points_per_item)
        items = (cnt + ppi - 1) // ppi
        f_exec = flop_model(P.degrees, 3, n_in, 0, 0)[0] + float((items * nt).sum()) * P.T * 2 * 4
    tile_bytes = (P.sep_vec.numel() + P.full_tiles.numel()) * 4 if (getattr(P, "sep_vec", None) is not None and args.formulation != "tile_net") else P.tiles.numel() * 4
    if args.formulation == "local_net":  # evaluation kernel: 16-byte point records in, geometry out, one net per work item
        alg_bytes = n * (16 + 4 * 3) + int(plan.counters[0].item()) * 4 * P.TS * 4
    else:
        alg_bytes = n * (16 + 4 * 3) + tile_bytes + cp.numel() * 4
    t_fp32 = f_exec / (fp32_peak * 1e12)
    t_hbm = alg_bytes / (hbm_peak * 1e9)
    t_roof = max(t_fp32, t_hbm)
    bound = "fp32" if t_fp32 >= t_hbm else "hbm"
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of this kernel
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_net_eval_c3_bytes" if args.formulation == "local_net" else "k_tile_fused_c3_bytes"]
    except Exception:
        pass
    roofline = {
        "bound": bound, "kernel": ("thb_tile::k_net_eval<4,4,4,4,3,0,MONO>" if mono else "thb_tile::k_net_eval<4,4,4,2,3,0>") if args.formulation == "local_net" else "thb_tile::k_tile<4,4,4,2,3,FUSED>",
        "achieved": (f_exec / (kern_ms * 1e-3) / 1e12) if bound == "fp32" else (alg_bytes / (kern_ms * 1e-3) / 1e9),
        "peak": fp32_peak if bound == "fp32" else hbm_peak, "unit": "TFLOP/s" if bound == "fp32" else "GB/s",
        "frac": t_roof / (kern_ms * 1e-3), "traffic": traffic,
        "kernel_ms": kern_ms, "plan_ms": plan_ms, "nets_ms": nets_ms, "kernel_share_of_step": kern_ms / (plan_ms + nets_ms + kern_ms),
        "nets_flops": f_nets,
        "flops_executed": f_exec, "flops_dense_model_8d": f_dense, "algorithmic_bytes": alg_bytes,
        "frac_vs_dense_model_8d": (f_dense / (fp32_peak * 1e12)) / (kern_ms * 1e-3),
        "frac_fp32": t_fp32 / (kern_ms * 1e-3), "frac_hbm": t_hbm / (kern_ms * 1e-3),
        "peak_source": "FP32: thb_fma_peak micro-kernel measured in this run (MEASURED_PEAKS.json has no FP32 entry); "
                       f"HBM: {'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'}",
    }
    # the whole step (plan + nets + evaluation) against the same two roofs: executed FLOPs of nets + evaluation, and the
    # algorithmic bytes of every pass over the points (count: params in, slot out; scatter: params + slot in, 16-byte
    # records out; evaluation: records in, geometry out) plus the nets written once and read once per work item
    n_items = int(plan.counters[0].item())
    n_cells_used = int((plan.cell_count[: P.nslots] > 0).sum())
    step_bytes = n * ((4 * P.ndim + 4) + (4 * P.ndim + 4 + 16) + (16 + 4 * 3)) + (n_cells_used + n_items) * 4 * P.TS * 4
    if args.formulation == "local_net" and mono:
        # the same kernel time against the FLOPs of the control-point form it replaces (Cox-de Boor + sum-factorised
        # contraction with tp(u), 612 per point for cubic 3-D: round 1's kernel): what the monomial form saves is work, not speed
        f_cp = flop_model_local_net(P, plan, 3, n_in, mono=False)[0]
        roofline["frac_fp32_vs_control_point_form_flops"] = (f_cp / (fp32_peak * 1e12)) / (kern_ms * 1e-3)
        roofline["note"] = "monomial nets: FP32 and HBM roofs of this kernel coincide (0.088 ms each on C3); frac = the slower roof / measured time"
    roofline["step_flops"] = f_exec + f_nets
    roofline["step_bytes"] = step_bytes
    roofline["step_frac"] = ((f_exec + f_nets) / (fp32_peak * 1e12)) / (ms_per_pass * 1e-3)
    roofline["hbm_step_frac"] = (step_bytes / (hbm_peak * 1e9)) / (ms_per_pass * 1e-3)
    roofline["ms_per_pass"] = ms_per_pass

    line = {
        "metric": "THB basis+eval points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: 3-D degree-3 THB, 4 active levels (~25% refined per level, graded), 256^3 param grid per GPU",
                   "points_per_gpu": n, "nnz": nnz, "nnz_tiled": nnz_tiled, "nnz_tiled_rank_one": nnz_sep, "mean_supports": nnz / max(n_in, 1), "nCP": P.nCP,
                   "table_bytes": P.nbytes(), "l2": "inputs (201 MB params + 201 MB output per pass) exceed the 126 MB L2",
                   "passes_per_step": inner, "ms_per_pass": ms_per_pass,
                   "tables": "dense one-hot tiles" if args.dense else "own-level supports direct, rank-one tiles as 3 factor vectors, full tiles otherwise", "formulation": args.formulation,
                   "nets": ("monomial form, nested Horner evaluation" if mono else "control points, Cox-de Boor + sum factorisation") if args.formulation == "local_net" else None,
                   "step": "plan || nets (second stream), then evaluation" if (args.formulation == "local_net" and not args.no_overlap) else "plan, nets, evaluation on one stream",
                   "sharding": f"y planes round-robin over {world} rank(s), no collective"},
        "clocks": clk.summary(), "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": n * P.ndim * 4, "d2h_bytes_per_step": n * 3 * 4,
                "ms_per_step": float(e2e_ms.item()), "passes_per_step": 1,
                "how": f"engine.HostPipeline: {len(pipe.bounds)} chunks, H2D / compute / D2H overlapped on 3 streams, pinned host buffers, "
                       f"{'one CUDA graph replay per step' if pipe._graph is not None else 'eager launches'}",
                "check_max_abs_diff_vs_device_path": e2e_diff,
                "host_copies_alone": copies, "host_cpus_bound_to_gpu_numa_node": numa},
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_variants:
        # the other two formulations of the same evaluation, for context (same plan, same inputs)
        def time_variant(Pv, form):
            pl = engine.Plan(Pv, params)
            o = [torch.zeros((n, 3), dtype=torch.float32, device=device)]
            for _ in range(3):
                engine.eval_fused(Pv, pl, cp, formulation=form, out=o)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                engine.eval_fused(Pv, pl, cp, formulation=form, out=o)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / 5, float((o[0] - out[0]).abs().max())

        var = {}
        if not args.dense and getattr(P, "sep_vec", None) is not None:
            keep = (P.sep_vec, P._desc_cache)
            P.sep_vec, P._desc_cache = None, None      # same tables without the rank-one split (every tiled support a full tile)
            ms, dev = time_variant(P, "per_point")
            f_full = flop_model(P.degrees, 3, n_in, nnz, nnz_tiled, 0)[0]
            var["all_full_tiles"] = {"kernel_ms": ms, "max_abs_diff_vs_headline": dev, "frac_of_fp32_roofline": (f_full / (fp32_peak * 1e12)) / (ms * 1e-3)}
            P.sep_vec, P._desc_cache = keep[0], None
        for form in ("local_net", "per_point", "tile_net"):
            if form != args.formulation and not args.dense:
                ms, dev = time_variant(P, form)
                var[form] = {"kernel_ms": ms, "max_abs_diff_vs_headline": dev}
                if form == "per_point":  # PHI of every support formed per point (the reference's gather-multiply, fused)
                    f_pp = flop_model(P.degrees, 3, n_in, nnz, nnz_tiled, nnz_sep)[0]
                    var[form]["frac_of_fp32_roofline"] = (f_pp / (fp32_peak * 1e12)) / (ms * 1e-3)
        if not args.dense:
            from thb_diff_b200.packed import PackedHierarchy

            Pd = PackedHierarchy.from_space(sp, device=device, onehot_direct=False)
            Pd.sep_vec, Pd._desc_cache = None, None   # literally dense: every support a full (p+1)^d tile dot
            ms, dev = time_variant(Pd, "per_point")
            var["dense_one_hot_tiles"] = {"kernel_ms": ms, "max_abs_diff_vs_headline": dev,
                                          "frac_of_fp32_roofline_8d": (f_dense / (fp32_peak * 1e12)) / (ms * 1e-3)}
            del Pd
        line["variants"] = var

    if rank == 0 and world == 1 and not args.no_extras:
        try:
            line["extras"] = ex = extras(sp, P, params, cp, hbm_peak)
            # the driver keeps flat scalars of `roofline`: the API-faithful (PHI-materialising) kernels and the comparison
            # with the reference's own CUDA kernels compiled for sm_100a go there too
            r = line["roofline"]
            r["basis_rows_ms"], r["basis_rows_hbm_frac"] = ex["THB_basis_fns"]["ms"], ex["THB_basis_fns"]["hbm_frac"]
            for k in ("int64", "int32"):
                e = ex[f"THB_eval_{k}"]
                r[f"csr_fwd_{k}_ms"], r[f"csr_fwd_{k}_hbm_frac"] = e["forward_ms"], e["forward_hbm_frac"]
                r[f"csr_bwd_{k}_ms"], r[f"csr_bwd_{k}_hbm_frac"] = e["backward_ms"], e["backward_hbm_frac"]
            if "reference_cuda_kernels_sm100a" in ex:
                rc = ex["reference_cuda_kernels_sm100a"]
                r["ref_cuda_fwd_ms"], r["ref_cuda_bwd_ms"] = rc["forward_ms"], rc["backward_ms"]
                r["speedup_vs_ref_cuda_fwd"], r["speedup_vs_ref_cuda_bwd"] = rc["speedup_forward"], rc["speedup_backward"]
                if "speedup_forward_int32_indices" in rc:
                    r["speedup_vs_ref_cuda_fwd_i32"], r["speedup_vs_ref_cuda_bwd_i32"] = rc["speedup_forward_int32_indices"], rc["speedup_backward_int32_indices"]
            r["fused_backward_ms"] = ex["fused_backward"]["ms"]
        except Exception as e:  # secondary numbers must never take the headline down -- but they must be seen
            import traceback

            line["extras"] = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
            line["roofline"]["extras_error"] = repr(e)[:120]
    if not args.no_blocks:
        # every N: strong scaling of the fixed grid, config 4 (fit with the gradient all-reduce), config 5 (5 levels, derivatives)
        blocks = {}
        torch.cuda.empty_cache()
        for name, fn in (("grid", lambda: grid_block(P, cp, rank, world, device)),
                         ("strong", lambda: strong_block(P, cp, rank, world, device, args.cell_cost)),
                         ("fit_c4", lambda: fit_block(sp, P, rank, world, device, args.cell_cost)),
                         ("c5", lambda: c5_block(rank, world, device, args.cell_cost))):
            try:
                blocks[name] = fn()
            except Exception as e:
                import traceback

                blocks[name] = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
                line["config"][name + "_error"] = repr(e)[:120]
            torch.cuda.empty_cache()
        line["blocks"] = blocks
        c = line["config"]
        b = blocks.get("grid", {})
        if "points_per_s" in b:
            c["grid_points_per_s"], c["grid_ms_per_step"] = b["points_per_s"], b["ms_per_step"]
            c["grid_what"] = "same points evaluated as a tensor-product grid (thb_eval_grid): no plan, cell-level sum factorisation"
        b = blocks.get("strong", {})
        if "points_per_s" in b:
            c["strong_points_per_s"], c["strong_ms_per_step"] = b["points_per_s"], b["ms_per_step"]
            c["strong_what"] = "fixed 256^3 grid in cost-balanced x slabs over the ranks, plan+nets+eval, one CUDA graph per step"
        b = blocks.get("fit_c4", {})
        if "adam_steps_per_s" in b:
            c["fit_c4_adam_steps_per_s"], c["fit_c4_ms_per_step"] = b["adam_steps_per_s"], b["ms_per_step"]
            c["fit_c4_allreduce_ms"], c["fit_c4_compute_ms"], c["fit_c4_adam_ms"] = (b["phases_ms"][k] for k in ("allreduce", "compute", "adam"))
            c["fit_c4_allreduce_bytes"], c["fit_c4_loss_first"], c["fit_c4_loss_last"] = b["allreduce_bytes"], b["eager"]["loss_first"], b["eager"]["loss_last"]
            c["fit_c4_collective"] = b["graph"]["exchange"]
            if "graph_nccl" in b:
                c["fit_c4_ms_per_step_nccl"] = b["graph_nccl"]["ms_per_step"]
                c["fit_c4_exchange_ms"], c["fit_c4_nccl_allreduce_ms"] = b.get("exchange_ms_min_over_ranks"), b.get("nccl_allreduce_ms_min_over_ranks")
        b = blocks.get("c5", {})
        if "points_per_s" in b:
            c["c5_points_per_s"], c["c5_ms_per_step"], c["c5_jacobian_max_abs_err"] = b["points_per_s"], b["ms_per_step"], b["jacobian_max_abs_err"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, threads, m = cpu_port_rate(P, params_host.numpy(), cp.cpu().numpy())
        line["cpu_baseline"] = {"value": rate, "unit": "points/s", "cores": threads, "kind": "port",
                                "sample": f"{m} strided points of the 256^3 grid, plain-C float64 oracle (oracle/thb_oracle_c.c, OpenMP)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
