"""world_size=2 on CPU (gloo): sharding covers every point exactly once and the gradient all-reduce used
by the fitting loop sums the per-rank partial gradients.  (The kernels themselves need a GPU; here the
per-rank partial gradients come from the oracle.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import load_golden
    from oracle import thb_oracle as O
    from thb_diff_b200.distributed import allreduce_grad_, shard_points

    g = load_golden("block3d")
    n = len(g["num_supp"])
    ids = shard_points(torch.arange(n), rank, world, mode).numpy()
    cs = np.cumsum(g["num_supp"])
    starts = cs - g["num_supp"]
    nCP = int(g["Jm"].max()) + 1
    go = np.random.default_rng(0).standard_normal((n, 3))
    # this rank's partial gradient from its own points only
    sel = np.concatenate([np.arange(starts[i], cs[i]) for i in ids])
    part = O.eval_backward(go[ids], (nCP, 3), g["Jm"][sel], g["PHI"][sel], np.cumsum(g["num_supp"][ids]))
    t = torch.from_numpy(part)
    allreduce_grad_(t)
    full = O.eval_backward(go, (nCP, 3), g["Jm"], g["PHI"], cs)
    ret[rank] = (float(np.abs(t.numpy() - full).max()), ids.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["contiguous", "strided"])
def test_sharded_gradient_allreduce(mode):
    world = 2
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "contiguous" else 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    ids = sorted(ret[0][1] + ret[1][1])
    assert ids == list(range(len(ids)))           # every point exactly once
    assert max(ret[0][0], ret[1][0]) < 1e-12      # all-reduced partial gradients == full gradient


def test_shard_bounds():
    from thb_diff_b200.distributed import shard_bounds

    for n, w in ((10, 3), (7, 8), (16777216, 8), (0, 2)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_three_pass_loss_matches_the_composed_expression():
    """distributed._ScaledSumSquaredError (the fit loop's loss) == scale * ((out - target) ** 2).sum(), value and gradient."""
    from thb_diff_b200.distributed import _ScaledSumSquaredError

    torch.manual_seed(0)
    out = torch.randn(1000, 3, dtype=torch.float64, requires_grad=True)
    tgt = torch.randn(1000, 3, dtype=torch.float64)
    scale = 1.0 / 3000
    a = _ScaledSumSquaredError.apply(out, tgt, scale)
    (ga,) = torch.autograd.grad(3.0 * a, out)
    out2 = out.detach().clone().requires_grad_(True)
    b = ((out2 - tgt) ** 2).sum() * scale
    (gb,) = torch.autograd.grad(3.0 * b, out2)
    assert abs(float(a) - float(b)) <= 1e-12 and float((ga - gb).abs().max()) <= 1e-14
