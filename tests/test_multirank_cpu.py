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


def _partition_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from thb_diff_b200.distributed import allreduce_active_rows_, partition_by_cell

    nslots = 37
    gen = torch.Generator().manual_seed(100 + rank)
    n = 500 + 77 * rank
    slot = torch.randint(-1, nslots, (n,), generator=gen, dtype=torch.int32)
    slot[slot == 5] = 6                       # an empty cell
    payload = torch.stack([slot.float(), torch.full((n,), float(rank)), torch.arange(n).float()], 1)
    recv, bounds = partition_by_cell(slot, nslots, payload)
    # dense gradient whose active rows differ per rank; rows outside the active set must not be touched
    grad = torch.zeros(20, 3)
    ids = torch.tensor([1, 4, 7, 19], dtype=torch.int32)
    grad[ids.long()] = float(rank + 1)
    grad[2] = 123.0 + rank                    # not an active row: stays local
    allreduce_active_rows_(grad, ids)
    ret[rank] = (recv.numpy(), bounds.numpy(), payload.numpy(), grad.numpy())
    dist.destroy_process_group()


def test_partition_by_cell_moves_every_point_to_the_owner_of_its_cell():
    world = 3
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_partition_worker, args=(world, port, ret), nprocs=world, join=True)
    bounds = ret[0][1]
    assert all(np.array_equal(ret[r][1], bounds) for r in range(world))
    assert bounds[0] == 0 and bounds[-1] == 37 and (np.diff(bounds) >= 0).all()
    sent = np.concatenate([ret[r][2] for r in range(world)])
    got = np.concatenate([ret[r][0] for r in range(world)])
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(sent), key(got))                      # nothing lost, nothing duplicated
    for r in range(world):
        s = ret[r][0][:, 0]
        inside = s >= 0
        assert ((s[inside] >= bounds[r]) & (s[inside] < bounds[r + 1])).all()
        if r:
            assert inside.all()                                     # points outside the domain go to rank 0
    # balance: cost = points + 100 per non-empty cell, contiguous ranges -> no rank above ~1.6x the mean here
    cnt = np.array([len(ret[r][0]) for r in range(world)])
    assert cnt.max() <= 1.6 * cnt.mean()
    for r in range(world):
        gr = ret[r][3]
        assert (gr[[1, 4, 7, 19]] == 6.0).all() and (gr[2] == 123.0 + r).all() and gr[[0, 3, 5]].sum() == 0


def test_cell_owner_bounds_balances_points_plus_cell_cost():
    from thb_diff_b200.distributed import cell_owner_bounds

    c = torch.tensor([0, 0, 1000, 10, 10, 10, 10, 0, 500, 500], dtype=torch.int64)
    b = cell_owner_bounds(c, 2, cell_cost=100.0)
    assert b.tolist()[0] == 0 and b.tolist()[-1] == 10
    cost = c.double() + (c > 0).double() * 100
    left = float(cost[: int(b[1])].sum())
    assert abs(left - float(cost.sum()) / 2) <= float(cost.max())
    assert cell_owner_bounds(c, 1).tolist() == [0, 10]
    assert cell_owner_bounds(torch.zeros(4, dtype=torch.int64), 3).tolist()[-1] == 4
