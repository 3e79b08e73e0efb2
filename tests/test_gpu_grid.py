"""thb_eval_grid (tensor-product grids of parameters: sum factorisation at the cell level, no plan) against the general
path on the same points and against the oracle on the golden grids of the reference's generate_parametric_coordinates."""
import numpy as np
import pytest
import torch

from helpers import load_golden, product_space_from_golden, rel_err
from oracle import thb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["readme2d", "graded2d", "block3d", "readme3d"])
def test_grid_equals_general_path_and_oracle(name):
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden(name)
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    d = h.ndim
    shape = (23, 17, 29)[:d]
    axes = [np.linspace(1e-5, 1.0, n, endpoint=False).astype(np.float32) for n in shape]
    pts = B.generate_parametric_coordinates(shape).astype(np.float32)          # the reference's 'xy' order
    cp = torch.from_numpy(np.random.default_rng(0).standard_normal((P.nCP, 3)).astype(np.float32)).cuda()
    want = engine.eval_fused(P, engine.Plan(P, pts), cp)[0]
    gp = engine.GridPlan(P, axes)
    got = engine.eval_grid(P, gp, cp)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= TOL * float(want.abs().max())
    # C order
    pts_ij = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, d)
    want_ij = engine.eval_fused(P, engine.Plan(P, pts_ij), cp)[0]
    got_ij = engine.eval_grid(P, engine.GridPlan(P, axes, order="ij"), cp)
    assert float((got_ij - want_ij).abs().max()) <= TOL * float(want.abs().max())
    # the golden parameter set of the fixture IS a grid of the reference: oracle contraction
    if "grid_shape" in g:
        gs = tuple(int(x) for x in g["grid_shape"])
        ax_g = [np.linspace(1e-5, 1.0, n, endpoint=False).astype(np.float32) for n in gs]
        ref = O.eval_forward(cp.cpu().numpy(), g["Jm"], g["PHI"], np.cumsum(g["num_supp"]))
        out = engine.eval_grid(P, engine.GridPlan(P, ax_g), cp).cpu().numpy()
        assert rel_err(out, ref, scale=np.abs(ref).max()).max() <= TOL


def test_grid_with_values_outside_the_domain_and_on_breakpoints():
    from thb_diff_b200 import engine
    from thb_diff_b200.packed import PackedHierarchy

    g = load_golden("block3d")
    h = product_space_from_golden(g)
    P = PackedHierarchy.from_space(h, device="cuda")
    U = [np.unique(h.knotvectors[max(h.knotvectors)][k]).astype(np.float32) for k in range(3)]
    axes = [np.unique(np.concatenate([[-0.25], U[k], np.nextafter(U[k], np.float32(2)), np.linspace(0.01, 0.99, 7).astype(np.float32), [1.5]])).astype(np.float32)
            for k in range(3)]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    cp = torch.from_numpy(np.random.default_rng(1).standard_normal((P.nCP, 3)).astype(np.float32)).cuda()
    plan = engine.Plan(P, pts)
    want = engine.eval_fused(P, plan, cp)[0]                      # zero rows for the points outside
    gp = engine.GridPlan(P, axes, order="ij")
    assert gp.n_inside == plan.stats()[0] < gp.n_points
    got = engine.eval_grid(P, gp, cp)
    assert float((got - want).abs().max()) <= TOL * float(want.abs().max())


def test_grid_c3_full_size():
    """BASELINE config 3: 256^3 grid, 4 levels -- Greville reproduction, partition of unity, equality with the general path."""
    from thb_diff_b200 import bspline_funcs as B
    from thb_diff_b200 import configs, engine
    from thb_diff_b200.packed import PackedHierarchy

    sp = configs.cubic3d_space(active_levels=4)
    P = PackedHierarchy.from_space(sp, device="cuda")
    ax = [np.linspace(1e-5, 1.0, 256, endpoint=False).astype(np.float32)] * 3
    gp = engine.GridPlan(P, ax)
    prm = B.grid_parametric_coordinates_device((256, 256, 256))
    grev = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    out = engine.eval_grid(P, gp, grev)
    assert float((out - prm).abs().max()) < 1e-5
    pou = engine.eval_grid(P, gp, torch.ones((P.nCP, 3), device="cuda"))
    assert float((pou - 1).abs().max()) < 1e-5
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp)).cuda()
    a = engine.eval_grid(P, gp, cp)
    b = engine.eval_fused(P, engine.Plan(P, prm), cp)[0]
    assert float((a - b).abs().max()) <= 2e-6
