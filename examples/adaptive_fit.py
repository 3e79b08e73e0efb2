"""Fit -> refine -> prolong control points -> refit, entirely on the fused GPU path.

    python examples/adaptive_fit.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200 import configs  # noqa: E402
from thb_diff_b200.multilevel_spline_space import BSpline, ControlPoints, Space, TensorProduct  # noqa: E402
from thb_diff_b200.torch_funcs import THB_fused_module  # noqa: E402


def target(p):
    return torch.stack([p[:, 0], p[:, 1], 0.2 * torch.exp(-60 * ((p[:, 0] - 0.5) ** 2 + (p[:, 1] - 0.5) ** 2))], dim=1)


def fit(space, cp, params, tgt, steps):
    mod = THB_fused_module(space, params, cellnet=True)
    x = cp.clone().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=5e-3, fused=True)
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = ((mod(x) - tgt) ** 2).mean()
        loss.backward()
        opt.step()
    return x.detach(), float(loss.detach())


def main(steps=200, n=200_000, verbose=True):
    sp = Space(TensorProduct([BSpline(configs.README_U0, 2), BSpline(configs.README_U1, 3)]), 4)
    sp.build_hierarchy_from_domain_sequence()
    cps = ControlPoints(sp)
    params = torch.rand((n, 2), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    tgt = target(params)
    cp = torch.from_numpy(configs.greville_ctrl_pts(sp, bump=0.0)).cuda()
    losses = []
    for rnd, box in enumerate([None, (0, (1, 0), (9, 7)), (1, (6, 4), (14, 10))]):
        if box is not None:
            lev, lo, hi = box
            sp.refine_box(lo, hi, lev)
            sp.build_hierarchy_from_domain_sequence()
            c = cps.update_CP(cp.cpu().numpy()[:, :2], sp.fns, sp.Coeff)   # x, y columns: prolongation is per column
            z = cps_z.update_CP(np.repeat(cp.cpu().numpy()[:, 2:3], 2, axis=1), sp.fns, sp.Coeff)
            cp = torch.from_numpy(np.concatenate([np.concatenate([c[l].reshape(-1, 2), z[l].reshape(-1, 2)[:, :1]], axis=1)
                                                  for l in range(4)]).astype(np.float32)).cuda()
        else:
            cps_z = ControlPoints(sp)
        cp, loss = fit(sp, cp, params, tgt, steps)
        losses.append(loss)
        if verbose:
            print(f"round {rnd}: active functions { {l: int(sp.fns[l].sum()) for l in sp.fns} }  loss {loss:.3e}")
    return losses


if __name__ == "__main__":
    main()
