"""The reference README's call sequence (README.md:58-111), with the imports switched to thb_diff_b200.

    python examples/readme_pipeline.py            # needs a B200 (no CPU fallback)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thb_diff_b200.bspline_funcs import generate_parametric_coordinates, grevilleAbscissae  # noqa: E402
from thb_diff_b200.core import (THB_basis_fns, compute_active_cells_active_supp, compute_active_span,  # noqa: E402
                                compute_refinement_operators)
from thb_diff_b200.multilevel_spline_space import BSpline, Space, TensorProduct  # noqa: E402
from thb_diff_b200.torch_funcs import THB_nn_module, prepare_data_for_CUDA_evaluation  # noqa: E402

bs1 = BSpline(knotvector=np.array([0, 0, 0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1, 1, 1]), degree=2)
bs2 = BSpline(knotvector=np.array([0, 0, 0, 0, 0.2, 0.4, 0.5, 0.6, 0.8, 0.9, 1, 1, 1, 1]), degree=3)
bs3 = BSpline(knotvector=np.array([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1, 1, 1]), degree=2)
h_space = Space(tensor_product=TensorProduct([bs1, bs2, bs3]), num_levels=3)
h_space.build_hierarchy_from_domain_sequence()
h_space._refine_cell(cellIdx=(2, 2, 2), level=0)
h_space._refine_cell(cellIdx=(3, 1, 0), level=0)
h_space.build_hierarchy_from_domain_sequence()

params = generate_parametric_coordinates((50, 50, 50))
ac_cells = compute_active_cells_active_supp(h_space.cells, h_space.fns, h_space.degrees)
fn_coeffs = compute_refinement_operators(h_space.fns, h_space.Coeff, h_space.degrees)
ac_cell_supp, num_supp = compute_active_span(params, h_space.knotvectors, h_space.cells, h_space.degrees, ac_cells)
PHI = THB_basis_fns(params, ac_cell_supp, num_supp, fn_coeffs, h_space.sh_fns, h_space.knotvectors, h_space.degrees)
print("points", len(num_supp), "nnz", PHI.shape[0], "supports per point", set(np.asarray(num_supp).tolist()))

ctrl_pts = {lev: grevilleAbscissae(h_space.sh_fns[lev], h_space.degrees, h_space.knotvectors[lev]) for lev in range(3)}
cp, Jm, phi, cumsum, device = prepare_data_for_CUDA_evaluation(PHI, ac_cell_supp, num_supp, ctrl_pts, ac_cells, h_space.sh_fns, "cuda")
cp.requires_grad_(True)
pts = THB_nn_module(Jm, phi, cumsum, device)(cp)
print("max |evaluated - params| with Greville control points:", float((pts.detach().cpu() - torch.from_numpy(params).float()).abs().max()))
pts.square().mean().backward()
print("grad w.r.t. control points:", tuple(cp.grad.shape))

# the adaptive grid as a .vtu file (reference: utils.plot3DAdaptiveGrid) and the Bezier extraction operators of a few points
from thb_diff_b200.bezier import compute_multilevel_bezier_extraction_operators  # noqa: E402
from thb_diff_b200.utils import plot3DAdaptiveGrid  # noqa: E402

path, vtu_pts, vtu_conn = plot3DAdaptiveGrid("/tmp/thb_adaptive_grid", ac_cells, ctrl_pts, h_space.knotvectors, fn_coeffs,
                                             h_space.sh_fns, h_space.degrees)
print("wrote", path, "with", len(vtu_conn), "hexahedra and", len(vtu_pts), "nodes")
few, few_ns = compute_active_span(params[:5], h_space.knotvectors, h_space.cells, h_space.degrees, ac_cells)
ops = compute_multilevel_bezier_extraction_operators(params[:5], few, ac_cells, fn_coeffs, h_space.sh_cells, h_space.sh_fns,
                                                     h_space.knotvectors, h_space.degrees)
print("Bezier operators of point 0:", ops[0].shape)
