"""thb_diff_b200 -- B200 (sm_100a) implementation of THB-Diff's evaluation hot path behind the reference's
API.  Importing the compute modules requires the built CUDA library (python -m thb_diff_b200.build);
there is no CPU fallback."""
__all__ = ["core", "bspline_funcs", "multilevel_spline_space", "torch_funcs", "THB_eval", "engine", "packed"]
