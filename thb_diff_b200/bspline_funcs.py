"""1-D B-spline primitives with the reference's names [THB/bspline_funcs.py]; the per-point functions run
on the GPU (csrc/thb_spans.cu), the set-up helpers are host numpy."""
import numpy as np
import torch

from . import engine
from .multilevel_spline_space import assemble_Tmatrix, refine_knotvector  # noqa: F401  (same public names)


def generate_parametric_coordinates(shape):
    """[THB/bspline_funcs.py:20-35]: numpy 'xy' meshgrid order (3-D: z fastest, then x, y slowest),
    values linspace(1e-5, 1, n, endpoint=False); float64 like the reference."""
    axes = [np.linspace(1e-5, 1.0, int(n), endpoint=False) for n in shape]
    grids = np.meshgrid(*axes)
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def grid_parametric_coordinates_device(shape, device="cuda", dtype=torch.float32, first=0, count=None, y_first=0, y_step=1):
    """The same grid generated on the device from the linear index (bit-equal to the float64 numpy values
    rounded to float32).  `y_first`/`y_step` select every y_step-th plane of the slowest axis (sharding)."""
    d = len(shape)
    ax = [torch.from_numpy(np.linspace(1e-5, 1.0, int(n), endpoint=False)).to(device=device, dtype=dtype) for n in shape]
    if d == 2:
        x, y = ax
        y = y[y_first::y_step]
        return torch.stack([x.repeat(len(y)), y.repeat_interleave(len(x))], dim=1).contiguous()
    x, y, z = ax
    y = y[y_first::y_step]
    nx, ny, nz = len(x), len(y), len(z)
    X = x.view(1, nx, 1).expand(ny, nx, nz).reshape(-1)
    Y = y.view(ny, 1, 1).expand(ny, nx, nz).reshape(-1)
    Z = z.view(1, 1, nz).expand(ny, nx, nz).reshape(-1)
    return torch.stack([X, Y, Z], dim=1).contiguous()


def grevilleAbscissae(fn_sh, degrees, knotvectors):
    """[THB/bspline_funcs.py:38-51] Greville points as a control net [*fn_sh, d]."""
    d = len(fn_sh)
    out = np.zeros((*fn_sh, d))
    for k in range(d):
        U = np.asarray(knotvectors[k], dtype=np.float64)
        c = np.convolve(U[1:], np.ones(degrees[k]), mode="valid")[: fn_sh[k]] / degrees[k]
        shp = [1] * d
        shp[k] = fn_sh[k]
        out[..., k] = c.reshape(shp)
    return out


def _dev_f32(x, device="cuda"):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(device)


def find_span_array_jax(params, U, degree):
    """[THB/bspline_funcs.py:77-83] knot span per point (int32 CUDA tensor); float32 comparison like JAX."""
    return engine.find_spans(_dev_f32(params).reshape(-1), _dev_f32(U), int(degree))


find_span_array = find_span_array_jax


def basis_fns_vmap(params, knotvector, degree):
    """[THB/bspline_funcs.py:185-188] the degree+1 non-vanishing basis values per point, always [N, p+1]
    (the reference squeezes N == 1 away)."""
    return engine.basis_1d(_dev_f32(params).reshape(-1), _dev_f32(knotvector), int(degree))


def der_basis_fns_vmap(params, knotvector, degree):
    """[THB/bspline_funcs.py:191-194] first derivatives of the same functions, [N, p+1]."""
    return engine.basis_1d(_dev_f32(params).reshape(-1), _dev_f32(knotvector), int(degree), derivs=True)[1]


def findSpan(n, p, u, U):
    """[THB/bspline_funcs.py:62-74] scalar span; evaluated with the vector rule (identical in-domain)."""
    return int(find_span_array_jax(np.array([u]), U, p)[0])


def basisFun(i, u, p, U):
    """[THB/bspline_funcs.py:86-101] scalar Cox-de Boor at the span of u."""
    return basis_fns_vmap(np.array([u]), U, p)[0].cpu().numpy()


def compute_tensor_product(args):
    """[THB/bspline_funcs.py:54-59] outer product of 2 or 3 vectors ("i,j->ij" / "i,j,k->ijk"), on the GPU
    (thb_tensor_product); [B, n] batches of vectors give [B, ...] batches of products.  numpy / CPU inputs are staged
    through the device and come back on the CPU."""
    a = [torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in args]
    src = a[0].device
    out = engine.tensor_product([x.to(device="cuda", dtype=torch.float32) for x in a])
    return out if src.type == "cuda" else out.to(src)


def basisFun_jax(param, knotvector, degree):
    """[THB/bspline_funcs.py:149-182] the degree+1 non-vanishing basis values of one parameter, shape [1, p+1]."""
    return basis_fns_vmap(np.atleast_1d(np.asarray(param, dtype=np.float32)).reshape(-1)[:1], knotvector, degree)
