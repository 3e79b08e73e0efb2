"""ctypes binding of libthb_b200.so (declared in include/thb_b200.h).

There is NO CPU fallback: importing this module fails loudly when the shared library has not been built
(python -m thb_diff_b200.build), and every call raises RuntimeError when the library reports an error --
the reference raises RuntimeError (c10::Error) on bad inputs as well.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("THB_LIB", os.path.join(_HERE, "libthb_b200.so"))  # THB_LIB: tuning builds only
MAX_LEVELS = 8
INFO_W = 12


class SpaceDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nlevels", C.c_int32), ("degree", C.c_int32 * 3), ("tile_stride", C.c_int32),
        ("nslots", C.c_int32), ("max_tiled", C.c_int32),
        ("knot_off", (C.c_int32 * 3) * MAX_LEVELS), ("knot_len", (C.c_int32 * 3) * MAX_LEVELS),
        ("ncells", (C.c_int32 * 3) * MAX_LEVELS), ("nfns", (C.c_int32 * 3) * MAX_LEVELS),
        ("fn_off", C.c_int64 * (MAX_LEVELS + 1)), ("cell_slot_off", C.c_int64 * MAX_LEVELS),
        ("knots", C.c_void_p), ("cell_slot", C.c_void_p), ("cellinfo", C.c_void_p), ("dmask", C.c_void_p),
        ("supp_jm", C.c_void_p), ("tiles", C.c_void_p),
        ("sep_vec", C.c_void_p), ("sep_jm", C.c_void_p), ("full_tiles", C.c_void_p), ("full_jm", C.c_void_p),
        ("span_lut", C.c_void_p), ("lut_off", C.c_int32 * 3), ("lut_n", C.c_int32 * 3), ("lut_scale", C.c_float * 3), ("lut_exact", C.c_int32),
        ("finest_slot", C.c_void_p), ("tiled_ref", C.c_void_p),
        ("poly", C.c_void_p), ("poly_off", (C.c_int32 * 3) * MAX_LEVELS),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("npoints", C.c_int64), ("points_per_item", C.c_int32), ("max_items", C.c_int32),
        ("slot", C.c_void_p), ("scratch", C.c_void_p), ("sorted_params", C.c_void_p), ("cell_count", C.c_void_p), ("cell_start", C.c_void_p),
        ("cell_cursor", C.c_void_p), ("items", C.c_void_p), ("counters", C.c_void_p), ("rows_in_plan_order", C.c_int32), ("light", C.c_int32),
        ("params", C.c_void_p),
    ]


class OrderedDesc(C.Structure):
    _fields_ = [
        ("item_ws", C.c_void_p), ("item0", C.c_void_p), ("contrib", C.c_void_p), ("tr_ptr", C.c_void_p), ("tr_idx", C.c_void_p),
        ("active_ids", C.c_void_p), ("nsupp", C.c_int64), ("nrows", C.c_int32), ("pad_", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -m thb_diff_b200.build). "
            "thb_diff_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    sig = {
        "thb_version": (i32, []),
        "thb_last_error": (C.c_char_p, []),
        "thb_launch_count": (i64, []),
        "thb_eval_forward": (i32, [vp, vp, i32, vp, vp, i32, vp, i64, i32, vp]),
        "thb_eval_backward": (i32, [vp, vp, i32, vp, vp, i32, vp, i64, i64, i32, vp]),
        "thb_find_spans": (i32, [vp, i64, i64, vp, i32, i32, vp, vp]),
        "thb_basis_1d": (i32, [vp, i64, i64, vp, i32, i32, vp, vp, vp]),
        "thb_active_span": (i32, [C.POINTER(SpaceDesc), vp, i64, vp, vp, vp]),
        "thb_expand_supports": (i32, [C.POINTER(SpaceDesc), vp, vp, i64, vp, i32, vp]),
        "thb_plan_build": (i32, [C.POINTER(SpaceDesc), vp, C.POINTER(PlanDesc), vp]),
        "thb_plan_build_phase": (i32, [C.POINTER(SpaceDesc), vp, C.POINTER(PlanDesc), C.c_int, vp]),
        "thb_basis_tiled": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, vp, C.POINTER(C.c_int32), i32, C.POINTER(vp), vp]),
        "thb_basis_dense": (i32, [vp, i64, i32, C.POINTER(C.c_int32), vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), vp, vp, i32, vp, i32, vp, vp]),
        "thb_eval_fused": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, vp, i32, C.POINTER(C.c_int32), i32, C.POINTER(vp), i32, vp, vp]),
        "thb_net_workspace_floats": (i64, [C.POINTER(SpaceDesc), i32]),
        "thb_cell_nets": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, i32, i32, vp, vp]),
        "thb_eval_nets": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, i32, i32, C.POINTER(C.c_int32), i32, C.POINTER(vp), vp]),
        "thb_eval_fused_backward": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, vp, i32, vp, i64, vp, vp]),
        "thb_eval_fused_backward_rows": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, vp, i32, vp, vp, i64, vp, vp]),
        "thb_eval_fused_backward_ordered": (i32, [C.POINTER(SpaceDesc), C.POINTER(PlanDesc), vp, vp, i32, C.POINTER(OrderedDesc), vp, i64, vp, vp]),
        "thb_mse_workspace_bytes": (i64, []),
        "thb_mse_grad": (i32, [vp, vp, i64, C.c_float, vp, vp, vp, vp]),
        "thb_adam_rows": (i32, [vp, vp, vp, vp, vp, i64, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32, vp, vp]),
        "thb_adam_rows_peer": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, i64, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp]),
        "thb_plan_records": (i32, [C.POINTER(SpaceDesc), vp, C.POINTER(PlanDesc), vp]),
        "thb_eval_grid": (i32, [C.POINTER(SpaceDesc), vp, vp, vp, i32, vp, vp, vp, C.POINTER(C.c_int64), i32, vp, vp, vp]),
        "thb_tensor_product": (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp]),
        "thb_fma_peak": (i32, [i32, i32, vp, C.POINTER(C.c_double), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sig.keys())


lib, EXPORTS = _load()


def check(rc, who=""):
    if rc != 0:
        msg = lib.thb_last_error().decode(errors="replace")
        raise RuntimeError(f"libthb_b200 {who} failed ({rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(lib.thb_launch_count())


def require_cuda(t, name, dtype=None):
    import torch

    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (thb_diff_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t
