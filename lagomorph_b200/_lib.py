"""ctypes binding of liblagomorph_b200.so (the C ABI in include/lagomorph_b200.h).

This is the only place the package touches native code. There is no CPU or
PyTorch fallback: if the shared library is missing the import raises, and every
operator raises for non-CUDA tensors.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGM_LIB_PATH") or os.path.join(_HERE, "liblagomorph_b200.so")  # override: kernel experiments only

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_double = ctypes.c_double
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/lagomorph_b200.h one to one
SIGNATURES = {
    "lgm_version": (c_int, []),
    "lgm_last_error": (ctypes.c_char_p, []),
    "lgm_set_debug_mode": (None, [c_int]),
    "lgm_get_debug_mode": (c_int, []),
    "lgm_launch_count": (c_i64, []),
    "lgm_profile_begin": (c_int, [c_void_p]),
    "lgm_profile_end": (c_int, [ctypes.c_char_p, c_i64]),
    "lgm_interp_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_i64_p, c_double, c_void_p]),
    "lgm_interp_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_i64_p, c_double, c_void_p]),
    "lgm_jtvf_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_int, c_int, c_void_p]),
    "lgm_jtvf_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_int, c_int, c_void_p]),
    "lgm_jtvf_adj_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_jtvf_adj_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_fluid_workspace_bytes": (c_i64, [c_int, c_i64, c_int, c_i64_p]),
    "lgm_fluid_apply": (c_int, [c_int, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_int, c_double, c_double, c_double, c_void_p, c_i64, c_void_p]),
    "lgm_fluid_operator": (c_int, [c_int, c_void_p, c_int, c_void_pp, c_void_pp, c_double, c_double, c_double, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_regrid_fwd": (c_int, [c_int, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_i64_p, c_double_p, c_double_p, c_void_p]),
    "lgm_regrid_bwd": (c_int, [c_int, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64_p, c_i64_p, c_double_p, c_double_p, c_void_p]),
    "lgm_affine_interp_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_affine_interp_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_Ad_star_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_ad_star_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_ad_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_void_p]),
    "lgm_compose_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_double, c_double, c_void_p]),
    "lgm_epdiff_scratch_bytes": (c_i64, [c_int, c_i64, c_int, c_i64_p]),
    "lgm_epdiff_step_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_double, c_double, c_double, c_double, c_void_p, c_i64, c_void_p]),
    "lgm_expmap_scratch_bytes": (c_i64, [c_int, c_i64, c_int, c_i64_p]),
    "lgm_expmap_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_double, c_int, c_double, c_double, c_double, c_void_p, c_i64, c_void_p]),
    "lgm_epdiff_bwd_scratch_bytes": (c_i64, [c_int, c_i64, c_int, c_i64_p]),
    "lgm_epdiff_step_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64_p, c_double, c_double, c_double, c_double, c_void_p, c_i64, c_int, c_int, c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "lagomorph_b200: native library %s is missing. Build it with "
            "`python -m lagomorph_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

F32, F64 = 0, 1


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError("lagomorph_b200 supports float32 and float64 tensors, got %s" % t.dtype)


def last_error():
    return lib.lgm_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise RuntimeError(last_error() or ("liblagomorph_b200 error %d" % rc))


def shape_arr(shape):
    return (ctypes.c_int64 * len(shape))(*[int(s) for s in shape])


def double_arr(vals):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


def require_cuda(*tensors):
    dev = None
    dt = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("lagomorph_b200 operators require CUDA tensors (no CPU fallback exists)")
        if dev is None:
            dev, dt = t.device, t.dtype
        elif t.device != dev:
            raise RuntimeError("all tensors must be on the same CUDA device")
        elif t.dtype != dt:
            raise RuntimeError("all tensors must have the same dtype")
    return dev


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def spatial_dim(t):
    d = t.dim() - 2
    if d not in (2, 3):
        raise RuntimeError("Only two- and three-dimensional fields are supported")
    return d


def aligned(t, n=16):
    """contiguous tensor whose data pointer is n-byte aligned (clones views that are not)"""
    t = t.contiguous()
    if t.data_ptr() % n:
        t = t.clone(memory_format=torch.contiguous_format)
    return t
