"""`lagomorph_ext` over liblagomorph_b200: the reference's pybind11 module, function for function.

The reference's Python layer (lagomorph/deform.py, diff.py, metric.py, affine.py) imports one
native module, `lagomorph_ext` (lagomorph/extension/extension.cpp:175-189, 13 functions). This
module has the same names, positional signatures, return conventions (a Tensor, or a list of
Tensors for the backward functions, zero-filled where the reference zero-fills) and argument
checks, and forwards every call to the C ABI of include/lagomorph_b200.h. Installing it as

    import sys, lagomorph_b200.lagomorph_ext as ext
    sys.modules["lagomorph_ext"] = ext

lets the reference's own `deform.py` / `diff.py` / `adjrep.py` / `affine.py` run unmodified on the
sm_100a kernels (tests/test_ext_shim.py does exactly that with the files under /root/reference).

`interp_hessian_diagonal_image` (extension.cpp:183) is out of scope and raises.
"""
import torch

from . import _lib as L
from . import affine as _affine
from . import deform as _deform
from . import diff as _diff
from . import metric as _metric


def _check_input(x, name):
    # CHECK_INPUT (extension.cpp:8-10)
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not x.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _zeros_if_none(t, like):
    return torch.zeros_like(like) if t is None else t


def set_debug_mode(mode):
    """extension.cpp:105-107"""
    L.lib.lgm_set_debug_mode(1 if mode else 0)


def interp_forward(Iv, u, dt=1.0):
    """extension.cpp:135-143 -> cuda/interp.cu:80-130"""
    _check_input(Iv, "Iv")
    _check_input(u, "u")
    return _deform.interp_forward(Iv, u, dt)


def interp_backward(grad_out, I, u, dt, need_I, need_u):
    """extension.cpp:145-156 -> cuda/interp.cu:246-313; returns [d_I, d_u], both always allocated and
    zero where not needed (cuda/interp.cu:262-264)"""
    _check_input(grad_out, "grad_out")
    _check_input(I, "I")
    _check_input(u, "u")
    d_I, d_u = _deform.interp_backward(grad_out, I, u, dt, bool(need_I), bool(need_u))
    return [_zeros_if_none(d_I, I), _zeros_if_none(d_u, u)]


def interp_hessian_diagonal_image(Iv, u, dt=1.0):
    return _deform.interp_hessian_diagonal_image(Iv, u, dt)


def jacobian_times_vectorfield_forward(g, v, displacement, transpose):
    """cuda/diff.cu:129-185 (first argument is the differentiated field)"""
    return _diff.jtvf_forward(g, v, displacement, transpose)


def jacobian_times_vectorfield_backward(grad_out, v, w, displacement, transpose, need_v, need_w):
    """cuda/diff.cu:475-540; the reference computes both gradients whatever need_* says (:483-484)"""
    d_v, d_w = _diff.jtvf_backward(grad_out, v, w, displacement, transpose, True, True)
    return [d_v, d_w]


def jacobian_times_vectorfield_adjoint_forward(g, v):
    """cuda/diff.cu:634-672"""
    return _diff.jtvf_adjoint_forward(g, v)


def jacobian_times_vectorfield_adjoint_backward(grad_out, v, w, need_v, need_w):
    """cuda/diff.cu:783-835; both gradients always (:789-790)"""
    d_z, d_w = _diff.jtvf_adjoint_backward(grad_out, v, w, True, True)
    return [d_z, d_w]


def fluid_operator(Fmv, inverse, cosluts, sinluts, alpha, beta, gamma):
    """extension.cpp:158-173 -> cuda/metric.cu:308-355; in place on the interleaved half spectrum"""
    _check_input(Fmv, "Fmv")
    dim = Fmv.dim() - 3
    if len(cosluts) != dim:
        raise RuntimeError("Must provide same number cosine LUTs (%d) as spatial dimension '%d'" % (len(cosluts), dim))
    if len(sinluts) != dim:
        raise RuntimeError("Must provide same number sine LUTs (%d) as spatial dimension '%d'" % (len(sinluts), dim))
    _metric.fluid_operator(Fmv, inverse, list(cosluts), list(sinluts), alpha, beta, gamma)


def regrid_forward(I, shape, origin, spacing):
    """cuda/affine.cu:683-734"""
    _check_input(I, "I")
    return _affine.regrid_forward(I, [int(s) for s in shape], [float(o) for o in origin], [float(s) for s in spacing])


def regrid_backward(grad_out, inshape, shape, origin, spacing):
    """cuda/affine.cu:802-855"""
    _check_input(grad_out, "grad_out")
    return _affine.regrid_backward(grad_out, [int(s) for s in inshape], [int(s) for s in shape],
                                   [float(o) for o in origin], [float(s) for s in spacing])


def affine_interp_forward(I, A, T):
    """extension.cpp:109-118 -> cuda/affine.cu:114-169. The reference also has a CPU branch here
    (cpu/affine.cpp); this library has no CPU path and raises for host tensors."""
    return _affine.affine_interp_forward(I, A, T)


def affine_interp_backward(grad_out, I, A, T, need_I, need_A, need_T):
    """extension.cpp:120-133 -> cuda/affine.cu:538-610; returns [d_I, d_A, d_T], zero where not needed"""
    _check_input(grad_out, "grad_out")
    _check_input(I, "I")
    _check_input(A, "A")
    _check_input(T, "T")
    d_I, d_A, d_T = _affine.affine_interp_backward(grad_out, I, A, T, bool(need_I), bool(need_A), bool(need_T))
    return [_zeros_if_none(d_I, I), _zeros_if_none(d_A, A), _zeros_if_none(d_T, T)]


__all__ = [
    "set_debug_mode", "affine_interp_forward", "affine_interp_backward", "regrid_forward", "regrid_backward",
    "fluid_operator", "interp_forward", "interp_backward", "interp_hessian_diagonal_image",
    "jacobian_times_vectorfield_forward", "jacobian_times_vectorfield_backward",
    "jacobian_times_vectorfield_adjoint_forward", "jacobian_times_vectorfield_adjoint_backward",
]
