"""Finite-difference Jacobian operators (mirror of lagomorph/diff.py)."""
import torch

from . import _lib as L


def _args(v, w):
    dev = L.require_cuda(v, w)
    d = L.spatial_dim(v)
    if w.dim() != v.dim() or w.shape[1] != d:
        raise RuntimeError("vector field is of wrong dimension")
    if v.shape[0] != w.shape[0]:
        raise RuntimeError("arguments must have same batch size dimension")
    if tuple(v.shape[2:]) != tuple(w.shape[2:]):
        raise RuntimeError("arguments must have the same spatial shape")
    return dev, d


def jtvf_forward(v, w, displacement, transpose):
    dev, d = _args(v, w)
    v = v.contiguous()
    w = w.contiguous()
    out = torch.empty_like(v)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_jtvf_fwd(L.dtype_code(v), L.ptr(out), L.ptr(v), L.ptr(w), v.shape[0], v.shape[1], d,
                                   L.shape_arr(v.shape[2:]), int(bool(displacement)), int(bool(transpose)),
                                   L.stream_ptr(dev)))
    return out


def jtvf_backward(gradout, v, w, displacement, transpose, need_v=True, need_w=True):
    dev, d = _args(v, w)
    L.require_cuda(gradout, v)
    gradout = gradout.contiguous()
    v = v.contiguous()
    w = w.contiguous()
    d_v = torch.empty_like(v) if need_v else None
    d_w = torch.empty_like(w) if need_w else None
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_jtvf_bwd(L.dtype_code(v), L.ptr(d_v), L.ptr(d_w), L.ptr(gradout), L.ptr(v), L.ptr(w),
                                   v.shape[0], v.shape[1], d, L.shape_arr(v.shape[2:]),
                                   int(bool(displacement)), int(bool(transpose)), L.stream_ptr(dev)))
    return d_v, d_w


def jtvf_adjoint_forward(z, w):
    dev, d = _args(z, w)
    z = z.contiguous()
    w = w.contiguous()
    out = torch.empty_like(z)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_jtvf_adj_fwd(L.dtype_code(z), L.ptr(out), L.ptr(z), L.ptr(w), z.shape[0], z.shape[1], d,
                                       L.shape_arr(z.shape[2:]), L.stream_ptr(dev)))
    return out


def jtvf_adjoint_backward(gradout, z, w, need_z=True, need_w=True):
    dev, d = _args(z, w)
    L.require_cuda(gradout, z)
    gradout = gradout.contiguous()
    z = z.contiguous()
    w = w.contiguous()
    d_z = torch.empty_like(z) if need_z else None
    d_w = torch.empty_like(w) if need_w else None
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_jtvf_adj_bwd(L.dtype_code(z), L.ptr(d_z), L.ptr(d_w), L.ptr(gradout), L.ptr(z), L.ptr(w),
                                       z.shape[0], z.shape[1], d, L.shape_arr(z.shape[2:]), L.stream_ptr(dev)))
    return d_z, d_w


class JacobianTimesVectorFieldFunction(torch.autograd.Function):
    """(Dv)w with the clamped central-difference Jacobian of v; `displacement` adds the
    identity to the Jacobian, `transpose` contracts with its transpose (reference: diff.py:7-35)."""

    @staticmethod
    def forward(ctx, v, w, displacement, transpose):
        ctx.displacement = displacement
        ctx.transpose = transpose
        ctx.save_for_backward(v, w)
        return jtvf_forward(v, w, displacement, transpose)

    @staticmethod
    def backward(ctx, gradout):
        v, w = ctx.saved_tensors
        d_v, d_w = jtvf_backward(gradout, v, w, ctx.displacement, ctx.transpose, *ctx.needs_input_grad[:2])
        return d_v, d_w, None, None


def jacobian_times_vectorfield(v, w, displacement=True, transpose=False):
    return JacobianTimesVectorFieldFunction.apply(v, w, displacement, transpose)


class JacobianTimesVectorFieldAdjointFunction(torch.autograd.Function):
    r"""The adjoint T(w)^\dagger z of the linear map T(w)v = (Dv)w (reference: diff.py:42-58)."""

    @staticmethod
    def forward(ctx, v, w):
        ctx.save_for_backward(v, w)
        return jtvf_adjoint_forward(v, w)

    @staticmethod
    def backward(ctx, gradout):
        v, w = ctx.saved_tensors
        d_v, d_w = jtvf_adjoint_backward(gradout, v, w, *ctx.needs_input_grad[:2])
        return d_v, d_w


jacobian_times_vectorfield_adjoint = JacobianTimesVectorFieldAdjointFunction.apply
