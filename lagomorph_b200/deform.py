"""Deformation-field operators: interp, its adjoint, and the compose family.

Mirrors lagomorph/deform.py of the reference (same names, argument meaning and
broadcasting); every call goes to liblagomorph_b200 through the C ABI.
"""
import numpy as np
import torch

from . import _lib as L


def identity(defshape, dtype=np.float32):
    """Identity deformation (numpy) for a shape in N C X Y (Z) order (reference: deform.py:10-21)."""
    dim = len(defshape) - 2
    ix = np.empty(defshape, dtype=dtype)
    for d in range(dim):
        ld = defshape[d + 2]
        shd = [1] * len(defshape)
        shd[d + 2] = ld
        ix[:, d, ...] = np.arange(ld, dtype=dtype).reshape(shd)
    return ix


def _check_interp_args(I, u):
    dev = L.require_cuda(I, u)
    d = L.spatial_dim(I)
    if u.dim() != I.dim() or u.shape[1] != d or tuple(u.shape[2:]) != tuple(I.shape[2:]):
        raise RuntimeError("interp: displacement must have shape (N, %d, *I.shape[2:])" % d)
    N, NI = u.shape[0], I.shape[0]
    if NI != N and NI != 1:
        # the reference reads u out of bounds here (cuda/interp.cu:90-92); refuse instead
        raise RuntimeError("interp: image batch must equal the displacement batch or be 1")
    return dev, d, N, NI


def interp_forward(I, u, dt=1.0):
    """out[n,c,x] = lerp_clamp(I[n or 0,c], x + dt*u[n,:,x]) (lagomorph_ext.interp_forward)."""
    dev, d, N, NI = _check_interp_args(I, u)
    I = I.contiguous()
    u = u.contiguous()
    C = I.shape[1]
    out = torch.empty((N, C) + tuple(I.shape[2:]), dtype=I.dtype, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_interp_fwd(L.dtype_code(I), L.ptr(out), L.ptr(I), L.ptr(u), N, NI, C, d,
                                     L.shape_arr(I.shape[2:]), float(dt), L.stream_ptr(dev)))
    return out


def interp_backward(gradout, I, u, dt, need_I=True, need_u=True):
    """(d_I, d_u) of interp; either is None when not needed (lagomorph_ext.interp_backward)."""
    dev, d, N, NI = _check_interp_args(I, u)
    L.require_cuda(gradout, I)
    gradout = gradout.contiguous()
    I = I.contiguous()
    u = u.contiguous()
    C = I.shape[1]
    d_I = torch.empty_like(I) if need_I else None  # zero-filled by the callee
    d_u = torch.empty_like(u) if need_u else None
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_interp_bwd(L.dtype_code(I), L.ptr(d_I), L.ptr(d_u), L.ptr(gradout), L.ptr(I),
                                     L.ptr(u), N, NI, C, d, L.shape_arr(I.shape[2:]), float(dt),
                                     L.stream_ptr(dev)))
    return d_I, d_u


class InterpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, u, dt):
        ctx.dt = dt
        ctx.save_for_backward(I, u)
        return interp_forward(I, u, dt)

    @staticmethod
    def backward(ctx, gradout):
        I, u = ctx.saved_tensors
        d_I, d_u = interp_backward(gradout, I, u, ctx.dt, *ctx.needs_input_grad[:2])
        return d_I, d_u, None


def interp(I, u, dt=1.0):
    return InterpFunction.apply(I, u, dt)


def interp_adjoint(g, u, dt=1.0, broadcast=False):
    """Adjoint (splat) of I -> interp(I, u, dt): returns d_I with d_I[n or 0] += w(x) g[n,:,x].

    This is the d_I output of the reference's interp_backward (cuda/interp.cu:220-223)
    exposed as an operator of its own. With broadcast=True all subjects splat into one image.
    """
    dev = L.require_cuda(g, u)
    d = L.spatial_dim(g)
    g = g.contiguous()
    u = u.contiguous()
    N, C = u.shape[0], g.shape[1]
    NI = 1 if broadcast else N
    d_I = torch.empty((NI, C) + tuple(g.shape[2:]), dtype=g.dtype, device=dev)
    with torch.cuda.device(dev):
        # I itself is only read for d_u, which is not requested: pass d_I's storage as a placeholder
        L.check(L.lib.lgm_interp_bwd(L.dtype_code(g), L.ptr(d_I), None, L.ptr(g), L.ptr(d_I), L.ptr(u),
                                     N, NI, C, d, L.shape_arr(g.shape[2:]), float(dt), L.stream_ptr(dev)))
    return d_I


def interp_hessian_diagonal_image(I, u, dt=1.0):
    raise NotImplementedError(
        "interp_hessian_diagonal_image is out of scope: 2-D only and aliased channels in the reference "
        "(cuda/interp.cu:342)")


class ComposeFunction(torch.autograd.Function):
    """ds*u(x) + dt*v(x + ds*u(x)) as one kernel; backward through the interp adjoint."""

    @staticmethod
    def forward(ctx, u, v, ds, dt):
        dev = L.require_cuda(u, v)
        d = L.spatial_dim(u)
        if u.shape != v.shape or u.shape[1] != d:
            raise RuntimeError("compose: u and v must both have shape (N, %d, ...)" % d)
        ctx.ds, ctx.dt = ds, dt
        ctx.save_for_backward(u, v)
        u = u.contiguous()
        v = v.contiguous()
        out = torch.empty_like(u)
        with torch.cuda.device(dev):
            L.check(L.lib.lgm_compose_fwd(L.dtype_code(u), L.ptr(out), L.ptr(u), L.ptr(v), u.shape[0], d,
                                          L.shape_arr(u.shape[2:]), float(ds), float(dt), L.stream_ptr(dev)))
        return out

    @staticmethod
    def backward(ctx, g):
        u, v = ctx.saved_tensors
        need_u, need_v = ctx.needs_input_grad[:2]
        d_v, d_u = interp_backward(g * ctx.dt, v, u, ctx.ds, need_v, need_u)
        if need_u:
            d_u = d_u + ctx.ds * g
        return d_u, d_v, None, None


def compose(u, v, ds=1.0, dt=1.0):
    """Return ds*u(x) + dt*v(x + ds*u(x)) (reference: deform.py:53-55)."""
    if u.shape == v.shape:
        return ComposeFunction.apply(u, v, ds, dt)
    return ds * u + dt * interp(v, u, dt=ds)


def compose_disp_vel(u, v, dt=1.0):
    """dt*v(x) + u(x + dt*v(x)) (reference: deform.py:58-62)."""
    return compose(v, u, ds=dt, dt=1.0)


def compose_vel_disp(v, u, dt=1.0):
    """u(x) + dt*v(x + u(x)) (reference: deform.py:65-70)."""
    return compose(u, v, ds=1.0, dt=dt)
