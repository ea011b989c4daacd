r"""Adjoint representation of Diff(R^d) (mirror of lagomorph/adjrep.py).

ad, ad_star and Ad_star run as ONE fused kernel each in the forward direction;
their backward passes reuse the component backward kernels, so gradients equal
those of the reference's unfused compositions.
"""
import torch

from . import _lib as L
from .deform import interp, interp_backward, interp_forward
from .diff import (jacobian_times_vectorfield, jtvf_adjoint_backward, jtvf_backward, jtvf_forward)


def _fused(fn, a, b):
    dev = L.require_cuda(a, b)
    d = L.spatial_dim(a)
    if a.shape != b.shape or a.shape[1] != d:
        raise RuntimeError("vector field is of wrong dimension")
    a = a.contiguous()
    b = b.contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(dev):
        L.check(fn(L.dtype_code(a), L.ptr(out), L.ptr(a), L.ptr(b), a.shape[0], d, L.shape_arr(a.shape[2:]),
                   L.stream_ptr(dev)))
    return out


class _AdFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, w):
        ctx.save_for_backward(v, w)
        return _fused(L.lib.lgm_ad_fwd, v, w)

    @staticmethod
    def backward(ctx, g):
        v, w = ctx.saved_tensors
        # ad(v,w) = J(v)w - J(w)v
        dv1, dw1 = jtvf_backward(g, v, w, False, False)
        dw2, dv2 = jtvf_backward(-g, w, v, False, False)
        return dv1 + dv2, dw1 + dw2


class _AdStarLittleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, m):
        ctx.save_for_backward(v, m)
        return _fused(L.lib.lgm_ad_star_fwd, v, m)

    @staticmethod
    def backward(ctx, g):
        v, m = ctx.saved_tensors
        # ad_star(v,m) = J(v)^T m - adj(m, v)
        dv1, dm1 = jtvf_backward(g, v, m, False, True)
        dm2, dv2 = jtvf_adjoint_backward(-g, m, v)
        return dv1 + dv2, dm1 + dm2


class _AdStarBigFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, phiinv, m):
        ctx.save_for_backward(phiinv, m)
        return _fused(L.lib.lgm_Ad_star_fwd, phiinv, m)

    @staticmethod
    def backward(ctx, g):
        phiinv, m = ctx.saved_tensors
        need_phi, need_m = ctx.needs_input_grad
        mphi = interp_forward(m, phiinv, 1.0)  # recomputed rather than saved
        d_phi1, d_mphi = jtvf_backward(g, phiinv, mphi, True, False, need_phi, True)
        d_m, d_phi2 = interp_backward(d_mphi, m, phiinv, 1.0, need_m, need_phi)
        d_phi = d_phi1 + d_phi2 if need_phi else None
        return d_phi, d_m


def ad(v, w):
    r"""ad(v,w) = -[v,w] = Dv w - Dw v (reference: adjrep.py:37-47)."""
    return _AdFunction.apply(v, w)


def Ad(phi, v):
    r"""Big adjoint action; not implemented in the reference either (adjrep.py:50-66)."""
    raise NotImplementedError


def ad_star(v, m):
    r"""ad^*(v,m) = (Dv)^T m + Dm v + m div v, the numerical adjoint of ad(v,.) (adjrep.py:69-83)."""
    return _AdStarLittleFunction.apply(v, m)


def Ad_star(phiinv, m):
    r"""Ad^*(phi,m)(x) = (D phi^{-1}(x) + I) m(x + phi^{-1}(x)) as the reference computes it:
    jacobian_times_vectorfield(phiinv, interp(m, phiinv), displacement=True) (adjrep.py:86-97).
    A momentum of batch 1 broadcasts against N deformations like the reference's interp does
    (cuda/interp.cu:90-92); that case runs the two-kernel composition, equal batches the fused kernel."""
    if m.dim() == phiinv.dim() and m.shape[0] != phiinv.shape[0] and m.shape[1:] == phiinv.shape[1:]:
        return jacobian_times_vectorfield(phiinv, interp(m, phiinv), displacement=True)
    return _AdStarBigFunction.apply(phiinv, m)


coad = Ad_star  # name used by the project brief for the big coadjoint action


def ad_dagger(x, y, metric):
    r"""ad^\dagger(x,y) = ad^*(x, y^\flat)^\sharp (adjrep.py:104-114)."""
    return metric.sharp(ad_star(x, metric.flat(y)))


def Ad_dagger(phi, y, metric):
    r"""Ad^\dagger(x,y) = Ad^*(x, y^\flat)^\sharp (adjrep.py:117-123)."""
    return metric.sharp(Ad_star(phi, metric.flat(y)))


def sym(x, y, metric):
    r"""sym(x,y) = -(ad^\dagger(x,y) + ad^\dagger(y,x)) (adjrep.py:126-136)."""
    return -(ad_dagger(x, y, metric) + ad_dagger(y, x, metric))


def sym_dagger(x, y, metric):
    r"""sym^\dagger(x,y) = ad^\dagger(y,x) - ad(x,y) (adjrep.py:139-145)."""
    return ad_dagger(y, x, metric) - ad(x, y)
