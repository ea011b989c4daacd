"""regrid and affine_interp (mirror of the operator part of lagomorph/affine.py)."""
import torch

from . import _lib as L


def affine_interp_forward(I, A, T):
    dev = L.require_cuda(I, A, T)
    d = L.spatial_dim(I)
    N = A.shape[0]
    if tuple(A.shape) != (N, d, d) or tuple(T.shape) != (N, d):
        raise RuntimeError("A must be (N,%d,%d) and T (N,%d)" % (d, d, d))
    NI = I.shape[0]
    if NI != N and NI != 1:
        raise RuntimeError("affine_interp: image batch must equal A's batch or be 1")
    I, A, T = I.contiguous(), A.contiguous(), T.contiguous()
    C = I.shape[1]
    out = torch.empty((N, C) + tuple(I.shape[2:]), dtype=I.dtype, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_affine_interp_fwd(L.dtype_code(I), L.ptr(out), L.ptr(I), L.ptr(A), L.ptr(T), N, NI, C, d,
                                            L.shape_arr(I.shape[2:]), L.stream_ptr(dev)))
    return out


def affine_interp_backward(grad_out, I, A, T, need_I=True, need_A=True, need_T=True):
    dev = L.require_cuda(grad_out, I, A, T)
    d = L.spatial_dim(I)
    if I.shape[1] != grad_out.shape[1]:
        raise RuntimeError("I and grad_out must have same number of channels")
    if A.shape[0] != T.shape[0]:
        raise RuntimeError("A and T must have same first dimension")
    grad_out, I, A, T = grad_out.contiguous(), I.contiguous(), A.contiguous(), T.contiguous()
    N, NI, C = A.shape[0], I.shape[0], I.shape[1]
    d_I = torch.empty_like(I) if need_I else None
    d_A = torch.empty_like(A) if need_A else None
    d_T = torch.empty_like(T) if need_T else None
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_affine_interp_bwd(L.dtype_code(I), L.ptr(d_I), L.ptr(d_A), L.ptr(d_T), L.ptr(grad_out),
                                            L.ptr(I), L.ptr(A), L.ptr(T), N, NI, C, d, L.shape_arr(I.shape[2:]),
                                            L.stream_ptr(dev)))
    return d_I, d_A, d_T


class AffineInterpFunction(torch.autograd.Function):
    """Interpolate an image under x -> A(x-o) + T + o, o the image centre (affine.py:11-36)."""

    @staticmethod
    def forward(ctx, I, A, T):
        ctx.save_for_backward(I, A, T)
        return affine_interp_forward(I, A, T)

    @staticmethod
    def backward(ctx, grad_out):
        I, A, T = ctx.saved_tensors
        return affine_interp_backward(grad_out, I, A, T, *ctx.needs_input_grad)


affine_interp = AffineInterpFunction.apply


class AffineInterp(torch.nn.Module):
    def forward(self, I, A, T):
        return AffineInterpFunction.apply(I, A, T)


def regrid_forward(I, outshape, origin, spacing):
    dev = L.require_cuda(I)
    d = L.spatial_dim(I)
    if len(outshape) != d:
        raise RuntimeError("Shape should be vector of size d (not 2+d)")
    if len(origin) != d:
        raise RuntimeError("Origin should be vector of size d (not 2+d)")
    if len(spacing) != d:
        raise RuntimeError("Spacing should be vector of size d (not 2+d)")
    I = I.contiguous()
    out = torch.empty(tuple(I.shape[:2]) + tuple(int(s) for s in outshape), dtype=I.dtype, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_regrid_fwd(L.dtype_code(I), L.ptr(out), L.ptr(I), I.shape[0], I.shape[1], d,
                                     L.shape_arr(I.shape[2:]), L.shape_arr(outshape), L.double_arr(origin),
                                     L.double_arr(spacing), L.stream_ptr(dev)))
    return out


def regrid_backward(grad_out, inshape, outshape, origin, spacing):
    dev = L.require_cuda(grad_out)
    d = L.spatial_dim(grad_out)
    grad_out = grad_out.contiguous()
    d_I = torch.empty(tuple(grad_out.shape[:2]) + tuple(int(s) for s in inshape), dtype=grad_out.dtype, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_regrid_bwd(L.dtype_code(grad_out), L.ptr(d_I), L.ptr(grad_out), grad_out.shape[0],
                                     grad_out.shape[1], d, L.shape_arr(inshape), L.shape_arr(outshape),
                                     L.double_arr(origin), L.double_arr(spacing), L.stream_ptr(dev)))
    return d_I


class RegridFunction(torch.autograd.Function):
    """Interpolate an image from one regular grid to another (affine.py:151-187)."""

    @staticmethod
    def forward(ctx, I, outshape, origin, spacing, displacement):
        outshape = [int(s) for s in outshape]
        origin = [float(o) for o in origin]
        spacing = [float(s) for s in spacing]
        ctx.inshape = I.shape[2:]
        ctx.outshape = outshape
        ctx.outorigin = origin
        ctx.outspacing = spacing
        ctx.displacement = displacement
        reg = regrid_forward(I, outshape, origin, spacing)
        if displacement:
            dim = len(I.shape) - 2
            if I.shape[1] != dim:
                raise ValueError("Incorrect num channels for regridding displacement")
            ctx.spacing_tensor = 1.0 / torch.tensor(spacing, dtype=reg.dtype, device=reg.device).view(
                1, dim, *[1] * dim)
            reg.mul_(ctx.spacing_tensor)
        return reg

    @staticmethod
    def backward(ctx, grad_out):
        d_I = regrid_backward(grad_out, ctx.inshape, ctx.outshape, ctx.outorigin, ctx.outspacing)
        if ctx.displacement:
            d_I.mul_(ctx.spacing_tensor)
        return d_I, None, None, None, None


def regrid_args(inshape, shape=None, origin=None, spacing=None):
    """Resolve (shape, origin, spacing) by the reference's rules (affine.py:190-272)."""
    if shape is None:
        if origin is None:
            if spacing is None:
                raise ValueError("At least one of shape, origin, or spacing required")
            raise NotImplementedError
        if spacing is None:
            raise NotImplementedError
        raise ValueError("Shape is required if specifying origin and spacing")
    d = len(inshape)
    if not isinstance(shape, (list, tuple, torch.Size)):
        shape = tuple([shape] * d)
    if origin is None:
        origin = tuple([(s - 1) * 0.5 for s in inshape])
        if spacing is None:
            spacing = tuple([(sI - 1) / (s - 1) for sI, s in zip(inshape, shape)])
    else:
        raise NotImplementedError
    if not isinstance(origin, (list, tuple)):
        origin = tuple([origin] * d)
    if not isinstance(spacing, (list, tuple)):
        spacing = tuple([spacing] * d)
    assert len(shape) == d
    assert len(origin) == d
    assert len(spacing) == d
    return tuple(shape), tuple(origin), tuple(spacing)


def regrid(I, shape=None, origin=None, spacing=None, displacement=False):
    """Interpolate from one regular grid to another; see the reference docstring
    (affine.py:190-243) for the argument rules, which are kept unchanged."""
    shape, origin, spacing = regrid_args(tuple(I.shape[2:]), shape, origin, spacing)
    return RegridFunction.apply(I, shape, origin, spacing, displacement)


class RegridModule(torch.nn.Module):
    def __init__(self, shape, origin, spacing):
        super(RegridModule, self).__init__()
        self.shape = shape
        self.origin = origin
        self.spacing = spacing

    def forward(self, I):
        return regrid(I, self.shape, self.origin, self.spacing)


# small batched algebra helpers of the reference's affine.py:49-148 (pure torch)
def det_2x2(A):
    return A[:, 0, 0] * A[:, 1, 1] - A[:, 0, 1] * A[:, 1, 0]


def affine_inverse(A, T):
    """Invert x -> A x + T: returns (A^-1, -A^-1 T) (reference: affine.py:101-111)."""
    Ainv = torch.linalg.inv(A)
    Tinv = -torch.matmul(Ainv, T.unsqueeze(2)).squeeze(2)
    return (Ainv, Tinv)
