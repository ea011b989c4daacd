"""oracle.py -- CPU ORACLE (test infrastructure only; never on the product path).

Python half of the oracle: a ctypes wrapper around oracle/lgm_oracle.cpp (the C++ restatement
of the reference's kernels) plus the reference's Python-level compositions restated on CPU
torch tensors:

  interp / compose*        lagomorph/deform.py:26-70
  jacobian_times_*         lagomorph/diff.py:7-61
  ad, ad_star, Ad_star ... lagomorph/adjrep.py:37-145
  FluidMetric              lagomorph/metric.py:9-97, with torch.rfft(x, d, normalized=True) ->
                           torch.fft.rfftn(x, dim=last d, norm="ortho") (same unitary one-sided DFT;
                           the removed API and cuFFT are the only third-party arithmetic on the path)
  EPDiff_step / expmap     lagomorph/lddmm.py:20-44, :73-91 (non-checkpointed branch)
  regrid                   lagomorph/affine.py:151-272

Pinned against: the reference's own per-point headers on the host (oracle/_ref/libref_points.so,
tests/test_oracle_pins.py), analytic known answers, the reference's property
tests, and the reference's own CUDA kernels run on the GPU box (oracle/_ref/libref_cuda.so ->
tests/golden/*.npz, tests/test_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblgm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "lgm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o",
                               _LIB_PATH, src])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _code(t):
    return {torch.float32: 0, torch.float64: 1}[t.dtype]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _sh(shape):
    return (ctypes.c_long * len(shape))(*[int(s) for s in shape])


def _c(t):
    assert not t.is_cuda, "the oracle runs on CPU tensors"
    return t.detach().contiguous()


L = ctypes.c_long
D = ctypes.c_double
I_ = ctypes.c_int


# ---- kernels ---------------------------------------------------------------------------
def interp_forward(I, u, dt=1.0):
    I, u = _c(I), _c(u)
    d = I.dim() - 2
    N = max(u.shape[0], I.shape[0])
    out = torch.empty((N, I.shape[1]) + tuple(I.shape[2:]), dtype=I.dtype)
    lib().orc_interp_fwd(I_(_code(I)), _p(out), _p(I), _p(u), L(N), L(I.shape[0]), L(I.shape[1]), I_(d),
                         _sh(I.shape[2:]), D(dt))
    return out


def interp_backward(go, I, u, dt=1.0, need_I=True, need_u=True):
    go, I, u = _c(go), _c(I), _c(u)
    d = I.dim() - 2
    N = max(u.shape[0], I.shape[0])
    d_I, d_u = torch.empty_like(I), torch.empty_like(u)
    lib().orc_interp_bwd(I_(_code(I)), _p(d_I), _p(d_u), _p(go), _p(I), _p(u), L(N), L(I.shape[0]),
                         L(I.shape[1]), I_(d), _sh(I.shape[2:]), D(dt), I_(int(need_I)), I_(int(need_u)))
    return d_I, d_u


def jtvf_forward(v, w, displacement=True, transpose=False):
    v, w = _c(v), _c(w)
    d = v.dim() - 2
    out = torch.empty_like(v)
    lib().orc_jtvf_fwd(I_(_code(v)), _p(out), _p(v), _p(w), L(v.shape[0]), L(v.shape[1]), I_(d),
                       _sh(v.shape[2:]), I_(int(displacement)), I_(int(transpose)))
    return out


def jtvf_backward(go, v, w, displacement=True, transpose=False):
    go, v, w = _c(go), _c(v), _c(w)
    d = v.dim() - 2
    d_v, d_w = torch.empty_like(v), torch.empty_like(w)
    lib().orc_jtvf_bwd(I_(_code(v)), _p(d_v), _p(d_w), _p(go), _p(v), _p(w), L(v.shape[0]), L(v.shape[1]),
                       I_(d), _sh(v.shape[2:]), I_(int(displacement)), I_(int(transpose)))
    return d_v, d_w


def jtvf_adjoint_forward(z, w):
    z, w = _c(z), _c(w)
    d = z.dim() - 2
    out = torch.empty_like(z)
    lib().orc_jtvf_adj_fwd(I_(_code(z)), _p(out), _p(z), _p(w), L(z.shape[0]), L(z.shape[1]), I_(d),
                           _sh(z.shape[2:]))
    return out


def jtvf_adjoint_backward(go, z, w):
    go, z, w = _c(go), _c(z), _c(w)
    d = z.dim() - 2
    d_z, d_w = torch.empty_like(z), torch.empty_like(w)
    lib().orc_jtvf_adj_bwd(I_(_code(z)), _p(d_z), _p(d_w), _p(go), _p(z), _p(w), L(z.shape[0]), L(z.shape[1]),
                           I_(d), _sh(z.shape[2:]))
    return d_z, d_w


def fluid_operator(Fmv, inverse, cosluts, sinluts, alpha, beta, gamma):
    """In place on the interleaved half spectrum (N,d,X,Y[,Zc],2); cuda/metric.cu:308-355."""
    assert Fmv.is_contiguous()
    d = Fmv.dim() - 3
    cl = [_c(c) for c in cosluts] + [None] * (3 - d)
    sl = [_c(s) for s in sinluts] + [None] * (3 - d)
    pp = lambda t: _p(t) if t is not None else None
    lib().orc_fluid_operator(I_(_code(Fmv)), _p(Fmv), I_(int(inverse)), pp(cl[0]), pp(sl[0]), pp(cl[1]),
                             pp(sl[1]), pp(cl[2]), pp(sl[2]), D(alpha), D(beta), D(gamma), L(Fmv.shape[0]),
                             I_(d), _sh(Fmv.shape[2:2 + d]))


def regrid_forward(I, outshape, origin, spacing):
    I = _c(I)
    d = I.dim() - 2
    out = torch.empty(tuple(I.shape[:2]) + tuple(outshape), dtype=I.dtype)
    lib().orc_regrid_fwd(I_(_code(I)), _p(out), _p(I), L(I.shape[0]), L(I.shape[1]), I_(d), _sh(I.shape[2:]),
                         _sh(outshape), (D * d)(*origin), (D * d)(*spacing))
    return out


def regrid_backward(go, inshape, outshape, origin, spacing):
    go = _c(go)
    d = go.dim() - 2
    d_I = torch.empty(tuple(go.shape[:2]) + tuple(inshape), dtype=go.dtype)
    lib().orc_regrid_bwd(I_(_code(go)), _p(d_I), _p(go), L(go.shape[0]), L(go.shape[1]), I_(d), _sh(inshape),
                         _sh(outshape), (D * d)(*origin), (D * d)(*spacing))
    return d_I


def affine_interp_forward(I, A, T):
    I, A, T = _c(I), _c(A), _c(T)
    d = I.dim() - 2
    N = A.shape[0]
    out = torch.empty((N, I.shape[1]) + tuple(I.shape[2:]), dtype=I.dtype)
    lib().orc_affine_interp_fwd(I_(_code(I)), _p(out), _p(I), _p(A), _p(T), L(N), L(I.shape[0]), L(I.shape[1]),
                                I_(d), _sh(I.shape[2:]))
    return out


# ---- compositions (reference Python layer) ------------------------------------------------
def interp(I, u, dt=1.0):
    return interp_forward(I, u, dt)


def compose(u, v, ds=1.0, dt=1.0):  # deform.py:53-55
    return ds * u + dt * interp(v, u, dt=ds)


def compose_disp_vel(u, v, dt=1.0):  # deform.py:58-62
    return compose(v, u, ds=dt, dt=1.0)


def compose_vel_disp(v, u, dt=1.0):  # deform.py:65-70
    return compose(u, v, ds=1.0, dt=dt)


def jacobian_times_vectorfield(v, w, displacement=True, transpose=False):
    return jtvf_forward(v, w, displacement, transpose)


def jacobian_times_vectorfield_adjoint(z, w):
    return jtvf_adjoint_forward(z, w)


def ad(v, w):  # adjrep.py:37-47
    return jtvf_forward(v, w, False, False) - jtvf_forward(w, v, False, False)


def ad_star(v, m):  # adjrep.py:69-83
    return jtvf_forward(v, m, False, True) - jtvf_adjoint_forward(m, v)


def Ad_star(phiinv, m):  # adjrep.py:86-97
    return jtvf_forward(phiinv, interp(m, phiinv), True, False)


class FluidMetric:
    """metric.py:37-97 on CPU: rfftn(ortho) -> restated fluid kernel -> irfftn(ortho)."""

    def __init__(self, params=[0.1, 0.0, 0.001]):
        assert len(params) == 3
        self.params = params

    @staticmethod
    def luts(shape, dtype):
        # metric.py:53-75: float64 numpy -> torch.Tensor (float32!) -> .type(dtype)
        cshape = list(shape)
        cshape[-1] = cshape[-1] // 2 + 1
        cos, sin = [], []
        for (Nf, N) in zip(cshape[2:], shape[2:]):
            cos.append(torch.Tensor(2.0 * (1.0 - np.cos(2 * np.pi * np.arange(Nf) / N))).type(dtype))
            sin.append(torch.Tensor(np.sin(2.0 * np.pi * np.arange(Nf) / N)).type(dtype))
        return cos, sin

    def operator(self, mv, inverse):
        mv = _c(mv)
        sh = mv.shape
        d = len(sh) - 2
        dims = tuple(range(2, 2 + d))
        Fmv = torch.view_as_real(torch.fft.rfftn(mv, dim=dims, norm="ortho")).contiguous()
        cos, sin = self.luts(sh, mv.dtype)
        fluid_operator(Fmv, inverse, cos, sin, *self.params)
        return torch.fft.irfftn(torch.view_as_complex(Fmv), s=sh[2:], dim=dims, norm="ortho")

    def sharp(self, m):
        return self.operator(m, True)

    def flat(self, m):
        return self.operator(m, False)


def EPDiff_step(metric, m0, dt, phiinv, mommask=None):  # lddmm.py:39-44
    m = Ad_star(phiinv, m0)
    if mommask is not None:
        m = m * mommask
    v = metric.sharp(m)
    return compose_disp_vel(phiinv, v, dt=-dt)


def expmap(metric, m0, T=1.0, num_steps=10, phiinv=None, mommask=None):  # lddmm.py:73-91
    if phiinv is None:
        phiinv = torch.zeros_like(m0)
    dt = T / num_steps
    for i in range(num_steps):
        phiinv = EPDiff_step(metric, m0, dt, phiinv, mommask=mommask)
    return phiinv


def expmap_advect(metric, m, T=1.0, num_steps=10, phiinv=None):  # lddmm.py:20-36
    if phiinv is None:
        phiinv = torch.zeros_like(m)
    dt = T / num_steps
    v = metric.sharp(m)
    phiinv = compose_disp_vel(phiinv, v, dt=-dt)
    for i in range(num_steps - 1):
        m = m - dt * ad_star(v, m)
        v = metric.sharp(m)
        phiinv = compose_disp_vel(phiinv, v, dt=-dt)
    return phiinv


def regrid(I, shape, origin=None, spacing=None, displacement=False):  # affine.py:151-272
    d = I.dim() - 2
    if not isinstance(shape, (list, tuple)):
        shape = tuple([shape] * d)
    if origin is None:
        origin = tuple([(s - 1) * 0.5 for s in I.shape[2:]])
        if spacing is None:
            spacing = tuple([(sI - 1) / (s - 1) for sI, s in zip(I.shape[2:], shape)])
    reg = regrid_forward(I, shape, origin, spacing)
    if displacement:
        reg = reg * (1.0 / torch.tensor(spacing, dtype=reg.dtype).view(1, d, *[1] * d))
    return reg


# dagger / sym compositions, adjrep.py:104-145
def ad_dagger(x, y, metric):
    return metric.sharp(ad_star(x, metric.flat(y)))


def Ad_dagger(phi, y, metric):
    return metric.sharp(Ad_star(phi, metric.flat(y)))


def sym(x, y, metric):
    return -(ad_dagger(x, y, metric) + ad_dagger(y, x, metric))


def sym_dagger(x, y, metric):
    return ad_dagger(y, x, metric) - ad(x, y)


# ---- autograd layer: the reference's torch.autograd.Functions over the oracle kernels -------------
class InterpFn(torch.autograd.Function):  # deform.py:26-41
    @staticmethod
    def forward(ctx, I, u, dt):
        ctx.dt = dt
        ctx.save_for_backward(I, u)
        return interp_forward(I, u, dt)

    @staticmethod
    def backward(ctx, go):
        I, u = ctx.saved_tensors
        d_I, d_u = interp_backward(go, I, u, ctx.dt, *ctx.needs_input_grad[:2])
        return d_I, d_u, None


class JtvfFn(torch.autograd.Function):  # diff.py:7-35
    @staticmethod
    def forward(ctx, v, w, displacement, transpose):
        ctx.displacement, ctx.transpose = displacement, transpose
        ctx.save_for_backward(v, w)
        return jtvf_forward(v, w, displacement, transpose)

    @staticmethod
    def backward(ctx, go):
        v, w = ctx.saved_tensors
        d_v, d_w = jtvf_backward(go, v, w, ctx.displacement, ctx.transpose)
        return d_v, d_w, None, None


class FluidFn(torch.autograd.Function):  # metric.py:9-34: the backward is the same operator on gout
    @staticmethod
    def forward(ctx, metric, inverse, mv):
        ctx.metric, ctx.inverse = metric, inverse
        return metric.operator(mv, inverse)

    @staticmethod
    def backward(ctx, go):
        return None, None, ctx.metric.operator(go, ctx.inverse)


def ag_expmap(metric, m0, T=1.0, num_steps=10):
    """differentiable expmap: lddmm.py:73-91 with Ad_star = adjrep.py:96-97, compose = deform.py:53-62"""
    phiinv = torch.zeros_like(m0)
    dt = T / num_steps
    for i in range(num_steps):
        m = JtvfFn.apply(phiinv, InterpFn.apply(m0, phiinv, 1.0), True, False)
        v = FluidFn.apply(metric, True, m)
        phiinv = (-dt) * v + 1.0 * InterpFn.apply(phiinv, v, -dt)
    return phiinv


def lddmm_step(metric, I, m, img, num_subjects, integration_steps=5, reg_weight=1e2, learning_rate_pose=2e2,
               momentum_preconditioning=False, need_image_grad=True):
    """One batch of LDDMMAtlasBuilder (lddmm.py:300-325, same-grid momenta). Returns
    (new m, loss * norm_factor, reg_term * norm_factor, dL/dI or None)."""
    m = m.detach().clone().requires_grad_(True)
    I = I.detach().clone().requires_grad_(need_image_grad)
    h = ag_expmap(metric, m, num_steps=integration_steps)
    Idef = InterpFn.apply(I, h, 1.0)
    v = FluidFn.apply(metric, True, m)
    reg_term = reg_weight * (v * m).sum() / img.numel()
    loss = ((Idef - img) ** 2).sum() / img.numel() + reg_term  # mse_loss(reduction="sum") / numel
    loss.backward()
    with torch.no_grad():
        norm_factor = img.shape[0] / num_subjects
        p = m.grad
        if momentum_preconditioning:
            p = metric.flat(p)
        m_new = m.detach() - learning_rate_pose * p  # m.add_(-lr, p)
    return m_new, (loss * norm_factor).item(), (reg_term * norm_factor).item(), (I.grad if need_image_grad else None)


def lddmm_epoch(metric, I, ms, batches, num_subjects, learning_rate_image=1e4, image_update_freq=0, world_size=1,
                **step_kwargs):
    """One epoch of one rank (lddmm.py:327-358) over in-memory batches, image update as
    lddmm.py:287-298 (gradient averaged over image_iters * world_size, plain SGD). Returns
    (new I, new ms, epoch loss, epoch reg term)."""
    I = I.detach().clone()
    acc = torch.zeros_like(I)
    iters = 0
    eloss = ereg = 0.0
    new_ms = []

    def update(force=False):
        nonlocal I, acc, iters
        if (iters < image_update_freq and not force) or iters == 0:
            return
        I = I - learning_rate_image * (acc / (iters * world_size))
        acc = torch.zeros_like(I)
        iters = 0

    for m, img in zip(ms, batches):
        m, l, r, gI = lddmm_step(metric, I, m, img, num_subjects, **step_kwargs)
        acc += gI
        new_ms.append(m)
        eloss += l
        ereg += r
        iters += 1
        update()
    update(force=True)
    return I, new_ms, eloss, ereg


# ---- affine atlas (config 4's caller) ---------------------------------------------------------
def affine_interp_backward(go, I, A, T, need_I=True, need_A=True, need_T=True):
    """cuda/affine.cu:538-610 -> kernels :171-536"""
    go, I, A, T = _c(go), _c(I), _c(A), _c(T)
    d = I.dim() - 2
    N = A.shape[0]
    d_I, d_A, d_T = torch.zeros_like(I), torch.zeros_like(A), torch.zeros_like(T)
    lib().orc_affine_interp_bwd(I_(_code(I)), _p(d_I), _p(d_A), _p(d_T), _p(go), _p(I), _p(A), _p(T), L(N),
                                L(I.shape[0]), L(I.shape[1]), I_(d), _sh(I.shape[2:]), I_(int(need_I)),
                                I_(int(need_A)), I_(int(need_T)))
    return d_I, d_A, d_T


class AffineFn(torch.autograd.Function):  # affine.py:11-36
    @staticmethod
    def forward(ctx, I, A, T):
        ctx.save_for_backward(I, A, T)
        return affine_interp_forward(I, A, T)

    @staticmethod
    def backward(ctx, go):
        I, A, T = ctx.saved_tensors
        return affine_interp_backward(go, I, A, T, *ctx.needs_input_grad)


def affine_atlas_epoch(I, As, Ts, imgs, batch_size, num_subjects, image_update_freq=0, affine_steps=1,
                       reg_weightA=0.0, reg_weightT=0.0, learning_rate_A=1e-3, learning_rate_T=1e-2,
                       learning_rate_I=1e5, world_size=1):
    """One epoch of one rank of affine_atlas (affine.py:345-405) over in-memory images. As holds
    A minus the identity, like the reference. Returns (I, As, Ts, epoch loss, iteration losses)."""
    I = I.detach().clone()
    As, Ts = As.detach().clone(), Ts.detach().clone()
    dim = Ts.shape[1]
    eye = torch.eye(dim, dtype=I.dtype).view(1, dim, dim)
    nvox = float(np.prod(I.shape[2:]))
    acc = torch.zeros_like(I)
    image_iters = 0
    epoch_loss = 0.0
    iter_losses = []

    def image_step():
        nonlocal I, acc, image_iters
        I = I - learning_rate_I * (acc / (image_iters * world_size))  # SGD step, affine.py:385-392
        acc = torch.zeros_like(I)
        image_iters = 0

    for b0 in range(0, imgs.shape[0], batch_size):
        sl = slice(b0, min(b0 + batch_size, imgs.shape[0]))
        img = imgs[sl]
        A, T = As[sl].clone(), Ts[sl].clone()
        for affit in range(affine_steps):
            A = A.detach().requires_grad_(True)
            T = T.detach().requires_grad_(True)
            last = affit == affine_steps - 1
            Iin = I.detach().clone().requires_grad_(last)
            Idef = AffineFn.apply(Iin, A + eye, T)
            regloss = 0.0
            if reg_weightA > 0:
                regloss = regloss + 0.5 * reg_weightA * torch.dot(A.reshape(-1), A.reshape(-1))
            if reg_weightT > 0:
                regloss = regloss + 0.5 * reg_weightT * torch.dot(T.reshape(-1), T.reshape(-1))
            loss = (((Idef - img) ** 2).sum() * (1.0 / nvox) + regloss) / img.shape[0]
            loss.backward()
            with torch.no_grad():
                li = loss.item() * (img.shape[0] / num_subjects)
                iter_losses.append(li)
                A = A.detach() - learning_rate_A * A.grad
                T = T.detach() - learning_rate_T * T.grad
                if last:
                    acc += Iin.grad
        image_iters += 1
        if image_iters == image_update_freq:
            image_step()
        epoch_loss += li
        As[sl], Ts[sl] = A, T
    if image_iters > 0:
        image_step()
    return I, As, Ts, epoch_loss, iter_losses
