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
