"""Shared helpers for the test-suite (seeded inputs, tolerances, reference-CUDA binding)."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
REF_POINTS = os.path.join(ROOT, "oracle", "_ref", "libref_points.so")

# parity tolerances: fp32 CUDA vs fp32 oracle on identical inputs
TOL32 = 1e-5   # interp / jacobian / metric, relative to max|reference|
TOL64 = 1e-12
TOL_SPLAT32 = 1e-4  # atomics reorder the sum


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def randn(shape, dtype=torch.float32, seed=1, scale=1.0):
    return (torch.randn(shape, generator=gen(seed), dtype=torch.float64) * scale).to(dtype)


def smooth_field(shape, dtype=torch.float32, seed=1, amp=3.0, sigma=2.0):
    """Gaussian-smoothed white noise scaled to max |.| = amp voxels."""
    x = torch.randn(shape, generator=gen(seed), dtype=torch.float64)
    d = len(shape) - 2
    r = max(1, int(3 * sigma))
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, dtype=torch.float64) / sigma) ** 2)
    k /= k.sum()
    for a in range(d):
        n = x.shape[2 + a]
        # circular convolution along axis a via FFT (sizes are tiny in tests)
        K = torch.zeros(n, dtype=torch.float64)
        for i, kv in enumerate(k):
            K[(i - r) % n] += kv
        x = torch.fft.ifft(torch.fft.fft(x, dim=2 + a) * torch.fft.fft(K).reshape([-1 if i == 2 + a else 1 for i in range(x.dim())]), dim=2 + a).real
    x = x / x.abs().max() * amp
    return x.to(dtype)


def relerr(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def l2err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def tol_for(dtype, splat=False):
    if dtype == torch.float64:
        return TOL64 if not splat else 1e-11
    return TOL_SPLAT32 if splat else TOL32


class RefCuda:
    """ctypes view of oracle/_ref/libref_cuda.so: the reference's own CUDA kernels (compiled
    unmodified for sm_100a against an ATen stand-in), taking CUDA torch tensors."""

    def __init__(self):
        self.lib = ctypes.CDLL(REF_CUDA)

    @staticmethod
    def available():
        return os.path.exists(REF_CUDA)

    @staticmethod
    def _code(t):
        return {torch.float32: 0, torch.float64: 1}[t.dtype]

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _sh(shape):
        return (ctypes.c_long * len(shape))(*[int(s) for s in shape])

    def _chk(self, rc):
        torch.cuda.synchronize()
        if rc == 701:  # cudaErrorLaunchOutOfResources
            import pytest
            pytest.skip("the reference kernel cannot launch on sm_100: its fixed 32x32 block needs more "
                        "registers than an SM has (no __launch_bounds__ in the reference)")
        assert rc == 0, "reference CUDA call failed: %d" % rc

    def interp_fwd(self, I, u, dt=1.0):
        I, u = I.contiguous(), u.contiguous()
        d = I.dim() - 2
        N = max(I.shape[0], u.shape[0])
        out = torch.empty((N, I.shape[1]) + tuple(I.shape[2:]), dtype=I.dtype, device=I.device)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_interp_fwd(self._code(I), self._p(out), self._p(I), self._p(u), ctypes.c_long(N),
                                            ctypes.c_long(I.shape[0]), ctypes.c_long(I.shape[1]), d,
                                            self._sh(I.shape[2:]), ctypes.c_double(dt)))
        return out

    def interp_bwd(self, go, I, u, dt=1.0):
        go, I, u = go.contiguous(), I.contiguous(), u.contiguous()
        d = I.dim() - 2
        N = max(I.shape[0], u.shape[0])
        d_I, d_u = torch.empty_like(I), torch.empty_like(u)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_interp_bwd(self._code(I), self._p(d_I), self._p(d_u), self._p(go), self._p(I),
                                            self._p(u), ctypes.c_long(N), ctypes.c_long(I.shape[0]),
                                            ctypes.c_long(I.shape[1]), d, self._sh(I.shape[2:]),
                                            ctypes.c_double(dt), 1, 1))
        return d_I, d_u

    def jtvf_fwd(self, v, w, disp, trans):
        v, w = v.contiguous(), w.contiguous()
        d = v.dim() - 2
        out = torch.empty_like(v)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_jtvf_fwd(self._code(v), self._p(out), self._p(v), self._p(w),
                                          ctypes.c_long(v.shape[0]), ctypes.c_long(v.shape[1]), d,
                                          self._sh(v.shape[2:]), int(disp), int(trans)))
        return out

    def jtvf_bwd(self, go, v, w, disp, trans):
        go, v, w = go.contiguous(), v.contiguous(), w.contiguous()
        d = v.dim() - 2
        d_v, d_w = torch.empty_like(v), torch.empty_like(w)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_jtvf_bwd(self._code(v), self._p(d_v), self._p(d_w), self._p(go), self._p(v),
                                          self._p(w), ctypes.c_long(v.shape[0]), ctypes.c_long(v.shape[1]), d,
                                          self._sh(v.shape[2:]), int(disp), int(trans)))
        return d_v, d_w

    def jtvf_adj_fwd(self, z, w):
        z, w = z.contiguous(), w.contiguous()
        d = z.dim() - 2
        out = torch.empty_like(z)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_jtvf_adj_fwd(self._code(z), self._p(out), self._p(z), self._p(w),
                                              ctypes.c_long(z.shape[0]), ctypes.c_long(z.shape[1]), d,
                                              self._sh(z.shape[2:])))
        return out

    def jtvf_adj_bwd(self, go, z, w):
        go, z, w = go.contiguous(), z.contiguous(), w.contiguous()
        d = z.dim() - 2
        d_z, d_w = torch.empty_like(z), torch.empty_like(w)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_jtvf_adj_bwd(self._code(z), self._p(d_z), self._p(d_w), self._p(go), self._p(z),
                                              self._p(w), ctypes.c_long(z.shape[0]), ctypes.c_long(z.shape[1]), d,
                                              self._sh(z.shape[2:])))
        return d_z, d_w

    def fluid_operator(self, Fmv, inverse, cosl, sinl, alpha, beta, gamma):
        d = Fmv.dim() - 3
        pp = lambda ts, i: self._p(ts[i]) if i < len(ts) else None
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_fluid_operator(self._code(Fmv), self._p(Fmv), int(inverse), pp(cosl, 0), pp(sinl, 0),
                                                pp(cosl, 1), pp(sinl, 1), pp(cosl, 2), pp(sinl, 2),
                                                ctypes.c_double(alpha), ctypes.c_double(beta), ctypes.c_double(gamma),
                                                ctypes.c_long(Fmv.shape[0]), d, self._sh(Fmv.shape[2:2 + d])))

    def regrid_fwd(self, I, outshape, origin, spacing):
        I = I.contiguous()
        d = I.dim() - 2
        out = torch.empty(tuple(I.shape[:2]) + tuple(outshape), dtype=I.dtype, device=I.device)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_regrid_fwd(self._code(I), self._p(out), self._p(I), ctypes.c_long(I.shape[0]),
                                            ctypes.c_long(I.shape[1]), d, self._sh(I.shape[2:]), self._sh(outshape),
                                            (ctypes.c_double * d)(*origin), (ctypes.c_double * d)(*spacing)))
        return out

    def regrid_bwd(self, go, inshape, outshape, origin, spacing):
        go = go.contiguous()
        d = go.dim() - 2
        d_I = torch.empty(tuple(go.shape[:2]) + tuple(inshape), dtype=go.dtype, device=go.device)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_regrid_bwd(self._code(go), self._p(d_I), self._p(go), ctypes.c_long(go.shape[0]),
                                            ctypes.c_long(go.shape[1]), d, self._sh(inshape), self._sh(outshape),
                                            (ctypes.c_double * d)(*origin), (ctypes.c_double * d)(*spacing)))
        return d_I

    def affine_fwd(self, I, A, T):
        I, A, T = I.contiguous(), A.contiguous(), T.contiguous()
        d = I.dim() - 2
        N = A.shape[0]
        out = torch.empty((N, I.shape[1]) + tuple(I.shape[2:]), dtype=I.dtype, device=I.device)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_affine_interp_fwd(self._code(I), self._p(out), self._p(I), self._p(A), self._p(T),
                                                   ctypes.c_long(N), ctypes.c_long(I.shape[0]),
                                                   ctypes.c_long(I.shape[1]), d, self._sh(I.shape[2:])))
        return out

    def affine_bwd(self, go, I, A, T):
        go, I, A, T = go.contiguous(), I.contiguous(), A.contiguous(), T.contiguous()
        d = I.dim() - 2
        N = A.shape[0]
        d_I, d_A, d_T = torch.empty_like(I), torch.empty_like(A), torch.empty_like(T)
        torch.cuda.synchronize()
        self._chk(self.lib.refcu_affine_interp_bwd(self._code(I), self._p(d_I), self._p(d_A), self._p(d_T),
                                                   self._p(go), self._p(I), self._p(A), self._p(T), ctypes.c_long(N),
                                                   ctypes.c_long(I.shape[0]), ctypes.c_long(I.shape[1]), d,
                                                   self._sh(I.shape[2:])))
        return d_I, d_A, d_T


def ref_luts(shape, dtype, device="cuda"):
    """cos / sin LUTs exactly as lagomorph/metric.py:53-75 builds them (float64 numpy -> dtype)"""
    cshape = list(shape)
    cshape[-1] = cshape[-1] // 2 + 1
    cos, sin = [], []
    for (Nf, N) in zip(cshape[2:], shape[2:]):
        cos.append(torch.Tensor(2.0 * (1.0 - np.cos(2 * np.pi * np.arange(Nf) / N))).type(dtype).to(device))
        sin.append(torch.Tensor(np.sin(2.0 * np.pi * np.arange(Nf) / N)).type(dtype).to(device))
    return cos, sin


class RefPipeline:
    """The reference's Python layer on top of its OWN CUDA kernels (RefCuda): FluidMetric
    (lagomorph/metric.py:11-19, torch.rfft(normalized=True) == torch.fft.rfftn(norm="ortho") on cuFFT),
    Ad_star (adjrep.py:86-97), ad_star (:69-83), compose (deform.py:53-55), EPDiff_step
    (lddmm.py:39-44) and the expmap loop (:87-91). Checker / baseline only."""

    def __init__(self, rc, params):
        self.rc, self.params = rc, params
        self._luts = {}

    def fluid(self, mv, inverse):
        d = mv.dim() - 2
        dims = tuple(range(2, 2 + d))
        F = torch.view_as_real(torch.fft.rfftn(mv, dim=dims, norm="ortho")).contiguous()
        key = (tuple(mv.shape[2:]), mv.dtype)
        if key not in self._luts:
            self._luts[key] = ref_luts(mv.shape, mv.dtype)
        cos, sin = self._luts[key]
        self.rc.fluid_operator(F, inverse, cos, sin, *self.params)
        return torch.fft.irfftn(torch.view_as_complex(F), s=mv.shape[2:], dim=dims, norm="ortho")

    def Ad_star(self, phiinv, m):
        return self.rc.jtvf_fwd(phiinv, self.rc.interp_fwd(m, phiinv, 1.0), True, False)

    def ad_star(self, v, m):
        return self.rc.jtvf_fwd(v, m, False, True) - self.rc.jtvf_adj_fwd(m, v)

    def compose(self, u, v, ds, dt):
        return ds * u + dt * self.rc.interp_fwd(v, u, ds)

    def step(self, m0, dt, phiinv):
        m = self.Ad_star(phiinv, m0)
        v = self.fluid(m, True)
        return self.compose(v, phiinv, -dt, 1.0)

    def expmap(self, m0, num_steps, T=1.0):
        phiinv = torch.zeros_like(m0)
        for _ in range(num_steps):
            phiinv = self.step(m0, T / num_steps, phiinv)
        return phiinv


def baseline_momenta(N, shape, sigma=4.0, vmax=4.0, seed=1, params=(0.1, 0.0, 0.01), device="cuda"):
    """BASELINE.md section 4 momenta on the GPU: white noise low-passed by a separable Gaussian
    (sigma voxels, periodic), scaled so that max|sharp(m0)| * T = vmax voxels."""
    d = len(shape)
    m = torch.randn((N, d) + tuple(shape), generator=gen(seed)).to(device)
    dims = tuple(range(2, 2 + d))
    F = torch.fft.rfftn(m, dim=dims)
    for a, n in enumerate(shape):
        nf = n // 2 + 1 if a == d - 1 else n
        k = torch.fft.rfftfreq(n) if a == d - 1 else torch.fft.fftfreq(n)
        g = torch.exp(-2.0 * (np.pi * sigma * k[:nf]) ** 2).to(device)
        F = F * g.view([-1 if i == 2 + a else 1 for i in range(F.dim())])
    m = torch.fft.irfftn(F, s=tuple(shape), dim=dims).contiguous()
    # scale with the reference composition of sharp (cuFFT + the Fourier symbol computed in torch)
    Fm = torch.fft.rfftn(m, dim=dims, norm="ortho")
    lam = torch.full(Fm.shape[2:], params[2], device=device, dtype=torch.float64)
    for a, n in enumerate(shape):
        nf = n // 2 + 1 if a == d - 1 else n
        w = 2.0 * (1.0 - torch.cos(2 * np.pi * torch.arange(nf, dtype=torch.float64) / n)).to(device)
        lam = lam + params[0] * w.view([-1 if i == a else 1 for i in range(d)])
    v = torch.fft.irfftn(Fm / (lam * lam).float(), s=tuple(shape), dim=dims, norm="ortho")
    return m * (vmax / v.abs().max())
