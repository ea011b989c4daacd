"""Golden vectors: outputs of the REFERENCE's own CUDA kernels (run on a B200 through
oracle/_ref/libref_cuda.so, see tests/golden/make_golden.py), committed as
tests/golden/golden_ref_cuda.npz together with their inputs.

  * `not gpu` half: the CPU oracle reproduces them  (pins the oracle against the reference itself)
  * `gpu` half:     the CUDA product reproduces them (parity with the reference)
"""
import os

import numpy as np
import pytest
import torch

from util import ROOT, l2err, relerr

GOLD = os.path.join(ROOT, "tests", "golden", "golden_ref_cuda.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture missing")

DN = [("f32", torch.float32), ("f64", torch.float64)]


@pytest.fixture(scope="module")
def G():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def tol(dtype, splat=False, chain=False):
    if dtype == torch.float64:
        return 1e-10 if chain else 1e-11
    return 1e-4 if (splat or chain) else 1e-5


class Impl:
    """uniform view over the oracle (CPU) and the product (CUDA)"""

    def __init__(self, kind, lm=None, orc=None):
        self.kind, self.lm, self.orc = kind, lm, orc

    def dev(self, t):
        return t.cuda() if self.kind == "gpu" else t

    def interp(self, I, u, dt):
        return self.lm.interp(self.dev(I), self.dev(u), dt) if self.kind == "gpu" else self.orc.interp(I, u, dt)

    def interp_bwd(self, go, I, u, dt):
        if self.kind == "cpu":
            return self.orc.interp_backward(go, I, u, dt)
        Ic, uc = I.cuda().requires_grad_(True), u.cuda().requires_grad_(True)
        return torch.autograd.grad(self.lm.interp(Ic, uc, dt), [Ic, uc], go.cuda())

    def jtvf(self, v, w, d, t):
        if self.kind == "cpu":
            return self.orc.jtvf_forward(v, w, d, t)
        return self.lm.jacobian_times_vectorfield(v.cuda(), w.cuda(), displacement=bool(d), transpose=bool(t))

    def jtvf_bwd(self, go, v, w, d, t):
        if self.kind == "cpu":
            return self.orc.jtvf_backward(go, v, w, d, t)
        vc, wc = v.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
        out = self.lm.jacobian_times_vectorfield(vc, wc, displacement=bool(d), transpose=bool(t))
        return torch.autograd.grad(out, [vc, wc], go.cuda())

    def adj(self, z, w):
        if self.kind == "cpu":
            return self.orc.jtvf_adjoint_forward(z, w)
        return self.lm.jacobian_times_vectorfield_adjoint(z.cuda(), w.cuda())

    def adj_bwd(self, go, z, w):
        if self.kind == "cpu":
            return self.orc.jtvf_adjoint_backward(go, z, w)
        zc, wc = z.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
        return torch.autograd.grad(self.lm.jacobian_times_vectorfield_adjoint(zc, wc), [zc, wc], go.cuda())

    def metric(self, params):
        return self.lm.FluidMetric(params) if self.kind == "gpu" else self.orc.FluidMetric(params)

    def mod(self):
        return self.lm if self.kind == "gpu" else self.orc


def check_all(G, im):
    for dname, dtype in DN:
        # interp
        for bc in (0, 1):
            k = "interp3_%s_bc%d_" % (dname, bc)
            I, u, go = G[k + "I"], G[k + "u"], G[k + "go"]
            assert relerr(im.interp(I, u, 0.6), G[k + "out"]) <= tol(dtype), k
            dI, du = im.interp_bwd(go, I, u, 0.6)
            assert relerr(dI, G[k + "dI"]) <= tol(dtype, splat=True), k
            assert relerr(du, G[k + "du"]) <= tol(dtype), k
        # jacobian family
        for dim in (2, 3):
            k = "jtvf%d_%s_" % (dim, dname)
            v, w, go = G[k + "v"], G[k + "w"], G[k + "go"]
            for d in (0, 1):
                for t in (0, 1):
                    kk = k + "d%dt%d_" % (d, t)
                    if kk + "out" not in G:
                        continue
                    assert relerr(im.jtvf(v, w, d, t), G[kk + "out"]) <= tol(dtype), kk
                    dv, dw = im.jtvf_bwd(go, v, w, d, t)
                    assert relerr(dv, G[kk + "dv"]) <= tol(dtype) and relerr(dw, G[kk + "dw"]) <= tol(dtype), kk
            assert relerr(im.adj(v, w), G[k + "adj_out"]) <= tol(dtype), k
            dz, dw = im.adj_bwd(go, v, w)
            assert relerr(dz, G[k + "adj_dz"]) <= tol(dtype) and relerr(dw, G[k + "adj_dw"]) <= tol(dtype), k
        # fluid metric (reference: cuFFT + reference multiplier kernel)
        for key in [q for q in G if q.startswith("fluid") and q.endswith("_%s_m" % dname)]:
            k = key[:-1]
            m = G[key]
            for pi, params in enumerate(([0.1, 0.0, 0.01], [0.1, 0.01, 0.001])):
                met = im.metric(params)
                assert l2err(met.sharp(im.dev(m)), G[k + "p%d_sharp" % pi]) <= tol(dtype), (k, pi)
                assert l2err(met.flat(im.dev(m)), G[k + "p%d_flat" % pi]) <= tol(dtype), (k, pi)
        # adjoint representation and the 3-step shoot
        k = "epdiff3_%s_" % dname
        m0, phi = G[k + "m0"], G[k + "phi"]
        M = im.mod()
        assert relerr(M.Ad_star(im.dev(phi), im.dev(m0)), G[k + "Ad_star"]) <= tol(dtype), k
        assert relerr(M.ad_star(im.dev(phi), im.dev(m0)), G[k + "ad_star"]) <= tol(dtype), k
        assert relerr(M.compose(im.dev(phi), im.dev(m0), -0.1, 1.0), G[k + "compose"]) <= tol(dtype), k
        met = im.metric([0.1, 0.0, 0.01])
        assert relerr(M.expmap(met, im.dev(m0), num_steps=3), G[k + "expmap3"]) <= tol(dtype, chain=True), k


def check_regrid_affine(G, im, lm=None, orc=None):
    for dname, dtype in DN:
        for dim, sh, osh in ((2, (9, 7), (13, 12)), (3, (6, 9, 7), (11, 13, 12))):
            k = "regrid%d_%s_" % (dim, dname)
            I, go = G[k + "I"], G[k + "go"]
            origin = tuple((s - 1) * 0.5 for s in sh)
            spacing = tuple((a - 1) / (b - 1) for a, b in zip(sh, osh))
            if im.kind == "cpu":
                out = orc.regrid_forward(I, osh, origin, spacing)
                dI = orc.regrid_backward(go, sh, osh, origin, spacing)
            else:
                Ic = I.cuda().requires_grad_(True)
                out = lm.regrid(Ic, shape=osh)
                (dI,) = torch.autograd.grad(out, [Ic], go.cuda())
            assert relerr(out, G[k + "out"]) <= tol(dtype), k
            assert relerr(dI, G[k + "dI"]) <= tol(dtype, splat=True), k
        for dim in (2, 3):
            k = "affine%d_%s_" % (dim, dname)
            I, A, T, go = G[k + "I"], G[k + "A"], G[k + "T"], G[k + "go"]
            ftol = 5e-5 if dtype == torch.float32 else 1e-11
            if im.kind == "cpu":
                assert relerr(orc.affine_interp_forward(I, A, T), G[k + "out"]) <= ftol, k
            else:
                Ic, Ac, Tc = (t.cuda().requires_grad_(True) for t in (I, A, T))
                out = lm.affine_interp(Ic, Ac, Tc)
                assert relerr(out, G[k + "out"]) <= ftol, k
                dI, dA, dT = torch.autograd.grad(out, [Ic, Ac, Tc], go.cuda())
                gt = 1e-4 if dtype == torch.float32 else 1e-10
                assert relerr(dI, G[k + "dI"]) <= gt and relerr(dA, G[k + "dA"]) <= gt and relerr(dT, G[k + "dT"]) <= gt, k


def test_oracle_reproduces_reference_cuda(G, orc):
    im = Impl("cpu", orc=orc)
    check_all(G, im)
    check_regrid_affine(G, im, orc=orc)


@pytest.mark.gpu
def test_product_reproduces_reference_cuda(G, lm):
    im = Impl("gpu", lm=lm)
    check_all(G, im)
    check_regrid_affine(G, im, lm=lm)
