"""Generate tests/golden/golden_ref_cuda.npz on a GPU box.

Every output stored here was produced by the REFERENCE's own CUDA kernels
(/root/reference/lagomorph/extension/cuda/*.cu compiled unmodified for sm_100a against the ATen
stand-in in oracle/ref_cuda/shim -> oracle/_ref/libref_cuda.so) plus, for the FluidMetric and the
EPDiff shoot, the reference's Python-level composition (lagomorph/metric.py:11-19,
lagomorph/adjrep.py:86-97, lagomorph/deform.py:53-62, lagomorph/lddmm.py:39-44,87-91) with
torch.rfft(normalized=True) replaced by torch.fft.rfftn(norm="ortho") on the GPU (cuFFT).

Run (GPU box):  python tests/golden/make_golden.py gpurun_out/golden_ref_cuda.npz
then copy the file to tests/golden/. Inputs are stored next to the outputs, so the consumers
(tests/test_golden.py) need neither this script nor the reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import RefCuda, randn  # noqa: E402


def luts(shape, dtype):  # lagomorph/metric.py:53-75
    cshape = list(shape)
    cshape[-1] = cshape[-1] // 2 + 1
    cos, sin = [], []
    for (Nf, N) in zip(cshape[2:], shape[2:]):
        cos.append(torch.Tensor(2.0 * (1.0 - np.cos(2 * np.pi * np.arange(Nf) / N))).type(dtype).cuda())
        sin.append(torch.Tensor(np.sin(2.0 * np.pi * np.arange(Nf) / N)).type(dtype).cuda())
    return cos, sin


class RefPipeline:
    """The reference's Python layer on top of its own CUDA kernels."""

    def __init__(self, rc, params):
        self.rc, self.params = rc, params

    def fluid(self, mv, inverse):  # metric.py:11-19
        d = mv.dim() - 2
        dims = tuple(range(2, 2 + d))
        F = torch.view_as_real(torch.fft.rfftn(mv, dim=dims, norm="ortho")).contiguous()
        cos, sin = luts(mv.shape, mv.dtype)
        self.rc.fluid_operator(F, inverse, cos, sin, *self.params)
        return torch.fft.irfftn(torch.view_as_complex(F), s=mv.shape[2:], dim=dims, norm="ortho")

    def Ad_star(self, phiinv, m):  # adjrep.py:86-97
        return self.rc.jtvf_fwd(phiinv, self.rc.interp_fwd(m, phiinv, 1.0), True, False)

    def ad_star(self, v, m):  # adjrep.py:69-83
        return self.rc.jtvf_fwd(v, m, False, True) - self.rc.jtvf_adj_fwd(m, v)

    def compose(self, u, v, ds, dt):  # deform.py:53-55
        return ds * u + dt * self.rc.interp_fwd(v, u, ds)

    def step(self, m0, dt, phiinv):  # lddmm.py:39-44
        m = self.Ad_star(phiinv, m0)
        v = self.fluid(m, True)
        return self.compose(v, phiinv, -dt, 1.0)

    def expmap(self, m0, num_steps):  # lddmm.py:87-91
        phiinv = torch.zeros_like(m0)
        for _ in range(num_steps):
            phiinv = self.step(m0, 1.0 / num_steps, phiinv)
        return phiinv


def main(out):
    rc = RefCuda()
    G = {}

    def put(name, t):
        G[name] = t.detach().cpu().numpy()

    for dname, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        # ---- interp (3-D; the 2-D forward kernel cannot launch on sm_100, see tests/util.py) ----
        sh = (6, 7, 9)
        for bc in (0, 1):
            I = randn((1 if bc else 2, 2) + sh, dtype, 101)
            u = randn((2, 3) + sh, dtype, 102, 2.0)
            u[:, :, 0] -= 3.0
            u[..., -1] += 3.5
            go = randn((2, 2) + sh, dtype, 103)
            k = "interp3_%s_bc%d_" % (dname, bc)
            put(k + "I", I); put(k + "u", u); put(k + "go", go)
            put(k + "out", rc.interp_fwd(I.cuda(), u.cuda(), 0.6))
            dI, du = rc.interp_bwd(go.cuda(), I.cuda(), u.cuda(), 0.6)
            put(k + "dI", dI); put(k + "du", du)
        # ---- jacobian family ----
        for dim, sh in ((2, (9, 11)), (3, (6, 7, 9))):
            v, w, go = (randn((2, dim) + sh, dtype, s) for s in (104, 105, 106))
            k = "jtvf%d_%s_" % (dim, dname)
            put(k + "v", v); put(k + "w", w); put(k + "go", go)
            for disp in (0, 1):
                for trans in (0, 1):
                    if dim == 2 and trans:
                        continue  # reference 2-D transposed kernel: launch out of resources on sm_100
                    kk = k + "d%dt%d_" % (disp, trans)
                    put(kk + "out", rc.jtvf_fwd(v.cuda(), w.cuda(), disp, trans))
                    dv, dw = rc.jtvf_bwd(go.cuda(), v.cuda(), w.cuda(), disp, trans)
                    put(kk + "dv", dv); put(kk + "dw", dw)
            put(k + "adj_out", rc.jtvf_adj_fwd(v.cuda(), w.cuda()))
            dz, dw = rc.jtvf_adj_bwd(go.cuda(), v.cuda(), w.cuda())
            put(k + "adj_dz", dz); put(k + "adj_dw", dw)
        # ---- fluid metric through the reference pipeline ----
        for dim, sh in ((2, (6, 10)), (3, (4, 6, 10)), (3, (16, 16, 16))):
            m = randn((2, dim) + sh, dtype, 107)
            k = "fluid%d_%s_%s_" % (dim, "x".join(map(str, sh)), dname)
            put(k + "m", m)
            for pi, params in enumerate(([0.1, 0.0, 0.01], [0.1, 0.01, 0.001])):
                rp = RefPipeline(rc, params)
                put(k + "p%d_sharp" % pi, rp.fluid(m.cuda(), True))
                put(k + "p%d_flat" % pi, rp.fluid(m.cuda(), False))
        # ---- regrid ----
        for dim, sh, osh in ((2, (9, 7), (13, 12)), (3, (6, 9, 7), (11, 13, 12))):
            I = randn((2, 2) + sh, dtype, 108)
            origin = tuple((s - 1) * 0.5 for s in sh)
            spacing = tuple((a - 1) / (b - 1) for a, b in zip(sh, osh))
            k = "regrid%d_%s_" % (dim, dname)
            out_ = rc.regrid_fwd(I.cuda(), osh, origin, spacing)
            go = randn(tuple(out_.shape), dtype, 109)
            put(k + "I", I); put(k + "go", go); put(k + "out", out_)
            put(k + "dI", rc.regrid_bwd(go.cuda(), sh, osh, origin, spacing))
        # ---- affine ----
        for dim, sh in ((2, (9, 7)), (3, (6, 9, 7))):
            I = randn((3, 2) + sh, dtype, 110)
            A = torch.eye(dim, dtype=dtype).repeat(3, 1, 1) + randn((3, dim, dim), dtype, 111, 0.1)
            T = randn((3, dim), dtype, 112, 1.5)
            go = randn((3, 2) + sh, dtype, 113)
            k = "affine%d_%s_" % (dim, dname)
            put(k + "I", I); put(k + "A", A); put(k + "T", T); put(k + "go", go)
            put(k + "out", rc.affine_fwd(I.cuda(), A.cuda(), T.cuda()))
            dI, dA, dT = rc.affine_bwd(go.cuda(), I.cuda(), A.cuda(), T.cuda())
            put(k + "dI", dI); put(k + "dA", dA); put(k + "dT", dT)
        # ---- adjoint representation + a 3-step EPDiff shoot, 16^3 ----
        params = [0.1, 0.0, 0.01]
        rp = RefPipeline(rc, params)
        sh = (16, 16, 16)
        m0 = randn((2, 3) + sh, dtype, 114).cuda()
        m0 = m0 * (3.0 / rp.fluid(m0, True).abs().max())
        phi = randn((2, 3) + sh, dtype, 115, 1.5).cuda()
        k = "epdiff3_%s_" % dname
        put(k + "m0", m0); put(k + "phi", phi)
        put(k + "Ad_star", rp.Ad_star(phi, m0))
        put(k + "ad_star", rp.ad_star(phi, m0))
        put(k + "compose", rp.compose(phi, m0, -0.1, 1.0))
        put(k + "expmap3", rp.expmap(m0, 3))
    np.savez_compressed(out, **G)
    print("wrote", out, len(G), "arrays", sum(a.nbytes for a in G.values()) >> 10, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_ref_cuda.npz"))
