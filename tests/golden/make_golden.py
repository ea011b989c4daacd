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
from util import RefCuda, RefPipeline, randn, ref_luts as luts  # noqa: E402


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
