"""Quarter-slab FFT path (csrc/qslab.cuh) against a torch.fft restatement of FluidMetric on the GPU and,
with LGM_NO_QSLAB=1 in the environment, the cluster slab kernels it replaces. Prints errors and timings.
Run on the GPU box:  python scripts/qslab_check.py [X ...]   (not a product path; a measurement helper)"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagomorph_b200 as lm  # noqa: E402


def torch_fluid(m, params, inverse):
    """metric.py:9-19 + cuda/metric.cu:220-306 for beta == 0 with torch.fft on the device (fp64)."""
    alpha, beta, gamma = params
    assert beta == 0
    sh = m.shape[2:]
    F = torch.fft.rfftn(m.double(), dim=(2, 3, 4), norm="ortho")
    ws = []
    for a, n in enumerate(sh):
        k = torch.arange(n if a < 2 else n // 2 + 1, device=m.device, dtype=torch.float64)
        w = (2.0 * (1.0 - torch.cos(2.0 * math.pi * k / n))).float().double()
        ws.append(w)
    sw = (ws[0][:, None, None] + ws[1][None, :, None] + ws[2][None, None, :]).float()
    lam = (gamma + alpha * sw.double()).float()
    L = (lam * lam).double()
    F = F / L if inverse else F * L
    return torch.fft.irfftn(F, s=sh, dim=(2, 3, 4), norm="ortho").float()


def main():
    xs = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    params = [0.1, 0.0, 0.01]
    met = lm.FluidMetric(params)
    tag = "cluster" if os.environ.get("LGM_NO_QSLAB") else "qslab"
    for X in xs:
        N = 2 if X == 256 else 4
        m = torch.randn((N, 3, X, 256, 256), generator=torch.Generator().manual_seed(5)).cuda()
        for name, inv in (("sharp", True), ("flat", False)):
            out = getattr(met, name)(m)
            ref = torch_fluid(m[:1], params, inv)
            err = ((out[:1] - ref).norm() / ref.norm()).item()
            emax = ((out[:1] - ref).abs().max() / ref.abs().max()).item()
            print("%s X=%d %s: rel L2 err %.3e, max err / max %.3e" % (tag, X, name, err, emax), flush=True)
        rt = ((met.flat(met.sharp(m)) - m).norm() / m.norm()).item()
        print("%s X=%d flat(sharp(m)) - m: %.3e" % (tag, X, rt), flush=True)
        # timing: sharp on 8 x 3 x X x 256 x 256 (C3 share at X = 256)
        Nb = 8
        mb = torch.randn((Nb, 3, X, 256, 256), device="cuda")
        for _ in range(3):
            met.sharp(mb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            met.sharp(mb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        V = Nb * X * 256 * 256
        print("%s X=%d sharp %d subjects: %.3f ms (%.3f of the 24 B/voxel HBM figure at 6551.7 GB/s)" % (
            tag, X, Nb, ms, 24.0 * V / (ms * 1e-3) / 6551.7e9), flush=True)
        import ctypes, json
        from lagomorph_b200 import _lib as L
        buf = ctypes.create_string_buffer(1 << 16)
        L.check(L.lib.lgm_profile_begin(L.stream_ptr(mb.device)))
        for _ in range(5):
            met.sharp(mb)
        L.check(L.lib.lgm_profile_end(buf, len(buf)))
        print("   per-kernel:", buf.value.decode(), flush=True)
        del m, mb, out, ref


if __name__ == "__main__":
    main()
