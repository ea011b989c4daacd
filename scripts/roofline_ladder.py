"""Practical ceilings for 36 B/voxel kernels at the C2 size (what a plain streaming kernel reaches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
dev = torch.device("cuda")
N, n = 16, 128
a = torch.randn(N, 3, n, n, n, device=dev)
b = torch.randn(N, 3, n, n, n, device=dev)
out = torch.empty_like(a)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
V = N * n ** 3
ms = t(lambda: torch.mul(a, 2.0, out=out)); print("copy-like  24 B/voxel: %.3f ms  %.0f GB/s" % (ms, 24 * V / ms / 1e6))
ms = t(lambda: torch.add(a, b, out=out)); print("add        36 B/voxel: %.3f ms  %.0f GB/s" % (ms, 36 * V / ms / 1e6))
z = torch.zeros_like(a)
ms = t(lambda: lm.compose(z, b, -0.1, 1.0)); print("compose u=0 (aligned gathers): %.3f ms  %.0f GB/s" % (ms, 36 * V / ms / 1e6))
u = lm.FluidMetric([0.1, 0, 0.01]).sharp(a); u = u * (4.0 / u.abs().max())
ms = t(lambda: lm.compose(u, b, -0.1, 1.0)); print("compose smooth u, ds=-0.1   : %.3f ms  %.0f GB/s" % (ms, 36 * V / ms / 1e6))
ms = t(lambda: lm.compose(u, b, 1.0, 1.0)); print("compose smooth u, ds=1      : %.3f ms  %.0f GB/s" % (ms, 36 * V / ms / 1e6))
ms = t(lambda: lm.interp(b, z)); print("interp u=0 C=3              : %.3f ms  %.0f GB/s" % (ms, 36 * V / ms / 1e6))
