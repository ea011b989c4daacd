"""sharp() timing on grids that are not powers of two (mixed-radix path) next to 128^3 / 256^3 and to
cuFFT doing the same work (torch.fft.rfftn + irfftn, no multiplier): python scripts/mixed_bench.py"""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from lagomorph_b200 import _lib as L
dev = torch.device("cuda")
met = lm.FluidMetric([0.1, 0.0, 0.01])
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for N, sh in [(16, (128, 128, 128)), (8, (192, 192, 192)), (8, (160, 192, 160)), (4, (182, 218, 182)), (2, (256, 256, 256)), (8, (120, 144, 120))]:
    m = torch.randn((N, 3) + sh, device=dev)
    ms = t(lambda: met.sharp(m))
    cu = t(lambda: torch.fft.irfftn(torch.fft.rfftn(m, dim=(2, 3, 4), norm="ortho"), s=sh, dim=(2, 3, 4), norm="ortho"))
    vox = N * sh[0] * sh[1] * sh[2]
    buf = ctypes.create_string_buffer(1 << 16)
    L.check(L.lib.lgm_profile_begin(L.stream_ptr(dev)))
    met.sharp(m)
    L.check(L.lib.lgm_profile_end(buf, len(buf)))
    ks = json.loads(buf.value.decode())
    print("%s x%d: sharp %.3f ms = %.0f GB/s algorithmic (24 B/voxel) | cuFFT rfftn+irfftn %.3f ms | %s" % (
        sh, N, ms, vox * 24 / ms / 1e6, cu, " ".join("%s %.3f" % (k, v["ms"]) for k, v in sorted(ks.items()))))
    del m
