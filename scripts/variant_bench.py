"""Time the EPDiff step kernels of the library selected by LGM_LIB_PATH (kernel experiments).
usage: python scripts/variant_bench.py [c2|c3]"""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from lagomorph_b200 import _lib as L
dev = torch.device("cuda")
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
N, shape, nsteps = {"c2": (16, (128,) * 3, 10), "c3": (8, (256,) * 3, 5)}[wl]
g = torch.Generator().manual_seed(1)
m0 = torch.randn((N, 3) + shape, generator=g).to(dev)
metric = lm.FluidMetric([0.1, 0.0, 0.01])
m0.mul_(4.0 / metric.sharp(m0).abs().max().item())
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: lm.expmap(metric, m0, num_steps=nsteps), 5)
V = N * shape[0] * shape[1] * shape[2]
buf = ctypes.create_string_buffer(1 << 16)
L.check(L.lib.lgm_profile_begin(L.stream_ptr(dev)))
for _ in range(3): lm.expmap(metric, m0, num_steps=nsteps)
L.check(L.lib.lgm_profile_end(buf, len(buf)))
ks = json.loads(buf.value.decode())
print("%s %s: shoot %.3f ms = %.2f G voxel-steps/s | " % (os.environ.get("LGM_LIB_PATH", "default").split("/")[-1], wl, ms, V * nsteps / ms / 1e6) +
      " ".join("%s %.4f" % (k, v["ms"] / (3 * nsteps)) for k, v in sorted(ks.items())))
