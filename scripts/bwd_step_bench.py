#!/usr/bin/env python
"""Times expmap forward+backward (fused backward) per kernel with the library's profile mode.
  python scripts/bwd_step_bench.py [--size 256] [--n 4] [--steps 3] [--reps 3]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from lagomorph_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--noprofile", action="store_true")
a = ap.parse_args()
n, N = a.size, a.n
torch.cuda.set_device(0)
g = torch.Generator().manual_seed(1)
metric = lm.FluidMetric([0.1, 0.0, 0.01])
w = torch.randn((N, 3, n, n, n), generator=g).cuda()
m0 = metric.flat(metric.sharp(metric.sharp(w)))  # smooth
v0 = metric.sharp(m0)
m0 = (m0 * (4.0 / v0.abs().max())).requires_grad_(True)
del w, v0


def once():
    h = lm.expmap(metric, m0, num_steps=a.steps)
    loss = (h * h).sum()
    loss.backward()
    m0.grad = None


once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    once()
e1.record()
torch.cuda.synchronize()
print("fwd+bwd ms per shoot:", e0.elapsed_time(e1) / a.reps, "G voxel-steps/s fwd+bwd:",
      N * n ** 3 * a.steps / (e0.elapsed_time(e1) / a.reps * 1e-3) / 1e9)
if not a.noprofile:
    L.check(L.lib.lgm_profile_begin(L.stream_ptr(m0.device)))
    once()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    L.check(L.lib.lgm_profile_end(buf, len(buf)))
    prof = json.loads(buf.value.decode())
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"] if isinstance(kv[1], dict) else 0):
        print(k, v)
