#!/usr/bin/env python
"""Kernel-time breakdown of one atlas epoch (config-3-like) under torch.profiler: every CUDA
kernel (this library's and torch's own pointwise / reduction kernels) summed by name."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--subjects", type=int, default=8)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--out", default="gpurun_out/atlas_profile.json")
a = ap.parse_args()
n, S = a.size, a.subjects
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


class Synth:
    def __len__(self):
        return S

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(100 + i)
        ax = torch.arange(n, dtype=torch.float32)
        c = n / 2 + (torch.rand(3, generator=g) - 0.5) * n / 8
        e = [torch.exp(-((ax - c[d]) ** 2) / (2 * (n / 6) ** 2)) for d in range(3)]
        return (e[0][:, None, None] * e[1][None, :, None] * e[2][None, None, :]).unsqueeze(0)


b = lm.LDDMMAtlasBuilder(Synth(), num_epochs=1, batch_size=a.batch, lddmm_integration_steps=5,
                         reg_weight=1e-2, learning_rate_pose=1.0, learning_rate_image=0.1, device=dev)
b.initialize()
b.epoch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b.epoch(); e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    b.epoch()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", 0) or getattr(ev, "cuda_time_total", 0)
    if t > 0 and ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.key, t / 1e3, ev.count))
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"epoch wall {wall:.2f} ms (unprofiled); kernel sum {tot:.2f} ms; {S} subjects of {n}^3")
for k, t, c in rows[:40]:
    print(f"{t:9.3f} ms {100*t/tot:5.1f}% x{c:4d}  {k[:110]}")
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump({"epoch_ms": wall, "kernel_sum_ms": tot, "size": n, "subjects": S,
           "kernels": [{"name": k, "ms": t, "count": c} for k, t, c in rows]}, open(a.out, "w"), indent=1)
