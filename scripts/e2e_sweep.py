"""expmap at small batches and expmap_host chunkings (what bounds the host-buffer path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
dev = torch.device("cuda")
shape, nsteps = (128,) * 3, 10
V = 128 ** 3
metric = lm.FluidMetric([0.1, 0.0, 0.01])
g = torch.Generator().manual_seed(1)
m_host = torch.randn((16, 3) + shape, generator=g).pin_memory()
m0 = m_host.to(dev)
s = 4.0 / metric.sharp(m0).abs().max().item()
m0.mul_(s); m_host.mul_(s)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1, 2, 3, 4, 8, 16):
    ms = t(lambda: lm.expmap(metric, m0[:n], num_steps=nsteps))
    print("expmap N=%2d: %.3f ms  %.3f ms/subject  %.2f G" % (n, ms, ms / n, n * V * nsteps / ms / 1e6))
out = torch.empty_like(m_host).pin_memory()
ms = t(lambda: m0.copy_(m_host, non_blocking=True)); print("H2D 384 MiB: %.2f ms = %.1f GB/s" % (ms, 0.4027 / ms * 1e3))
ms = t(lambda: out.copy_(m0, non_blocking=True)); print("D2H 384 MiB: %.2f ms = %.1f GB/s" % (ms, 0.4027 / ms * 1e3))
cfgs = [("auto", True), (2, True), ([1, 2, 3, 3, 3, 2, 1, 1], True), ([1, 2, 3, 4, 3, 2, 1], True), ([1, 1, 2, 3, 3, 3, 2, 1], True),
        ([1, 2, 4, 4, 3, 1, 1], True), ([1, 2, 4, 5, 2, 1, 1], True)]
res = {i: [] for i in range(len(cfgs))}
for rep in range(3):
    for i, (chunk, gr) in enumerate(cfgs):
        res[i].append(t(lambda: lm.expmap_host(metric, m_host, num_steps=nsteps, out=out, device=dev, chunk=chunk, graphs=gr)))
for i, (chunk, gr) in enumerate(cfgs):
    ms = sorted(res[i])[1]
    print("expmap_host chunk=%s graphs=%s: %s ms  median %.2f G" % (chunk, gr, ["%.2f" % x for x in res[i]], 16 * V * nsteps / ms / 1e6))
