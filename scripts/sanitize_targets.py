"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_targets.py
The 256 x 256 slabs go through the 4-CTA cluster kernels (DSMEM all-to-all, csrc/fluid.cu) for beta != 0 or
X < 64, through the quarter-slab kernels (csrc/qslab.cuh) otherwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm

torch.manual_seed(0)
dev = "cuda"
for params in ([0.1, 0.0, 0.01], [0.1, 0.02, 0.01]):
    met = lm.FluidMetric(params)
    for sh in [(8, 256, 256), (16, 128, 128), (8, 32, 64), (128, 128), (12, 10, 14), (24, 40, 48), (48, 96)]:
        d = len(sh)
        m = torch.randn((2, d) + sh, device=dev)
        v = met.sharp(m)
        back = met.flat(v)
        print("fluid", params, sh, float((back - m).abs().max()))
met = lm.FluidMetric([0.1, 0.0, 0.01])
mq = torch.randn((1, 3, 64, 256, 256), device=dev)     # quarter-slab kernels + their X pass
print("qslab", float((met.flat(met.sharp(mq)) - mq).abs().max()))
print("qslab shoot", float(lm.expmap(met, mq * (2.0 / met.sharp(mq).abs().max()), num_steps=2).abs().max()))
del mq
# Ad_star through the shared-memory ring (stencil operand staged by bulk async copies), every row length
for shp in [(6, 10, 32), (9, 12, 64), (20, 20, 128), (5, 9, 256)]:
    phi = (torch.rand((2, 3) + shp, device=dev) - 0.5) * 6.0
    mm = torch.randn((2, 3) + shp, device=dev)
    print("adstar ring", shp, float(lm.Ad_star(phi, mm).abs().max()))
sh = (16, 24, 32)
m0 = torch.randn((2, 3) + sh, device=dev)
m0 = m0 * (3.0 / met.sharp(m0).abs().max())
h = lm.expmap(met, m0, num_steps=3)
print("shoot", float(h.abs().max()))
# compose through the shared-memory ring (bulk async copies + mbarriers) for every supported row length,
# incl. warps that fall back to the global gather and tiles at the volume border
for shp in [(6, 10, 32), (9, 12, 64), (7, 20, 128), (5, 9, 256)]:
    u = (torch.rand((2, 3) + shp, device=dev) - 0.5) * 6.0
    u[0, :, :, :2, 3:7] *= 8.0
    v = torch.randn((2, 3) + shp, device=dev)
    print("ring", shp, float(lm.compose(u, v, ds=-0.1, dt=1.0).abs().max()))
m0g = m0.clone().requires_grad_(True)
I = torch.randn((1, 1) + sh, device=dev, requires_grad=True)
hh = lm.expmap(met, m0g, num_steps=2)
loss = (lm.interp(I, hh) ** 2).sum() + (lm.ad_star(hh, m0g) * lm.ad(hh, m0g)).sum()
loss.backward()
print("bwd", float(m0g.grad.abs().max()), float(I.grad.abs().max()))
A = (torch.eye(3, device=dev)[None] + 0.05 * torch.randn(2, 3, 3, device=dev)).requires_grad_(True)
T = torch.randn(2, 3, device=dev, requires_grad=True)
out = lm.affine_interp(I, A, T)
out.sum().backward()
r = lm.regrid(I.detach(), shape=(20, 20, 20))
print("affine", float(A.grad.abs().max()), tuple(r.shape))
m2 = torch.randn(2, 2, 64, 64, device=dev)
print("2d", float(lm.expmap(lm.FluidMetric([0.1, 0.0, 0.1]), m2 * 0.01, num_steps=2).abs().max()))
torch.cuda.synchronize()
print("done")
