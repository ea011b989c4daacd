"""GPU: whole atlas epochs of the product against the CPU oracle's restatement of the reference
drivers from the same initial state -- LDDMMAtlasBuilder (lagomorph/lddmm.py:287-358) and
affine_atlas (lagomorph/affine.py:345-405): atlas image, momenta / poses and losses after each epoch."""
import pytest
import torch

from util import RefCuda, randn, relerr

pytestmark = pytest.mark.gpu


def _blobs(S, shape, dtype, seed=1):
    g = torch.Generator().manual_seed(seed)
    grid = torch.meshgrid(*[torch.arange(n, dtype=torch.float64) for n in shape], indexing="ij")
    out = []
    for s in range(S):
        c = [n / 2 + (torch.rand(1, generator=g).item() - 0.5) * n / 4 for n in shape]
        r2 = sum((gg - cc) ** 2 for gg, cc in zip(grid, c))
        out.append(torch.exp(-r2 / (2 * (min(shape) / 5) ** 2)))
    return torch.stack(out).unsqueeze(1).to(dtype)


@pytest.mark.parametrize("shape", [(32, 32), (16, 16, 16)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.float64, 1e-9)])
@pytest.mark.parametrize("freq,precond", [(0, False), (1, True)])
def test_lddmm_atlas_epochs_vs_oracle(lm, orc, shape, dtype, tol, freq, precond):
    S, B, steps = 4, 2, 3
    params = [0.1, 0.0, 0.1]
    data = _blobs(S, shape, dtype)
    kw = dict(reg_weight=1e-2, learning_rate_pose=2.0)
    b = lm.LDDMMAtlasBuilder(data, num_epochs=1, batch_size=B, lddmm_integration_steps=steps, learning_rate_image=0.5,
                             image_update_freq=freq, momentum_preconditioning=precond,
                             metric=lm.FluidMetric(params), **kw)
    b.initialize()
    om = orc.FluidMetric(params)
    I = data.mean(0, keepdim=True)
    assert relerr(b.I, I) <= (1e-6 if dtype == torch.float32 else 1e-14)
    ms = [torch.zeros(B, len(shape), *shape, dtype=dtype) for _ in range(S // B)]
    batches = [data[i:i + B] for i in range(0, S, B)]
    for ep in range(2):
        l, r = b.epoch()
        I, ms, lo, ro = orc.lddmm_epoch(om, I, ms, batches, S, learning_rate_image=0.5, image_update_freq=freq,
                                        integration_steps=steps, momentum_preconditioning=precond, **kw)
        assert abs(l - lo) <= tol * abs(lo) and abs(r - ro) <= tol * max(abs(ro), 1e-12)
        assert relerr(b.I, I) <= tol
        for mg, mo in zip(b.ms, ms):
            assert relerr(mg, mo) <= 10 * tol  # gradients through 3 steps of splats (atomic order)


@pytest.mark.parametrize("shape", [(32, 32), (16, 16, 16)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.float64, 1e-9)])
def test_affine_atlas_epochs_vs_oracle(lm, orc, shape, dtype, tol):
    S, B, d = 4, 2, len(shape)
    data = _blobs(S, shape, dtype, seed=3)
    As = (0.02 * randn((S, d, d), dtype, 4))
    Ts = (0.5 * randn((S, d), dtype, 5))
    kw = dict(affine_steps=2, reg_weightA=1e-2, reg_weightT=1e-3, learning_rate_A=1e-3, learning_rate_T=5e-2,
              learning_rate_I=0.5)
    I0 = data.mean(0, keepdim=True)
    Ig, Ag, Tg, el, il = lm.affine_atlas(data, As.clone(), Ts.clone(), I=I0.clone(), num_epochs=2, batch_size=B,
                                         image_update_freq=0, **kw)
    I, A, T = I0, As, Ts
    elo, ilo = [], []
    for ep in range(2):
        I, A, T, l, its = orc.affine_atlas_epoch(I, A, T, data, B, S, image_update_freq=0, **kw)
        elo.append(l)
        ilo.extend(its)
    assert relerr(Ig, I) <= tol and relerr(Ag, A) <= 10 * tol and relerr(Tg, T) <= 10 * tol
    assert all(abs(a - b) <= tol * abs(b) for a, b in zip(el, elo))
    assert len(il) == len(ilo) and all(abs(a - b) <= tol * abs(b) for a, b in zip(il, ilo))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_affine_backward_vs_reference_cuda(orc, dtype):
    """pins the oracle's new affine backward to the reference's own kernel (cuda/affine.cu:330-610)"""
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    rc = RefCuda()
    for sh in [(9, 8), (6, 7, 5)]:
        d = len(sh)
        I, go = randn((2, 1) + sh, dtype, 1), randn((2, 1) + sh, dtype, 2)
        A = torch.eye(d, dtype=dtype)[None] + 0.05 * randn((2, d, d), dtype, 3)
        T = 0.7 * randn((2, d), dtype, 4)
        want = rc.affine_bwd(go.cuda(), I.cuda(), A.cuda(), T.cuda())
        got = orc.affine_interp_backward(go, I, A, T)
        tol = 1e-5 if dtype == torch.float32 else 1e-12
        for g, w in zip(got, want):
            assert relerr(g, w) <= tol
