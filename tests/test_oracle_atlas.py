"""CPU: the oracle's restatement of the two atlas drivers' steps is self-consistent -- the
hand-written backward kernels it chains (interp / Jacobian / fluid / affine) give the gradient of
the loss they implement (central finite differences in fp64), so that tests/test_atlas_oracle_gpu.py
can compare whole epochs of the product against it."""
import torch

from util import randn


def _blobs(S, shape, dtype, seed=1):
    g = torch.Generator().manual_seed(seed)
    grid = torch.meshgrid(*[torch.arange(n, dtype=torch.float64) for n in shape], indexing="ij")
    out = []
    for s in range(S):
        c = [n / 2 + (torch.rand(1, generator=g).item() - 0.5) * n / 4 for n in shape]
        r2 = sum((gg - cc) ** 2 for gg, cc in zip(grid, c))
        out.append(torch.exp(-r2 / (2 * (min(shape) / 5) ** 2)))
    return torch.stack(out).unsqueeze(1).to(dtype)


def test_affine_backward_matches_finite_differences(orc):
    for sh in [(9, 8), (6, 7, 5)]:
        d = len(sh)
        I = _blobs(2, sh, torch.float64, 3)
        go = randn((2, 1) + sh, torch.float64, 4)
        A = torch.eye(d, dtype=torch.float64)[None] + 0.05 * randn((2, d, d), torch.float64, 5)
        T = 0.3 * randn((2, d), torch.float64, 6)
        f = lambda A_, T_, I_: (orc.affine_interp_forward(I_, A_, T_) * go).sum().item()
        d_I, d_A, d_T = orc.affine_interp_backward(go, I, A, T)
        eps = 1e-6
        for ix in [(0, 0, 0), (1, d - 1, 0), (0, 1, d - 1)]:
            Ap, Am = A.clone(), A.clone()
            Ap[ix] += eps
            Am[ix] -= eps
            fd = (f(Ap, T, I) - f(Am, T, I)) / (2 * eps)
            assert abs(fd - d_A[ix].item()) <= 1e-6 * max(1.0, abs(fd))
        for ix in [(0, 0), (1, d - 1)]:
            Tp, Tm = T.clone(), T.clone()
            Tp[ix] += eps
            Tm[ix] -= eps
            fd = (f(A, Tp, I) - f(A, Tm, I)) / (2 * eps)
            assert abs(fd - d_T[ix].item()) <= 1e-6 * max(1.0, abs(fd))
        # d_I is the exact adjoint of the (linear in I) forward
        J = randn(I.shape, torch.float64, 7)
        assert abs((orc.affine_interp_forward(J, A, T) * go).sum().item() - (J * d_I).sum().item()) <= 1e-10
        # broadcast image: one d_I for both subjects
        dIb, _, _ = orc.affine_interp_backward(go, I[:1], A, T)
        assert dIb.shape[0] == 1
        assert abs((orc.affine_interp_forward(J[:1], A, T) * go).sum().item() - (J[:1] * dIb).sum().item()) <= 1e-10


def test_lddmm_step_gradient_matches_finite_differences(orc):
    sh = (8, 8)
    data = _blobs(2, sh, torch.float64)
    metric = orc.FluidMetric([0.5, 0.0, 0.5])
    I = data.mean(0, keepdim=True)
    m = 0.05 * randn((2, 2) + sh, torch.float64, 2)
    kw = dict(integration_steps=2, reg_weight=0.1, learning_rate_pose=1.0)

    def loss_fn(mm):
        h = orc.expmap(metric, mm, num_steps=2)
        Idef = orc.interp(I, h)
        v = metric.sharp(mm)
        return (((Idef - data) ** 2).sum() / data.numel() + 0.1 * (v * mm).sum() / data.numel()).item()

    m_new, loss, reg, gI = orc.lddmm_step(metric, I, m, data, 2, **kw)
    g = m - m_new  # learning rate 1: the update is the gradient
    assert abs(loss - loss_fn(m)) <= 1e-12
    eps = 1e-6
    for ix in [(0, 0, 3, 4), (1, 1, 5, 2), (0, 1, 0, 0), (1, 0, 7, 7)]:
        mp, mm_ = m.clone(), m.clone()
        mp[ix] += eps
        mm_[ix] -= eps
        fd = (loss_fn(mp) - loss_fn(mm_)) / (2 * eps)
        assert abs(fd - g[ix].item()) <= 1e-6 * max(1.0, abs(fd)), (ix, fd, g[ix].item())
    # image gradient: loss is quadratic in I, check one direction
    J = randn(I.shape, torch.float64, 9)
    h = orc.expmap(metric, m, num_steps=2)
    r = orc.interp(I, h) - data
    fdI = (2 * r * orc.interp(J, h)).sum().item() / data.numel()
    assert abs(fdI - (gI * J).sum().item()) <= 1e-10


def test_atlas_epochs_decrease_loss(orc):
    sh = (16, 16)
    data = _blobs(4, sh, torch.float32)
    metric = orc.FluidMetric([0.1, 0.0, 0.1])
    I = data.mean(0, keepdim=True)
    ms = [torch.zeros(2, 2, *sh), torch.zeros(2, 2, *sh)]
    batches = [data[:2], data[2:]]
    losses = []
    for ep in range(3):
        I, ms, l, r = orc.lddmm_epoch(metric, I, ms, batches, 4, learning_rate_image=0.5, integration_steps=3,
                                      reg_weight=1e-2, learning_rate_pose=2.0)
        losses.append(l)
    assert losses[-1] < losses[0]
    As, Ts = torch.zeros(4, 2, 2), torch.zeros(4, 2)
    I = data.mean(0, keepdim=True)
    al = []
    for ep in range(3):
        I, As, Ts, l, _ = orc.affine_atlas_epoch(I, As, Ts, data, 2, 4, learning_rate_A=1e-3, learning_rate_T=1e-1,
                                                 learning_rate_I=1.0)
        al.append(l)
    assert al[-1] < al[0]
