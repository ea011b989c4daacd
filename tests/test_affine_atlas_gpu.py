"""GPU: affine atlas driver (lagomorph_b200/affine_atlas.py, restating lagomorph/affine.py:288-438)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _blobs(shape, shifts, sigma=4.0):
    axes = [torch.arange(n, dtype=torch.float32) for n in shape]
    out = []
    for s in shifts:
        e = [torch.exp(-((ax - (n - 1) / 2 - float(sd)) ** 2) / (2 * sigma ** 2)) for ax, n, sd in zip(axes, shape, s)]
        img = e[0]
        for f in e[1:]:
            img = img.unsqueeze(-1) * f
        out.append(img.unsqueeze(0))
    return torch.stack(out)  # (S, 1, ...)


@pytest.mark.parametrize("shape", [(32, 32), (16, 16, 32)])
def test_one_epoch_equals_manual_step(lm, shape):
    """one epoch, one batch: A and T move by -lr * the autograd gradients of the reference's loss, the
    atlas by -lr_I * its gradient (affine.py:355-405)"""
    d = len(shape)
    S = 4
    g = torch.Generator().manual_seed(3)
    data = _blobs(shape, (torch.rand(S, d, generator=g) - 0.5) * 3)
    As = 0.02 * torch.randn(S, d, d, generator=g)
    Ts = 0.5 * torch.randn(S, d, generator=g)
    lrA, lrT, lrI = 0.3, 2.0, 5.0
    I1, A1, T1, el, il = lm.affine_atlas(data, As.clone(), Ts.clone(), num_epochs=1, batch_size=S, learning_rate_A=lrA,
                                         learning_rate_T=lrT, learning_rate_I=lrI, reg_weightA=0.1, reg_weightT=0.2)
    # manual
    dev = torch.device("cuda")
    img = data.to(dev)
    I = img.mean(0, keepdim=True).requires_grad_(True)
    A = As.to(dev).requires_grad_(True)
    T = Ts.to(dev).requires_grad_(True)
    eye = torch.eye(d, device=dev).view(1, d, d)
    Idef = lm.affine_interp(I, A + eye, T)
    nvox = float(torch.tensor(shape).prod())
    loss = (((Idef - img) ** 2).sum() / nvox + 0.5 * 0.1 * (A * A).sum() + 0.5 * 0.2 * (T * T).sum()) / S
    gI, gA, gT = torch.autograd.grad(loss, [I, A, T])
    assert torch.allclose(A1.to(dev), A.detach() - lrA * gA, rtol=1e-5, atol=1e-7)
    assert torch.allclose(T1.to(dev), T.detach() - lrT * gT, rtol=1e-5, atol=1e-7)
    assert torch.allclose(I1, I.detach() - lrI * gI, rtol=1e-5, atol=1e-7)
    assert abs(el[0] - loss.item()) <= 1e-5 * abs(loss.item()) and len(il) == 1


def test_translations_are_recovered_and_standardized(lm):
    """subjects are one blob shifted by known amounts: the loss falls, the translations approach the
    (mean-centred) negative shifts, and StandardizedDataset undoes the pose"""
    shape = (32, 32)
    shifts = torch.tensor([[2.0, -1.0], [-2.0, 1.0], [1.0, 2.0], [-1.0, -2.0]])
    data = _blobs(shape, shifts)
    S = len(shifts)
    As, Ts = torch.zeros(S, 2, 2), torch.zeros(S, 2)
    I, As, Ts, el, il = lm.affine_atlas(data, As, Ts, num_epochs=60, batch_size=2, learning_rate_A=0.0,
                                        learning_rate_T=40.0, learning_rate_I=2.0)
    assert el[-1] < 0.25 * el[0], (el[0], el[-1])
    assert all(b <= a * 1.05 for a, b in zip(el, el[1:])), "loss not decreasing"
    # Idef(x) = I(x + T): a subject shifted by s is matched by T = -s (the shifts have zero mean)
    assert (Ts + shifts).abs().max().item() < 0.5, Ts
    std = lm.StandardizedDataset(data, As, Ts)
    a = std[0]
    b = std[1]
    assert (a - b).abs().max().item() < 0.15 * data.max().item()   # both land on the atlas frame
