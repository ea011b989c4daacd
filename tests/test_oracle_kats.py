"""CPU-only: the oracle against analytic known answers, the reference's property
tests (adjointness, transposition, 2D==3D, flat(sharp)=id, expmap(0)=0) and finite differences
of its own forward kernels (pins the restated backward kernels)."""
import numpy as np
import pytest
import torch

from util import randn

DT = torch.float64


def test_trilerp_known_answer(orc):
    I = torch.arange(8, dtype=torch.float32).reshape(1, 1, 2, 2, 2)
    u = torch.zeros(1, 3, 2, 2, 2)
    u[0, 0], u[0, 1], u[0, 2] = 0.5, 0.25, 0.75
    assert orc.interp(I, u)[0, 0, 0, 0, 0].item() == 3.25
    go = torch.zeros(1, 1, 2, 2, 2)
    go[0, 0, 0, 0, 0] = 1.0
    dI, du = orc.interp_backward(go, I, u)
    assert du[0, :, 0, 0, 0].tolist() == [4.0, 2.0, 1.0]
    assert abs(dI.sum().item() - 1.0) < 1e-6


def test_interp_identity_shift_and_clamp(orc):
    I = randn((2, 3, 5, 6, 7), DT, 1)
    z = torch.zeros(2, 3, 5, 6, 7, dtype=DT)
    assert torch.equal(orc.interp(I, z), I)
    u = z.clone()
    u[:, 2] = 1.0  # integer shift along z with constant extension at the border
    out = orc.interp(I, u)
    assert torch.equal(out[..., :-1], I[..., 1:]) and torch.equal(out[..., -1], I[..., -1])
    u[:, 2] = -2.3  # below range at k=0..2 -> both corners clamp to v[0]: (1-t)*v0 + t*v0, t = 0.7
    assert torch.allclose(orc.interp(I, u)[..., 0], I[..., 0], rtol=1e-15, atol=1e-15)
    u[:, 2] = 50.0
    assert torch.equal(orc.interp(I, u)[..., 3], I[..., -1])


def test_gradient_known_answer(orc):
    ii, jj, kk = torch.meshgrid(torch.arange(5.), torch.arange(5.), torch.arange(5.), indexing="ij")
    f = (3 * ii + 5 * jj + 7 * kk).reshape(1, 1, 5, 5, 5)
    for d, (inner, corner) in enumerate([(3.0, 1.5), (5.0, 2.5), (7.0, 3.5)]):
        w = torch.zeros(1, 3, 5, 5, 5)
        w[0, d] = 1
        o = orc.jtvf_forward(f, w, False, False)
        assert o[0, 0, 2, 2, 2].item() == inner
        assert o[0, 0, 0, 0, 0].item() == corner and o[0, 0, 4, 4, 4].item() == corner


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("disp", [True, False])
def test_transpose_and_adjoint_properties(orc, dim, disp):
    sh = (2, dim) + (4,) * dim
    g, u, v, m = (randn(sh, DT, s) for s in (1, 2, 3, 4))
    a = (orc.jtvf_forward(g, u, disp, False) * v).sum()
    b = (u * orc.jtvf_forward(g, v, disp, True)).sum()
    assert torch.allclose(a, b)
    a = (orc.jtvf_forward(u, v, False, False) * m).sum()
    b = (u * orc.jtvf_adjoint_forward(m, v)).sum()
    assert torch.allclose(a, b)
    # ad_star is the exact discrete adjoint of ad(v, .)
    assert torch.allclose((orc.ad(v, u) * m).sum(), (u * orc.ad_star(v, m)).sum())


def test_2d_matches_3d(orc):
    I = randn((2, 2, 3, 4), DT, 5)
    u = randn((2, 2, 3, 4), DT, 6)
    u3 = torch.zeros(2, 3, 3, 4, 1, dtype=DT)
    u3[:, :2] = u.unsqueeze(4)
    assert torch.allclose(orc.interp(I, u).unsqueeze(4), orc.interp(I.unsqueeze(4), u3))
    v2, m2 = randn((2, 2, 2, 2), DT, 7), randn((2, 2, 2, 2), DT, 8)
    rep = lambda x: torch.cat([x.unsqueeze(4)] * 2, 4)
    v3 = torch.zeros(2, 3, 2, 2, 2, dtype=DT)
    m3 = torch.zeros(2, 3, 2, 2, 2, dtype=DT)
    v3[:, :2], m3[:, :2] = rep(v2), rep(m2)
    for disp in (True, False):
        for trans in (True, False):
            assert torch.allclose(orc.jtvf_forward(v3, m3, disp, trans)[:, :2, :, :, 0],
                                  orc.jtvf_forward(v2, m2, disp, trans))
    assert torch.allclose(orc.jtvf_adjoint_forward(v3, m3)[:, :2, :, :, 0], orc.jtvf_adjoint_forward(v2, m2))


def _fd(f, x, go, eps=1e-6):
    """finite-difference gradient of <f(x), go> w.r.t. x"""
    g = torch.zeros_like(x)
    xf = x.reshape(-1)
    for i in range(xf.numel()):
        old = xf[i].item()
        xf[i] = old + eps
        fp = (f(x) * go).sum()
        xf[i] = old - eps
        fm = (f(x) * go).sum()
        xf[i] = old
        g.reshape(-1)[i] = (fp - fm) / (2 * eps)
    return g


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("bcast", [False, True])
def test_interp_backward_matches_finite_differences(orc, dim, bcast):
    sh = (3,) * dim
    I = randn((1 if bcast else 2, 2) + sh, DT, 11)
    u = randn((2, dim) + sh, DT, 12, 0.4) + 0.25  # keep away from cell boundaries' kinks mostly
    go = randn((2, 2) + sh, DT, 13)
    dI, du = orc.interp_backward(go, I, u, 0.8)
    assert torch.allclose(dI, _fd(lambda x: orc.interp(x, u, 0.8), I.clone(), go), atol=1e-7)
    assert torch.allclose(du, _fd(lambda x: orc.interp(I, x, 0.8), u.clone(), go), atol=1e-5)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("disp", [True, False])
@pytest.mark.parametrize("trans", [True, False])
def test_jtvf_backward_matches_finite_differences(orc, dim, disp, trans):
    sh = (2, dim) + (3,) * dim
    v, w, go = randn(sh, DT, 21), randn(sh, DT, 22), randn(sh, DT, 23)
    dv, dw = orc.jtvf_backward(go, v, w, disp, trans)
    assert torch.allclose(dv, _fd(lambda x: orc.jtvf_forward(x, w, disp, trans), v.clone(), go), atol=1e-7)
    assert torch.allclose(dw, _fd(lambda x: orc.jtvf_forward(v, x, disp, trans), w.clone(), go), atol=1e-7)
    dz, dw2 = orc.jtvf_adjoint_backward(go, v, w)
    assert torch.allclose(dz, _fd(lambda x: orc.jtvf_adjoint_forward(x, w), v.clone(), go), atol=1e-7)
    assert torch.allclose(dw2, _fd(lambda x: orc.jtvf_adjoint_forward(v, x), w.clone(), go), atol=1e-7)


@pytest.mark.parametrize("dim", [2, 3])
def test_fluid_metric_properties(orc, dim):
    sh = (2, dim) + (6,) * dim
    met = orc.FluidMetric([0.1, 0.01, 0.001])  # test_metric.py parameters (beta != 0)
    m = randn(sh, DT, 31)
    assert torch.allclose(met.flat(met.sharp(m)), m, atol=1e-3)
    # symmetric: <sharp(a), b> == <a, sharp(b)>
    b = randn(sh, DT, 32)
    assert torch.allclose((met.sharp(m) * b).sum(), (m * met.sharp(b)).sum())
    # constants: only gamma acts; symbol is squared
    c = torch.full(sh, 2.0, dtype=DT)
    g = 0.01
    assert torch.allclose(orc.FluidMetric([0.1, 0, g]).sharp(c), c / g ** 2)
    # single Fourier mode along x, beta = 0
    N, k, a = 6, 1, 0.1
    x = torch.arange(N, dtype=DT)
    mode = torch.zeros(sh, dtype=DT)
    mode[:, 0] = torch.cos(2 * np.pi * k * x / N).reshape([N] + [1] * (dim - 1))
    lam = g + a * 2 * (1 - np.cos(2 * np.pi * k / N))
    assert torch.allclose(orc.FluidMetric([a, 0, g]).sharp(mode), mode / lam ** 2, atol=1e-6)


@pytest.mark.parametrize("dim,res", [(2, 128), (3, 16)])
@pytest.mark.parametrize("steps", [1, 5])
def test_expmap_zero(orc, dim, res, steps):
    m = torch.zeros((1, dim) + (res,) * dim, dtype=DT)
    h = orc.expmap(orc.FluidMetric([1.0, 0.1, 0.01]), m, num_steps=steps)
    assert torch.equal(h, m)


def test_regrid_identity_and_adjoint(orc):
    I = randn((2, 3, 4, 5, 6), DT, 41)
    assert torch.allclose(orc.regrid(I, (4, 5, 6)), I)
    osh = (7, 6, 9)
    origin = tuple((s - 1) * 0.5 for s in (4, 5, 6))
    spacing = tuple((a - 1) / (b - 1) for a, b in zip((4, 5, 6), osh))
    go = randn((2, 3) + osh, DT, 42)
    lhs = (orc.regrid_forward(I, osh, origin, spacing) * go).sum()
    rhs = (I * orc.regrid_backward(go, (4, 5, 6), osh, origin, spacing)).sum()
    assert torch.allclose(lhs, rhs)


def test_c1_config_runs(orc):
    """BASELINE config 1: 2-D 128x128 batch 8, expmap 10 steps, FluidMetric(0.1, 0, 0.01), CPU."""
    from util import smooth_field
    met = orc.FluidMetric([0.1, 0.0, 0.01])
    m0 = smooth_field((8, 2, 128, 128), torch.float32, 1, amp=1.0, sigma=4.0)
    m0 = m0 * (4.0 / met.sharp(m0).abs().max())
    h = orc.expmap(met, m0, num_steps=10)
    assert torch.isfinite(h).all() and 1.0 < h.abs().max().item() < 16.0
