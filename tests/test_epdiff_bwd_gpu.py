"""Fused backward of the EPDiff step (lgm_epdiff_step_bwd) against
  (a) the CPU oracle's backward chain in fp64: interp / jtvf backward kernels restated from
      cuda/interp.cu:185-244 and cuda/diff.cu:286-473, chained as autograd chains them for
      lddmm.py:39-44 (sharp is self-adjoint: metric.py:21-34), and
  (b) the unfused CUDA autograd path (per-operator backward kernels, themselves pinned to the oracle).
Tolerance: 1e-4 of max|grad| (fp32 splats accumulate in atomic order)."""
import pytest
import torch

from util import relerr, smooth_field

pytestmark = pytest.mark.gpu

PARAMS = [0.1, 0.0, 0.01]


def relerr_q(a, b, allow=1e-4):
    """max|a-b| / max|b| after dropping the `allow` fraction of largest deviations. The derivative of
    trilinear interpolation jumps across cell faces, and the fp32 trajectory (GPU) and the fp64 one
    (oracle) can put an isolated sample on different sides of a face; those samples are not a
    rounding-level comparison."""
    a = a.detach().cpu().double().flatten()
    b = b.detach().cpu().double().flatten()
    e = (a - b).abs().sort().values
    drop = int(allow * e.numel() + 0.999)
    return (e[e.numel() - 1 - drop] / b.abs().max()).item()


def oracle_shoot_grad(orc, om, m0, phi0, steps, T, gout, mask=None):
    """d/dm0 and d/dphi0 of <expmap(m0, phi0), gout> by the hand-chained oracle backward."""
    dt = T / steps
    phis, vs, mis = [], [], []
    phi = phi0
    for k in range(steps):
        mi = orc.interp_forward(m0, phi, 1.0)
        m = orc.jtvf_forward(phi, mi, True, False)
        if mask is not None:
            m = m * mask
        v = om.sharp(m)
        phis.append(phi); vs.append(v); mis.append(mi)
        phi = -dt * v + orc.interp_forward(phi, v, -dt)
    G = gout.clone()
    d_m0 = torch.zeros_like(m0)
    for k in reversed(range(steps)):
        phi, v, mi = phis[k], vs[k], mis[k]
        d_phi_splat, d_v = orc.interp_backward(G, phi, v, -dt)
        d_v = d_v + (-dt) * G
        d_m = om.sharp(d_v)
        if mask is not None:
            d_m = d_m * mask
        d_phi_j, d_mi = orc.jtvf_backward(d_m, phi, mi, True, False)
        d_m0_k, d_phi_i = orc.interp_backward(d_mi, m0, phi, 1.0)
        d_m0 += d_m0_k
        G = d_phi_splat + d_phi_j + d_phi_i
    return phi if steps == 0 else None, d_m0, G


def make_inputs(N, sh, seed, vmax=2.5):
    import oracle.oracle as orc
    om = orc.FluidMetric(PARAMS)
    m0 = smooth_field((N, 3) + sh, torch.float64, seed, amp=1.0, sigma=2.0)
    v0 = om.sharp(m0)
    m0 = (m0 * (vmax / v0.abs().max())).float()
    phi0 = smooth_field((N, 3) + sh, torch.float64, seed + 1, amp=1.5, sigma=2.0).float()
    phi0[:, :, 0] -= 2.0  # push a border band out of range (clamp path)
    phi0[..., -1] += 2.5
    gout = smooth_field((N, 3) + sh, torch.float64, seed + 2, amp=1.0, sigma=1.0).float()
    return om, m0, phi0, gout


@pytest.mark.parametrize("sh", [(8, 6, 32), (16, 16, 32), (5, 9, 64)])
@pytest.mark.parametrize("steps", [1, 3])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_backward_vs_oracle(lm, orc, sh, steps, masked):
    N = 2
    om, m0, phi0, gout = make_inputs(N, sh, 300 + steps)
    mask = None
    if masked:
        mask = (torch.rand((N, 3) + sh, generator=torch.Generator().manual_seed(5)) > 0.3).float()
    _, ref_m0, ref_phi = oracle_shoot_grad(orc, om, m0.double(), phi0.double(), steps, 1.0, gout.double(),
                                           None if mask is None else mask.double())
    gm = lm.FluidMetric(PARAMS)
    a = m0.cuda().requires_grad_(True)
    p = phi0.cuda().requires_grad_(True)
    n0 = lm.launch_count()
    out = lm.expmap(gm, a, T=1.0, num_steps=steps, phiinv=p, mommask=None if mask is None else mask.cuda())
    assert type(out.grad_fn).__name__ == "EPDiffShootFunctionBackward"
    d_m0, d_phi = torch.autograd.grad(out, [a, p], gout.cuda())
    assert lm.launch_count() > n0
    assert relerr_q(d_m0, ref_m0) <= 1e-4
    assert relerr_q(d_phi, ref_phi) <= 1e-4
    # and nothing is wildly off anywhere (a face crossing changes one sample's gradient, not its scale)
    assert relerr(d_m0, ref_m0) <= 0.05 and relerr(d_phi, ref_phi) <= 0.05


@pytest.mark.parametrize("sh", [(16, 16, 32), (32, 32, 32)])
@pytest.mark.parametrize("mode", ["expmap", "checkpoint", "step", "m0_only", "phi_only"])
def test_fused_backward_vs_unfused(lm, monkeypatch, sh, mode):
    N, steps = 2, 4
    _, m0, phi0, gout = make_inputs(N, sh, 410)
    gm = lm.FluidMetric(PARAMS)

    def run(fused):
        monkeypatch.setenv("LGM_FUSED_BWD", "1" if fused else "0")
        a = m0.cuda().requires_grad_(mode != "phi_only")
        p = phi0.cuda().requires_grad_(mode != "m0_only")
        if mode == "checkpoint":
            out = lm.expmap(gm, a, num_steps=steps, phiinv=p, checkpoints=2)
        elif mode == "step":
            out = lm.EPDiff_step(gm, a, 0.25, p)
        else:
            out = lm.expmap(gm, a, num_steps=steps, phiinv=p)
        ins = [t for t in (a, p) if t.requires_grad]
        return out.detach(), torch.autograd.grad(out, ins, gout.cuda())

    o1, g1 = run(True)
    o0, g0 = run(False)
    assert relerr(o1, o0) <= 1e-5
    assert len(g1) == len(g0)
    for x, y in zip(g1, g0):
        assert relerr(x, y) <= 1e-4


def test_fused_backward_default_phiinv_and_loss(lm, monkeypatch):
    """expmap(m0) with phiinv=None (the atlas builder's call): gradient of a sum-of-squares loss."""
    _, m0, _, _ = make_inputs(2, (16, 16, 32), 77)
    gm = lm.FluidMetric(PARAMS)
    gr = []
    for fused in ("1", "0"):
        monkeypatch.setenv("LGM_FUSED_BWD", fused)
        a = m0.cuda().requires_grad_(True)
        h = lm.expmap(gm, a, num_steps=5)
        (h * h).sum().backward()
        gr.append(a.grad.clone())
    assert relerr(gr[0], gr[1]) <= 1e-4


@pytest.mark.parametrize("steps", [1, 2, 4])
@pytest.mark.parametrize("masked", [False, True])
def test_identity_start_shortcut_gradient(lm, monkeypatch, steps, masked):
    """phiinv=None: the first step runs as sharp + scaling and its backward as scaling + sharp
    (lddmm._steps_saving / _steps_backward). Output and d_m0 against the per-operator autograd chain."""
    _, m0, _, gout = make_inputs(2, (16, 16, 32), 91)
    gm = lm.FluidMetric(PARAMS)
    mask = None
    if masked:
        mask = (torch.rand(m0.shape, generator=torch.Generator().manual_seed(6)) > 0.3).float().cuda()
    res = []
    for fused in ("1", "0"):
        monkeypatch.setenv("LGM_FUSED_BWD", fused)
        a = m0.cuda().requires_grad_(True)
        h = lm.expmap(gm, a, num_steps=steps, mommask=mask)
        assert (type(h.grad_fn).__name__ == "EPDiffShootFunctionBackward") == (fused == "1")
        (g,) = torch.autograd.grad(h, [a], gout.cuda())
        res.append((h.detach(), g))
    assert relerr(res[0][0], res[1][0]) <= 1e-5
    assert relerr(res[0][1], res[1][1]) <= 1e-4


def test_unsupported_shapes_fall_back(lm):
    """Z not a multiple of 32, 2-D and fp64 keep the per-operator autograd chain (still CUDA)."""
    gm = lm.FluidMetric(PARAMS)
    for shape, dt in (((1, 3, 8, 8, 12), torch.float32), ((1, 2, 16, 32), torch.float32),
                      ((1, 3, 8, 8, 32), torch.float64)):
        a = torch.randn(shape, dtype=dt, generator=torch.Generator().manual_seed(3)).cuda().requires_grad_(True)
        out = lm.expmap(gm, a, num_steps=2)
        assert type(out.grad_fn).__name__ != "EPDiffShootFunctionBackward"
        out.sum().backward()
        assert torch.isfinite(a.grad).all()
