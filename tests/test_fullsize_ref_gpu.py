"""GPU: the product against the REFERENCE'S OWN CUDA KERNELS (oracle/_ref/libref_cuda.so + the
reference's Python composition, tests/util.py RefPipeline) at the BASELINE.json sizes -- not just
size-independent properties:
  C2  2 x 3 x 128^3, 10-step shoot, T = 1, max|v| = 4 voxels (BASELINE.md section 4 momenta)
  C3  1 x 3 x 256^3, 5-step shoot (lddmm.py:119,305)
  C1  8 x 2 x 128^2, 10 steps: GPU vs the CPU oracle (tests/test_shoot_gpu.py)
  C4  2 x 1 x 192^3 affine_interp forward + all three gradients
plus the single operators of one step at 128^3 / 256^3. Tolerances are BASELINE.md section 2:
interp / Jacobian / metric 1e-5 relative to max|ref|, the shoot 1e-4, splats compared as a sum with
a 1e-4 elementwise bound."""
import pytest
import torch

from util import RefCuda, RefPipeline, baseline_momenta, relerr, l2err

pytestmark = pytest.mark.gpu

PARAMS = [0.1, 0.0, 0.01]


@pytest.fixture(scope="module")
def ref():
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    return RefPipeline(RefCuda(), PARAMS)


@pytest.mark.parametrize("N,n,steps", [(2, 128, 10), (1, 256, 5)])
def test_shoot_fullsize_vs_reference_kernels(lm, ref, N, n, steps):
    sh = (n, n, n)
    m0 = baseline_momenta(N, sh, sigma=4.0, vmax=4.0, seed=1)
    want = ref.expmap(m0, steps)
    got = lm.expmap(lm.FluidMetric(PARAMS), m0, num_steps=steps)
    assert want.abs().max().item() > 1.0          # a real deformation, several voxels
    assert relerr(got, want) <= 1e-4
    assert l2err(got, want) <= 1e-5


@pytest.mark.parametrize("N,n", [(2, 128), (1, 256)])
def test_step_operators_fullsize_vs_reference_kernels(lm, ref, N, n):
    sh = (n, n, n)
    metric = lm.FluidMetric(PARAMS)
    m0 = baseline_momenta(N, sh, seed=2)
    # a displacement of a few voxels with a border band pushed out of range (clamped corners)
    phi = metric.sharp(m0)
    phi = phi * (5.0 / phi.abs().max())
    phi[:, :, :2] -= 4.0
    phi[:, :, :, :, -2:] += 4.0
    want = ref.Ad_star(phi, m0)
    assert relerr(lm.Ad_star(phi, m0), want) <= 1e-5
    v = metric.sharp(want)
    vr = ref.fluid(want, True)
    assert relerr(v, vr) <= 1e-5 and l2err(v, vr) <= 1e-5
    # flat on a broad-band field (on the smooth vr the output is a cancellation ~1e-4 of the input and
    # fp32 rounding of EITHER implementation is 1e-4 of it)
    w = torch.randn(vr.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    assert relerr(metric.flat(w), ref.fluid(w, False)) <= 1e-5
    fs, fr = metric.flat(vr), ref.fluid(vr, False)
    assert (fs - fr).abs().max().item() <= 1e-5 * vr.abs().max().item() * (0.01 + 0.1 * 12) ** 2
    assert relerr(lm.compose(vr, phi, ds=-0.1, dt=1.0), ref.compose(vr, phi, -0.1, 1.0)) <= 1e-5
    assert relerr(lm.ad_star(vr, m0), ref.ad_star(vr, m0)) <= 1e-5
    I = baseline_momenta(N, sh, seed=3)[:, :1].contiguous()
    assert relerr(lm.interp(I, phi), ref.rc.interp_fwd(I, phi, 1.0)) <= 1e-5


@pytest.mark.parametrize("N,n", [(2, 128)])
def test_interp_backward_fullsize_vs_reference_kernels(lm, ref, N, n):
    sh = (n, n, n)
    g = torch.Generator(device="cuda").manual_seed(4)
    I = torch.randn((N, 1) + sh, device="cuda", generator=g)
    go = torch.randn((N, 1) + sh, device="cuda", generator=g)
    u = lm.FluidMetric(PARAMS).sharp(torch.randn((N, 3) + sh, device="cuda", generator=g))
    u = u * (5.0 / u.abs().max())
    dI_r, du_r = ref.rc.interp_bwd(go, I, u, 1.0)
    Ic, uc = I.clone().requires_grad_(True), u.clone().requires_grad_(True)
    lm.interp(Ic, uc).backward(go)
    assert relerr(uc.grad, du_r) <= 1e-5
    # splat: atomic order differs -> compare the sum, then elementwise with the looser bound
    s, sr = Ic.grad.double().sum().item(), dI_r.double().sum().item()
    assert abs(s - sr) <= 1e-5 * go.double().abs().sum().item()
    assert relerr(Ic.grad, dI_r) <= 1e-4


@pytest.mark.parametrize("N,n", [(1, 128), (1, 256)])
def test_stress_variant_vs_reference_kernels(lm, ref, N, n):
    """SURVEY 8(d) stress variant: white-noise displacement uniform in +-8 voxels plus border bands pushed
    +-(n + 2) voxels out of range -- no locality for the gathers (every warp of the staged-ring kernels leaves
    its window or sits at a clamped border), all splats of a band land on one face (atomic contention)."""
    sh = (n, n, n)
    g = torch.Generator(device="cuda").manual_seed(1)
    u = (torch.rand((N, 3) + sh, device="cuda", generator=g) - 0.5) * 16.0
    u[:, 0, :3] -= n + 2
    u[:, 1, :, -3:] += n + 2
    u[:, 2, :, :, :2] -= n + 2
    u[:, 2, :, :, -2:] += n + 2
    m = torch.randn((N, 3) + sh, device="cuda", generator=g)
    assert relerr(lm.Ad_star(u, m), ref.Ad_star(u, m)) <= 1e-5
    assert relerr(lm.compose(m, u, ds=1.0, dt=1.0), ref.compose(m, u, 1.0, 1.0)) <= 1e-5
    assert relerr(lm.compose(u, m, ds=-0.1, dt=1.0), ref.compose(u, m, -0.1, 1.0)) <= 1e-5   # ring kernel, all fallbacks
    I = m[:, :1].contiguous()
    go = torch.randn((N, 1) + sh, device="cuda", generator=g)
    assert relerr(lm.interp(I, u), ref.rc.interp_fwd(I, u, 1.0)) <= 1e-5
    dI_r, du_r = ref.rc.interp_bwd(go, I, u, 1.0)
    Ic, uc = I.clone().requires_grad_(True), u.clone().requires_grad_(True)
    lm.interp(Ic, uc).backward(go)
    assert relerr(uc.grad, du_r) <= 1e-5
    s, sr = Ic.grad.double().sum().item(), dI_r.double().sum().item()
    assert abs(s - sr) <= 1e-5 * go.double().abs().sum().item()
    assert relerr(Ic.grad, dI_r) <= 1e-4


def test_affine_interp_c4_vs_reference_kernels(lm, ref):
    """BASELINE config 4 at its grid (192^3), 2 subjects: forward and d_I, d_A, d_T."""
    N, sh = 2, (192, 192, 192)
    g = torch.Generator(device="cuda").manual_seed(5)
    I = baseline_momenta(N, sh, seed=6)[:, :1].contiguous()
    A = (torch.eye(3, device="cuda")[None] + 0.05 * torch.randn((N, 3, 3), device="cuda", generator=g)).contiguous()
    T = 2 * torch.randn((N, 3), device="cuda", generator=g)
    go = torch.randn((N, 1) + sh, device="cuda", generator=g)
    want = ref.rc.affine_fwd(I, A, T)
    Ic, Ac, Tc = I.clone().requires_grad_(True), A.clone().requires_grad_(True), T.clone().requires_grad_(True)
    out = lm.affine_interp(Ic, Ac, Tc)
    assert relerr(out, want) <= 1e-5
    out.backward(go)
    dI_r, dA_r, dT_r = ref.rc.affine_bwd(go, I, A, T)
    # d_A / d_T are sums over 7 M voxels of +-O(1) terms: fp32 accumulation order differs (tree vs atomics)
    assert relerr(Ac.grad, dA_r) <= 1e-3 and relerr(Tc.grad, dT_r) <= 1e-3
    assert abs(Ic.grad.double().sum().item() - dI_r.double().sum().item()) <= 1e-5 * go.double().abs().sum().item()
    assert relerr(Ic.grad, dI_r) <= 1e-4
