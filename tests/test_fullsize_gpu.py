"""GPU: size-independent properties at the BASELINE.json sizes (C2: 16 x 128^3, one 256^3 subject
of C3), where the CPU oracle would take too long: identities, exact discrete adjointness,
mass conservation, operator inverse, zero momentum, linearity."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g) * scale


def _smooth_disp(lm, shape, seed, amp):
    u = lm.FluidMetric([0.1, 0.0, 0.01]).sharp(_rand(shape, seed))
    return u * (amp / u.abs().max())


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


@pytest.mark.parametrize("N,n", [(16, 128), (1, 256)])
def test_interp_identity_adjointness_mass(lm, N, n):
    sh = (n, n, n)
    I = _rand((N, 1) + sh, 1)
    z = torch.zeros((N, 3) + sh, device="cuda")
    assert torch.equal(lm.interp(I, z), I)                       # interp(I, 0) == I, bit exact
    u = _smooth_disp(lm, (N, 3) + sh, 2, 6.0)
    u[:, :, :2] -= 5.0                                           # push a border band out of range
    g = _rand((N, 1) + sh, 3)
    Iu = lm.interp(I, u)
    adj = lm.interp_adjoint(g, u)
    lhs, rhs = _dot(Iu, g), _dot(I, adj)
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), I.numel() ** 0.5)   # <interp(I,u), g> == <I, splat(g,u)>
    assert abs(adj.double().sum().item() - g.double().sum().item()) <= 1e-5 * g.double().abs().sum().item()
    # broadcast image: all subjects splat into one image
    adjb = lm.interp_adjoint(g, u, broadcast=True)
    assert abs(adjb.double().sum().item() - g.double().sum().item()) <= 1e-5 * g.double().abs().sum().item()
    assert torch.allclose(adjb, adj.sum(0, keepdim=True), rtol=1e-3, atol=1e-3 * adj.abs().max().item())


@pytest.mark.parametrize("N,n", [(16, 128), (1, 256)])
def test_jacobian_family_adjointness(lm, N, n):
    sh = (N, 3, n, n, n)
    v, w, m = _rand(sh, 4), _rand(sh, 5), _rand(sh, 6)
    tol = lambda a, b: 2e-5 * max(abs(a), abs(b), v.numel() ** 0.5)
    a = _dot(lm.jacobian_times_vectorfield(v, w, displacement=True, transpose=False), m)
    b = _dot(w, lm.jacobian_times_vectorfield(v, m, displacement=True, transpose=True))
    assert abs(a - b) <= tol(a, b)
    a = _dot(lm.jacobian_times_vectorfield(v, w, displacement=False), m)
    b = _dot(v, lm.jacobian_times_vectorfield_adjoint(m, w))
    assert abs(a - b) <= tol(a, b)
    a = _dot(lm.ad(v, w), m)
    b = _dot(w, lm.ad_star(v, m))
    assert abs(a - b) <= tol(a, b)
    # fused Ad_star == its two-kernel definition
    phi = _smooth_disp(lm, sh, 7, 4.0)
    ref = lm.jacobian_times_vectorfield(phi, lm.interp(m, phi), displacement=True)
    out = lm.Ad_star(phi, m)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # fused compose == its definition
    ref = -0.1 * phi + 1.0 * lm.interp(v, phi, dt=-0.1)
    out = lm.compose(phi, v, ds=-0.1, dt=1.0)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("N,n", [(16, 128), (1, 256)])
@pytest.mark.parametrize("params", [[0.1, 0.0, 0.01], [0.1, 0.01, 0.001]])
def test_fluid_metric_inverse_symmetry_linearity(lm, N, n, params):
    sh = (N, 3, n, n, n)
    met = lm.FluidMetric(params)
    m, b = _rand(sh, 8), _rand(sh, 9)
    v = met.sharp(m)
    back = met.flat(v)
    assert ((back - m).norm() / m.norm()).item() <= 2e-4       # flat(sharp(m)) == m (reference test: atol 1e-3)
    x, y = _dot(v, b), _dot(m, met.sharp(b))
    assert abs(x - y) <= 1e-4 * max(abs(x), abs(y))             # symmetric operator
    lin = met.sharp(2.0 * m + b)
    assert ((lin - (2.0 * v + met.sharp(b))).norm() / lin.norm()).item() <= 1e-5
    c = torch.full(sh, 3.0, device="cuda")                      # constants: only gamma acts, squared
    assert torch.allclose(met.sharp(c), c / params[2] ** 2, rtol=1e-5)


@pytest.mark.parametrize("N,n,steps", [(16, 128, 10), (1, 256, 5)])
def test_expmap_zero_and_small_momentum(lm, N, n, steps):
    sh = (N, 3, n, n, n)
    met = lm.FluidMetric([0.1, 0.0, 0.01])
    z = torch.zeros(sh, device="cuda")
    assert torch.equal(lm.expmap(met, z, num_steps=steps), z)   # expmap(0) == 0
    # first-order behaviour: for tiny momentum phi^-1 ~ -sharp(m)
    m = _rand(sh, 10)
    v = met.sharp(m)
    m = m * (1e-3 / v.abs().max())
    h = lm.expmap(met, m, num_steps=steps)
    v = met.sharp(m)
    assert ((h + v).abs().max() / v.abs().max()).item() <= 1e-2


def test_extreme_displacements_and_ragged_sizes(lm, orc):
    """far out-of-range / huge / NaN coordinates stay memory safe; sizes that are not multiples of
    the warp width take the same fast path correctly"""
    sh = (9, 10, 40)
    I = torch.randn((2, 3) + sh, generator=torch.Generator().manual_seed(11))
    u = torch.randn((2, 3) + sh, generator=torch.Generator().manual_seed(12)) * 3
    u[0, :, :3] = 1.0e6
    u[1, :, -3:] = -3.0e5
    ref = orc.interp(I, u)
    out = lm.interp(I.cuda(), u.cuda())
    assert (out.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    ref = orc.Ad_star(u, I)
    assert (lm.Ad_star(u.cuda(), I.cuda()).cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    ref = orc.compose(u, I, -0.3, 1.0)
    assert (lm.compose(u.cuda(), I.cuda(), -0.3, 1.0).cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    un = u.clone()
    un[0, 0, 4, 4, 4] = float("nan")
    un[1, 2, 1, 1, 1] = float("inf")
    out = lm.interp(I.cuda(), un.cuda())          # must not fault
    adj = lm.interp_adjoint(I.cuda(), un.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(out[0, :, 0, 0, 0]).all()


@pytest.mark.parametrize("params", [[0.1, 0.0, 0.01], [0.1, 0.01, 0.001]])
def test_fluid_metric_256_matches_oracle(lm, orc, params):
    """256^3 takes the two-CTA cluster slab kernels (DSMEM transposition between the Z and the Y
    transform): sharp and flat against the CPU oracle (torch.fft), tolerance 1e-5 relative L2."""
    sh = (1, 3, 256, 256, 256)
    m = torch.randn(sh, generator=torch.Generator().manual_seed(21))
    om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
    for name in ("sharp", "flat"):
        ref = getattr(om, name)(m)
        out = getattr(gm, name)(m.cuda()).cpu()
        assert ((out - ref).norm() / ref.norm()).item() <= 1e-5, name
