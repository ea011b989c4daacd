"""Parity of the CUDA product (through the Python mirror -> C ABI) against the CPU oracle,
on the same seeded inputs. fp64 to 1e-12, fp32 to 1e-5 relative (splat 1e-4: atomic order)."""
import pytest
import torch

import util
from util import randn, relerr, l2err, smooth_field, tol_for

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.float64]
SHAPES = {2: [(7, 9), (32, 16)], 3: [(5, 6, 7), (16, 8, 32)]}


def cases():
    for dim in (2, 3):
        for sh in SHAPES[dim]:
            for dt in DTYPES:
                yield dim, sh, dt


def disp_like(N, dim, sh, dtype, seed, amp=2.5):
    """displacement with out-of-range excursions at the border (clamp path)"""
    u = randn((N, dim) + sh, dtype, seed, scale=amp)
    u[:, :, 0] -= 3.0
    u[..., -1] += 3.5
    return u


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
@pytest.mark.parametrize("bcast", [False, True])
@pytest.mark.parametrize("C", [1, 3])
def test_interp_forward(lm, orc, dim, sh, dtype, bcast, C):
    N = 3
    I = randn((1 if bcast else N, C) + sh, dtype, 11)
    u = disp_like(N, dim, sh, dtype, 12)
    for dt in (1.0, -0.1):
        ref = orc.interp(I, u, dt)
        out = lm.interp(I.cuda(), u.cuda(), dt)
        assert relerr(out, ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
@pytest.mark.parametrize("bcast", [False, True])
def test_interp_backward(lm, orc, dim, sh, dtype, bcast):
    N, C = 2, 2
    I = randn((1 if bcast else N, C) + sh, dtype, 21)
    u = disp_like(N, dim, sh, dtype, 22)
    go = randn((N, C) + sh, dtype, 23)
    dI_ref, du_ref = orc.interp_backward(go, I, u, 0.7)
    Ic, uc = I.cuda().requires_grad_(True), u.cuda().requires_grad_(True)
    out = lm.interp(Ic, uc, 0.7)
    dI, du = torch.autograd.grad(out, [Ic, uc], go.cuda())
    assert relerr(du, du_ref) <= tol_for(dtype)
    assert relerr(dI, dI_ref) <= tol_for(dtype, splat=True)
    # splat conserves mass: sum d_I == sum gout (clamp keeps everything inside)
    assert abs(dI.double().sum().item() - go.double().sum().item()) <= 1e-5 * go.double().abs().sum().item()
    # interp_adjoint == d_I
    adj = lm.interp_adjoint(go.cuda(), u.cuda(), 0.7, broadcast=bcast)
    assert relerr(adj, dI_ref) <= tol_for(dtype, splat=True)


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
@pytest.mark.parametrize("disp", [False, True])
@pytest.mark.parametrize("trans", [False, True])
def test_jtvf_forward_backward(lm, orc, dim, sh, dtype, disp, trans):
    N = 2
    v = randn((N, dim) + sh, dtype, 31)
    w = randn((N, dim) + sh, dtype, 32)
    go = randn((N, dim) + sh, dtype, 33)
    ref = orc.jtvf_forward(v, w, disp, trans)
    dv_ref, dw_ref = orc.jtvf_backward(go, v, w, disp, trans)
    vc, wc = v.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    out = lm.jacobian_times_vectorfield(vc, wc, displacement=disp, transpose=trans)
    assert relerr(out, ref) <= tol_for(dtype)
    dv, dw = torch.autograd.grad(out, [vc, wc], go.cuda())
    assert relerr(dv, dv_ref) <= tol_for(dtype)
    assert relerr(dw, dw_ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
def test_jtvf_scalar_channels(lm, orc, dim, sh, dtype):
    """non-transposed, non-displacement mode accepts any channel count for v"""
    N, C = 2, 4
    v = randn((N, C) + sh, dtype, 34)
    w = randn((N, dim) + sh, dtype, 35)
    ref = orc.jtvf_forward(v, w, False, False)
    out = lm.jacobian_times_vectorfield(v.cuda(), w.cuda(), displacement=False, transpose=False)
    assert relerr(out, ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
def test_jtvf_adjoint_forward_backward(lm, orc, dim, sh, dtype):
    N = 2
    z = randn((N, dim) + sh, dtype, 41)
    w = randn((N, dim) + sh, dtype, 42)
    go = randn((N, dim) + sh, dtype, 43)
    ref = orc.jtvf_adjoint_forward(z, w)
    dz_ref, dw_ref = orc.jtvf_adjoint_backward(go, z, w)
    zc, wc = z.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    out = lm.jacobian_times_vectorfield_adjoint(zc, wc)
    assert relerr(out, ref) <= tol_for(dtype)
    dz, dw = torch.autograd.grad(out, [zc, wc], go.cuda())
    assert relerr(dz, dz_ref) <= tol_for(dtype)
    assert relerr(dw, dw_ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh,dtype", list(cases()))
def test_fused_adjrep(lm, orc, dim, sh, dtype):
    N = 2
    v = randn((N, dim) + sh, dtype, 51)
    m = randn((N, dim) + sh, dtype, 52)
    phi = disp_like(N, dim, sh, dtype, 53, amp=1.5)
    assert relerr(lm.ad(v.cuda(), m.cuda()), orc.ad(v, m)) <= tol_for(dtype)
    assert relerr(lm.ad_star(v.cuda(), m.cuda()), orc.ad_star(v, m)) <= tol_for(dtype)
    assert relerr(lm.Ad_star(phi.cuda(), m.cuda()), orc.Ad_star(phi, m)) <= tol_for(dtype)
    for ds, dt in ((1.0, 1.0), (-0.1, 1.0), (0.3, -2.0)):
        assert relerr(lm.compose(phi.cuda(), v.cuda(), ds, dt), orc.compose(phi, v, ds, dt)) <= tol_for(dtype)


# incl. shapes whose Z transform takes the global-memory edge stage (two radix stages: Z >= 64) in the
# line kernels (Y != Z) and in the slab kernels (Y == Z)
FLUID_SHAPES = {2: [(3, 3), (6, 10), (16, 32), (64, 64), (32, 128), (16, 256)],
                3: [(3, 3, 3), (4, 6, 5), (8, 16, 32), (32, 8, 16), (16, 32, 64), (8, 64, 64), (8, 16, 128)]}
FLUID_PARAMS = [[0.1, 0.0, 0.01], [0.1, 0.01, 0.001], [1.0, 0.1, 0.01]]


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("params", FLUID_PARAMS)
def test_fluid_metric(lm, orc, dim, dtype, params):
    for sh in FLUID_SHAPES[dim]:
        m = randn((2, dim) + sh, dtype, 61)
        om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
        tol = 1e-5 if dtype == torch.float32 else 1e-11
        assert l2err(gm.sharp(m.cuda()), om.sharp(m)) <= tol, (sh, "sharp")
        assert l2err(gm.flat(m.cuda()), om.flat(m)) <= tol, (sh, "flat")


# sizes that are not powers of two: the mixed-radix path (csrc/mixfft.cuh): radices 4/2/3/5/7, generic
# prime stages (13, 109, 11*11), odd last axis, mixed with power-of-two axes, typical MRI grids
MIXED_SHAPES = {2: [(5, 7), (9, 15), (24, 40), (48, 96), (218, 26), (121, 22), (30, 1000)],
                3: [(5, 6, 7), (12, 10, 14), (24, 40, 48), (20, 24, 20), (16, 48, 80), (26, 22, 21), (40, 48, 40)]}


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("params", FLUID_PARAMS[:2])
def test_fluid_metric_mixed_radix(lm, orc, dim, dtype, params):
    for sh in MIXED_SHAPES[dim]:
        m = randn((2, dim) + sh, dtype, 63)
        om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
        tol = 1e-5 if dtype == torch.float32 else 1e-11
        assert l2err(gm.sharp(m.cuda()), om.sharp(m)) <= tol, (sh, "sharp")
        assert l2err(gm.flat(m.cuda()), om.flat(m)) <= tol, (sh, "flat")


def test_fluid_metric_mri_grid(lm, orc):
    """one subject on a 160 x 192 x 160 grid (5*32, 3*64) and on 91 x 109 x 91 (MNI 2 mm: 7*13, prime 109)"""
    for sh in [(160, 192, 160), (91, 109, 91)]:
        m = randn((1, 3) + sh, torch.float32, 64)
        om, gm = orc.FluidMetric([0.1, 0.0, 0.01]), lm.FluidMetric([0.1, 0.0, 0.01])
        assert l2err(gm.sharp(m.cuda()), om.sharp(m)) <= 1e-5, sh


@pytest.mark.parametrize("dtype", DTYPES)
def test_fluid_metric_partial_line_block(lm, orc, dtype):
    """fewer lines than one CTA of the Z pass holds (16 < 32): the padded fill path, not the edge stage"""
    m = randn((1, 2, 8, 64), dtype, 62)
    om, gm = orc.FluidMetric([0.1, 0.01, 0.001]), lm.FluidMetric([0.1, 0.01, 0.001])
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    assert l2err(gm.sharp(m.cuda()), om.sharp(m)) <= tol
    assert l2err(gm.flat(m.cuda()), om.flat(m)) <= tol


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("params", FLUID_PARAMS)
@pytest.mark.parametrize("inverse", [True, False])
def test_fluid_operator_boundary(lm, orc, dim, dtype, params, inverse):
    """the standalone multiplier (reference FFI fluid_operator) on a torch.fft spectrum"""
    sh = (6, 5) if dim == 2 else (4, 6, 5)
    m = randn((2, dim) + sh, dtype, 62)
    dims = tuple(range(2, 2 + dim))
    F = torch.view_as_real(torch.fft.rfftn(m, dim=dims, norm="ortho")).contiguous()
    cos, sin = orc.FluidMetric.luts(m.shape, dtype)
    Fg = F.cuda()
    lm.fluid_operator(Fg, inverse, [c.cuda() for c in cos], [s.cuda() for s in sin], *params)
    orc.fluid_operator(F, inverse, cos, sin, *params)
    assert relerr(Fg, F) <= (1e-5 if dtype == torch.float32 else 1e-12)


@pytest.mark.parametrize("dim,sh", [(2, (32, 32)), (3, (16, 16, 16))])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("steps", [1, 5])
def test_expmap(lm, orc, dim, sh, dtype, steps):
    params = [0.1, 0.0, 0.01]
    N = 2
    om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
    m0 = smooth_field((N, dim) + sh, torch.float64, 71, amp=1.0, sigma=2.0)
    v0 = om.sharp(m0)
    m0 = (m0 * (3.0 / v0.abs().max())).to(dtype)  # max |v| = 3 voxels
    ref = orc.expmap(om, m0, num_steps=steps)
    out = lm.expmap(gm, m0.cuda(), num_steps=steps)
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    assert relerr(out, ref) <= tol
    # the autograd (unfused) path computes the same thing
    out2 = lm.expmap(gm, m0.cuda().requires_grad_(True), num_steps=steps)
    assert relerr(out2, ref) <= tol
    # expmap_advect
    ref3 = orc.expmap_advect(om, m0, num_steps=steps)
    out3 = lm.expmap_advect(gm, m0.cuda(), num_steps=steps)
    assert relerr(out3, ref3) <= tol


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("disp", [False, True])
def test_regrid(lm, orc, dim, dtype, disp):
    sh = (9, 7) if dim == 2 else (6, 9, 7)
    osh = (13, 12) if dim == 2 else (11, 13, 12)
    I = randn((2, dim) + sh, dtype, 81)
    ref = orc.regrid(I, osh, displacement=disp)
    Ic = I.cuda().requires_grad_(True)
    out = lm.regrid(Ic, shape=osh, displacement=disp)
    assert relerr(out, ref) <= tol_for(dtype)
    go = randn(tuple(out.shape), dtype, 82)
    (dI,) = torch.autograd.grad(out, [Ic], go.cuda())
    origin = tuple((s - 1) * 0.5 for s in sh)
    spacing = tuple((a - 1) / (b - 1) for a, b in zip(sh, osh))
    gs = go
    if disp:
        gs = go * (1.0 / torch.tensor(spacing, dtype=dtype).view(1, dim, *[1] * dim))
    dI_ref = orc.regrid_backward(gs, sh, osh, origin, spacing)
    assert relerr(dI, dI_ref) <= tol_for(dtype, splat=True)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bcast", [False, True])
def test_affine_interp_forward(lm, orc, dim, dtype, bcast):
    sh = (9, 7) if dim == 2 else (6, 9, 7)
    N, C = 3, 2
    I = randn((1 if bcast else N, C) + sh, dtype, 91)
    A = torch.eye(dim, dtype=dtype).repeat(N, 1, 1) + randn((N, dim, dim), dtype, 92, 0.1)
    T = randn((N, dim), dtype, 93, 1.5)
    ref = orc.affine_interp_forward(I, A, T)
    out = lm.affine_interp(I.cuda(), A.cuda(), T.cuda())
    # coordinates are a 3-term fp32 dot product whose contraction order differs between
    # compilers: allow one coordinate ulp times the image gradient
    assert relerr(out, ref) <= (5e-5 if dtype == torch.float32 else 1e-12)


@pytest.mark.parametrize("bcast", [False, True])
@pytest.mark.parametrize("C", [1, 2])
def test_affine_interp_fast_paths_3d(lm, orc, bcast, C):
    """fp32 3-D fast kernels (csrc/affine3.cu; the backward one needs Z % 32 == 0): forward against the
    oracle, all three gradients against the fp64 generic kernels on the same inputs (incl. samples
    pushed outside the volume by the rotation / translation)"""
    sh = (6, 10, 32)
    N = 3
    I = randn((1 if bcast else N, C) + sh, torch.float32, 94)
    A = torch.eye(3).repeat(N, 1, 1) + randn((N, 3, 3), torch.float32, 95, 0.1)
    T = randn((N, 3), torch.float32, 96, 1.5)
    go = randn((N, C) + sh, torch.float32, 97)
    assert relerr(lm.affine_interp(I.cuda(), A.cuda(), T.cuda()), orc.affine_interp_forward(I, A, T)) <= 5e-5
    grads = {}
    for dt in (torch.float32, torch.float64):
        Ic, Ac, Tc = (t.to(dt).cuda().requires_grad_(True) for t in (I, A, T))
        grads[dt] = torch.autograd.grad(lm.affine_interp(Ic, Ac, Tc), [Ic, Ac, Tc], go.to(dt).cuda())
    for g32, g64 in zip(grads[torch.float32], grads[torch.float64]):
        assert relerr(g32.double(), g64) <= 1e-4
    # single gradients take other template instances (d_I only / d_A, d_T only)
    Ic, Ac, Tc = (t.cuda().requires_grad_(True) for t in (I, A, T))
    out = lm.affine_interp(Ic, Ac, Tc)
    (dI,) = torch.autograd.grad(out, [Ic], go.cuda(), retain_graph=True)
    dA, dT = torch.autograd.grad(out, [Ac, Tc], go.cuda())
    assert relerr(dI.double(), grads[torch.float64][0]) <= 1e-4
    assert relerr(dA.double(), grads[torch.float64][1]) <= 1e-4 and relerr(dT.double(), grads[torch.float64][2]) <= 1e-4


def test_empty_and_errors(lm):
    z = torch.zeros(0, 3, 4, 4, 4, device="cuda")
    assert lm.interp(z, z).shape == z.shape
    assert lm.jacobian_times_vectorfield(z, z).shape == z.shape
    with pytest.raises(RuntimeError):
        lm.interp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))  # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        lm.jacobian_times_vectorfield(torch.zeros(1, 3, 4, 4, 1).cuda(), torch.zeros(1, 3, 4, 4, 1).cuda())  # thin
    with pytest.raises(RuntimeError):
        lm.interp(torch.zeros(2, 1, 4, 4).cuda(), torch.zeros(3, 2, 4, 4).cuda())  # batch mismatch


def test_debug_mode_and_launch_count(lm):
    lm.set_debug_mode(True)
    n0 = lm.launch_count()
    x = torch.randn(1, 2, 8, 8, device="cuda")
    lm.interp(x, x)
    assert lm.launch_count() == n0 + 1
    lm.set_debug_mode(False)


def test_expmap_host_pipeline(lm, orc):
    """host-buffer API (chunked, three streams) == plain device shoot, incl. a ragged last chunk"""
    params = [0.1, 0.0, 0.01]
    gm = lm.FluidMetric(params)
    m0 = smooth_field((5, 3, 16, 16, 16), torch.float32, 72, amp=1.0, sigma=2.0)
    m0 = (m0 * (3.0 / orc.FluidMetric(params).sharp(m0).abs().max())).pin_memory()
    ref = lm.expmap(gm, m0.cuda(), num_steps=3).cpu()
    out = lm.expmap_host(gm, m0, num_steps=3, chunk=2, graphs=False)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    # the cached CUDA-graph plan: first call captures, later calls replay on new data
    for chunk in (2, "auto"):
        out = lm.expmap_host(gm, m0, num_steps=3, chunk=chunk)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    m1 = (m0 * 0.5).pin_memory()
    ref1 = lm.expmap(gm, m1.cuda(), num_steps=3).cpu()
    out1 = lm.expmap_host(gm, m1, num_steps=3, chunk="auto")
    torch.cuda.synchronize()
    assert torch.equal(out1, ref1)
    from lagomorph_b200 import lddmm
    assert len(lddmm._HOST_PLANS) >= 1, "expmap_host fell back to the eager path"
    # two and three compute streams (consecutive chunks' graphs in flight at once), repeated calls on
    # alternating data so that a chunk's graph is replayed while its predecessor's results are still in flight
    for ns, chunk in ((2, [1, 1, 1, 1, 1]), (2, [1, 2, 2]), (3, 1), (2, "auto")):
        for rep in range(3):
            a, r = (m0, ref) if rep % 2 == 0 else (m1, ref1)
            out = lm.expmap_host(gm, a, num_steps=3, chunk=chunk, streams=ns)
            torch.cuda.synchronize()
            assert torch.equal(out, r), (ns, chunk, rep)


@pytest.mark.parametrize("dim,sh", [(2, (16, 16)), (3, (8, 16, 16))])
@pytest.mark.parametrize("dtype", DTYPES)
def test_dagger_sym_and_compose_variants(lm, orc, dim, sh, dtype):
    """adjrep.py:104-145 compositions and the compose_* wrappers (deform.py:58-70)"""
    params = [0.5, 0.0, 0.5]
    om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
    x = randn((2, dim) + sh, dtype, 201)
    y = randn((2, dim) + sh, dtype, 202)
    phi = randn((2, dim) + sh, dtype, 203, 0.5)
    tol = 5e-5 if dtype == torch.float32 else 1e-10
    xc, yc, pc = x.cuda(), y.cuda(), phi.cuda()
    assert relerr(lm.ad_dagger(xc, yc, gm), orc.ad_dagger(x, y, om)) <= tol
    assert relerr(lm.Ad_dagger(pc, yc, gm), orc.Ad_dagger(phi, y, om)) <= tol
    assert relerr(lm.sym(xc, yc, gm), orc.sym(x, y, om)) <= tol
    assert relerr(lm.sym_dagger(xc, yc, gm), orc.sym_dagger(x, y, om)) <= tol
    assert relerr(lm.compose_disp_vel(pc, xc, dt=-0.2), orc.compose_disp_vel(phi, x, dt=-0.2)) <= tol_for(dtype)
    assert relerr(lm.compose_vel_disp(xc, pc, dt=0.3), orc.compose_vel_disp(x, phi, dt=0.3)) <= tol_for(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_expmap_mommask(lm, orc, dtype):
    params = [0.1, 0.0, 0.01]
    om, gm = orc.FluidMetric(params), lm.FluidMetric(params)
    sh = (16, 16, 16)
    m0 = smooth_field((2, 3) + sh, torch.float64, 73, amp=1.0, sigma=2.0)
    m0 = (m0 * (2.0 / om.sharp(m0).abs().max())).to(dtype)
    mask = (randn((2, 3) + sh, dtype, 74) > 0).to(dtype)
    ref = orc.expmap(om, m0, num_steps=3, mommask=mask)
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    assert relerr(lm.expmap(gm, m0.cuda(), num_steps=3, mommask=mask.cuda()), ref) <= tol          # fused path
    bm = mask[:, :1]                                                                                # broadcast mask
    ref = orc.expmap(om, m0, num_steps=3, mommask=bm)
    assert relerr(lm.expmap(gm, m0.cuda(), num_steps=3, mommask=bm.cuda()), ref) <= tol            # unfused path
