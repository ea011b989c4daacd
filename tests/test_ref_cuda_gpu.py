"""The reference's OWN CUDA kernels (oracle/_ref/libref_cuda.so, built unmodified from
/root/reference against an ATen stand-in) run side by side with the oracle and the product on the
GPU box. This is the strongest pin: reference output == oracle == product."""
import pytest
import torch

import util
from util import RefCuda, randn, relerr, tol_for

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RefCuda.available(), reason="oracle/_ref/libref_cuda.so not built")]

DTYPES = [torch.float32, torch.float64]


@pytest.fixture(scope="module")
def rc():
    return RefCuda()


def _u(N, dim, sh, dtype, seed):
    u = randn((N, dim) + sh, dtype, seed, 2.0)
    u[:, :, 0] -= 3.0
    u[..., -1] += 3.5
    return u


@pytest.mark.parametrize("dim,sh", [(2, (9, 11)), (3, (6, 7, 9))])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bcast", [False, True])
def test_interp(lm, orc, rc, dim, sh, dtype, bcast):
    N, C = 2, 3
    I = randn((1 if bcast else N, C) + sh, dtype, 1)
    u = _u(N, dim, sh, dtype, 2)
    go = randn((N, C) + sh, dtype, 3)
    ref = rc.interp_fwd(I.cuda(), u.cuda(), 0.6)
    assert relerr(orc.interp(I, u, 0.6), ref) <= tol_for(dtype)
    assert relerr(lm.interp(I.cuda(), u.cuda(), 0.6), ref) <= tol_for(dtype)
    dI_ref, du_ref = rc.interp_bwd(go.cuda(), I.cuda(), u.cuda(), 0.6)
    dI_o, du_o = orc.interp_backward(go, I, u, 0.6)
    assert relerr(dI_o, dI_ref) <= tol_for(dtype, True) and relerr(du_o, du_ref) <= tol_for(dtype)
    Ic, uc = I.cuda().requires_grad_(True), u.cuda().requires_grad_(True)
    dI, du = torch.autograd.grad(lm.interp(Ic, uc, 0.6), [Ic, uc], go.cuda())
    assert relerr(dI, dI_ref) <= tol_for(dtype, True) and relerr(du, du_ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh", [(2, (9, 11)), (3, (6, 7, 9))])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("disp", [False, True])
@pytest.mark.parametrize("trans", [False, True])
def test_jtvf(lm, orc, rc, dim, sh, dtype, disp, trans):
    N = 2
    v, w, go = (randn((N, dim) + sh, dtype, s) for s in (4, 5, 6))
    ref = rc.jtvf_fwd(v.cuda(), w.cuda(), disp, trans)
    assert relerr(orc.jtvf_forward(v, w, disp, trans), ref) <= tol_for(dtype)
    assert relerr(lm.jacobian_times_vectorfield(v.cuda(), w.cuda(), disp, trans), ref) <= tol_for(dtype)
    dv_ref, dw_ref = rc.jtvf_bwd(go.cuda(), v.cuda(), w.cuda(), disp, trans)
    dv_o, dw_o = orc.jtvf_backward(go, v, w, disp, trans)
    assert relerr(dv_o, dv_ref) <= tol_for(dtype) and relerr(dw_o, dw_ref) <= tol_for(dtype)
    vc, wc = v.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    dv, dw = torch.autograd.grad(lm.jacobian_times_vectorfield(vc, wc, disp, trans), [vc, wc], go.cuda())
    assert relerr(dv, dv_ref) <= tol_for(dtype) and relerr(dw, dw_ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh", [(2, (9, 11)), (3, (6, 7, 9))])
@pytest.mark.parametrize("dtype", DTYPES)
def test_jtvf_adjoint(lm, orc, rc, dim, sh, dtype):
    N = 2
    z, w, go = (randn((N, dim) + sh, dtype, s) for s in (7, 8, 9))
    ref = rc.jtvf_adj_fwd(z.cuda(), w.cuda())
    assert relerr(orc.jtvf_adjoint_forward(z, w), ref) <= tol_for(dtype)
    assert relerr(lm.jacobian_times_vectorfield_adjoint(z.cuda(), w.cuda()), ref) <= tol_for(dtype)
    dz_ref, dw_ref = rc.jtvf_adj_bwd(go.cuda(), z.cuda(), w.cuda())
    dz_o, dw_o = orc.jtvf_adjoint_backward(go, z, w)
    assert relerr(dz_o, dz_ref) <= tol_for(dtype) and relerr(dw_o, dw_ref) <= tol_for(dtype)
    zc, wc = z.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    dz, dw = torch.autograd.grad(lm.jacobian_times_vectorfield_adjoint(zc, wc), [zc, wc], go.cuda())
    assert relerr(dz, dz_ref) <= tol_for(dtype) and relerr(dw, dw_ref) <= tol_for(dtype)


@pytest.mark.parametrize("dim,sh", [(2, (6, 10)), (3, (4, 6, 10))])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("params", [[0.1, 0.0, 0.01], [0.1, 0.01, 0.001]])
@pytest.mark.parametrize("inverse", [True, False])
def test_fluid_operator_and_metric(lm, orc, rc, dim, sh, dtype, params, inverse):
    """reference pipeline on the GPU: torch.fft.rfftn(ortho) -> reference fluid kernel -> irfftn"""
    m = randn((2, dim) + sh, dtype, 10)
    dims = tuple(range(2, 2 + dim))
    cos, sin = orc.FluidMetric.luts(m.shape, dtype)
    F = torch.view_as_real(torch.fft.rfftn(m.cuda(), dim=dims, norm="ortho")).contiguous()
    F_o = F.cpu().clone()
    F_p = F.clone()
    rc.fluid_operator(F, inverse, [c.cuda() for c in cos], [s.cuda() for s in sin], *params)
    orc.fluid_operator(F_o, inverse, cos, sin, *params)
    lm.fluid_operator(F_p, inverse, [c.cuda() for c in cos], [s.cuda() for s in sin], *params)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert relerr(F_o, F) <= tol and relerr(F_p, F) <= tol
    ref = torch.fft.irfftn(torch.view_as_complex(F), s=m.shape[2:], dim=dims, norm="ortho")
    gm = lm.FluidMetric(params)
    out = gm.sharp(m.cuda()) if inverse else gm.flat(m.cuda())
    assert util.l2err(out, ref) <= (1e-5 if dtype == torch.float32 else 1e-11)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
def test_regrid(lm, orc, rc, dim, dtype):
    sh = (9, 7) if dim == 2 else (6, 9, 7)
    osh = (13, 12) if dim == 2 else (11, 13, 12)
    I = randn((2, 2) + sh, dtype, 11)
    origin = tuple((s - 1) * 0.5 for s in sh)
    spacing = tuple((a - 1) / (b - 1) for a, b in zip(sh, osh))
    ref = rc.regrid_fwd(I.cuda(), osh, origin, spacing)
    assert relerr(orc.regrid_forward(I, osh, origin, spacing), ref) <= tol_for(dtype)
    assert relerr(lm.regrid(I.cuda(), shape=osh), ref) <= tol_for(dtype)
    go = randn(tuple(ref.shape), dtype, 12)
    dref = rc.regrid_bwd(go.cuda(), sh, osh, origin, spacing)
    assert relerr(orc.regrid_backward(go, sh, osh, origin, spacing), dref) <= tol_for(dtype, True)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bcast", [False, True])
def test_affine(lm, orc, rc, dim, dtype, bcast):
    sh = (9, 7) if dim == 2 else (6, 9, 7)
    N, C = 3, 2
    I = randn((1 if bcast else N, C) + sh, dtype, 13)
    A = torch.eye(dim, dtype=dtype).repeat(N, 1, 1) + randn((N, dim, dim), dtype, 14, 0.1)
    T = randn((N, dim), dtype, 15, 1.5)
    go = randn((N, C) + sh, dtype, 16)
    ref = rc.affine_fwd(I.cuda(), A.cuda(), T.cuda())
    tol = 5e-5 if dtype == torch.float32 else 1e-12
    assert relerr(orc.affine_interp_forward(I, A, T), ref) <= tol
    assert relerr(lm.affine_interp(I.cuda(), A.cuda(), T.cuda()), ref) <= tol
    dI_ref, dA_ref, dT_ref = rc.affine_bwd(go.cuda(), I.cuda(), A.cuda(), T.cuda())
    Ic, Ac, Tc = (t.cuda().requires_grad_(True) for t in (I, A, T))
    dI, dA, dT = torch.autograd.grad(lm.affine_interp(Ic, Ac, Tc), [Ic, Ac, Tc], go.cuda())
    tolg = 1e-4 if dtype == torch.float32 else 1e-11
    assert relerr(dI, dI_ref) <= tolg and relerr(dA, dA_ref) <= tolg and relerr(dT, dT_ref) <= tolg
