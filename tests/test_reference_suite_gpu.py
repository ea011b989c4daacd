"""The reference's own test-suite (testing/test_{interp,diff,metric,adjrep,lddmm,affine}.py),
restated against `lagomorph_b200 as lm` with the same seeds, sizes (res 2-3, float64 CUDA) and
assertions. gradcheck tests are included (the reference deselects them by default, setup.cfg:5)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TF = [True, False]


def catch_gradcheck(message, *args, **kw):  # testing/utils.py:4-10
    check = False
    msg = message
    try:
        check = torch.autograd.gradcheck(*args, **kw)
    except RuntimeError as e:
        msg = f"{str(e)} {message}"
    assert check, msg


@pytest.fixture(autouse=True)
def _seed(lm):
    lm.set_debug_mode(True)
    np.random.seed(1)
    torch.manual_seed(1)
    yield
    lm.set_debug_mode(False)


# ---- testing/test_interp.py ---------------------------------------------------------------
@pytest.mark.parametrize("nc", [1, 2, 4])
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("testI,testu", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("broadcastI", TF)
def test_interp_gradcheck(lm, bs, nc, dim, testI, testu, broadcastI):
    res = 2
    imsh = tuple([1 if broadcastI else bs, nc] + [res] * dim)
    defsh = tuple([bs, dim] + [res] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda().requires_grad_(testI)
    u = torch.randn(defsh, dtype=I.dtype).to(I.device).requires_grad_(testu)
    catch_gradcheck("Failed interp gradcheck", lm.interp, (I, u))


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("nc", [1, 2, 4])
@pytest.mark.parametrize("broadcastI", TF)
def test_interp_2d_match_3d(lm, bs, nc, broadcastI):
    res = 2
    imsh = tuple([1 if broadcastI else bs, nc] + [res] * 2)
    defsh = tuple([bs, 2] + [res] * 2)
    defsh3 = tuple([bs, 3] + [res] * 2 + [1])
    I = torch.randn(imsh, dtype=torch.float64).cuda()
    u = torch.randn(defsh, dtype=I.dtype).to(I.device)
    I3 = I.unsqueeze(4)
    u3 = torch.zeros(defsh3, dtype=u.dtype, device=u.device)
    u3[:, :2, ...] = u.unsqueeze(4)
    assert torch.allclose(lm.interp(I, u).unsqueeze(4), lm.interp(I3, u3)), "Failed interp 2d match 3d"


# ---- testing/test_diff.py -----------------------------------------------------------------
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("disp", TF)
@pytest.mark.parametrize("trans", TF)
@pytest.mark.parametrize("testphi,testm", [(True, True), (True, False), (False, True)])
def test_jacobian_times_vectorfield_gradcheck(lm, bs, dim, disp, trans, testphi, testm):
    defsh = tuple([bs, dim] + [2] * dim)
    phiinv = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(testphi)
    m = torch.randn_like(phiinv).requires_grad_(testm)
    foo = lambda v, w: lm.jacobian_times_vectorfield(v, w, displacement=disp, transpose=trans)
    catch_gradcheck("Failed jacobian_times_vectorfield gradcheck", foo, (phiinv, m))


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("disp", TF)
def test_jacobian_times_vectorfield_transpose(lm, bs, dim, disp):
    defsh = tuple([bs, dim] + [2] * dim)
    g = torch.randn(defsh, dtype=torch.float64).cuda()
    u = torch.randn_like(g)
    v = torch.randn_like(g)
    Dguv = (lm.jacobian_times_vectorfield(g, u, displacement=disp, transpose=False) * v).sum()
    uDgTv = (u * lm.jacobian_times_vectorfield(g, v, displacement=disp, transpose=True)).sum()
    assert torch.allclose(Dguv, uDgTv), "Failed jacobian_times_vectorfield_transpose"


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_jacobian_times_vectorfield_adjoint(lm, bs, dim):
    defsh = tuple([bs, dim] + [2] * dim)
    u = torch.randn(defsh, dtype=torch.float64).cuda()
    v = torch.randn_like(u)
    m = torch.randn_like(u)
    Duvm = (lm.jacobian_times_vectorfield(u, v, displacement=False, transpose=False) * m).sum()
    uadjvm = (u * lm.jacobian_times_vectorfield_adjoint(m, v)).sum()
    assert torch.allclose(Duvm, uadjvm), "Failed jacobian_times_vectorfield_adjoint"


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_jacobian_times_vectorfield_adjoint_gradcheck(lm, bs, dim):
    defsh = tuple([bs, dim] + [2] * dim)
    v = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    m = torch.randn_like(v).requires_grad_(True)
    catch_gradcheck("Failed jacobian_times_vectorfield_adjoint gradcheck",
                    lm.jacobian_times_vectorfield_adjoint, (v, m))


def _replicate(v2):
    bs = v2.shape[0]
    v3 = torch.zeros((bs, 3, 2, 2, 2), dtype=torch.float64).cuda()
    v3[:, :2, :, :, 0] = v2
    v3[:, :2, :, :, 1] = v2
    return v3


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("disp", TF)
@pytest.mark.parametrize("trans", TF)
def test_jacobian_times_vectorfield_2d_match_3d(lm, bs, disp, trans):
    v2 = torch.randn((bs, 2, 2, 2), dtype=torch.float64).cuda()
    m2 = torch.randn_like(v2)
    dvm2 = lm.jacobian_times_vectorfield(v2, m2, displacement=disp, transpose=trans)
    dvm3 = lm.jacobian_times_vectorfield(_replicate(v2), _replicate(m2), displacement=disp, transpose=trans)
    assert torch.allclose(dvm3[:, :2, :, :, 0], dvm2), "Failed jacobian_times_vectorfield 2D match 3D"


@pytest.mark.parametrize("bs", [1, 2])
def test_jacobian_times_vectorfield_adjoint_2d_match_3d(lm, bs):
    v2 = torch.randn((bs, 2, 2, 2), dtype=torch.float64).cuda()
    m2 = torch.randn_like(v2)
    dvm2 = lm.jacobian_times_vectorfield_adjoint(v2, m2)
    dvm3 = lm.jacobian_times_vectorfield_adjoint(_replicate(v2), _replicate(m2))
    assert torch.allclose(dvm3[:, :2, :, :, 0], dvm2), "Failed jacobian_times_vectorfield_adjoint 2D match 3D"


# ---- testing/test_metric.py ------------------------------------------------------------------
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_fluid_sharp_gradcheck(lm, bs, dim):
    defsh = tuple([bs, dim] + [3] * dim)
    m = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    metric = lm.FluidMetric([0.1, 0.01, 0.001])
    catch_gradcheck(f"Failed fluid sharp gradcheck with batch size {bs} dim {dim}", metric.sharp, (m,), eps=1e-4)


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_fluid_flat_gradcheck(lm, bs, dim):
    defsh = tuple([bs, dim] + [3] * dim)
    v = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    metric = lm.FluidMetric([0.1, 0.01, 0.001])
    catch_gradcheck(f"Failed fluid flat gradcheck with batch size {bs} dim {dim}", metric.flat, (v,))


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("res", [3, 16])
def test_fluid_inverse(lm, bs, dim, res):
    defsh = tuple([bs, dim] + [res] * dim)
    m = torch.randn(defsh, dtype=torch.float64).cuda()
    metric = lm.FluidMetric([0.1, 0.01, 0.001])
    vm = metric.flat(metric.sharp(m))
    assert torch.allclose(vm, m, atol=1e-3), f"Failed fluid inverse check with batch size {bs} dim {dim}"


# ---- testing/test_adjrep.py --------------------------------------------------------------------
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_Ad_star_gradcheck(lm, bs, dim):
    defsh = tuple([bs, dim] + [2] * dim)
    phiinv = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    m = torch.randn_like(phiinv)
    catch_gradcheck(f"Failed Ad_star gradcheck with batch size {bs} dim {dim}", lm.Ad_star, (phiinv, m))


# not in the reference suite: the other fused operators' backward passes
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("name", ["ad", "ad_star", "Ad_star"])
def test_fused_gradcheck(lm, dim, name):
    defsh = tuple([2, dim] + [3] * dim)
    a = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    b = torch.randn_like(a).requires_grad_(True)
    catch_gradcheck(f"Failed {name} gradcheck", getattr(lm, name), (a, b))


@pytest.mark.parametrize("dim", [2, 3])
def test_compose_gradcheck(lm, dim):
    defsh = tuple([2, dim] + [3] * dim)
    a = torch.randn(defsh, dtype=torch.float64).cuda().requires_grad_(True)
    b = torch.randn_like(a).requires_grad_(True)
    catch_gradcheck("Failed compose gradcheck", lambda u, v: lm.compose(u, v, ds=0.7, dt=-1.3), (a, b))


@pytest.mark.parametrize("dim", [2, 3])
def test_ad_star_is_adjoint_of_ad(lm, dim):
    """<ad(v,w), m> == <w, ad_star(v,m)> exactly (discrete adjoint by construction, adjrep.py:78-79)"""
    defsh = tuple([2, dim] + [5] * dim)
    v = torch.randn(defsh, dtype=torch.float64).cuda()
    w = torch.randn_like(v)
    m = torch.randn_like(v)
    assert torch.allclose((lm.ad(v, w) * m).sum(), (w * lm.ad_star(v, m)).sum())


# ---- testing/test_lddmm.py ---------------------------------------------------------------------
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("step", [1, 5])
def test_expmap_zero(lm, bs, dim, step):
    res = 128 if dim == 2 else 64  # reference uses 128 for both; 64^3 float64 keeps the run short
    defsh = tuple([bs, dim] + [res] * dim)
    m = torch.zeros(defsh, dtype=torch.float64).cuda()
    metric = lm.FluidMetric([1.0, 0.1, 0.01])
    h = lm.expmap(metric, m, num_steps=step)
    assert torch.allclose(m, h), "Failed expmap of zero is identity check"


def test_expmap_gradcheck_and_checkpointing(lm):
    """gradient through a 3-step shoot; the checkpointed path gives identical gradients"""
    metric = lm.FluidMetric([0.5, 0.0, 0.5])
    m = (0.3 * torch.randn((1, 2, 4, 4), dtype=torch.float64)).cuda().requires_grad_(True)
    catch_gradcheck("Failed expmap gradcheck", lambda x: lm.expmap(metric, x, num_steps=3), (m,))
    g = torch.randn((1, 2, 4, 4), dtype=torch.float64).cuda()
    (g1,) = torch.autograd.grad(lm.expmap(metric, m, num_steps=5), [m], g)
    (g2,) = torch.autograd.grad(lm.expmap(metric, m, num_steps=5, checkpoints=2), [m], g)
    (g3,) = torch.autograd.grad(lm.expmap(metric, m, num_steps=5, checkpoints=True), [m], g)
    assert torch.allclose(g1, g2, rtol=1e-10, atol=1e-12) and torch.allclose(g1, g3, rtol=1e-10, atol=1e-12)


# ---- testing/test_affine.py (regrid + affine parts) -------------------------------------------
@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_affine_interp_identity(lm, bs, dim):
    imsh = tuple([bs, 1] + [2] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda()
    A = torch.eye(dim, dtype=I.dtype).view(1, dim, dim).repeat(bs, 1, 1).cuda()
    T = torch.zeros((bs, dim), dtype=I.dtype).cuda()
    assert torch.allclose(lm.affine_interp(I, A, T), I), "Failed affine interp identity check"


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("c", [1, 2, 4])
@pytest.mark.parametrize("which", ["I", "A", "T", "all"])
def test_affine_interp_gradcheck(lm, bs, dim, c, which):
    imsh = tuple([bs, c] + [2] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda().requires_grad_(which in ("I", "all"))
    A = torch.randn((bs, dim, dim), dtype=I.dtype).cuda().requires_grad_(which in ("A", "all"))
    T = torch.randn((bs, dim), dtype=I.dtype).cuda().requires_grad_(which in ("T", "all"))
    catch_gradcheck("Failed affine interp gradcheck", lm.affine_interp, (I, A, T))


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_affine_inverse(lm, bs, dim):
    A = torch.randn((bs, dim, dim), dtype=torch.float64).cuda()
    T = torch.randn((bs, dim), dtype=torch.float64).cuda()
    Ainv, Tinv = lm.affine_inverse(A, T)
    x = torch.randn((bs, dim, 1), dtype=torch.float64).cuda()
    y = torch.matmul(A, x) + T.unsqueeze(2)
    assert torch.allclose(torch.matmul(Ainv, y) + Tinv.unsqueeze(2), x)


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("disp", TF)
def test_regrid_identity(lm, bs, dim, disp):
    imsh = tuple([bs, dim] + [2] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda()
    assert torch.allclose(I, lm.regrid(I, shape=imsh[2:], displacement=disp)), "Failed regrid identity check"


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("c", [1, 2, 4])
def test_regrid_gradcheck(lm, bs, dim, c):
    imsh = tuple([bs, c] + [2] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda().requires_grad_(True)
    catch_gradcheck("Failed regrid gradcheck", lambda J: lm.regrid(J, shape=[3] * dim, displacement=False), (I,))


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_regrid_displacement_gradcheck(lm, bs, dim):
    imsh = tuple([bs, dim] + [2] * dim)
    I = torch.randn(imsh, dtype=torch.float64).cuda().requires_grad_(True)
    catch_gradcheck("Failed regrid displacement gradcheck",
                    lambda J: lm.regrid(J, shape=[3] * dim, displacement=True), (I,))
