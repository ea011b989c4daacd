"""lgm_expmap_fwd (the whole shoot as one library call) against the step loop it replaces
(lgm_epdiff_step_fwd per step): identical bit for bit, for every dtype / dim, with and without an
initial displacement and a momentum mask (reference: lagomorph/lddmm.py:73-91); and BASELINE
config 1 on the GPU against the CPU oracle."""
import pytest
import torch

from util import smooth_field, relerr

pytestmark = pytest.mark.gpu

PARAMS = [0.1, 0.0, 0.01]


def _inputs(lm, N, sh, dtype, seed=5, vmax=3.0):
    d = len(sh)
    m0 = smooth_field((N, d) + sh, torch.float64, seed, amp=1.0, sigma=2.0).to(dtype).cuda()
    metric = lm.FluidMetric(PARAMS)
    m0 = m0 * (vmax / metric.sharp(m0).abs().max())
    return metric, m0


def _loop(lm, metric, m0, T, steps, phiinv, mommask):
    p = torch.zeros_like(m0) if phiinv is None else phiinv
    for _ in range(steps):
        p = lm.EPDiff_step(metric, m0, T / steps, p, mommask=mommask)
    return p


@pytest.mark.parametrize("sh", [(16, 16, 16), (8, 16, 32), (32, 16, 64), (12, 10, 14), (128, 128)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("steps", [1, 2, 5])
def test_shoot_equals_step_loop(lm, sh, dtype, steps):
    metric, m0 = _inputs(lm, 2, sh, dtype)
    ref = _loop(lm, metric, m0, 1.0, steps, None, None)
    out = lm.expmap(metric, m0, T=1.0, num_steps=steps)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("sh", [(16, 16, 32), (12, 10, 14), (32, 32)])
def test_shoot_initial_displacement_and_mask(lm, sh):
    metric, m0 = _inputs(lm, 3, sh, torch.float32, seed=6)
    d = len(sh)
    phi0 = smooth_field((3, d) + sh, torch.float32, 7, amp=2.0, sigma=2.0).cuda()
    mask = (torch.rand((3, d) + sh, generator=torch.Generator().manual_seed(3)) > 0.2).float().cuda()
    for p0, mk in ((phi0, None), (None, mask), (phi0, mask)):
        ref = _loop(lm, metric, m0, 0.7, 3, p0, mk)
        out = lm.expmap(metric, m0, T=0.7, num_steps=3, phiinv=p0, mommask=mk)
        assert torch.equal(out, ref)


def test_shoot_border_flow(lm):
    """momenta strong enough to push samples across the volume border (clamped corners)"""
    metric, m0 = _inputs(lm, 1, (16, 32, 32), torch.float32, seed=8, vmax=40.0)
    ref = _loop(lm, metric, m0, 1.0, 4, None, None)
    out = lm.expmap(metric, m0, T=1.0, num_steps=4)
    assert torch.isfinite(out).all() and torch.equal(out, ref)


def test_shoot_vs_oracle_c1(lm, orc):
    """BASELINE config 1 (2-D 128x128, batch 8, 10 steps): product on the GPU vs the CPU oracle."""
    N, sh = 8, (128, 128)
    m0 = smooth_field((N, 2) + sh, torch.float64, 11, amp=1.0, sigma=4.0)
    om = orc.FluidMetric(PARAMS)
    m0 = (m0 * (4.0 / om.sharp(m0).abs().max())).float()
    ref = orc.expmap(om, m0, num_steps=10)
    out = lm.expmap(lm.FluidMetric(PARAMS), m0.cuda(), num_steps=10)
    assert relerr(out, ref) <= 1e-4


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_nonfinite_displacement_propagates(lm, axis):
    """A diverged deformation (NaN / Inf displacement) must give NaN samples, as the reference's
    t = x - floor(x) does (include/interp.h:64-92), in the fp32 3-D fast paths too -- not a finite
    border value that would mask the blow-up."""
    sh = (8, 16, 32)
    g = torch.Generator().manual_seed(2)
    I = torch.randn((1, 3) + sh, generator=g).cuda()
    for bad in (float("nan"), float("inf")):
        u = torch.zeros((1, 3) + sh, device="cuda")
        u[0, axis, 3, 5, 7] = bad
        out = lm.interp(I, u)
        assert torch.isnan(out[0, :, 3, 5, 7]).all()
        assert torch.isfinite(out).sum().item() == out.numel() - 3
        assert torch.isnan(lm.compose(u, I)[0, :, 3, 5, 7]).all()
        assert torch.isnan(lm.Ad_star(u, I)[0, :, 3, 5, 7]).all()


_RING_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[2])
import lagomorph_b200 as lm
outs = {}
for i, sh in enumerate([(8, 16, 32), (20, 24, 64), (16, 40, 128), (8, 12, 256), (5, 9, 128)]):
    g = torch.Generator().manual_seed(10 + i)
    u = (torch.rand((2, 3) + sh, generator=g) - 0.5) * 6.0      # |ds*u| < 0.3 voxel with ds = -0.1
    u[0, :, :, :2, 5:9] *= 8.0                                   # a patch that leaves the staged window
    u[1, 0, -1] += 30.0                                          # samples pushed across the x border
    if sh[2] == 256:
        u[0, 2, :, 4:6, 118:140] *= 40.0                         # across the seam of the two z tiles, past their halo
        u[1, 2, :, :, 250:] += 25.0                              # clamped at the upper z border
    v = torch.randn((2, 3) + sh, generator=g)
    outs["c%d" % i] = lm.compose(u.cuda(), v.cuda(), ds=-0.1, dt=1.0).cpu()
    outs["d%d" % i] = lm.compose_disp_vel(v.cuda(), u.cuda(), dt=0.05).cpu()
torch.save(outs, sys.argv[1])
"""


def test_ring_compose_bit_identical_to_planar_gather(tmp_path):
    """compose with the gather source staged in shared memory by bulk async copies (csrc/compose_ring.cu)
    == the planar L1 gather kernel (LGM_NO_RING=1, read once per process: hence two subprocesses),
    bit for bit, including warps that fall back because a sample leaves the staged window."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ring.py"
    script.write_text(_RING_SCRIPT)
    res = {}
    for tag, env in (("ring", {}), ("planar", {"LGM_NO_RING": "1"})):
        out = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        e.pop("LGM_NO_RING", None)
        e.update(env)
        subprocess.check_call([sys.executable, str(script), out, root], env=e)
        res[tag] = torch.load(out)
    assert res["ring"].keys() == res["planar"].keys() and len(res["ring"]) == 10
    for k in res["ring"]:
        assert torch.equal(res["ring"][k], res["planar"][k]), k


_SHORTCUT_SCRIPT = """
import sys, torch
sys.path.insert(0, sys.argv[2])
import lagomorph_b200 as lm
outs = {}
g = torch.Generator().manual_seed(9)
for name, sh, params, steps in (("c128", (2, 3, 128, 128, 128), [0.1, 0.0, 0.01], 3),
                                ("q256", (1, 3, 64, 256, 256), [0.1, 0.0, 0.01], 2),
                                ("beta", (2, 3, 32, 32, 32), [0.1, 0.01, 0.001], 3),
                                ("odd", (1, 3, 24, 20, 36), [0.1, 0.0, 0.01], 2),
                                ("2d", (3, 2, 64, 64), [0.1, 0.0, 0.01], 4),
                                ("f64", (1, 3, 16, 16, 16), [0.1, 0.0, 0.01], 2)):
    dt = torch.float64 if name == "f64" else torch.float32
    m0 = torch.randn(sh, generator=g, dtype=dt).cuda()
    met = lm.FluidMetric(params)
    m0 = m0 * (3.0 / met.sharp(m0).abs().max())
    outs[name] = lm.expmap(met, m0, num_steps=steps).cpu()
    outs[name + "_1"] = lm.expmap(met, m0, num_steps=1).cpu()
torch.save(outs, sys.argv[1])
"""


def test_first_step_shortcut_bit_identical(tmp_path):
    """lgm_expmap_fwd from the identity runs its first step as ONE sharp with the -dt scaling in the last
    kernel (csrc/shoot3.cu). Same bits as the full step (LGM_NO_FIRST_STEP_SHORTCUT=1, read once per
    process: two subprocesses) on every FFT path: 128^2 slabs, quarter slabs, beta != 0, mixed radix, 2-D,
    fp64 -- up to the sign of exact zeros, hence the comparison with ==."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "shortcut.py"
    script.write_text(_SHORTCUT_SCRIPT)
    res = {}
    for tag, env in (("short", {}), ("full", {"LGM_NO_FIRST_STEP_SHORTCUT": "1"})):
        out = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        e.pop("LGM_NO_FIRST_STEP_SHORTCUT", None)
        e.update(env)
        subprocess.check_call([sys.executable, str(script), out, root], env=e)
        res[tag] = torch.load(out)
    assert res["short"].keys() == res["full"].keys() and len(res["short"]) == 12
    for k in res["short"]:
        a, b = res["short"][k], res["full"][k]
        assert torch.isfinite(a).all() and a.abs().max() > 0, k
        assert bool((a == b).all()), k


_QSLAB_SCRIPT = """
import sys, torch
sys.path.insert(0, sys.argv[2])
import lagomorph_b200 as lm
outs = {}
g = torch.Generator().manual_seed(13)
met = lm.FluidMetric([0.1, 0.0, 0.01])
for X in (64, 128, 256):
    m = torch.randn((1, 3, X, 256, 256), generator=g).cuda()
    outs["sharp%d" % X] = met.sharp(m).cpu()
    outs["flat%d" % X] = met.flat(m).cpu()
    if X == 64:
        outs["roundtrip_err"] = ((met.flat(met.sharp(m)) - m).norm() / m.norm()).cpu()
torch.save(outs, sys.argv[1])
"""


def test_quarter_slab_path_matches_cluster_path(tmp_path):
    """256 x 256 planes, beta == 0: the quarter-slab kernels (csrc/qslab.cuh: Y split 4 x 64, radix-4 Y
    stage inside the X pass) against the cluster slab kernels they replace (LGM_NO_QSLAB=1), X = 64, 128,
    256. Different factorisations of the same transform: equal to fp32 round-off (1e-6 relative L2; both
    are checked against the CPU oracle at 1e-5 in test_fullsize_gpu.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "qslab.py"
    script.write_text(_QSLAB_SCRIPT)
    res = {}
    for tag, env in (("qslab", {}), ("cluster", {"LGM_NO_QSLAB": "1"})):
        out = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        e.pop("LGM_NO_QSLAB", None)
        e.update(env)
        subprocess.check_call([sys.executable, str(script), out, root], env=e)
        res[tag] = torch.load(out)
    for k in res["qslab"]:
        if k == "roundtrip_err":
            assert res["qslab"][k].item() <= 5e-5
            continue
        a, b = res["qslab"][k], res["cluster"][k]
        assert ((a - b).norm() / b.norm()).item() <= 1e-6, k


_SAFE_SCRIPT = """
import sys, torch
sys.path.insert(0, sys.argv[2])
import lagomorph_b200 as lm
outs = {}
g = torch.Generator().manual_seed(19)
for name, sh, params in (("c128", (2, 3, 128, 128, 128), [0.1, 0.0, 0.01]), ("q256", (1, 3, 64, 256, 256), [0.1, 0.0, 0.01]),
                         ("x64", (2, 3, 64, 32, 64), [0.5, 0.0, 0.5]), ("x512", (1, 3, 512, 16, 32), [0.1, 0.0, 0.001]),
                         ("2d", (3, 2, 128, 128), [0.1, 0.0, 0.01]), ("edge", (1, 3, 32, 32, 32), [1e14, 0.0, 2e-4]),
                         ("tiny", (1, 3, 32, 32, 32), [0.1, 0.0, 1e-5])):
    m = torch.randn(sh, generator=g).cuda()
    met = lm.FluidMetric(params)
    outs[name] = met.sharp(m).cpu()
    outs[name + "_rt"] = met.flat(met.sharp(m)).cpu()
torch.save(outs, sys.argv[1])
"""


def test_safe_range_multiplier_bit_identical(tmp_path):
    """beta == 0, fp32: when the host can bound lambda = gamma + alpha * sw inside [2e-4, 1e15] the X passes run
    the multiplier without its safe_sqrt selects and without the reciprocal's range test (csrc/fluid.cu
    lambda_range_safe / oo_lambda_fast<R, true>). Same bits as the guarded form (LGM_NO_SAFE_LAMBDA=1, read once
    per process: two subprocesses) on the 128^2 slab path, the quarter slabs, two- and three-stage X transforms,
    2-D, parameters at the edge of the range, and parameters outside it (gamma = 1e-5: guarded form either way)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "safe.py"
    script.write_text(_SAFE_SCRIPT)
    res = {}
    for tag, env in (("safe", {}), ("guarded", {"LGM_NO_SAFE_LAMBDA": "1"})):
        out = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        e.pop("LGM_NO_SAFE_LAMBDA", None)
        e.update(env)
        subprocess.check_call([sys.executable, str(script), out, root], env=e)
        res[tag] = torch.load(out)
    assert res["safe"].keys() == res["guarded"].keys() and len(res["safe"]) == 14
    for k in res["safe"]:
        assert torch.isfinite(res["safe"][k]).all(), k
        assert torch.equal(res["safe"][k], res["guarded"][k]), k


_ARING_SCRIPT = """
import sys, torch
sys.path.insert(0, sys.argv[2])
import lagomorph_b200 as lm
from util import smooth_field
outs = {}
for i, (sh, amp) in enumerate((((2, 3, 20, 24, 128), 3.0), ((1, 3, 17, 13, 64), 5.0), ((2, 3, 9, 10, 32), 2.0),
                               ((1, 3, 6, 11, 256), 4.0), ((1, 3, 36, 8, 128), 60.0), ((1, 3, 21, 19, 256), 6.0))):
    phi = smooth_field(sh, torch.float32, 40 + i, amp=amp, sigma=2.0)
    phi[:, :, 0] -= 2.0       # border bands pushed out of range: clamped corners, clamped stencil rows
    phi[..., -1] += 2.5
    m = torch.randn(sh, generator=torch.Generator().manual_seed(50 + i))
    outs["a%d" % i] = lm.Ad_star(phi.cuda(), m.cuda()).cpu()
torch.save(outs, sys.argv[1])
"""


def test_ring_adstar_bit_identical_to_planar_kernel(tmp_path):
    """Ad_star with the stencil operand staged in a shared-memory ring (csrc/adstar_ring.cu) == the planar
    L1 kernel (LGM_NO_ADSTAR_RING=1, read once per process: two subprocesses), bit for bit: partial y
    tiles, x extents that are not a multiple of the march length, Z = 32 ... 256, borders."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "aring.py"
    script.write_text("import sys\nsys.path.insert(0, %r)\n" % os.path.join(root, "tests") + _ARING_SCRIPT)
    res = {}
    # ring: the defaults (256-voxel rows as z tiles staged by TMA tensor copies); rows: the z tiles with one
    # bulk copy per staged row (the path taken without a tensor-map entry point); wide: round 2's full rows
    variants = (("ring", {}), ("rows", {"LGM_ADSTAR_RING_TMAP": "0"}), ("wide", {"LGM_ADSTAR_RING_256": "2"}),
                ("planar", {"LGM_NO_ADSTAR_RING": "1"}))
    for tag, env in variants:
        out = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        for k in ("LGM_NO_ADSTAR_RING", "LGM_ADSTAR_RING_256", "LGM_ADSTAR_RING_TMAP"):
            e.pop(k, None)
        e.update(env)
        subprocess.check_call([sys.executable, str(script), out, root], env=e)
        res[tag] = torch.load(out)
    assert res["ring"].keys() == res["planar"].keys() and len(res["ring"]) == 6
    for tag in ("ring", "rows", "wide"):
        for k in res[tag]:
            assert torch.isfinite(res[tag][k]).all()
            assert torch.equal(res[tag][k], res["planar"][k]), (tag, k)
