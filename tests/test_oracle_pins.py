"""CPU-only: pin the oracle's per-point arithmetic against the reference's OWN headers
(include/interp.h, extrap.h, diff.h) instantiated on the host: oracle/_ref/libref_points.so.
Bit-exact on random and edge coordinates, float and double."""
import ctypes
import os

import numpy as np
import pytest
import torch

from util import REF_POINTS

pytestmark = pytest.mark.skipif(not os.path.exists(REF_POINTS), reason="oracle/_ref/libref_points.so not built")


@pytest.fixture(scope="module")
def ref():
    return ctypes.CDLL(REF_POINTS)


def _coords(rng, n, lo, hi):
    c = rng.uniform(lo, hi, size=n)
    edge = np.array([-2.3, -1.0, -1e-9, 0.0, 0.5, 1.0, hi - 1.0, hi - 1e-7, hi, hi + 3.7, -0.0])
    return np.concatenate([edge, c])


@pytest.mark.parametrize("dtype,suffix,ct", [(np.float32, "f32", ctypes.c_float), (np.float64, "f64", ctypes.c_double)])
def test_points_3d(orc, ref, dtype, suffix, ct):
    rng = np.random.RandomState(1)
    nx, ny, nz = 4, 5, 6
    img = rng.randn(nx, ny, nz).astype(dtype)
    I = torch.from_numpy(img).reshape(1, 1, nx, ny, nz)
    tri = getattr(ref, "ref_trilerp_" + suffix)
    tri.restype = ct
    trig = getattr(ref, "ref_trilerp_grad_" + suffix)
    spl = getattr(ref, "ref_splat3_" + suffix)
    xs, ys, zs = _coords(rng, 60, -2, nx + 1), _coords(rng, 60, -2, ny + 1), _coords(rng, 60, -2, nz + 1)
    ip = img.ctypes.data_as(ctypes.c_void_p)
    for x, y, z in zip(xs, ys, zs):
        x, y, z = dtype(x), dtype(y), dtype(z)
        # oracle: voxel (0,0,0) displaced by (x,y,z) with dt = 1
        u = torch.zeros(1, 3, nx, ny, nz, dtype=I.dtype)
        u[0, 0, 0, 0, 0], u[0, 1, 0, 0, 0], u[0, 2, 0, 0, 0] = float(x), float(y), float(z)
        val = orc.interp(I, u)[0, 0, 0, 0, 0].item()
        assert val == tri(ip, ct(x), ct(y), ct(z), nx, ny, nz)
        go = torch.zeros_like(I)
        go[0, 0, 0, 0, 0] = 1.5
        dI, du = orc.interp_backward(go, I, u)
        o = (ct * 4)()
        trig(o, ip, ct(x), ct(y), ct(z), nx, ny, nz)
        assert [du[0, d, 0, 0, 0].item() for d in range(3)] == [dtype(o[1]) * dtype(1.5), dtype(o[2]) * dtype(1.5), dtype(o[3]) * dtype(1.5)]
        d = np.zeros((nx, ny, nz), dtype=dtype)
        spl(d.ctypes.data_as(ctypes.c_void_p), ct(1.5), ct(x), ct(y), ct(z), nx, ny, nz)
        # only voxel (0,0,0) carries mass 1.5; every other voxel splats 0 at its own position
        assert np.array_equal(dI[0, 0].numpy(), d)


@pytest.mark.parametrize("dtype,suffix,ct", [(np.float32, "f32", ctypes.c_float), (np.float64, "f64", ctypes.c_double)])
def test_points_2d(orc, ref, dtype, suffix, ct):
    rng = np.random.RandomState(2)
    nx, ny = 5, 7
    img = rng.randn(nx, ny).astype(dtype)
    I = torch.from_numpy(img).reshape(1, 1, nx, ny)
    bil = getattr(ref, "ref_bilerp_" + suffix)
    bil.restype = ct
    bilg = getattr(ref, "ref_bilerp_grad_" + suffix)
    spl = getattr(ref, "ref_splat2_" + suffix)
    ip = img.ctypes.data_as(ctypes.c_void_p)
    for x, y in zip(_coords(rng, 60, -2, nx + 1), _coords(rng, 60, -2, ny + 1)):
        x, y = dtype(x), dtype(y)
        u = torch.zeros(1, 2, nx, ny, dtype=I.dtype)
        u[0, 0, 0, 0], u[0, 1, 0, 0] = float(x), float(y)
        assert orc.interp(I, u)[0, 0, 0, 0].item() == bil(ip, ct(x), ct(y), nx, ny)
        go = torch.zeros_like(I)
        go[0, 0, 0, 0] = -0.75
        dI, du = orc.interp_backward(go, I, u)
        o = (ct * 3)()
        bilg(o, ip, ct(x), ct(y), nx, ny)
        assert [du[0, d, 0, 0].item() for d in range(2)] == [dtype(o[1]) * dtype(-0.75), dtype(o[2]) * dtype(-0.75)]
        d = np.zeros((nx, ny), dtype=dtype)
        spl(d.ctypes.data_as(ctypes.c_void_p), ct(-0.75), ct(x), ct(y), nx, ny)
        assert np.array_equal(dI[0, 0].numpy(), d)


@pytest.mark.parametrize("dtype,suffix,ct", [(np.float32, "f32", ctypes.c_float), (np.float64, "f64", ctypes.c_double)])
def test_grad_point(orc, ref, dtype, suffix, ct):
    rng = np.random.RandomState(3)
    nx, ny, nz = 3, 4, 5
    a = rng.randn(nx, ny, nz).astype(dtype)
    f = torch.from_numpy(a).reshape(1, 1, nx, ny, nz)
    g3 = getattr(ref, "ref_grad3_" + suffix)
    outs = []
    for d in range(3):
        w = torch.zeros(1, 3, nx, ny, nz, dtype=f.dtype)
        w[0, d] = 1
        outs.append(orc.jtvf_forward(f, w, False, False)[0, 0])
    o = (ct * 3)()
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                g3(o, a.ctypes.data_as(ctypes.c_void_p), nx, ny, nz, i, j, k)
                assert [outs[d][i, j, k].item() for d in range(3)] == [o[0], o[1], o[2]]
    a2 = rng.randn(nx, ny).astype(dtype)
    f2 = torch.from_numpy(a2).reshape(1, 1, nx, ny)
    g2 = getattr(ref, "ref_grad2_" + suffix)
    o2 = (ct * 2)()
    for d in range(2):
        w = torch.zeros(1, 2, nx, ny, dtype=f2.dtype)
        w[0, d] = 1
        out = orc.jtvf_forward(f2, w, False, False)[0, 0]
        for i in range(nx):
            for j in range(ny):
                g2(o2, a2.ctypes.data_as(ctypes.c_void_p), nx, ny, i, j)
                assert out[i, j].item() == o2[d]
