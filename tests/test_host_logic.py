"""CPU-only: host-side logic of the Python mirror (no kernels are launched)."""
import numpy as np
import pytest
import torch


def test_identity(lm):
    ix = lm.identity((2, 3, 4, 5, 6))
    assert ix.shape == (2, 3, 4, 5, 6) and ix.dtype == np.float32
    assert ix[1, 0, 3, 0, 0] == 3 and ix[0, 1, 0, 4, 0] == 4 and ix[0, 2, 0, 0, 5] == 5


def test_regrid_argument_rules(lm):
    from lagomorph_b200.affine import regrid_args
    sh, o, s = regrid_args((5, 9), shape=(9, 17))
    assert sh == (9, 17) and o == (2.0, 4.0) and s == (0.5, 0.5)
    sh, o, s = regrid_args((5, 9, 3), shape=4)
    assert sh == (4, 4, 4)
    sh, o, s = regrid_args((5, 9), shape=(9, 17), spacing=2.0)
    assert s == (2.0, 2.0) and o == (2.0, 4.0)
    with pytest.raises(ValueError):
        regrid_args((5, 9))
    with pytest.raises(NotImplementedError):
        regrid_args((5, 9), spacing=1.0)
    with pytest.raises(ValueError):
        regrid_args((5, 9), origin=1.0, spacing=1.0)
    with pytest.raises(NotImplementedError):
        regrid_args((5, 9), shape=(3, 3), origin=(0, 0))


def test_cpu_tensors_fail_loudly(lm):
    x = torch.zeros(1, 2, 4, 4)
    for call in (lambda: lm.interp(x, x), lambda: lm.jacobian_times_vectorfield(x, x),
                 lambda: lm.jacobian_times_vectorfield_adjoint(x, x), lambda: lm.FluidMetric().sharp(x),
                 lambda: lm.Ad_star(x, x), lambda: lm.ad_star(x, x), lambda: lm.regrid(x, shape=(8, 8)),
                 lambda: lm.expmap(lm.FluidMetric(), x)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_api_surface(lm):
    names = ["set_debug_mode", "interp", "interp_adjoint", "jacobian_times_vectorfield",
             "jacobian_times_vectorfield_adjoint", "FluidMetric", "Metric", "Ad_star", "ad_star", "ad", "Ad",
             "coad", "ad_dagger", "Ad_dagger", "sym", "sym_dagger", "expmap", "expmap_advect", "EPDiff_step",
             "EPDiffStep", "affine_interp", "affine_inverse", "regrid", "compose", "compose_disp_vel",
             "compose_vel_disp", "identity", "LDDMMAtlasBuilder"]
    for n in names:
        assert hasattr(lm, n), n
    with pytest.raises(NotImplementedError):
        lm.Ad(None, None)
    m = lm.FluidMetric([0.1, 0.0, 0.01])
    assert m.params == [0.1, 0.0, 0.01]


def test_metric_from_args(lm):
    import argparse
    p = argparse.ArgumentParser()
    lm.Metric.add_args(p)
    a = p.parse_args(["--fluid_alpha", "0.5"])
    m = lm.Metric.from_args(a)
    assert m.params == [0.5, 0.0, 0.01]


def test_shard_indices():
    from lagomorph_b200.atlas import shard_indices
    assert shard_indices(8, 1, 0) == list(range(8))
    assert shard_indices(8, 4, 1, shuffle=False) == [1, 5]
    # the reference builds DistributedSampler(dataset, num_replicas, rank) with its defaults
    # (lddmm.py:163-178, affine.py:309-312): shuffle=True, seed 0, epoch 0, pad by wrapping
    from torch.utils.data.distributed import DistributedSampler
    for n, w in ((10, 4), (64, 8), (7, 2)):
        seen = []
        for r in range(w):
            assert list(DistributedSampler(list(range(n)), num_replicas=w, rank=r)) == shard_indices(n, w, r)
            ds = DistributedSampler(list(range(n)), num_replicas=w, rank=r, shuffle=False)
            assert list(ds) == shard_indices(n, w, r, shuffle=False)
            seen += shard_indices(n, w, r)
        assert sorted(set(seen)) == list(range(n))


def test_expmap_host_chunk_schedule(lm):
    """_auto_chunks: sizes are positive, add up to the batch, ramp up from one subject and down to one
    when copies are faster than compute, and stay at one when they are not"""
    from lagomorph_b200.lddmm import _auto_chunks
    for steps in (1, 3, 5, 10, 20, 50):
        for N in list(range(1, 40)) + [64, 100]:
            c = _auto_chunks(N, steps)
            assert sum(c) == N and min(c) >= 1, (N, steps, c)
            assert c[0] == 1 and c[-1] == 1 or N <= 3, (N, steps, c)
    assert _auto_chunks(16, 10) == [1, 2, 4, 5, 3, 1]
    assert _auto_chunks(8, 5) == [1] * 8
    # the same schedule from a measured copy / compute ratio (what expmap_host passes)
    assert _auto_chunks(16, 10, ratio=0.48) == [1, 2, 4, 5, 3, 1]
    assert _auto_chunks(8, 99, ratio=1.3) == [1] * 8
    for r in (0.01, 0.1, 0.3, 0.7, 1.0, 5.0):
        for N in (1, 2, 7, 16, 64):
            c = _auto_chunks(N, 10, ratio=r)
            assert sum(c) == N and min(c) >= 1 and max(c) <= 6  # cap 5, +1 when a stray single is merged
    # the second candidate expmap_host times against it: 1, 2, 3, ... up and down again
    from lagomorph_b200.lddmm import _ramp_chunks
    assert _ramp_chunks(16) == [1, 2, 3, 4, 3, 2, 1] and _ramp_chunks(9) == [1, 2, 3, 2, 1]
    for N in range(1, 80):
        c = _ramp_chunks(N)
        assert sum(c) == N and min(c) >= 1 and c == sorted(c[:len(c) // 2 + 1]) + sorted(c[len(c) // 2 + 1:], reverse=True), (N, c)
    # the two-stream candidates: pairs with single subjects at both ends, and single subjects throughout
    from lagomorph_b200.lddmm import _pair_chunks
    assert _pair_chunks(16) == [1, 1, 2, 2, 2, 2, 2, 2, 1, 1] and _pair_chunks(7) == [1, 1, 2, 1, 1, 1]
    for N in range(1, 40):
        c = _pair_chunks(N)
        assert sum(c) == N and min(c) >= 1 and max(c) <= 2 and c[0] == 1 and c[-1] == 1, (N, c)
    with pytest.raises(RuntimeError, match="host tensors"):
        lm.expmap_host(lm.FluidMetric(), torch.zeros(1, 3, 4, 4, 4, device="meta") if False else _FakeCuda())


class _FakeCuda:
    is_cuda = True
