#!/usr/bin/env python
"""Operator-level micro-benchmark (secondary to bench.py): every hot-path operator of SURVEY.md
section 8(a) at the C2 size (N=16, 128^3, fp32), CUDA-event timed, reported as algorithmic GB/s
against the measured HBM peak. Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import lagomorph_b200 as lm  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N, n = int(os.environ.get("OPS_N", 16)), int(os.environ.get("OPS_SIZE", 128))
    peak = 6551.7
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    V = n ** 3
    f = lambda c=3, b=N: torch.randn(b, c, n, n, n, device=dev, generator=g)
    metric = lm.FluidMetric([0.1, 0.0, 0.01])
    v, w, m = f(), f(), f()
    with torch.no_grad():   # smooth displacement, max 3 voxels (realistic gather locality)
        u = metric.sharp(f())
        u = u * (3.0 / u.abs().max())
    I1, I3 = f(1), f(3)
    Ib = f(1, 1)
    go1, go3 = f(1), f(3)
    res = {}

    def rec(name, ms, bytes_per_voxel):
        gbs = bytes_per_voxel * N * V / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "alg_B_per_voxel": bytes_per_voxel, "GBps": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / peak, 3)}

    with torch.no_grad():
        rec("interp C=3", timeit(lambda: lm.interp(I3, u)), 36)
        rec("interp C=1", timeit(lambda: lm.interp(I1, u)), 20)
        rec("interp C=1 broadcast image", timeit(lambda: lm.interp(Ib, u)), 16)
        rec("interp_adjoint (splat) C=1", timeit(lambda: lm.interp_adjoint(go1, u)), 20)
        rec("interp_adjoint (splat) C=1 broadcast", timeit(lambda: lm.interp_adjoint(go1, u, broadcast=True)), 16)
        rec("jacobian_times_vectorfield", timeit(lambda: lm.jacobian_times_vectorfield(v, w)), 36)
        rec("jacobian_times_vectorfield transpose", timeit(lambda: lm.jacobian_times_vectorfield(v, w, False, True)), 36)
        rec("jacobian_times_vectorfield_adjoint", timeit(lambda: lm.jacobian_times_vectorfield_adjoint(v, w)), 36)
        rec("ad", timeit(lambda: lm.ad(v, w)), 36)
        rec("ad_star", timeit(lambda: lm.ad_star(v, m)), 36)
        rec("Ad_star", timeit(lambda: lm.Ad_star(u, m)), 36)
        rec("compose", timeit(lambda: lm.compose(u, v, -0.1, 1.0)), 36)
        rec("sharp (beta=0)", timeit(lambda: metric.sharp(m)), 24)
        rec("flat (beta=0)", timeit(lambda: metric.flat(m)), 24)
        mb = lm.FluidMetric([0.1, 0.01, 0.001])
        rec("sharp (beta!=0)", timeit(lambda: mb.sharp(m)), 24)
    # backward passes (autograd Functions)
    ug = u.clone().requires_grad_(True)
    Ig = I3.clone().requires_grad_(True)
    out = lm.interp(Ig, ug)
    rec("interp backward (d_I + d_u) C=3", timeit(lambda: torch.autograd.grad(out, [Ig, ug], go3, retain_graph=True)), 72)
    vg, wg = v.clone().requires_grad_(True), w.clone().requires_grad_(True)
    out2 = lm.jacobian_times_vectorfield(vg, wg)
    rec("jtvf backward (d_v + d_w)", timeit(lambda: torch.autograd.grad(out2, [vg, wg], go3, retain_graph=True)), 60)
    del out, out2, ug, Ig, vg, wg
    torch.cuda.empty_cache()
    # BASELINE config 4: affine_interp at 192^3, batch 32 (forward 8 B/voxel; backward reads gout and I,
    # splats d_I: 16 B/voxel + the d_A / d_T reductions)
    N4, n4 = int(os.environ.get("OPS_N4", 32)), int(os.environ.get("OPS_SIZE4", 192))
    V4 = n4 ** 3
    I4 = torch.randn(N4, 1, n4, n4, n4, device=dev, generator=g)
    A4 = torch.eye(3, device=dev).repeat(N4, 1, 1) + 0.05 * torch.randn(N4, 3, 3, device=dev, generator=g)
    T4 = 2.0 * torch.randn(N4, 3, device=dev, generator=g)
    go4 = torch.randn(N4, 1, n4, n4, n4, device=dev, generator=g)

    def rec4(name, ms, bpv):
        gbs = bpv * N4 * V4 / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "alg_B_per_voxel": bpv, "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                     "size": [N4, 1, n4, n4, n4]}
    with torch.no_grad():
        rec4("affine_interp forward (C4)", timeit(lambda: lm.affine_interp(I4, A4, T4)), 8)
    Ig4, Ag4, Tg4 = I4.clone().requires_grad_(True), A4.clone().requires_grad_(True), T4.clone().requires_grad_(True)
    out4 = lm.affine_interp(Ig4, Ag4, Tg4)
    rec4("affine_interp backward (d_I + d_A + d_T) (C4)",
         timeit(lambda: torch.autograd.grad(out4, [Ig4, Ag4, Tg4], go4, retain_graph=True)), 16)
    print(json.dumps({"N": N, "size": n, "hbm_peak_gbs": peak, "ops": res}))


if __name__ == "__main__":
    main()
