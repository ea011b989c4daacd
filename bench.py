#!/usr/bin/env python
"""bench.py -- EPDiff-shoot throughput (3-D voxel-steps/s) on N B200s of one node.

One bench "step" = one forward geodesic shoot (`expmap`, num_steps EPDiff steps) of one batch of
synthetic momenta per GPU. Default workload is BASELINE.json configs[1] (C2: 3-D 128^3, batch 16,
10 steps); `--workload c3` is the per-GPU share of configs[2] (256^3, batch 8, 5 steps).

  value  : whole-job voxel-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    : the same through the public API with HOST buffers: pinned-host momenta -> device, shoot,
           deformation -> pinned host, all inside the timed region
  roofline: dominant kernel of the step, timed live with CUDA events on the launch stream
  cpu_baseline: the CPU oracle (port of the reference; the reference has no CPU path for this,
           SURVEY.md F1) on a bounded sample, rank 0 / N=1 only

  also   : c3_256 (the shoot at 256^3), c3_atlas (config 3: one atlas epoch per GPU with the NCCL
           all_reduce of the atlas gradient inside the timed region -- at every N), and at N = 1
           c2_fwd_bwd, c4_affine, c5_deep (configs 2 / 4 / 5 with autograd) and ref_cuda (the
           reference's own CUDA kernels on this GPU, oracle/_ref/libref_cuda.so)

`--impl reference` times that CPU port alone (all host threads) on the same metric/config.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (batch per GPU, (X,Y,Z), EPDiff steps per shoot)
    "c2": (16, (128, 128, 128), 10),
    "c3": (8, (256, 256, 256), 5),
    "tiny": (2, (32, 32, 32), 3),
}
PARAMS = [0.1, 0.0, 0.01]           # FluidMetric used by the atlas builder (lddmm.py:213)
ALG_BYTES_PER_VOXEL_STEP = 96       # Ad*(36) + sharp(24) + compose(36), fp32 3-D
# A shoot from the identity (phiinv=None, what expmap and the atlas builder do) runs its FIRST step as one
# sharp with the -dt scaling inside (csrc/shoot3.cu: Ad_star(0, m0) = m0, compose(0, v) = -dt v): that
# step's compulsory traffic is read m0 + write phiinv = 24 B per voxel, and the roofline fractions below
# count it as 24 B, not 96 (forward + backward: 24 + 36 instead of 324). `value` stays the plain metric,
# N * V * num_steps / time. LGM_NO_FIRST_STEP_SHORTCUT=1 runs the full first step.
FIRST_STEP_SHORTCUT = os.environ.get("LGM_NO_FIRST_STEP_SHORTCUT") is None
FIRST_STEP_BYTES = 24 if FIRST_STEP_SHORTCUT else 96
FIRST_STEP_FWD_BWD_BYTES = 60 if FIRST_STEP_SHORTCUT else 324


def shoot_bytes_per_voxel(nsteps):
    """algorithmic bytes per voxel of a whole forward shoot from the identity"""
    return FIRST_STEP_BYTES + ALG_BYTES_PER_VOXEL_STEP * (nsteps - 1)


def shoot_fwd_bwd_bytes_per_voxel(nsteps):
    return FIRST_STEP_FWD_BWD_BYTES + 324 * (nsteps - 1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_momenta(N, shape, seed, device=None, sigma=4.0):
    """BASELINE.md section 4 momenta: seeded white noise (CPU generator, deterministic per rank)
    low-passed by a separable periodic Gaussian of sigma = 4 voxels. The smoothing is input
    generation, not the product: torch.fft on `device` (or on the CPU when device is None).
    The caller scales the result so that max|sharp(m0)| = 4 voxels."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    d = len(shape)
    out = []
    for n in range(N):                      # one subject at a time bounds the FFT workspace
        m = torch.randn((1, d) + tuple(shape), generator=g, dtype=torch.float32)
        if device is not None:
            m = m.to(device)
        dims = tuple(range(2, 2 + d))
        F = torch.fft.rfftn(m, dim=dims)
        for a, n_a in enumerate(shape):
            k = torch.fft.rfftfreq(n_a) if a == d - 1 else torch.fft.fftfreq(n_a)
            w = torch.exp(-2.0 * (math.pi * sigma * k) ** 2).to(F.device)
            F = F * w.view([-1 if i == 2 + a else 1 for i in range(F.dim())])
        out.append(torch.fft.irfftn(F, s=tuple(shape), dim=dims))
    return torch.cat(out).contiguous()


def gpu_numa_node(index):
    """(node, why): the NUMA node the GPU hangs off, or (None, reason)"""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        path = "/sys/bus/pci/devices/%s/numa_node" % bus
        if not os.path.exists(path):
            return None, "no %s" % path
        node = int(open(path).read())
        if node < 0:
            return None, "sysfs reports numa_node = %d for %s (single-node box or virtualised topology)" % (node, bus)
        return node, None
    except Exception as e:
        return None, "%s: %s" % (type(e).__name__, e)


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers
    are allocated (first touch places them on that node): with 8 ranks copying at once the
    host<->device path, not the GPU, bounds the e2e number. Returns (node or None, reason)."""
    node, why = gpu_numa_node(index)
    if node is None:
        return None, why
    try:
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node, None
        return None, "node %d has no CPU in this process's affinity mask" % node
    except Exception as e:
        return None, "%s: %s" % (type(e).__name__, e)


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}



# ---- secondary workloads ("also") ---------------------------------------------------------------
FWD_BWD_BYTES_PER_VOXEL_STEP = 324  # 96 forward + 228 backward (estimate), fp32 3-D


def _timed(torch, fn, K, W, barrier, reduce_max):
    for _ in range(W):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    barrier()
    return reduce_max(e0.elapsed_time(e1)) / K


def _blob_dataset(torch, S, n, dev, seed0=100):
    """S synthetic subjects (1, n, n, n): one Gaussian blob each, centre jittered per subject;
    generated lazily on the device so that a rank only materialises its own shard."""
    ax = torch.arange(n, dtype=torch.float32, device=dev)

    class Synth:
        def __len__(self):
            return S

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(seed0 + int(i))
            c = (n - 1) / 2 + (torch.rand(3, generator=g) - 0.5) * n / 8
            e = [torch.exp(-((ax - float(c[d])) ** 2) / (2 * (n / 6) ** 2)) for d in range(3)]
            return (e[0][:, None, None] * e[1][None, :, None] * e[2][None, None, :]).unsqueeze(0)

    return Synth()


def nccl_log_summary():
    """what NCCL's own INIT log of this process says about the communicator (file set in main())"""
    import glob
    import re
    raw = os.environ.get("NCCL_DEBUG_FILE", "")
    pat = raw.replace("%p", str(os.getpid())).replace("%h", "*")
    out = {"log": pat or None}
    try:
        files = glob.glob(pat) if pat else []
        if not files and raw:   # NCCL may expand %p / %h differently: any rank's log of this launch will do
            files = sorted(glob.glob(raw.replace("%p", "*").replace("%h", "*")))[:1]
        out["log_bytes"] = sum(os.path.getsize(f) for f in files)
        txt = "".join(open(f, errors="replace").read() for f in files)
        m = re.search(r"NCCL version ([0-9.+a-z]+)", txt)
        out["version"] = m.group(1) if m else None
        m = re.search(r"nranks (\d+)", txt)
        out["nranks"] = int(m.group(1)) if m else None
        out["nvls"] = bool(re.search(r"NVLS", txt)) if txt else None
        m = re.search(r"(\d+) coll channels, (\d+) collnet channels, (\d+) nvls channels, (\d+) p2p channels", txt)
        out["channels"] = m.group(0) if m else None
        out["transport"] = sorted(set(re.findall(r"via (P2P/[A-Za-z/]+|SHM[A-Za-z/]*|NET/[A-Za-z]+)", txt)))[:4]
    except Exception as e:
        out["error"] = str(e)[:120]
    return out


def bench_c3_atlas(torch, dist, lm, dev, world, rank, hbm, barrier, reduce_max, K=2, W=1, n=256, per_gpu=8):
    """BASELINE config 3: one atlas epoch (lagomorph/lddmm.py:287-358) over this rank's 8 subjects of
    256^3 in one batch: 5-step expmap, deform, loss, backward through the shoot, momentum update, and
    the image update with the NCCL all_reduce of the 64 MiB atlas gradient (+ one 2-scalar all_reduce
    of the losses) INSIDE the timed region."""
    S = per_gpu * world
    b = lm.LDDMMAtlasBuilder(_blob_dataset(torch, S, n, dev), num_epochs=1, batch_size=per_gpu,
                             lddmm_integration_steps=5, reg_weight=1e-2, learning_rate_pose=1.0,
                             learning_rate_image=0.1, device=dev, world_size=world, rank=rank)
    b.initialize()
    n0 = lm.launch_count()
    ms = _timed(torch, b.epoch, K, W, barrier, reduce_max)
    launches = (lm.launch_count() - n0) // (K + W)
    vs = S * n ** 3 * 5 / (ms * 1e-3)
    out = {"value": S / (ms * 1e-3), "unit": "subjects/s", "ms_per_epoch": ms, "n_gpus": world,
           "voxel_steps_per_s_fwd_bwd": vs,
           "hbm_roofline_frac_324B": vs / world * (shoot_fwd_bwd_bytes_per_voxel(5) / 5) / 1e9 / hbm,
           "collective": "NCCL all_reduce of the %d MiB atlas gradient (async, overlapping the momentum update) "
                         "+ one all_reduce of 2 scalars, inside the timed region" % (n ** 3 * 4 >> 20) if world > 1
                         else "none (1 rank)",
           "nccl": nccl_log_summary() if world > 1 else None,
           "gpu_launches_per_epoch": launches, "last_epoch_loss": b.iter_losses[-1] if b.iter_losses else None,
           "config": {"shape": [n, n, n], "subjects": S, "subjects_per_gpu": per_gpu, "batch": per_gpu,
                      "epdiff_steps": 5}}
    del b
    torch.cuda.empty_cache()
    return out


def bench_c2_fwd_bwd(torch, lm, dev, hbm, metric, barrier, reduce_max, K=3, W=1, n=128, batch=16, steps=10):
    """BASELINE config 2, second figure: one pairwise-registration iteration = lddmm_step
    (lagomorph/lddmm.py:300-325) with the image held fixed: shoot, deform, loss, full backward,
    momentum update."""
    data = _blob_dataset(torch, batch, n, dev, seed0=300)
    b = lm.LDDMMAtlasBuilder(data, num_epochs=1, batch_size=batch, lddmm_integration_steps=steps, reg_weight=1e-2,
                             learning_rate_pose=1.0, learning_rate_image=0.0, device=dev, metric=metric)
    b.initialize()
    m = make_momenta(batch, (n, n, n), 7, device=dev)
    m.mul_(4.0 / metric.sharp(m).abs().max().item())
    img = b.images[b.batches[0]]

    def it():
        b.lddmm_step(m, img, need_image_grad=False)

    ms = _timed(torch, it, K, W, barrier, reduce_max)
    vs = batch * n ** 3 * steps / (ms * 1e-3)
    del b
    torch.cuda.empty_cache()
    return {"value": vs, "unit": "voxel-steps/s fwd+bwd", "ms_per_iteration": ms,
            "hbm_roofline_frac_324B": vs * (shoot_fwd_bwd_bytes_per_voxel(steps) / steps) / 1e9 / hbm,
            "config": {"shape": [n, n, n], "batch": batch, "epdiff_steps": steps}}


def bench_c4_affine(torch, lm, dev, hbm, barrier, reduce_max, K=5, W=2, n=192, batch=32):
    """BASELINE config 4: affine_interp forward + backward with all three gradients (d_I, d_A, d_T)."""
    g = torch.Generator(device=dev).manual_seed(1)
    data = _blob_dataset(torch, batch, n, dev, seed0=500)
    I = torch.stack([data[i] for i in range(batch)]).requires_grad_(True)
    A = (torch.eye(3, device=dev)[None] + 0.05 * torch.randn((batch, 3, 3), device=dev, generator=g)).requires_grad_(True)
    T = (2 * torch.randn((batch, 3), device=dev, generator=g)).requires_grad_(True)
    go = torch.randn((batch, 1, n, n, n), device=dev, generator=g)
    fwd_ms = _timed(torch, lambda: lm.affine_interp(I, A, T), K, W, barrier, reduce_max)

    def both():
        out = lm.affine_interp(I, A, T)
        torch.autograd.grad(out, [I, A, T], go)

    ms = _timed(torch, both, K, W, barrier, reduce_max)
    vox = batch * n ** 3
    return {"value": vox / (ms * 1e-3), "unit": "voxels/s fwd+bwd", "ms_fwd": fwd_ms, "ms_fwd_bwd": ms,
            "fwd_frac_of_hbm_8B": vox * 8 / (fwd_ms * 1e-3) / 1e9 / hbm,
            "bwd_frac_of_hbm_16B": vox * 16 / ((ms - fwd_ms) * 1e-3) / 1e9 / hbm,
            "config": {"shape": [n, n, n], "batch": batch}}


def bench_c5_deep(torch, lm, dev, hbm, metric, barrier, reduce_max, K=3, W=1, n=128, batch=4, steps=5):
    """BASELINE config 5 (deep LDDMM): a small 3-D CNN (3 conv layers, 16 channels; cuDNN, a library,
    timed separately) predicts the momenta (4, 3, 128^3) from the (atlas, subject) pair; the loss goes
    through expmap, interp and the metric and loss.backward() reaches the CNN weights."""
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Conv3d(2, 16, 3, padding=1), torch.nn.ReLU(),
                              torch.nn.Conv3d(16, 16, 3, padding=1), torch.nn.ReLU(),
                              torch.nn.Conv3d(16, 3, 3, padding=1)).to(dev)
    data = _blob_dataset(torch, batch + 1, n, dev, seed0=700)
    J = torch.stack([data[i] for i in range(batch)])
    I = data[batch].unsqueeze(0)
    x = torch.cat([I.expand(batch, -1, -1, -1, -1), J], dim=1).contiguous()
    with torch.no_grad():  # scale the (random-init) net's output to a flow of ~2 voxels
        gain = 2.0 / metric.sharp(net(x)).abs().max().item()

    def cnn_only():
        m = net(x) * gain
        m.backward(torch.ones_like(m))
        net.zero_grad(set_to_none=True)

    def full():
        m = net(x) * gain
        h = lm.expmap(metric, m, num_steps=steps)
        Idef = lm.interp(I, h)
        v = metric.sharp(m)
        loss = ((Idef - J) ** 2).sum() / J.numel() + 1e-2 * (v * m).sum() / J.numel()
        loss.backward()
        net.zero_grad(set_to_none=True)

    cnn_ms = _timed(torch, cnn_only, K, W, barrier, reduce_max)
    ms = _timed(torch, full, K, W, barrier, reduce_max)
    vs = batch * n ** 3 * steps / ((ms - cnn_ms) * 1e-3)
    del net
    torch.cuda.empty_cache()
    return {"value": vs, "unit": "voxel-steps/s fwd+bwd (CNN time excluded)", "ms_total": ms, "ms_cnn_fwd_bwd": cnn_ms,
            "hbm_roofline_frac_324B": vs * (shoot_fwd_bwd_bytes_per_voxel(steps) / steps) / 1e9 / hbm,
            "config": {"shape": [n, n, n], "batch": batch, "epdiff_steps": steps, "cnn": "3 x Conv3d(3^3), 16 ch (cuDNN)"}}


def bench_ref_cuda(torch, metric, lm, dev, n=128, batch=2, steps=10):
    """The reference's OWN CUDA kernels on this GPU: lagomorph/extension/cuda/*.cu compiled unmodified
    for sm_100a against an ATen stand-in (oracle/_ref/libref_cuda.so) under the reference's Python
    composition with cuFFT for torch.rfft (tests/util.py RefPipeline). Checker-side baseline."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import RefCuda, RefPipeline
    if not RefCuda.available():
        return {"unavailable": "oracle/_ref/libref_cuda.so not built (needs /root/reference at build time)"}
    ref = RefPipeline(RefCuda(), PARAMS)
    m0 = make_momenta(batch, (n, n, n), 1, device=dev)
    m0.mul_(4.0 / metric.sharp(m0).abs().max().item())
    ref.expmap(m0, 1)
    dt = None
    for _ in range(2):   # best of two whole shoots (the first still pays cuFFT plan / allocator warm-up)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        want = ref.expmap(m0, steps)
        torch.cuda.synchronize()
        dt = min(dt, time.perf_counter() - t0) if dt is not None else time.perf_counter() - t0
    got = lm.expmap(metric, m0, num_steps=steps)
    err = ((got - want).abs().max() / want.abs().max()).item()
    return {"value": batch * n ** 3 * steps / dt, "unit": "voxel-steps/s", "kind": "reference kernels, ATen shim",
            "ms_per_shoot": dt * 1e3, "sample": "%d subjects of %d^3, %d EPDiff steps" % (batch, n, steps),
            "product_vs_reference_max_rel_err": err}


def cpu_port_throughput(shape, steps, reps=2, warmup=1, batch=1):
    """voxel-steps/s of the CPU oracle (OpenMP kernels + torch.fft/MKL) on a bounded sample: `batch`
    subjects, `warmup` one-step shoots (FFT plans, page faults), then `reps` full shoots timed together.
    ONE procedure for the cpu_baseline leg of the main arm and for `--impl reference`, so that the two
    report the same number on the same box. Returns (value, seconds, cores, sample text)."""
    # torchrun exports OMP_NUM_THREADS=1; the baseline is supposed to use every host core
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    os.environ["MKL_NUM_THREADS"] = str(ncpu)
    import torch
    torch.set_num_threads(ncpu)
    from oracle import oracle as orc
    orc.lib()
    met = orc.FluidMetric(PARAMS)
    m0 = make_momenta(batch, shape, 1)
    m0 = m0 * (4.0 / met.sharp(m0).abs().max())
    for _ in range(warmup):
        orc.expmap(met, m0, num_steps=1)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.expmap(met, m0, num_steps=steps)
    dt = time.perf_counter() - t0
    V = shape[0] * shape[1] * shape[2]
    sample = "%d subject(s) of %dx%dx%d, %d EPDiff steps, %d shoot(s) after %d warm-up step(s) (%.1f s)" % (
        (batch,) + tuple(shape) + (steps, reps, warmup, dt))
    return batch * V * steps * reps / dt, dt, torch.get_num_threads(), sample


def run_reference(args):
    """The reference arm: the CPU port of the reference path, all host threads, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, shape, nsteps = WORKLOADS[args.workload]
    val, dt, cores, sample = cpu_port_throughput(shape, nsteps, reps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "3D voxel-steps/sec (EPDiff shoot)", "value": val, "unit": "voxel-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "shape": list(shape), "batch_per_gpu": batch, "epdiff_steps": nsteps,
                   "metric_params": PARAMS, "note": "CPU port of the reference path (oracle/): the reference has no CPU "
                   "implementation of interp/jacobian/fluid kernels (SURVEY.md F1) and does not build on torch 2.x (F2)"},
        "cpu_baseline": {"value": val, "unit": "voxel-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary 256^3 measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # NCCL's communicator log goes to a file (stdout must stay ONE JSON line); rank 0 summarises it in
        # also.c3_atlas.nccl. Set BEFORE torch is imported: NCCL latches its debug settings on first use.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"   # (the image presets NCCL_DEBUG=VERSION)
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/lgm_nccl_%d.%%p.log" % os.getppid())
    import torch
    import torch.distributed as dist
    import lagomorph_b200 as lm
    from lagomorph_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa, numa_why = bind_to_gpu_numa_node(local_rank) if world > 1 else (None, "single rank: not bound")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    metric = lm.FluidMetric(PARAMS)

    def measure(workload, K, W, with_e2e=True):
        batch, shape, nsteps = WORKLOADS[workload]
        V = shape[0] * shape[1] * shape[2]
        m0 = make_momenta(batch, shape, 1 + rank, device=dev)
        # scale so the flow moves ~4 voxels (diffeomorphic, realistic gather locality)
        m0.mul_(4.0 / metric.sharp(m0).abs().max().item())
        m_host = torch.empty(m0.shape, dtype=m0.dtype).pin_memory()
        m_host.copy_(m0)
        torch.cuda.synchronize()
        shoot = lambda: lm.expmap(metric, m0, num_steps=nsteps)
        sampler = ClockSampler(local_rank)
        sampler.start()  # runs through warm-up (same workload) and the timed region
        for _ in range(W):
            h = shoot()
        barrier()
        n0 = lm.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(K):
            h = shoot()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # nvidia-smi delivers a sample every 100 ms and the timed region of the default run is
        # shorter than that: keep the same workload running (untimed) until a few samples exist, so
        # the reported clocks / throttle reasons are always "under this load"
        t_tail = time.perf_counter()
        while len(sampler.rows) < 4 and time.perf_counter() - t_tail < 3.0 and sampler.proc is not None:
            h = shoot()
            torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["window"] = "warm-up + timed region + untimed tail of the same workload"
        launches = lm.launch_count() - n0
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        value = world * batch * V * nsteps * K / (ms * 1e-3)
        res = {"value": value, "ms_per_step": ms / K, "launches": launches, "clocks": clocks,
               "batch": batch, "shape": shape, "nsteps": nsteps, "V": V}
        # ---- per-kernel breakdown (profile mode: an event after every library launch) ----
        buf = ctypes.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        L.check(L.lib.lgm_profile_begin(L.stream_ptr(dev)))
        for _ in range(max(1, K // 2)):
            shoot()
        L.check(L.lib.lgm_profile_end(buf, len(buf)))
        res["kernels"] = json.loads(buf.value.decode())
        res["kernel_reps"] = max(1, K // 2)
        # ---- end to end through the public API with host buffers ----
        if with_e2e:
            h_host = torch.empty((batch, 3) + tuple(shape), dtype=torch.float32).pin_memory()
            def e2e_step():
                # public host-buffer API: pipelined H2D / shoot / D2H over chunks of subjects
                lm.expmap_host(metric, m_host, num_steps=nsteps, out=h_host, device=dev)
            for _ in range(min(W, 2)):
                e2e_step()
            barrier()
            e0.record()
            for _ in range(K):
                e2e_step()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["e2e"] = {"value": world * batch * V * nsteps * K / (t.item() * 1e-3), "unit": "voxel-steps/s",
                          "h2d_bytes_per_step": m_host.numel() * 4, "d2h_bytes_per_step": h_host.numel() * 4}
            # what the box's host<->device path allows: the same two buffers copied in both directions at
            # once by ALL ranks together, no compute (the ceiling of any pipelined host-buffer API)
            s2 = torch.cuda.Stream(dev)
            d_tmp = torch.empty_like(m0)
            barrier()
            e0.record()
            for _ in range(3):
                m0.copy_(m_host, non_blocking=True)
                with torch.cuda.stream(s2):
                    h_host.copy_(d_tmp, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["e2e"]["copy_only_ms_per_step"] = t.item()
            res["e2e"]["copy_gbs_per_rank_both_directions"] = 2 * m_host.numel() * 4 / (t.item() * 1e-3) / 1e9
            res["e2e"]["copy_bound_ceiling"] = world * batch * V * nsteps / (t.item() * 1e-3)
            del d_tmp
        del m0, h
        torch.cuda.empty_cache()
        return res

    # algorithmic bytes per voxel of the launch, fp32 3-D (DESIGN.md "Kernels")
    def alg_bytes(name, shape):
        zc = shape[2] // 2 + 1
        spec = 3 * 8.0 * zc / shape[2]          # one pass over the 3-channel half spectrum, per voxel
        return {"Ad_star": 36.0, "compose": 36.0, "zfwd": 12.0 + spec, "zinv": 12.0 + spec,
                "ypass": 2 * spec, "xpass": 2 * spec, "slab_fwd": 12.0 + spec, "slab_inv": 12.0 + spec}.get(name)

    main_res = measure(args.workload, args.steps, args.warmup)
    hbm, peak_src = peaks()
    batch, shape, nsteps, V = main_res["batch"], main_res["shape"], main_res["nsteps"], main_res["V"]
    ks = main_res["kernels"]
    dom = max(ks, key=lambda k: ks[k]["ms"]) if ks else None
    roofline = None
    breakdown = {}
    for k, v in ks.items():
        per_launch_ms = v["ms"] / v["launches"]
        ab = alg_bytes(k, shape)
        entry = {"launches_per_shoot": v["launches"] // main_res["kernel_reps"], "ms_per_launch": per_launch_ms,
                 "ms_per_epdiff_step": v["ms"] / (main_res["kernel_reps"] * nsteps),
                 "share": v["ms"] / sum(x["ms"] for x in ks.values())}
        if ab is not None:
            # algorithmic bytes of all launches of this kernel in the profiled shoots / their total time
            total_bytes = ab * batch * V * v["launches"]   # every launch processes the whole batch
            entry["achieved_gbs"] = total_bytes / (v["ms"] * 1e-3) / 1e9
            entry["frac"] = entry["achieved_gbs"] / hbm
            entry["alg_bytes_per_launch"] = total_bytes / v["launches"]
        breakdown[k] = entry
    if dom is not None and "achieved_gbs" in breakdown[dom]:
        traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
        tname = {"c2": "r2_traffic.json", "c3": "r2_traffic_c3.json"}.get(args.workload)
        tpath = os.path.join(ROOT, "profiles", tname) if tname else None
        if tpath and os.path.exists(tpath):
            t = json.load(open(tpath)).get(dom)
            if t:
                traffic = t["dram_read_bytes"] + t["dram_write_bytes"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": breakdown[dom]["achieved_gbs"], "peak": hbm,
                    "unit": "GB/s", "frac": breakdown[dom]["frac"], "traffic": traffic,
                    "traffic_source": "profiles/%s (ncu --set full)" % tname if traffic else None,
                    "algorithmic_bytes_per_launch": breakdown[dom].get("alg_bytes_per_launch"),
                    "peak_source": peak_src, "algorithmic_bytes_per_voxel": alg_bytes(dom, shape)}
    step_frac = main_res["value"] / world * (shoot_bytes_per_voxel(nsteps) / nsteps) / 1e9 / hbm

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    extra = {}

    def also(name, fn):  # never lose the main line
        try:
            extra[name] = fn()
        except Exception as e:
            extra[name] = {"error": ("%s: %s" % (type(e).__name__, e))[:300]}
            torch.cuda.empty_cache()

    if not args.no_extra and args.workload == "c2":
        def c3_shoot():
            r3 = measure("c3", max(2, args.steps // 2), 2, with_e2e=False)
            return {"value": r3["value"], "unit": "voxel-steps/s", "ms_per_step": r3["ms_per_step"],
                    "hbm_roofline_frac_96B": r3["value"] / world * (shoot_bytes_per_voxel(r3["nsteps"]) / r3["nsteps"]) / 1e9 / hbm,
                    "kernel_ms_per_epdiff_step": {k: v["ms"] / (r3["kernel_reps"] * r3["nsteps"]) for k, v in r3["kernels"].items()},
                    "config": {"shape": list(r3["shape"]), "batch_per_gpu": r3["batch"], "epdiff_steps": r3["nsteps"]}}
        also("c3_256", c3_shoot)
        also("c3_atlas", lambda: bench_c3_atlas(torch, dist, lm, dev, world, rank, hbm, barrier, reduce_max))
        if world == 1:
            also("c2_fwd_bwd", lambda: bench_c2_fwd_bwd(torch, lm, dev, hbm, metric, barrier, reduce_max))
            also("c4_affine", lambda: bench_c4_affine(torch, lm, dev, hbm, barrier, reduce_max))
            also("c5_deep", lambda: bench_c5_deep(torch, lm, dev, hbm, metric, barrier, reduce_max))
            also("ref_cuda", lambda: bench_ref_cuda(torch, metric, lm, dev))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, secs, cores, sample = cpu_port_throughput(shape, nsteps, reps=3, warmup=3)
        cpu = {"value": val, "unit": "voxel-steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "3D voxel-steps/sec (EPDiff shoot)", "value": main_res["value"], "unit": "voxel-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "shape": list(shape), "batch_per_gpu": batch,
                       "epdiff_steps": nsteps, "metric_params": PARAMS,
                       "l2": "inputs larger than L2 (%d MiB per field vs 126 MB)" % (batch * 3 * V * 4 >> 20),
                       "parallelism": "subjects sharded over ranks, no data-path collective",
                       "host_numa_node": numa, "host_numa_note": numa_why,
                       "momenta": "white noise low-passed by a sigma=4 Gaussian, max|sharp(m0)| = 4 voxels (BASELINE.md 4)",
                       "first_step": ("from the identity: one sharp with the -dt scaling inside, counted as 24 B/voxel "
                                      "(not 96) in hbm_roofline_frac_96B" if FIRST_STEP_SHORTCUT else "full step")},
            "hbm_roofline_frac_96B": step_frac,
            "alg_bytes_per_voxel_shoot": shoot_bytes_per_voxel(nsteps),
            "hbm_roofline_frac_nominal": main_res["value"] / world * ALG_BYTES_PER_VOXEL_STEP / 1e9 / hbm,
            "hbm_roofline_frac_note": "hbm_roofline_frac_96B counts the bytes of what ran: 96 B per voxel-step, 24 B for "
                                      "the first step from the identity; _nominal is value x 96 B / peak (every step "
                                      "counted as a full one)",
            "roofline": roofline, "kernel_breakdown": breakdown, "cpu_baseline": cpu,
            "e2e": main_res.get("e2e"), "gpu_launches": main_res["launches"], "clocks": main_res["clocks"],
        }
        if extra:
            line["also"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
