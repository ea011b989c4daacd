"""LDDMM vector-momentum geodesic shooting (mirror of lagomorph/lddmm.py:20-105)."""
import math
import os

import torch

from . import _lib as L
from . import adjrep, deform
from .metric import FluidMetric


def expmap_advect(metric, m, T=1.0, num_steps=10, phiinv=None):
    """EPDiff without the integrated form: Euler steps of d/dt m = -ad_v^* m (lddmm.py:20-36)."""
    if phiinv is None:
        phiinv = torch.zeros_like(m)
    dt = T / num_steps
    v = metric.sharp(m)
    phiinv = deform.compose_disp_vel(phiinv, v, dt=-dt)
    for i in range(num_steps - 1):
        m = m - dt * adjrep.ad_star(v, m)
        v = metric.sharp(m)
        phiinv = deform.compose_disp_vel(phiinv, v, dt=-dt)
    return phiinv


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class _StepWorkspace:
    """Scratch for lgm_epdiff_step_fwd, reused across the steps of one shoot."""

    def __init__(self, m0):
        self.dev = m0.device
        self.d = L.spatial_dim(m0)
        self.N = m0.shape[0]
        self.sh = L.shape_arr(m0.shape[2:])
        self.code = L.dtype_code(m0)
        self.nbytes = int(L.lib.lgm_epdiff_scratch_bytes(self.code, self.N, self.d, self.sh))
        if self.nbytes < 0:
            raise RuntimeError("lgm_epdiff_scratch_bytes rejected the arguments")
        self.buf = torch.empty(max(self.nbytes, 16), dtype=torch.uint8, device=self.dev)


def _fused_step(metric, m0, dt, phiinv, mommask, ws, out):
    alpha, beta, gamma = [float(p) for p in metric.params]
    with torch.cuda.device(ws.dev):
        L.check(L.lib.lgm_epdiff_step_fwd(ws.code, L.ptr(out), L.ptr(phiinv), L.ptr(m0), L.ptr(mommask), ws.N,
                                          ws.d, ws.sh, float(dt), alpha, beta, gamma, L.ptr(ws.buf), ws.nbytes,
                                          L.stream_ptr(ws.dev)))
    return out


def _fused_shoot(metric, m0, dt, num_steps, phiinv, mommask):
    """The whole no-grad shoot as one library call (lgm_expmap_fwd); phiinv None = zeros."""
    dev = m0.device
    d, N, sh, code = L.spatial_dim(m0), m0.shape[0], L.shape_arr(m0.shape[2:]), L.dtype_code(m0)
    alpha, beta, gamma = [float(p) for p in metric.params]
    out = torch.empty_like(m0)
    with torch.cuda.device(dev):
        nbytes = int(L.lib.lgm_expmap_scratch_bytes(code, N, d, sh))
        if nbytes < 0:
            raise RuntimeError("lgm_expmap_scratch_bytes rejected the arguments")
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        L.check(L.lib.lgm_expmap_fwd(code, L.ptr(out), L.ptr(phiinv), L.ptr(m0), L.ptr(mommask), N, d, sh, float(dt),
                                     int(num_steps), alpha, beta, gamma, L.ptr(buf), nbytes, L.stream_ptr(dev)))
    return out


def _fusable(metric, m0, phiinv, mommask):
    if not isinstance(metric, FluidMetric) or _needs_grad(m0, phiinv, mommask):
        return False
    if m0.dim() not in (4, 5) or m0.shape[1] != m0.dim() - 2:
        return False  # the unfused chain raises "vector field is of wrong dimension" (adjrep._fused)
    if not (m0.is_cuda and phiinv.is_cuda and m0.shape == phiinv.shape and m0.dtype == phiinv.dtype):
        return False
    return mommask is None or (mommask.shape == m0.shape and mommask.dtype == m0.dtype and mommask.is_cuda)


def _fused_bwd_ok(metric, m0, phiinv, mommask):
    """fp32 3-D CUDA fields with Z % 32 == 0: the step's backward runs as lgm_epdiff_step_bwd."""
    if os.environ.get("LGM_FUSED_BWD", "1") == "0" or not isinstance(metric, FluidMetric):
        return False
    if not (m0.is_cuda and phiinv.is_cuda and m0.shape == phiinv.shape and m0.dtype == phiinv.dtype):
        return False
    if m0.dtype != torch.float32 or m0.dim() != 5 or m0.shape[1] != 3 or m0.shape[0] < 1:
        return False
    if mommask is not None and not (mommask.shape == m0.shape and mommask.dtype == m0.dtype and mommask.is_cuda):
        return False
    return int(L.lib.lgm_epdiff_bwd_scratch_bytes(L.dtype_code(m0), m0.shape[0], 3, L.shape_arr(m0.shape[2:]))) >= 0


def _identity_shortcut():
    return os.environ.get("LGM_NO_FIRST_STEP_SHORTCUT") is None


def _steps_saving(metric, m0, dt, N, phiinv, mommask, from_identity=False):
    """N forward steps keeping each step's input displacement and velocity (what the backward reads).
    from_identity: phiinv is the all-zero field expmap made for phiinv=None; the first step is then
    Ad_star(0, m0) = m0, compose_disp_vel(0, v, -dt) = -dt*v: one sharp and a scaling (as lgm_expmap_fwd
    does, csrc/shoot3.cu), and its backward is a scaling and one sharp (_steps_backward)."""
    phis, vs = [], []
    with torch.no_grad():
        for n in range(N):
            if n == 0 and from_identity:
                v = metric.sharp(m0 if mommask is None else m0 * mommask)
                phis.append(phiinv)
                vs.append(phiinv)       # placeholder: the shortcut backward reads neither
                phiinv = v.mul_(-dt)
                continue
            m = adjrep.Ad_star(phiinv, m0)
            if mommask is not None:
                m = m * mommask
            v = metric.sharp(m)
            phis.append(phiinv)
            vs.append(v)
            phiinv = deform.compose_disp_vel(phiinv, v, dt=-dt)
    return phiinv, phis, vs


def _steps_backward(metric, m0, dt, phis, vs, mommask, gradout, need_m0, need_phi, from_identity=False):
    """Backward of the steps recorded by _steps_saving, last step first: one lgm_epdiff_step_bwd per
    step; dL/dm0 accumulates in place over the steps, dL/dphiinv is carried in `g`."""
    dev = m0.device
    code, N = L.dtype_code(m0), m0.shape[0]
    sh = L.shape_arr(m0.shape[2:])
    nbytes = int(L.lib.lgm_epdiff_bwd_scratch_bytes(code, N, 3, sh))
    scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    g = gradout.contiguous().clone()
    d_m0 = torch.zeros_like(m0) if need_m0 else None
    acc = torch.zeros_like(m0) if (need_phi or len(phis) > 1) else None
    alpha, beta, gamma = [float(p) for p in metric.params]
    with torch.cuda.device(dev):
        for k in reversed(range(len(phis))):
            want_phi = need_phi or k > 0
            if not (want_phi or need_m0):
                break
            if k == 0 and from_identity and not need_phi:
                # phi_1 = -dt * sharp(m0 * mask): d_m0 += mask * sharp(-dt * g) (sharp is self-adjoint)
                with torch.no_grad():
                    w = metric.sharp(g.mul_(-dt))
                    if mommask is not None:
                        w.mul_(mommask)
                    d_m0.add_(w)
                break
            L.check(L.lib.lgm_epdiff_step_bwd(code, L.ptr(g), L.ptr(d_m0), L.ptr(acc), L.ptr(phis[k]), L.ptr(vs[k]),
                                              L.ptr(m0), L.ptr(mommask), N, 3, sh, float(dt), alpha, beta, gamma,
                                              L.ptr(scratch), nbytes, int(want_phi), int(need_m0),
                                              L.stream_ptr(dev)))
    return d_m0, (g if need_phi else None)


class EPDiffShootFunction(torch.autograd.Function):
    """N EPDiff steps with a hand-written backward (fp32 3-D): forward keeps (phiinv_k, v_k) per step
    -- or, with save=False, only the inputs and replays the block in backward (activation
    checkpointing, the working form of lddmm.py:47-70)."""

    @staticmethod
    def forward(ctx, metric, m0, dt, N, phiinv, mommask, save, from_identity=False):
        m0c, pc = L.aligned(m0.detach()), L.aligned(phiinv.detach())
        mk = None if mommask is None else mommask.detach().contiguous()
        ctx.metric, ctx.dt, ctx.N, ctx.save, ctx.mk = metric, dt, N, save, mk
        ctx.from_identity = bool(from_identity) and _identity_shortcut()
        if save:
            out, phis, vs = _steps_saving(metric, m0c, dt, N, pc, mk, ctx.from_identity)
            ctx.nsaved = len(phis)
            ctx.save_for_backward(m0c, *phis, *vs)
        else:
            out = pc
            with torch.no_grad():
                for n in range(N):
                    out = EPDiff_step(metric, m0c, dt, out, mommask=mk)
            ctx.save_for_backward(m0c, pc)
        return out

    @staticmethod
    def backward(ctx, gradout):
        need_m0, need_phi = ctx.needs_input_grad[1], ctx.needs_input_grad[4]
        if ctx.save:
            m0c = ctx.saved_tensors[0]
            phis = list(ctx.saved_tensors[1:1 + ctx.nsaved])
            vs = list(ctx.saved_tensors[1 + ctx.nsaved:])
        else:
            m0c, pc = ctx.saved_tensors
            _, phis, vs = _steps_saving(ctx.metric, m0c, ctx.dt, ctx.N, pc, ctx.mk, ctx.from_identity)
        d_m0, d_phi = _steps_backward(ctx.metric, m0c, ctx.dt, phis, vs, ctx.mk, gradout, need_m0, need_phi,
                                      ctx.from_identity)
        return None, d_m0, None, None, d_phi, None, None, None


def EPDiff_step(metric, m0, dt, phiinv, mommask=None):
    """phiinv <- -dt*v + phiinv(x - dt*v), v = sharp(Ad_star(phiinv, m0)) (lddmm.py:39-44).

    Without autograd this is one library call (three fused stages); with autograd it is
    the same three stages as differentiable Functions."""
    if _fusable(metric, m0, phiinv, mommask):
        m0c, pc = L.aligned(m0), L.aligned(phiinv)
        mk = None if mommask is None else mommask.contiguous()
        return _fused_step(metric, m0c, dt, pc, mk, _StepWorkspace(m0c), torch.empty_like(pc))
    if _needs_grad(m0, phiinv) and not _needs_grad(mommask) and _fused_bwd_ok(metric, m0, phiinv, mommask):
        return EPDiffShootFunction.apply(metric, m0, dt, 1, phiinv, mommask, True)
    m = adjrep.Ad_star(phiinv, m0)
    if mommask is not None:
        m = m * mommask
    v = metric.sharp(m)
    return deform.compose_disp_vel(phiinv, v, dt=-dt)


EPDiffStep = EPDiff_step  # alias used by the project brief


class EPDiffStepsFunction(torch.autograd.Function):
    """N EPDiff steps with activation recomputation: only (m0, phiinv) are kept and the block is
    replayed in backward. (The reference's version, lddmm.py:47-70, has swapped arguments and is
    unreachable; this is the working equivalent.)"""

    @staticmethod
    def forward(ctx, metric, m0, dt, N, phiinv, mommask):
        ctx.metric, ctx.dt, ctx.N, ctx.mommask = metric, dt, N, mommask
        ctx.save_for_backward(m0, phiinv)
        with torch.no_grad():
            for n in range(N):
                phiinv = EPDiff_step(metric, m0, dt, phiinv, mommask=mommask)
        return phiinv

    @staticmethod
    def backward(ctx, gradout):
        m0, phiinv = ctx.saved_tensors
        with torch.enable_grad():
            m0_ = m0.detach().requires_grad_(ctx.needs_input_grad[1])
            p_ = phiinv.detach().requires_grad_(ctx.needs_input_grad[4])
            p = p_
            for n in range(ctx.N):
                p = EPDiff_step(ctx.metric, m0_, ctx.dt, p, mommask=ctx.mommask)
            inputs = [t for t in (m0_, p_) if t.requires_grad]
            grads = list(torch.autograd.grad(p, inputs, gradout)) if inputs else []
        g_m0 = grads.pop(0) if m0_.requires_grad else None
        g_p = grads.pop(0) if p_.requires_grad else None
        return None, g_m0, None, None, g_p, None


def EPDiff_steps(metric, m0, dt, N, phiinv, mommask=None):
    if not _needs_grad(mommask) and _fused_bwd_ok(metric, m0, phiinv, mommask):
        return EPDiffShootFunction.apply(metric, m0, dt, N, phiinv, mommask, False)
    return EPDiffStepsFunction.apply(metric, m0, dt, N, phiinv, mommask)


def expmap(metric, m0, T=1.0, num_steps=10, phiinv=None, mommask=None, checkpoints=False):
    """Exponential map of an initial momentum; returns the displacement of phi^{-1}
    (reference: lddmm.py:73-105; the non-checkpointed branch :87-91 is the parity target).

    checkpoints: False/None -> plain loop; int k -> recompute in blocks of k steps;
    True -> blocks of ~sqrt(num_steps) steps (num_steps unchanged, last block shorter)."""
    dt = T / num_steps
    if checkpoints is None or checkpoints is False:
        if num_steps > 0 and _fusable(metric, m0, m0 if phiinv is None else phiinv, mommask):
            mk = None if mommask is None else mommask.contiguous()
            return _fused_shoot(metric, L.aligned(m0), dt, num_steps, None if phiinv is None else L.aligned(phiinv), mk)
    from_identity = phiinv is None
    if phiinv is None:
        phiinv = torch.zeros_like(m0)
    if checkpoints is None or checkpoints is False:
        if (num_steps > 0 and _needs_grad(m0, phiinv) and not _needs_grad(mommask)
                and _fused_bwd_ok(metric, m0, phiinv, mommask)):
            return EPDiffShootFunction.apply(metric, m0, dt, num_steps, phiinv, mommask, True, from_identity)
        for i in range(num_steps):
            phiinv = EPDiff_step(metric, m0, dt, phiinv, mommask=mommask)
        return phiinv
    cps = int(math.sqrt(num_steps)) if checkpoints is True else int(checkpoints)
    cps = max(1, cps)
    done = 0
    while done < num_steps:
        k = min(cps, num_steps - done)
        if _needs_grad(m0, phiinv):
            phiinv = EPDiff_steps(metric, m0, dt, k, phiinv, mommask)
        else:
            for i in range(k):
                phiinv = EPDiff_step(metric, m0, dt, phiinv, mommask=mommask)
        done += k
    return phiinv


def _auto_chunks(N, num_steps=10, cap=5, ratio=None):
    """Chunk sizes for expmap_host. Every chunk costs a fixed ~0.25 ms of launch tails, so chunks
    should be few and large; but the first host->device copy and the last device->host copy are
    exposed, and a chunk's copy has to hide behind its neighbour's compute. With ratio r = copy time /
    compute time per subject the sizes may grow by 1/r per chunk from 1 at the head, shrink to 1 at
    the tail, and are capped in the middle: r = 0.48 (N = 16, 10 steps on a PCIe 5 x16 B200) ->
    [1, 2, 4, 5, 3, 1]; r >= 1 (copy as slow as compute) -> single subjects.
    expmap_host MEASURES r on this box (_copy_compute_ratio); ratio=None falls back to the figure of
    the development box (12 B per voxel over ~55 GB/s against ~22 G voxel-steps/s: 4.8 / num_steps)."""
    if ratio is None:
        ratio = 4.8 / num_steps
    g = max(1.0, 1.0 / max(ratio, 1e-3))
    head, tail = [], []
    h, t, left = 1.0, 1.0, N
    while left > 0:
        k = min(int(h), cap, left)
        head.append(k)
        left -= k
        h *= g
        if left > 0:
            k = min(int(t), cap, left)
            tail.insert(0, k)
            left -= k
            t *= 1.5 * g if g >= 1.5 else g
    sizes = head + tail
    i = 1
    while i < len(sizes) - 1:           # no stray single subject between larger chunks
        if sizes[i] == 1 and sizes[i - 1] > 1 and sizes[i + 1] > 1:
            j = i - 1 if sizes[i - 1] <= sizes[i + 1] else i + 1
            sizes[j] += 1
            del sizes[i]
        else:
            i += 1
    return sizes


def _ramp_chunks(N):
    """1, 2, 3, ... up and down again: [1, 2, 3, 4, 3, 2, 1] for N = 16. Small chunks at both ends (their
    copies are exposed), every copy hidden behind a neighbour's shoot as long as a subject's copy is
    shorter than its shoot; the remainder that does not fit the triangle widens the middle."""
    k = 1
    while (k + 1) * (k + 1) <= N:
        k += 1
    sizes = list(range(1, k + 1)) + list(range(k - 1, 0, -1))   # sums to k*k
    left = N - k * k
    mid = len(sizes) // 2
    order = sorted(range(len(sizes)), key=lambda i: (abs(i - mid), i))   # the rest widens the middle first
    for t in range(left):
        sizes[order[t % len(order)]] += 1
    return [c for c in sizes if c > 0]


def _pair_chunks(N):
    """1, 1, 2, 2, ..., 2, 1, 1: pairs in the middle, single subjects at both ends (whose copies are exposed)"""
    if N < 4:
        return [1] * N
    return [1, 1] + [2] * ((N - 4) // 2) + [1] * ((N - 4) % 2) + [1, 1]


_BEST_CHUNKS = {}   # (device, subject shape, dtype, steps, N, streams) -> (chunk schedule, compute streams) measured best


def _measured_chunks(metric, m0_host, T, num_steps, out, dev, streams=None):
    """chunk="auto": candidate (schedule, compute streams) pairs are each run once on the real data the first
    time a (device, shape, steps, batch) combination is seen, and the fastest is kept: the schedule from the
    copy/compute model (_auto_chunks) and the ramp (_ramp_chunks) on one compute stream, and -- when the
    caller leaves `streams` open -- the ramp, pairs and single subjects on two. (The model assumes a shoot's
    time is proportional to its batch; small chunks are slower than that on ONE stream because none of their
    ~50 launches fills the GPU, but two of them in flight do: C2 on a B200, profiles/r2_e2e_streams.log:
    ramp x 1 stream 14.07 ms, single subjects x 1 stream 16.1 ms, single subjects x 2 streams 12.6 ms.)"""
    N = m0_host.shape[0]
    key = (dev.index, tuple(m0_host.shape[1:]), m0_host.dtype, int(num_steps), N, streams)
    if key in _BEST_CHUNKS:
        return _BEST_CHUNKS[key]
    one = 1 if streams is None else max(1, int(streams))
    model = _auto_chunks(N, num_steps, ratio=_copy_compute_ratio(metric, m0_host, T, num_steps, dev))
    cands = [(model, one)]
    ramp = _ramp_chunks(N)
    if ramp != model and N >= 4:
        cands.append((ramp, one))
    if N >= 4 and (streams is None or one > 1):
        two = 2 if streams is None else one
        for sizes in (ramp, _pair_chunks(N), [1] * N):
            if (sizes, two) not in cands:
                cands.append((sizes, two))
    if len(cands) == 1 or torch.cuda.is_current_stream_capturing():
        return cands[0]
    best, best_ms = cands[0], None
    for sizes, ns in cands:
        expmap_host(metric, m0_host, T=T, num_steps=num_steps, out=out, device=dev, chunk=sizes, streams=ns)   # plan, warm-up
        ms = None
        for _ in range(2):      # best of two: one run is ~1.3 x the shoot, candidates differ by a few per cent
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            expmap_host(metric, m0_host, T=T, num_steps=num_steps, out=out, device=dev, chunk=sizes, streams=ns)
            e1.record()
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1)
            ms = t if ms is None else min(ms, t)
        if best_ms is None or ms < best_ms:
            best, best_ms = (sizes, ns), ms
    mine = (dev.index, tuple(m0_host.shape[1:]), m0_host.dtype, int(num_steps))
    for k in [k for k in _HOST_PLANS if k[:4] == mine and (k[-2] != tuple(best[0]) or k[-1] != best[1] + 2)]:
        _HOST_PLANS.pop(k)            # the losing candidates' graphs and buffers
    _BEST_CHUNKS[key] = best
    return best


def expmap_host(metric, m0_host, T=1.0, num_steps=10, out=None, device=None, chunk="auto", graphs=True, streams=None):
    """Shoot momenta that live in (pinned) HOST memory and return the deformations in host memory.

    Subjects are independent, so the batch is cut into chunks that flow through a three-stage
    pipeline on separate CUDA streams: host->device copy of chunk i+1, EPDiff shoot of chunk i,
    device->host copy of chunk i-1 all overlap (PCIe is full duplex). With graphs=True (default) every
    chunk's shoot is replayed as a cached CUDA graph (_host_plan). streams = number of compute streams:
    consecutive chunks' graphs alternate between them, so that two small shoots are in flight at once
    (graphs only: each graph owns its scratch memory). chunk="auto" with streams=None measures a few
    (schedule, streams) candidates on first use and keeps the fastest (_measured_chunks).
    Same result as `expmap(metric, m0_host.cuda(), ...).cpu()`. No autograd.
    """
    if m0_host.is_cuda:
        raise RuntimeError("expmap_host takes host tensors; use expmap for device tensors")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    N = m0_host.shape[0]
    if out is None:
        out = torch.empty_like(m0_host, pin_memory=True)
    if N == 0:
        return out
    # chunk: "auto", an int (uniform chunks) or a list of chunk sizes.
    if isinstance(chunk, str):
        sizes, auto_ns = _measured_chunks(metric, m0_host, T, num_steps, out, dev, streams)
        if streams is None:
            streams = auto_ns
    elif isinstance(chunk, (list, tuple)):
        sizes = [int(c) for c in chunk if int(c) > 0]
        assert sum(sizes) == N, "chunk sizes must add up to the batch"
    else:
        # uniform chunks; the first and last are halved (down to 1 subject) because their copy
        # cannot hide behind any compute
        chunk = max(1, min(int(chunk), N))
        edge = max(1, chunk // 2)
        sizes = []
        left = N
        if N > 2 * edge:
            sizes.append(edge)
            left -= 2 * edge
            while left > 0:
                sizes.append(min(chunk, left))
                left -= sizes[-1]
            sizes.append(edge)
        else:
            while left > 0:
                sizes.append(min(chunk, left))
                left -= sizes[-1]
    maxc = max(sizes)
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    # streams: how many chunk shoots may be in flight at once. A chunk of one to four subjects does not
    # fill the GPU through the ramp and the tail of each of its ~50 launches; with two compute streams the
    # neighbouring chunk's kernels fill those gaps. Only with graphs (each graph owns its scratch memory).
    ns = (1 if streams is None else max(1, int(streams))) if graphs else 1
    nbuf = ns + 2
    plan = _host_plan(metric, m0_host, T, num_steps, dev, sizes, nbuf) if graphs else None
    if plan is not None:
        dbuf = plan["dbuf"]
    else:
        ns = 1
        dbuf = [torch.empty((maxc,) + tuple(m0_host.shape[1:]), dtype=m0_host.dtype, device=dev) for _ in range(nbuf)]
    s_comp = [cur] if ns == 1 else [torch.cuda.Stream(dev) for _ in range(ns)]
    for sc in s_comp:
        if sc is not cur:
            sc.wait_stream(cur)
    in_done = [torch.cuda.Event() for _ in range(nbuf)]
    comp_done = [None] * nbuf  # input buffer of slot b has been consumed
    starts = [sum(sizes[:i]) for i in range(len(sizes))]

    def issue_h2d(ci):
        b = ci % nbuf
        n = sizes[ci]
        with torch.cuda.stream(s_in):
            if comp_done[b] is not None:
                s_in.wait_event(comp_done[b])
            dbuf[b][:n].copy_(m0_host[starts[ci]:starts[ci] + n], non_blocking=True)
            in_done[b].record(s_in)

    depth = nbuf - 1          # copies issued ahead of the chunk being shot
    for ci in range(min(depth, len(starts))):
        issue_h2d(ci)
    for ci, st in enumerate(starts):
        b = ci % nbuf
        n = sizes[ci]
        sc = s_comp[ci % ns]
        with torch.cuda.stream(sc):
            sc.wait_event(in_done[b])
            if plan is not None:
                plan["graphs"][ci].replay()      # the chunk's num_steps x 5 launches as one CUDA graph
                h = plan["outs"][ci]
            else:
                with torch.no_grad():
                    h = expmap(metric, dbuf[b][:n], T=T, num_steps=num_steps)
            ev = torch.cuda.Event()
            ev.record(sc)
        comp_done[b] = ev
        if ci + depth < len(starts):
            issue_h2d(ci + depth)     # its buffer was chunk ci - 1's: that shoot is already in its stream
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev)
            out[st:st + n].copy_(h, non_blocking=True)
            if plan is None:
                h.record_stream(s_out)
    for sc in s_comp:
        if sc is not cur:
            cur.wait_stream(sc)
    cur.wait_stream(s_out)
    cur.wait_stream(s_in)
    return out


_RATIOS = {}   # (device, subject shape, dtype, steps) -> measured copy / compute time ratio


def _copy_compute_ratio(metric, m0_host, T, num_steps, dev):
    """Host->device copy time of ONE subject over the time of its shoot, measured with CUDA events
    the first time a (device, shape, dtype, steps) combination is seen; None if it cannot be
    measured (stream capture in progress)."""
    key = (dev.index, tuple(m0_host.shape[1:]), m0_host.dtype, int(num_steps))
    if key in _RATIOS:
        return _RATIOS[key]
    if torch.cuda.is_current_stream_capturing() or m0_host.shape[0] == 0:
        return None
    with torch.cuda.device(dev), torch.no_grad():
        one = m0_host[:1]
        d = one.to(dev, non_blocking=True)
        expmap(metric, d, T=T, num_steps=num_steps)            # warm-up: tables, allocator
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize(dev)
        ev[0].record()
        d.copy_(one, non_blocking=True)
        ev[1].record()
        expmap(metric, d, T=T, num_steps=num_steps)
        ev[2].record()
        torch.cuda.synchronize(dev)
        copy_ms, shoot_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    _RATIOS[key] = copy_ms / max(shoot_ms, 1e-6)
    return _RATIOS[key]


_HOST_PLANS = {}   # key -> plan; at most _HOST_PLANS_MAX entries (each holds its chunks' device buffers)
_HOST_PLANS_MAX = 2


def _host_plan(metric, m0_host, T, num_steps, dev, sizes, nbuf):
    """CUDA graphs of the chunk shoots of expmap_host, cached per (device, shape, dtype, steps, metric,
    chunk schedule). A chunk of one or two subjects is fifty short launches: replayed as a graph they
    run back to back without launch gaps and without the Python / ctypes work per step. The plan owns
    the staging buffers the graphs read and the deformation buffers they write. Returns None (eager
    path) when capture is not possible."""
    if os.environ.get("LGM_HOST_GRAPHS", "1") == "0" or not isinstance(metric, FluidMetric):
        return None
    if torch.cuda.is_current_stream_capturing():
        return None
    key = (dev.index, tuple(m0_host.shape[1:]), m0_host.dtype, int(num_steps), float(T),
           tuple(float(p) for p in metric.params), tuple(sizes), nbuf)
    plan = _HOST_PLANS.get(key)
    if plan is not None:
        return plan
    try:
        maxc = max(sizes)
        dbuf = [torch.zeros((maxc,) + tuple(m0_host.shape[1:]), dtype=m0_host.dtype, device=dev) for _ in range(nbuf)]
        with torch.no_grad():
            expmap(metric, dbuf[0][:1], T=T, num_steps=1)   # builds the FFT tables outside any capture
        torch.cuda.synchronize(dev)
        graphs, outs = [], []
        for ci, n in enumerate(sizes):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                h = expmap(metric, dbuf[ci % nbuf][:n], T=T, num_steps=num_steps)
            graphs.append(g)
            outs.append(h)
        torch.cuda.synchronize(dev)
        plan = {"dbuf": dbuf, "graphs": graphs, "outs": outs}
    except Exception:
        torch.cuda.synchronize(dev)
        return None
    while len(_HOST_PLANS) >= _HOST_PLANS_MAX:
        _HOST_PLANS.pop(next(iter(_HOST_PLANS)))
    _HOST_PLANS[key] = plan
    return plan
