"""expmap_host: chunk schedules x compute streams (1, 2, 3), C2 (16 x 128^3, 10 steps) and 8 x 256^3, 5 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
dev = torch.device("cuda")
metric = lm.FluidMetric([0.1, 0.0, 0.01])
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def run(N, side, nsteps, scheds):
    shape = (side,) * 3
    V = side ** 3
    g = torch.Generator().manual_seed(1)
    m_host = torch.randn((N, 3) + shape, generator=g).pin_memory()
    m0 = m_host.to(dev)
    s = 4.0 / metric.sharp(m0).abs().max().item()
    m0.mul_(s); m_host.mul_(s)
    ref = lm.expmap(metric, m0, num_steps=nsteps).cpu()
    ms = t(lambda: lm.expmap(metric, m0, num_steps=nsteps))
    print("%d x %d^3: device-resident shoot %.2f ms = %.2f G" % (N, side, ms, N * V * nsteps / ms / 1e6), flush=True)
    out = torch.empty_like(m_host).pin_memory()
    cfgs = [(c, ns) for c in scheds for ns in (1, 2, 3)]
    res = {i: [] for i in range(len(cfgs))}
    for rep in range(3):
        for i, (chunk, ns) in enumerate(cfgs):
            res[i].append(t(lambda: lm.expmap_host(metric, m_host, num_steps=nsteps, out=out, device=dev, chunk=chunk, streams=ns), reps=4))
            if rep == 0:
                assert torch.equal(out, ref), "result differs: %s %d" % (chunk, ns)
    for i, (chunk, ns) in enumerate(cfgs):
        ms = sorted(res[i])[1]
        print("  chunk=%-28s streams=%d: %s ms  median %.2f G" % (chunk, ns, ["%.2f" % x for x in res[i]], N * V * nsteps / ms / 1e6), flush=True)
run(16, 128, 10, [[1, 2, 3, 4, 3, 2, 1], [1, 2, 4, 5, 3, 1], [1, 3, 4, 4, 3, 1], [2, 4, 4, 4, 2], [1, 2, 2, 2, 2, 2, 2, 2, 1], [1, 1, 2, 2, 2, 2, 2, 2, 1, 1], [1] * 16, [2] * 8, [4] * 4])
run(8, 256, 5, [[1, 2, 2, 2, 1], [1] * 8, [1, 1, 2, 2, 1, 1], [2, 2, 2, 2]])

# device-resident: the batch as k concurrent graph replays on k streams (do tails / ramps of one group's
# kernels fill with the other group's work?)
def split_run(N, side, nsteps, ks):
    shape = (side,) * 3
    V = side ** 3
    g = torch.Generator().manual_seed(1)
    m0 = torch.randn((N, 3) + shape, generator=g).to(dev)
    m0.mul_(4.0 / metric.sharp(m0).abs().max().item())
    with torch.no_grad():
        lm.expmap(metric, m0, num_steps=nsteps)
    for k in ks:
        parts = list(m0.chunk(k))
        graphs, outs = [], []
        for p in parts:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr), torch.no_grad():
                outs.append(lm.expmap(metric, p, num_steps=nsteps))
            graphs.append(gr)
        streams = [torch.cuda.Stream(dev) for _ in range(k)]
        cur = torch.cuda.current_stream(dev)
        def go():
            for s_, g_ in zip(streams, graphs):
                s_.wait_stream(cur)
                with torch.cuda.stream(s_):
                    g_.replay()
            for s_ in streams:
                cur.wait_stream(s_)
        ms = t(go, reps=10)
        print("%d x %d^3 as %d concurrent graphs: %.3f ms = %.2f G" % (N, side, k, ms, N * V * nsteps / ms / 1e6), flush=True)
split_run(16, 128, 10, [1, 2, 4])
split_run(8, 256, 5, [1, 2, 4])
