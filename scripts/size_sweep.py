"""compose / Ad_star / torch add at several sizes: is the power-of-two plane stride hurting DRAM?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
dev = torch.device("cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for shape in [(128,128,128),(120,120,128),(128,128,160),(136,136,128),(128,129,128),(96,160,128)]:
    N = 16
    a = torch.randn(N, 3, *shape, device=dev) * 0.5
    b = torch.randn(N, 3, *shape, device=dev)
    out = torch.empty_like(a)
    V = N * shape[0]*shape[1]*shape[2]
    ms0 = t(lambda: torch.add(a, b, out=out))
    ms1 = t(lambda: lm.compose(a, b, -0.1, 1.0))
    ms2 = t(lambda: lm.Ad_star(a, b))
    print("%-16s add %.0f GB/s | compose %.3f ms %.0f GB/s | Ad_star %.3f ms %.0f GB/s" % (shape, 36*V/ms0/1e6, ms1, 36*V/ms1/1e6, ms2, 36*V/ms2/1e6))
