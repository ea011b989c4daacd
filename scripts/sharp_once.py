"""A few sharp() calls at one size, for ncu captures. usage: python scripts/sharp_once.py N n"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
N, n = int(sys.argv[1]), int(sys.argv[2])
m = torch.randn(N, 3, n, n, n, device="cuda")
met = lm.FluidMetric([0.1, 0.0, 0.01])
for _ in range(3):
    v = met.sharp(m)
torch.cuda.synchronize()
