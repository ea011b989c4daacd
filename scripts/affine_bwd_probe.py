"""affine_interp backward at config-4 size, split by requested gradients (which part costs what)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from lagomorph_b200.affine import affine_interp_backward, affine_interp_forward
dev = "cuda"
N, n = 16, 192
g = torch.Generator(device=dev).manual_seed(1)
I = torch.rand((N, 1, n, n, n), device=dev, generator=g)
go = torch.randn((N, 1, n, n, n), device=dev, generator=g)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, A, T in (("identity", torch.eye(3, device=dev)[None].repeat(N, 1, 1), torch.zeros(N, 3, device=dev)),
                   ("shift 0.3", torch.eye(3, device=dev)[None].repeat(N, 1, 1), torch.full((N, 3), 0.3, device=dev)),
                   ("c4: I + 0.05 randn, 2 randn", torch.eye(3, device=dev)[None] + 0.05 * torch.randn((N, 3, 3), device=dev, generator=g), 2 * torch.randn((N, 3), device=dev, generator=g))):
    A, T = A.contiguous(), T.contiguous()
    print(name, "fwd %.3f" % t(lambda: affine_interp_forward(I, A, T)),
          "bwd all %.3f" % t(lambda: affine_interp_backward(go, I, A, T, True, True, True)),
          "d_I only %.3f" % t(lambda: affine_interp_backward(go, I, A, T, True, False, False)),
          "d_A,d_T only %.3f" % t(lambda: affine_interp_backward(go, I, A, T, False, True, True)),
          "memset %.3f" % t(lambda: torch.zeros_like(I)),
          "interp_adjoint (splat3) %.3f" % t(lambda: lm.interp_adjoint(go, torch.zeros((N, 3, n, n, n), device=dev))))
