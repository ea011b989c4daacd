"""torchrun --nproc-per-node 2 scripts/affine_atlas_2rank_check.py: the sharded affine atlas (2 ranks x 4
subjects, one batch each) must reproduce the single-process run (8 subjects, two batches of 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lagomorph_b200 as lm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(5)
S, shape = 8, (16, 16, 32)
ax = [torch.arange(n, dtype=torch.float32) for n in shape]
sh = (torch.rand(S, 3, generator=g) - 0.5) * 4
data = torch.stack([(torch.exp(-((ax[0] - 7.5 - s[0]) ** 2) / 18)[:, None, None] * torch.exp(-((ax[1] - 7.5 - s[1]) ** 2) / 18)[None, :, None]
                     * torch.exp(-((ax[2] - 15.5 - s[2]) ** 2) / 18)[None, None, :]).unsqueeze(0) for s in sh])
As0, Ts0 = 0.01 * torch.randn(S, 3, 3, generator=g), 0.3 * torch.randn(S, 3, generator=g)
kw = dict(num_epochs=3, batch_size=4, learning_rate_A=0.05, learning_rate_T=5.0, learning_rate_I=1.0, reg_weightA=0.1)
# DistributedSampler order puts subjects 0,2,4,6 on rank 0: reorder the single-process run the same way so
# that its two batches are exactly the two ranks' batches
perm = [0, 2, 4, 6, 1, 3, 5, 7]
I2, A2, T2, el2, _ = lm.affine_atlas(data, As0.clone(), Ts0.clone(), world_size=world, rank=rank, device=dev, **kw)
if rank == 0:
    I1, A1, T1, el1, _ = lm.affine_atlas(data[perm], As0[perm].clone(), Ts0[perm].clone(), device=dev, **kw)
    inv = torch.argsort(torch.tensor(perm))
    A1, T1 = A1[inv], T1[inv]
    ok = (torch.allclose(I1, I2, rtol=1e-4, atol=1e-6) and torch.allclose(A1, A2.cpu(), rtol=1e-4, atol=1e-6)
          and torch.allclose(T1, T2.cpu(), rtol=1e-4, atol=1e-6) and all(abs(a - b) <= 1e-4 * abs(a) for a, b in zip(el1, el2)))
    print("affine_atlas 2-rank check:", "OK" if ok else "MISMATCH", el1, el2, (A1 - A2.cpu()).abs().max().item(), (T1 - T2.cpu()).abs().max().item())
dist.barrier()
dist.destroy_process_group()
