#!/usr/bin/env python
"""Secondary bench: BASELINE config 3 (3-D LDDMM atlas building, subjects sharded over GPUs,
NCCL all_reduce of the atlas gradient). One "step" = one epoch over this rank's subjects:
5-step expmap + deform + loss + backward + momentum update per batch, then the image update.
Prints one JSON line (rank 0): subjects/s and voxel-steps/s (fwd+bwd) aggregate over ranks.

  python bench_atlas.py [--size 128] [--subjects-per-gpu 8] [--batch 4] [--steps 2] [--warmup 1]
  torchrun --nproc-per-node N ... bench_atlas.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--subjects-per-gpu", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import lagomorph_b200 as lm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = a.subjects_per_gpu * world
    n = a.size

    class Synth:  # template blob + per-subject shift, generated on the fly (deterministic)
        def __len__(self):
            return S

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(100 + i)
            ax = torch.arange(n, dtype=torch.float32)
            c = n / 2 + (torch.rand(3, generator=g) - 0.5) * n / 8
            e = [torch.exp(-((ax - c[d]) ** 2) / (2 * (n / 6) ** 2)) for d in range(3)]
            return (e[0][:, None, None] * e[1][None, :, None] * e[2][None, None, :]).unsqueeze(0)

    b = lm.LDDMMAtlasBuilder(Synth(), num_epochs=1, batch_size=a.batch, lddmm_integration_steps=5,
                             reg_weight=1e-2, learning_rate_pose=1.0, learning_rate_image=0.1,
                             device=dev, world_size=world, rank=rank)
    b.initialize()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        b.epoch()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lm.launch_count()
    e0.record()
    for _ in range(a.steps):
        loss, reg = b.epoch()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / a.steps
    if rank == 0:
        print(json.dumps({
            "metric": "atlas epoch: subjects/s (5-step expmap fwd+bwd, image + momentum update)",
            "value": S / (ms * 1e-3), "unit": "subjects/s", "voxel_steps_per_s_fwd_bwd": S * n ** 3 * 5 / (ms * 1e-3),
            "n_gpus": world, "ms_per_epoch": ms, "scaling": "weak",
            "config": {"workload": "c3-like", "shape": [n, n, n], "subjects_per_gpu": a.subjects_per_gpu,
                       "batch": a.batch, "epdiff_steps": 5}, "gpu_launches": lm.launch_count() - n0,
            "last_epoch_loss": loss}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
