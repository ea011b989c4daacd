#!/usr/bin/env python
"""Secondary bench: BASELINE config 4 (3-D 192^3 affine atlas building, batch 32 per GPU). One "step" =
one epoch of lagomorph_b200.affine_atlas over this rank's subjects: affine_interp forward, MSE,
affine_interp backward (d_I splat + d_A / d_T reductions), parameter and atlas updates, NCCL
all_reduce of the atlas gradient. Prints one JSON line (rank 0).

  python bench_affine_atlas.py [--size 192] [--subjects-per-gpu 32] [--batch 32] [--steps 3] [--warmup 1]
  torchrun --nproc-per-node N ... bench_affine_atlas.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--subjects-per-gpu", type=int, default=32)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
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
    n, S = a.size, a.subjects_per_gpu * world
    g = torch.Generator(device=dev).manual_seed(1)
    # synthetic subjects: one blob, per-subject shift (generated on the device; every rank builds all
    # S small descriptors but only materialises its own shard inside affine_atlas via indexing)
    ax = torch.arange(n, dtype=torch.float32, device=dev)

    class Synth:
        def __len__(self):
            return S

        def __getitem__(self, i):
            gi = torch.Generator().manual_seed(100 + int(i))
            c = (n - 1) / 2 + (torch.rand(3, generator=gi) - 0.5) * n / 8
            e = [torch.exp(-((ax - float(c[d])) ** 2) / (2 * (n / 6) ** 2)) for d in range(3)]
            return (e[0][:, None, None] * e[1][None, :, None] * e[2][None, None, :]).unsqueeze(0)

    As = torch.zeros(S, 3, 3)
    Ts = torch.zeros(S, 3)
    kw = dict(batch_size=a.batch, learning_rate_A=1e-3, learning_rate_T=1.0, learning_rate_I=1.0,
              world_size=world, rank=rank, device=dev)
    data = Synth()
    n0 = lm.launch_count()
    for _ in range(a.warmup):
        I, As, Ts, el, _ = lm.affine_atlas(data, As, Ts, num_epochs=1, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    I, As, Ts, el, _ = lm.affine_atlas(data, As, Ts, I=I, num_epochs=a.steps, **kw)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / a.steps
    if rank == 0:
        print(json.dumps({"metric": "affine atlas epoch: subjects/s (affine_interp fwd + bwd, pose and image updates)",
                          "value": S / (ms * 1e-3), "unit": "subjects/s", "n_gpus": world, "ms_per_epoch": ms,
                          "note": "epoch time includes staging this rank's shard on the device (synthetic generator)",
                          "scaling": "weak", "config": {"workload": "c4", "shape": [n, n, n], "subjects_per_gpu": a.subjects_per_gpu,
                                                        "batch": a.batch},
                          "gpu_launches": lm.launch_count() - n0, "epoch_losses": el}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
