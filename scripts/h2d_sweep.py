"""Host<->device copy bandwidth per rank with all ranks copying at once (torchrun, 1/2/4/8 ranks):
what bounds the e2e (host-buffer) number at N > 1. Pinned buffers of one C2 batch (402 MB), H2D,
D2H and both directions together; with and without binding the rank to its GPU's NUMA node.
  python -m torch.distributed.run --nproc-per-node N scripts/h2d_sweep.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import bind_to_gpu_numa_node, gpu_numa_node

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 16 * 3 * 128 ** 3 * 4

def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def measure(tag):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)          # first touch on the current CPU set
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(dev)
    res = {}
    for name in ("h2d", "d2h", "both"):
        for rep in range(3):
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                if name in ("h2d", "both"):
                    d_a.copy_(h_in, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            sync()
            ms = e0.elapsed_time(e1)
        res[name] = 4 * nbytes * (2 if name == "both" else 1) / (ms * 1e-3) / 1e9
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device=dev)
    if world > 1:
        all_t = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
    else:
        all_t = [t]
    if rank == 0:
        rows = [[round(float(x), 1) for x in a] for a in all_t]
        print(json.dumps({"ranks": world, "binding": tag, "gbs_per_rank_[h2d,d2h,both]": rows,
                          "aggregate_gbs": [round(sum(r[i] for r in rows), 1) for i in range(3)]}))

node, why = gpu_numa_node(lr)
if rank == 0:
    print(json.dumps({"gpu0_numa_node": node, "why": why, "cpus": len(os.sched_getaffinity(0)),
                      "nodes": sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else None}))
measure("none")
b, why = bind_to_gpu_numa_node(lr)
measure("numa node %s" % b if b is not None else "unbound (%s)" % why)
if world > 1:
    dist.destroy_process_group()
