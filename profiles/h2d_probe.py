"""Concurrent pinned host->device copies at N ranks: the ceiling of the `e2e` figure per N.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/h2d_probe.py

Every rank copies its share of the config-2 matrix (819.2 MB / N, pinned) to its GPU, all ranks at once, 10 times;
prints per-rank GB/s (min / max), the aggregate, the time of the slowest rank, and each GPU's CPU/NUMA affinity."""
import os
import subprocess

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 819_200_000 // world
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
ms = []
for _ in range(10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d.copy_(h, non_blocking=True)
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
best = torch.tensor([min(ms)], dtype=torch.float64, device="cuda")
allms = [torch.zeros_like(best) for _ in range(world)]
if world > 1:
    dist.all_gather(allms, best)
else:
    allms = [best]
if rank == 0:
    t = [float(x.item()) for x in allms]
    gbs = [nbytes / (x * 1e-3) / 1e9 for x in t]
    print("ranks %d  bytes/rank %.1f MB  slowest %.3f ms  per-rank GB/s min %.1f max %.1f  aggregate %.1f GB/s"
          % (world, nbytes / 1e6, max(t), min(gbs), max(gbs), 819.2e6 / (max(t) * 1e-3) / 1e9))
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:2500])
    except Exception as e:
        print("topo unavailable", e)
if world > 1:
    dist.destroy_process_group()
