# all-reduce latency of the MBD result (int64[100000] = 800 KB) over the GPUs of one box:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/allreduce_probe.py
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(100_000, dtype=torch.int64, device="cuda")
for _ in range(20):
    dist.all_reduce(x)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if dist.get_rank() == 0:
    print("all_reduce int64[100000], %d GPUs, NCCL_ALGO=%s NCCL_PROTO=%s: %.1f us" % (
        dist.get_world_size(), os.environ.get("NCCL_ALGO", "default"), os.environ.get("NCCL_PROTO", "default"), t.item()))
dist.destroy_process_group()
