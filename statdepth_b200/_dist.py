"""Multi-GPU sharding of one depth call: one process per GPU, torch.distributed for the plumbing.

* relaxed band depth is additive over time rows  -> every rank ranks a contiguous block of rows
  and the int64 numerators are summed with ONE all-reduce (800 KB at n = 100k);
* strict band depth / simplex / point-cloud depths have independent queries -> every rank takes a
  contiguous block of queries and the results are joined with ONE all-gather.

Nothing here touches the GPU: the compute callable is injected, so the partition / collective /
re-assembly logic is exercised on CPU with the gloo backend (tests/test_distributed.py).
"""
import os

import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when not distributed."""
    if os.environ.get("STATDEPTH_DISTRIBUTED", "1") == "0":
        return 0, 1
    try:
        import torch.distributed as dist
    except Exception:  # torch missing: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block(total: int, rank: int, size: int):
    """Balanced contiguous block [lo, hi) of `total` items for `rank` of `size`."""
    base, rem = divmod(int(total), int(size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for_collectives():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum(arr: np.ndarray) -> np.ndarray:
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allgather_blocks(local: np.ndarray, total: int) -> np.ndarray:
    """Concatenate per-rank blocks (sizes given by block()) along axis 0."""
    import torch
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()
    dev = _device_for_collectives()
    sizes = [block(total, r, size) for r in range(size)]
    width = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(parts, mine)
    out = [p.cpu().numpy()[: hi - lo] for p, (lo, hi) in zip(parts, sizes)]
    return np.concatenate(out, axis=0)


def relaxed_counts(compute, X, queries, j):
    """compute(X_rows, queries, j) -> int64[nq]; rows sharded, counts all-reduced."""
    rank, size = world()
    if size == 1:
        return compute(X, queries, j)
    T = X.shape[0]
    lo, hi = block(T, rank, size)
    nq = X.shape[1] if queries is None else len(queries)
    local = compute(X[lo:hi], queries, j) if hi > lo else np.zeros(nq, dtype=np.int64)
    return allreduce_sum(np.asarray(local, dtype=np.int64))


def query_sharded(compute, queries, dtype):
    """compute(query_block) -> array[len(block)]; queries sharded, results all-gathered."""
    rank, size = world()
    queries = np.asarray(queries, dtype=np.int64)
    if size == 1:
        return compute(queries)
    lo, hi = block(len(queries), rank, size)
    local = compute(queries[lo:hi]) if hi > lo else np.zeros(0, dtype=dtype)
    return allgather_blocks(np.asarray(local, dtype=dtype), len(queries))
