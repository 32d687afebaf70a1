"""Multi-GPU plumbing: replicate index-range sharding and the summary all-reduce (the path's only collective).

One process per GPU. Replicates are independent, so rank g of G takes the contiguous global index range
``[g*n/G, (g+1)*n/G)`` and passes its start as ``replicate_offset``; Philox coordinates are global ids, hence the union of the
shards is bit-identical to a single-GPU run. Afterwards the int64 summary vectors are summed across ranks
(NCCL on GPUs, gloo in the CPU test)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """(offset, count) of rank's contiguous share of n_total replicates; shares differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def allreduce_summary(vec: np.ndarray, device: str = "cpu") -> np.ndarray:
    """Sums an int64 summary vector over all ranks of the default process group (no-op without one)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.int64)).to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
