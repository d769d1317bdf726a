"""Multi-GPU plumbing: the batch shards by contiguous instance-index ranges, one rank (process) per GPU, with no
data-path collective (instances never interact: each tiny_solve touches one workspace, admm.cpp:111-152).  The only
collective is the final statistics gather (iteration histogram, totals, max time) over torch.distributed -- NCCL on
GPUs, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(rank: int, world: int, per_rank: int | None = None, total: int | None = None):
    """Instance indices [b0, b1) owned by `rank`.  Weak scaling: give per_rank; strong scaling: give total."""
    if per_rank is not None:
        return rank * per_rank, (rank + 1) * per_rank
    assert total is not None
    return total * rank // world, total * (rank + 1) // world


def local_stats(iters: np.ndarray, status: np.ndarray, max_iter: int):
    """Per-shard statistics vector: [sum(iter), #solved, #instances, hist[0..max_iter]] as float64."""
    hist = np.bincount(np.asarray(iters, dtype=np.int64), minlength=max_iter + 1)[: max_iter + 1]
    return np.concatenate([[float(np.sum(iters, dtype=np.int64)), float(np.sum(status == 1)), float(len(iters))],
                           hist.astype(np.float64)])


def gather_stats(vec, times_ms, dist=None, device=None):
    """Sum the statistics vector and take the max of the timing vector over all ranks (no-op without a group)."""
    import torch
    v = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=device)
    t = torch.as_tensor(np.asarray(times_ms, dtype=np.float64), device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return v.cpu().numpy(), t.cpu().numpy()
