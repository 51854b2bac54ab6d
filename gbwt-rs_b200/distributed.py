"""Multi-GPU plumbing: query / path sharding and the optional result gather.

The hot path has no exchange step: the index is replicated in every GPU's HBM and each rank processes a
contiguous block of the batch. `torch.distributed` is used only to gather fixed-size results (24 bytes per
SearchState) to one rank when a caller wants them in one place; it works with NCCL (GPU tensors over NVLink)
and with gloo (CPU tensors, used by the tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous block [lo, hi) of n items owned by `rank`; block sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_states(local: np.ndarray, n_total: int, dst: int = 0, device=None):
    """Gathers per-rank result blocks (rows of u64 words, e.g. SearchState = 3 words) on rank `dst`.

    `local` holds this rank's shard_range block. Returns the (n_total, words) array on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    # words per row from the dtype (also right for a rank whose block is empty): SearchState = 3, BidirectionalState = 6
    local = np.asarray(local)
    words = max(1, local.dtype.itemsize // 8) * (int(np.prod(local.shape[1:])) if local.ndim > 1 else 1)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    width = max(sizes) if sizes else 0
    buf = torch.zeros((width, words), dtype=torch.int64, device=device)
    if len(local):
        buf[: len(local)] = torch.from_numpy(np.ascontiguousarray(local).view(np.int64).reshape(len(local), words)).to(buf.device)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    if rank != dst:
        return None
    out = np.zeros((n_total, words), dtype=np.uint64)
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi] = gathered[r][: hi - lo].cpu().numpy().view(np.uint64)
    return out
