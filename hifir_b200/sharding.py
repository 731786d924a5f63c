"""Host-side sharding for the multi-GPU runs: one apply does not shard (sequential levels), so
every rank holds a full factor replica and owns a contiguous range of right-hand-side columns
(or of independent systems).  No collective is on the hot path; results are gathered once."""
from __future__ import annotations

import numpy as np


def shard_range(total, world, rank):
    """Contiguous, balanced [begin, end) of `total` items for `rank` of `world` (first ranks get the
    remainder)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def local_columns(B, world, rank):
    """The rank's slice of a row-interleaved block B[n, nrhs] (copy, contiguous)."""
    b, e = shard_range(B.shape[1], world, rank)
    return np.ascontiguousarray(B[:, b:e])


def gather_columns(X_local, nrhs, world, rank, dist=None, device=None):
    """all_gather the per-rank column slices back into X[n, nrhs] (torch.distributed; NCCL on GPUs,
    gloo in the CPU tests).  Slices may be ragged: they are padded to the widest one."""
    import torch
    n = X_local.shape[0]
    width = max(shard_range(nrhs, world, r)[1] - shard_range(nrhs, world, r)[0] for r in range(world))
    t = torch.zeros(n, width, dtype=torch.float64, device=device)
    xl = torch.as_tensor(X_local, dtype=torch.float64, device=device)
    t[:, : xl.shape[1]] = xl
    parts = [torch.empty_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, t)
    else:
        parts = [t]
    X = torch.empty(n, nrhs, dtype=torch.float64, device=device)
    for r in range(world):
        b, e = shard_range(nrhs, world, r)
        X[:, b:e] = parts[r][:, : e - b]
    return X
