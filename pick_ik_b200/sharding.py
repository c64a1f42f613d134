"""Pose-batch sharding across ranks (one process per GPU).

The hot path partitions trivially: problems are independent and every random stream is keyed by the
GLOBAL problem index, so a shard solved with ``first_problem_index = shard start`` reproduces the
unsharded result bit for bit.  The only exchange is one all-gather of the packed per-problem results
(n + 3 doubles each) at the end -- torch.distributed over NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the pose batch owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_results(solution, error_code, cost, iterations):
    """[B, n + 3] float64: joints, cost, error_code, iterations (exact in binary64).  Works on numpy arrays
    and on torch tensors (device-side packing for the NCCL gather)."""
    if isinstance(solution, np.ndarray):
        return np.concatenate([solution, cost[:, None], error_code[:, None].astype(np.float64),
                               iterations[:, None].astype(np.float64)], axis=1)
    import torch

    return torch.cat([solution, cost[:, None], error_code.double()[:, None], iterations.double()[:, None]], dim=1)


def unpack_results(packed):
    n = packed.shape[1] - 3
    if isinstance(packed, np.ndarray):
        return dict(solution=packed[:, :n].copy(), cost=packed[:, n].copy(),
                    error_code=packed[:, n + 1].astype(np.int32), iterations=packed[:, n + 2].astype(np.int32))
    import torch

    return dict(solution=packed[:, :n], cost=packed[:, n], error_code=packed[:, n + 1].to(torch.int32),
                iterations=packed[:, n + 2].to(torch.int32))


def all_gather_results(packed, world: int, max_rows: int):
    """All-gathers the packed results of every rank (torch tensor on the backend's device).  Shards may
    differ by one row: each is padded to max_rows for the collective.  Returns [world, max_rows, n + 3]."""
    import torch
    import torch.distributed as dist

    rows, width = packed.shape
    if rows < max_rows:
        pad = torch.zeros((max_rows - rows, width), dtype=packed.dtype, device=packed.device)
        packed = torch.cat([packed, pad], dim=0)
    out = torch.empty((world, max_rows, width), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out.view(-1), packed.contiguous().view(-1))
    return out


def concat_shards(gathered, total: int, world: int):
    """Drops the padding rows of all_gather_results -> [total, n + 3] in global problem order."""
    parts = []
    for r in range(world):
        a, b = shard_range(total, r, world)
        parts.append(gathered[r, : b - a])
    if isinstance(gathered, np.ndarray):
        return np.concatenate(parts, axis=0)
    import torch

    return torch.cat(parts, dim=0)
