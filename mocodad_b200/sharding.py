"""Window sharding across GPUs.

Windows are independent (no cross-window op in mocodad.py:129-184; BatchNorm is in eval mode), so
the dataset index space is split into contiguous per-rank ranges, every rank scores its range with
the full (tiny) weight set, and ONE all-gather of the per-window fp32 scores ends the epoch.  The
Philox noise is keyed by the global window index, so scores do not depend on the rank count.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first ``n_items % world_size`` ranks get one extra item."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_scores(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-window scores of every rank's ``shard_bounds`` range -> [n_total] on all ranks.

    One collective on a padded buffer (``all_gather_into_tensor``; shards differ by at most one item)."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.numel() != n_total:
            raise ValueError("single process: local scores must cover the whole range")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_total, rank, world)
    if local.numel() != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} local scores, got {local.numel()}")
    width = -(-n_total // world)  # ceil
    if n_total % world == 0:   # equal shards (the bench's weak scaling, any dataset whose size divides): no padding, no re-assembly
        recv = torch.empty(n_total, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, local.reshape(-1).contiguous(), group=group)
        return recv
    send = torch.zeros(width, dtype=local.dtype, device=local.device)
    send[: hi - lo] = local.reshape(-1)
    recv = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_total, r, world)
        parts.append(recv[r * width: r * width + (b - a)])
    return torch.cat(parts)


def score_sharded(engine, data: torch.Tensor, n_generated_samples: int, *, seed: int = 0,
                  first_window: int = 0, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Score the rank's contiguous slice of ``data`` (the same full batch on every rank, on the
    engine's device) and return all ``data.shape[0]`` scores on every rank."""
    n = data.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(n, rank, world)
    local = engine.reverse_diffusion(data[lo:hi].contiguous(), n_generated_samples, seed=seed,
                                     first_window=first_window + lo)["best"]
    return gather_scores(local, n, group)
