"""Multi-GPU plumbing of the sampling path: the batch shards by image (independent chains, as the
reference's DistributedSampler evaluation does, /root/reference/ldmseg/trainers/trainers_ldm_cond.py:243-245),
weights are replicated, and the only collective is one all-gather of the decoded ids per global batch."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_ids(local_ids: torch.Tensor, global_batch: int) -> torch.Tensor:
    """all-gather per-rank id maps [b_r, H, W] (uint8) into the global batch order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ids
    world = dist.get_world_size()
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    maxb = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxb,) + tuple(local_ids.shape[1:]), dtype=local_ids.dtype, device=local_ids.device)
    pad[: local_ids.shape[0]] = local_ids
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
