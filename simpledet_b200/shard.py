"""Image sharding of the hot path across ranks: every operator is independent per
image, so a global batch is split contiguously, B images per rank, with NO data-path collective —
exactly the reference's DataParallelExecutorGroup slicing (core/detection_module.py:443).
The only cross-rank traffic of a benchmark is the max-reduction of the device-timed duration."""
from __future__ import annotations

import torch
import torch.distributed as dist


def image_shard(global_images: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of [0, global_images) owned by `rank` (first ranks get the
    remainder, like np.array_split)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(global_images, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, device=None) -> float:
    """max of a per-rank scalar (elapsed ms); identity when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(units_per_rank: int, elapsed_ms_local: float, device=None) -> float:
    """units all ranks processed / max-over-ranks time (weak scaling: fixed work per rank)."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    return world * units_per_rank / (max_over_ranks(elapsed_ms_local, device) / 1e3)
