"""Data-parallel sharding of utterances across ranks (one process per GPU).

The enhance path has no cross-utterance coupling (per-clip normalisation, per-clip attention, BatchNorm in eval
mode), so a batch is split contiguously by rank and each rank runs its own plan; the only collective is an
OPTIONAL gather of the enhanced waveforms (NCCL on GPUs, gloo in the CPU tests).  SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the items owned by `rank`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def enhance_sharded(clips: torch.Tensor, enhance_fn: Callable[[torch.Tensor], torch.Tensor], gather: bool = True,
                    micro_batch: int = 64) -> torch.Tensor:
    """Every rank holds the same [N, n] batch, enhances its shard in micro-batches with `enhance_fn`
    ([b, n] -> [b, n]) and, if `gather`, all ranks receive the full [N, n] result."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(clips.shape[0], rank, world)
    outs: List[torch.Tensor] = [enhance_fn(clips[i:min(i + micro_batch, hi)]) for i in range(lo, hi, micro_batch)]
    mine = torch.cat(outs) if outs else clips.new_empty((0, clips.shape[1]))
    if not gather or world == 1:
        return mine
    sizes = [shard_range(clips.shape[0], r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    padded = clips.new_zeros((biggest, clips.shape[1]))
    padded[:mine.shape[0]] = mine
    bucket = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bucket, padded)
    return torch.cat([bucket[r][:b - a] for r, (a, b) in enumerate(sizes)])
