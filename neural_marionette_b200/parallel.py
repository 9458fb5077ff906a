"""Clip-level data parallelism: clips are independent through the whole path (GroupNorm is per
sample, the dynamics are per clip), so inference shards the clip dimension across ranks with NO data-path
collective; the only communication is the optional gather of the (tiny) keypoint outputs and, for timing, a
barrier / max-reduce.  One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first (n_items % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float, device=None) -> float:
    """Device-side time of a multi-GPU step = the slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_clip_outputs(local: torch.Tensor, n_clips_total: int) -> torch.Tensor:
    """All-gather per-clip outputs (e.g. keypoints (b_local, T, K, 4)) of a `shard_range` split back into clip
    order on every rank.  Shards may differ by one clip: they are padded to the largest for the collective."""
    rank, size = world()
    if size == 1:
        return local
    sizes = shard_sizes(n_clips_total, size)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(size)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
