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
    """Drain this rank's device queue, meet the other ranks, and drain again (the barrier itself may enqueue work)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
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


class GradientBuckets:
    """Flat fp32 gradient storage + bucketed all-reduce for the data-parallel training step.

    The parameters are laid out in REVERSE registration order (the order in which a backward pass produces their
    gradients) in one flat buffer cut into `n_buckets` contiguous buckets of roughly equal size; `param.grad` of every
    parameter is a view into it, so the wgrad kernels write straight into the communication buffer.  The backward
    calls `ready(param)` as gradients are produced: when the last parameter of a bucket is ready its all-reduce (sum) is
    launched asynchronously and overlaps the rest of the backward; `finish()` waits for all buckets and divides by the
    world size.  Stage 1 of the reference trains 8 557 044 detector values (34.2 MB): 4 buckets of ≈ 8.6 MB, each
    far above NCCL's latency floor over NVLink and small enough to overlap.  Backend: NCCL on GPUs, gloo in the CPU tests.

    Aliasing contract: `zero()` (or `zero_grad(set_to_none=False)`) keeps `param.grad` inside the flat buffer.  If a
    caller drops the alias (`optimizer.zero_grad()` with torch's default `set_to_none=True`, `p.grad = None`), the next
    `ready(param)` / `finish()` notices that `param.grad` no longer points at its slot, copies the fresh gradient into
    the flat buffer and re-attaches it - so the all-reduce never runs on a stale buffer.  A parameter whose gradient is
    still `None` at `finish()` contributed nothing on this rank: its slot stays zero (as with an unused parameter).
    `attach_hooks()` registers post-accumulate hooks so that a plain `loss.backward()` drives `ready()`.
    """

    def __init__(self, params, n_buckets: int = 4, device=None):
        self.params = [p for p in params if p.requires_grad][::-1]
        if not self.params:
            raise ValueError("GradientBuckets: no trainable parameters")
        device = device if device is not None else self.params[0].device
        sizes = [p.numel() for p in self.params]
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.bounds: List[Tuple[int, int]] = []            # [lo, hi) offsets of the buckets in `flat`
        self._bucket_of, self._pending = {}, []
        target, lo, off, b = total / max(1, n_buckets), 0, 0, 0
        self._slot = {}
        for p, n in zip(self.params, sizes):
            p.grad = self.flat[off:off + n].view_as(p)
            self._slot[id(p)] = (off, n)
            self._bucket_of[id(p)] = b
            off += n
            last = p is self.params[-1]
            if last or (off >= target * (b + 1) and b < n_buckets - 1):     # cumulative thresholds: balanced cuts
                self.bounds.append((lo, off))
                lo, b = off, b + 1
        self._need = [0] * len(self.bounds)
        for p in self.params:
            self._need[self._bucket_of[id(p)]] += 1
        self._left, self._work, self._launched = list(self._need), [None] * len(self.bounds), [False] * len(self.bounds)
        self._finished = False

    def zero(self) -> None:
        self.flat.zero_()
        for p in self.params:
            self._attach(p, copy=False)
        self._left, self._work, self._launched = list(self._need), [None] * len(self.bounds), [False] * len(self.bounds)
        self._finished = False

    def _attach(self, p, copy: bool = True) -> None:
        """Make `p.grad` the view of its slot again (copying a detached gradient into the slot first)."""
        off, n = self._slot[id(p)]
        slot = self.flat[off:off + n]
        if p.grad is not None and p.grad.data_ptr() == slot.data_ptr() and p.grad.numel() == n:
            return
        if p.grad is not None and copy:
            slot.copy_(p.grad.reshape(-1))
        p.grad = slot.view_as(p)

    def attach_hooks(self):
        """Drive `ready()` from autograd: one post-accumulate hook per parameter (returns the handles)."""
        return [p.register_post_accumulate_grad_hook(lambda q: self.ready(q)) for p in self.params]

    def ready(self, param) -> None:
        """The gradient of `param` is complete on this rank."""
        if id(param) not in self._bucket_of:
            return
        self._attach(param)
        b = self._bucket_of[id(param)]
        self._left[b] -= 1
        if self._left[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        if self._launched[b]:
            return
        self._launched[b] = True
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            lo, hi = self.bounds[b]
            self._work[b] = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)

    def finish(self) -> None:
        """Launch whatever was not reported ready, wait for every bucket, average over the ranks.  Idempotent until the
        next `zero()`."""
        if self._finished:
            return
        self._finished = True
        for b, launched in enumerate(self._launched):
            if not launched:                                   # gradients nobody reported: make sure they are in the buffer
                for p in self.params:
                    if self._bucket_of[id(p)] == b and p.grad is not None:
                        self._attach(p)
        for b in range(len(self.bounds)):
            self._launch(b)
        for w in self._work:
            if w is not None:
                w.wait()
        _, size = world()
        if size > 1:
            self.flat.div_(size)
