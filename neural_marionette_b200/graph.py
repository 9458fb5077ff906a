"""CUDA-graph replay of the detector for a fixed input shape.

A batch-1 `KyptDetector.forward` (BASELINE.json configs[0] shape: one 20-frame clip) enqueues ≈350 kernels whose
total device time is well below the time Python + the driver need to launch them; capturing the launch sequence
once and replaying it removes that gap.  Every C-ABI entry point only enqueues work on the stream it is given
(no allocation, no synchronisation, tensor maps passed by value), so the whole forward — including the
once-per-clip branch on its auxiliary stream — is capturable as is.  The replayed kernels are the same launches
with the same arguments: results are bit-identical to the eager call.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from . import ops


class CapturedCall:
    """Capture `fn(static_input)` (a function of ONE tensor returning a tensor / tuple / dict of tensors) into a
    CUDA graph; `__call__(x)` copies `x` into the static input, replays, and returns clones of the outputs
    (`clone=False`: the static buffers themselves, overwritten by the next call)."""

    def __init__(self, fn: Callable, example: torch.Tensor, warmup: int = 3):
        if not example.is_cuda:
            raise ops.L.NmError("CapturedCall needs a CUDA tensor (there is no CPU fallback)")
        self.static_in = example.clone()
        self.stream = torch.cuda.Stream(device=example.device)
        self.stream.wait_stream(torch.cuda.current_stream(example.device))
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):          # builds weight packs, scratch buffers and tables outside the graph
                fn(self.static_in)
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=self.stream):
            self.static_out = fn(self.static_in)

    def __call__(self, x: torch.Tensor, clone: bool = True):
        if x.shape != self.static_in.shape:
            raise ValueError(f"captured for input shape {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x)
        self.graph.replay()
        return _map(self.static_out, (lambda t: t.clone()) if clone else (lambda t: t))


def _map(out, f):
    if isinstance(out, torch.Tensor):
        return f(out)
    if isinstance(out, dict):
        return {k: _map(v, f) for k, v in out.items()}
    if isinstance(out, (tuple, list)):
        return type(out)(_map(v, f) for v in out)
    return out


def capture_detector(detector, example_seq: torch.Tensor, warmup: int = 3) -> CapturedCall:
    """`KyptDetector.forward` for occupancy clips shaped like `example_seq` (B, T, 1, G, G, G)."""
    return CapturedCall(lambda seq: detector(seq), example_seq, warmup)


def capture_detector_from_points(detector, example_points: torch.Tensor, grid_size: int, warmup: int = 3) -> CapturedCall:
    """Fused normalise + voxelize + `KyptDetector.forward` for raw fp32 point clips (B, T, N, 3) on the GPU."""
    def fn(raw) -> Dict[str, torch.Tensor]:
        vox = ops.normalize_voxelize(raw, grid_size, check=False)
        out = detector(vox)
        out["voxel"] = vox
        return out
    return CapturedCall(fn, example_points, warmup)


class CapturedTrainStep:
    """One training step - fused normalise + voxelize, `zero_grad`, forward, `loss_fn`, backward, `FusedAdam.step_device` -
    captured into a CUDA graph and replayed: the ~1 800 launches of the step are enqueued without Python or driver work
    between them (eager: 7 ms of the 132 ms step are GPU idle time).  The captured launches are the eager ones with the same
    arguments, so a replayed step is bit-identical to an eager `step_device()` step.  `warmup` REAL training steps run
    first (they build the scratch buffers the graph then refers to).  Frozen into the graph: the input shape, the loss
    scale and the learning rate; overflowed steps are skipped on the device and counted (`optimizer.sync_counters()`).
    Single-GPU only (the NCCL gradient exchange is driven from Python hooks)."""

    def __init__(self, detector, optimizer, loss_fn: Callable, example_points: torch.Tensor, grid_size: int, warmup: int = 2):
        if not example_points.is_cuda:
            raise ops.L.NmError("CapturedTrainStep needs CUDA tensors (there is no CPU fallback)")
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            raise ops.L.NmError("CapturedTrainStep: single-process training only")
        self.optimizer = optimizer
        self.static_in = example_points.clone()
        dev = example_points.device

        def one() -> torch.Tensor:
            vox = ops.normalize_voxelize(self.static_in, grid_size, check=False)
            optimizer.zero_grad()
            loss = loss_fn(detector(vox))
            loss.backward()
            optimizer.step_device()
            return loss.detach()

        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):
                one()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss = one()

    def __call__(self, points: torch.Tensor) -> torch.Tensor:
        """Copy `points` (B, T, N, 3) into the static input, replay the step, return the (static) loss tensor."""
        if points.shape != self.static_in.shape:
            raise ValueError(f"captured for input shape {tuple(self.static_in.shape)}, got {tuple(points.shape)}")
        self.static_in.copy_(points, non_blocking=True)
        self.graph.replay()
        return self.loss
