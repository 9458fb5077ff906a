"""Optimizer of the training loop (reference train.py:380-409: `torch.optim.Adam(lr)` with default betas / eps and no
weight decay) as ONE kernel launch over a flat parameter buffer, fused with the data-parallel gradient exchange.

`FusedAdam` re-homes the parameters it is given as views of one flat fp32 buffer (reverse registration order = the
order in which the backward produces their gradients) and their `.grad` as views of `parallel.GradientBuckets.flat`;
`step()` waits for the bucketed all-reduce (NCCL over NVLink when torch.distributed is initialised), checks the
gradients for inf / NaN (loss-scaled fp16 activation gradients can overflow: the step is skipped then, as with AMP),
runs `nm_adam_step`, and drops the packed-weight caches of `owner` (the kernel edits the parameters through raw
pointers, which does not bump `Tensor._version`).  Create it AFTER moving the model to its device."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import autograd as AG
from . import ops
from .parallel import GradientBuckets


class FusedAdam:
    def __init__(self, params, lr: float = 4e-4, betas=(0.9, 0.999), eps: float = 1e-8, owner: Optional[torch.nn.Module] = None,
                 n_buckets: int = 4, overlap: bool = True):
        self.lr, self.betas, self.eps, self.owner = float(lr), (float(betas[0]), float(betas[1])), float(eps), owner
        self.buckets = GradientBuckets(params, n_buckets=n_buckets)
        self.params = self.buckets.params
        dev = self.params[0].device
        if dev.type != "cuda":
            raise L.NmError("FusedAdam needs CUDA parameters (there is no CPU fallback)")
        total = self.buckets.flat.numel()
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                if p.dtype != torch.float32:
                    raise L.NmError("FusedAdam: fp32 parameters only")
                n = p.numel()
                view = self.flat_param[off:off + n].view_as(p)
                view.copy_(p)
                p.data = view
                off += n
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)   # step_device(): steps applied / skipped, on the device
        self._counters_live = False      # set by the first step_device(): the device counters are authoritative from then on
                                         # (a captured graph advances them without Python); host-side steps mirror into them
        self.steps = 0
        self.skipped = 0
        self._good = 0
        self._hooks = self.buckets.attach_hooks() if overlap else []
        if owner is not None:
            ops.invalidate_caches(owner)

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Zero the flat gradient buffer; the `.grad` views stay attached whatever `set_to_none` says."""
        self.buckets.zero()

    def step_device(self) -> None:
        """`step()` without any host synchronisation: the inf / NaN check, the skip decision and the step count (for the
        bias corrections) stay on the device (`nm_adam_step_dev`), so the call can be captured in a CUDA graph together
        with the forward and the backward (`graph.CapturedTrainStep`).  The loss scale is not adapted in this mode;
        `sync_counters()` brings `steps` / `skipped` back to the host."""
        if not self._counters_live:
            self._upload_counters()
            self._counters_live = True
        self.buckets.finish()
        g = self.buckets.flat
        self.flag.zero_()
        L.call("nm_grad_nonfinite", L.ptr(g), g.numel(), L.ptr(self.flag), L.stream())
        L.call("nm_adam_step_dev", L.ptr(self.flat_param), L.ptr(g), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), g.numel(),
               self.lr, self.betas[0], self.betas[1], self.eps, L.ptr(self.counters), 1.0, L.ptr(self.flag), L.stream())
        if self.owner is not None:
            ops.invalidate_caches(self.owner)

    def _upload_counters(self) -> None:
        self.counters.copy_(torch.tensor([self.steps, self.skipped], dtype=torch.int32), non_blocking=False)

    def sync_counters(self) -> None:
        """After `step_device()` calls or graph replays: read the device-side step / skip counts back (one synchronisation)."""
        if self._counters_live:
            c = self.counters.tolist()
            self.steps, self.skipped = int(c[0]), int(c[1])

    def step(self) -> bool:
        """-> False when a gradient was inf / NaN and the update was skipped."""
        self.sync_counters()
        self.buckets.finish()
        g = self.buckets.flat
        self.flag.zero_()
        L.call("nm_grad_nonfinite", L.ptr(g), g.numel(), L.ptr(self.flag), L.stream())
        if int(self.flag.item()) != 0:
            # an fp16 activation gradient overflowed somewhere in the backward: skip, and scale the next pass 4x lower
            self.skipped += 1
            self._good = 0
            AG.adjust_headroom(-2)
            if self._counters_live:
                self._upload_counters()
            return False
        self.steps += 1
        self._good += 1
        if self._good % 500 == 0 and AG.headroom_log2() < AG.GRAD_HEADROOM_LOG2:
            AG.adjust_headroom(+1)
        L.call("nm_adam_step", L.ptr(self.flat_param), L.ptr(g), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), g.numel(),
               self.lr, self.betas[0], self.betas[1], self.eps, self.steps, 1.0, None, L.stream())
        if self.owner is not None:
            ops.invalidate_caches(self.owner)
        if self._counters_live:
            self._upload_counters()
        return True

    def state_dict(self) -> dict:
        self.sync_counters()
        return dict(steps=self.steps, lr=self.lr, betas=self.betas, eps=self.eps, exp_avg=self.exp_avg.clone(),
                    exp_avg_sq=self.exp_avg_sq.clone())

    def load_state_dict(self, state: dict) -> None:
        self.steps, self.lr, self.betas, self.eps = int(state["steps"]), float(state["lr"]), tuple(state["betas"]), float(state["eps"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        if self._counters_live:
            self._upload_counters()
