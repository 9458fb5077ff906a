"""Gradient oracle for the config #4 training step (TEST INFRASTRUCTURE ONLY).

The oracle's forward (`nm_oracle.detector_forward`) is a functional graph over a `state_dict`, so
`torch.autograd` differentiates it as is.  `oracle/make_golden_grad.py` pins these gradients against the
REFERENCE's own `loss.backward()` (bit-identical for all 314 trainable detector tensors) and writes
`tests/golden/detector_grad_g32.npz`; the future CUDA backward is tested against this file / this function.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import nm_oracle as O

# train.py:68-78, 177-181 — stage-1 (detector) loss weights of the reference's argparse defaults
DETECTOR_LOSS_WEIGHTS = {
    "recon_loss": 100.0, "sparsity_loss": 5.0, "separation_loss": 0.1, "vol_fit_reg": 10.0, "kypt_const_loss": 0.0,
    "local_const_loss": 1e-3, "time_const_loss": 1.0, "sparsity_const_loss": 0.01, "intensity_const_loss": 0.01,
    "graph_traj_loss": 1.0, "graph_vol_loss": 0.0,
}


def detector_loss(out: dict, recon_only: bool) -> torch.Tensor:
    """train.py:400-420: the weighted sum of the detector's loss terms (config #4 names the reconstruction term)."""
    if recon_only:
        return DETECTOR_LOSS_WEIGHTS["recon_loss"] * out["recon_loss"]
    total = 0.0
    for name, w in DETECTOR_LOSS_WEIGHTS.items():
        if w != 0.0:
            total = total + w * out[name]
    return total


def detector_gradients(vox: torch.Tensor, sd: Dict[str, torch.Tensor], hp, recon_only: bool = True):
    """(loss, {key: dLoss/dParam}) of `KyptDetector.forward` on vox (B, T, 1, G, G, G) for every floating-point
    tensor of the `kypt_detector.*` part of the state dict that takes part in the graph."""
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k.startswith("kypt_detector.") else v)
              for k, v in sd.items()}
    out = O.detector_forward(vox, leaves, hp)
    loss = detector_loss(out, recon_only)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaves.items() if getattr(v, "grad", None) is not None}
