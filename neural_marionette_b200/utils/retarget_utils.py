"""Retarget post-processing over the path's outputs — the library part of the reference's `vis_retarget.py`
(SURVEY.md §8f#4): skin-weight extraction (a Python loop over every point there), forward kinematics of the
retargeted skeleton and linear blend skinning, as `nm_skin_weights` / `nm_retarget_fk` /
`nm_linear_blend_skinning` launches.  CUDA tensors only.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _dev_f32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


def extract_skin_weights(A, priority, parents, points, keypoints, HARDNESS=8.0, THRESHOLD=0.2):
    """vis_retarget.py:21-62 — same arguments and return value: points (N, 3) numpy, keypoints (K, 4) torch ->
    (N, K) numpy float32.  `A` only supplies the device, as in the reference."""
    dev = A.device
    pts = _dev_f32(points, dev)
    kp = _dev_f32(keypoints, dev)
    par = torch.as_tensor(parents).to(device=dev, dtype=torch.int32).contiguous()
    root = int(priority.indices[0])
    skin, _, err = ops.skin_weights(pts, kp, par, root, HARDNESS, THRESHOLD)
    out = skin.cpu().numpy()
    if int(err.item()) != 0:
        raise RuntimeError("extract_skin_weights: the parent walk over invalid joints does not terminate "
                           "(the root joint is below THRESHOLD)")
    return out


def retarget_keypoints(R, offset, source_keypoints, priority, parents):
    """vis_retarget.py:275-301 (`--ours` branch): R (T, K, 3, 3) source rotations, offset (1, K, 3, 1) target bone
    offsets (`get_offset`), source_keypoints (1, T, K, 4) -> new_keypoints (1, T, K, 4): FK positions clipped to
    [-1, 1] with the source intensities."""
    dev = R.device
    T, K = R.shape[0], R.shape[1]
    order = priority.indices.to(device=dev, dtype=torch.int32).contiguous()
    par = torch.as_tensor(parents).to(device=dev, dtype=torch.int32).contiguous()
    root = int(priority.indices[0])
    root_pos = source_keypoints[0, :, root, :3].float().contiguous()
    pos = ops.retarget_fk(R.float().contiguous(), offset.float().reshape(K, 3).contiguous(), root_pos, order, par, True)
    return torch.cat([pos[None], source_keypoints[..., 3:].float()], dim=-1)


def linear_blend_skinning(points, joints, R_inv, R, pos, skin_weights):
    """vis_retarget.py:263-270 + :303-322: points (N, 3), joints (K, 3) bind-pose joint positions, R_inv (K, 3, 3)
    bind-pose inverse rotations (None: identity, the non-`--ours` branch), R (T, K, 3, 3) / pos (T, K, 3) posed
    transforms, skin_weights (N, K) -> (T, N, 3) numpy float32 (the reference returns float64 from numpy einsums)."""
    dev = R.device
    T3x4 = torch.cat([R.float(), pos.float().reshape(R.shape[0], R.shape[1], 3, 1)], dim=-1).contiguous()
    out = ops.linear_blend_skinning(_dev_f32(points, dev), _dev_f32(joints, dev),
                                    None if R_inv is None else _dev_f32(R_inv, dev), T3x4, _dev_f32(skin_weights, dev))
    return out.cpu().numpy()
