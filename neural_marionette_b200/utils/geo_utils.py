"""Rotation helpers (drop-in for the reference's utils/geo_utils.py).  The fused HSVRNN step kernel does
this arithmetic in-kernel; these torch versions serve callers that use the helpers directly."""
from __future__ import annotations

import torch


def normalize_vector(v, return_mag=False):
    mag = v.pow(2).sum(1).sqrt() + 1e-10
    out = v / mag[:, None]
    return (out, mag) if return_mag else out


def cross_product(u, v):
    return torch.stack((u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1],
                        u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2],
                        u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]), dim=1)


def compute_rotation_matrix_from_6d(param):
    """(..., 6) -> (..., 3, 3) by Gram-Schmidt; columns are (x, y, z) (reference geo_utils.py:56-78)."""
    lead = param.shape[:-1]
    flat = param.reshape(-1, 6)
    x = normalize_vector(flat[:, :3])
    z = normalize_vector(cross_product(x, flat[:, 3:]))
    y = cross_product(z, x)
    return torch.stack((x, y, z), dim=2).reshape(*lead, 3, 3)


def compute_global_rot_from_local_rot(params, priority, parents, inverse=False):
    """Chain local rotations down the skeleton in priority order -> {joint: (B, 3, 3)} (reference geo_utils.py:3-27)."""
    local = compute_rotation_matrix_from_6d(params)
    order = [int(i) for i in priority.indices]
    out = {order[0]: local[:, order[0]]}
    for j in order[1:]:
        up = out[int(parents[j])]
        out[j] = torch.bmm(local[:, j], up) if inverse else torch.bmm(up, local[:, j])
    return out
