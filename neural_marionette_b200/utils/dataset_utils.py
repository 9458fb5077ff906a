"""Drop-in for the reference's utils/dataset_utils.py (crop / normalise / voxelize).

``voxelize`` keeps the reference contract — numpy (N, >=3) in, numpy float32
(1, G, G, G) out, bit-exact — but the scatter runs on the GPU through
``nm_voxelize``.  The batched entry points (`voxelize_clip`, `voxelize_raw_clips`)
keep the data on the device, which is what the detector wants.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def crop_sequence(seq, start, T, sample_rate=1):
    """Strided slice of the frame axis (reference dataset_utils.py:6-7)."""
    stop = start + T * sample_rate
    return seq[start:stop:sample_rate]


def episodic_normalization(seq, scale=1.0, x_trans=0.0, z_trans=0.0, joints=None):
    """Clip-global isotropic bounding-box normalisation to [-1, 1) (reference dataset_utils.py:9-19).
    Host-side numpy, kept for callers that want the normalised points themselves; the fused device
    path is `voxelize_raw_clips`."""
    lo = seq.min(axis=(0, 1))
    span = (seq.max(axis=(0, 1)) - lo).max() + 1e-5
    shift = np.array([x_trans, 0, z_trans])
    out = ((seq - lo[None, None]) * scale / span) * 2 - 1 + shift
    if joints is None:
        return out
    return out, ((joints - lo[None, None]) * scale / span) * 2 - 1


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("neural_marionette_b200.voxelize needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def voxelize(pos_coords, output_shape, is_binarized=True, device=None):
    """points (N, >=3) numpy -> (1, G, G, G) float32 numpy occupancy (reference dataset_utils.py:21-31)."""
    shape = tuple(int(s) for s in output_shape)
    if len(set(shape)) != 1:
        raise ValueError("only cubic grids are supported (the model is hard-wired to grid_size^3)")
    pts = np.ascontiguousarray(np.asarray(pos_coords)[..., :3])
    if pts.dtype not in (np.float64, np.float32):
        pts = pts.astype(np.float64)
    dev = _device(device)
    grid = ops.voxelize_points(torch.from_numpy(pts).to(dev)[None], shape[0])
    return grid.cpu().numpy()


def voxelize_clip(points, grid_size, device=None):
    """(T, N, 3) normalised points (numpy or tensor) -> (T, 1, G, G, G) fp32 CUDA tensor: the per-frame
    loop of the callers (reference vis_generation.py:19-23, dataset/dataset.py:170-183) in one launch."""
    if isinstance(points, np.ndarray):
        points = torch.from_numpy(np.ascontiguousarray(points[..., :3]))
    dev = _device(device if device is not None else (points.device if points.is_cuda else None))
    grid = ops.voxelize_points(points.to(dev), grid_size)
    return grid[:, None]


def voxelize_raw_clips(raw_points, grid_size, scale=1.0, x_trans=0.0, z_trans=0.0, device=None):
    """(B, T, N, 3) raw float32 points -> (B, T, 1, G, G, G) fp32 CUDA tensor; normalisation and scatter fused
    on the device (bit-exact with episodic_normalization + voxelize for float32 inputs)."""
    if isinstance(raw_points, np.ndarray):
        raw_points = torch.from_numpy(np.ascontiguousarray(raw_points[..., :3], dtype=np.float32))
    dev = _device(device if device is not None else (raw_points.device if raw_points.is_cuda else None))
    return ops.normalize_voxelize(raw_points.to(dev), grid_size, scale, x_trans, z_trans)
