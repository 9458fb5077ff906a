"""Evaluation metrics over the path's outputs — drop-in for the reference's `utils/eval_utils.py`.

Same function names, arguments, returned dicts and in-place side effects (`params['recon']` is binarised,
invalid rows of `params['keypoints']` are overwritten), but the per-(b, t) Python loops with their `.item()`
syncs are one `nm_voxel_chamfer` / `nm_semantic_nearest` launch sequence and ONE device->host copy.
CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops


def evaluate(name, scores_dict, params):
    """Running update of one metric (`utils/eval_utils.py:4-10`): `scores_dict[name]` is the accumulator returned by
    the previous call (None the first time)."""
    metric = _METRICS.get(name)
    if metric is None:
        raise ValueError("invalid evaluation metric.")
    return metric(scores_dict[name], params)


def evaluate_final(name, scores_dict, result_dir='pretrained/results'):
    """Final figure of one metric over the whole evaluation set (`utils/eval_utils.py:12-27`); the per-joint /
    per-clip table goes to `<result_dir>/semantic/semantic_result.csv` or `<result_dir>/chamfer/chamfer_result.csv`
    (the reference's relative paths; the directory is created here)."""
    if name not in _METRICS:
        raise ValueError("invalid evaluation metric.")
    sub, fname = ('semantic', 'semantic_result.csv') if name == 'semantic' else ('chamfer', 'chamfer_result.csv')
    os.makedirs(os.path.join(result_dir, sub), exist_ok=True)
    target = os.path.join(result_dir, sub, fname)
    if name == 'semantic':
        hist = scores_dict[name]                       # (K', K) match counts, normalised IN PLACE like the reference
        hist /= hist[0].sum()
        best = hist.max(axis=-1)                       # (K',): share of frames on the most frequent detected keypoint
        np.savetxt(target, best, delimiter=",")
        return best.mean()
    per_clip = np.array(scores_dict[name])             # (clips, 1)
    np.savetxt(target, per_clip, delimiter=",")
    return per_clip.mean() * 1e4                       # reported x 1e4


def _dense_view(t: torch.Tensor, shape):
    """A dense fp32 alias of `t` when possible (so that in-place effects reach the caller's tensor), else a copy."""
    if t.dtype == torch.float32 and t.is_contiguous():
        return t.view(shape), True
    return t.float().contiguous().view(shape), False


def voxel_chamfer_distance(scores, params):
    """utils/eval_utils.py:29-56.  params['voxel'], params['recon']: (B, T, 1, G, G, G) on the GPU."""
    B, T, C, *X = params['voxel'].size()
    if scores is None:
        scores = []
    G = X[0]
    gt, _ = _dense_view(params['voxel'], (B * T, G, G, G))
    recon, aliased = _dense_view(params['recon'], (B * T, G, G, G))
    chamfer, _, err = ops.voxel_chamfer(gt, recon, binarize=True)
    if not aliased:                                   # recon[recon >= 0.5] = 1; recon[recon < 0.5] = 0 (:37-38)
        params['recon'].copy_(recon.view(params['recon'].shape))
    host = torch.cat([chamfer, err.float()]).cpu().numpy()      # one device->host copy
    if host[-1] != 0:
        raise IndexError("voxel_chamfer_distance: a frame has no occupied voxel in gt or recon "
                         "(the reference's min over an empty dimension raises here too)")
    per_frame = host[:-1].astype(np.float64).reshape(B, T)
    for b in range(B):
        scores.append([float(per_frame[b].sum() / T)])
    return dict(scores=scores, scores_log=float(per_frame.sum() / (B * T)))


def semantic_scores(scores, params):
    """utils/eval_utils.py:60-90.  params['keypoints'] (B, T, K, 4), params['gt_keypoints'] (B, T, K', 3)."""
    B, T, K, _ = params['keypoints'].size()
    gt = params['gt_keypoints']
    K_gt = gt.size(2)
    kypt, aliased = _dense_view(params['keypoints'], (B * T, K, 4))
    _, hist = ops.semantic_nearest(kypt, gt.float().contiguous().view(B * T, K_gt, 3), 0.2)
    if not aliased:
        params['keypoints'].copy_(kypt.view(params['keypoints'].shape))
    if scores is None:
        scores = np.zeros((K_gt, K))
    counts = hist.cpu().numpy()                       # (K', K) over the B*T frames
    scores += counts
    temp = np.array([(counts[k] / counts[k].sum()).max() for k in range(K_gt)], dtype=np.float32)
    return dict(scores=scores, scores_log=temp.mean())


_METRICS = {'semantic': semantic_scores, 'voxel_chamfer': voxel_chamfer_distance}
