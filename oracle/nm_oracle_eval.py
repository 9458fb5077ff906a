"""CPU oracle for the consumers of the path's outputs: the evaluation metrics of
`utils/eval_utils.py` and the retarget post-processing of `vis_retarget.py`.

TEST INFRASTRUCTURE ONLY (same rule as nm_oracle.py).  Restated with numpy / torch on the CPU, side-effect free
(the reference mutates its inputs; the functions here return the mutated copies instead).  Pinned by
`oracle/make_golden_eval.py` against the reference's own functions run in the build container
(`tests/golden/eval_retarget.npz`).
"""
from __future__ import annotations

import numpy as np
import torch


def voxel_chamfer_per_frame(gt: torch.Tensor, recon: torch.Tensor):
    """utils/eval_utils.py:29-56.  gt, recon (n, G, G, G) -> (per-frame chamfer (n,) float64, binarised recon)."""
    n, G = gt.shape[0], gt.shape[-1]
    recon = recon.clone()
    recon[recon >= 0.5] = 1                                             # :37
    recon[recon < 0.5] = 0                                              # :38
    out = np.zeros(n)
    for f in range(n):
        a = torch.stack(torch.where(gt[f]), dim=-1) / ((G - 1) / 2) - 1            # :43
        b = torch.stack(torch.where(recon[f]), dim=-1) / ((G - 1) / 2) - 1         # :44
        d = (a[:, None] - b[None]).pow(2).sum(dim=-1)                               # :45
        out[f] = (d.min(dim=-1).values.mean() + d.min(dim=0).values.mean()).item()  # :46
    return out, recon


def voxel_chamfer_distance(scores, voxel: torch.Tensor, recon: torch.Tensor):
    """The dict of utils/eval_utils.py:52-56 for voxel / recon (B, T, 1, G, G, G)."""
    B, T, _, G = voxel.shape[:4]
    per, _ = voxel_chamfer_per_frame(voxel.reshape(B * T, G, G, G), recon.reshape(B * T, G, G, G))
    per = per.reshape(B, T)
    scores = [] if scores is None else scores
    for b in range(B):
        scores.append([per[b].sum() / T])
    return dict(scores=scores, scores_log=per.sum() / (B * T))


def semantic_nearest(keypoints: torch.Tensor, gt_keypoints: torch.Tensor, threshold: float = 0.2):
    """utils/eval_utils.py:65-84.  keypoints (B, T, K, 4), gt (B, T, K', 3) ->
    (masked keypoints, closest index (B*T, K') int64, histogram (K', K) int64)."""
    B, T, K, _ = keypoints.shape
    kp = keypoints.clone()
    kp[torch.where(kp[..., -1] < threshold)] = torch.tensor([1e4, 1e4, 1e4, 1.0])    # :68-69
    d = (gt_keypoints[:, :, :, None] - kp[:, :, None, :, :3]).pow(2).sum(-1)         # :79
    idx = d.min(dim=-1).indices.reshape(B * T, -1)                                   # :80-81
    Kgt = idx.shape[1]
    hist = np.zeros((Kgt, K), dtype=np.int64)
    for k in range(Kgt):
        hist[k] = np.bincount(idx[:, k].numpy(), minlength=K)                        # :84 one_hot[...].sum(0)
    return kp, idx, hist


def semantic_scores(scores, keypoints, gt_keypoints):
    """The dict of utils/eval_utils.py:86-90."""
    _, _, hist = semantic_nearest(keypoints, gt_keypoints)
    scores = np.zeros(hist.shape) if scores is None else scores
    scores += hist
    temp = np.array([(h / h.sum()).max() for h in hist], dtype=np.float32)
    return dict(scores=scores, scores_log=temp.mean())


def extract_skin_weights(priority_indices, parents, points: np.ndarray, keypoints: torch.Tensor, hardness=8.0,
                         threshold=0.2) -> np.ndarray:
    """vis_retarget.py:21-62 vectorised over the points (the reference loops over them in Python)."""
    parents = [int(p) for p in parents]
    K = keypoints.shape[0]
    invalid = (keypoints[:, -1] < threshold)                                         # :33
    pts = torch.from_numpy(points).to(keypoints.dtype)
    bones = torch.zeros(K, 3, dtype=keypoints.dtype)
    for k in range(K):
        p = parents[k]
        if p == k:
            bones[k] = keypoints[k, :3]                                               # :39
        else:
            while bool(invalid[p]):                                                   # :41-42
                p = parents[p]
            bones[k] = (keypoints[k, :3] + keypoints[p, :3]) / 2                      # :44
    dist = (pts[:, None] - bones[None]).pow(2).sum(dim=-1).sqrt()                    # :46
    dist[:, invalid] = 1e4                                                            # :48
    dist[:, int(priority_indices[0])] = 1e4                                           # :49
    child = dist.argmin(dim=-1)                                                       # :52
    parent = torch.tensor(parents)[child]                                             # :56
    cd = ((pts - keypoints[child, :3]).pow(2).sum(-1).sqrt() * hardness).exp()        # :57
    pd = ((pts - keypoints[parent, :3]).pow(2).sum(-1).sqrt() * hardness).exp()       # :58
    w = torch.zeros(len(pts), K, dtype=keypoints.dtype)
    rows = torch.arange(len(pts))
    w[rows, parent] = cd / (cd + pd)                                                  # :59
    w[rows, child] = pd / (cd + pd)                                                   # :60
    return w.numpy()


def retarget_fk(R: torch.Tensor, offset: torch.Tensor, root_pos: torch.Tensor, order, parents, clip=True):
    """vis_retarget.py:279-287, :300.  R (T, K, 3, 3), offset (K, 3), root_pos (T, 3) -> (T, K, 3)."""
    T, K = R.shape[:2]
    order = [int(i) for i in order]
    pos = torch.zeros(T, K, 3, dtype=R.dtype)
    pos[:, order[0]] = root_pos
    for idx in order[1:]:
        pos[:, idx] = (R[:, idx] @ offset[idx][:, None]).squeeze(-1) + pos[:, int(parents[idx])]
    return pos.clip(-1, 1) if clip else pos


def linear_blend_skinning(points: np.ndarray, joints: np.ndarray, R_inv, T3x4: np.ndarray, skin: np.ndarray):
    """vis_retarget.py:263-270 + :315-322 in float64.  T3x4 (T, K, 3, 4) -> (T, N, 3)."""
    points = np.asarray(points, np.float64)
    off = points[:, None] - np.asarray(joints, np.float64)[None]                      # (N, K, 3)  :264
    local = off if R_inv is None else np.einsum('kij,nkj->nki', np.asarray(R_inv, np.float64), off)   # :265
    homo = np.concatenate([local, np.ones(local.shape[:2] + (1,))], axis=-1)          # :315
    out = []
    for t in range(T3x4.shape[0]):
        kin = np.einsum('kij,nkj->kin', np.asarray(T3x4[t], np.float64), homo)        # :319
        out.append(np.einsum('nk,kin->ni', np.asarray(skin, np.float64), kin)[..., :3])   # :320
    return np.stack(out, axis=0)
