"""Heat-map / keypoint helpers and the detector's auxiliary losses (drop-in for the reference's
utils/kypt_detector_utils.py: same function names and tensor contracts).

On the detector's hot path the first three are fused into CUDA kernels (coordinate channels are synthesised
inside the first conv; soft-argmax and the Gaussian render live in ``nm_heatmap_head``); the functions here
serve code that calls the helpers directly.  The auxiliary losses act on (B, T, K, .) tensors — a few kB — and
are plain torch expressions, except the chamfer volume-fitting term which has its own kernel
(the reference materialises a (B, K, 3, 64^3) tensor for it).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops


def _axis(n, device):
    return torch.linspace(-1.0, 1.0, n, device=device)


def add_coord_channels(vox):
    """(B, C, X1..XD) -> (B, C + D, X1..XD): one linspace(-1, 1) channel per spatial axis."""
    sizes = vox.shape[2:]
    chans = []
    for d, n in enumerate(sizes):
        view = [1] * len(sizes)
        view[d] = n
        chans.append(_axis(n, vox.device).view(view).expand(*sizes))
    coords = torch.stack(chans, dim=0)[None].expand(vox.shape[0], -1, *sizes)
    return torch.cat([vox, coords], dim=1)


def extract_keypoints_from_heatmap(heatmap):
    """(B, K, G1..GD) -> (B, K, D + 1): soft-argmax coordinates of the sum-normalised marginals of
    (heatmap + 1e-6) plus intensity = mean / (max_k mean + 1e-6)."""
    nd = heatmap.dim() - 2
    spatial = tuple(range(2, 2 + nd))
    inten = heatmap.mean(dim=spatial)
    inten = inten / (inten.amax(dim=-1, keepdim=True) + 1e-6)
    shifted = heatmap + 1e-6
    cols = []
    for d in range(nd):
        marg = shifted.sum(dim=tuple(a for a in spatial if a != 2 + d))
        marg = marg / marg.sum(dim=-1, keepdim=True)
        cols.append((marg * _axis(heatmap.shape[2 + d], heatmap.device)).sum(dim=-1))
    return torch.stack(cols + [inten], dim=-1)


def extract_gaussian_map_from_keypoints(keypoint, sigma=1.0, G=None):
    """(B, K, 4) keypoints -> (B, K, G, G, G) separable Gaussians of width 2 (sigma/G)^2 times intensity."""
    if isinstance(sigma, (list, tuple)):
        if len(set(float(s) for s in sigma)) != 1:
            raise NotImplementedError("per-keypoint sigmas (fixed_sigma=0) are not implemented")
        sigma = float(sigma[0])
    if keypoint.shape[-1] != 4:
        raise NotImplementedError("only 3-D keypoints (x, y, z, intensity) are implemented")
    B, K = keypoint.shape[:2]
    return ops.gaussian_render(keypoint.reshape(B * K, 1, 4).reshape(B, K, 4), float(sigma), int(G))


# ---------------------------------------------------------------- auxiliary losses
def get_keypoint_sparsity_loss(weight_matrix):
    """(B, T, K, G..) heat-maps -> (B, T): mean over K of |mean heat-map activation|."""
    spatial = tuple(range(3, weight_matrix.dim()))
    return weight_matrix.mean(dim=spatial).abs().mean(dim=2)


def sparsity_loss_from_means(heat_mean):
    """Same quantity from the per-(frame, keypoint) means the head kernel already produced: (B, T, K) -> (B, T)."""
    return heat_mean.abs().mean(dim=2)


def get_temporal_separation_loss(keypoints, sep_sigma):
    """(B, T, K, D+1) -> (B,): keypoints whose trajectories (about their temporal mean) coincide are penalised."""
    xyz = keypoints[..., :-1]
    K = xyz.shape[2]
    motion = xyz - xyz.mean(dim=1, keepdim=True)
    gap = (motion.unsqueeze(3) - motion.unsqueeze(2)).pow(2).sum(-1).mean(dim=1)
    penalty = torch.exp(-gap / (2.0 * sep_sigma ** 2.0)).sum(dim=(1, 2)) - K
    return penalty / (K * (K - 1))


def get_volume_fitting_loss(seq, keypoints, sigmas, vol_fit_type):
    """'chamfer': mean over occupied voxels of the squared distance to the nearest keypoint -> (B, T)."""
    B, T = seq.shape[:2]
    if vol_fit_type == "none":
        return torch.zeros(B, T, device=seq.device)
    if vol_fit_type != "chamfer":
        raise NotImplementedError(f"vol_fit_type={vol_fit_type!r} is not implemented (shipped config uses 'chamfer')")
    G = seq.shape[-1]
    frames = seq.reshape(B * T, G, G, G).float().contiguous()
    kp = keypoints.reshape(B * T, keypoints.shape[2], 4).float().contiguous()
    return ops.chamfer_vol_fit(frames, kp).view(B, T)


def get_graph_consistency_loss(keypoints, affinity, local_const=True, time_const=True, sparsity_const=True,
                               intensity_const=True, ver=0):
    """Graph regularisers on (B, T, K, 4) keypoints and the (n, K, K, 1) affinity: returns
    (local (B,T), time (B,T), sparsity (1,1), intensity (1,1))."""
    dev = keypoints.device
    blank = torch.zeros(1, 1, device=dev)
    local, timec = blank, blank
    if local_const or time_const:
        infl = affinity.amax(dim=0)
        if ver == 2:
            infl = infl + infl.transpose(0, 1)
        infl = infl[None, None]
        xyz = keypoints[..., :3]
        d2 = (xyz.unsqueeze(3) - xyz.unsqueeze(2)).pow(2).sum(-1, keepdim=True)
        weight = infl if ver == 1 else infl * keypoints[..., -1][..., None, None]
        if local_const:
            local = (d2 * weight).mean(dim=(2, 3, 4))
        if time_const:
            timec = ((d2 - d2.mean(dim=1, keepdim=True)).abs() * weight).mean(dim=(2, 3, 4))
    sparse = blank
    if sparsity_const:
        a = affinity.squeeze(-1)
        cross = (a[:, None] * a[None]).pow(2).sum(dim=1, keepdim=True) - a[:, None].pow(4)
        sparse = cross.sum(dim=(0, 1)).mean(dim=(0, 1), keepdim=True)
    return local, timec, sparse, blank


def get_graph_traj_loss(keypoints, affinity, ver=0):
    """Velocity / acceleration direction agreement between affine keypoints -> (1, 1)."""
    infl = affinity.squeeze(-1).amax(dim=0)
    if ver == 2:
        infl = infl + infl.transpose(0, 1)
    infl = infl[None, None]
    vel = keypoints[:, 1:, :, :3] - keypoints[:, :-1, :, :3]
    acc = vel[:, 1:] - vel[:, :-1]

    def disagreement(v):
        return (1 - F.cosine_similarity(v.unsqueeze(3), v.unsqueeze(2), dim=-1, eps=1e-6)) / 2

    wv, wa = infl, infl
    if ver in (0, 2):
        inten = keypoints[..., -1].unsqueeze(-1)
        iv = (inten[:, 1:] + inten[:, :-1]) / 2
        wv = infl * iv
        wa = infl * ((iv[:, 1:] + iv[:, :-1]) / 2)
    total = (disagreement(vel) * wv).mean(dim=(0, 1)) + (disagreement(acc) * wa).mean(dim=(0, 1))
    return total.mean(dim=(0, 1), keepdim=True)
