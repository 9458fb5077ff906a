"""CPU oracle for the Neural Marionette keypoint-detection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``neural_marionette_b200/`` may import
this file: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have something to
check the CUDA path against (and to time on the host cores).

It is a *restatement* of the reference algorithm as plain functions over a
``state_dict`` (a ``{key: fp32 tensor}`` mapping with the reference's key
names) written against numpy / ``torch.nn.functional`` on the
CPU.  The arithmetic itself lives in ATen exactly as it does for the reference
(which is pure Python over ``torch.nn``); what is restated here is the graph.
Each function cites the reference file:line it follows.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against *outputs of the reference itself run in the build
container* — ``oracle/make_golden.py`` imports ``/root/reference``, checks every
function below against the corresponding reference function on seeded inputs
and writes the small fixtures under ``tests/golden/`` that the CPU test-suite
replays (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# ----------------------------------------------------------------------------
# hyper-parameters (reference: pretrained/aist/opt.pickle)
# ----------------------------------------------------------------------------
def default_hparams(**over) -> SimpleNamespace:
    """The subset of opt.pickle the model constructors read
    (model/kypt_detector.py:18-50, model/hsvrnn_bvh.py:14-20,
    model/neural_marionette.py:11,15)."""
    hp = dict(
        grid_size=64, nkeypoints=24, input_dim=3, gaussian_sigma=1.5,
        const_intensity=3, gaussian_cat_type="none", fixed_sigma=1,
        affinity_ver=3, nneighbor=2, keypoints_graph="affinity_params",
        graph_loss_ver=1, vol_fit_type="chamfer", sep_sigma=0.02,
        nhidden_kypt=512, nlatent_kypt=128, transition_type="dl",
        Tcond=5, Ttot=20, is_binarized=1,
        # flags the detector constructor reads but that do not touch the path
        keypoints_detach=0, graph_random_init=0, using_local_const=1,
        using_time_const=1, using_sparsity_const=1, using_intensity_const=1,
        graph_traj_weight=1.0, graph_vol_weight=0.0, affinity_anneal=0,
        state_mode="no_cat", action_mode="pose",
    )
    hp.update(over)
    return SimpleNamespace(**hp)


# ----------------------------------------------------------------------------
# a1 + the step in front of it: utils/dataset_utils.py
# ----------------------------------------------------------------------------
def crop_sequence(seq: np.ndarray, start: int, T: int, sample_rate: int = 1) -> np.ndarray:
    """utils/dataset_utils.py:6-7 — strided slice of the frame axis."""
    return seq[start:start + T * sample_rate:sample_rate]


def episodic_normalization(seq: np.ndarray, scale: float = 1.0, x_trans: float = 0.0,
                           z_trans: float = 0.0) -> np.ndarray:
    """utils/dataset_utils.py:9-19 — clip-global isotropic bbox normalisation to
    [-1, 1).  The arithmetic runs in the input dtype up to the final
    ``+ np.array([x_trans, 0, z_trans])`` which promotes to float64
    (numpy >= 2 scalar rules: ``blen + 1e-5`` stays float32 for float32 input)."""
    lo = np.amin(seq, axis=(0, 1))
    hi = np.amax(seq, axis=(0, 1))
    extent = (hi - lo).max()
    out = ((seq - lo[None, None]) * scale / (extent + 1e-5)) * 2 - 1
    return out + np.array([x_trans, 0, z_trans])


def voxelize(points: np.ndarray, output_shape: Sequence[int]) -> np.ndarray:
    """utils/dataset_utils.py:21-31 (is_binarized=True branch; the other branch
    is dead code).  float64 quotient by ``step + 1e-5``, truncation toward zero,
    idempotent scatter of 1.0.  Returns (1, G, G, G) float32, index [ix, iy, iz]."""
    shape = tuple(int(s) for s in output_shape)
    lo = np.array([-1, -1, -1])
    hi = np.array([1, 1, 1])
    step = (hi - lo) / np.asarray(shape)          # float64
    grid = np.zeros(shape, dtype=np.float32)
    xyz = points[..., :3]
    cell = ((xyz - lo) / (step + 1e-5)).astype(np.int32)
    grid[cell[:, 0], cell[:, 1], cell[:, 2]] = 1.0
    return grid[None]


def voxelize_clip(points: np.ndarray, grid_size: int) -> np.ndarray:
    """Callers' loop (vis_generation.py:19-23, dataset/dataset.py:170-183):
    (T, N, 3) normalised points -> (T, 1, G, G, G) float32."""
    return np.stack([voxelize(points[t], (grid_size,) * 3) for t in range(len(points))], axis=0)


# ----------------------------------------------------------------------------
# a2, a10, a11: utils/kypt_detector_utils.py
# ----------------------------------------------------------------------------
def add_coord_channels(vox: Tensor) -> Tensor:
    """utils/kypt_detector_utils.py:4-26 — append one linspace(-1,1,X_d) channel
    per spatial axis ('ij' meshgrid: channel d varies along axis d)."""
    B = vox.shape[0]
    dims = vox.shape[2:]
    axes = [torch.linspace(-1.0, 1.0, n, device=vox.device) for n in dims]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)
    return torch.cat([vox, mesh[None].expand(B, *mesh.shape)], dim=1)


def keypoints_from_heatmap(hm: Tensor) -> Tensor:
    """utils/kypt_detector_utils.py:28-55 — intensity = mean / (max_k mean + 1e-6);
    coordinate d = <sum-normalised marginal of (hm + 1e-6) along d, linspace>."""
    B, K = hm.shape[:2]
    nd = hm.dim() - 2
    inten = hm.mean(dim=tuple(range(2, 2 + nd)))
    inten = inten / (inten.max(dim=-1, keepdim=True).values + 1e-6)
    out = []
    for d in range(nd):
        lin = torch.linspace(-1.0, 1.0, hm.shape[2 + d], device=hm.device)
        other = tuple(i for i in range(2, 2 + nd) if i != 2 + d)
        marg = (hm + 1e-6).sum(dim=other)                    # (B, K, G_d)
        marg = marg / marg.sum(dim=-1, keepdim=True)
        out.append((marg * lin).sum(dim=-1))
    return torch.cat([torch.stack(out, dim=-1), inten[..., None]], dim=-1)


def gaussian_map(kp: Tensor, sigma: float, G: int) -> Tensor:
    """utils/kypt_detector_utils.py:57-90 — separable Gaussian of width
    2 (sigma/G)^2 in normalised coordinates times the intensity; the product is
    accumulated in axis order ((1*ex)*ey)*ez, then * I."""
    xyz, inten = kp[..., :-1], kp[..., -1]
    B, K, D = xyz.shape
    width = 2.0 * (sigma / G) ** 2.0
    lin = torch.linspace(-1.0, 1.0, G, device=kp.device)
    vol = torch.ones(B, K, *([G] * D), device=kp.device)
    for d in range(D):
        e = (-(lin[None, None] - xyz[:, :, d, None]).pow(2) / width).exp()
        shape = [B, K] + [1] * D
        shape[2 + d] = G
        vol = vol * e.reshape(shape)
    return vol * inten.reshape(B, K, *([1] * D))


def render_gaussians(kp: Tensor, sigma: float, G: int) -> Tensor:
    """The K-loop at model/kypt_detector.py:349-353 / :222-230 (one call per
    keypoint then cat) is the batched call: every keypoint shares sigma when
    fixed_sigma=1 (kypt_detector.py:39-40)."""
    return gaussian_map(kp, sigma, G)


# ----------------------------------------------------------------------------
# a3-a7: modules/vox_modules.py
# ----------------------------------------------------------------------------
def _conv(x, sd, p, stride=1, padding=0):
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _gn(x, sd, p):
    c = x.shape[1]
    return F.group_norm(x, c // 16, sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def basic_block(x, sd, p, k):
    """modules/vox_modules.py:8-19 — conv(k, pad (k-1)//2) -> GN(C//16) -> LeakyReLU(0.01)."""
    return F.leaky_relu(_gn(_conv(x, sd, p + ".block.0", 1, (k - 1) // 2), sd, p + ".block.1"), 0.01)


def res_block(x, sd, p):
    """modules/vox_modules.py:22-47 — [conv3,GN,LReLU,conv3,GN](x) + skip(x); the
    trailing ``F.leaky_relu(., True)`` has slope 1.0, i.e. it is the identity."""
    r = F.leaky_relu(_gn(_conv(x, sd, p + ".res_branch.0", 1, 1), sd, p + ".res_branch.1"), 0.01)
    r = _gn(_conv(r, sd, p + ".res_branch.3", 1, 1), sd, p + ".res_branch.4")
    if (p + ".skip_con.0.weight") in sd:
        s = _gn(_conv(x, sd, p + ".skip_con.0"), sd, p + ".skip_con.1")
    else:
        s = x
    return r + s


def pool_block(x, sd, p):
    """modules/vox_modules.py:49-61 — learned 2x down-sample: conv(k2,s2) -> GN -> LReLU."""
    return F.leaky_relu(_gn(_conv(x, sd, p + ".stride_conv.0", 2, 0), sd, p + ".stride_conv.1"), 0.01)


def upsample_block(x, sd, p, output_padding=0):
    """modules/vox_modules.py:63-75 — ConvTranspose3d(k2,s2) -> GN -> LReLU."""
    y = F.conv_transpose3d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], stride=2,
                           padding=0, output_padding=output_padding)
    return F.leaky_relu(_gn(y, sd, p + ".block.1"), 0.01)


def hourglass(x, sd, p, N):
    """modules/vox_modules.py:78-120 — 3-level hour-glass with additive skips."""
    op = [(N // 4) % 2, (N // 2) % 2, N % 2]
    s1 = res_block(x, sd, p + ".skip_res1")
    x = res_block(pool_block(x, sd, p + ".encoder_pool1"), sd, p + ".encoder_res1")
    s2 = res_block(x, sd, p + ".skip_res2")
    x = res_block(pool_block(x, sd, p + ".encoder_pool2"), sd, p + ".encoder_res2")
    s3 = res_block(x, sd, p + ".skip_res3")
    x = res_block(pool_block(x, sd, p + ".encoder_pool3"), sd, p + ".encoder_res3")
    x = res_block(x, sd, p + ".decoder_res3")
    x = upsample_block(x, sd, p + ".decoder_upsample3", op[0]) + s3
    x = res_block(x, sd, p + ".decoder_res2")
    x = upsample_block(x, sd, p + ".decoder_upsample2", op[1]) + s2
    x = res_block(x, sd, p + ".decoder_res1")
    x = upsample_block(x, sd, p + ".decoder_upsample1", op[2]) + s1
    return x


def feature_net(x, sd, p, grid_size):
    """model/kypt_detector.py:264-272 — Basic(k5) -> Pool -> Res -> Pool -> HG -> Res."""
    x = basic_block(x, sd, p + ".0", 5)
    x = pool_block(x, sd, p + ".1")
    x = res_block(x, sd, p + ".2")
    x = pool_block(x, sd, p + ".3")
    x = hourglass(x, sd, p + ".4", grid_size // 4)
    return res_block(x, sd, p + ".5")


# ----------------------------------------------------------------------------
# a8-a9: VoxToKyptNet (const_intensity == 3, fixed_sigma == 1: the shipped config)
# ----------------------------------------------------------------------------
def vox_to_kypt(seq: Tensor, sd: StateDict, hp, prefix="kypt_detector.vox_to_kypt"):
    """model/kypt_detector.py:299-364.  Returns heatmaps (B,T,K,g,g,g), keypoints
    (B,T,K,4), gaussians (B,T,K,g,g,g), first_feature (B,128,g,g,g)."""
    assert hp.const_intensity == 3 and hp.fixed_sigma, "oracle restates the shipped configuration only"
    B, T = seq.shape[:2]
    K, G = hp.nkeypoints, hp.grid_size
    g = G // 4
    # once per clip (:311-316): spatio-temporal heatmap from the frame mean
    st = feature_net(add_coord_channels(seq.mean(dim=1)), sd,
                     prefix + ".extract_spatio_temporal_features", G)
    prev = F.leaky_relu(_conv(st, sd, prefix + ".extract_spatio_temporal_heatmaps_from_features.0"), 0.01)
    prev = prev.reshape(B * K, 1, g, g, g)
    hms, kps, gss, first_feature = [], [], [], None
    for t in range(T):
        feat = feature_net(add_coord_channels(seq[:, t]), sd, prefix + ".extract_features", G)
        if t == 0:
            first_feature = feat
        hm = F.leaky_relu(_conv(feat, sd, prefix + ".extract_heatmaps_from_features.0"), 0.01)
        hm = hm.view(B * K, 1, g, g, g)
        hm = F.softplus(_conv(torch.cat([hm, prev], dim=1), sd, prefix + ".propagate_heatmaps.0"))
        hm = hm.view(B, K, g, g, g)
        kp = keypoints_from_heatmap(hm)
        hms.append(hm)
        kps.append(kp)
        gss.append(render_gaussians(kp, hp.gaussian_sigma, g))
    return torch.stack(hms, 1), torch.stack(kps, 1), torch.stack(gss, 1), first_feature


# ----------------------------------------------------------------------------
# a12: KyptToVoxNet
# ----------------------------------------------------------------------------
def kypt_to_vox(gaussians: Tensor, first_feature: Tensor, first_frame: Tensor, sd: StateDict, hp,
                prefix="kypt_detector.kypt_to_vox", sharpness=10.0, translation=0.5) -> Tensor:
    """model/kypt_detector.py:388-415 + :417-460 (gaussian_cat_type == 'none')."""
    assert hp.gaussian_cat_type == "none"
    T = gaussians.shape[1]
    d = prefix + ".decode_voxel_from_combined_representation"
    out = []
    for t in range(T):
        x = torch.cat([gaussians[:, t], first_feature, gaussians[:, 0]], dim=1)
        x = F.leaky_relu(_conv(add_coord_channels(x), sd, prefix + ".adjust_combined_representation.0"), 0.01)
        x = F.interpolate(x, scale_factor=2.0, mode="trilinear", align_corners=False)
        x = F.leaky_relu(_gn(_conv(x, sd, d + ".1", 1, 1), sd, d + ".2"), 0.01)
        x = F.leaky_relu(_gn(_conv(x, sd, d + ".4", 1, 1), sd, d + ".5"), 0.01)
        x = F.interpolate(x, scale_factor=2.0, mode="trilinear", align_corners=False)
        x = F.leaky_relu(_gn(_conv(x, sd, d + ".8", 1, 1), sd, d + ".9"), 0.01)
        x = F.leaky_relu(_gn(_conv(x, sd, d + ".11", 1, 1), sd, d + ".12"), 0.01)
        x = _conv(x, sd, d + ".14")
        out.append(torch.sigmoid(sharpness * (torch.tanh(x) + first_frame - translation)))
    return torch.stack(out, dim=1)


# ----------------------------------------------------------------------------
# a13: KyptDetector.forward / get_affinity / decode_from_dyna + the losses
# ----------------------------------------------------------------------------
def get_affinity(sd: StateDict, hp, prefix="kypt_detector") -> Tensor:
    """model/kypt_detector.py:171-211, affinity_ver == 3: row-softmax of the
    (n, K, K-1) parameters re-inserted around a zero diagonal -> (n, K, K, 1)."""
    assert hp.affinity_ver == 3
    w = torch.softmax(sd[prefix + ".affinity_params"], dim=-1)
    n, K, _ = w.shape
    full = torch.zeros(n, K, K)
    col = torch.arange(K - 1)
    for i in range(K):
        dst = col + (col >= i).long()       # skip the diagonal
        full[:, i, dst] = w[:, i, :]
    return full.unsqueeze(-1)


def sparsity_loss(heatmaps: Tensor) -> Tensor:
    """utils/kypt_detector_utils.py:92-103."""
    return heatmaps.mean(dim=(3, 4, 5)).abs().mean(dim=2)


def separation_loss(keypoints: Tensor, sep_sigma: float) -> Tensor:
    """utils/kypt_detector_utils.py:105-133."""
    xyz = keypoints[..., :-1]
    K = xyz.shape[2]
    disp = xyz - xyz.mean(dim=1, keepdim=True)
    diff = (disp[:, :, :, None] - disp[:, :, None]).pow(2).sum(-1).mean(dim=1)
    loss = (-diff / (2.0 * sep_sigma ** 2.0)).exp().sum(dim=(1, 2))
    return (loss - K) / (K * (K - 1))


def chamfer_vol_fit(seq: Tensor, keypoints: Tensor) -> Tensor:
    """utils/kypt_detector_utils.py:135-157 ('chamfer'): mean over occupied voxels
    of the squared distance (normalised coords) to the nearest keypoint."""
    B, T = seq.shape[:2]
    out = []
    for t in range(T):
        obs = add_coord_channels(seq[:, t])[:, None]                     # (B,1,4,X,X,X)
        kp = keypoints[:, t, :, :3][:, :, :, None, None, None]          # (B,K,3,1,1,1)
        dist = (obs[:, :, 1:] - kp).pow(2).sum(dim=2).min(dim=1, keepdim=True).values
        dist = dist * seq[:, t]
        out.append(dist.sum(dim=(1, 2, 3, 4)) / seq[:, t].sum(dim=(1, 2, 3, 4)))
    return torch.stack(out, dim=1)


def graph_consistency_losses(keypoints: Tensor, affinity: Tensor):
    """utils/kypt_detector_utils.py:172-225 with ver == 1 and all four flags on
    (the shipped config).  Returns (local, time, sparsity, intensity)."""
    infl = affinity.max(dim=0).values[None, None]                        # (1,1,K,K,1)
    pos = keypoints[..., :3]
    dist = (pos[:, :, :, None] - pos[:, :, None]).pow(2).sum(dim=-1, keepdim=True)
    local = (dist * infl).mean(dim=(2, 3, 4))
    time_c = ((dist - dist.mean(dim=1, keepdim=True)).abs() * infl).mean(dim=(2, 3, 4))
    a = affinity.squeeze(-1)
    sp = (a[:, None] * a[None]).pow(2).sum(dim=1, keepdim=True) - a[:, None].pow(4)
    sp = sp.sum(dim=(0, 1)).mean(dim=(0, 1), keepdim=True)
    return local, time_c, sp, torch.zeros(1, 1)


def graph_traj_loss(keypoints: Tensor, affinity: Tensor) -> Tensor:
    """utils/kypt_detector_utils.py:228-265 with ver == 1."""
    infl = affinity.squeeze(-1).max(dim=0).values[None, None]
    vel = keypoints[:, 1:, :, :3] - keypoints[:, :-1, :, :3]
    acc = vel[:, 1:] - vel[:, :-1]
    cos = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
    vc = ((-cos(vel[:, :, :, None], vel[:, :, None]) + 1) / 2 * infl).mean(dim=(0, 1))
    ac = ((-cos(acc[:, :, :, None], acc[:, :, None]) + 1) / 2 * infl).mean(dim=(0, 1))
    return (vc + ac).mean(dim=(0, 1), keepdim=True)


def detector_forward(seq: Tensor, sd: StateDict, hp, affinity_started: bool = True) -> dict:
    """model/kypt_detector.py:81-169 — the dict the reference returns (same keys).
    ``affinity_started`` mirrors ``self.affinity_start`` (set by ``anneal``,
    :71-78); every caller anneals once before the first forward."""
    B, T = seq.shape[:2]
    heatmaps, keypoints, gaussians, first_feature = vox_to_kypt(seq, sd, hp)
    recon = kypt_to_vox(gaussians, first_feature, seq[:, 0], sd, hp)
    recon_loss = F.binary_cross_entropy(recon, seq, reduction="none").mean(dim=(2, 3, 4, 5))
    zeros = torch.zeros(B, T)
    if hp.vol_fit_type == "chamfer":
        vol = chamfer_vol_fit(seq, keypoints)
    else:
        vol = zeros
    if hp.keypoints_graph == "none" or not affinity_started:
        affinity = None
        local = time_c = sp = inten = traj = zeros
    else:
        affinity = get_affinity(sd, hp)
        local, time_c, sp, inten = graph_consistency_losses(keypoints, affinity)
        traj = graph_traj_loss(keypoints, affinity) if hp.graph_traj_weight > 0 else zeros
    return dict(
        recon=recon, keypoints=keypoints, heatmaps=heatmaps, affinity=affinity,
        recon_loss=recon_loss.mean(), vol_fit_reg=vol.mean(), kypt_const_loss=zeros.mean(),
        separation_loss=separation_loss(keypoints, hp.sep_sigma).mean(),
        sparsity_loss=sparsity_loss(heatmaps).mean(),
        local_const_loss=local.mean(), time_const_loss=time_c.mean(),
        sparsity_const_loss=sp.mean(), intensity_const_loss=inten.mean(),
        graph_traj_loss=traj.mean(), graph_vol_loss=zeros.mean(),
        first_feature=first_feature, gaussians=gaussians,
    )


def decode_from_dyna(keypoints: Tensor, first_feature: Tensor, first_frame: Tensor, sd, hp) -> Tensor:
    """model/kypt_detector.py:213-241."""
    g = hp.grid_size // 4
    gs = torch.stack([render_gaussians(keypoints[:, t], hp.gaussian_sigma, g)
                      for t in range(keypoints.shape[1])], dim=1)
    return kypt_to_vox(gs, first_feature, first_frame, sd, hp)


# ----------------------------------------------------------------------------
# skeleton extraction: utils/dyna_utils.py:6-171 (host-side, one-off, K small).
# The reference leans on networkx for all-pairs shortest paths; restated here
# with a plain Dijkstra so the oracle has no dependency beyond numpy.
# ----------------------------------------------------------------------------
def _all_pairs(K: int, edges: Dict[Tuple[int, int], float], big: float) -> np.ndarray:
    """All-pairs shortest path lengths on an undirected weighted graph;
    unreachable pairs keep ``big`` (dyna_utils.py:21-35 and its repeats)."""
    import heapq
    adj: List[List[Tuple[int, float]]] = [[] for _ in range(K)]
    for (a, b), w in edges.items():
        adj[a].append((b, w))
    out = np.ones((K, K)) * big
    for s in range(K):
        dist = {s: 0.0}
        heap = [(0.0, s)]
        done = set()
        while heap:
            d, u = heapq.heappop(heap)
            if u in done:
                continue
            done.add(u)
            for v, w in adj[u]:
                nd = d + w
                if v not in dist or nd < dist[v]:
                    dist[v] = nd
                    heapq.heappush(heap, (nd, v))
        for v, d in dist.items():
            out[s, v] = d
    return out


def _edges(mask: np.ndarray, weight: Optional[np.ndarray] = None) -> Dict[Tuple[int, int], float]:
    e = {}
    for a, b in np.stack(np.where(mask), axis=-1):
        w = 1.0 if weight is None else float(weight[a, b])
        # an undirected graph keeps one weight per edge: the last one added wins
        e[(int(a), int(b))] = w
        e[(int(b), int(a))] = w
    return e


def _n_components(K: int, edges) -> int:
    parent = list(range(K))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for (a, b) in edges:
        parent[find(a)] = find(b)
    return len({find(i) for i in range(K)})


def skeleton_from_affinity(affinity: Tensor, big: float = 1e4):
    """utils/dyna_utils.py:6-171 ``process_affinity_glob``: affinity (n,K,K,1) ->
    (A (K,K) float64 tree adjacency, priority (values, indices) ascending tree
    distance from the root, parents (K,) int64)."""
    n, K = affinity.shape[:2]
    infl_t = affinity.max(dim=0).values.squeeze(-1)
    infl = infl_t.detach().cpu().numpy()
    top = infl_t.topk(n, dim=-1).indices.cpu().numpy()
    adj = np.zeros((K, K), dtype=np.float32)
    adj[np.arange(K)[:, None], top] = 1
    adj = np.maximum(adj, adj.T)

    dij = _all_pairs(K, _edges(adj), big)
    if _n_components(K, _edges(adj)) > 1:                     # :38-66
        tot = dij.sum(axis=-1)
        root = tot.argmin()
        order = tot.copy().argsort()
        rank = np.zeros(K)
        rank[order] = np.arange(K)
        cand = np.where(dij[root] == big)[0]
        pick = cand[0]
        for c in cand[1:]:
            if rank[pick] > rank[c]:
                pick = c
        adj[root, pick] = 1
        adj[pick, root] = 1
        dij = _all_pairs(K, _edges(adj), big)

    # break ties between equal total-distance nodes with 1e-5 edge bumps (:68-81)
    tot = dij.sum(axis=-1)
    wadj = adj.copy()
    for k in range(K - 1):
        for q in range(k + 1, K):
            if tot[k] == tot[q]:
                qs = np.where(adj[q])[0]
                for m in np.where(adj[k])[0]:
                    if m in qs:
                        l = q if infl[m, k] > infl[m, q] else k
                        wadj[m, l] += 1e-5
                        wadj[l, m] += 1e-5
    dij = torch.from_numpy(_all_pairs(K, _edges(adj, wadj), big))

    root = dij.sum(dim=-1).topk(K, dim=-1, largest=False).indices[0]      # :100-102
    priority = dij[root].topk(K, largest=False)
    rank = dij[root]
    parents = []
    for k in range(K):                                                    # :105-140
        if k == root:
            parents.append(k)
            continue
        nbrs = np.where(adj[k])[0]
        best, best_d = None, -1e3
        for m in nbrs:
            dd = rank[m] - rank[k]
            if dd < 0 and dd > best_d:
                best, best_d = m, dd
            elif dd < 0 and dd == best_d:
                if infl[k, m] > infl[k, best]:
                    best, best_d = m, dd
            elif dd == 0:
                co, co_rank = None, 1e4
                for mm in np.where(adj[m])[0]:
                    if mm in nbrs and rank[mm] < rank[m] and co_rank > rank[mm]:
                        co, co_rank = mm, rank[mm]
                if co is not None and infl[co, m] > infl[co, k]:
                    best, best_d = m, dd
        if best is None:
            best = priority.indices[0]
            adj[k, best] = 1
            adj[best, k] = 1
        parents.append(int(best))
    parents = torch.LongTensor(parents)

    A = torch.zeros_like(dij)
    for k in range(K):
        if k != parents[k]:
            A[k, parents[k]] = 1
            A[parents[k], k] = 1
    dij = torch.from_numpy(_all_pairs(K, _edges(A.numpy(), wadj), big))   # :151-168
    priority = dij[root].topk(K, dim=-1, largest=False)
    return A, priority, parents


# ----------------------------------------------------------------------------
# a14-a15: HSVRNNBVH (model/hsvrnn_bvh.py, utils/geo_utils.py)
# ----------------------------------------------------------------------------
def _mlp(x, sd, p):
    """Linear -> LeakyReLU(0.01) -> Linear (hsvrnn_bvh.py:29-54)."""
    h = F.leaky_relu(F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"]), 0.01)
    return F.linear(h, sd[p + ".2.weight"], sd[p + ".2.bias"])


def rot6d_to_matrix(p: Tensor) -> Tensor:
    """utils/geo_utils.py:30-78 — Gram-Schmidt on two 3-vectors; columns (x, y, z)."""
    lead = p.shape[:-1]
    p = p.reshape(-1, 6)
    a, b = p[:, 0:3], p[:, 3:6]

    def unit(v):
        return v / (torch.sqrt(v.pow(2).sum(1)) + 1e-10).reshape(-1, 1)

    def cross(u, v):
        return torch.stack((u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1],
                            u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2],
                            u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]), dim=1)
    x = unit(a)
    z = unit(cross(x, b))
    y = cross(z, x)
    return torch.stack((x, y, z), dim=2).reshape(*lead, 3, 3)


def gru_cell(x: Tensor, h: Tensor, sd: StateDict, p: str) -> Tensor:
    """torch.nn.GRUCell semantics, gate order r,z,n (hsvrnn_bvh.py:57-58)."""
    gi = F.linear(x, sd[p + ".weight_ih"], sd[p + ".bias_ih"])
    gh = F.linear(h, sd[p + ".weight_hh"], sd[p + ".bias_hh"])
    ir, iz, inn = gi.chunk(3, dim=-1)
    hr, hz, hn = gh.chunk(3, dim=-1)
    r = torch.sigmoid(ir + hr)
    z = torch.sigmoid(iz + hz)
    n = torch.tanh(inn + r * hn)
    return (1 - z) * n + z * h


def bone_offsets(keypoints: Tensor, parents: Tensor, sd: StateDict, prefix="dyna_module") -> Tensor:
    """hsvrnn_bvh.py:236-253 — (lower) median over T of the keypoint-parent
    distance times the unit offset_param direction -> (B, K, 3, 1)."""
    pos = keypoints[..., :3]
    dist = (pos[:, :, :, None] - pos[:, :, None]).pow(2).sum(dim=-1).sqrt()
    med = dist.median(dim=1).values
    K = pos.shape[2]
    scale = torch.stack([med[:, k, parents[k]] for k in range(K)], dim=-1)
    op = sd[prefix + ".offset_param"]
    unit = op / (op.pow(2).sum(dim=-1, keepdim=True).sqrt() + 1e-10)
    return (unit[None] * scale[..., None])[..., None]


def decode_pose(dec_in: Tensor, offset: Tensor, order: Tensor, parents: Tensor, sd: StateDict,
                K: int, prefix="dyna_module") -> Tuple[Tensor, Tensor]:
    """hsvrnn_bvh.py:255-286 + geo_utils.py:3-27 — root/intensity MLP (tanh),
    6-D rotations, global rotations down the tree in ``order`` (priority.indices),
    forward kinematics.  Returns (flat (B, 4K), R (B, K, 3, 3))."""
    B = dec_in.shape[0]
    raw = torch.tanh(_mlp(dec_in, sd, prefix + ".root_intensity_decoder"))
    root_pos = raw[:, :3]
    inten = ((raw[:, 3:] + 1) * 0.5).unsqueeze(-1)
    Rloc = rot6d_to_matrix(_mlp(dec_in, sd, prefix + ".joint_matrix_decoder").reshape(B, K, 6))
    root = int(order[0])
    Rg: Dict[int, Tensor] = {root: Rloc[:, root]}
    for i in order[1:].tolist():
        Rg[i] = torch.bmm(Rg[int(parents[i])], Rloc[:, i])
    pos = torch.zeros(B, K, 3)
    pos[:, root] = root_pos
    for i in order[1:].tolist():
        pos[:, i] = torch.bmm(Rg[i], offset[:, i]).squeeze(-1) + pos[:, int(parents[i])]
    R = torch.stack([Rg[i] for i in range(K)], dim=1)
    return torch.cat([pos, inten], dim=-1).reshape(B, -1), R


def _softplus_std(raw):
    return F.softplus(raw) + 1e-4


def posterior_step(h, kp_flat, eps, offset, order, parents, sd, K, prefix="dyna_module"):
    """One conditioned step (hsvrnn_bvh.py:98-128 / :175-199): posterior MLP,
    ``z_i = mean + std * eps_i`` for the S injected ``eps`` (rsample), decode all
    S, keep the one nearest to the detected keypoints, GRU update."""
    post = _mlp(torch.cat([h, kp_flat], dim=-1), sd, prefix + ".extract_post_dist")
    mean, std = post.chunk(2, dim=-1)
    std = _softplus_std(std)
    z = mean[None] + std[None] * eps                       # (S, B, Z)
    flats, Rs = [], []
    for i in range(eps.shape[0]):
        f, R = decode_pose(torch.cat([h, z[i]], dim=-1), offset, order, parents, sd, K, prefix)
        flats.append(f)
        Rs.append(R)
    flats = torch.stack(flats, 0)
    pick = (kp_flat[None] - flats).pow(2).sum(-1).argmin(dim=0)
    bi = torch.arange(h.shape[0])
    best_z, best_f = z[pick, bi], flats[pick, bi]
    best_R = torch.stack(Rs, 0)[pick, bi]
    h_new = gru_cell(torch.cat([best_f, best_z], dim=-1), h, sd, prefix + ".kypt_rnn_cell")
    return h_new, best_f, best_z, best_R, (mean, std)


def prior_params(h, sd, prefix="dyna_module"):
    mean, std = _mlp(h, sd, prefix + ".extract_prior_dist").chunk(2, dim=-1)
    return mean, _softplus_std(std)


def prior_step(h, eps, offset, order, parents, sd, K, prefix="dyna_module"):
    """One generated step (hsvrnn_bvh.py:208-225)."""
    mean, std = prior_params(h, sd, prefix)
    z = mean + std * eps
    f, _ = decode_pose(torch.cat([h, z], dim=-1), offset, order, parents, sd, K, prefix)
    h_new = gru_cell(torch.cat([f, z], dim=-1), h, sd, prefix + ".kypt_rnn_cell")
    return h_new, f, z


def kl_normal(m1, s1, m2, s2):
    """torch.distributions.kl.kl_divergence(Normal(m1,s1), Normal(m2,s2))."""
    var_ratio = (s1 / s2).pow(2)
    t1 = ((m1 - m2) / s2).pow(2)
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def dyna_encode(keypoints, skeleton, sd, hp, eps=None, S=10, prefix="dyna_module") -> dict:
    """hsvrnn_bvh.py:67-156.  ``eps`` (T, S, B, Z) replaces the draws of
    ``Normal.rsample`` (which are ``torch.normal(0, 1)`` of shape (S, B, Z), one
    call per step); when None they are drawn here in the same order."""
    B, T, K, _ = keypoints.shape
    _, priority, parents = skeleton
    order = priority.indices
    h = sd[prefix + ".init_kypt_rnn_state"].expand(B, -1)
    offset = bone_offsets(keypoints, parents, sd, prefix)
    hs, zs, kps, Rs, kls = [h], [], [], [], []
    for t in range(T):
        pm, ps = prior_params(h, sd, prefix)
        e = eps[t] if eps is not None else torch.normal(torch.zeros(S, B, hp.nlatent_kypt),
                                                        torch.ones(S, B, hp.nlatent_kypt))
        h, f, z, R, (qm, qs) = posterior_step(h, keypoints[:, t].reshape(B, -1), e, offset, order,
                                              parents, sd, K, prefix)
        kls.append(kl_normal(qm, qs, pm, ps))
        hs.append(h), zs.append(z), kps.append(f.view(B, K, -1)), Rs.append(R)
    kp_inf = torch.stack(kps, 1)
    return dict(kypt_recon=kp_inf[..., :4], R=torch.stack(Rs, 1), z_kypts=torch.stack(zs, 1),
                h_kypts=torch.stack(hs, 1), kl_kypt=torch.stack(kls, 1).mean(),
                kypt_recon_loss=(kp_inf - keypoints).pow(2).sum(dim=(2, 3)).mean())


def dyna_generate(keypoints_cond, skeleton, sd, hp, Ttot, Tcond, eps_cond=None, eps_gen=None, S=10,
                  prefix="dyna_module") -> dict:
    """hsvrnn_bvh.py:158-234.  eps_cond (Tcond, S, B, Z), eps_gen (Ttot-Tcond, B, Z)."""
    B, _, K, _ = keypoints_cond.shape
    _, priority, parents = skeleton
    order = priority.indices
    Z = hp.nlatent_kypt
    h = sd[prefix + ".init_kypt_rnn_state"].expand(B, -1)
    offset = bone_offsets(keypoints_cond, parents, sd, prefix)
    cond, gen = [], []
    for t in range(Tcond):
        e = eps_cond[t] if eps_cond is not None else torch.normal(torch.zeros(S, B, Z), torch.ones(S, B, Z))
        h, f, _, _, _ = posterior_step(h, keypoints_cond[:, t].reshape(B, -1), e, offset, order,
                                       parents, sd, K, prefix)
        cond.append(f.view(B, K, -1))
    for t in range(Tcond, Ttot):
        e = eps_gen[t - Tcond] if eps_gen is not None else torch.normal(torch.zeros(B, Z), torch.ones(B, Z))
        h, f, _ = prior_step(h, e, offset, order, parents, sd, K, prefix)
        gen.append(f.view(B, K, -1))
    return dict(keypoints_cond=torch.stack(cond, 1)[..., :4], keypoints_gen=torch.stack(gen, 1)[..., :4],
                h_last=h)


def dyna_interpolate(keypoints, skeleton, sd, hp, sample_num, sample_rate, eps, prefix="dyna_module"):
    """vis_interpolation.py:90-137 — key-frame interpolation with ``sample_num`` hypotheses: at key frames
    (``t % sample_rate == 0`` or the last frame) every hypothesis draws one posterior sample, the one nearest to the
    detected keypoints survives (state, latent and pose collapse onto it) and the pending in-between poses of the
    hypothesis whose *prior* sample lands nearest to it are emitted; between key frames the prior is rolled out.
    keypoints (1, T, K, 4); eps (T, 2, sample_num, Z) replaces the rsample draws in call order ([t, 0]: posterior at
    key frames / prior otherwise, [t, 1]: the prior draw "for choosing").  Returns (selected (1, T, K, 4), picks)."""
    _, T, K, _ = keypoints.shape
    _, priority, parents = skeleton
    order = priority.indices
    h = sd[prefix + ".init_kypt_rnn_state"].expand(sample_num, -1)                           # :90
    offset = bone_offsets(keypoints, parents, sd, prefix).expand(sample_num, -1, -1, -1)     # :91
    selected, pending, picks = [], [], []
    for t in range(T):
        kp_flat = keypoints[:, t].reshape(1, -1).expand(sample_num, -1)                      # :96-97
        if t % sample_rate == 0 or t == T - 1:
            qm, qs = _mlp(torch.cat([h, kp_flat], dim=-1), sd, prefix + ".extract_post_dist").chunk(2, dim=-1)
            qs = _softplus_std(qs)
            pm, ps = prior_params(h, sd, prefix)
            z = qm + qs * eps[t, 0]                                                          # :106
            zc = pm + ps * eps[t, 1]                                                         # :108
            f, _ = decode_pose(torch.cat([h, z], dim=-1), offset, order, parents, sd, K, prefix)
            fc, _ = decode_pose(torch.cat([h, zc], dim=-1), offset, order, parents, sd, K, prefix)
            i = (f - kp_flat).pow(2).sum(dim=-1).argmin()                                    # :111-112
            f = f[i][None].expand(sample_num, -1)
            z = z[i][None].expand(sample_num, -1)
            h = h[i][None].expand(sample_num, -1)
            j = (fc - f).pow(2).sum(dim=-1).argmin()                                         # :116-117
            pending.append(kp_flat)
            selected += [s[j].view(K, 4) for s in pending]                                   # :118-120
            pending = []
            picks.append((int(i), int(j)))
        else:
            pm, ps = prior_params(h, sd, prefix)
            z = pm + ps * eps[t, 0]
            f, _ = decode_pose(torch.cat([h, z], dim=-1), offset, order, parents, sd, K, prefix)
            pending.append(f)                                                                # :129
        h = gru_cell(torch.cat([f, z], dim=-1), h, sd, prefix + ".kypt_rnn_cell")            # :131-132
    sel = torch.stack(selected, dim=0)[None].clone()
    sel[0, :, :, -1] = sel[0, 0, :, -1]                                                      # :135
    return sel, picks


# ----------------------------------------------------------------------------
# C11 façade: NeuralMarionette.forward / generate
# ----------------------------------------------------------------------------
def marionette_generate(vox_seq, sd, hp, skeleton=None, eps_cond=None, eps_gen=None) -> dict:
    """model/neural_marionette.py:58-103 (transition_type == 'dl')."""
    T = vox_seq.shape[1]
    det = detector_forward(vox_seq[:, :hp.Tcond].contiguous(), sd, hp)
    if skeleton is None:
        skeleton = skeleton_from_affinity(det["affinity"])
    dyn = dyna_generate(det["keypoints"], skeleton, sd, hp, Ttot=T, Tcond=hp.Tcond,
                        eps_cond=eps_cond, eps_gen=eps_gen)
    gen = decode_from_dyna(dyn["keypoints_gen"], det["first_feature"], vox_seq[:, 0], sd, hp)
    return dict(gen=torch.cat([det["recon"][:, :hp.Tcond], gen], dim=1),
                keypoints=torch.cat([det["keypoints"][:, :hp.Tcond], dyn["keypoints_gen"]], dim=1),
                A_hats=None)


# ----------------------------------------------------------------------------
# synthetic inputs / weights shared by tests, smoke and bench
# ----------------------------------------------------------------------------
def synthetic_clip(seed: int, T: int, N: int) -> np.ndarray:
    """Seeded anisotropic blob with rigidly moving limb clusters, (T, N, 3) float32."""
    rng = np.random.default_rng(seed)
    base = (rng.standard_normal((N, 3)) * np.array([0.3, 0.9, 0.2])).astype(np.float32)
    limb = rng.integers(0, 5, size=N)
    axis = rng.standard_normal((5, 3)).astype(np.float32)
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    piv = (rng.standard_normal((5, 3)) * 0.3).astype(np.float32)
    out = np.empty((T, N, 3), dtype=np.float32)
    for t in range(T):
        ang = np.float32(0.05 * t)
        c, s = np.cos(ang), np.sin(ang)
        p = base.copy()
        for j in range(1, 5):
            m = limb == j
            k = axis[j]
            v = p[m] - piv[j]
            # Rodrigues rotation about axis k through pivot piv[j]
            v = v * c + np.cross(k, v) * s + np.outer(v @ k, k) * (1 - c)
            p[m] = v + piv[j]
        out[t] = p + np.float32(0.02 * t) * np.array([1.0, 0.0, 0.5], dtype=np.float32)
    return out


def synthetic_state_dict(hp, seed: int = 0, peaked: bool = True) -> StateDict:
    """Seeded synthetic checkpoint with the reference's exact key/shape layout
    (337 tensors at the shipped hyper-parameters).  Built from an
    explicit shape table so it needs neither the reference nor the product
    package.  ``peaked`` scales the heatmap head so the heatmaps are sharply
    localised (SURVEY.md §7 'random-init weights make weak parity tests')."""
    gen = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    K, Z, H = hp.nkeypoints, hp.nlatent_kypt, hp.nhidden_kypt

    def conv(p, co, ci, k, transposed=False):
        fan_in = ci * k ** 3
        bound = 1.0 / math.sqrt(fan_in)
        shape = (ci, co, k, k, k) if transposed else (co, ci, k, k, k)
        sd[p + ".weight"] = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        sd[p + ".bias"] = (torch.rand(co, generator=gen) * 2 - 1) * bound

    def norm(p, c):
        sd[p + ".weight"] = 1.0 + 0.2 * torch.randn(c, generator=gen)
        sd[p + ".bias"] = 0.2 * torch.randn(c, generator=gen)

    def res(p, ci, co):
        conv(p + ".res_branch.0", co, ci, 3), norm(p + ".res_branch.1", co)
        conv(p + ".res_branch.3", co, co, 3), norm(p + ".res_branch.4", co)
        if ci != co:
            conv(p + ".skip_con.0", co, ci, 1), norm(p + ".skip_con.1", co)

    def pool(p, c):
        conv(p + ".stride_conv.0", c, c, 2), norm(p + ".stride_conv.1", c)

    def up(p, ci, co):
        conv(p + ".block.0", co, ci, 2, transposed=True), norm(p + ".block.1", co)

    def hg(p, ci, co):
        pool(p + ".encoder_pool1", ci), res(p + ".encoder_res1", ci, 32)
        pool(p + ".encoder_pool2", 32), res(p + ".encoder_res2", 32, 48)
        pool(p + ".encoder_pool3", 48), res(p + ".encoder_res3", 48, 72)
        res(p + ".decoder_res3", 72, 72), up(p + ".decoder_upsample3", 72, 48)
        res(p + ".decoder_res2", 48, 48), up(p + ".decoder_upsample2", 48, 32)
        res(p + ".decoder_res1", 32, 32), up(p + ".decoder_upsample1", 32, co)
        res(p + ".skip_res1", ci, co), res(p + ".skip_res2", 32, 32), res(p + ".skip_res3", 48, 48)

    def fnet(p, cin, cout):
        conv(p + ".0.block.0", cout // 4, 1 + cin, 5), norm(p + ".0.block.1", cout // 4)
        pool(p + ".1", cout // 4)
        res(p + ".2", cout // 4, cout // 2)
        pool(p + ".3", cout // 2)
        hg(p + ".4", cout // 2, cout // 2)
        res(p + ".5", cout // 2, cout)

    def lin(p, co, ci):
        bound = 1.0 / math.sqrt(ci)
        sd[p + ".weight"] = (torch.rand(co, ci, generator=gen) * 2 - 1) * bound
        sd[p + ".bias"] = (torch.rand(co, generator=gen) * 2 - 1) * bound

    kd = "kypt_detector"
    # non-uniform affinity so the skeleton is a deep tree, not the all-ones default
    sd[kd + ".affinity_params"] = torch.randn(hp.nneighbor, K, K - 1, generator=gen) * 2.0
    v = kd + ".vox_to_kypt"
    fnet(v + ".extract_features", hp.input_dim, 128)
    conv(v + ".extract_heatmaps_from_features.0", K, 128, 1)
    fnet(v + ".extract_spatio_temporal_features", hp.input_dim, 256)
    conv(v + ".extract_spatio_temporal_heatmaps_from_features.0", K, 256, 1)
    conv(v + ".propagate_heatmaps.0", 1, 2, 1)
    if peaked:
        sd[v + ".extract_heatmaps_from_features.0.weight"] *= 12.0
        sd[v + ".extract_spatio_temporal_heatmaps_from_features.0.weight"] *= 6.0
        sd[v + ".propagate_heatmaps.0.weight"] = torch.tensor([1.5, 0.75]).reshape(1, 2, 1, 1, 1)
        sd[v + ".propagate_heatmaps.0.bias"] = torch.tensor([-2.0])
    d = kd + ".kypt_to_vox"
    conv(d + ".adjust_combined_representation.0", 128, 128 + 2 * K + hp.input_dim, 1)
    dd = d + ".decode_voxel_from_combined_representation"
    conv(dd + ".1", 64, 128, 3), norm(dd + ".2", 64)
    conv(dd + ".4", 64, 64, 3), norm(dd + ".5", 64)
    conv(dd + ".8", 32, 64, 3), norm(dd + ".9", 32)
    conv(dd + ".11", 32, 32, 3), norm(dd + ".12", 32)
    conv(dd + ".14", 1, 32, 1)
    m = "dyna_module"
    sd[m + ".init_kypt_rnn_state"] = torch.randn(1, H, generator=gen)
    sd[m + ".offset_param"] = torch.randn(K, 3, generator=gen)
    S = K * (hp.input_dim + 1)
    lin(m + ".extract_post_dist.0", 128, H + S), lin(m + ".extract_post_dist.2", 2 * Z, 128)
    lin(m + ".extract_prior_dist.0", 128, H), lin(m + ".extract_prior_dist.2", 2 * Z, 128)
    lin(m + ".root_intensity_decoder.0", 128, H + Z), lin(m + ".root_intensity_decoder.2", 3 + K, 128)
    lin(m + ".joint_matrix_decoder.0", 128, H + Z), lin(m + ".joint_matrix_decoder.2", 6 * K, 128)
    b = 1.0 / math.sqrt(H)
    sd[m + ".kypt_rnn_cell.weight_ih"] = (torch.rand(3 * H, S + Z, generator=gen) * 2 - 1) * b
    sd[m + ".kypt_rnn_cell.weight_hh"] = (torch.rand(3 * H, H, generator=gen) * 2 - 1) * b
    sd[m + ".kypt_rnn_cell.bias_ih"] = (torch.rand(3 * H, generator=gen) * 2 - 1) * b
    sd[m + ".kypt_rnn_cell.bias_hh"] = (torch.rand(3 * H, generator=gen) * 2 - 1) * b
    return sd
