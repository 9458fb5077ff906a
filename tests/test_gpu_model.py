"""End-to-end GPU parity: the drop-in modules against the oracle / the reference's golden outputs.

north_star tolerances: keypoints within 1e-3 of the grid extent (= 2e-3 absolute), heat-maps within 1e-2
relative (to the heat-map peak), occupancy bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import nm_oracle as O

pytestmark = pytest.mark.gpu

KP_TOL = 2e-3        # 1e-3 of the grid extent [-1, 1]
HM_TOL = 1e-2        # relative to the heat-map peak


def build(hp, seed):
    import neural_marionette_b200 as nm
    sd = O.synthetic_state_dict(hp, seed=seed)
    net = nm.NeuralMarionette(hp)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net = net.cuda().eval()
    net.anneal(1)
    return net, sd


def clips(seed, B, T, N, G):
    import neural_marionette_b200 as nm
    raw = np.stack([O.synthetic_clip(seed + b, T, N) for b in range(B)], 0)
    vox = nm.voxelize_raw_clips(raw, G)
    ref = np.stack([O.voxelize_clip(O.episodic_normalization(raw[b]), G) for b in range(B)], 0)
    assert np.array_equal(vox.cpu().numpy(), ref)
    return vox, torch.from_numpy(ref)


@pytest.mark.parametrize("tag", ["g32", "g64"])
def test_detector_vs_reference_golden(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"detector_{tag}.npz"))
    G, B, T = int(z["G"]), int(z["B"]), int(z["T"])
    hp = O.default_hparams(grid_size=G)
    net, sd = build(hp, int(z["seed"]))
    vox, _ = clips(int(z["vox_seed"]), B, T, 20000, G)
    with torch.no_grad():
        out = net.kypt_detector(vox)
        gen = net.kypt_detector.decode_from_dyna(torch.from_numpy(z["gen_kp"]).cuda(), out["first_feature"],
                                                 vox[:, 0])["gen"]
    assert set(out.keys()) == {"recon", "keypoints", "heatmaps", "affinity", "recon_loss", "vol_fit_reg",
                               "kypt_const_loss", "separation_loss", "sparsity_loss", "local_const_loss",
                               "time_const_loss", "sparsity_const_loss", "intensity_const_loss", "graph_traj_loss",
                               "graph_vol_loss", "first_feature"}
    kp_err = np.abs(out["keypoints"].cpu().numpy() - z["keypoints"]).max()
    hm_ref = z["heatmaps_t0"]
    hm_err = np.abs(out["heatmaps"][:, :1].cpu().numpy() - hm_ref).max() / hm_ref.max()
    print(f"[{tag}] keypoint max err {kp_err:.3e}  heat-map rel err {hm_err:.3e}")
    assert kp_err <= KP_TOL
    assert hm_err <= HM_TOL
    ff = out["first_feature"][:, ::8].cpu().numpy()
    ff_ref = z["first_feature_sub_f16"].astype(np.float32)
    assert np.abs(ff - ff_ref).max() <= 2e-2 * np.abs(ff_ref).max()
    rec = out["recon"][..., ::2, ::2, ::2].cpu().numpy()
    assert np.abs(rec - z["recon_sub_f16"].astype(np.float32)).mean() <= 5e-3
    assert np.abs(gen[..., ::2, ::2, ::2].cpu().numpy() - z["gen_sub_f16"].astype(np.float32)).mean() <= 5e-3
    np.testing.assert_allclose(out["affinity"].cpu().numpy(), z["affinity"], atol=1e-6)
    got = np.array([float(out[k]) for k in ["recon_loss", "vol_fit_reg", "separation_loss", "sparsity_loss",
                                            "local_const_loss", "time_const_loss", "sparsity_const_loss",
                                            "graph_traj_loss"]])
    np.testing.assert_allclose(got, z["losses"], rtol=2e-2, atol=1e-5)


def test_detector_batched_vs_oracle_and_chunking():
    """B*T larger than one frame chunk; VoxToKyptNet / KyptToVoxNet public forwards; batch independence."""
    import neural_marionette_b200.model.kypt_detector as kd
    G, B, T = 32, 3, 4
    hp = O.default_hparams(grid_size=G)
    net, sd = build(hp, 41)
    vox, vox_ref = clips(5000, B, T, 20000, G)
    old = kd.FRAME_CHUNK
    try:
        kd.FRAME_CHUNK = 0          # forces clip-sized chunks (one clip per pass)
        with torch.no_grad():
            hm, kp, gs, ff = net.kypt_detector.vox_to_kypt(vox)
            rec = net.kypt_detector.kypt_to_vox(gs, ff, vox[:, 0])
        kd.FRAME_CHUNK = 64
        with torch.no_grad():
            out = net.kypt_detector(vox)
    finally:
        kd.FRAME_CHUNK = old
    with torch.no_grad():
        ref = O.detector_forward(vox_ref, sd, hp)
    assert (kp.cpu() - ref["keypoints"]).abs().max() <= KP_TOL
    assert float((hm.cpu() - ref["heatmaps"]).abs().max() / ref["heatmaps"].max()) <= HM_TOL
    assert float((gs.cpu() - ref["gaussians"]).abs().max()) <= 2e-2
    assert (rec.cpu() - ref["recon"]).abs().mean() <= 5e-3
    # chunking must not change results (same kernels, same per-sample arithmetic)
    assert torch.equal(out["keypoints"], kp)
    assert (out["recon"] - rec).abs().max() <= 2e-2      # decoder fed keypoints vs gaussians: fp32 re-render only
    # outputs are ordinary writable tensors (callers threshold them in place, vis_generation.py:138-139)
    out["recon"][out["recon"] < 0.5] = 0


def test_detector_is_bit_reproducible():
    """Two runs on the same input give identical bits (no float atomics, fixed reduction orders, fixed MMA order):
    the reference's callers set cudnn.deterministic = True and compare runs."""
    G, B, T = 64, 2, 3
    hp = O.default_hparams(grid_size=G)
    net, _ = build(hp, 43)
    vox, _ = clips(6000, B, T, 20000, G)
    with torch.no_grad():
        a = net.kypt_detector(vox)
        b = net.kypt_detector(vox)
    for key in ("heatmaps", "keypoints", "recon", "first_feature", "affinity", "vol_fit_reg", "sparsity_loss"):
        assert torch.equal(a[key], b[key]), key
    assert float(a["recon_loss"]) == float(b["recon_loss"])


def test_generate_vs_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "generate_g32.npz"))
    hp = O.default_hparams(grid_size=32)
    net, sd = build(hp, int(z["seed"]))
    T = int(z["T"])
    vox, _ = clips(int(z["vox_seed"]), 1, T, 20000, 32)
    act = {"detector": True, "learner": True}
    with torch.no_grad():
        with pytest.raises(TypeError):
            net.generate(vox, act)               # skeleton not built yet: same failure mode as the reference
        log = net(vox, act)
        assert {"kypt_recon", "R", "z_kypts", "h_kypts", "kl_kypt", "kypt_recon_loss"} <= set(log.keys())
        out = net.generate(vox, act, eps_cond=torch.from_numpy(z["eps_cond"]).cuda(),
                           eps_gen=torch.from_numpy(z["eps_gen"]).cuda())
    kp_ref = z["keypoints"]
    Tc = hp.Tcond
    cond_err = np.abs(out["keypoints"][:, :Tc].cpu().numpy() - kp_ref[:, :Tc]).max()
    gen_err = np.abs(out["keypoints"][:, Tc:].cpu().numpy() - kp_ref[:, Tc:]).max()
    vol_err = np.abs(out["gen"][..., ::2, ::2, ::2].cpu().numpy() - z["gen_sub_f16"].astype(np.float32)).mean()
    print(f"[generate] detected keypoints max err {cond_err:.3e}, generated keypoints (15-step roll-out) max err "
          f"{gen_err:.3e}, generated volumes mean abs err {vol_err:.3e}")
    assert cond_err <= KP_TOL
    # generated frames go through a recurrent network (GRU + best-of-10 sample pick) fed by the detected keypoints; they
    # stay inside the same north-star keypoint budget (measured 4.0e-4; detected 9.6e-4; volumes 2.9e-5)
    assert gen_err <= KP_TOL
    assert vol_err <= 1e-3
    assert out["gen"].shape == (1, T, 1, 32, 32, 32) and out["A_hats"] is None


def test_state_dict_roundtrip_and_cache_invalidation():
    hp = O.default_hparams(grid_size=32)
    net, sd = build(hp, 51)
    vox, _ = clips(5100, 1, 2, 5000, 32)
    with torch.no_grad():
        k1 = net.kypt_detector(vox)["keypoints"].clone()
        sd2 = O.synthetic_state_dict(hp, seed=52)
        net.load_state_dict(sd2, strict=True)          # in-place copy: packed weights must be rebuilt
        k2 = net.kypt_detector(vox)["keypoints"].clone()
        net.load_state_dict(sd, strict=True)
        k3 = net.kypt_detector(vox)["keypoints"]
    assert (k1 - k2).abs().max() > 1e-3
    assert torch.equal(k1, k3)
    for k, v in net.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_deepcopy_and_pickle_after_a_forward():
    """The derived-weight caches (packed fp16 weights, ctypes structs of the HSVRNN matrices) live outside the modules:
    a model that has run can be deep-copied / pickled, and the copy computes the same result."""
    import copy
    import io
    hp = O.default_hparams(grid_size=32)
    net, _ = build(hp, 57)
    vox, _ = clips(5700, 1, 3, 20000, 32)
    act = {"detector": True, "learner": True}
    with torch.no_grad():
        net(vox, act)
        ref = net.kypt_detector(vox)["keypoints"]
        twin = copy.deepcopy(net)
        buf = io.BytesIO()
        torch.save(net, buf)
        assert torch.equal(twin.kypt_detector(vox)["keypoints"], ref)
        twin.dyna_module.encode(ref, net.kypt_detector.get_affinity())


def test_training_mode_submodules_raise_but_detector_trains():
    """Autograd is wired through KyptDetector.forward; the sub-networks called on their own in training mode with
    autograd enabled still refuse (no silent graph-less result)."""
    hp = O.default_hparams(grid_size=32)
    net, _ = build(hp, 53)
    net.train()
    vox = torch.zeros(1, 2, 1, 32, 32, 32, device="cuda")
    vox[:, :, :, 8:20, 8:20, 8:20] = 1.0
    with pytest.raises(NotImplementedError):
        net.kypt_detector.vox_to_kypt(vox)
    out = net.kypt_detector(vox)
    assert out["recon_loss"].requires_grad and out["keypoints"].requires_grad
    with torch.no_grad():                                   # the caller's no_grad is respected in training mode
        out = net(vox, {"detector": True, "learner": False})
    assert not out["recon_loss"].requires_grad


@pytest.mark.parametrize("scale", [64.0, 1.0 / 64.0])
def test_fp16_range_scaled_prenorm_tensors(scale):
    """fp16 range stress (the raw pre-GroupNorm conv outputs are stored as fp16): every conv that feeds a GroupNorm has
    its weight and bias multiplied by 64 or 1/64, which scales those raw tensors by the same factor while the reference
    (fp32) output only changes through GroupNorm's eps.  The CUDA path must stay finite and inside the north-star
    tolerances against the oracle run on the SAME scaled checkpoint."""
    import neural_marionette_b200 as nm
    G, B, T = 32, 1, 4
    hp = O.default_hparams(grid_size=G)
    sd = O.synthetic_state_dict(hp, seed=77)
    gn_fed = [k[:-len(".weight")] for k in sd if k.endswith(".weight") and sd[k].dim() == 5 and
              any(tag in k for tag in (".block.0", ".stride_conv.0", ".res_branch.0", ".res_branch.3", ".skip_con.0",
                                       "decode_voxel_from_combined_representation.1.",
                                       "decode_voxel_from_combined_representation.4.",
                                       "decode_voxel_from_combined_representation.8.",
                                       "decode_voxel_from_combined_representation.11."))]
    assert len(gn_fed) >= 70
    for k in gn_fed:
        sd[k + ".weight"] = sd[k + ".weight"] * scale
        sd[k + ".bias"] = sd[k + ".bias"] * scale
    net = nm.NeuralMarionette(hp)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.anneal(1)
    vox, ref_vox = clips(7700, B, T, 20000, G)
    with torch.no_grad():
        out = net.kypt_detector(vox)
        ref = O.detector_forward(ref_vox, sd, hp)
    for k in ("keypoints", "heatmaps", "recon"):
        assert torch.isfinite(out[k]).all(), k
    kp_err = float((out["keypoints"].cpu() - ref["keypoints"]).abs().max())
    hm_err = float((out["heatmaps"].cpu() - ref["heatmaps"]).abs().max() / ref["heatmaps"].max())
    rec_err = float((out["recon"].cpu() - ref["recon"]).abs().mean())
    print(f"[scale {scale:g}] keypoint max err {kp_err:.3e}, heat-map rel err {hm_err:.3e}, recon mean abs err {rec_err:.3e}")
    assert kp_err <= KP_TOL and hm_err <= HM_TOL and rec_err <= 5e-3


def test_stress_resolution_g128_vs_oracle():
    """Config #5: 2x the grid per axis (128^3, heat-map grid 32^3) with 100k points per frame."""
    import neural_marionette_b200 as nm
    G, T, N = 128, 2, 100000
    hp = O.default_hparams(grid_size=G)
    net, sd = build(hp, 61)
    raw = O.synthetic_clip(6100, T, N)[None]
    vox = nm.voxelize_raw_clips(raw, G)
    ref_vox = O.voxelize_clip(O.episodic_normalization(raw[0]), G)[None]
    assert np.array_equal(vox.cpu().numpy(), ref_vox)                  # scatter contention: bit-exact
    with torch.no_grad():
        hm, kp, gs, ff = net.kypt_detector.vox_to_kypt(vox)
        ref = O.vox_to_kypt(torch.from_numpy(ref_vox), sd, hp)
    assert (kp.cpu() - ref[1]).abs().max() <= KP_TOL
    assert float((hm.cpu() - ref[0]).abs().max() / ref[0].max()) <= HM_TOL
    with torch.no_grad():
        rec = net.kypt_detector.kypt_to_vox.decode(net.kypt_detector.vox_to_kypt.detect(vox)["first_feature_act"],
                                                   vox[:, 0], keypoints=kp[:, :1], sigma=1.5)
    assert rec.shape == (1, 1, 1, G, G, G) and bool(torch.isfinite(rec).all())


def test_full_size_config2_properties():
    """BASELINE.json configs[1] at its full size (64 clips x 20 frames x 20 000 points, grid 64^3) through
    size-independent properties: occupancy bit-exact against the oracle on every distinct clip; duplicated clips give
    identical bits wherever they sit in the batch (clips are independent: GroupNorm is per sample, reductions have a
    fixed order); permuting the clips permutes the outputs; value ranges of the reference's definitions; one clip
    checked against the CPU oracle within the north-star tolerances."""
    import neural_marionette_b200 as nm
    G, B, T, N, D = 64, 64, 20, 20000, 4
    hp = O.default_hparams(grid_size=G)
    net, sd = build(hp, 47)
    distinct = np.stack([O.synthetic_clip(7000 + d, T, N) for d in range(D)], 0)
    owner = np.array([(3 * b + b // 7) % D for b in range(B)])              # duplicates spread over both 32-clip passes
    raw = distinct[owner]
    vox = nm.voxelize_raw_clips(raw, G)
    assert vox.shape == (B, T, 1, G, G, G)
    ref_vox = np.stack([O.voxelize_clip(O.episodic_normalization(distinct[d]), G) for d in range(D)], 0)
    first = [int(np.argmax(owner == d)) for d in range(D)]
    for d in range(D):
        assert np.array_equal(vox[first[d]].cpu().numpy(), ref_vox[d])      # bit-exact occupancy
    assert torch.equal(vox, vox[torch.as_tensor(first)][torch.as_tensor(owner)])
    with torch.no_grad():
        out = net.kypt_detector(vox)
        kp, hm, rec = out["keypoints"], out["heatmaps"], out["recon"]
        assert kp.shape == (B, T, 24, 4) and hm.shape == (B, T, 24, 16, 16, 16) and rec.shape == vox.shape
        # value ranges: soft-argmax coordinates inside the cube, intensities in (0, 1], Softplus heat-maps > 0, sigmoid
        assert float(kp[..., :3].abs().max()) < 1.0 and float(kp[..., 3].min()) > 0.0 and float(kp[..., 3].max()) <= 1.0
        assert float(kp[..., 3].amax(dim=-1).min()) > 0.999                 # the brightest keypoint of a frame has I ~ 1
        assert float(hm.min()) > 0.0 and float(rec.min()) >= 0.0 and float(rec.max()) <= 1.0
        assert all(bool(torch.isfinite(out[k]).all()) for k in ("recon_loss", "vol_fit_reg", "separation_loss",
                                                                 "sparsity_loss", "graph_traj_loss"))
        # duplicates: identical bits for every copy of a clip
        rep = torch.as_tensor(first, device=kp.device)[torch.as_tensor(owner, device=kp.device)]
        for key in ("keypoints", "heatmaps", "first_feature"):
            assert torch.equal(out[key], out[key][rep]), key
        assert torch.equal(rec[40:], rec[rep[40:]])
        # permutation equivariance
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).cuda()
        out_p = net.kypt_detector(vox[perm].contiguous())
        assert torch.equal(out_p["keypoints"], kp[perm]) and torch.equal(out_p["heatmaps"], hm[perm])
        # one full clip against the CPU oracle (north-star tolerances)
        ref = O.detector_forward(torch.from_numpy(ref_vox[0])[None], sd, hp)
    b0 = first[0]
    assert (kp[b0].cpu() - ref["keypoints"][0]).abs().max() <= KP_TOL
    assert float((hm[b0].cpu() - ref["heatmaps"][0]).abs().max() / ref["heatmaps"].max()) <= HM_TOL
    assert (rec[b0].cpu() - ref["recon"][0]).abs().mean() <= 5e-3


def test_cuda_graph_replay_is_bit_identical():
    """`graph.capture_detector*`: the captured launch sequence replays to the same bits as the eager call, also on a
    different input of the same shape (config #1 shape is launch-bound; the graph removes the launch gaps)."""
    from neural_marionette_b200 import graph as NG
    import neural_marionette_b200 as nm
    G, B, T = 32, 1, 4
    hp = O.default_hparams(grid_size=G)
    net, _ = build(hp, 53)
    det = net.kypt_detector
    raw_a = torch.from_numpy(O.synthetic_clip(8100, T, 20000))[None].cuda()
    raw_b = torch.from_numpy(O.synthetic_clip(8101, T, 20000))[None].cuda()
    vox_a, vox_b = (nm.voxelize_raw_clips(r, G) for r in (raw_a, raw_b))
    with torch.no_grad():
        eager_a, eager_b = det(vox_a), det(vox_b)
    cap = NG.capture_detector(det, vox_a)
    for vox, eager in ((vox_a, eager_a), (vox_b, eager_b), (vox_a, eager_a)):
        got = cap(vox)
        for key in ("keypoints", "heatmaps", "recon", "first_feature", "recon_loss", "vol_fit_reg", "sparsity_loss",
                    "graph_traj_loss"):
            assert torch.equal(got[key], eager[key]), key
    cap2 = NG.capture_detector_from_points(det, raw_a, G)
    got = cap2(raw_b)
    assert torch.equal(got["voxel"], vox_b) and torch.equal(got["keypoints"], eager_b["keypoints"])
    assert torch.equal(got["recon"], eager_b["recon"])
    with pytest.raises(ValueError):
        cap(vox_a[:, :2])
    # eager calls still work after captures (scratch buffers are per stream)
    with torch.no_grad():
        again = det(vox_b)
    assert torch.equal(again["keypoints"], eager_b["keypoints"])
