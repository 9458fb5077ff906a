"""Pin the oracle against the reference and (re)generate tests/golden/.

Run in the BUILD container only (it imports /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

It (1) loads ``oracle.synthetic_state_dict`` into the *reference*
``NeuralMarionette`` with ``load_state_dict(strict=True)`` — which proves the
key/shape layout is the reference's checkpoint layout, (2) runs reference and
oracle on the same seeded inputs and asserts agreement (bit-exact where the
op order is identical, <= 2e-6 otherwise), (3) writes small fixtures with the
REFERENCE's outputs and a JSON report (tests/golden/ORACLE_PIN.json).
"""
from __future__ import annotations

import hashlib
import json
import os
import pickle
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import nm_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
report = {"torch": torch.__version__, "numpy": np.__version__, "checks": {}}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def maxdiff(a, b) -> float:
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    return float(torch.nan_to_num(a - b).abs().max())


def check(name, a, b, tol):
    d = maxdiff(a, b) / max(1.0, float(torch.nan_to_num(torch.as_tensor(a).double()).abs().max()))
    report["checks"][name] = {"max_diff_rel_to_max1": d, "tol": tol}
    assert d <= tol, f"{name}: oracle differs from reference by {d} > {tol}"
    print(f"  ok  {name:55s} max|diff| = {d:.3e}")


def ref_opt(**over):
    with open(os.path.join(REF, "pretrained/aist/opt.pickle"), "rb") as f:
        opt = pickle.load(f)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


# --------------------------------------------------------------------------- voxelize
def golden_voxelize():
    from utils.dataset_utils import episodic_normalization, voxelize
    print("voxelize / episodic_normalization")
    out = {}
    cases = [(1000, 4, 20000, 64), (1001, 3, 20000, 64), (1002, 2, 100000, 128), (1003, 2, 777, 32)]
    hashes = []
    for seed, T, N, G in cases:
        clip = O.synthetic_clip(seed, T, N)
        ref_n = episodic_normalization(clip)
        ora_n = O.episodic_normalization(clip)
        assert ref_n.dtype == ora_n.dtype == np.float64 and np.array_equal(ref_n, ora_n)
        for t in range(T):
            r = voxelize(ref_n[t], (G, G, G), True)
            o = O.voxelize(ora_n[t], (G, G, G))
            assert r.dtype == o.dtype == np.float32 and np.array_equal(r, o)
            hashes.append(dict(seed=seed, t=t, N=N, G=G, occupied=int(r.sum()), sha256=sha(r)))
    report["checks"]["voxelize_bit_exact_frames"] = len(hashes)
    print(f"  ok  {len(hashes)} frames bit-exact vs reference")
    # scale / translation variants (vis_retarget.py:92-100 uses scale 0.8)
    clip = O.synthetic_clip(1004, 2, 5000)
    for scale, xt, zt in [(0.8, 0.0, 0.0), (0.7, 0.1, 0.05)]:
        assert np.array_equal(episodic_normalization(clip, scale, xt, zt),
                              O.episodic_normalization(clip, scale, xt, zt))
    # the only real geometry in the reference tree: target.obj vertices
    verts = []
    with open(os.path.join(REF, "data/demo/target/ninja/target.obj")) as f:
        for line in f:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
    verts = np.asarray(verts, dtype=np.float32)
    norm = episodic_normalization(verts[None], 0.8)[0]
    grid = voxelize(norm, (64, 64, 64), True)
    assert np.array_equal(grid, O.voxelize(O.episodic_normalization(verts[None], 0.8)[0], (64,) * 3))
    out["obj_points_f32"] = verts            # raw vertices (float32), ~150 kB
    out["obj_grid_packed"] = np.packbits(grid.astype(np.uint8).ravel())
    out["obj_occupied"] = np.int64(grid.sum())
    np.savez_compressed(os.path.join(GOLD, "voxelize_obj.npz"), **out)
    with open(os.path.join(GOLD, "voxelize_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)


# --------------------------------------------------------------------------- detector
def build_reference(hp_over, seed):
    from model.neural_marionette import NeuralMarionette
    opt = ref_opt(**hp_over)
    net = NeuralMarionette(opt).eval()
    net.anneal(1)
    hp = O.default_hparams(**hp_over)
    sd = O.synthetic_state_dict(hp, seed=seed)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net, sd, hp


def make_vox(seed, B, T, N, G):
    clips = []
    for b in range(B):
        pts = O.episodic_normalization(O.synthetic_clip(seed + b, T, N))
        clips.append(O.voxelize_clip(pts, G))
    return torch.from_numpy(np.stack(clips, 0)).float()


def golden_detector():
    print("KyptDetector.forward / decode_from_dyna")
    for tag, G, B, T, seed in [("g64", 64, 1, 3, 11), ("g32", 32, 2, 3, 12)]:
        net, sd, hp = build_reference(dict(grid_size=G), seed)
        if tag == "g64":
            n_t = len(sd)
            n_p = sum(v.numel() for v in sd.values())
            report["state_dict"] = {"tensors": n_t, "params": n_p}
            assert (n_t, n_p) == (337, 10087015), (n_t, n_p)
        vox = make_vox(2000 + seed, B, T, 20000, G)
        with torch.no_grad():
            ref = net.kypt_detector(vox)
            ora = O.detector_forward(vox, sd, hp)
            gen_kp = ref["keypoints"] * 0.9
            ref_gen = net.kypt_detector.decode_from_dyna(gen_kp, ref["first_feature"], vox[:, 0])["gen"]
            ora_gen = O.decode_from_dyna(gen_kp, ora["first_feature"], vox[:, 0], sd, hp)
        for k in ["keypoints", "heatmaps", "recon", "first_feature", "affinity"]:
            check(f"detector[{tag}].{k}", ref[k], ora[k], 2e-6 if k != "recon" else 2e-5)
        for k in ["recon_loss", "vol_fit_reg", "separation_loss", "sparsity_loss", "local_const_loss",
                  "time_const_loss", "sparsity_const_loss", "intensity_const_loss", "graph_traj_loss",
                  "graph_vol_loss", "kypt_const_loss"]:
            check(f"detector[{tag}].{k}", ref[k], ora[k], 2e-6)
        check(f"detector[{tag}].decode_from_dyna", ref_gen, ora_gen, 2e-5)
        kp = ref["keypoints"]
        report[f"keypoint_spread[{tag}]"] = dict(
            xyz_min=float(kp[..., :3].min()), xyz_max=float(kp[..., :3].max()),
            inten_min=float(kp[..., 3].min()), hm_min=float(ref["heatmaps"].min()),
            hm_max=float(ref["heatmaps"].max()))
        np.savez_compressed(
            os.path.join(GOLD, f"detector_{tag}.npz"),
            seed=seed, vox_seed=2000 + seed, B=B, T=T, G=G,
            keypoints=kp.numpy(), heatmaps_t0=ref["heatmaps"][:, :1].numpy().astype(np.float32),
            recon_sub_f16=ref["recon"][..., ::2, ::2, ::2].numpy().astype(np.float16),
            first_feature_sub_f16=ref["first_feature"][:, ::8].numpy().astype(np.float16),
            gen_kp=gen_kp.numpy(), gen_sub_f16=ref_gen[..., ::2, ::2, ::2].numpy().astype(np.float16),
            losses=np.array([float(ref[k]) for k in ["recon_loss", "vol_fit_reg", "separation_loss",
                                                      "sparsity_loss", "local_const_loss",
                                                      "time_const_loss", "sparsity_const_loss",
                                                      "graph_traj_loss"]], dtype=np.float64),
            affinity=ref["affinity"].numpy())


# --------------------------------------------------------------------------- unit functions
def golden_units():
    from utils.kypt_detector_utils import (add_coord_channels, extract_gaussian_map_from_keypoints,
                                           extract_keypoints_from_heatmap)
    from utils.geo_utils import compute_rotation_matrix_from_6d
    print("unit functions")
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 6, 6, 6, generator=g)
    check("add_coord_channels", add_coord_channels(x), O.add_coord_channels(x), 0.0)
    hm = torch.nn.functional.softplus(4 * torch.randn(3, 24, 16, 16, 16, generator=g))
    kp_r = extract_keypoints_from_heatmap(hm.clone())
    check("extract_keypoints_from_heatmap", kp_r, O.keypoints_from_heatmap(hm), 0.0)
    gs_r = torch.cat([extract_gaussian_map_from_keypoints(kp_r[:, k:k + 1], sigma=1.5, G=16)
                      for k in range(24)], dim=1)
    check("extract_gaussian_map_from_keypoints", gs_r, O.render_gaussians(kp_r, 1.5, 16), 0.0)
    p = torch.randn(4, 24, 6, generator=g)
    check("compute_rotation_matrix_from_6d", compute_rotation_matrix_from_6d(p), O.rot6d_to_matrix(p), 0.0)
    np.savez_compressed(os.path.join(GOLD, "units.npz"), hm_seed=5, kp=kp_r.numpy(),
                        gauss_sum=gs_r.sum(dim=(2, 3, 4)).numpy(), rot6d_in=p.numpy(),
                        rot6d_out=compute_rotation_matrix_from_6d(p).numpy())


# --------------------------------------------------------------------------- skeleton + dynamics
def golden_dynamics():
    from utils.dyna_utils import process_affinity_glob
    print("skeleton + HSVRNNBVH")
    hp = O.default_hparams()
    skel = []
    g = torch.Generator().manual_seed(77)
    n_ok = 0
    for trial in range(60):
        scale = [0.0, 0.5, 2.0, 6.0][trial % 4]
        sd = {"kypt_detector.affinity_params": torch.randn(2, 24, 23, generator=g) * scale + 1.0}
        aff = O.get_affinity(sd, hp)
        A_r, pr_r, pa_r = process_affinity_glob(aff)
        A_o, pr_o, pa_o = O.skeleton_from_affinity(aff)
        assert torch.equal(pa_r, pa_o), (trial, pa_r, pa_o)
        assert torch.equal(pr_r.indices, pr_o.indices), trial
        assert torch.equal(pr_r.values, pr_o.values), trial
        assert torch.equal(A_r, A_o), trial
        n_ok += 1
        if trial < 8:
            skel.append(dict(affinity_params=sd["kypt_detector.affinity_params"].numpy().tolist(),
                             parents=pa_r.tolist(), order=pr_r.indices.tolist(),
                             values=pr_r.values.tolist()))
    report["checks"]["skeleton_trials_exact"] = n_ok
    print(f"  ok  skeleton: {n_ok} random affinities identical (parents, priority, A)")
    with open(os.path.join(GOLD, "skeleton.json"), "w") as f:
        json.dump(skel, f)

    net, sd, hp = build_reference({}, seed=21)
    B, T, K, Z = 3, 6, 24, hp.nlatent_kypt
    gk = torch.Generator().manual_seed(99)
    kp = torch.cat([torch.rand(B, T, K, 3, generator=gk) * 1.2 - 0.6,
                    torch.rand(B, T, K, 1, generator=gk)], dim=-1)
    kp = kp + 0.02 * torch.arange(T).float()[None, :, None, None]
    with torch.no_grad():
        aff = net.kypt_detector.get_affinity()
        check("get_affinity", aff, O.get_affinity(sd, hp), 0.0)
        torch.manual_seed(2)
        ref = net.dyna_module.encode(kp, aff)
        skeleton = (net.dyna_module.A, net.dyna_module.priority, net.dyna_module.parents)
        torch.manual_seed(2)
        ora = O.dyna_encode(kp, O.skeleton_from_affinity(aff), sd, hp)
        for k in ["kypt_recon", "R", "z_kypts", "h_kypts", "kl_kypt", "kypt_recon_loss"]:
            check(f"encode.{k}", ref[k], ora[k], 5e-6)
        torch.manual_seed(3)
        refg = net.dyna_module.generate(kp[:, :3], aff, Ttot=9, Tcond=3)
        torch.manual_seed(3)
        eps_c = torch.stack([torch.normal(torch.zeros(10, B, Z), torch.ones(10, B, Z)) for _ in range(3)])
        eps_g = torch.stack([torch.normal(torch.zeros(B, Z), torch.ones(B, Z)) for _ in range(6)])
        orag = O.dyna_generate(kp[:, :3], skeleton, sd, hp, Ttot=9, Tcond=3, eps_cond=eps_c, eps_gen=eps_g)
        for k in ["keypoints_cond", "keypoints_gen"]:
            check(f"generate.{k} (injected eps == rsample order)", refg[k], orag[k], 5e-6)
        # demo-style pieces reached into by vis_generation.py:86-127
        off_r = net.dyna_module.get_offset(kp)
        check("get_offset", off_r, O.bone_offsets(kp, skeleton[2], sd), 0.0)
        dec_in = torch.randn(B, 640, generator=gk)
        f_r, R_r = net.dyna_module.extract_kypt_from_latent_and_state(dec_in, off_r)
        f_o, R_o = O.decode_pose(dec_in, off_r, skeleton[1].indices, skeleton[2], sd, K)
        check("extract_kypt_from_latent_and_state.flat", f_r, f_o, 1e-6)
        check("extract_kypt_from_latent_and_state.R", R_r, R_o, 1e-6)
        x = torch.randn(B, 224, generator=gk)
        h = torch.randn(B, 512, generator=gk)
        check("kypt_rnn_cell", net.dyna_module.kypt_rnn_cell(x, h), O.gru_cell(x, h, sd, "dyna_module.kypt_rnn_cell"), 2e-6)
    np.savez_compressed(
        os.path.join(GOLD, "dynamics.npz"), seed=21, kp=kp.numpy(), eps_cond=eps_c.numpy(),
        eps_gen=eps_g.numpy(), keypoints_cond=refg["keypoints_cond"].numpy(),
        keypoints_gen=refg["keypoints_gen"].numpy(), parents=skeleton[2].numpy(),
        order=skeleton[1].indices.numpy(), offset=off_r.numpy(),
        enc_kypt_recon=ref["kypt_recon"].numpy(), enc_h_last=ref["h_kypts"][:, -1].numpy(),
        enc_kl=float(ref["kl_kypt"]), dec_in=dec_in.numpy(), dec_flat=f_r.numpy(), dec_R=R_r.numpy(),
        gru_x=x.numpy(), gru_h=h.numpy(), gru_out=net.dyna_module.kypt_rnn_cell(x, h).detach().numpy())


def golden_generate():
    print("NeuralMarionette.generate (config #1 call sequence)")
    net, sd, hp = build_reference(dict(grid_size=32), seed=31)
    B, T = 1, 8
    vox = make_vox(4000, B, T, 20000, 32)
    Z = hp.nlatent_kypt
    with torch.no_grad():
        # HSVRNNBVH.generate needs the skeleton that only encode() builds
        # (hsvrnn_bvh.py:75-79): every caller tracks once before generating
        net(vox, {"detector": True, "learner": True})
        torch.manual_seed(2)
        ref = net.generate(vox, {"detector": True, "learner": True})
        torch.manual_seed(2)
        eps_c = torch.stack([torch.normal(torch.zeros(10, B, Z), torch.ones(10, B, Z)) for _ in range(hp.Tcond)])
        eps_g = torch.stack([torch.normal(torch.zeros(B, Z), torch.ones(B, Z)) for _ in range(T - hp.Tcond)])
        ora = O.marionette_generate(vox, sd, hp, eps_cond=eps_c, eps_gen=eps_g)
    check("generate.keypoints", ref["keypoints"], ora["keypoints"], 1e-5)
    check("generate.gen", ref["gen"], ora["gen"], 1e-4)
    np.savez_compressed(os.path.join(GOLD, "generate_g32.npz"), seed=31, vox_seed=4000, T=T,
                        eps_cond=eps_c.numpy(), eps_gen=eps_g.numpy(), keypoints=ref["keypoints"].numpy(),
                        gen_sub_f16=ref["gen"][..., ::2, ::2, ::2].numpy().astype(np.float16))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    golden_voxelize()
    golden_units()
    golden_detector()
    golden_dynamics()
    golden_generate()
    with open(os.path.join(GOLD, "ORACLE_PIN.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("all oracle checks passed; fixtures written to", GOLD)
