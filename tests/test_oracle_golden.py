"""CPU suite: the oracle replays the fixtures that oracle/make_golden.py wrote
from the REFERENCE's outputs (tests/golden/).  This is what pins the oracle on
machines where /root/reference does not exist (the GPU box)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import nm_oracle as O


def _vox(seed, B, T, N, G):
    clips = [O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(seed + b, T, N)), G) for b in range(B)]
    return torch.from_numpy(np.stack(clips, 0)).float()


def test_pin_report_present(golden_dir):
    rep = json.load(open(os.path.join(golden_dir, "ORACLE_PIN.json")))
    assert rep["state_dict"] == {"tensors": 337, "params": 10087015}
    assert rep["checks"]["voxelize_bit_exact_frames"] >= 10
    assert rep["checks"]["skeleton_trials_exact"] >= 50


def test_voxelize_hashes(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "voxelize_hashes.json")))
    clips = {}
    for c in cases:
        key = (c["seed"], c["N"])
        if key not in clips:
            T = 1 + max(x["t"] for x in cases if x["seed"] == c["seed"])
            clips[key] = O.episodic_normalization(O.synthetic_clip(c["seed"], T, c["N"]))
        g = O.voxelize(clips[key][c["t"]], (c["G"],) * 3)
        assert g.shape == (1, c["G"], c["G"], c["G"]) and g.dtype == np.float32
        assert int(g.sum()) == c["occupied"]
        assert hashlib.sha256(g.tobytes()).hexdigest() == c["sha256"]


def test_voxelize_real_geometry(golden_dir):
    z = np.load(os.path.join(golden_dir, "voxelize_obj.npz"))
    pts = O.episodic_normalization(z["obj_points_f32"][None], 0.8)[0]
    g = O.voxelize(pts, (64, 64, 64))
    assert int(g.sum()) == int(z["obj_occupied"])
    assert np.array_equal(np.packbits(g.astype(np.uint8).ravel()), z["obj_grid_packed"])


def test_voxelize_edge_cases():
    # empty cloud -> empty grid; duplicates are idempotent; p just below 1 lands in the last cell
    assert O.voxelize(np.zeros((0, 3)), (8, 8, 8)).sum() == 0
    p = np.array([[0.0, 0.0, 0.0]] * 5 + [[np.nextafter(1.0, 0.0)] * 3, [-1.0, -1.0, -1.0]])
    g = O.voxelize(p, (8, 8, 8))[0]
    assert g.sum() == 3 and g[7, 7, 7] == 1 and g[0, 0, 0] == 1 and g[3, 3, 3] == 1


def test_units(golden_dir):
    z = np.load(os.path.join(golden_dir, "units.npz"))
    g = torch.Generator().manual_seed(int(z["hm_seed"]))
    torch.rand(2, 3, 6, 6, 6, generator=g)
    hm = torch.nn.functional.softplus(4 * torch.randn(3, 24, 16, 16, 16, generator=g))
    kp = O.keypoints_from_heatmap(hm)
    np.testing.assert_allclose(kp.numpy(), z["kp"], atol=2e-6)
    gs = O.render_gaussians(torch.from_numpy(z["kp"]), 1.5, 16)
    np.testing.assert_allclose(gs.sum(dim=(2, 3, 4)).numpy(), z["gauss_sum"], rtol=1e-5)
    np.testing.assert_allclose(O.rot6d_to_matrix(torch.from_numpy(z["rot6d_in"])).numpy(), z["rot6d_out"], atol=1e-6)


def test_skeleton(golden_dir):
    hp = O.default_hparams()
    for c in json.load(open(os.path.join(golden_dir, "skeleton.json"))):
        sd = {"kypt_detector.affinity_params": torch.tensor(c["affinity_params"])}
        _, pr, pa = O.skeleton_from_affinity(O.get_affinity(sd, hp))
        assert pa.tolist() == c["parents"] and pr.indices.tolist() == c["order"]
        np.testing.assert_allclose(pr.values.numpy(), np.array(c["values"]), rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", ["g32", "g64"])
def test_detector(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"detector_{tag}.npz"))
    G, B, T = int(z["G"]), int(z["B"]), int(z["T"])
    hp = O.default_hparams(grid_size=G)
    sd = O.synthetic_state_dict(hp, seed=int(z["seed"]))
    assert len(sd) == 337
    vox = _vox(int(z["vox_seed"]), B, T, 20000, G)
    with torch.no_grad():
        out = O.detector_forward(vox, sd, hp)
        gen = O.decode_from_dyna(torch.from_numpy(z["gen_kp"]), out["first_feature"], vox[:, 0], sd, hp)
    # different host CPUs pick different oneDNN kernels: allow fp32 noise, not more
    np.testing.assert_allclose(out["keypoints"].numpy(), z["keypoints"], atol=2e-5)
    ref_hm = z["heatmaps_t0"]
    assert np.abs(out["heatmaps"][:, :1].numpy() - ref_hm).max() <= 1e-4 * ref_hm.max()
    np.testing.assert_allclose(out["recon"][..., ::2, ::2, ::2].numpy(), z["recon_sub_f16"].astype(np.float32), atol=2e-3)
    np.testing.assert_allclose(gen[..., ::2, ::2, ::2].numpy(), z["gen_sub_f16"].astype(np.float32), atol=2e-3)
    np.testing.assert_allclose(out["affinity"].numpy(), z["affinity"], atol=1e-7)
    got = np.array([float(out[k]) for k in ["recon_loss", "vol_fit_reg", "separation_loss", "sparsity_loss",
                                            "local_const_loss", "time_const_loss", "sparsity_const_loss",
                                            "graph_traj_loss"]])
    np.testing.assert_allclose(got, z["losses"], rtol=1e-4, atol=1e-6)


def test_dynamics(golden_dir):
    z = np.load(os.path.join(golden_dir, "dynamics.npz"))
    hp = O.default_hparams()
    sd = O.synthetic_state_dict(hp, seed=int(z["seed"]))
    skel = O.skeleton_from_affinity(O.get_affinity(sd, hp))
    assert skel[2].tolist() == z["parents"].tolist() and skel[1].indices.tolist() == z["order"].tolist()
    kp = torch.from_numpy(z["kp"])
    with torch.no_grad():
        out = O.dyna_generate(kp[:, :3], skel, sd, hp, Ttot=9, Tcond=3,
                              eps_cond=torch.from_numpy(z["eps_cond"]), eps_gen=torch.from_numpy(z["eps_gen"]))
        np.testing.assert_allclose(out["keypoints_cond"].numpy(), z["keypoints_cond"], atol=2e-5)
        np.testing.assert_allclose(out["keypoints_gen"].numpy(), z["keypoints_gen"], atol=2e-5)
        off = O.bone_offsets(kp, skel[2], sd)
        np.testing.assert_allclose(off.numpy(), z["offset"], atol=1e-7)
        f, R = O.decode_pose(torch.from_numpy(z["dec_in"]), off, skel[1].indices, skel[2], sd, 24)
        np.testing.assert_allclose(f.numpy(), z["dec_flat"], atol=2e-6)
        np.testing.assert_allclose(R.numpy(), z["dec_R"], atol=2e-6)
        h = O.gru_cell(torch.from_numpy(z["gru_x"]), torch.from_numpy(z["gru_h"]), sd, "dyna_module.kypt_rnn_cell")
        np.testing.assert_allclose(h.numpy(), z["gru_out"], atol=2e-6)


def test_generate(golden_dir):
    z = np.load(os.path.join(golden_dir, "generate_g32.npz"))
    hp = O.default_hparams(grid_size=32)
    sd = O.synthetic_state_dict(hp, seed=int(z["seed"]))
    vox = _vox(int(z["vox_seed"]), 1, int(z["T"]), 20000, 32)
    with torch.no_grad():
        out = O.marionette_generate(vox, sd, hp, eps_cond=torch.from_numpy(z["eps_cond"]),
                                    eps_gen=torch.from_numpy(z["eps_gen"]))
    np.testing.assert_allclose(out["keypoints"].numpy(), z["keypoints"], atol=5e-5)
    np.testing.assert_allclose(out["gen"][..., ::2, ::2, ::2].numpy(), z["gen_sub_f16"].astype(np.float32), atol=3e-3)


# ---- plain-C oracle (oracle/nm_oracle_c.c): same fixtures, no numpy in the arithmetic ------------------------------
def test_c_oracle_voxelize_hashes(golden_dir):
    from oracle import c_oracle as C
    cases = json.load(open(os.path.join(golden_dir, "voxelize_hashes.json")))
    clips = {}
    for c in cases:
        key = (c["seed"], c["N"])
        if key not in clips:
            T = 1 + max(x["t"] for x in cases if x["seed"] == c["seed"])
            raw = O.synthetic_clip(c["seed"], T, c["N"])
            clips[key] = (raw, C.episodic_normalization(raw))
            assert np.array_equal(clips[key][1], O.episodic_normalization(raw))        # float64, bit for bit
        g = C.voxelize(clips[key][1][c["t"]], c["G"])
        assert int(g.sum()) == c["occupied"]
        assert hashlib.sha256(g.tobytes()).hexdigest() == c["sha256"]


def test_c_oracle_real_geometry_and_clip(golden_dir):
    from oracle import c_oracle as C
    z = np.load(os.path.join(golden_dir, "voxelize_obj.npz"))
    g = C.normalize_voxelize_clip(z["obj_points_f32"][None], 64, scale=0.8)[0]
    assert int(g.sum()) == int(z["obj_occupied"])
    assert np.array_equal(np.packbits(g.astype(np.uint8).ravel()), z["obj_grid_packed"])
    raw = O.synthetic_clip(77, 3, 5000)
    for kw in (dict(), dict(scale=0.8, x_trans=0.05, z_trans=0.1)):
        want = O.voxelize_clip(O.episodic_normalization(raw, **kw), 32)
        assert np.array_equal(C.normalize_voxelize_clip(raw, 32, **kw), want)


def test_c_oracle_edge_cases():
    from oracle import c_oracle as C
    assert C.voxelize(np.zeros((0, 3)), 8).sum() == 0
    p = np.array([[0.0, 0.0, 0.0]] * 5 + [[np.nextafter(1.0, 0.0)] * 3, [-1.0, -1.0, -1.0]])
    g = C.voxelize(p, 8)[0]
    assert g.sum() == 3 and g[7, 7, 7] == 1 and g[0, 0, 0] == 1 and g[3, 3, 3] == 1
    with pytest.raises(IndexError):
        C.voxelize(np.array([[1.5, 0.0, 0.0]]), 8)
    with pytest.raises(ValueError):
        C.episodic_normalization(np.zeros((0, 4, 3), np.float32))
    # extra columns (normals) are ignored, utils/dataset_utils.py:27
    q = np.concatenate([p, np.ones((len(p), 3))], 1)
    assert np.array_equal(C.voxelize(q, 8), g[None])


def test_interpolation_fixture(golden_dir):
    """oracle.dyna_interpolate replays the reference's vis_interpolation.py loop (fixture written from the reference)."""
    z = np.load(os.path.join(golden_dir, "interpolation.npz"))
    hp = O.default_hparams()
    sd = O.synthetic_state_dict(hp, seed=int(z["seed"]))
    skeleton = O.skeleton_from_affinity(O.get_affinity(sd, hp))
    sel, picks = O.dyna_interpolate(torch.from_numpy(z["kp"]), skeleton, sd, hp, int(z["sample_num"]), int(z["sample_rate"]),
                                    torch.from_numpy(z["eps"]))
    assert [list(p) for p in picks] == z["picks"].tolist()
    assert (sel - torch.from_numpy(z["selected"])).abs().max() <= 5e-6
