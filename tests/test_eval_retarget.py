"""Consumers of the path's outputs: evaluation metrics (utils/eval_utils.py) and retarget
post-processing (vis_retarget.py).  tests/golden/eval_retarget.npz holds the REFERENCE's outputs
(oracle/make_golden_eval.py); the CPU tests replay them through the oracle, the GPU tests through the product.

Tolerances: nearest-joint indices / histograms / in-place masks bit-exact (integer work); chamfer 1e-5 relative
(the reference sums float32 distances, the kernel exact integers); skin weights 2e-6; FK 1e-6; skinned points 2e-6
(fp32 kernel vs the reference's float64 einsums, coordinates in [-1, 1])."""
import os
from collections import namedtuple

import numpy as np
import pytest
import torch

from oracle import nm_oracle as O
from oracle import nm_oracle_eval as E

Priority = namedtuple("Priority", ["values", "indices"])


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "eval_retarget.npz"))


def _vc_inputs(gold):
    seed, T, N, G, B = (int(v) for v in gold["vc_seeds"])
    vox = torch.from_numpy(np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(seed + b, T, N)), G)
                                     for b in range(B)], 0)).float()
    return vox, torch.from_numpy(gold["vc_soft"].astype(np.float32))


# ------------------------------------------------------------------------------------------------ CPU: oracle vs fixture
def test_oracle_voxel_chamfer(gold):
    vox, soft = _vc_inputs(gold)
    log = E.voxel_chamfer_distance(None, vox, soft)
    assert np.array_equal(np.array(log["scores"]), gold["vc_ref_scores"])
    assert log["scores_log"] == float(gold["vc_ref_log"])


def test_oracle_semantic(gold):
    kp, gt = torch.from_numpy(gold["sem_kp"]), torch.from_numpy(gold["sem_gt"])
    log = E.semantic_scores(None, kp, gt)
    masked, idx, _ = E.semantic_nearest(kp, gt)
    assert np.array_equal(log["scores"], gold["sem_ref_scores"]) and log["scores_log"] == float(gold["sem_ref_log"])
    assert np.array_equal(masked.numpy(), gold["sem_ref_masked"]) and np.array_equal(idx.numpy(), gold["sem_ref_idx"])


def test_oracle_skin_fk_lbs(gold):
    w = E.extract_skin_weights(gold["sk_order"], gold["sk_parents"], gold["sk_points"], torch.from_numpy(gold["sk_kp"]))
    assert np.abs(w - gold["sk_ref_weights"]).max() <= 1e-6
    assert np.allclose(w.sum(1), 1.0, atol=1e-6) and ((w != 0).sum(1) <= 2).all()
    pos = E.retarget_fk(torch.from_numpy(gold["fk_R"]), torch.from_numpy(gold["fk_off"]), torch.from_numpy(gold["fk_root"]),
                        gold["sk_order"], gold["sk_parents"])
    assert np.abs(pos.numpy() - gold["fk_pos"]).max() <= 1e-6
    T3x4 = np.concatenate([gold["fk_R"], gold["fk_pos"][..., None]], -1)
    lbs = E.linear_blend_skinning(gold["sk_points"], gold["sk_kp"][:, :3], gold["fk_Rinv"], T3x4, gold["sk_ref_weights"])
    assert np.abs(lbs - gold["lbs_out"]).max() <= 1e-6


def test_eval_api_rejects_cpu_tensors_and_bad_metric():
    from neural_marionette_b200.utils import eval_utils as U
    with pytest.raises(ValueError):
        U.evaluate("nope", {}, {})
    with pytest.raises(Exception):
        U.voxel_chamfer_distance(None, dict(voxel=torch.zeros(1, 1, 1, 8, 8, 8), recon=torch.zeros(1, 1, 1, 8, 8, 8)))


# ------------------------------------------------------------------------------------------------ GPU: product vs fixture
@pytest.mark.gpu
def test_gpu_voxel_chamfer_matches_reference(gold):
    from neural_marionette_b200.utils import eval_utils as U
    vox, soft = _vc_inputs(gold)
    recon = soft.cuda()
    log = U.evaluate("voxel_chamfer", {"voxel_chamfer": None}, dict(voxel=vox.cuda(), recon=recon))
    assert np.allclose(np.array(log["scores"]), gold["vc_ref_scores"], rtol=1e-5, atol=0)
    assert abs(log["scores_log"] - float(gold["vc_ref_log"])) <= 1e-5 * float(gold["vc_ref_log"])
    # the reference binarises params['recon'] in place (utils/eval_utils.py:37-38)
    assert torch.equal(recon.cpu(), (soft >= 0.5).float())
    # appending to an existing score list, as train.py does batch after batch
    log2 = U.voxel_chamfer_distance(log["scores"], dict(voxel=vox.cuda(), recon=soft.cuda()))
    assert len(log2["scores"]) == 2 * vox.shape[0]


@pytest.mark.gpu
def test_gpu_voxel_chamfer_properties():
    from neural_marionette_b200 import ops
    G = 64
    vox = torch.from_numpy(O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(41, 5, 20000)), G)).float()[:, 0].cuda()
    # identical volumes -> 0; symmetric in its arguments; chunked calls agree with one call; per-frame oracle
    same, occ, err = ops.voxel_chamfer(vox, vox.clone())
    assert float(same.abs().max()) == 0.0 and int(err.item()) == 0
    assert torch.equal(occ[:, 0].cpu(), vox.flatten(1).sum(1).int().cpu()) and torch.equal(occ[:, 0], occ[:, 1])
    rolled = torch.roll(vox, shifts=(2, 1), dims=(1, 3)).contiguous()
    ab, _, _ = ops.voxel_chamfer(vox, rolled.clone())
    ba, _, _ = ops.voxel_chamfer(rolled, vox.clone())
    assert torch.equal(ab, ba)
    ab2, _, _ = ops.voxel_chamfer(vox, rolled.clone(), frames_per_call=2)
    assert torch.equal(ab, ab2)
    want, _ = E.voxel_chamfer_per_frame(vox[:2].cpu(), rolled[:2].cpu())
    assert np.allclose(ab[:2].cpu().numpy(), want, rtol=1e-5, atol=0)
    # an empty reconstruction: the reference raises, so does the product
    from neural_marionette_b200.utils import eval_utils as U
    with pytest.raises(IndexError):
        U.voxel_chamfer_distance(None, dict(voxel=vox[None, :1, None], recon=torch.zeros(1, 1, 1, G, G, G).cuda()))


@pytest.mark.gpu
def test_gpu_semantic_scores_match_reference(gold):
    from neural_marionette_b200.utils import eval_utils as U
    from neural_marionette_b200 import ops
    kp = torch.from_numpy(gold["sem_kp"]).cuda()
    log = U.evaluate("semantic", {"semantic": None}, dict(keypoints=kp, gt_keypoints=torch.from_numpy(gold["sem_gt"]).cuda()))
    assert np.array_equal(log["scores"], gold["sem_ref_scores"])
    assert log["scores_log"] == float(gold["sem_ref_log"])
    assert np.array_equal(kp.cpu().numpy(), gold["sem_ref_masked"])          # in-place mask, :68-69
    kp2 = torch.from_numpy(gold["sem_kp"]).cuda().view(-1, kp.shape[2], 4)
    idx, _ = ops.semantic_nearest(kp2, torch.from_numpy(gold["sem_gt"]).cuda().view(kp2.shape[0], -1, 3))
    assert np.array_equal(idx.cpu().numpy(), gold["sem_ref_idx"])
    log2 = U.semantic_scores(log["scores"], dict(keypoints=kp, gt_keypoints=torch.from_numpy(gold["sem_gt"]).cuda()))
    assert np.array_equal(log2["scores"], 2 * gold["sem_ref_scores"])


@pytest.mark.gpu
def test_gpu_skin_weights_fk_lbs_match_reference(gold):
    from neural_marionette_b200.utils import retarget_utils as RT
    order = torch.from_numpy(gold["sk_order"])
    parents = torch.from_numpy(gold["sk_parents"])
    kp = torch.from_numpy(gold["sk_kp"]).cuda()
    w = RT.extract_skin_weights(torch.zeros(1).cuda(), Priority(None, order), parents, gold["sk_points"], kp, 8.0, 0.2)
    ref = gold["sk_ref_weights"]
    assert w.shape == ref.shape and w.dtype == np.float32
    same_bone = ((w != 0) == (ref != 0)).all(1)
    assert same_bone.mean() >= 0.999                     # a point equidistant from two bones may flip on the last ulp
    assert np.abs(w[same_bone] - ref[same_bone]).max() <= 2e-6

    R = torch.from_numpy(gold["fk_R"]).cuda()
    src = torch.zeros(1, R.shape[0], R.shape[1], 4).cuda()
    src[0, :, int(order[0]), :3] = torch.from_numpy(gold["fk_root"]).cuda()
    src[..., 3] = 0.7
    new_kp = RT.retarget_keypoints(R, torch.from_numpy(gold["fk_off"]).cuda().view(1, -1, 3, 1), src, Priority(None, order), parents)
    assert new_kp.shape == src.shape and float((new_kp[..., 3] - 0.7).abs().max()) == 0.0
    assert np.abs(new_kp[0, ..., :3].cpu().numpy() - gold["fk_pos"]).max() <= 1e-6

    pos = torch.from_numpy(gold["fk_pos"]).cuda()
    out = RT.linear_blend_skinning(gold["sk_points"], gold["sk_kp"][:, :3], gold["fk_Rinv"], R, pos, ref)
    assert out.shape == gold["lbs_out"].shape
    assert np.abs(out - gold["lbs_out"]).max() <= 2e-6
    # identity bind rotations (the non-`--ours` branch) against the oracle
    T3x4 = np.concatenate([gold["fk_R"], gold["fk_pos"][..., None]], -1)
    want = E.linear_blend_skinning(gold["sk_points"], gold["sk_kp"][:, :3], None, T3x4, ref)
    out = RT.linear_blend_skinning(gold["sk_points"], gold["sk_kp"][:, :3], None, R, pos, ref)
    assert np.abs(out - want).max() <= 2e-6


@pytest.mark.gpu
def test_gpu_skin_weights_invalid_root_raises():
    from neural_marionette_b200.utils import retarget_utils as RT
    kp = torch.rand(4, 4).cuda()
    kp[:, 3] = torch.tensor([0.1, 0.9, 0.1, 0.9])        # the root (0) and joint 2 are invalid: 3 -> 2 -> 0 -> 0 -> ...
    with pytest.raises(RuntimeError):
        RT.extract_skin_weights(kp, Priority(None, torch.tensor([0, 1, 2, 3])), torch.tensor([0, 0, 0, 2]),
                                np.random.rand(10, 3).astype(np.float32), kp)


def test_evaluate_final_matches_reference_arithmetic(tmp_path):
    """utils/eval_utils.py:12-27: pure host arithmetic (the CSV files go to <result_dir>/semantic|chamfer)."""
    from neural_marionette_b200.utils import eval_utils as U
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, size=(17, 24)).astype(np.float64)
    counts[:, 0] += 1
    counts *= 600.0 / counts.sum(1, keepdims=True)            # every gt joint was matched in the same number of frames
    want = (counts / counts[0].sum()).max(axis=-1)
    got = U.evaluate_final("semantic", {"semantic": counts.copy()}, result_dir=str(tmp_path))
    assert got == want.mean()
    assert np.allclose(np.loadtxt(tmp_path / "semantic" / "semantic_result.csv", delimiter=","), want)
    per_clip = [[0.0012], [0.0034], [0.0005]]
    got = U.evaluate_final("voxel_chamfer", {"voxel_chamfer": per_clip}, result_dir=str(tmp_path))
    assert got == np.array(per_clip).mean() * 1e4
    assert np.allclose(np.loadtxt(tmp_path / "chamfer" / "chamfer_result.csv", delimiter=","), np.array(per_clip)[:, 0])
    with pytest.raises(ValueError):
        U.evaluate_final("nope", {})


def test_c_oracle_metrics_replay_reference_fixture(gold):
    """Plain-C restatement (oracle/nm_oracle_c.c) of the two evaluation metrics against the reference-derived fixture."""
    from oracle import c_oracle as C
    vox, soft = _vc_inputs(gold)
    B, T, G = vox.shape[0], vox.shape[1], vox.shape[-1]
    per = np.array([C.voxel_chamfer_frame(vox[b, t, 0].numpy(), soft[b, t, 0].numpy()) for b in range(B) for t in range(T)])
    assert np.allclose(per, gold["vc_ref_per_frame"], rtol=1e-6, atol=0)
    assert np.allclose(per.reshape(B, T).mean(1), gold["vc_ref_scores"][:, 0], rtol=1e-6, atol=0)
    kp, gt = gold["sem_kp"], gold["sem_gt"]
    idx = np.stack([C.semantic_nearest_frame(kp.reshape(-1, *kp.shape[2:])[f], gt.reshape(-1, *gt.shape[2:])[f])
                    for f in range(kp.shape[0] * kp.shape[1])])
    assert np.array_equal(idx, gold["sem_ref_idx"])
    with pytest.raises(IndexError):
        C.voxel_chamfer_frame(np.zeros((8, 8, 8), np.float32), np.ones((8, 8, 8), np.float32))
