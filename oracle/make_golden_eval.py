"""Pin oracle/nm_oracle_eval.py against the reference's own functions and write tests/golden/eval_retarget.npz.

BUILD container only (imports /root/reference):  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_eval.py

`utils/eval_utils.py` imports cleanly.  `vis_retarget.py` needs open3d / matplotlib at import time, so
`extract_skin_weights` is lifted out of its source by `ast` and executed unmodified with the names it uses
(`torch`, `deepcopy`); the FK / skinning statements of its `__main__` block (:279-322) cannot be lifted as a
function and are pinned through the oracle's restatement only (checked here against an independent 4x4
homogeneous-matrix formulation).
"""
from __future__ import annotations

import ast
import json
import os
import sys
from collections import namedtuple
from copy import deepcopy

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import nm_oracle as O            # noqa: E402
from oracle import nm_oracle_eval as E       # noqa: E402
from utils import eval_utils as R            # noqa: E402  (the reference)

GOLD = os.path.join(ROOT, "tests", "golden")
report = {}


def lift(path, name):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "deepcopy": deepcopy, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


torch.manual_seed(0)
rng = np.random.default_rng(0)

# ---- voxel chamfer: gt = voxelized synthetic clips, recon = a noisy soft version of a shifted gt -----------------
B, T, G = 2, 3, 32
vox = torch.from_numpy(np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(300 + b, T, 4000)), G)
                                 for b in range(B)], 0)).float()                       # (B, T, 1, G, G, G)
soft = torch.roll(vox, shifts=(1, -1), dims=(3, 5)) * 0.7 + torch.rand(vox.shape) * 0.35
soft = soft.half().float()                                 # stored as fp16 in the fixture: quantise BEFORE the reference sees it
ref_recon = soft.clone()
ref_log = R.voxel_chamfer_distance(None, dict(voxel=vox.clone(), recon=ref_recon))    # mutates ref_recon
ora_log = E.voxel_chamfer_distance(None, vox, soft)
per, binar = E.voxel_chamfer_per_frame(vox.reshape(B * T, G, G, G), soft.reshape(B * T, G, G, G))
assert torch.equal(binar.reshape(ref_recon.shape), ref_recon), "in-place binarisation differs"
assert np.allclose(np.array(ref_log["scores"]), np.array(ora_log["scores"]), rtol=0, atol=0)
assert ref_log["scores_log"] == ora_log["scores_log"]
report["voxel_chamfer"] = "oracle == reference (bit-identical) on %d frames at G=%d" % (B * T, G)

# ---- semantic scores ----------------------------------------------------------------------------------------------
Bs, Ts, K, Kgt = 3, 4, 24, 17
kp = torch.rand(Bs, Ts, K, 4) * 2 - 1
kp[..., 3] = torch.rand(Bs, Ts, K)                        # intensities, ~20 % below the 0.2 threshold
gtk = torch.rand(Bs, Ts, Kgt, 3) * 2 - 1
ref_kp = kp.clone()
ref_sem = R.semantic_scores(None, dict(keypoints=ref_kp, gt_keypoints=gtk.clone()))
ora_sem = E.semantic_scores(None, kp, gtk)
masked, idx, hist = E.semantic_nearest(kp, gtk)
assert torch.equal(masked, ref_kp)
assert np.array_equal(ref_sem["scores"], ora_sem["scores"]) and ref_sem["scores_log"] == ora_sem["scores_log"]
report["semantic"] = "oracle == reference (bit-identical)"

# ---- skin weights: reference function lifted from vis_retarget.py -------------------------------------------------
ref_skin = lift(os.path.join(REF, "vis_retarget.py"), "extract_skin_weights")
Priority = namedtuple("Priority", ["values", "indices"])
aff = torch.rand(2, 24, 24, 1)
A, prio, parents = O.skeleton_from_affinity(aff)
order = prio.indices.long()
parents_t = torch.as_tensor(parents).long()
root = int(order[0])
z = np.load(os.path.join(GOLD, "voxelize_obj.npz"))
pts = O.episodic_normalization(z["obj_points_f32"][None], 0.8)[0][::6].copy()          # float64 (2078, 3), like o3d points
sk_kp = torch.rand(24, 4) * 1.2 - 0.6
sk_kp[:, 3] = torch.rand(24)
sk_kp[root, 3] = 0.9                                       # the reference spins forever on an invalid root
ref_w = ref_skin(torch.zeros(1), Priority(None, order), parents_t, pts.astype(np.float32), sk_kp.clone(), 8.0, 0.2)
ora_w = E.extract_skin_weights(order, parents_t, pts.astype(np.float32), sk_kp, 8.0, 0.2)
d = float(np.abs(ref_w - ora_w).max())
assert d <= 1e-6, d
report["skin_weights"] = "oracle vs reference max |diff| = %.2e on %d points" % (d, len(pts))

# ---- FK + LBS: oracle vs an independent homogeneous-matrix formulation --------------------------------------------
Tn = 5
Rm = O.rot6d_to_matrix(torch.randn(Tn * 24, 6)).reshape(Tn, 24, 3, 3)
Rinv = O.rot6d_to_matrix(torch.randn(24, 6)).reshape(24, 3, 3).transpose(1, 2)
off = torch.randn(24, 3) * 0.1
rootp = torch.randn(Tn, 3) * 0.1
pos = E.retarget_fk(Rm, off, rootp, order, parents_t, clip=True)
chk = torch.zeros(Tn, 24, 3)
for t in range(Tn):
    chk[t, root] = rootp[t]
    for i in order[1:]:
        chk[t, i] = torch.bmm(Rm[None, t, int(i)], off[None, int(i), :, None]).squeeze(-1)[0] + chk[t, int(parents_t[i])]
assert torch.allclose(pos, chk.clip(-1, 1), atol=1e-6)
T3x4 = torch.cat([Rm, pos[..., None]], -1).numpy()
joints = sk_kp[:, :3].numpy()
lbs = E.linear_blend_skinning(pts, joints, Rinv.numpy(), T3x4, ref_w)
M = np.zeros((Tn, 24, 4, 4)); M[:, :, :3] = T3x4; M[:, :, 3, 3] = 1
Binv = np.zeros((24, 4, 4)); Binv[:, :3, :3] = Rinv.numpy(); Binv[:, 3, 3] = 1
Binv[:, :3, 3] = -np.einsum('kij,kj->ki', Rinv.numpy().astype(np.float64), joints.astype(np.float64))
ph = np.concatenate([pts, np.ones((len(pts), 1))], 1)
alt = np.einsum('nk,tkij,kjl,nl->tni', ref_w.astype(np.float64), M, Binv, ph)[..., :3]
assert np.abs(alt - lbs).max() <= 1e-9
report["fk_lbs"] = "oracle == homogeneous-matrix formulation (<= 1e-9)"

# ---- key-frame interpolation: the loop of vis_interpolation.py:86-135 executed verbatim on the reference network ----
import pickle           # noqa: E402
import textwrap         # noqa: E402
from torch.distributions import Normal   # noqa: E402
from model.neural_marionette import NeuralMarionette   # noqa: E402  (the reference)

opt = pickle.load(open(os.path.join(REF, "pretrained/aist/opt.pickle"), "rb"))
hp = O.default_hparams()
sd = O.synthetic_state_dict(hp, seed=21)
network = NeuralMarionette(opt).eval()
network.anneal(1)
network.load_state_dict(sd, strict=True)
Ti, sample_num, sample_rate, Z = 12, 16, 5, hp.nlatent_kypt
gk = torch.Generator().manual_seed(7)
ikp = torch.cat([torch.rand(1, Ti, 24, 3, generator=gk) * 1.2 - 0.6, torch.rand(1, Ti, 24, 1, generator=gk)], dim=-1)
ikp = ikp + 0.02 * torch.arange(Ti).float()[None, :, None, None]
ieps = torch.randn(Ti, 2, sample_num, Z, generator=gk)
queue = []
for t in range(Ti):
    queue += [ieps[t, 0], ieps[t, 1]] if (t % sample_rate == 0 or t == Ti - 1) else [ieps[t, 0]]


class QueuedNormal(Normal):
    """Normal whose rsample() consumes the injected draws: loc + eps * scale is what Normal.rsample computes."""
    def rsample(self, sample_shape=torch.Size()):
        return self.loc + queue.pop(0) * self.scale


lines = open(os.path.join(REF, "vis_interpolation.py")).read().split("\n")[85:136]     # :86-136
assert lines[0].strip().startswith("_ = network.dyna_module.encode") and "selected_keypoints[0, :, :, -1]" in lines[-1]
ns = dict(network=network, torch=torch, Normal=QueuedNormal, keypoints=ikp.clone(), T=Ti, K=24, sample_num=sample_num,
          sample_rate=sample_rate)
with torch.no_grad():
    ns["affinity"] = network.kypt_detector.get_affinity()
    exec(textwrap.dedent("\n".join(lines)), ns)
    assert not queue
    skeleton = (network.dyna_module.A, network.dyna_module.priority, network.dyna_module.parents)
    ora_sel, picks = O.dyna_interpolate(ikp, skeleton, sd, hp, sample_num, sample_rate, ieps)
d = float((ns["selected_keypoints"] - ora_sel).abs().max())
assert d <= 5e-6, d
report["interpolation"] = "oracle vs reference loop max |diff| = %.2e (T=%d, %d hypotheses, key frames every %d), picks %s" % (
    d, Ti, sample_num, sample_rate, picks)
np.savez_compressed(os.path.join(GOLD, "interpolation.npz"), seed=21, kp=ikp.numpy(), eps=ieps.numpy(),
                    sample_num=sample_num, sample_rate=sample_rate, selected=ns["selected_keypoints"].numpy(),
                    picks=np.array(picks))

np.savez_compressed(
    os.path.join(GOLD, "eval_retarget.npz"),
    vc_soft=soft.numpy().astype(np.float16), vc_seeds=np.array([300, T, 4000, G, B]),
    vc_ref_scores=np.array(ref_log["scores"]), vc_ref_log=np.array(ref_log["scores_log"]), vc_ref_per_frame=per,
    sem_kp=kp.numpy(), sem_gt=gtk.numpy(), sem_ref_masked=ref_kp.numpy(), sem_ref_scores=ref_sem["scores"],
    sem_ref_log=np.array(ref_sem["scores_log"]), sem_ref_idx=idx.numpy(),
    sk_points=pts.astype(np.float32), sk_kp=sk_kp.numpy(), sk_order=order.numpy(), sk_parents=parents_t.numpy(),
    sk_ref_weights=ref_w,
    fk_R=Rm.numpy(), fk_Rinv=Rinv.numpy(), fk_off=off.numpy(), fk_root=rootp.numpy(), fk_pos=pos.numpy(), lbs_out=lbs)
json.dump(report, open(os.path.join(GOLD, "EVAL_PIN.json"), "w"), indent=1)
print(json.dumps(report, indent=1))
