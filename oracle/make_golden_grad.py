"""Pin the gradient oracle (oracle/nm_oracle_grad.py) against the reference's own backward pass and write
tests/golden/detector_grad_g32.npz.  BUILD container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_grad.py

Config #4 shape scaled to the CPU: 2 clips x 3 frames, grid 32^3, reference `KyptDetector` in train mode,
loss = 100 * recon_loss (and, second entry, the full stage-1 weighted sum of train.py:177-181), `loss.backward()`.
The fixture keeps, per parameter tensor, the gradient's L2 norm, its sum and 8 sampled entries (fixed positions),
plus the complete gradients of the small tensors — enough to test a CUDA backward layer by layer.
"""
from __future__ import annotations

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

from oracle import nm_oracle as O            # noqa: E402
from oracle import nm_oracle_grad as OG      # noqa: E402
from model.neural_marionette import NeuralMarionette   # noqa: E402  (the reference)

G, B, T, SEED, VSEED = 32, 2, 3, 61, 400
hp = O.default_hparams(grid_size=G, Tcond=3, Ttot=10)
sd = O.synthetic_state_dict(hp, seed=SEED)
opt = pickle.load(open(os.path.join(REF, "pretrained/aist/opt.pickle"), "rb"))
opt.grid_size = G
vox = torch.from_numpy(np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(VSEED + b, T, 5000)), G)
                                 for b in range(B)], 0)).float()


def sample_positions(numel, n=8):
    return (np.arange(n, dtype=np.int64) * 2654435761 % max(numel, 1)).astype(np.int64)


fixture, report = {}, {}
for tag, recon_only in (("recon", True), ("full", False)):
    net = NeuralMarionette(opt)
    net.load_state_dict(sd, strict=True)
    net.train()
    net.anneal(1)
    out = net.kypt_detector(vox)
    loss = OG.detector_loss(out, recon_only)
    loss.backward()
    ref = {"kypt_detector." + k: p.grad for k, p in net.kypt_detector.named_parameters() if p.grad is not None}
    ora_loss, ora = OG.detector_gradients(vox, sd, hp, recon_only)
    assert float(ora_loss) == float(loss), (float(ora_loss), float(loss))
    assert set(ref) == set(ora), set(ref) ^ set(ora)
    worst = max(float((ref[k] - ora[k]).abs().max() / (ref[k].abs().max() + 1e-30)) for k in ref)
    assert worst <= 1e-5, worst            # recon-only: bit-identical; the full sum differs by fp32 summation order
    report[tag] = dict(loss=float(loss), tensors=len(ref), params=int(sum(g.numel() for g in ref.values())),
                       worst_rel_diff_oracle_vs_reference=worst)
    keys = sorted(ref)
    fixture[f"{tag}_loss"] = np.float64(float(loss))
    fixture[f"{tag}_keys"] = np.array(keys)
    fixture[f"{tag}_norm"] = np.array([float(ref[k].double().norm()) for k in keys])
    fixture[f"{tag}_sum"] = np.array([float(ref[k].double().sum()) for k in keys])
    fixture[f"{tag}_samples"] = np.stack([ref[k].reshape(-1)[torch.from_numpy(sample_positions(ref[k].numel()))].numpy()
                                          for k in keys])
    for k in keys:
        if ref[k].numel() <= 4096 and tag == "recon":
            fixture["recon_grad::" + k] = ref[k].numpy()

fixture["meta"] = np.array([G, B, T, SEED, VSEED, 5000])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "detector_grad_g32.npz"), **fixture)
json.dump(report, open(os.path.join(ROOT, "tests", "golden", "GRAD_PIN.json"), "w"), indent=1)
print(json.dumps(report, indent=1))
