"""Per-tensor comparison of the CUDA training step's gradients with torch.autograd over the fp32 CPU oracle
(oracle/nm_oracle_grad.py, pinned bit-exactly to the reference's loss.backward()).  Test-side tool (runs the oracle).

    python tools/grad_report.py [recon|full] [G] [B] [T]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nm_oracle as O            # noqa: E402
from oracle import nm_oracle_grad as OG      # noqa: E402
import neural_marionette_b200 as nm          # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "recon"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T = int(sys.argv[4]) if len(sys.argv) > 4 else 3
hp = O.default_hparams(grid_size=G, Tcond=3, Ttot=10)
sd = O.synthetic_state_dict(hp, seed=61)
vox = torch.from_numpy(np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(400 + b, T, 5000)), G)
                                 for b in range(B)], 0)).float()
loss_ref, ref = OG.detector_gradients(vox, sd, hp, recon_only=(tag == "recon"))
net = nm.NeuralMarionette(hp)
net.load_state_dict(sd, strict=True)
net = net.cuda().train()
net.anneal(1)
out = net.kypt_detector(vox.cuda())
loss = OG.detector_loss(out, recon_only=(tag == "recon"))
loss.backward()
print(f"loss {float(loss):.6f} ref {float(loss_ref):.6f}")
rows = []
for k, p in net.kypt_detector.named_parameters():
    key = "kypt_detector." + k
    if p.grad is None or key not in ref:
        print("missing", key, p.grad is None, key in ref)
        continue
    g, r = p.grad.float().cpu().double(), ref[key].double()
    err = float((g - r).norm() / r.norm().clamp_min(1e-30))
    cos = float((g * r).sum() / (g.norm() * r.norm()).clamp_min(1e-30))
    rows.append((err, cos, g.numel(), float(r.norm()), key))
rows.sort(reverse=True)
for err, cos, numel, rn, key in rows[:40]:
    print(f"{err:9.3e} cos {cos:.5f} numel {numel:7d} |ref| {rn:9.3e} {key}")
errs = np.array([r[0] for r in rows])
big = np.array([r[0] for r in rows if r[2] >= 4096])
print(f"tensors {len(rows)}  median {np.median(errs):.3e}  p90 {np.percentile(errs, 90):.3e}  max {errs.max():.3e};"
      f"  >=4096 elements: median {np.median(big):.3e} max {big.max():.3e}")
allg = torch.cat([p.grad.float().cpu().double().reshape(-1) for k, p in net.kypt_detector.named_parameters() if p.grad is not None])
allr = torch.cat([ref["kypt_detector." + k].double().reshape(-1) for k, p in net.kypt_detector.named_parameters() if p.grad is not None])
print(f"whole gradient vector: rel L2 error {float((allg - allr).norm() / allr.norm()):.3e}  cos {float((allg * allr).sum() / (allg.norm() * allr.norm())):.6f}")
print("grad scale used:", nm.ops.grad_scale())
