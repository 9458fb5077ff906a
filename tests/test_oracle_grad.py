"""Config #4 groundwork: the gradient oracle (oracle/nm_oracle_grad.py = torch.autograd over the oracle's functional
forward) replays tests/golden/detector_grad_g32.npz, which oracle/make_golden_grad.py wrote from the REFERENCE's own
`loss.backward()` (train-mode KyptDetector, 2 clips x 3 frames, grid 32^3).  A CUDA backward is tested against the
same file."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nm_oracle as O
from oracle import nm_oracle_grad as OG


@pytest.mark.parametrize("tag,tensors,params", [("recon", 314, 8555940), ("full", 315, 8557044)])
def test_gradient_oracle_replays_reference_backward(golden_dir, tag, tensors, params):
    z = np.load(os.path.join(golden_dir, "detector_grad_g32.npz"))
    G, B, T, seed, vseed, N = (int(v) for v in z["meta"])
    hp = O.default_hparams(grid_size=G, Tcond=3, Ttot=10)
    sd = O.synthetic_state_dict(hp, seed=seed)
    vox = torch.from_numpy(np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(vseed + b, T, N)), G)
                                     for b in range(B)], 0)).float()
    loss, grads = OG.detector_gradients(vox, sd, hp, recon_only=(tag == "recon"))
    keys = [str(k) for k in z[f"{tag}_keys"]]
    assert sorted(grads) == keys and len(keys) == tensors
    assert sum(g.numel() for g in grads.values()) == params          # 8 557 044 gradient values in stage 1
    assert abs(float(loss) - float(z[f"{tag}_loss"])) <= 1e-5 * abs(float(z[f"{tag}_loss"]))
    pos = lambda n: torch.from_numpy((np.arange(8, dtype=np.int64) * 2654435761 % max(n, 1)).astype(np.int64))
    for i, k in enumerate(keys):
        g = grads[k]
        scale = max(float(z[f"{tag}_norm"][i]), 1e-12)
        assert abs(float(g.double().norm()) - float(z[f"{tag}_norm"][i])) <= 2e-5 * scale, k
        got = g.reshape(-1)[pos(g.numel())].numpy()
        assert np.abs(got - z[f"{tag}_samples"][i]).max() <= 2e-5 * max(float(g.abs().max()), 1e-12), k
        if tag == "recon" and ("recon_grad::" + k) in z.files:
            assert np.abs(g.numpy() - z["recon_grad::" + k]).max() <= 2e-5 * max(float(g.abs().max()), 1e-12), k


def test_gradient_pin_report(golden_dir):
    rep = json.load(open(os.path.join(golden_dir, "GRAD_PIN.json")))
    assert rep["recon"]["worst_rel_diff_oracle_vs_reference"] == 0.0
    assert rep["full"]["worst_rel_diff_oracle_vs_reference"] <= 1e-5
