"""Model façade (drop-in for the reference's model/neural_marionette.py): owns the detector and the
dynamics module, routes `forward` (track) and `generate` (condition + roll-out + decode)."""
from __future__ import annotations

import contextlib

import torch
import torch.nn as nn

from .hsvrnn_bvh import HSVRNNBVH
from .kypt_detector import KyptDetector


class NeuralMarionette(nn.Module):
    def __init__(self, options=None):
        super().__init__()
        self.options = options
        self.kypt_detector = KyptDetector(options)
        self.Tcond = options.Tcond
        self.dyna_module = HSVRNNBVH(options)
        self.current_actives = {"detector": True, "learner": True}
        self.transition_type = options.transition_type

    def anneal(self, nepoch, nbatch=None):
        if nbatch is None:           # once per epoch
            self.kypt_detector.anneal(nepoch)

    def control_active(self, module_actives):
        """Freeze / unfreeze a sub-model when its flag flips (reference neural_marionette.py:22-32)."""
        owners = {"detector": self.kypt_detector, "learner": self.dyna_module}
        for name, owner in owners.items():
            want = module_actives[name]
            if self.current_actives[name] == want:
                continue
            for p in owner.parameters():
                p.requires_grad = want
            self.current_actives[name] = want

    def forward(self, vox_seq, module_actives=None):
        """vox_seq (B, T, 1, G, G, G).  Detector (no grad if inactive) then `dyna_module.encode`."""
        log = {}
        if module_actives["detector"] or module_actives["learner"]:
            # the caller's grad mode is inherited when the detector is active (reference neural_marionette.py:40-44)
            ctx = contextlib.nullcontext() if module_actives["detector"] else torch.no_grad()
            with ctx:
                det = self.kypt_detector(vox_seq)
            log.update(det)
        if module_actives["learner"]:
            log.update(self.dyna_module.encode(det["keypoints"].detach(), det["affinity"].detach()))
        return log

    def generate(self, vox_seq, module_actives=None, eps_cond=None, eps_gen=None):
        """Detector on the first Tcond frames, latent roll-out to T frames, decode the generated keypoints
        (reference neural_marionette.py:58-103)."""
        T = vox_seq.shape[1]
        assert self.Tcond < T
        log = {}
        if not module_actives["learner"]:
            return log
        det = self.kypt_detector(vox_seq[:, :self.Tcond].contiguous())
        keypoints = det["keypoints"]
        dyn = self.dyna_module.generate(keypoints, det.get("affinity"), Ttot=T, Tcond=self.Tcond,
                                        eps_cond=eps_cond, eps_gen=eps_gen)
        gen_kp = dyn["keypoints_gen"]
        gen = self.kypt_detector.decode_from_dyna(gen_kp, det["first_feature"], vox_seq[:, 0])["gen"]
        log.update(gen=torch.cat([det["recon"][:, :self.Tcond], gen], dim=1),
                   keypoints=torch.cat([keypoints[:, :self.Tcond], gen_kp], dim=1), A_hats=None)
        return log
