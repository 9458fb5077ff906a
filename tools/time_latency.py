"""Config #1 shape (one 20-frame clip, 20 000 points per frame, grid 64^3): latency of normalise + voxelize +
KyptDetector.forward, eager vs CUDA-graph replay (neural_marionette_b200/graph.py).  CUDA events, 5 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_marionette_b200 as nm               # noqa: E402
from neural_marionette_b200 import graph as NG, ops    # noqa: E402
from oracle import nm_oracle as O                 # noqa: E402  (synthetic weights / inputs only)

G, T = 64, 20
hp = O.default_hparams(grid_size=G)
net = nm.NeuralMarionette(hp)
net.load_state_dict(O.synthetic_state_dict(hp, 0))
net = net.cuda().eval()
net.anneal(1)
det = net.kypt_detector


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for B in (1, 2, 4):
    raw = torch.stack([torch.from_numpy(O.synthetic_clip(100 + b, T, 20000)) for b in range(B)]).cuda()
    with torch.no_grad():
        eager = timed(lambda: det(ops.normalize_voxelize(raw, G, check=False)))
    cap = NG.capture_detector_from_points(det, raw, G)
    replay = timed(lambda: cap(raw, clone=False))
    print(f"B={B}: eager {eager:.2f} ms/clip-batch ({B * T / eager * 1e3:.0f} frames/s)  "
          f"graph replay {replay:.2f} ms ({B * T / replay * 1e3:.0f} frames/s)  x{eager / replay:.2f}")
