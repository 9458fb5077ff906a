"""Time the §8f#4 kernels at config #2 sizes: voxel chamfer over 1 280 frames at 64^3, skin weights + LBS on the
12 465-vertex demo mesh size x 40 frames.  CUDA events, 3 warm-ups."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_marionette_b200 import ops            # noqa: E402
from oracle import nm_oracle as O                 # noqa: E402  (input generator only)


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


G, n = 64, 1280
clips = [O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(500 + b, 20, 20000)), G)[:, 0] for b in range(4)]
vox = torch.from_numpy(np.concatenate(clips, 0)).float().cuda().repeat(n // 80, 1, 1, 1).contiguous()
rec = torch.roll(vox, shifts=(1, 1), dims=(1, 2)).contiguous()
ms = timed(lambda: ops.voxel_chamfer(vox, rec, binarize=False, frames_per_call=256))
occ = float(vox.flatten(1).sum(1).mean())
print(f"voxel_chamfer: {n} frames @ {G}^3, {occ:.0f} occupied voxels/frame: {ms:.2f} ms "
      f"({n / ms * 1e3:.0f} frames/s; grids read at {2 * n * G**3 * 4 / ms / 1e6:.0f} GB/s incl. the search)")

N, K, T = 12465, 24, 40
pts = (torch.rand(N, 3) * 1.6 - 0.8).cuda()
kp = torch.rand(K, 4).cuda()
kp[:, 3] = 0.9
par = torch.tensor([0] + [max(0, i - 1) for i in range(1, K)], dtype=torch.int32).cuda()
ms = timed(lambda: ops.skin_weights(pts, kp, par, 0, 8.0, 0.2))
print(f"skin_weights: {N} points x {K} joints: {ms * 1e3:.1f} us")
skin = ops.skin_weights(pts, kp, par, 0, 8.0, 0.2)[0]
R = O.rot6d_to_matrix(torch.randn(T * K, 6)).reshape(T, K, 3, 3).cuda()
T3x4 = torch.cat([R, torch.rand(T, K, 3, 1).cuda()], -1).contiguous()
ms = timed(lambda: ops.linear_blend_skinning(pts, kp[:, :3].contiguous(), R[0].contiguous(), T3x4, skin))
print(f"linear_blend_skinning: {T} frames x {N} points: {ms * 1e3:.1f} us")
