"""Timing of the config #4 backward bricks on the decoder shapes (80 frames): data gradient (forward tcgen05 kernels),
weight gradient (tcgen05 kernel; NM_TIME_SLAB=1 also times the first mma.sync version), GroupNorm + LeakyReLU backward.  CUDA events, 3 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_marionette_b200 import ops    # noqa: E402


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


n = 80
SHAPES = (("dec.8", 64, 64, 32, 80), ("dec.11", 64, 32, 32, 80), ("dec.1", 32, 128, 64, 80), ("dec.4", 32, 64, 64, 80),
          ("enc.res2.0", 32, 32, 64, 80), ("enc.res5.3", 16, 128, 128, 240), ("st.res2.3", 32, 128, 128, 24),
          ("st.res5.3", 16, 256, 256, 24))
for name, grid, cin, cout, n in SHAPES:
    conv = torch.nn.Conv3d(cin, cout, 3, 1, 1).cuda()
    x = torch.randn(n, grid, grid, grid, cin, device="cuda", dtype=torch.float16)
    gy = torch.randn(n, grid, grid, grid, cout, device="cuda", dtype=torch.float16)
    flops = 2.0 * n * grid ** 3 * cin * cout * 27
    t_f = timed(lambda: ops.conv3d(x, conv))
    t_d = timed(lambda: ops.conv3d_input_grad(gy, conv))
    t_w = timed(lambda: ops.conv3d_weight_grad(x, gy, impl="tc"))
    t_ws = timed(lambda: ops.conv3d_weight_grad(x, gy, impl="slab"), reps=2) if os.environ.get("NM_TIME_SLAB") else float("nan")
    gn = torch.nn.GroupNorm(cout // 16, cout).cuda()
    t_g = timed(lambda: ops.groupnorm_backward(gy, gy, gn))
    print(f"{name} {cin}->{cout} @ {grid}^3 x {n} frames: fwd {t_f:.2f} ms ({flops / t_f / 1e9:.0f} TF/s)  dgrad {t_d:.2f} ms "
          f"({flops / t_d / 1e9:.0f} TF/s)  wgrad tcgen05 {t_w:.2f} ms ({flops / t_w / 1e9:.0f} TF/s)  [mma.sync {t_ws:.2f} ms]  "
          f"GN+LReLU bwd {t_g:.2f} ms "
          f"({5 * gy.numel() * 2 / t_g / 1e6:.0f} GB/s)")
