"""Timing of one k3 conv layer through ops.conv3d, plain and with the fused input GroupNorm + LeakyReLU.
usage: time_conv.py n grid cin cout [fused_only]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_marionette_b200 import ops

n, grid, cin, cout = (int(v) for v in sys.argv[1:5])
fused_only = len(sys.argv) > 5
torch.manual_seed(0)
conv = torch.nn.Conv3d(cin, cout, 3, 1, 1).cuda()
gn = torch.nn.GroupNorm(cout // 16, cout).cuda()
x = (torch.randn(n, grid, grid, grid, cin, device="cuda") * 1.5 + 0.3).to(ops.ACT_DTYPE)
a = 0.5 + torch.rand(n, cin, device="cuda")
b = torch.randn(n, cin, device="cuda")


def timeit(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flops = 2.0 * n * grid ** 3 * cin * cout * 27
if not fused_only:
    t = timeit(lambda: ops.conv3d(x, conv, gn))
    print(f"plain : {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
    t = timeit(lambda: ops.affine_act(x, a, b, True))
    print(f"affine: {t:.3f} ms")
if ops.can_fuse_input(x, conv):
    t = timeit(lambda: ops.conv3d(x, conv, gn, in_affine=(a, b, True)))
    print(f"fused : {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
else:
    print("fused : not supported for this shape")
