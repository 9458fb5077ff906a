"""Timing of the memory-pipe conv kernel on the pool / 1x1 shapes of the encoder (n frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_marionette_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 320
for grid, cin, cout, k, s, fused in ((64, 32, 32, 2, 2, True), (32, 32, 64, 1, 1, False), (32, 64, 64, 2, 2, True)):
    conv = torch.nn.Conv3d(cin, cout, k, s, 0).cuda()
    gn = torch.nn.GroupNorm(cout // 16, cout).cuda()
    x = (torch.randn(n, grid, grid, grid, cin, device="cuda")).to(ops.ACT_DTYPE)
    a = 0.5 + torch.rand(n, cin, device="cuda"); b = torch.randn(n, cin, device="cuda")
    ia = (a, b, True) if fused else None
    if fused and cin == 64:
        ia = (a, b, False, x.clone(), a, b)
    fn = lambda: ops.conv3d(x, conv, gn, in_affine=ia)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nb = x.numel() * 2 * (2 if (fused and cin == 64) else 1) + n * (grid // s) ** 3 * cout * 2
    print(f"grid {grid} {cin}->{cout} k{k}s{s} fused={fused}: {ms:.3f} ms  {nb / ms / 1e9:.2f} TB/s (incl. finalize)")
