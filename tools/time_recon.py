"""Timing of nm_final_recon on the bench shape (n frames, 64^3, 32 channels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_marionette_b200 import ops
n, G, C, T = int(sys.argv[1]) if len(sys.argv) > 1 else 320, 64, 32, 20
x = torch.randn(n, G, G, G, C, device="cuda").to(ops.ACT_DTYPE)
a = 0.5 + torch.rand(n, C, device="cuda"); b = torch.randn(n, C, device="cuda")
conv = torch.nn.Conv3d(C, 1, 1).cuda()
first = (torch.rand(n // T, G, G, G, device="cuda") < 0.1).float()
tgt = (torch.rand(n, G, G, G, device="cuda") < 0.1).float()
out = torch.empty(n, G, G, G, device="cuda"); bce = torch.empty(n, device="cuda")
fn = lambda: ops.final_recon(x, a, b, conv, first, T, 10.0, 0.5, target=tgt, out=out, bce_out=bce)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"final_recon n={n}: {ms:.3f} ms  {(x.numel() * 2 + 3 * n * G ** 3 * 4) / ms / 1e9:.2f} TB/s")
