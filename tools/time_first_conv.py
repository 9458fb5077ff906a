"""Timing + parity of the CoordConv first layer on bench-shaped occupancy (n frames at 64^3)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neural_marionette_b200 as nm
from neural_marionette_b200 import ops
from oracle import nm_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 320
G, T = 64, 20
raw = np.stack([O.synthetic_clip(100 + b, T, 20000) for b in range(2)], 0)
vox = nm.voxelize_raw_clips(raw, G)                  # (2, T, 1, G, G, G)
occ = vox.reshape(-1, G, G, G)
occ = occ.repeat((n + occ.shape[0] - 1) // occ.shape[0], 1, 1, 1)[:n].contiguous()
for cout in (32, 64):
    torch.manual_seed(cout)
    conv = torch.nn.Conv3d(4, cout, 5, 1, 2).cuda()
    out = ops.first_conv(occ, conv)
    ref = conv(O.add_coord_channels(occ[:2, None].cpu()).cuda()).permute(0, 2, 3, 4, 1)
    err = (out[:2].float() - ref).abs().max().item() / ref.abs().max().item()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): ops.first_conv(occ, conv)
    ev0.record()
    for _ in range(10): ops.first_conv(occ, conv)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    gb = out.numel() * 2 / 1e9
    gn = torch.nn.GroupNorm(cout // 16, cout).cuda()
    for _ in range(3): ops.first_conv(occ, conv, gn)
    ev0.record()
    for _ in range(10): ops.first_conv(occ, conv, gn)
    ev1.record(); torch.cuda.synchronize()
    print(f"   with GroupNorm statistics + finalize: {ev0.elapsed_time(ev1) / 10:.3f} ms")
    print(f"Cout={cout} n={n}: {ms:.3f} ms  {ms / n * 1e3:.2f} us/frame  write {gb / ms * 1e3:.0f} GB/s  rel err {err:.2e}")
