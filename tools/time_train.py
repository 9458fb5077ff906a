"""Time the config #4 training step (D-FAUST shape: T = 10, 24 clips per GPU, grid 64^3; reference train.py:387-409)
with CUDA events: forward, backward, optimizer, and the per-kernel device time from torch.profiler (our kernels only
launch through the C ABI; the profiler just lists what ran on the stream).

    python tools/time_train.py [clips] [frames] [grid] [--profile]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nm_oracle as O            # synthetic weights / clips only  # noqa: E402
from oracle import nm_oracle_grad as OG      # the loss weights of train.py    # noqa: E402
import neural_marionette_b200 as nm          # noqa: E402
from neural_marionette_b200 import ops, optim  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 24
T = int(args[1]) if len(args) > 1 else 10
G = int(args[2]) if len(args) > 2 else 64
hp = O.default_hparams(grid_size=G, Tcond=3, Ttot=10)
net = nm.NeuralMarionette(hp)
net.load_state_dict(O.synthetic_state_dict(hp, seed=0), strict=True)
net = net.cuda().train()
net.anneal(1)
opt = optim.FusedAdam(net.kypt_detector.parameters(), lr=4e-4, owner=net)
raw = np.stack([O.synthetic_clip(1000 + b % 8, T, 20000) for b in range(B)], 0)
raw_dev = torch.from_numpy(raw).cuda()


def step(events=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    vox = ops.normalize_voxelize(raw_dev, G, check=False)
    opt.zero_grad()
    out = net.kypt_detector(vox)
    loss = OG.detector_loss(out, recon_only=False)
    ev[1].record()
    loss.backward()
    ev[2].record()
    ok = opt.step()
    ev[3].record()
    if events is not None:
        events.append(ev)
    return float(loss.detach()), ok


for _ in range(2):
    print("warm-up", step())
torch.cuda.synchronize()
evs = []
t0 = time.perf_counter()
for _ in range(3):
    l = step(evs)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
f = np.mean([e[0].elapsed_time(e[1]) for e in evs])
b = np.mean([e[1].elapsed_time(e[2]) for e in evs])
o = np.mean([e[2].elapsed_time(e[3]) for e in evs])
print(f"{B} clips x {T} frames @ {G}^3: forward {f:.1f} ms, backward {b:.1f} ms, optimizer {o:.1f} ms, total "
      f"{f + b + o:.1f} ms (wall {wall * 1e3:.1f} ms) = {B * T / (f + b + o) * 1e3:.0f} frames/s; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; loss {l[0]:.4f}; grad scale {ops.grad_scale()}")
if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
    tot = sum(r.device_time_total for r in rows)
    print(f"device time of one step: {tot / 1e3:.1f} ms")
    for r in rows[:45]:
        print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}% x{r.count:4d}  {r.key[:110]}")
if "--gaps" in sys.argv:
    # where the GPU idles inside a step: gaps between consecutive kernels on the device timeline, by the kernel that follows
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
                key=lambda t: t[0])
    span = ev[-1][1] - ev[0][0]
    busy = sum(e[1] - e[0] for e in ev)
    gaps = {}
    last_end = ev[0][1]
    for st, en, name in ev[1:]:
        if st > last_end:
            key = name.replace("void ", "").replace("(anonymous namespace)::", "").split("(")[0][:60]
            g = gaps.setdefault(key, [0, 0.0])
            g[0] += 1
            g[1] += st - last_end
        last_end = max(last_end, en)
    print(f"device span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, idle {(span - busy) / 1e3:.1f} ms over {len(ev)} launches")
    for k, (c, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  idle before {k:60s} x{c:4d} {t / 1e3:7.2f} ms ({t / max(c, 1):6.1f} us each)")
if "--graph" in sys.argv:
    # the same step captured in a CUDA graph (graph.CapturedTrainStep: device-side Adam step counter / skip flag)
    from neural_marionette_b200 import graph
    cap = graph.CapturedTrainStep(net.kypt_detector, opt, lambda out: OG.detector_loss(out, recon_only=False), raw_dev, G, warmup=2)
    for _ in range(2):
        cap(raw_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        loss = cap(raw_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    opt.sync_counters()
    print(f"CUDA-graph replay: {ms:.1f} ms per step = {B * T / ms * 1e3:.0f} frames/s; loss {float(loss):.4f}; "
          f"steps applied {opt.steps}, skipped {opt.skipped}")
