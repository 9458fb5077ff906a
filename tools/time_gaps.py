"""GPU idle time inside ONE inference step of bench.py's headline workload (64 clips x 20 frames): gaps between
consecutive kernels on the device timeline (torch.profiler, CUDA activities only), by the kernel that follows the gap."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_marionette_b200 as nm          # noqa: E402
from neural_marionette_b200 import ops       # noqa: E402
from oracle import nm_oracle as O            # synthetic weights / clips only  # noqa: E402

_args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, T, G = int(_args[0]) if _args else 64, 20, 64
hp = O.default_hparams(grid_size=G)
net = nm.NeuralMarionette(hp)
net.load_state_dict(O.synthetic_state_dict(hp, 0))
net = net.cuda().eval()
net.anneal(1)
raw = torch.from_numpy(np.stack([O.synthetic_clip(1000 + b % 8, T, 20000) for b in range(B)], 0)).cuda()


def step():
    return net.kypt_detector(ops.normalize_voxelize(raw, G, check=False))


with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
            key=lambda t: t[0])
# two streams overlap: merge intervals to get the busy time
busy, cur_s, cur_e = 0.0, ev[0][0], ev[0][1]
gaps = {}
for st, en, name in ev[1:]:
    if st > cur_e:
        busy += cur_e - cur_s
        key = name.replace("void ", "").replace("(anonymous namespace)::", "").split("(")[0][:60]
        g = gaps.setdefault(key, [0, 0.0])
        g[0] += 1
        g[1] += st - cur_e
        cur_s, cur_e = st, en
    else:
        cur_e = max(cur_e, en)
busy += cur_e - cur_s
span = ev[-1][1] - ev[0][0]
print(f"{B} clips: device span {span / 1e3:.1f} ms, busy (union over streams) {busy / 1e3:.1f} ms, idle {(span - busy) / 1e3:.1f} ms over {len(ev)} launches")
for k, (c, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  idle before {k:60s} x{c:4d} {t / 1e3:7.2f} ms ({t / max(c, 1):6.1f} us each)")
if "--kernels" in sys.argv:
    # device time by kernel (torch.profiler durations; in-step, warm caches - the ncu launch lists are cold and serialised)
    tot = {}
    for st, en, name in ev:
        key = name.replace("void ", "").replace("(anonymous namespace)::", "")[:110]
        t = tot.setdefault(key, [0, 0.0])
        t[0] += 1
        t[1] += en - st
    total = sum(v[1] for v in tot.values())
    print(f"kernel time (sum over both streams) {total / 1e3:.1f} ms")
    for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:36]:
        print(f"  {t / 1e3:8.2f} ms {100 * t / total:5.1f}% x{c:4d}  {k}")
