"""List the remaining element-wise passes (affine_act / upsample2x / gn_scale_shift) of one detector forward
with their tensor shapes and CUDA-event times."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neural_marionette_b200 as nm
from neural_marionette_b200 import ops
from oracle import nm_oracle as O

B, T, G = 16, 20, 64
hp = O.default_hparams(grid_size=G)
net = nm.NeuralMarionette(hp); net.load_state_dict(O.synthetic_state_dict(hp, 0)); net = net.cuda().eval(); net.anneal(1)
raw = np.stack([O.synthetic_clip(100 + b % 8, T, 20000) for b in range(B)], 0)
vox = nm.voxelize_raw_clips(raw, G)
log = []
def wrap(name):
    fn = getattr(ops, name)
    def inner(x, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(x, *a, **k); e1.record()
        log.append((name, tuple(x.shape), k.get("x2") is not None, e0, e1))
        return r
    setattr(ops, name, inner)
for n_ in ("affine_act", "upsample2x", "gn_scale_shift", "conv_transpose3d"):
    wrap(n_)
with torch.no_grad():
    net.kypt_detector(vox); log.clear()
    net.kypt_detector(vox)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, shape, dual, e0, e1 in log:
    k = (name, shape, dual)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = 0
for (name, shape, dual), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tot += ms
    print(f"{name:18s} {str(shape):28s} dual={int(dual)} x{cnt:2d}  {ms:7.3f} ms")
print("total %.3f ms" % tot)
