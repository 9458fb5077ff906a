"""Timing of NeuralMarionette.generate (config #3 shape) and its parts on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neural_marionette_b200 as nm
from oracle import nm_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, G = 20, 64
hp = O.default_hparams(grid_size=G)
net = nm.NeuralMarionette(hp); net.load_state_dict(O.synthetic_state_dict(hp, 0)); net = net.cuda().eval(); net.anneal(1)
raw = np.stack([O.synthetic_clip(100 + b % 8, T, 20000) for b in range(B)], 0)
vox = nm.voxelize_raw_clips(raw, G)
act = {"detector": True, "learner": True}

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

with torch.no_grad():
    net(vox[:2], act)
    det = net.kypt_detector(vox[:, :5].contiguous())
    kp, aff = det["keypoints"], det["affinity"]
    print(f"B={B}")
    print("generate total        %.1f ms" % timeit(lambda: net.generate(vox, act)))
    print("detector (5 frames)   %.1f ms" % timeit(lambda: net.kypt_detector(vox[:, :5].contiguous())))
    print("dyna.generate 5+15    %.1f ms" % timeit(lambda: net.dyna_module.generate(kp, aff, Ttot=T, Tcond=5)))
    gen_kp = net.dyna_module.generate(kp, aff, Ttot=T, Tcond=5)["keypoints_gen"]
    print("decode_from_dyna (15) %.1f ms" % timeit(lambda: net.kypt_detector.decode_from_dyna(gen_kp, det["first_feature"], vox[:, 0])))
    print("dyna.encode T=20      %.1f ms" % timeit(lambda: net.dyna_module.encode(net.kypt_detector(vox)["keypoints"] if False else kp.repeat(1, 4, 1, 1), aff)))
    # key-frame interpolation (vis_interpolation.py shape: T = 21, key frames every 10) with many hypotheses
    kp21 = kp[:1].repeat(1, 5, 1, 1)[:, :21].contiguous()
    for S in (32, 512, 4096):
        print("dyna.interpolate T=21 S=%-5d %.2f ms" % (S, timeit(lambda: net.dyna_module.interpolate(kp21, aff, sample_num=S, sample_rate=10))))
