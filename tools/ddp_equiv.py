"""Data-parallel equivalence of the training step: the gradients a world-size-W NCCL run leaves in
`param.grad` after the bucketed all-reduce equal the single-GPU gradients on the concatenated batch, and one Adam step
moves every rank's parameters identically.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_equiv.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nm_oracle as O            # synthetic weights / clips   # noqa: E402
from oracle import nm_oracle_grad as OG      # train.py's loss weights     # noqa: E402
import neural_marionette_b200 as nm          # noqa: E402
from neural_marionette_b200 import optim     # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G, T, per_rank = 32, 3, 2
hp = O.default_hparams(grid_size=G, Tcond=3, Ttot=10)
sd = O.synthetic_state_dict(hp, seed=61)
clips = np.stack([O.voxelize_clip(O.episodic_normalization(O.synthetic_clip(400 + b, T, 5000)), G)
                  for b in range(world * per_rank)], 0)
vox_all = torch.from_numpy(clips).float().cuda()


def build():
    net = nm.NeuralMarionette(hp)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    net.anneal(1)
    return net


# ---- data-parallel: this rank's shard, bucketed all-reduce driven by the backward, one Adam step
net = build()
opt = optim.FusedAdam(net.kypt_detector.parameters(), lr=4e-4, owner=net)
opt.zero_grad()
out = net.kypt_detector(vox_all[rank * per_rank:(rank + 1) * per_rank])
OG.detector_loss(out, recon_only=False).backward()
opt.buckets.finish()
g_dp = opt.buckets.flat.clone()
assert opt.step()
p_dp = opt.flat_param.clone()

# ---- single process on the concatenated batch (no collective: a second model outside the process group's buckets)
ref = build()
params = [p for p in ref.kypt_detector.parameters() if p.requires_grad][::-1]
out = ref.kypt_detector(vox_all)
OG.detector_loss(out, recon_only=False).backward()
g_ref = torch.cat([p.grad.reshape(-1) for p in params])

err = float((g_dp - g_ref).norm() / g_ref.norm())
mx = float((g_dp - g_ref).abs().max() / g_ref.abs().max())
gathered = [torch.empty_like(p_dp) for _ in range(world)]
dist.all_gather(gathered, p_dp)
same = all(torch.equal(gathered[0], q) for q in gathered)
moved = float((p_dp - torch.cat([p.detach().reshape(-1) for p in params])).abs().max())
if rank == 0:
    print(f"DDP_EQUIV world={world} rel_l2={err:.3e} max_rel={mx:.3e} params_identical_across_ranks={same} max_param_step={moved:.3e}")
ok = err < 2e-3 and same and moved > 0
dist.destroy_process_group()
sys.exit(0 if ok else 1)
