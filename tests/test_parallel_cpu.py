"""world_size-2 gloo test of the clip-sharding host logic (the N>1 path of bench.py / multi-GPU inference)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_marionette_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_keypoints(clip_ids):
    # deterministic "output" of clip i, so that any mis-ordering of the gather is visible
    return torch.stack([torch.full((3, 24, 4), float(i)) + torch.arange(4.0) for i in clip_ids]) if len(clip_ids) \
        else torch.zeros(0, 3, 24, 4)


def _worker(rank, world, port, n_clips, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = P.shard_range(n_clips, rank, world)
    local = _fake_keypoints(range(lo, hi))
    full = P.gather_clip_outputs(local, n_clips)
    slowest = P.max_over_ranks(10.0 + rank)
    P.barrier()
    ok = torch.equal(full, _fake_keypoints(range(n_clips))) and slowest == 10.0 + world - 1
    out.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly_once():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            r = [P.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
            assert P.shard_sizes(n, w) == [h - l for l, h in r]


def test_gloo_world2_gather_and_timing():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 4), (4, 7)]


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 11))
    net[1].bias.requires_grad_(False)                       # frozen tensors (offset_param in the reference) are skipped
    gb = P.GradientBuckets(net.parameters(), n_buckets=3)
    trainable = [p for p in net.parameters() if p.requires_grad]
    ok = sum(hi - lo for lo, hi in gb.bounds) == sum(p.numel() for p in trainable) and 2 <= len(gb.bounds) <= 3
    ok = ok and gb.params[0] is trainable[-1]               # reverse order: the last layer's gradients come first
    for step in range(2):
        gb.zero()
        for i, p in enumerate(gb.params):                   # "backward": rank r produces (r + 1) * (i + 1 + step) everywhere
            p.grad.fill_(float((rank + 1) * (i + 1 + step)))
            gb.ready(p)
        gb.finish()
        mean = sum(r + 1 for r in range(world)) / world
        for i, p in enumerate(gb.params):
            ok = ok and bool(torch.all(p.grad == mean * (i + 1 + step))) and p.grad.data_ptr() >= gb.flat.data_ptr()
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_gradient_buckets():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _autograd_worker(rank, world, port, out):
    """A real backward drives the buckets through post-accumulate hooks; between the steps the caller drops the `.grad`
    aliases the way torch's default `optimizer.zero_grad(set_to_none=True)` does."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    ref.load_state_dict(net.state_dict())
    gb = P.GradientBuckets(net.parameters(), n_buckets=2)
    gb.attach_hooks()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    x_all = torch.randn(3, world * 4, 6, generator=g)        # the same data on every rank; rank r takes its 4 rows
    ok = True
    for step in range(3):
        if step == 1:
            opt.zero_grad()                                   # set_to_none=True: every alias is dropped
            ok = ok and all(p.grad is None for p in net.parameters())
        gb.zero()                                             # re-attaches (and zeroes) whatever the caller did
        if step == 2:
            for p in net.parameters():
                p.grad = None                                 # dropped AFTER zero(): the hooks must re-home the fresh gradients
        net(x_all[step, rank * 4:(rank + 1) * 4]).pow(2).sum().backward()
        gb.finish()
        ref.zero_grad()
        (ref(x_all[step]).pow(2).sum() / world).backward()    # mean over the ranks of the per-rank sums
        for p, q in zip(net.parameters(), ref.parameters()):
            ok = ok and bool(torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6))
            off, n = gb._slot[id(p)]
            ok = ok and p.grad.data_ptr() == gb.flat[off:off + n].data_ptr()
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_buckets_survive_dropped_grad_aliases():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_autograd_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_gradient_buckets_single_process():
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    gb = P.GradientBuckets(net.parameters(), n_buckets=8)   # more buckets than tensors: one tensor per bucket at most
    assert 1 <= len(gb.bounds) <= 4 and gb.bounds[0][0] == 0 and gb.bounds[-1][1] == gb.flat.numel()
    assert all(a[1] == b[0] for a, b in zip(gb.bounds, gb.bounds[1:]))
    for p in gb.params:
        p.grad.fill_(2.0)
        gb.ready(p)
    gb.finish()                                             # world size 1: no collective, no scaling
    assert bool(torch.all(gb.flat == 2.0))


def test_gradient_bucket_bounds_property():
    """Whatever the tensor sizes and bucket count: buckets tile the flat buffer exactly, follow the reverse parameter order,
    never split a tensor, and every `.grad` aliases its slice."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 40), min_size=1, max_size=12), st.integers(1, 9))
    def check(sizes, n_buckets):
        params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
        gb = P.GradientBuckets(params, n_buckets=n_buckets)
        assert gb.flat.numel() == sum(sizes) and 1 <= len(gb.bounds) <= min(n_buckets, len(sizes))
        assert gb.bounds[0][0] == 0 and gb.bounds[-1][1] == sum(sizes)
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(gb.bounds, gb.bounds[1:] + [(sum(sizes), None)]))
        off = 0
        for p in params[::-1]:
            assert p.grad.data_ptr() == gb.flat.data_ptr() + 4 * off
            lo, hi = gb.bounds[gb._bucket_of[id(p)]]
            assert lo <= off and off + p.numel() <= hi
            off += p.numel()
    check()
