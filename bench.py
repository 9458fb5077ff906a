#!/usr/bin/env python
"""Headline benchmark: voxel frames/sec of keypoint detection (BASELINE.json `metric`).

Workload (`configs[1]`): B = 64 synthetic AIST-shape clips x T = 20 frames x 20 000 points at grid 64^3 per GPU;
one step = fused normalise+voxelize of the raw point clouds + `KyptDetector.forward` (spatio-temporal branch,
per-frame encoder, heat-map head / soft-argmax, decoder, reconstruction BCE and auxiliary losses — what the
reference's forward does).  `value` = frames / second with the raw points resident in HBM; `e2e` = the same
through the public API from pinned host memory (H2D of the points, D2H of the keypoints and the loss inside the
timed region).  Multi-GPU (torchrun): clips shard across ranks, no data-path collective -> weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic conv FLOPs (2*MAC) at G = 64, SURVEY.md §8(d) / BASELINE.md §5
GF_ENC_FRAME, GF_ST_CLIP, GF_DEC_FRAME = 27.848, 93.815, 65.434


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--cpu-baseline-frames", type=int, default=20,
                    help="frames of the ONE clip the CPU arm processes per step (default = a full clip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="detector", choices=["detector", "generate"],
                    help="detector = configs[1] (default, the headline); generate = configs[2] shape: "
                         "NeuralMarionette.generate (detector on Tcond frames + HSVRNN roll-out + decode of the rest)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synthetic_raw(seed0, B, T, N):
    from oracle import nm_oracle as O   # input generator shared with the tests (not on the measured path)
    base = [O.synthetic_clip(seed0 + b, T, N) for b in range(min(B, 8))]
    rng = np.random.default_rng(seed0)
    out = np.empty((B, T, N, 3), dtype=np.float32)
    for b in range(B):       # 8 distinct clips, re-posed by a random rigid offset/scale: distinct occupancy per clip
        out[b] = base[b % len(base)] * np.float32(rng.uniform(0.8, 1.2)) + rng.uniform(-0.2, 0.2, size=3).astype(np.float32)
    return out


def cpu_reference_run(args, steps, warmup, frames):
    """The reference path's CPU implementation (oracle port: /root/reference cannot travel to the GPU box)."""
    from oracle import nm_oracle as O
    torch.set_num_threads(os.cpu_count())
    hp = O.default_hparams(grid_size=args.grid)
    sd = O.synthetic_state_dict(hp, seed=0)
    raw = O.synthetic_clip(1000, frames, args.points)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vox = torch.from_numpy(O.voxelize_clip(O.episodic_normalization(raw), args.grid))[None]
            if args.workload == "generate":
                float(O.marionette_generate(vox, sd, hp)["gen"][:, -1].mean())
            else:
                out = O.detector_forward(vox, sd, hp)
                float(out["recon_loss"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return frames / t, t, torch.get_num_threads()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    B, T, N, G = args.clips, args.frames, args.points, args.grid
    config = {"workload": f"KyptDetector.forward (voxelize + ST branch + encoder + head/soft-argmax + decoder + losses) "
                          f"on {B} synthetic AIST-shape clips x {T} frames x {N} pts per GPU, grid {G}^3, K=24",
              "clips_per_gpu": B, "frames_per_clip": T, "points_per_frame": N, "grid": G,
              "l2": "inputs (307 MB of points, >20 GB of activations per step) exceed the 126 MB L2",
              "precision": "fp16 activations/weights at rest, fp32 accumulation (tcgen05 kind::f16), fp32 GroupNorm "
                           "statistics / heads / losses",
              "parallelism": f"clip-sharded x{world}, no data-path collective"}

    if args.workload == "generate":
        Tc = 5   # opt.pickle / dataset/config.py:55-56 (oracle.default_hparams().Tcond)
        config["workload"] = (f"NeuralMarionette.generate (voxelize + detector on Tcond={Tc} frames + {T}-step HSVRNN "
                              f"roll-out + decode of the {T - Tc} generated frames) on {B} synthetic clips x {T} frames "
                              f"x {N} pts per GPU, grid {G}^3, K=24")
    if args.impl == "reference":
        if rank != 0:
            return
        frames = max(2, args.cpu_baseline_frames)
        fps, t, cores = cpu_reference_run(args, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)), frames)
        print(json.dumps({
            "impl": "reference", "metric": "voxel frames/sec keypoint detection", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"1 clip x {frames} frames per step (voxelize + {args.workload}; same per-frame work "
                                       f"as the GPU arm)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    import neural_marionette_b200 as nm
    from neural_marionette_b200 import _lib, ops
    from oracle import nm_oracle as O   # synthetic weights/inputs + the cpu_baseline leg only

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.call("nm_device_supported")

    hp = O.default_hparams(grid_size=G)
    net = nm.NeuralMarionette(hp)
    net.load_state_dict(O.synthetic_state_dict(hp, seed=0), strict=True)
    net = net.to(dev).eval()
    net.anneal(1)
    det = net.kypt_detector

    raw_host = torch.from_numpy(synthetic_raw(1000 + 100 * rank, B, T, N)).pin_memory()
    raw_dev = raw_host.to(dev)

    def step_resident():
        vox = ops.normalize_voxelize(raw_dev, G, check=False)
        return det(vox)

    def step_e2e():
        vox = ops.normalize_voxelize(raw_host.to(dev, non_blocking=True), G, check=False)
        out = det(vox)
        kp = out["keypoints"].cpu()
        return kp, float(out["recon_loss"])

    if args.workload == "generate":
        act = {"detector": True, "learner": True}
        with torch.no_grad():
            net(ops.normalize_voxelize(raw_dev[:2], G, check=False), act)      # builds the skeleton once, as the reference requires
        def step_resident():  # noqa: F811
            return net.generate(ops.normalize_voxelize(raw_dev, G, check=False), act)

        def step_e2e():  # noqa: F811
            out = net.generate(ops.normalize_voxelize(raw_host.to(dev, non_blocking=True), G, check=False), act)
            return out["keypoints"].cpu(), float(out["gen"][:, -1].mean())

    from neural_marionette_b200.parallel import barrier, max_over_ranks

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev) / steps      # device time, slowest rank

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_resident()
        sampler = ClockSampler(local)
        sampler.start()
        calls0 = _lib.CALLS
        ops.PROFILE = {}
        ms = timed(step_resident, args.steps)
        prof = ops.PROFILE
        ops.PROFILE = None
        launches = (_lib.CALLS - calls0)
        clocks = sampler.result()
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

        # SURVEY.md §8(d) config #2 asks for the encoder-only figure next to the full forward: voxelize + VoxToKyptNet
        # (ST branch, per-frame encoder, heat-map head, soft-argmax, Gaussian render) without decoder and losses
        def step_encoder():
            return det.vox_to_kypt(ops.normalize_voxelize(raw_dev, G, check=False))
        step_encoder()
        ms_enc = timed(step_encoder, args.steps)

    frames_total = world * B * T
    value = frames_total / (ms / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)

    # roofline of the dominant kernel: the tcgen05 implicit-GEMM conv, timed per launch with CUDA events
    pk = peaks()
    torch.cuda.synchronize()
    by_shape = {}
    for key, evs in (prof or {}).items():
        t_ms = sum(a.elapsed_time(b) for a, b in evs)
        by_shape[key] = (t_ms, len(evs))
    roof = None
    if by_shape:
        tot_ms = sum(v[0] for v in by_shape.values())
        tot_flop = sum(k[-1] * v[1] for k, v in by_shape.items())
        top = max(by_shape.items(), key=lambda kv: kv[1][0])
        (tn, tD, tci, tco, tk, ts, tflop), (t_ms, cnt) = top
        ach = tflop * cnt / (t_ms / 1e3) / 1e12
        slab3 = tk == 3 and ts == 1 and tD % 16 == 0 and tci in (32, 64, 128) and tco <= 128 and \
            tco % (32 if tci <= 64 else 16) == 0
        roof = {"bound": "tensor",
                "kernel": ("conv3d_slab3_kernel" if slab3 else "conv3d_tc_kernel") + " (nm_conv3d_tc, tcgen05 implicit GEMM)",
                "layer": f"n={tn} grid={tD} Cin={tci} Cout={tco} k={tk} s={ts}", "achieved": ach,
                "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch (dec.8 with the up-sampling produced in its
                # operand path) from the ncu pass in profiles/r01_step_launches_v3.md: 1.40 GB read + 5.34 GB written
                # for 320 frames = 21.08 MB/frame (algorithmic: 4.19 MB low-resolution input + 16.78 MB output)
                "traffic": 21.08e6 * tn if (tD, tci, tco, tk) == (64, 64, 32, 3) else None,
                "note": "dec.8: conv3d_k3(upsample2x(LeakyReLU(GroupNorm(x)))) in one kernel; FLOPs counted are the "
                        "conv's only" if (tD, tci, tco, tk) == (64, 64, 32, 3) else None,
                "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "launch_ms": t_ms / cnt, "share_of_step": t_ms / (ms * args.steps),
                "all_tc_convs": {"achieved": tot_flop / (tot_ms / 1e3) / 1e12, "share_of_step": tot_ms / (ms * args.steps)},
                "by_layer": [{"layer": f"n={k[0]} grid={k[1]} {k[2]}->{k[3]} k{k[4]}s{k[5]}", "launches": v[1],
                              "ms_per_launch": round(v[0] / v[1], 4),
                              "tflops": round(k[6] * v[1] / (v[0] / 1e3) / 1e12, 1),
                              "share_of_step": round(v[0] / (ms * args.steps), 4)}
                             for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:30]]}

    line = {
        "metric": "voxel frames/sec keypoint detection", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(raw_host.numel() * 4),
                "d2h_bytes_per_step": int(B * T * 24 * 4 * 4 + 4), "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "gpu_launches_note": "C-ABI calls inside the timed region (each enqueues >= 1 of our kernels)",
        "model_tflops": (world * (B * (GF_ST_CLIP + T * (GF_ENC_FRAME + GF_DEC_FRAME))) / (ms / 1e3) / 1e3
                         if args.workload == "detector" else
                         world * B * (GF_ST_CLIP + hp.Tcond * (GF_ENC_FRAME + GF_DEC_FRAME) + (T - hp.Tcond) * GF_DEC_FRAME)
                         / (ms / 1e3) / 1e3) if G == 64 else None,
        "encoder_only": {"value": world * B * T / (ms_enc / 1e3), "unit": "frames/s", "ms_per_step": ms_enc,
                         "workload": "voxelize + VoxToKyptNet.forward (no decoder, no losses), same clips",
                         "model_tflops": world * B * (GF_ST_CLIP + T * GF_ENC_FRAME) / (ms_enc / 1e3) / 1e3 if G == 64 else None},
        "roofline": roof,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, t, cores = cpu_reference_run(args, 2, 1, args.cpu_baseline_frames)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"1 clip x {args.cpu_baseline_frames} frames (voxelize + {args.workload}), "
                                          f"1 warm-up + 2 timed reps of {t:.1f} s; same per-frame work as the GPU arm"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
