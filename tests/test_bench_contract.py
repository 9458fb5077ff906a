"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads,
non-zero ranks of a multi-process launch exit without work, and the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    # small grid / 2 frames so that the CPU port finishes in seconds; the default run uses grid 64 and a full clip
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--grid", "32", "--points", "2000",
              "--cpu-baseline-frames", "2"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    # kind = "reference" when oracle/_ref holds the staged reference (build container / GPU box), else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # the line says what actually ran: one timed step of ONE clip x 2 frames, not the GPU arm's 64-clip step
    assert d["steps"] == 1 and d["warmup"] == 1 and d["config"]["clips_per_step"] == 1 and d["config"]["frames_per_clip"] == 2
    assert abs(d["value"] - 2 / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype"):
        assert k in d


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1"], timeout=300)
    assert p.returncode != 0 and "needs a GPU" in (p.stderr + p.stdout)
