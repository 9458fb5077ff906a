"""Multi-GPU (NCCL) check of the data-parallel training step; needs >= 2 GPUs (skipped on a single-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  The CPU-side plumbing is covered with gloo in
tests/test_parallel_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allreduced_gradients_equal_single_gpu_gradients_on_the_concatenated_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tools", "ddp_equiv.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(p.stdout[-2000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "DDP_EQUIV world=2" in p.stdout and "params_identical_across_ranks=True" in p.stdout
