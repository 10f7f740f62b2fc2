"""Launches the one-process-per-GPU parity run (tests/run_multi_gpu.py) when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sharded_q6_q1_q3_match_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (NCCL path); the merge kernels themselves are covered by test_gpu_multi.py")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "run_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "multi-GPU parity ok" in out.stdout
