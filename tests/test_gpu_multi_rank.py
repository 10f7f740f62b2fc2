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


def test_library_comm_sharded_and_partitioned_plans_match_single_gpu():
    """The same with every collective inside the library (pgf_comm_*, NCCL behind the C ABI), including the
    hash-partitioned join / GROUP BY exchange of SURVEY 8e rows 4-5."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs; the partition kernels, row sets and the plan are covered on one GPU by test_gpu_exchange.py")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(root, "tests", "run_multi_gpu_lib.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "library-comm multi-GPU parity ok" in out.stdout
