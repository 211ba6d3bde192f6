"""GPU, >= 2 devices: the NCCL path (sharded tICA / KCenters) equals the single-GPU
estimators.  Spawns tools/check_parallel.py under torchrun; skipped on 1-GPU boxes
(the same protocol is covered on CPU by tests/test_parallel_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "check_parallel.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PARALLEL_OK world_size=2" in out.stdout
