"""Data-parallel parity on real GPUs over NCCL: after the overlapped bucketed all-reduce every rank's
gradients equal the MEAN over ranks of the single-process gradients of each rank's shard -- eager micro-steps and CUDA-graphed
micro-steps (exchange queued behind in-graph events, and the default: exchange after the last replay).  Runs scripts/dp_parity.py under torchrun on 2 GPUs; skipped on a box
with fewer than 2 (the world-size-2 host logic is covered on CPU by tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_gradients_equal_mean_of_shards_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs 2 GPUs, this box has {n}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "scripts", "dp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("-> OK") == 3, r.stdout[-2000:]
