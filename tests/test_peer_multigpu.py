"""Peer-memory exchange path (csrc/peer.cuh) on real GPUs: runs scripts/dp_check.py under torchrun on 2 GPUs -- raw
all-gather, sharded == single-process training for comm=nccl / peer / peer+graph, bit-identical replicas.  Skipped on
boxes with fewer than 2 GPUs (the host logic of the sharded trainer is covered on CPU by the gloo test)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_check_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_ALLGATHER PASS" in r.stdout and "DP_CHECK PASS" in r.stdout
