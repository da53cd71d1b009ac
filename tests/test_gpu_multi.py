"""GPU, >= 2 devices (-m gpu; skipped on a single-GPU box): the fused NVLink peer-memory all-reduce of dL/dh against the
NCCL path, bitwise agreement between ranks, and CUDA-graph replay - see tools/check_peer_allreduce.py."""
import socket
import subprocess
import sys

import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_allreduce_matches_nccl():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "check_peer_allreduce.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "PEER_ALLREDUCE" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_timeout_is_loud():
    """A rank whose peer never joins gets NaN + a RuntimeError, never a silent partial sum (tools/check_peer_timeout.py)."""
    import os
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "check_peer_timeout.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, B200CAM_COMM_TIMEOUT_S="1"))
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "PEER_TIMEOUT nan=True raised=True" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_caption_camera_sharded_batch_matches_single_gpu():
    """OpticsZernike.data_parallel(): the batch-global max (Lens.py:312) all-reduced between the sensor kernels - two ranks
    reproduce the one-GPU sensor images and coefficient gradient (tools/check_lens_sharded.py)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "check_lens_sharded.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "LENS_SHARDED" in res.stdout and "ok=True" in res.stdout
