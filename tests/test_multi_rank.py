"""N > 1: the sort-last plumbing and exchange kernels with one process per rank.

CPU (gloo, world_size 2): host-side logic of ascent_b200.distributed vs the oracle.
GPU (nccl, needs >= 2 B200): path A / path B through the P2P compositing kernels, bit-exact vs
the oracle at the same rank/block layout.  Runs under ``gpurun --gpus 2 -- pytest -m gpu``;
skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(mode, world, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "multi_rank_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="2", VR_COMM_TIMEOUT_MS=os.environ.get("VR_COMM_TIMEOUT_MS", "8000"))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "checks OK" in r.stdout


def test_host_logic_world2_gloo():
    _run("host", 2, 300)


def test_host_logic_world4_gloo():
    _run("host", 4, 300)


@pytest.mark.gpu
def test_sort_last_p2p_ranks_sharing_one_gpu():
    """2 and 3 rank processes time-sharing cuda:0, arenas mapped through CUDA IPC: every cross-rank kernel
    (pull and push image folds, partial merge, layer fold, z-buffer, depth broadcast, abort protocol) runs
    against the oracle even on a one-GPU box."""
    _run("gpu-shared", 2, 900)
    _run("gpu-shared", 3, 900)


@pytest.mark.gpu
def test_sort_last_p2p_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _run("gpu", 2, 600)
    if n >= 4:
        _run("gpu", 4, 600)
    if n >= 8:
        _run("gpu", 8, 600)
