"""The NVLink peer-memory paths on a box with at least two GPUs (skipped on a single-GPU box; the tools print one
line that states the comparison they made):
  * likelihood step with the exchange fused into the kernel == kernel + NCCL all-gather, on every rank;
  * sharded device sampler == single-GPU chain with the same seed, bit for bit, on every rank."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, *args, nproc=2, timeout=300):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")


@pytest.mark.gpu
def test_fused_exchange_equals_allgather_two_gpus():
    _two_gpus()
    out = _torchrun("peer_check.py")
    assert "fused exchange == NCCL all-gather on every rank: True" in out and "peer_error=False" in out, out


@pytest.mark.gpu
def test_sharded_sampler_bit_identical_two_gpus():
    _two_gpus()
    out = _torchrun("mcmc_peer_check.py", "64")
    assert "bit-identical to the single-GPU chain on every rank: True" in out and "peer_error=False" in out, out
