"""Multi-GPU tests (need >= 2 GPUs; skipped otherwise): one process per GPU launched with torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def run_ranks(script, n, timeout=900, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    return r.stdout


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
def test_shot_sharding_and_nccl_gradient_allreduce():
    out = run_ranks("_nccl_worker.py", 2)
    assert "NCCL_SHOTS_OK" in out
    print(out.strip().splitlines()[-1])


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("tma", ["1", "2"])
def test_zslab_decomposition_matches_single_gpu(tma):
    """GPI_TMA3=1: kernel family chosen by tile utilisation (narrow slabs -> register-staged kernels);
    GPI_TMA3=2: the TMA-pipelined kernels on the slab windows too."""
    out = run_ranks("_slab_worker.py", 2, extra_env={"GPI_TMA3": tma})
    assert "SLAB_OK" in out
    print(out.strip().splitlines()[-1])


@pytest.mark.skipif(ngpus() < 4, reason="needs 4 GPUs")
def test_zslab_decomposition_four_ranks():
    """Four slabs: the two middle ranks exchange with both neighbours in both phases."""
    out = run_ranks("_slab_worker.py", 4)
    assert "SLAB_OK world=4" in out
    print(out.strip().splitlines()[-1])


@pytest.mark.skipif(ngpus() < 4, reason="needs 4 GPUs")
def test_shot_sharding_four_ranks():
    out = run_ranks("_nccl_worker.py", 4)
    assert "NCCL_SHOTS_OK" in out
    print(out.strip().splitlines()[-1])
