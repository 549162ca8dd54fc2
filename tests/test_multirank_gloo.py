"""world_size-2 `gloo` test of the multi-rank host path on CPU: contiguous shot chunks per
rank, record gather to rank 0, gradient sum over ranks, ncclUniqueId broadcast."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_sschunks_match_reference_rounding(G):
    """fdtd.jl:251-255: ssi = round.(Int, range(0, nss, length=nworker+1))"""
    assert [list(c) for c in G.sschunks(64, 8)] == [list(range(8 * i, 8 * i + 8)) for i in range(8)]
    assert [len(c) for c in G.sschunks(5, 2)] == [2, 3]          # round(2.5) = 2 (ties to even, as in Julia)
    assert [len(c) for c in G.sschunks(7, 2)] == [4, 3]          # round(3.5) = 4
    assert sum(len(c) for c in G.sschunks(33, 8)) == 33


def test_two_ranks_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MULTIRANK_OK" in r.stdout
