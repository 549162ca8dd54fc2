"""CPU emulation of the order-4 CUDA kernels (tests/emu/): kernels4.cuh and kernels4v.cuh compiled as host C++ behind a thin shim
(no shared memory, barriers or PTX in these kernels, so running the threads one after the other is exact).  The four-cells-per-thread
kernels must reproduce the one-cell-per-thread kernels -- which the GPU tests pin to the oracle bit for bit -- on random fields with
CPML slabs on every face, a free surface, several resident shots and the ragged z extents of the float4 groups."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vectorised_order4_kernels_match_the_scalar_ones(tmp_path):
    exe = str(tmp_path / "emu_k4")
    src = os.path.join(ROOT, "tests", "emu", "emu_kernels4.cpp")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and "EMU_OK" in r.stdout, r.stdout[-3000:]
