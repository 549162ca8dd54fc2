"""CPU emulation of the CUDA kernels (tests/emu/): the .cuh sources compiled as host C++ behind a thin shim (tests/emu/cuda_shim.h).
Every kernel family except the TMA-pipelined one is free of shared memory, barriers and PTX outside four accessor helpers, so running
its threads one after the other is exact.  The one-thread-per-cell reference-order kernels are pinned to the oracle bit for bit by the
GPU tests; here the vectorised families must reproduce them on random fields -- CPML slabs on every face, rigid faces, a free surface,
several resident shots, ragged z extents of the float4 groups -- without a GPU, and under AddressSanitizer / UBSan (no access outside
an operand array):

  * order 2: k_vel2v / k_stress2v and k_vel3v / k_stress3v  vs  k_vel / k_stress        (emu_kernels2.cpp, three time steps)
  * order 4: k_vel4v / k_stress4v                             vs  k_vel4 / k_stress4    (emu_kernels4.cpp)
  * bounds + round trip: k_grad2d_el, k_grad3d_el, k_grad3d (orders 2 and 4), k_boundary save -> wipe -> force with three / six
    fields and a shot batch                                                              (emu_kernels_misc.cpp)
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


NAMES = ["emu_kernels2", "emu_kernels4", "emu_kernels_misc"]


def _build_and_run(name, out_dir):
    exe = os.path.join(out_dir, name)
    src = os.path.join(ROOT, "tests", "emu", name + ".cpp")
    # AddressSanitizer + UBSan: the operands live in exactly-sized host vectors, so any read or write outside a field, a coefficient
    # array or a CPML memory block (a halo load at the edge of the box, a float4 straddling a row) aborts the run
    flags = ["-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w"]
    san = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    r = subprocess.run(["g++"] + flags + san + ["-o", exe, src], capture_output=True, text=True)
    if r.returncode != 0:                       # toolchain without the sanitizer runtimes: plain build
        r = subprocess.run(["g++"] + flags + ["-o", exe, src], capture_output=True, text=True)
    if r.returncode != 0:
        return r.returncode, "compile failed:\n" + r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    return r.returncode, r.stdout


@pytest.fixture(scope="module")
def harness_runs(tmp_path_factory):
    """The three harnesses are single-threaded compile-and-run jobs: started together, collected one by one."""
    from concurrent.futures import ThreadPoolExecutor
    out_dir = str(tmp_path_factory.mktemp("emu_kernels"))
    pool = ThreadPoolExecutor(max_workers=len(NAMES))
    futures = {name: pool.submit(_build_and_run, name, out_dir) for name in NAMES}
    yield futures
    pool.shutdown(wait=True)


@pytest.mark.parametrize("name", NAMES)
def test_vectorised_kernels_match_the_reference_order_kernels(harness_runs, name):
    rc, out = harness_runs[name].result(timeout=1200)
    sys.stdout.write(out)
    assert rc == 0 and "EMU_OK" in out, out[-3000:]
