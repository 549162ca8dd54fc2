"""The whole C-ABI library on the CPU: engine.cu -- handle life cycle, uploads and layout conversion, the time loop of gpi_run, every
kernel launch -- compiled as host C++ behind a stand-in for the CUDA runtime (tests/emu/cuda_rt_shim.h, tests/emu/make_emu_engine.py:
two textual substitutions, launches become loops over blocks and threads) and driven through the SAME ctypes binding and the SAME
parity tests the B200 runs (`-m gpu` tests of tests/test_parity_gpu.py, tests/test_order4.py, ...), in a child pytest with
GPI_LIB pointing at the emulated library.  Every one of them must hold bit for bit against the oracle without a GPU.

What this covers that tests/test_emu_kernels.py does not: the host side of the engine (batching, descriptor tables, source / receiver
row lists, boundary store slots, two-wavefield merged launches, the ping-pong time levels of GPI_PINGPONG=1, gradient stacking) and
the kernels in their real launch geometry -- including the TMA-pipelined 3-D elastic kernels (kernels3t.cuh) through the host forms
of their PTX primitives: the tile decomposition and the shell, the descriptors (an emulated cuTensorMapEncodeTiled with the driver's
argument checks), the producer's box list and expect_tx byte accounting (a mismatch aborts instead of hanging), the tile header,
the consumers' shared-memory indexing.  What it cannot cover: the asynchrony of that pipeline and its warp shuffles, NCCL, and
anything about timing.  All 69 single-GPU parity tests pass under the emulation (24 min on 8 cores, profiles/r01/emu_parity_final.log); the no-GPU suite runs the
subset below.

TEST INFRASTRUCTURE: the emulated library is built into a temporary directory, is never installed next to the package, and
`engine.py` cannot pick it up by itself (it loads libgpifdtd.so or fails); `test_product_library_is_not_the_emulation` checks that.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

PARITY = "tests/test_parity_gpu.py::"
SELECTION = [
    PARITY + "test_c1_acoustic2d_records[p-rfields0]",
    PARITY + "test_acoustic2d_multishot_batches",
    PARITY + "test_elastic2d_records[True]",
    PARITY + "test_elastic2d_stress_source",
    PARITY + "test_elastic3d_partial_pml_faces[faces1]",        # TMA tiles forced (GPI_TMA3=2), CPML boxes on three faces
    PARITY + "test_elastic3d_partial_pml_faces[faces2]",
    PARITY + "test_dmod_matches_oracle",
    PARITY + "test_medium_padded_on_device",
    PARITY + "test_simultaneous_and_coincident_sources[acou2d]",
    PARITY + "test_boundary_save_and_force_match_oracle[acou2d_batched]",
    PARITY + "test_boundary_save_and_force_match_oracle[elastic2d]",
    PARITY + "test_small_and_degenerate_cases",
    PARITY + "test_axis_shorter_than_npml_is_rejected_with_a_message",
    PARITY + "test_fwi_gradient_acoustic2d",                    # default path: ping-pong time levels + the fused pass of kernels2a.cuh
    PARITY + "test_born_records_match_oracle[p-rfields0]",
    PARITY + "test_illumination_matches_oracle[acou2d]",
    PARITY + "test_illumination_matches_oracle[fwi2d]",
    PARITY + "test_pingpong_adjoint_equals_the_copy_path[elastic]",
    "tests/test_order4.py::test_order4_elastic2d[False-vz]",
    # the engine against fixtures evaluated from the REFERENCE'S OWN KERNEL TEXT (no oracle in the loop): every physics / order / dimension
    "tests/test_reference_pinned.py::test_cuda_engine_reproduces_the_reference_text",
]


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    import make_emu_engine
    return make_emu_engine.build(str(tmp_path_factory.mktemp("emu_engine")))


def test_launch_rewriting_covers_every_launch():
    """Every `<<<...>>>` of engine.cu becomes an emu::launch; nothing else of the file changes but the include lines."""
    import make_emu_engine
    with open(os.path.join(ROOT, "geophyinv.jl_b200", "csrc", "engine.cu")) as f:
        src = f.read()
    out, n = make_emu_engine.transform(src)
    assert n == src.count("<<<") and n >= 40
    assert out.count("emu::launch_mt(") == src.count("k_post<<<")
    assert "cuda_runtime.h" not in out and "<<<" not in out


def test_product_library_is_not_the_emulation(emu_lib):
    nm = lambda p: subprocess.run(["nm", "-D", "--defined-only", p], capture_output=True, text=True).stdout
    assert "gpi_emu_marker" in nm(emu_lib)
    prod = os.path.join(ROOT, "geophyinv.jl_b200", "libgpifdtd.so")
    if os.path.exists(prod):
        assert "gpi_emu_marker" not in nm(prod)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "geophyinv.jl_b200")):
        assert not any("emu" in f for f in files), "the package must not ship an emulated library"
        for f in files:
            if f.endswith(".py"):
                # the only mention the package may make of the emulation is engine.py's guard that REFUSES to load it
                txt = "\n".join(l for l in open(os.path.join(dirpath, f)).read().lower().splitlines()
                                if not any(k in l for k in ("gpi_emu_marker", "gpi_tests_allow_emu", "never the cpu emulation", "is the cpu emulation")))
                assert "emu" not in txt.replace("enumerate", ""), f


def test_engine_host_code_and_kernels_match_the_oracle_on_the_cpu(emu_lib):
    env = dict(os.environ, GPI_LIB=emu_lib, GPI_TESTS_ALLOW_EMU="1", OMP_WAIT_POLICY="passive")
    env.pop("GPI_PINGPONG", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-s", "-p", "no:cacheprovider"] + SELECTION,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and " deselected" not in r.stdout          # every selected test ran (exit code 0: none failed)
