"""Rewrites geophyinv.jl_b200/csrc/engine.cu into a host C++ translation unit (test infrastructure; see tests/emu/cuda_rt_shim.h).

Only two textual changes are made, so what the no-GPU suite runs IS the engine's host code and kernels:
  1. <cuda.h> / <cuda_runtime.h> / <nvtx3/nvToolsExt.h>  ->  "cuda_rt_shim.h"  (host stand-in for the runtime API);
  2. every launch  k<<<grid, block, smem, stream>>>(args);  ->  emu::launch(grid, block, [&] { k(args); });
     (emu::launch_mt for the kernels that use __syncthreads).
build(out_dir) compiles the result to libgpifdtd_emu.so and returns its path.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "geophyinv.jl_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
BARRIER_KERNELS = {"k_post"}          # kernels with __syncthreads


def _matching(src: str, i: int, open_ch: str, close_ch: str) -> int:
    """index of the bracket that closes the one at src[i]"""
    depth = 0
    for j in range(i, len(src)):
        if src[j] == open_ch:
            depth += 1
        elif src[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced brackets")


def _split_top(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src: str):
    out, pos, n = [], 0, 0
    pat = re.compile(r"((?:t3::)?k_\w+(?:<[^<>;(){}]*>)?)<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            break
        kernel = m.group(1)
        cfg_end = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():cfg_end])
        assert 2 <= len(cfg) <= 4, cfg
        a0 = cfg_end + 3
        assert src[a0] == "(", src[a0:a0 + 20]
        a1 = _matching(src, a0, "(", ")")
        args = src[a0 + 1:a1]
        base = kernel.split("<")[0].split("::")[-1]
        fn = "emu::launch_mt" if base in BARRIER_KERNELS else "emu::launch"
        out.append(src[pos:m.start()])
        out.append(f"{fn}({cfg[0]}, {cfg[1]}, [&] {{ {kernel}({args}); }})")
        pos = a1 + 1
        n += 1
    out.append(src[pos:])
    return "".join(out), n


def transform(src: str):
    src = src.replace("#include <cuda.h>\n", "").replace("#include <cuda_runtime.h>", '#include "cuda_rt_shim.h"')
    src = src.replace("#include <nvtx3/nvToolsExt.h>\n", "")          # nvtxRangePushA / nvtxRangePop: no-op stand-ins in the shim
    src = src.replace('#include "../../include/gpifdtd.h"', f'#include "{os.path.join(ROOT, "include", "gpifdtd.h")}"')
    src, n = rewrite_launches(src)
    assert "<<<" not in src
    return src, n


def build(out_dir: str, sanitize: bool = False, opt: str = "-O2") -> str:
    with open(os.path.join(CSRC, "engine.cu")) as f:
        src, n = transform(f.read())
    cpp = os.path.join(out_dir, "emu_engine.cpp")
    with open(cpp, "w") as f:
        f.write(src)
    lib = os.path.join(out_dir, "libgpifdtd_emu.so")
    cmd = ["g++", opt, "-g", "-std=c++17", "-shared", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-w", "-DEMU_PLAIN_ARITH",
           "-I", EMU, "-I", CSRC, "-o", lib, cpp, "-ldl", "-lpthread"]
    if sanitize:
        cmd[1:1] = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulated engine failed to compile:\n" + r.stderr[-6000:])
    return lib


if __name__ == "__main__":
    import sys
    d = sys.argv[1] if len(sys.argv) > 1 else "/tmp/emu_engine"
    os.makedirs(d, exist_ok=True)
    print(build(d))
