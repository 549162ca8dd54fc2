"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gpifdtd.h declares, agrees with the Python binding's struct layouts, and refuses to run
without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpifdtd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpi_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib(G):
    if not os.path.exists(G.engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return G.load_library()


def test_every_declared_symbol_is_exported(G, lib):
    syms = header_symbols()
    assert len(syms) >= 29
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gpifdtd.h but not exported by libgpifdtd.so"
    assert sorted(G.engine.EXPORTS) == syms, "Python binding and header disagree on the export list"


def test_struct_layouts_match_header(G):
    # gpi_config: 22 int32 (n[3] counted thrice) + 8 doubles; gpi_timers: 12 doubles (ABI 2)
    assert C.sizeof(G.engine.GpiConfig) == 4 * 22 + 8 * 8
    assert C.sizeof(G.engine.GpiTimers) == 8 * 12
    assert G.engine.GpiConfig.dt.offset == 88


def test_field_shapes_follow_fields_jl(G, lib):
    """fields.jl:92-671 via SURVEY App. A: staggered array shapes."""
    n = (C.c_int32 * 3)(20, 30, 40)
    out = (C.c_int32 * 3)()
    want3 = {"p": (20, 30, 40), "vx": (20, 30, 41), "vy": (20, 31, 40), "vz": (21, 30, 40),
             "tauxy": (18, 29, 39), "tauxz": (19, 28, 39), "tauyz": (19, 29, 38),
             "dtauxxdx": (18, 28, 39), "dtauyydy": (18, 29, 38), "dtauzzdz": (19, 28, 38)}
    for f, shp in want3.items():
        phys = G.engine.ACOUSTIC if f == "p" else G.engine.ELASTIC
        assert lib.gpi_field_shape(3, phys, G.engine.FIELD[f], n, out) == 0
        assert tuple(out) == shp, f
    n2 = (C.c_int32 * 3)(20, 1, 40)
    want2 = {"p": (20, 1, 40), "vx": (20, 1, 41), "vz": (21, 1, 40), "dpdx": (18, 1, 39), "dpdz": (19, 1, 38)}
    for f, shp in want2.items():
        assert lib.gpi_field_shape(2, G.engine.ACOUSTIC, G.engine.FIELD[f], n2, out) == 0
        assert tuple(out) == shp, f
    # fields that do not exist for the physics / dimensionality are an error, not a guess
    assert lib.gpi_field_shape(2, G.engine.ACOUSTIC, G.engine.FIELD["vy"], n2, out) != 0
    assert lib.gpi_field_shape(3, G.engine.ACOUSTIC, G.engine.FIELD["tauxy"], n, out) != 0
    assert lib.gpi_field_shape(3, G.engine.ELASTIC, G.engine.FIELD["p"], n, out) != 0
    # host-side shape table agrees with the library
    for f, shp in want3.items():
        assert G.field_shape(f, (20, 30, 40)) == shp


def test_no_gpu_means_error_not_fallback(G, lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = G.engine.GpiConfig()
    cfg.abi_version, cfg.ndims, cfg.physics, cfg.order = 2, 2, 0, 2
    cfg.n[0], cfg.n[1], cfg.n[2] = 100, 1, 100
    cfg.nt, cfg.npml, cfg.nbound, cfg.npw, cfg.nshots = 10, 41, 3, 1, 1
    h = C.c_void_p()
    assert lib.gpi_create(C.byref(cfg), C.byref(h)) != 0
    assert not h.value
    msg = lib.gpi_last_error(None).decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    with pytest.raises(G.EngineError):
        G.Engine(cfg)


def test_bad_config_rejected(G, lib):
    h = C.c_void_p()
    cfg = G.engine.GpiConfig()
    cfg.abi_version = 99
    assert lib.gpi_create(C.byref(cfg), C.byref(h)) != 0
    assert "ABI" in lib.gpi_last_error(None).decode()
    cfg.abi_version, cfg.order, cfg.ndims = 2, 8, 2
    assert lib.gpi_create(C.byref(cfg), C.byref(h)) != 0
    assert "order" in lib.gpi_last_error(None).decode()
    assert lib.gpi_create(None, C.byref(h)) != 0


def test_product_never_imports_oracle():
    """The product package must have no path to the CPU oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "geophyinv.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "from oracle" not in txt, f


def test_julia_shim_binds_every_export():
    """julia/GPIFdtdB200.jl is the reference-side binding; Julia is not installed here, so the check is textual:
    every function include/gpifdtd.h declares has a `ccall((:name, LIB), ...)` in the shim."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gpifdtd.h")).read()
    shim = open(os.path.join(root, "julia", "GPIFdtdB200.jl")).read()
    exports = sorted(set(re.findall(r"\b(gpi_[a-z0-9_]+)\s*\(", hdr)))
    assert len(exports) >= 30
    missing = [e for e in exports if f"(:{e}, LIB)" not in shim]
    assert not missing, missing


def test_no_fma_contraction_in_the_ptx(tmp_path):
    """Bit-identity with the reference's Base.Threads path rests on evaluating every product and sum separately (Julia does not
    contract a*b+c outside @fastmath).  The kernels use __fmul_rn / __fadd_rn / __dmul_rn / ... for that; this compiles the
    library to PTX and checks that nothing was contracted behind their back: no fma.rn.f32 at all, and the only fma.rn.f64 are the
    four `x + 2.0 * y` of update_dmod! (medium.jl:165), whose product is exact."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ptx = str(tmp_path / "engine.ptx")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-ptx", "-o", ptx,
                        os.path.join(root, "geophyinv.jl_b200", "csrc", "engine.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    text = open(ptx).read()
    assert text.count("fma.rn.f32") == 0
    f64 = [l for l in text.splitlines() if "fma.rn.f64" in l]
    assert len(f64) <= 4 and all("0d4000000000000000" in l for l in f64), f64[:6]


def test_machine_code_of_the_measured_kernels_is_unchanged():
    """The numbers under profiles/r02 were measured with exactly this machine code: the SASS of every kernel at the round's last GPU run
    (tests/golden/sass_r02.txt: cuobjdump -sass hashed per kernel by scripts/sass_hashes.py, kernel-parameter offsets masked; written by
    scripts/pin_sass.py) must come out of the current sources unchanged.  A deliberate kernel change re-pins the fixture in the same
    commit -- and its numbers need re-measuring."""
    import shutil
    import subprocess
    import sys
    lib = os.path.join(ROOT, "geophyinv.jl_b200", "libgpifdtd.so")
    pin = os.path.join(ROOT, "tests", "golden", "sass_r02.txt")
    if not (os.path.exists(lib) and os.path.exists(pin) and shutil.which("cuobjdump") and shutil.which("c++filt")):
        pytest.skip("needs the built library, the pinned hashes and cuobjdump")
    src_newer = max(os.path.getmtime(os.path.join(ROOT, "geophyinv.jl_b200", "csrc", f))
                    for f in os.listdir(os.path.join(ROOT, "geophyinv.jl_b200", "csrc")) if f.endswith((".cu", ".cuh")))
    if src_newer > os.path.getmtime(lib):
        pytest.skip("libgpifdtd.so is older than its sources (run __graft_entry__.build())")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_hashes.py"), lib], capture_output=True, text=True, check=True).stdout
    now = {}
    for line in out.splitlines():
        h, _, name = line.split(" ", 2)
        now[name.strip()] = h
    want = {}
    for line in open(pin):
        h, name = line.rstrip("\n").split("  ", 1)
        want[name.strip()] = h
    assert len(want) >= 50
    changed = [k for k in want if now.get(k) != want[k]]
    assert not changed, "SASS changed for: " + "; ".join(changed)


def _julia_struct(shim, name):
    """(field, type) list of `struct name ... end` in the Julia shim."""
    import re
    body = re.search(r"struct\s+" + name + r"\s*\n(.*?)\nend", shim, re.S).group(1)
    out = []
    for stmt in re.split(r"[;\n]", body):
        stmt = stmt.split("#")[0].strip()
        if stmt:
            f, t = stmt.split("::")
            out.append((f.strip(), t.strip()))
    return out


def _c_struct(hdr, name):
    """(field, ctype, count) list of `typedef struct name { ... } name;` in include/gpifdtd.h."""
    import re
    body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + ";", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ctype, rest = stmt.split(None, 1)
        for decl in rest.split(","):
            m = re.match(r"\s*(\w+)\s*(?:\[(\d+)\])?\s*$", decl)
            out.append((m.group(1), ctype, int(m.group(2) or 1)))
    return out


def test_struct_layouts_agree_across_header_ctypes_and_julia(G):
    """Struct drift is the silent failure of a ccall binding: `gpi_config` and `gpi_timers` must have the same fields, in the same order,
    at the same byte offsets in include/gpifdtd.h, in the ctypes mirror (engine.py) and in julia/GPIFdtdB200.jl (isbits structs follow
    the C layout rules), and the three must agree on the ABI version."""
    import re
    hdr = open(os.path.join(ROOT, "include", "gpifdtd.h")).read()
    shim = open(os.path.join(ROOT, "julia", "GPIFdtdB200.jl")).read()
    csize = {"int32_t": 4, "double": 8}
    jsize = {"Int32": (4, 4), "Float64": (8, 8), "NTuple{3,Int32}": (12, 4), "NTuple{3,Float64}": (24, 8)}
    for cname, jname, ct in (("gpi_config", "GpiConfig", G.engine.GpiConfig), ("gpi_timers", "GpiTimers", G.engine.GpiTimers)):
        # header -> offsets (natural alignment)
        off, want = 0, []
        for f, t, n in _c_struct(hdr, cname):
            a = csize[t]
            off = (off + a - 1) // a * a
            want.append((f, off, a * n))
            off += a * n
        got = [(f, getattr(ct, f).offset, getattr(ct, f).size) for f, _ in ct._fields_]
        assert got == want, (cname, got, want)
        off, jl = 0, []
        for f, t in _julia_struct(shim, jname):
            size, a = jsize[t]
            off = (off + a - 1) // a * a
            jl.append((f, off, size))
            off += size
        assert jl == want, (jname, jl, want)
    abi_h = int(re.search(r"#define GPI_ABI_VERSION (\d+)", hdr).group(1))
    abi_j = int(re.search(r"const ABI_VERSION = Int32\((\d+)\)", shim).group(1))
    assert abi_h == abi_j == G.engine.ABI_VERSION


def test_julia_shim_is_loadable_as_a_submodule():
    """Textual checks of what made the round-1 shim unloadable: a submodule does not see its parent's bindings, so every GeoPhyInv
    name the shim uses must be imported with `using ..GeoPhyInv: ...`; Born / unshifted runs must be reachable through mod_x_proc!."""
    import re
    shim = open(os.path.join(ROOT, "julia", "GPIFdtdB200.jl")).read()
    code = "\n".join(l.split("#")[0] for l in shim.splitlines())
    imported = re.search(r"using \.\.GeoPhyInv:\s*([^\n]+)", code).group(1).replace(" ", "").split(",")
    for name in ("FdtdElastic", "_fd_order", "_fd_npml", "_fd_nbound", "_fd_npextend"):
        assert (re.search(r"\b" + name + r"\b", code.replace("using ..GeoPhyInv:", "")) is None) or name in imported, name
    used = set(re.findall(r"\b(_fd_\w+|Fdtd\w+)\b", code))
    assert used <= set(imported), used - set(imported)
    assert "mode_flags" in code and "GPI_RUN_BORN" in code and "GPI_RUN_UNSHIFTED_RHO" in code
    hdr = open(os.path.join(ROOT, "include", "gpifdtd.h")).read()
    for name in ("GPI_RUN_BORN", "GPI_RUN_UNSHIFTED_RHO"):
        hv = int(re.search(r"#define " + name + r" (0x[0-9a-fA-F]+)", hdr).group(1), 16)
        jv = int(re.search(r"const " + name + r" = Int32\((0x[0-9a-fA-F]+)\)", shim).group(1), 16)
        assert hv == jv


def test_every_environment_switch_is_documented():
    """Every GPI_* variable the library or the binding reads appears in INTEGRATION.md's switch table."""
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = set()
    for rel in ("geophyinv.jl_b200/csrc/engine.cu", "geophyinv.jl_b200/engine.py", "geophyinv.jl_b200/host/fdtd.py"):
        txt = open(os.path.join(ROOT, rel)).read()
        names |= set(re.findall(r'getenv\("(GPI_[A-Z0-9_]+)"\)', txt)) | set(re.findall(r'environ(?:\.get)?[\[(]"(GPI_[A-Z0-9_]+)"', txt))
    assert len(names) >= 15
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
