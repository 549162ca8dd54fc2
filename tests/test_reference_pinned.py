"""Parity pinned to the REFERENCE'S SOURCE TEXT (VERDICT r01 task 2).

tests/golden/ref_<case>_<lit>.npz were produced by tests/golden/from_reference.py, which parses the `@parallel` kernel bodies, the
finite-difference macros, the templated CPML / Dirichlet / boundary kernels, the `update_*!` call sequences and the `get_mgrid` shape
methods out of /root/reference/src and evaluates them with numpy for a few dozen time steps (nothing of oracle/ or of the CUDA engine
is involved in making them).  Here the same inputs (tests/golden/refcases.py) go through the C ABI into

  * the CPU oracle (oracle/fdtd_oracle.c), in BOTH literal typings           -> must reproduce the fixtures BIT FOR BIT  (no GPU needed)
  * the CUDA engine (libgpifdtd.so), literal typing = the reference's (f32)  -> BIT FOR BIT as well                      (-m gpu)

so the chain  reference text -> oracle -> CUDA engine  has no unpinned link.  Compared: every record sample, a strided sample and two
bit-pattern checksums (over every entry; -0.0 counted as +0.0, see from_reference.checksum) of every final wavefield, the derived medium coefficients (`dmod`), the array shapes, and for the FWI case the gradients of
invK and rho after a forward_save + adjoint pass (boundary store, save_tp!, adjoint sources, imaging).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200 import engine as E  # noqa: E402
import refcases  # noqa: E402

F32 = np.float32
GOLD = os.path.join(ROOT, "tests", "golden")
DMOD_ORACLE = {"dtinvavxirho": 0, "dtinvavyirho": 1, "dtinvavzirho": 2, "dtK": 3, "dtlambda": 4, "dtM": 5,
               "dtavmu": 6, "dtavxzimu": 6, "dtavxyimu": 7, "dtavyzimu": 8}


def make_cfg(case):
    cfg = E.GpiConfig()
    nd = case["ndims"]
    cfg.abi_version, cfg.ndims, cfg.order = E.ABI_VERSION, nd, case["order"]
    cfg.physics = E.ACOUSTIC if case["physics"] == "acoustic" else E.ELASTIC
    n = case["n"]
    cfg.n[0], cfg.n[1], cfg.n[2] = (n[0], n[1], n[2]) if nd == 3 else (n[0], 1, n[1])
    cfg.nt, cfg.npml, cfg.nbound = case["nt"], refcases.npml_of(case["order"]), 3
    cfg.pml_faces = E.face_mask(case["pml_faces"])
    cfg.rigid_faces = E.face_mask(set(case["pml_faces"]) | set(case.get("rigid_faces", [])))        # fdtd.jl:215
    cfg.stressfree_faces = E.face_mask(case.get("stressfree_faces", []))
    cfg.npw, cfg.nshots, cfg.device = case.get("npw", 1), 1, -1
    cfg.store_boundary = 1 if case.get("gradient") else 0
    fc = refcases.fc(case)
    cfg.dt, cfg.dtI = float(fc["dt"]), float(fc["dtI"])
    for q, d in enumerate(["z", "y", "x"]):
        if nd == 2 and d == "y":
            cfg.d[q], cfg.dI[q] = 1.0, 1.0
        else:
            cfg.d[q], cfg.dI[q] = float(fc["d" + d]), float(fc["d" + d + "I"])
    return cfg


def drive(eng, case):
    """the reference's call order through the C ABI: medium, dmod, CPML, acquisition, wavelets, mod_x_proc!, results"""
    nd = case["ndims"]
    from geophyinv_jl_b200.host.grids import fields_of
    fields = fields_of(case["physics"], nd)
    shape = {f: eng.field_shape(f) for f in fields}
    inp = refcases.build_inputs(case, shape)
    for name, a in inp["mod"].items():
        eng.set_medium(name, np.asfortranarray(a))
    eng.update_dmod()
    for df, v in inp["pml"].items():
        eng.set_pml(df, v["a"], v["b"], v["kI"])
    for f, csc in inp["spray"].items():
        eng.set_sparse(E.SPRAY, 0, 0, f, *csc)
    for f, csc in inp["recv"].items():
        for ipw in range(case.get("npw", 1)):
            eng.set_sparse(E.INTERP, ipw, 0, f, *csc)
    out = {"shape": shape}
    nr = {f: len(csc[0]) - 1 for f, csc in inp["recv"].items()}
    fwd = {f: np.asfortranarray(w) for f, w in inp["wavelets"].items()}
    if case.get("gradient"):
        for f, w in fwd.items():
            eng.set_wavelets(0, 0, f, w)
        eng.reset(E.RESET_WAVEFIELDS | E.RESET_RECORDS | E.RESET_GRADIENTS | E.RESET_BOUNDARY)
        eng.run("forward_save", [1], [True, False])
        out["rec"] = {f: eng.get_records(0, 0, f, nr[f]) for f in nr}
        for f, w in refcases.backward_wavelets(fwd).items():
            eng.set_wavelets(0, 0, f, np.asfortranarray(w))
        for f, w in refcases.adjoint_wavelets(case, out["rec"]).items():
            eng.set_wavelets(1, 0, f, np.asfortranarray(w))
        eng.reset(E.RESET_WAVEFIELDS | E.RESET_GRADIENTS)
        eng.run("adjoint", [1, 2], [True, True])
        out["grad"] = {k: eng.get_gradient(k) for k in ("invK", "rho")}
    else:
        for f, w in fwd.items():
            eng.set_wavelets(0, 0, f, w)
        eng.reset(E.RESET_WAVEFIELDS | E.RESET_RECORDS)
        eng.run("forward", [1], [True])
        out["rec"] = {f: eng.get_records(0, 0, f, nr[f]) for f in nr}
    from geophyinv_jl_b200.host.grids import wavefields_of
    out["fld"] = [{f: eng.get_field(ipw, f) for f in wavefields_of(case["physics"], nd)} for ipw in range(case.get("npw", 1))]
    return out


def sample(a, stride=None):
    stride = stride or (3 if a.ndim == 2 else 5)
    return np.ascontiguousarray(a[tuple(slice(None, None, stride) for _ in range(a.ndim))])


def compare(out, gold, name, what):
    bad = []
    shapes = dict(s.split(":") for s in gold["shapes"])
    for f, sh in out["shape"].items():
        if tuple(int(x) for x in shapes[f].split(",")) != tuple(sh):
            bad.append(f"shape of {f}: {sh} vs reference {shapes[f]}")
    for f, r in out["rec"].items():
        g = gold[f"rec_{f}"]
        assert np.abs(g).max() > 0 or f not in ("p", "vz"), f"{name}: empty reference record {f}"
        if not np.array_equal(r, g):
            bad.append(f"records {f}: max abs diff {np.abs(r.astype(np.float64) - g).max():.3e} (max |ref| {np.abs(g).max():.3e}), {np.count_nonzero(r != g)} of {g.size} differ")
    for ipw, flds in enumerate(out["fld"]):
        for f, a in flds.items():
            g = gold[f"fld{ipw}_{f}"]
            if not np.array_equal(sample(a), g):
                bad.append(f"final {f} (pw {ipw + 1}): max abs diff {np.abs(sample(a).astype(np.float64) - g).max():.3e} (max |ref| {np.abs(g).max():.3e})")
            bits = (np.ascontiguousarray(a, F32) + F32(0)).view(np.uint32).astype(np.uint64).ravel()      # every entry, not only the sample
            cs = np.array([np.add.reduce(bits), np.bitwise_xor.reduce(bits)], np.uint64)
            if not np.array_equal(cs, gold[f"sum{ipw}_{f}"]):
                bad.append(f"checksum of final {f} (pw {ipw + 1}): {cs} vs {gold[f'sum{ipw}_{f}']}")
    for k, g in out.get("grad", {}).items():
        if not np.array_equal(g, gold[f"grad_{k}"]):
            bad.append(f"gradient {k}: rel-L2 {np.linalg.norm(g.astype(np.float64) - gold[f'grad_{k}']) / np.linalg.norm(gold[f'grad_{k}']):.3e}")
    for k, a in out.get("dmod", {}).items():
        if not np.array_equal(sample(a), gold[f"dmod_{k}"]):
            bad.append(f"dmod {k}: {np.count_nonzero(sample(a) != gold[f'dmod_{k}'])} sampled entries differ")
    assert not bad, f"{name} [{what}] differs from the reference-text fixture:\n  " + "\n  ".join(bad)


def oracle_run(name, lit):
    import oracle as O
    case = refcases.CASES[name]
    eng = O.OracleEngine(make_cfg(case), np.float32, literals=lit)
    out = drive(eng, case)
    gold = np.load(os.path.join(GOLD, f"ref_{name}_{lit}.npz"))
    out["dmod"] = {}
    for k in [x[5:] for x in gold.files if x.startswith("dmod_")]:
        out["dmod"][k] = eng.get_dmod(DMOD_ORACLE[k])
    compare(out, gold, name, f"oracle, {lit} literals")
    eng.close()


@pytest.mark.parametrize("lit", ["f32", "f64"])
@pytest.mark.parametrize("name", list(refcases.CASES))
def test_oracle_reproduces_the_reference_text(name, lit):
    oracle_run(name, lit)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(refcases.CASES))
def test_cuda_engine_reproduces_the_reference_text(name):
    """the product path against the fixtures directly -- no oracle in the loop"""
    case = refcases.CASES[name]
    eng = G.Engine(make_cfg(case))
    out = drive(eng, case)
    gold = np.load(os.path.join(GOLD, f"ref_{name}_f32.npz"))
    compare(out, gold, name, "CUDA engine")


def test_fixture_provenance():
    """every fixture names the digest of the reference files it was generated from, and both literal typings exist for every case"""
    digests = set()
    for name in refcases.CASES:
        for lit in ("f32", "f64"):
            g = np.load(os.path.join(GOLD, f"ref_{name}_{lit}.npz"))
            digests.add(str(g["reference_sha256"]))
    assert len(digests) == 1 and len(next(iter(digests))) == 64
