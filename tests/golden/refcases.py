"""TEST INFRASTRUCTURE -- the cases of the reference-pinned golden vectors and their INPUTS.

Shared by tests/golden/from_reference.py (which evaluates the reference's source text on these inputs) and by
tests/test_reference_pinned.py (which feeds the same inputs to the CPU oracle / the CUDA engine through the C ABI), so
that the fixtures hold outputs only.  Inputs are closed-form functions of the array indices (no RNG stream, nothing
from host/): medium parameters on the EXTENDED grid, CPML coefficient vectors, spray / interpolation matrices as CSC
with 1-based rows into the field's own array, wavelets [nt, ns].  They are deliberately generic (heterogeneous medium,
sources inside the CPML slabs and next to corners, receivers everywhere) so that every term of every kernel matters
within the few time steps a fixture holds.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
ALL4 = ["zmin", "zmax", "xmin", "xmax"]
ALL6 = ["zmin", "zmax", "ymin", "ymax", "xmin", "xmax"]


def npml_of(order):
    return 40 + order - 1


def _ext(interior, faces, dims, order):
    return tuple(m + npml_of(order) * sum(1 for f in faces if f[0] == d) for m, d in zip(interior, dims))


def _case(physics, ndims, order, interior, nt, faces, sources, receivers, **kw):
    dims = ["z", "y", "x"] if ndims == 3 else ["z", "x"]
    return dict(physics=physics, ndims=ndims, order=order, n=_ext(interior, faces, dims, order), nt=nt, pml_faces=list(faces),
                sources=sources, receivers=receivers, **kw)


# sources / receivers: field -> list of (fractional position in the field's own array, per axis, in [0, 1])
_S2 = [(0.50, 0.52), (0.93, 0.06)]
_R2 = [(0.45, 0.50), (0.55, 0.58), (0.97, 0.05), (0.90, 0.10), (0.02, 0.50), (0.50, 0.985)]
_S3 = [(0.50, 0.52, 0.48), (0.93, 0.07, 0.10)]
_R3 = [(0.45, 0.50, 0.50), (0.55, 0.58, 0.44), (0.96, 0.05, 0.08), (0.90, 0.10, 0.12), (0.03, 0.50, 0.50), (0.50, 0.98, 0.52)]

CASES = {
    # 2-D acoustic, CPML on four faces; pressure and body-force sources; order 2 and 4
    "acou2d_o2": _case("acoustic", 2, 2, (16, 18), 28, ALL4, {"p": _S2, "vz": _S2[:1]}, {"p": _R2, "vx": _R2, "vz": _R2}),
    "acou2d_o4": _case("acoustic", 2, 4, (16, 18), 24, ALL4, {"p": _S2, "vx": _S2[1:]}, {"p": _R2, "vx": _R2, "vz": _R2}),
    # 2-D elastic with a free surface on zmin (no CPML there; one source right under it so that the mirror / zeroing of the surface rows acts
    # on a live field), and with four CPML faces at order 4
    "el2d_o2": _case("elastic", 2, 2, (40, 18), 28, ["zmax", "xmin", "xmax"], {"vz": [(0.04, 0.50)] + _S2[1:], "tauxx": [(0.05, 0.45)]}, {"vx": _R2, "vz": _R2},
                     stressfree_faces=["zmin"]),
    "el2d_o4": _case("elastic", 2, 4, (16, 18), 24, ALL4, {"vx": _S2, "vz": _S2[:1]}, {"vx": _R2, "vz": _R2}),
    # 3-D acoustic: six faces (order 2); three faces (order 4: keeps the grid small, exercises the no-CPML side of every axis)
    "acou3d_o2": _case("acoustic", 3, 2, (10, 12, 11), 14, ALL6, {"p": _S3, "vy": _S3[:1]}, {"p": _R3, "vx": _R3, "vy": _R3, "vz": _R3}),
    "acou3d_o4": _case("acoustic", 3, 4, (50, 52, 51), 12, ["zmax", "ymin", "xmax"], {"p": _S3}, {"p": _R3, "vz": _R3}),
    # 3-D elastic: the roofline case's operators, six CPML faces (order 2); free surface + three faces (order 4)
    "el3d_o2": _case("elastic", 3, 2, (10, 12, 11), 14, ALL6, {"vz": _S3, "tauxx": _S3[:1]}, {"vx": _R3, "vy": _R3, "vz": _R3}),
    "el3d_o4": _case("elastic", 3, 4, (52, 50, 51), 12, ["zmax", "ymax", "xmin"], {"vz": [(0.06, 0.50, 0.50)] + _S3[1:], "vx": _S3[1:]}, {"vx": _R3, "vy": _R3, "vz": _R3},
                     stressfree_faces=["zmin"]),
    # 2-D acoustic FWI: forward_save (boundary store) then adjoint (save_tp!, boundary force, adjoint sources, imaging of invK and rho)
    "fwi2d_o2": _case("acoustic", 2, 2, (16, 18), 28, ALL4, {"vz": _S2[:1], "vx": _S2[:1]}, {"vz": _R2[:4], "vx": _R2[:4]}, npw=2, gradient=True),
}


def dims_of(case):
    return ["z", "y", "x"] if case["ndims"] == 3 else ["z", "x"]


def _idx(shape):
    return np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")


def medium(case):
    """independent parameters `mod` on the extended grid (medium.jl:81-95), Float32"""
    g = _idx(case["n"])
    ph = [0.37, 0.21, 0.13][: len(g)] if case["ndims"] == 3 else [0.37, 0.13]
    a = sum(p * x for p, x in zip(ph, g))
    b = sum(p * x for p, x in zip(ph[::-1], g))
    vp = 2500.0 * (1 + 0.08 * np.sin(a + 0.5))
    rho = 2200.0 * (1 + 0.06 * np.cos(b - 0.3))
    if case["physics"] == "acoustic":
        return {"invK": (1.0 / (vp * vp * rho)).astype(F32), "rho": rho.astype(F32)}
    vs = vp / 1.8 * (1 + 0.05 * np.cos(0.5 * a + 0.2 * b))
    mu = vs * vs * rho
    lam = vp * vp * rho - 2 * mu
    return {"invlambda": (1.0 / lam).astype(F32), "invmu": (1.0 / mu).astype(F32), "rho": rho.astype(F32)}


def fc(case):
    """get_fc (fdtd.jl:313-332): Float64 on the host, then Data.Number (Float32); d?I = inv(d? * 24) at order 4"""
    ds = [10.0, 12.5, 8.0] if case["ndims"] == 3 else [10.0, 8.0]
    scale = {2: 1.0, 4: 24.0}[case["order"]]
    dt = 1e-3
    out = {"dt": F32(dt), "dtI": F32(1.0 / dt)}
    for d, s in zip(dims_of(case), ds):
        out["d" + d] = F32(s)
        out["d" + d + "I"] = F32(1.0 / (s * scale))
    return out


def pml_vectors(case, dfields):
    """a, b, kI (2 npml each: min slab then max slab) per derivative field; damping-like values, a little different per field"""
    npml = npml_of(case["order"])
    out = {}
    for f in sorted(dfields):
        h = (sum(ord(c) * (i + 1) for i, c in enumerate(f)) % 17) / 170.0
        s = np.arange(npml, dtype=np.float64)
        an = np.concatenate([(npml - s) / npml, (s + 1) / npml])
        d = (0.55 + h) * an ** 2
        b = np.exp(-(d + 0.02))
        a = -(0.4 - h) * an ** 2 * (1 - b)
        kI = 1.0 / (1.0 + (0.25 + h) * an ** 2)
        out[f] = {"a": a.astype(F32), "b": b.astype(F32), "kI": kI.astype(F32)}
    return out


def proj(shape, points):
    """CSC (colptr, rowval, nzval), 1-based, one column per point: the 2^N cells around a fractional position of the FIELD's own array,
    multilinear weights, stored entries sorted by row (what `sparsevec` / `get_proj_matrix` produce, proj_mat.jl:229-247)"""
    colptr, rowval, nzval = [1], [], []
    nd = len(shape)
    for p in points:
        base, frac = [], []
        for q in range(nd):
            x = p[q] * (shape[q] - 1)
            i0 = min(int(np.floor(x)), shape[q] - 2)
            base.append(i0)
            frac.append(x - i0)
        ent = []
        for corner in range(1 << nd):
            w, lin, stride = 1.0, 0, 1
            for q in range(nd):
                bit = (corner >> q) & 1
                w *= frac[q] if bit else (1 - frac[q])
                lin += (base[q] + bit) * stride
                stride *= shape[q]
            ent.append((lin + 1, F32(w)))
        ent.sort()
        rowval += [e[0] for e in ent]
        nzval += [e[1] for e in ent]
        colptr.append(len(rowval) + 1)
    return np.asarray(colptr, np.int64), np.asarray(rowval, np.int64), np.asarray(nzval, F32)


def wavelet(case, field, ns):
    nt = case["nt"]
    t = np.arange(nt, dtype=np.float64) * 1e-3
    out = np.zeros((nt, ns), F32)
    for s in range(ns):
        f0, t0 = 90.0 + 25.0 * s, 0.006 + 0.002 * s
        arg = (np.pi * f0 * (t - t0)) ** 2
        amp = (1e6 if field.startswith("v") else 1e3) * (1.0 - 0.3 * s)
        out[:, s] = (amp * (1 - 2 * arg) * np.exp(-arg)).astype(F32)
    return out


def build_inputs(case, shape):
    """`shape`: {field: array shape} for the case's physics / order (from the reference's fields.jl, or from gpi_field_shape_order)"""
    dfields = [f for f in shape if f.startswith("d") and len(f) > 3]
    return {"fc": fc(case), "mod": medium(case), "pml": pml_vectors(case, dfields),
            "spray": {f: proj(shape[f], pts) for f, pts in case["sources"].items()},
            "recv": {f: proj(shape[f], pts) for f, pts in case["receivers"].items()},
            "wavelets": {f: wavelet(case, f, len(pts)) for f, pts in case["sources"].items()}}


def backward_wavelets(fwd):
    """get_source(w, ::vz|vx|vy, ::Val{-1}) (source.jl:8-19): negate, circshift by -1, reverse, zero the first sample"""
    out = {}
    for f, w in fwd.items():
        ww = -np.asarray(w, F32)
        ww = np.roll(ww, -1, axis=0)
        ww = ww[::-1, :].copy()
        ww[0, :] = 0
        out[f] = ww
    return out


def adjoint_wavelets(case, records):
    """pw 2's sources of the adjoint pass: the forward records, time-reversed and scaled (any [nt, nr] block would do; this one has the
    right support in time)"""
    return {f: (np.asarray(r, F32)[::-1, :] * F32(1e3)).astype(F32) for f, r in records.items() if f.startswith("v")}
