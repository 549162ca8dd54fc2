#!/usr/bin/env python
"""Generates the golden vectors in tests/golden/ from the CPU oracle (oracle/fdtd_oracle.c, Float32 build).

The reference ships no golden vectors for this path and Julia is absent from this image, so
these are regression pins of the restatement -- itself pinned by the reference's own invariants in
tests/test_oracle_invariants.py -- not outputs of the Julia code.  Each fixture stores a decimated view of the
receiver records / gradients plus float64 checksums of the full arrays, so the files stay small.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import gallery  # noqa: E402
import oracle as O  # noqa: E402


def summary(a):
    a = np.asarray(a, np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), np.abs(a).max()])


CASES = {
    # name: (attrib factory, gallery kwargs builder, record fields)
    "c1_acou2d_p": (G.FdtdAcoustic, lambda: gallery.c1_acou2d_homo(sfield="p", rfields=("p",)), ("p",)),
    "c1_acou2d_vz": (G.FdtdAcoustic, lambda: gallery.c1_acou2d_homo(sfield="vz", rfields=("vz", "vx", "p")), ("vz", "vx", "p")),
    "elastic2d": (G.FdtdElastic, lambda: gallery.elastic2d(), ("vz", "vx")),
    "elastic2d_freesurface": (G.FdtdElastic, lambda: gallery.elastic2d(stressfree=True), ("vz", "vx")),
    "acou3d": (G.FdtdAcoustic, lambda: gallery.acou3d(), ("p", "vx")),
    "c3_elastic3d_n40": (G.FdtdElastic, lambda: gallery.c3_elastic3d(n=40, nt=150, nr=12, fq=25.0, rfields=("vz", "vx", "vy")), ("vz", "vx", "vy")),
    # _fd_order = 4 (SeisForwExpt(...; order=4)): heterogeneous 2-D acoustic, 2-D elastic with the free surface, 3-D elastic
    "o4_acou2d": (G.FdtdAcoustic, lambda: dict(gallery.c2_acou2d_layered(nz=90, nx=140, nt=400, nss=2, nr=20, fq=15.0, rfields=("p", "vx")), order=4), ("p", "vx")),
    "o4_elastic2d_freesurface": (G.FdtdElastic, lambda: dict(gallery.elastic2d(nt=350, stressfree=True), order=4), ("vz", "vx")),
    "o4_elastic3d_n30": (G.FdtdElastic, lambda: dict(gallery.c3_elastic3d(n=30, nt=110, nr=10, fq=25.0, rfields=("vz", "vx", "vy")), order=4), ("vz", "vx", "vy")),
}


def records_case(name):
    attrib, build, fields = CASES[name]
    po = O.OraclePFdtd(attrib(), **build())
    po.update()
    out = {}
    for iss, rec in enumerate(po.c.data[0]):
        for f in fields:
            d = rec.d[f]
            out[f"s{iss}_{f}_dec"] = d[::8, ::3].copy()             # every 8th sample, every 3rd receiver
            out[f"s{iss}_{f}_sum"] = summary(d)
    return out


def gradient_case():
    kw, true = gallery.c4_fwi2d(nz=70, nx=110, nt=500, nss=3, nr=24, fq=10.0)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true}); pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw)
    m = pa.get_modelvector()
    g = np.zeros_like(m)
    loss = G.gradient(g, m, dobs, pa)
    half = g.size // 2
    return {"loss": np.array([loss]), "gK_dec": g[:half:7].copy(), "gR_dec": g[half::7].copy(),
            "gK_sum": summary(g[:half]), "gR_sum": summary(g[half:])}


if __name__ == "__main__":
    O.build()
    only = sys.argv[1:]                      # optional: names of the fixtures to (re)write
    for name in CASES:
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **records_case(name))
        print("wrote", name)
    if not only or "c4_fwi2d_gradient" in only:
        np.savez_compressed(os.path.join(HERE, "c4_fwi2d_gradient.npz"), **gradient_case())
        print("wrote c4_fwi2d_gradient")
