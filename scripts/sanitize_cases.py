#!/usr/bin/env python
"""Small end-to-end runs of every kernel family, meant to be executed under compute-sanitizer
(memcheck / racecheck / initcheck) on the GPU box:  compute-sanitizer --tool memcheck python scripts/sanitize_cases.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import gallery  # noqa: E402

which = sys.argv[1:] or ["all"]


def want(name):
    return "all" in which or name in which


def run(label, pa):
    pa.update()
    tot = sum(float(np.abs(r.d[f]).sum()) for r in pa.c.data[0] for f in pa.c.rfields)
    assert np.isfinite(tot)
    print(f"{label}: ok, sum|records| = {tot:.6e}", flush=True)


if want("acou2d"):
    for order in (2, 4):
        run(f"2-D acoustic order {order}", G.SeisForwExpt(G.FdtdAcoustic(), **gallery.c2_acou2d_layered(nz=60, nx=75, nt=40, nss=3, nr=8, fq=20.0, rfields=("p", "vx")), shot_batch=2, order=order))
if want("elastic2d"):
    for order in (2, 4):
        run(f"2-D elastic free surface order {order}", G.SeisForwExpt(G.FdtdElastic(), **gallery.elastic2d(nz=50, nx=61, nt=40, nr=6, stressfree=True), order=order))
if want("acou3d"):
    for order in (2, 4):
        run(f"3-D acoustic order {order}", G.SeisForwExpt(G.FdtdAcoustic(), **gallery.acou3d(n=17, nt=15, nr=4), order=order))
if want("elastic3d"):
    for order in (2, 4):
        run(f"3-D elastic order {order} (TMA tiles + shell at order 2)", G.SeisForwExpt(G.FdtdElastic(), **gallery.c3_elastic3d(n=21, nt=12, nr=4, fq=40.0, rfields=("vz", "vx"), stressfree=(order == 2)), order=order))
    os.environ["GPI_TMA3"] = "0"
    run("3-D elastic register-staged float4 kernels", G.SeisForwExpt(G.FdtdElastic(), **gallery.c3_elastic3d(n=19, nt=10, nr=4, fq=40.0)))
    os.environ["GPI_SCALAR3D"] = "1"
    run("3-D elastic scalar kernels", G.SeisForwExpt(G.FdtdElastic(), **gallery.c3_elastic3d(n=18, nt=8, nr=4, fq=40.0)))
    del os.environ["GPI_TMA3"], os.environ["GPI_SCALAR3D"]
if want("illum"):      # illum_flag: k_illum per step (Float64 accumulators per resident shot), k_axpy1d stack; 2-D launches carry the programmatic-launch attribute
    pa = G.SeisForwExpt(G.FdtdAcoustic(), **gallery.c2_acou2d_layered(nz=60, nx=75, nt=40, nss=3, nr=8, fq=20.0), shot_batch=2, illum_flag=True)
    run("2-D acoustic with illumination", pa)
    assert np.isfinite(pa["illum"]).all() and pa["illum"].max() > 0
if want("gradient"):
    for order in (2, 4):
        kw, true = gallery.c4_fwi2d(nz=40, nx=50, nt=60, nss=3, nr=8, fq=15.0)
        pa = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, shot_batch=2, order=order)
        dobs = [d.copy() for d in pa.c.data[0]]
        m = pa.get_modelvector(); g = np.zeros_like(m)
        G.gradient(g, m, dobs, pa)
        assert np.isfinite(g).all()
        print(f"2-D FWI gradient order {order}: ok, |g|max = {np.abs(g).max():.3e}", flush=True)
    kw, true = gallery.fwi3d(n=14, nt=30, nr=4, nss=1)
    pa = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw)
    dobs = [d.copy() for d in pa.c.data[0]]
    m = pa.get_modelvector(); g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
    print(f"3-D FWI gradient: ok, |g|max = {np.abs(g).max():.3e}", flush=True)
if want("born"):
    kw = gallery.c1_acou2d_homo(nz=41, nx=47, nt=50, nr=6, sfield="p", rfields=("p", "vz"), fq=12.0, dt=1.8e-3)
    pb = G.SeisForwExpt(G.FdtdAcoustic(born=True), **kw)
    mp = kw["medium"].copy(); mp.vp[15:25, 18:28] *= np.float32(1.02)
    G.update(pb, kw["medium"], mp)
    pb.update()
    print(f"FD-Born: ok, sum|scattered| = {float(np.abs(pb.c.data[1][0].d['p']).sum()):.6e}", flush=True)
print("ALL CASES DONE", flush=True)
