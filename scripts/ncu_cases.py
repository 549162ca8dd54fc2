#!/usr/bin/env python
"""Short runs of one kernel family each, sized like the BASELINE configs, for `ncu --set full` captures of EVERY stencil kernel
(scripts/gpu_ncu_all.sh).  usage: ncu_cases.py <case>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import gallery  # noqa: E402

case = sys.argv[1]
if case == "acou2d":            # C2 grid, 8 resident supersources: k_vel2v<0>, k_stress2v<0>
    pa = G.SeisForwExpt(G.FdtdAcoustic(), **gallery.c2_acou2d_layered(nt=40, nss=8))
    pa.update()
elif case == "elastic2d":       # C2-sized elastic grid, 8 supersources: k_vel2v<1>, k_stress2v<1>
    pa = G.SeisForwExpt(G.FdtdElastic(), **gallery.elastic2d(nz=350, nx=1700, nt=40, nr=32, nss=8))
    pa.update()
elif case == "acou3d":          # 256^3 acoustic: k_vel3v<0>, k_stress3v<0>
    pa = G.SeisForwExpt(G.FdtdAcoustic(), **gallery.acou3d(n=256, nt=24, nr=16))
    pa.update()
elif case == "elastic3d_o4":    # C3 grid at order 4: k_vel4<3,1>, k_dirichlet4<3>, k_stress4<3,1>
    pa = G.SeisForwExpt(G.FdtdElastic(), **gallery.c3_elastic3d(n=256, nt=12, nr=16), order=4)
    pa.update()
elif case == "acou2d_o4":       # C2 grid at order 4: k_vel4<2,0>, k_stress4<2,0>
    pa = G.SeisForwExpt(G.FdtdAcoustic(), **gallery.c2_acou2d_layered(nt=40, nss=8), order=4)
    pa.update()
elif case == "gradient2d":      # C4 grid, 8 supersources, short: k_boundary<0|1>, k_grad2d, merged two-wavefield launches
    kw, true = gallery.c4_fwi2d(nt=40, nss=8)
    pa = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw)
    dobs = [d.copy() for d in pa.c.data[0]]
    m = pa.get_modelvector(); g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
elif case == "gradient3d_el":   # 3-D elastic adjoint on a 160^3 medium: six-field boundary store, k_grad3d_el
    kw, true = gallery.fwi3d_elastic(n=160, nt=16, nr=8)
    pa = G.PFdtd(G.FdtdElastic("forward_save"), **kw)
    dobs = [d.copy() for d in pa.c.data[0]]
    m = pa.get_modelvector(); g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
else:
    raise SystemExit(f"unknown case {case}")
print(case, "done", flush=True)
