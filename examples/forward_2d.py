#!/usr/bin/env python
"""Forward modelling with the B200 engine behind the reference's interface (needs a CUDA device).

The reference (README of pawbz/GeoPhyInv.jl, test/fdtd) writes

    pa = SeisForwExpt(FdtdAcoustic(); medium, ageom, srcwav, tgrid, rfields=[:p], pml_faces=[...])
    update!(pa)
    d = pa[:data, 1][1].d[:p]

and this is the same call sequence through the Python twin of the Julia shim."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geophyinv_jl_b200 as G

# medium: 2 km x 2 km, 10 m cells, vp = 2500 m/s, rho = 2500 kg/m^3  (Medium(:acou_homo2D), media/gallery.jl:14-21)
grid = [G.StepRange.from_stop(-1000.0, 1000.0, 201)] * 2
medium = G.Medium(grid, np.full((201, 201), 2500.0, np.float32), np.full((201, 201), 2500.0, np.float32))
# acquisition: one source, 64 receivers in a well on the other side  (AGeom(mgrid, :xwell, SSrcs(1), Recs(64)))
ageom = G.ageom_xwell(grid, nss=1, nr=64)
tgrid = G.StepRange(0.0, 2e-3, 1000)
srcwav = G.make_srcwav(tgrid, ageom, ["p"], G.ricker(10.0, tgrid, tpeak=0.15))

pa = G.SeisForwExpt(G.FdtdAcoustic(), medium=medium, ageom=ageom, srcwav=srcwav, tgrid=tgrid, rfields=["p"],
                    pml_faces=["zmin", "zmax", "xmin", "xmax"])
t = G.update(pa)                                   # update!(pa)
d = pa["data", 1][0].d["p"]                        # pa[:data, 1][1].d[:p], (nt x nr)
print(f"records {d.shape}, max |p| = {np.abs(d).max():.4e}; {t['run_ms']:.1f} ms on the device, "
      f"{t['cell_updates'] / t['run_ms'] / 1e6:.1f} Gcell-updates/s, kernels: {pa.engine.kernel_family()}")

# a new medium, same experiment: update!(pa, medium); update!(pa)
medium.vp[80:120, 80:120] *= np.float32(1.1)
G.update(pa, medium)
G.update(pa)
print(f"after update!(pa, medium): max |p| = {np.abs(pa['data', 1][0].d['p']).max():.4e}")
