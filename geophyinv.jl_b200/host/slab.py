"""Large 3-D experiments over z-slabs without ever materialising the global medium on one host.

`PFdtd(..., zslab=(rank, nranks))` (host/fdtd.py) is the drop-in route: every rank holds the reference's
global host objects and the engine picks its rows.  At BASELINE config 5 (594 x 1106 x 1106 extended cells,
8.7 GB of medium per copy) that is wasteful, so this module drives the same C ABI with per-rank rows:
the caller supplies `rows(name, k0, k1) -> [k1-k0, ny, nx]` for the EXTENDED medium parameters
(`invlambda`, `invmu`, `rho` or `invK`, `rho`) and everything else (CPML profiles, spray / interpolation
weights, wavelets) is built exactly as `PFdtd` builds it (fdtd.jl:137-280, cpml.jl:144-155, ageom.jl:33-58).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from .. import engine as E
from .cpml import pml_coefficients
from .data import AGeomss, findfreq, padmgrid
from .grids import NBOUND, NPML, ORDER, StepRange, dfields_of
from .proj import get_proj_matrix

F32 = np.float32
ALL6 = ("zmin", "zmax", "ymin", "ymax", "xmin", "xmax")


class SlabExpt:
    def __init__(self, physics: str, grid: Sequence[StepRange], tgrid: StepRange, ageomss: AGeomss, sfield: str,
                 wavelet: np.ndarray, rfields: Sequence[str], rows: Callable[[str, int, int], np.ndarray],
                 vp_bounds, rank: int, nranks: int, device: int = -1, pml_faces=ALL6, chunk: int = 32):
        assert len(grid) == 3
        self.physics, self.rank, self.nranks = physics, rank, nranks
        self.exgrid = padmgrid(list(grid), NPML, pml_faces)
        n = [len(g) for g in self.exgrid]
        self.n, self.nt, self.rfields, self.ageomss = n, len(tgrid), list(rfields), ageomss
        cfg = E.GpiConfig()
        cfg.abi_version, cfg.ndims, cfg.order = E.ABI_VERSION, 3, ORDER
        cfg.physics = E.ACOUSTIC if physics == "acoustic" else E.ELASTIC
        cfg.n[0], cfg.n[1], cfg.n[2] = n
        cfg.nt, cfg.npml, cfg.nbound = self.nt, NPML, NBOUND
        cfg.pml_faces = cfg.rigid_faces = E.face_mask(pml_faces)
        cfg.npw, cfg.nshots, cfg.device = 1, 1, device
        cfg.slab_rank, cfg.slab_nranks = rank, nranks
        dt = F32(tgrid.step)
        cfg.dt, cfg.dtI = float(dt), float(F32(1.0 / tgrid.step))
        for q, g in enumerate(self.exgrid):
            cfg.d[q], cfg.dI[q] = float(F32(g.step)), float(F32(1.0 / g.step))
        self.cfg = cfg
        self.engine = E.Engine(cfg)
        self.ka, self.kb = self.engine.slab_range()
        # medium rows: owned planes plus one halo plane on each side, in chunks to bound host memory
        names = ["invK", "rho"] if physics == "acoustic" else ["invlambda", "invmu", "rho"]
        lo, hi = max(self.ka - 1, 0), min(self.kb + 1, n[0])
        for name in names:
            for k0 in range(lo, hi, chunk):
                k1 = min(k0 + chunk, hi)
                self.engine.set_medium_rows(name, np.asarray(rows(name, k0, k1), F32), k0)
        self.engine.update_dmod()
        # acquisition and wavelets (one supersource)
        pts = lambda c, m: [[c[d][i] for d in ("z", "y", "x")] for i in range(m)]
        cp, rv, nz, _ = get_proj_matrix(sfield, self.exgrid, pts(ageomss.s, ageomss.ns))
        self.engine.set_sparse(E.SPRAY, 0, 0, sfield, cp, rv, nz)
        for rf in self.rfields:
            cp, rv, nz, _ = get_proj_matrix(rf, self.exgrid, pts(ageomss.r, ageomss.nr))
            self.engine.set_sparse(E.INTERP, 0, 0, rf, cp, rv, nz)
        w = np.zeros((self.nt, ageomss.ns), F32, order="F")
        w[:, :] = np.asarray(wavelet, np.float64)[: self.nt, None].astype(F32)
        self.engine.set_wavelets(0, 0, sfield, w)
        freqpeak = F32(findfreq(w, tgrid, "peak")) if np.abs(w).max() > 0 else F32(0)
        velavg = F32((F32(vp_bounds[0]) + F32(vp_bounds[1])) / F32(2))
        pml = pml_coefficients(dfields_of(physics, 3), self.exgrid, list(grid), pml_faces, float(dt), float(velavg), float(freqpeak), NPML)
        for df, (a, b, kI) in pml.items():
            self.engine.set_pml(df, a, b, kI)

    def attach_nccl(self, dist):
        from . import dist as D
        if self.nranks > 1:
            uid = D.share_unique_id(self.engine.nccl_unique_id, dist)
            self.engine.nccl_init(uid, self.rank, self.nranks)

    def update(self):
        """`update!(pa)` for the one supersource: reset, run, records (complete on every rank)."""
        self.engine.reset(E.RESET_WAVEFIELDS | E.RESET_RECORDS)
        self.engine.run("forward", [1], [True])
        return self.engine.timers()

    def records(self, field: str):
        return self.engine.get_records(0, 0, field, self.ageomss.nr)


def synthetic_rows(n_interior, exn, seed=1234, vp0=3000.0, vs0=1732.0, rho0=2300.0, sigma=0.02, faces=ALL6):
    """Row generator of the C3 / C5 synthetic elastic medium: each interior z plane is drawn from its own
    seeded stream (so every rank produces identical halo rows), then replicate-padded into the PML like
    `padarray` (media.jl:260-304).  Returns rows(name, k0, k1) for invlambda / invmu / rho."""
    nzi, nyi, nxi = n_interior
    padz = NPML if "zmin" in faces else 0
    pady = (NPML if "ymin" in faces else 0, NPML if "ymax" in faces else 0)
    padx = (NPML if "xmin" in faces else 0, NPML if "xmax" in faces else 0)
    cache = {}

    def plane(kint):
        if kint not in cache:
            if len(cache) > 40:
                cache.clear()
            rng = np.random.default_rng([seed, kint])
            vp = (vp0 * (1 + sigma * rng.standard_normal((nyi, nxi), dtype=np.float32))).astype(F32)
            vs = (vs0 * (1 + sigma * rng.standard_normal((nyi, nxi), dtype=np.float32))).astype(F32)
            rho = (rho0 * (1 + sigma * rng.standard_normal((nyi, nxi), dtype=np.float32))).astype(F32)
            mu = vs * vs * rho
            lam = (vp * vp - F32(2) * (vs * vs)) * rho
            pad = lambda a: np.pad(a, (pady, padx), mode="edge")
            cache[kint] = {"invlambda": pad(F32(1) / lam), "invmu": pad(F32(1) / mu), "rho": pad(rho)}
        return cache[kint]

    def rows(name, k0, k1):
        out = np.empty((k1 - k0, exn[1], exn[2]), F32)
        for k in range(k0, k1):
            out[k - k0] = plane(min(max(k - padz, 0), nzi - 1))[name]
        return out

    return rows
