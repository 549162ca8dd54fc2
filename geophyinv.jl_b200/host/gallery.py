"""Synthetic experiment gallery: the five BASELINE.json configurations (C1..C5)
plus down-sized variants for parity tests.  Mirrors the role of the reference's galleries
(src/media/gallery.jl:14-21, src/ageom/gallery.jl:64-115, src/fdtd/gallery.jl:2-32); the large
reference media (Marmousi2, overthrust) are git-LFS blobs that are absent, so everything is synthetic.

Each builder returns the keyword arguments of `SeisForwExpt` (so the same dict drives the CUDA engine
and, in tests, the CPU oracle).
"""
from __future__ import annotations

import numpy as np

from .data import AGeomss, Medium, ageom_xwell, make_srcwav, ricker
from .grids import StepRange

F32 = np.float32


def _ricker(fq, tgrid, tpeak):
    """Ricker on `tgrid`; when `tgrid` is shorter than the wavelet (bounded CPU samples of a big case),
    the wavelet is generated on a long enough grid of the same step and truncated."""
    need = int(np.ceil((tpeak + 1.5 / fq) / tgrid.step)) + 2
    if len(tgrid) >= need:
        return ricker(fq, tgrid, tpeak=tpeak)
    return ricker(fq, StepRange(tgrid.first, tgrid.step, need), tpeak=tpeak)[: len(tgrid)]


def c1_acou2d_homo(nz=201, nx=201, nt=1000, nr=64, sfield="p", rfields=("p",), nss=1, dt=2e-3, fq=10.0):
    """C1: 2-D acoustic homogeneous 201x201 (10 m), vp = rho = 2500, Ricker 10 Hz, :xwell geometry,
    PML on four faces, 1000 steps of 2 ms (media/gallery.jl:14-21, fdtd/gallery.jl:8-13)."""
    grid = [StepRange.from_stop(-1000.0, 1000.0, nz), StepRange.from_stop(-1000.0, 1000.0, nx)]
    medium = Medium.homogeneous(grid, 2500.0, 2500.0)
    tgrid = StepRange(0.0, dt, nt)
    ageom = ageom_xwell(grid, nss=nss, nr=nr)
    wav = _ricker(fq, tgrid, max(0.15, 1.5 / fq))
    if sfield != "p":
        wav = wav * 1e6                                   # fdtd/gallery.jl:12
    srcwav = make_srcwav(tgrid, ageom, [sfield], wav)
    return dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=list(rfields),
                pml_faces=["zmin", "zmax", "xmin", "xmax"])


def layered_vp(nz, nx, dz, nlayers=12, vp0=1500.0, vp1=4500.0, undulation=0.05, periods=1.5):
    """12 horizontal layers vp0 -> vp1 with +-5 % sinusoidal lateral undulation of the interfaces."""
    depth = np.arange(nz)[:, None] * dz
    total = (nz - 1) * dz
    x = np.arange(nx)[None, :] / max(nx - 1, 1)
    shift = undulation * total * np.sin(2 * np.pi * periods * x)
    layer = np.clip(np.floor((depth + shift) / total * nlayers), 0, nlayers - 1)
    return (vp0 + (vp1 - vp0) * layer / (nlayers - 1)).astype(F32)


def c2_acou2d_layered(nz=350, nx=1700, nt=4000, nss=64, nr=256, dt=1e-3, fq=8.0, sfield="p", rfields=("p",), d=10.0):
    """C2: Marmousi-sized layered synthetic: rho = 310 vp^0.25, sources and receivers at z = 20 m."""
    grid = [StepRange(0.0, d, nz), StepRange(0.0, d, nx)]
    vp = layered_vp(nz, nx, d)
    rho = (310.0 * vp.astype(np.float64) ** 0.25).astype(F32)
    medium = Medium(grid, vp, rho)
    tgrid = StepRange(0.0, dt, nt)
    xs = np.linspace(grid[1].first + 0.05 * (grid[1].last - grid[1].first), grid[1].last - 0.05 * (grid[1].last - grid[1].first), nss)
    xr = np.linspace(grid[1].first + 0.02 * (grid[1].last - grid[1].first), grid[1].last - 0.02 * (grid[1].last - grid[1].first), nr)
    zs = 2.0 * d
    ageom = [AGeomss({"z": [zs], "x": [x]}, {"z": np.full(nr, zs), "x": xr}) for x in xs]
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.02)
    if sfield != "p":
        wav = wav * 1e6
    srcwav = make_srcwav(tgrid, ageom, [sfield], wav)
    return dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=list(rfields),
                pml_faces=["zmin", "zmax", "xmin", "xmax"])


def c3_elastic3d(n=256, nt=2000, nr=64, dt=1e-3, fq=10.0, d=10.0, seed=1234, sfield="vz", rfields=("vz",), stressfree=False):
    """C3: 3-D isotropic elastic n^3 (10 m), vp 3000 / vs 1732 / rho 2300 each x (1 + 0.02 N(0,1)),
    PML on six faces, one :vz source at the centre, receivers on a line."""
    grid = [StepRange(0.0, d, n)] * 3
    rng = np.random.default_rng(seed)
    shp = (n, n, n)
    vp = (3000.0 * (1 + 0.02 * rng.standard_normal(shp))).astype(F32)
    vs = (1732.0 * (1 + 0.02 * rng.standard_normal(shp))).astype(F32)
    rho = (2300.0 * (1 + 0.02 * rng.standard_normal(shp))).astype(F32)
    medium = Medium(grid, vp, rho, vs)
    tgrid = StepRange(0.0, dt, nt)
    L = grid[0].last
    # off-node positions so all 8 trilinear taps are exercised
    src = {"z": [0.5 * L + 0.3 * d], "y": [0.5 * L + 0.2 * d], "x": [0.5 * L + 0.1 * d]}
    rec = {"z": np.full(nr, 0.25 * L + 0.4 * d), "y": np.full(nr, 0.5 * L + 0.2 * d), "x": np.linspace(0.1 * L, 0.9 * L, nr)}
    ageom = [AGeomss(src, rec)]
    tpeak = 1.5 / fq + 0.01
    wav = _ricker(fq, tgrid, tpeak) * (1e6 if sfield.startswith("v") else 1.0)
    srcwav = make_srcwav(tgrid, ageom, [sfield], wav)
    kw = dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=list(rfields))
    if stressfree:
        kw["pml_faces"] = ["zmax", "ymin", "ymax", "xmin", "xmax"]
        kw["rigid_faces"] = ["zmax", "ymin", "ymax", "xmin", "xmax"]
        kw["stressfree_faces"] = ["zmin"]
    return kw


def acou3d(n=48, nt=200, nr=16, dt=1e-3, fq=20.0, d=10.0, seed=7, sfield="p", rfields=("p", "vx")):
    """Small 3-D acoustic case (same template as C3)."""
    grid = [StepRange(0.0, d, n)] * 3
    rng = np.random.default_rng(seed)
    vp = (2500.0 * (1 + 0.03 * rng.standard_normal((n, n, n)))).astype(F32)
    rho = (2200.0 * (1 + 0.03 * rng.standard_normal((n, n, n)))).astype(F32)
    medium = Medium(grid, vp, rho)
    tgrid = StepRange(0.0, dt, nt)
    L = grid[0].last
    src = {"z": [0.5 * L + 0.3 * d], "y": [0.45 * L + 0.2 * d], "x": [0.4 * L + 0.1 * d]}
    rec = {"z": np.full(nr, 0.3 * L + 0.4 * d), "y": np.linspace(0.2 * L, 0.8 * L, nr), "x": np.linspace(0.1 * L, 0.9 * L, nr)}
    ageom = [AGeomss(src, rec)]
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.005) * (1e6 if sfield.startswith("v") else 1.0)
    srcwav = make_srcwav(tgrid, ageom, [sfield], wav)
    return dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=list(rfields))


def elastic2d(nz=120, nx=150, nt=400, nr=24, nss=2, dt=1e-3, fq=12.0, d=10.0, seed=3, sfield="vz", rfields=("vz", "vx"), stressfree=False):
    """Small 2-D elastic case (two supersources)."""
    grid = [StepRange(0.0, d, nz), StepRange(0.0, d, nx)]
    rng = np.random.default_rng(seed)
    vp = layered_vp(nz, nx, d, nlayers=5, vp0=2200.0, vp1=3600.0) * (1 + 0.01 * rng.standard_normal((nz, nx))).astype(F32)
    vs = (vp / 1.8).astype(F32)
    rho = (310.0 * vp.astype(np.float64) ** 0.25).astype(F32)
    medium = Medium(grid, vp.astype(F32), rho, vs)
    tgrid = StepRange(0.0, dt, nt)
    Lz, Lx = grid[0].last, grid[1].last
    ageom = []
    for iss in range(nss):
        sx = (0.3 + 0.4 * iss / max(nss - 1, 1)) * Lx + 0.37 * d
        ageom.append(AGeomss({"z": [0.15 * Lz + 0.21 * d], "x": [sx]}, {"z": np.full(nr, 0.1 * Lz + 0.6 * d), "x": np.linspace(0.05 * Lx, 0.95 * Lx, nr)}))
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.005) * (1e6 if sfield.startswith("v") else 1.0)
    srcwav = make_srcwav(tgrid, ageom, [sfield], wav)
    kw = dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=list(rfields),
              pml_faces=["zmin", "zmax", "xmin", "xmax"])
    if stressfree:
        kw["pml_faces"] = ["zmax", "xmin", "xmax"]
        kw["rigid_faces"] = ["zmax", "xmin", "xmax"]
        kw["stressfree_faces"] = ["zmin"]
    return kw


def c4_fwi2d(nz=350, nx=1700, nt=3000, nss=32, nr=128, dt=1e-3, fq=8.0, d=10.0, box=0.05):
    """C4: 2-D acoustic FWI gradient.  Model = C2 medium; 'observed' data come from the C2 medium with a
    +5 % vp box.  Source and records are :vz (adjoint injection exists only for velocity fields,
    source.jl:142-156).  Returns (kwargs for the model experiment, true medium)."""
    kw = c2_acou2d_layered(nz=nz, nx=nx, nt=nt, nss=nss, nr=nr, dt=dt, fq=fq, sfield="vz", rfields=("vz",), d=d)
    true = kw["medium"].copy()
    z0, z1, x0, x1 = int(0.4 * nz), int(0.6 * nz), int(0.4 * nx), int(0.6 * nx)
    true.vp[z0:z1, x0:x1] *= F32(1 + box)
    return kw, true


def fwi3d(n=24, nt=220, nr=12, nss=2, dt=1e-3, fq=18.0, d=10.0, box=0.05, seed=11):
    """3-D acoustic FWI gradient case: smooth random model, 'observed' data from the same medium
    with a +5 % vp and rho box.  Source and records are :vz / :vx (adjoint injection exists only for velocity fields)."""
    from scipy.ndimage import gaussian_filter
    grid = [StepRange(0.0, d, n)] * 3
    rng = np.random.default_rng(seed)
    vp = (2500.0 * (1 + 0.04 * gaussian_filter(rng.standard_normal((n, n, n)), 2.0) / 0.1)).astype(F32)
    rho = (2200.0 * (1 + 0.04 * gaussian_filter(rng.standard_normal((n, n, n)), 2.0) / 0.1)).astype(F32)
    medium = Medium(grid, vp, rho)
    tgrid = StepRange(0.0, dt, nt)
    L = grid[0].last
    ageom = []
    for iss in range(nss):
        sx = (0.25 + 0.5 * iss / max(nss - 1, 1)) * L + 0.13 * d
        src = {"z": [0.2 * L + 0.3 * d], "y": [0.45 * L + 0.2 * d], "x": [sx]}
        rec = {"z": np.full(nr, 0.8 * L + 0.4 * d), "y": np.linspace(0.2 * L, 0.8 * L, nr), "x": np.linspace(0.1 * L, 0.9 * L, nr)}
        ageom.append(AGeomss(src, rec))
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.005) * 1e6
    srcwav = make_srcwav(tgrid, ageom, ["vz"], wav)
    kw = dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=["vz", "vx"])
    true = medium.copy()
    a, b = int(0.35 * n), int(0.65 * n)
    true.vp[a:b, a:b, a:b] *= F32(1 + box)
    true.rho[a:b, a:b, a:b] *= F32(1 + box)
    return kw, true


def fwi2d_elastic(nz=44, nx=56, nt=420, nr=16, nss=2, dt=1e-3, fq=12.0, d=10.0, box=0.05, seed=13):
    """2-D elastic FWI gradient case: smooth random model, 'observed' data from the same medium with a
    +5 % box in vp, vs and rho.  Sources :vz, records :vz and :vx (adjoint injection exists only for velocity fields)."""
    from scipy.ndimage import gaussian_filter
    grid = [StepRange(0.0, d, nz), StepRange(0.0, d, nx)]
    rng = np.random.default_rng(seed)
    sm = lambda: gaussian_filter(rng.standard_normal((nz, nx)), 3.0) / 0.08
    vp = (3000.0 * (1 + 0.015 * sm())).astype(F32)
    vs = (1600.0 * (1 + 0.015 * sm())).astype(F32)
    rho = (2300.0 * (1 + 0.015 * sm())).astype(F32)
    medium = Medium(grid, vp, rho, vs)
    tgrid = StepRange(0.0, dt, nt)
    Lz, Lx = grid[0].last, grid[1].last
    ageom = []
    for iss in range(nss):
        sx = (0.25 + 0.5 * iss / max(nss - 1, 1)) * Lx + 0.37 * d
        ageom.append(AGeomss({"z": [0.15 * Lz + 0.21 * d], "x": [sx]}, {"z": np.full(nr, 0.85 * Lz + 0.6 * d), "x": np.linspace(0.05 * Lx, 0.95 * Lx, nr)}))
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.005) * 1e6
    srcwav = make_srcwav(tgrid, ageom, ["vz"], wav)
    kw = dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=["vz", "vx"], pml_faces=["zmin", "zmax", "xmin", "xmax"])
    true = medium.copy()
    z0, z1, x0, x1 = int(0.4 * nz), int(0.6 * nz), int(0.4 * nx), int(0.6 * nx)
    for a in (true.vp, true.vs, true.rho):
        a[z0:z1, x0:x1] *= F32(1 + box)
    return kw, true


def fwi3d_elastic(n=16, nt=150, nr=9, nss=1, dt=1e-3, fq=20.0, d=10.0, box=0.05, seed=17):
    """3-D elastic FWI gradient case: smooth random model, 'observed' data with a +5 % box in vp, vs, rho."""
    from scipy.ndimage import gaussian_filter
    grid = [StepRange(0.0, d, n)] * 3
    rng = np.random.default_rng(seed)
    sm = lambda: gaussian_filter(rng.standard_normal((n, n, n)), 2.0) / 0.05
    vp = (3000.0 * (1 + 0.015 * sm())).astype(F32)
    vs = (1600.0 * (1 + 0.015 * sm())).astype(F32)
    rho = (2300.0 * (1 + 0.015 * sm())).astype(F32)
    medium = Medium(grid, vp, rho, vs)
    tgrid = StepRange(0.0, dt, nt)
    L = grid[0].last
    ageom = []
    for iss in range(nss):
        sx = (0.3 + 0.4 * iss / max(nss - 1, 1)) * L + 0.13 * d
        src = {"z": [0.2 * L + 0.3 * d], "y": [0.45 * L + 0.2 * d], "x": [sx]}
        rec = {"z": np.full(nr, 0.8 * L + 0.4 * d), "y": np.linspace(0.2 * L, 0.8 * L, nr), "x": np.linspace(0.1 * L, 0.9 * L, nr)}
        ageom.append(AGeomss(src, rec))
    wav = _ricker(fq, tgrid, 1.5 / fq + 0.005) * 1e6
    srcwav = make_srcwav(tgrid, ageom, ["vz"], wav)
    kw = dict(medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=["vz", "vx", "vy"])
    true = medium.copy()
    a, b = int(0.35 * n), int(0.65 * n)
    for arr_ in (true.vp, true.vs, true.rho):
        arr_[a:b, a:b, a:b] *= F32(1 + box)
    return kw, true
