"""Host data model that stays unchanged across the engine swap: Medium, AGeom, Srcs/Recs, wavelets.

Python restatement of the parts of the reference's L0 layer that the FDTD experiment consumes
(citations relative to /root/reference).  Arrays keep the reference's index order [z,(y),x] and are
Float32 (`Data.Number`, src/GeoPhyInv.jl:95-100); numpy arrays are stored Fortran-ordered so that
`a[iz, ix]` and a column-major flat view both mean what they mean in Julia.
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field
from typing import Dict, List, Sequence

import numpy as np

from .grids import NPML, StepRange, dim_names

F32 = np.float32


def _farr(a, dtype=F32):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


# --------------------------------------------------------------------------------------------------
# Medium (src/media/media.jl:18-132)
# --------------------------------------------------------------------------------------------------
class Medium:
    """AcousticMedium (vp, rho) or ElasticMedium (vp, vs, rho) on `grid = [mz, (my,) mx]`."""

    def __init__(self, grid: Sequence[StepRange], vp, rho, vs=None, validate: bool = True):
        self.grid = list(grid)
        shp = tuple(len(g) for g in self.grid)
        self.vp, self.rho = _farr(vp), _farr(rho)
        self.vs = None if vs is None else _farr(vs)
        for a in (self.vp, self.rho) + (() if self.vs is None else (self.vs,)):
            assert a.shape == shp, f"medium array shape {a.shape} != grid {shp}"
            if validate:
                assert np.all(a >= 0), "negative medium parameters"

    @property
    def elastic(self) -> bool:
        return self.vs is not None

    @property
    def ndims(self) -> int:
        return len(self.grid)

    @staticmethod
    def homogeneous(grid, vp=2500.0, rho=2500.0, vs=None):
        shp = tuple(len(g) for g in grid)
        return Medium(grid, np.full(shp, vp, F32), np.full(shp, rho, F32), None if vs is None else np.full(shp, vs, F32))

    # derived parameters, all evaluated in Float32 (media.jl:103-130)
    def __getitem__(self, name: str) -> np.ndarray:
        vp, rho, vs = self.vp, self.rho, self.vs
        one = F32(1)
        if name in ("vp", "rho", "vs"):
            return getattr(self, name)
        if not self.elastic:
            K = vp * vp * rho
            table = {"K": K, "invK": one / K, "lambda": K, "invlambda": one / K, "M": K, "invrho": one / rho}
        else:
            mu = vs * vs * rho
            lam = (vp * vp - F32(2) * (vs * vs)) * rho
            table = {"mu": mu, "invmu": one / mu, "lambda": lam, "invlambda": one / lam, "M": vp * vp * rho,
                     "K": (vp * vp - F32(4) / F32(3) * (vs * vs)) * rho, "invrho": one / rho}
            table["invK"] = one / table["K"]
        return _farr(table[name])

    def ref(self, name: str) -> F32:
        """`mean(m)` (media.jl:31)."""
        return F32(np.mean(self[name], dtype=np.float32))

    def bounds(self, name: str, frac: float = 0.1):
        """media.jl:24-26"""
        m = self[name]
        r = self.ref(name)
        return [max(F32(0), F32(m.min() - F32(frac) * r)), F32(m.max() + F32(frac) * r)]

    def copy(self) -> "Medium":
        return Medium(self.grid, self.vp.copy(order="F"), self.rho.copy(order="F"),
                      None if self.vs is None else self.vs.copy(order="F"), validate=False)

    def copy_from(self, other: "Medium") -> "Medium":
        """`copyto!(pac.medium, medium)` (medium.jl:133): in place, no allocation."""
        assert self.elastic == other.elastic and self.vp.shape == other.vp.shape, "medium shape / physics mismatch"
        self.grid = list(other.grid)
        np.copyto(self.vp, other.vp); np.copyto(self.rho, other.rho)
        if self.elastic:
            np.copyto(self.vs, other.vs)
        return self


def _face_flags(faces, ndims):
    names = dim_names(ndims)
    fs = {str(f).lstrip(":") for f in faces}
    return [(d + "min") in fs for d in names], [(d + "max") in fs for d in names]


def padmgrid(grid, npml, faces):
    """media.jl:278-296"""
    fmin, fmax = _face_flags(faces, len(grid))
    out = []
    for g, a, b in zip(grid, fmin, fmax):
        lo, hi = (npml if a else 0), (npml if b else 0)
        out.append(StepRange(g.start - lo * g.step, g.step, g.length + lo + hi))
    return out


def padarray(medium: Medium, npml: int = NPML, faces=("zmin", "zmax", "ymin", "ymax", "xmin", "xmax")) -> Medium:
    """Replicate-pad the medium into the PML (media.jl:260-304): `Pad(:replicate)` on every axis,
    then a view that keeps the padding only on PML faces."""
    fmin, fmax = _face_flags(faces, medium.ndims)

    def pad(a):
        idx = []
        for ax, (n, a_min, a_max) in enumerate(zip(a.shape, fmin, fmax)):
            lo, hi = (npml if a_min else 0), (npml if a_max else 0)
            idx.append(np.clip(np.arange(-lo, n + hi), 0, n - 1))
        return _farr(a[np.ix_(*idx)])

    return Medium(padmgrid(medium.grid, npml, faces), pad(medium.vp), pad(medium.rho), None if medium.vs is None else pad(medium.vs), validate=False)


def pad_widths(ndims: int, npml: int, faces):
    """cells of replicate padding on the (min, max) face of each axis"""
    fmin, fmax = _face_flags(faces, ndims)
    return [npml if a else 0 for a in fmin], [npml if b else 0 for b in fmax]


# --------------------------------------------------------------------------------------------------
# acquisition geometry (src/ageom/ageom.jl:4-67)
# --------------------------------------------------------------------------------------------------
@dataclass
class AGeomss:
    """One supersource: `ns` simultaneous sources and `nr` receivers; coordinates keyed 'z','y','x'."""
    s: Dict[str, np.ndarray]
    r: Dict[str, np.ndarray]

    def __post_init__(self):
        self.s = {k: np.asarray(v, np.float64) for k, v in self.s.items()}
        self.r = {k: np.asarray(v, np.float64) for k, v in self.r.items()}
        assert list(self.s) == list(self.r), "AGeomss construct"
        self.ns = len(next(iter(self.s.values())))
        self.nr = len(next(iter(self.r.values())))
        assert all(len(v) == self.ns for v in self.s.values()) and all(len(v) == self.nr for v in self.r.values())

    def inside(self, grid) -> bool:
        names = dim_names(len(grid))
        ok = True
        for d, g in zip(names, grid):
            lo, hi = min(g.first, g.last), max(g.first, g.last)
            for c in (self.s[d], self.r[d]):
                ok &= bool(np.all((c >= lo) & (c <= hi)))
        return ok


AGeom = list  # Vector{AGeomss}


def get_adjoint_ageom(ageom):
    """Receivers become simultaneous sources (src/fdtd/ageom.jl:10-14)."""
    return [AGeomss(a.r, a.r) for a in ageom]


def ageom_xwell(grid, nss: int = 1, nr: int = 10, sp=((0.9, 0.9), (0.1, 0.9)), rp=((0.9, 0.1), (0.1, 0.1))):
    """`AGeom(mgrid, :xwell, SSrcs(nss), Recs(nr))` (src/ageom/gallery.jl:64-115): one source per
    supersource spread between sp[0] and sp[1], receivers on a line between rp[0] and rp[1];
    fractions are the `getp` weights a*m[1] + (1-a)*m[end]; 3-D uses the middle of y."""
    nd = len(grid)
    names = dim_names(nd)

    def getp(frac):
        fr = [frac[0], 0.5, frac[1]] if nd == 3 else list(frac)
        return [a * g.first + (1 - a) * g.last for a, g in zip(fr, grid)]

    s0, s1, r0, r1 = getp(sp[0]), getp(sp[1]), getp(rp[0]), getp(rp[1])
    out = []
    for iss in range(nss):
        w = 0.0 if nss == 1 else iss / (nss - 1)
        s = {d: np.array([s0[i] + w * (s1[i] - s0[i])]) for i, d in enumerate(names)}
        r = {d: np.linspace(r0[i], r1[i], nr) for i, d in enumerate(names)}
        out.append(AGeomss(s, r))
    return out


# --------------------------------------------------------------------------------------------------
# Srcs / Recs (src/database/database.jl:19-43): per supersource, field -> (nt x n) Float32 matrix
# --------------------------------------------------------------------------------------------------
class Records:
    def __init__(self, n: int, grid: StepRange, fields: Sequence[str], dtype=F32):
        self.n, self.grid, self.fields = int(n), grid, list(fields)
        self.d = {f: np.zeros((len(grid), self.n), dtype, order="F") for f in self.fields}

    def __getitem__(self, key):
        return self.n if key in ("n", "ns", "nr") else self.d[key]

    def copy(self):
        out = type(self)(self.n, self.grid, self.fields)
        for f in self.fields:
            out.d[f][...] = self.d[f]
        return out

    def fill(self, v=0.0):
        for a in self.d.values():
            a[...] = v

    def reverse(self):
        """`reverse!` along time (database.jl:180-190)."""
        for f in self.fields:
            self.d[f][...] = self.d[f][::-1, :]

    def update(self, fields, wav):
        """`update!(srcwav, fields, wav)`: the same wavelet for every source (database.jl)."""
        w = np.asarray(wav, np.float64)
        for f in fields:
            if f not in self.d:
                self.fields.append(f)
                self.d[f] = np.zeros((len(self.grid), self.n), F32, order="F")
            self.d[f][...] = w[:, None].astype(F32)


class Srcs(Records):
    pass


class Recs(Records):
    pass


def make_srcwav(tgrid: StepRange, ageom, fields, wav=None) -> List[Srcs]:
    """`Srcs(tgrid, ageom, fields)` + `update!(srcwav, fields, wav)`: Vector{Srcs} over supersources."""
    out = [Srcs(a.ns, tgrid, fields) for a in ageom]
    if wav is not None:
        for s in out:
            s.update(fields, wav)
    return out


def make_recs(tgrid: StepRange, ageom, fields) -> List[Recs]:
    return [Recs(a.nr, tgrid, fields) for a in ageom]


# --------------------------------------------------------------------------------------------------
# wavelets (src/srcwav/wavelets.jl:14-72) and the source transform (src/fdtd/source.jl:3-19)
# --------------------------------------------------------------------------------------------------
def ricker(fqdom: float, tgrid: StepRange, tpeak: float | None = None, maxamp: float = 1.0) -> np.ndarray:
    t = tgrid.values
    if tpeak is None:
        tpeak = tgrid.first + 1.5 / fqdom
    if tpeak < tgrid.first + 1.5 / fqdom or tpeak > tgrid.last - 1.5 / fqdom:
        raise ValueError("cannot output Ricker for given tgrid and tpeak")
    pf = (np.pi * np.pi) * fqdom ** 2.0
    tsq = (t - tpeak) * (t - tpeak)
    return (1.0 - 2.0 * pf * tsq) * np.exp(-1.0 * pf * tsq) * maxamp


def get_source(w: np.ndarray, field: str, src_type: int) -> np.ndarray:
    """source.jl:3-19.  w is (nt, ns) Float32."""
    w = np.asarray(w, F32)
    if src_type == 0:
        return np.zeros_like(w)
    if src_type == 1:
        return w.copy()
    if src_type == -1:
        if field not in ("vx", "vy", "vz"):
            raise TypeError(f"no method get_source(w, ::{field}, ::Val{{-1}}) (source.jl:8)")
        ww = -w                                    # rmul!(ww, -1)
        ww = np.roll(ww, -1, axis=0)               # circshift(ww, (-1, 0))
        ww = ww[::-1, :].copy()                    # reverse!(ww, dims=1)
        ww[0, :] = 0
        return ww
    raise TypeError(f"no method get_source for src_type {src_type}")


def findfreq_all(x: np.ndarray, tgrid: StepRange, threshold: float = -50.0):
    """(min, max, peak) of src/Utils/freq.jl:28-57 from ONE spectrum (the reference transforms the wavelets once per attribute)."""
    x = np.asarray(x, np.float64)
    cx = np.fft.rfft(x, axis=0)
    fgrid = np.fft.rfftfreq(len(tgrid), tgrid.step)
    ax = np.abs(cx) ** 2
    if ax.max() == 0.0:
        return 0.0, 0.0, 0.0
    ax = 10.0 * np.log10(np.maximum(ax / ax.max(), 1e-300))
    if ax.ndim == 2:
        ipeak = np.unravel_index(np.argmax(ax.T), ax.T.shape)[::-1][0]
        cols = np.nonzero(ax >= threshold)
        order = np.lexsort((cols[0], cols[1]))          # findfirst / findlast run in column-major order
        ifirst, ilast = cols[0][order[0]], cols[0][order[-1]]
    else:
        ipeak = int(np.argmax(ax))
        rows = np.nonzero(ax >= threshold)[0]
        ifirst, ilast = rows[0], rows[-1]
    return float(fgrid[ifirst]), float(fgrid[ilast]), float(fgrid[ipeak])


def findfreq(x: np.ndarray, tgrid: StepRange, attrib: str = "peak", threshold: float = -50.0) -> float:
    """src/Utils/freq.jl:28-57"""
    fmin, fmax, fpeak = findfreq_all(x, tgrid, threshold)
    return {"min": fmin, "max": fmax, "peak": fpeak}[attrib]
