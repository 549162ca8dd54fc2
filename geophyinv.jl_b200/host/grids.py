"""Ranges and staggered-grid geometry (host side).

Mirrors Julia's `StepRangeLen` as used for `medium.grid` / `tgrid`, and `get_mgrid`
(reference src/fields.jl:92-671) reduced to per-axis node types, O = order - 1 (orders 2 and 4):

    'I'  tauii / p nodes      offset  0     length n
    'V'  velocity nodes       offset -O/2   length n+O
    'H'  half nodes           offset +O/2   length n-O
    'J'  inner integer nodes  offset +O     length n-2O

The reference fixes the order at compile time (`_fd_order`, a Preferences.jl constant, src/GeoPhyInv.jl:85-92); here
it is a per-experiment keyword (`SeisForwExpt(...; order=4)`), `ORDER` below being the default.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ORDER = 2                      # _fd_order     (src/GeoPhyInv.jl:85)
NPML = 40 + (ORDER - 1)        # _fd_npml = _fd_npextend (src/GeoPhyInv.jl:90-91)
NBOUND = 3                     # _fd_nbound    (src/GeoPhyInv.jl:92)


def npml_of(order: int) -> int:
    """`_fd_npml = _fd_npextend = 40 + (_fd_order - 1)` (src/GeoPhyInv.jl:90-91)."""
    if order not in (2, 4):
        raise NotImplementedError("orders 2 and 4 are implemented (6 and 8 are broken upstream)")
    return 40 + (order - 1)


@dataclass(frozen=True)
class StepRange:
    """`range(start, step=step, length=length)`."""
    start: float
    step: float
    length: int

    @staticmethod
    def from_stop(start: float, stop: float, length: int) -> "StepRange":
        return StepRange(float(start), (float(stop) - float(start)) / (length - 1), int(length))

    @property
    def values(self) -> np.ndarray:
        return self.start + self.step * np.arange(self.length, dtype=np.float64)

    @property
    def first(self) -> float:
        return self.start

    @property
    def last(self) -> float:
        return self.start + self.step * (self.length - 1)

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, i: int) -> float:
        if i < 0:
            i += self.length
        return self.start + self.step * i


# node types per field, (z, y, x)  -- src/fields.jl:92-671
FIELD_TYPES = {
    **{f: "III" for f in ("p", "tauxx", "tauyy", "tauzz", "dvxdx", "dvydy", "dvzdz")},
    "vx": "IIV", "vy": "IVI", "vz": "VII",
    **{f: "JJH" for f in ("dpdx", "dtauxxdx", "dtauxydy", "dtauxzdz")},
    **{f: "JHJ" for f in ("dpdy", "dtauyydy", "dtauxydx", "dtauyzdz")},
    **{f: "HJJ" for f in ("dpdz", "dtauzzdz", "dtauxzdx", "dtauyzdy")},
    **{f: "JHH" for f in ("tauxy", "dvxdy", "dvydx")},
    **{f: "HJH" for f in ("tauxz", "dvxdz", "dvzdx")},
    **{f: "HHJ" for f in ("tauyz", "dvydz", "dvzdy")},
}
_OFFSET = {"I": 0.0, "V": -0.5, "H": 0.5, "J": 1.0}
_DLEN = {"I": 0, "V": 1, "H": -1, "J": -2}


def dim_names(ndims: int):
    return ["z", "x"] if ndims == 2 else ["z", "y", "x"]


def _has_y(field: str) -> bool:
    return "y" in field


def fields_of(physics: str, ndims: int, contains: str = ""):
    """`Fields(attrib_mod, s; ndims)` (src/fields.jl:12-26)."""
    out = []
    for f in FIELD_TYPES:
        if contains not in f:
            continue
        if ndims == 2 and _has_y(f):
            continue
        if physics == "acoustic":
            if "tau" in f or f in ("dvxdy", "dvxdz", "dvydx", "dvydz", "dvzdx", "dvzdy"):
                continue
        else:
            if "p" in f:
                continue
        out.append(f)
    return out


def wavefields_of(physics: str, ndims: int):
    return [f for f in fields_of(physics, ndims) if not f.startswith("d")]


def dfields_of(physics: str, ndims: int):
    return [f for f in fields_of(physics, ndims) if f.startswith("d")]


def get_mgrid(field: str, grids, order: int = ORDER):
    """Grid of `field` given the tauii grids `[mz, (my,) mx]` (src/fields.jl:92-671)."""
    ndims = len(grids)
    t = FIELD_TYPES[field]
    if ndims == 2:
        t = t[0] + t[2]
    out = []
    for ty, m in zip(t, grids):
        out.append(StepRange(m.start + _OFFSET[ty] * m.step * (order - 1), m.step, m.length + _DLEN[ty] * (order - 1)))
    return out


def field_shape(field: str, n, order: int = ORDER):
    ndims = len(n)
    t = FIELD_TYPES[field]
    if ndims == 2:
        t = t[0] + t[2]
    return tuple(int(nn + _DLEN[ty] * (order - 1)) for ty, nn in zip(t, n))


def dfield_axis(field: str, ndims: int) -> int:
    """Index into `dim_names(ndims)` of the axis a derivative field is taken along (cpml.jl:131-132)."""
    return dim_names(ndims).index(field[-1])
