"""Inversion-grid layer of `SeisInvExpt` (reference: src/fwi/fwi.jl:11-38, src/fwi/func_grad.jl:1-68, src/proj_mat.jl:18-26,
208-247) -- SURVEY 8f rank 4, the callers on the model side of the hot path.

The optimiser's model vector lives on a coarse inversion grid `migrid`; the engine's medium on the modelling grid `mmgrid`.  Upstream
keeps ONE dense interpolation matrix per axis, P_a = get_proj_matrix([mm_a], [mi_a]) (column j = weights of modelling node j on the
inversion nodes), and applies them separably with Tullio:

    lossvalue(m)  : mfull[i, j(, o)] = sum P[k, i] m[k, l(, n)] Q[l, j] (R[n, o])          (inversion -> modelling grid)
    gradient!(g,m): the same for m, then the modelling-grid gradient, then g = the same contraction with the TRANSPOSED matrices
    get_modelvector: the other way round with P'_a = get_proj_matrix([mi_a], [mm_a])        (modelling grid sampled at the inversion nodes)

`apply_proj_matrix` is that contraction (Float32, one axis after the other: every output of the forward direction has two non-zero
terms per axis, so the order of the axes only moves the last bit).  The wave equation solves in between are `lossvalue` / `gradient`
of host/fdtd.py, i.e. the CUDA engine."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .grids import StepRange
from .proj import get_neighbour_indices, _frac

F32 = np.float32


def proj_matrix_1d(grid: StepRange, points: np.ndarray) -> np.ndarray:
    """`get_proj_matrix([points], [grid])` for one axis (proj_mat.jl:229-247 with the 1-D `bilinear_interp`, proj_mat.jl:85-100): dense
    (len(grid), len(points)), column j = the two linear-interpolation weights of `points[j]` on `grid`; points outside the grid take
    the end node (weight 1), as `_frac` clamps."""
    g = grid.values
    P = np.zeros((g.size, len(points)), F32)
    for j, v in enumerate(np.asarray(points, np.float64)):
        i1, i2 = get_neighbour_indices(g, float(v))
        d = _frac(g, i1, i2, float(v))
        P[i1 - 1, j] += F32(1 - d)
        P[i2 - 1, j] += F32(d)
    return P


def apply_proj_matrix(m: np.ndarray, mats: Sequence[np.ndarray]) -> np.ndarray:
    """m1[i, j(, o)] = sum_k,l(,n) P[k, i] m[k, l(, n)] Q[l, j] (R[n, o])  (proj_mat.jl:18-26), Float32."""
    out = np.asarray(m, F32)
    for axis, P in enumerate(mats):
        out = np.moveaxis(np.tensordot(P.T.astype(F32), out, axes=([1], [axis])), 0, axis).astype(F32)
    return out


class SeisInvExpt:
    """`SeisInvExpt(paf, dobs, migrid, mparams)` (fwi.jl:11-29; `migrid` may be a list of node counts, fwi.jl:32-38)."""

    def __init__(self, paf, dobs, migrid=None, mparams: Optional[List[str]] = None):
        self.paf, self.dobs = paf, dobs
        self.mmgrid = list(paf.c.medium.grid)
        if migrid is None:
            migrid = self.mmgrid
        if all(isinstance(n, (int, np.integer)) for n in migrid):        # fwi.jl:34-36: four cells in from either end
            migrid = [StepRange.from_stop(mm.first + 4 * mm.step, mm.last - 4 * mm.step, int(n)) for mm, n in zip(self.mmgrid, migrid)]
        assert len(migrid) == len(self.mmgrid)
        self.migrid = list(migrid)
        self.mparams = list(paf.c.mparams if mparams is None else mparams)
        # fwi.jl:15-17: modelling nodes interpolated on the inversion grid
        self.P = [proj_matrix_1d(mi, mm.values) for mm, mi in zip(self.mmgrid, self.migrid)]
        self.mfull = paf.get_modelvector(self.mparams)
        self.gmfull = np.zeros_like(self.mfull)

    # shapes (column-major, as the reference reshapes its chunks)
    def _mm_shape(self):
        return tuple(len(g) for g in self.mmgrid)

    def _mi_shape(self):
        return tuple(len(g) for g in self.migrid)

    def _to_modelling_grid(self, m: np.ndarray) -> np.ndarray:
        chunks = np.split(np.asarray(m, F32), len(self.mparams))
        out = [apply_proj_matrix(c.reshape(self._mi_shape(), order="F"), self.P).ravel(order="F") for c in chunks]
        self.mfull[...] = np.concatenate(out)
        return self.mfull

    def get_modelvector(self) -> np.ndarray:
        """func_grad.jl:1-24: the modelling-grid model sampled at the inversion nodes."""
        Pback = [proj_matrix_1d(mm, mi.values) for mm, mi in zip(self.mmgrid, self.migrid)]
        self.mfull[...] = self.paf.get_modelvector(self.mparams)
        chunks = np.split(self.mfull, len(self.mparams))
        return np.concatenate([apply_proj_matrix(c.reshape(self._mm_shape(), order="F"), Pback).ravel(order="F") for c in chunks])

    def lossvalue(self, m: np.ndarray) -> float:
        """func_grad.jl:26-43"""
        from .fdtd import lossvalue
        return lossvalue(self._to_modelling_grid(m), self.dobs, self.paf, self.mparams)

    def gradient(self, g: np.ndarray, m: np.ndarray) -> float:
        """func_grad.jl:45-68: g = P' (gradient on the modelling grid); returns the loss of the forward pass."""
        from .fdtd import gradient
        loss = gradient(self.gmfull, self._to_modelling_grid(m), self.dobs, self.paf, self.mparams)
        gch = np.split(self.gmfull, len(self.mparams))
        Pt = [P.T for P in self.P]
        g[...] = np.concatenate([apply_proj_matrix(c.reshape(self._mm_shape(), order="F"), Pt).ravel(order="F") for c in gch])
        return loss
