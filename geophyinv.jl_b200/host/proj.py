"""Spray / interpolation matrices (host side), reference src/proj_mat.jl:69-247 and
src/fdtd/ageom.jl:16-58.  The engine consumes them as CSC through `gpi_set_sparse`.

Returned CSC uses Julia conventions: 1-based `colptr` / `rowval`, rows = LinearIndices of the
FIELD's own staggered array (z fastest), stored entries of a column sorted by row (`sparsevec`).
"""
from __future__ import annotations

import numpy as np

from .grids import ORDER, get_mgrid


def get_neighbour_indices(arr: np.ndarray, val: float):
    """proj_mat.jl:69-79 (1-based)."""
    n = arr.size
    idx = int(np.searchsorted(arr, val, side="left")) + 1      # searchsortedfirst
    if idx == 1:
        return 1, 2
    if idx >= n:
        return n - 1, n
    return idx - 1, idx


def _frac(arr, i1, i2, v):
    if v < arr.min():
        return 0.0
    if v > arr.max():
        return 1.0
    return (v - arr[i1 - 1]) / (arr[i2 - 1] - arr[i1 - 1])


def bilinear_interp(grids, point, upstream_3d_swap: bool = True):
    """`bilinear_interp(mmgrid..., P...)` (proj_mat.jl:85-205).  grids/point are in [z,(y),x] order.

    2-D: consistent.  3-D: the reference's signature is (x, y, z, zi, yi, xi) but it is called as
    (mz, my, mx, Pz, Py, Px), so it searches the z grid with the x coordinate and the x grid with
    the z coordinate.  `upstream_3d_swap=True` reproduces that."""
    g = [gr.values for gr in grids]
    nd = len(g)
    if nd == 2:
        (a1, a2), (b1, b2) = get_neighbour_indices(g[0], point[0]), get_neighbour_indices(g[1], point[1])
        da, db = _frac(g[0], a1, a2, point[0]), _frac(g[1], b1, b2, point[1])
        n = g[0].size
        lin = lambda ia, ib: ia + (ib - 1) * n
        idx = [lin(a1, b1), lin(a1, b2), lin(a2, b1), lin(a2, b2)]
        w = [(1 - da) * (1 - db), (1 - da) * db, da * (1 - db), da * db]
    else:
        pz, py, px = point
        cz, cx = (px, pz) if upstream_3d_swap else (pz, px)   # coordinate used on the z grid / x grid
        (a1, a2) = get_neighbour_indices(g[0], cz)
        (b1, b2) = get_neighbour_indices(g[1], py)
        (c1, c2) = get_neighbour_indices(g[2], cx)
        da, db, dc = _frac(g[0], a1, a2, cz), _frac(g[1], b1, b2, py), _frac(g[2], c1, c2, cx)
        n, m = g[0].size, g[1].size
        lin = lambda ia, ib, ic: ia + (ib - 1) * n + (ic - 1) * n * m
        idx = [lin(a1, b1, c1), lin(a1, b1, c2), lin(a1, b2, c1), lin(a1, b2, c2),
               lin(a2, b1, c1), lin(a2, b1, c2), lin(a2, b2, c1), lin(a2, b2, c2)]
        w = [(1 - da) * (1 - db) * (1 - dc), (1 - da) * (1 - db) * dc, (1 - da) * db * (1 - dc), (1 - da) * db * dc,
             da * (1 - db) * (1 - dc), da * (1 - db) * dc, da * db * (1 - dc), da * db * dc]
    idx = np.asarray(idx, np.int64)
    w = np.asarray(w, np.float64).astype(np.float32)           # weights buffer is zeros(number, npt)
    order = np.argsort(idx, kind="stable")                     # sparsevec sorts by index
    return idx[order], w[order]


def _neighbours_vec(arr: np.ndarray, vals: np.ndarray):
    """`get_neighbour_indices` and the clamped fraction of `bilinear_interp` for all points of one axis."""
    n = arr.size
    idx = np.searchsorted(arr, vals, side="left") + 1
    i1 = np.where(idx == 1, 1, np.where(idx >= n, n - 1, idx - 1))
    i2 = i1 + 1
    d = (vals - arr[i1 - 1]) / (arr[i2 - 1] - arr[i1 - 1])
    d = np.where(vals < arr.min(), 0.0, np.where(vals > arr.max(), 1.0, d))
    return i1.astype(np.int64), i2.astype(np.int64), d


def get_proj_matrix(field: str, exgrid, points, upstream_3d_swap: bool = True, order: int = ORDER):
    """`get_proj_matrix(field, attrib_mod, exmedium.grid..., Ps)` (fdtd/ageom.jl:16-22,
    proj_mat.jl:229-247): one column per point.  Returns (colptr, rowval, nzval, nrows).

    All points are processed at once (the reference builds one sparse column per point and `sparse_hcat`s them;
    SURVEY 8f rank 4 names this the host-side bottleneck once the engine is fast): same operations per element as
    `bilinear_interp`, hence the same bits (tests/test_oracle_invariants.py::test_weights_sum_to_one)."""
    grids = get_mgrid(field, exgrid, order)
    g = [gr.values for gr in grids]
    nd = len(g)
    nrows = int(np.prod([len(x) for x in g]))
    P = np.asarray(points, np.float64).reshape(-1, nd)
    npt = P.shape[0]
    if npt == 0:
        return np.ones(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32), nrows
    if nd == 2:
        a1, a2, da = _neighbours_vec(g[0], P[:, 0])
        b1, b2, db = _neighbours_vec(g[1], P[:, 1])
        n = g[0].size
        lin = lambda ia, ib: ia + (ib - 1) * n
        idx = np.stack([lin(a1, b1), lin(a1, b2), lin(a2, b1), lin(a2, b2)], axis=1)
        w = np.stack([(1 - da) * (1 - db), (1 - da) * db, da * (1 - db), da * db], axis=1)
    else:
        cz, cx = (P[:, 2], P[:, 0]) if upstream_3d_swap else (P[:, 0], P[:, 2])
        a1, a2, da = _neighbours_vec(g[0], cz)
        b1, b2, db = _neighbours_vec(g[1], P[:, 1])
        c1, c2, dc = _neighbours_vec(g[2], cx)
        n, m = g[0].size, g[1].size
        lin = lambda ia, ib, ic: ia + (ib - 1) * n + (ic - 1) * n * m
        idx = np.stack([lin(a1, b1, c1), lin(a1, b1, c2), lin(a1, b2, c1), lin(a1, b2, c2),
                        lin(a2, b1, c1), lin(a2, b1, c2), lin(a2, b2, c1), lin(a2, b2, c2)], axis=1)
        w = np.stack([(1 - da) * (1 - db) * (1 - dc), (1 - da) * (1 - db) * dc, (1 - da) * db * (1 - dc), (1 - da) * db * dc,
                      da * (1 - db) * (1 - dc), da * (1 - db) * dc, da * db * (1 - dc), da * db * dc], axis=1)
    o = np.argsort(idx, axis=1, kind="stable")                 # sparsevec sorts by index
    idx = np.take_along_axis(idx, o, axis=1)
    w = np.take_along_axis(w, o, axis=1).astype(np.float32)    # weights buffer is zeros(number, npt)
    k = idx.shape[1]
    colptr = 1 + k * np.arange(npt + 1, dtype=np.int64)
    return colptr, np.ascontiguousarray(idx.ravel()), np.ascontiguousarray(w.ravel()), nrows


def get_proj_matrix_pointwise(field: str, exgrid, points, upstream_3d_swap: bool = True, order: int = ORDER):
    """The same matrix built one point at a time, as the reference does (kept as the check of the batched build)."""
    grids = get_mgrid(field, exgrid, order)
    colptr, rowval, nzval = [1], [], []
    for P in points:
        idx, w = bilinear_interp(grids, P, upstream_3d_swap)
        rowval.extend(idx.tolist()); nzval.extend(w.tolist())
        colptr.append(len(rowval) + 1)
    nrows = int(np.prod([len(g) for g in grids]))
    return np.asarray(colptr, np.int64), np.asarray(rowval, np.int64), np.asarray(nzval, np.float32), nrows
