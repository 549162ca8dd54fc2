"""CPML damping profiles (host side, Float64 -> Float32), reference src/fdtd/cpml.jl:8-155.

The engine never recomputes these: they depend on `freqpeak` scanned from the wavelets
(src/fdtd/source.jl:227-237) and are passed verbatim through `gpi_set_pml`.
"""
from __future__ import annotations

import numpy as np

from .grids import NPML, ORDER, StepRange, dfield_axis, dim_names, get_mgrid


def pml_profile(exmgrid: StepRange, mgrid: StepRange, flags, dt: float, velavg: float, freqpeak: float, npml: int = NPML):
    """`update_pml!(pml, exmgrid, mgrid, flags, dt, velavg, freqpeak)` (cpml.jl:8-103).
    Returns (a, b, kI), each of length 2*npml: first npml entries = min face, last npml = max face."""
    # dt, velavg and freqpeak reach this function as `Data.Number` (Float32) values (cpml.jl:144-155);
    # `pi * freqpeak` with a Float32 argument is a Float32 product in Julia.
    dt = float(np.float32(dt)); velavg = float(np.float32(velavg))
    x = exmgrid.values
    nx = x.size
    xoriginleft = mgrid.first - 2 * mgrid.step
    xoriginright = mgrid.last + 2 * mgrid.step
    NPOWER, K_MAX_PML = 2.0, 1.0
    ALPHA_MAX_PML = float(np.float32(np.pi) * np.float32(freqpeak))
    thickness = (npml - 1) * mgrid.step
    Rcoef = 0.001
    d0 = -(NPOWER + 1) * velavg * np.log(Rcoef) / (2.0 * thickness)

    k = np.ones(nx); d = np.zeros(nx); alpha = np.zeros(nx); a = np.zeros(nx); b = np.zeros(nx)
    for ix in range(nx):
        if flags[0]:
            ab = xoriginleft - x[ix]
            if ab >= 0.0:
                an = ab / thickness
                d[ix] = d0 * an ** NPOWER
                k[ix] = 1.0 + (K_MAX_PML - 1.0) * an ** NPOWER
                alpha[ix] = ALPHA_MAX_PML * (1.0 - an) + 0.01 * ALPHA_MAX_PML
        if flags[1]:
            ab = x[ix] - xoriginright
            if ab >= 0.0:
                an = ab / thickness
                d[ix] = d0 * an ** NPOWER
                k[ix] = 1.0 + (K_MAX_PML - 1.0) * an ** NPOWER
                alpha[ix] = ALPHA_MAX_PML * (1.0 - an) + 0.01 * ALPHA_MAX_PML
        if alpha[ix] < 0.0:
            alpha[ix] = 0.0
        b[ix] = np.exp(-(d[ix] / k[ix] + alpha[ix]) * dt)
        if abs(d[ix]) > 1.0e-6:
            a[ix] = d[ix] * (b[ix] - 1.0) / (k[ix] * (d[ix] + k[ix] * alpha[ix]))
    if nx < npml:
        # the reference slices the first and last npml entries of every axis whether or not it has a CPML face (cpml.jl:100-105):
        # a BoundsError there, a message here
        raise ValueError(f"an extended grid axis has {nx} nodes, fewer than npml = {npml}: update_pml! (cpml.jl:100-105) cannot take its slabs")
    pkI = np.concatenate([1.0 / k[:npml], 1.0 / k[nx - npml:]])
    pa = np.concatenate([a[:npml], a[nx - npml:]])
    pb = np.concatenate([b[:npml], b[nx - npml:]])
    return pa.astype(np.float32), pb.astype(np.float32), pkI.astype(np.float32)


def pml_coefficients(dfields, exgrid, mgrid, pml_faces, dt: float, velavg: float, freqpeak: float, npml: int = NPML, order: int = ORDER):
    """Loop over the derivative fields (cpml.jl:125-142): each uses ITS OWN staggered 1-D grid along
    its last letter.  Returns {dfield: (a, b, kI)}."""
    nd = len(exgrid)
    faces = {str(f).lstrip(":") for f in pml_faces}
    out = {}
    for df in dfields:
        i = dfield_axis(df, nd)
        dim = dim_names(nd)[i]
        exg = get_mgrid(df, exgrid, order)[i]
        flags = [dim + "min" in faces, dim + "max" in faces]
        out[df] = pml_profile(exg, mgrid[i], flags, dt, velavg, freqpeak, npml)
    return out
