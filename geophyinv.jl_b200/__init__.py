"""geophyinv.jl_b200 -- B200-native drop-in backend for GeoPhyInv.jl's `src/fdtd` hot path.

Layout:
  csrc/            CUDA kernels (sm_100a) + the C ABI of include/gpifdtd.h  -> libgpifdtd.so
  engine.py        ctypes binding of the ABI (Python twin of julia/GPIFdtdB200.jl)
  host/            mirror of the reference's host interface for this path
                   (Medium, AGeom, Srcs/Recs, SeisForwExpt, update!, lossvalue, gradient!, SeisInvExpt's inversion grid)

The directory name contains a dot, so it is imported through the shim `geophyinv_jl_b200.py`
at the repository root (or any importlib spec that names this directory).
"""
from . import engine
from .engine import Engine, EngineError, load_library
from .host.data import (AGeomss, Medium, Recs, Srcs, ageom_xwell, get_source, make_recs, make_srcwav, padarray,
                        ricker)
from .host.fdtd import (FdtdAcoustic, FdtdElastic, LinearMap, PFdtd, SeisForwExpt, adjoint_map, check_stability, forward_map, gradient,
                        l2_adjoint_source, l2_lossvalue, lossvalue, sschunks, update, view_inner)
from .host.grids import NBOUND, NPML, ORDER, StepRange, field_shape, get_mgrid
from .host.inv import SeisInvExpt, apply_proj_matrix, proj_matrix_1d

__all__ = [
    "engine", "Engine", "EngineError", "load_library", "AGeomss", "Medium", "Recs", "Srcs", "ageom_xwell",
    "get_source", "make_recs", "make_srcwav", "padarray", "ricker", "FdtdAcoustic", "FdtdElastic", "PFdtd",
    "SeisForwExpt", "LinearMap", "check_stability", "forward_map", "adjoint_map", "gradient", "l2_adjoint_source", "l2_lossvalue", "lossvalue", "sschunks", "update",
    "view_inner", "SeisInvExpt", "apply_proj_matrix", "proj_matrix_1d", "NBOUND", "NPML", "ORDER", "StepRange", "field_shape", "get_mgrid",
]
