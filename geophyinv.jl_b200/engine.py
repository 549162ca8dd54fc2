"""ctypes binding of libgpifdtd.so -- the C ABI declared in include/gpifdtd.h.

This is the Python twin of julia/GPIFdtdB200.jl: each method is one `ccall` seam of the reference
(SURVEY.md section 8b).  There is NO CPU fallback: importing works anywhere, but constructing an
`Engine` without the built CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPI_LIB") or os.path.join(_HERE, "libgpifdtd.so")   # GPI_LIB: a tuning variant of the same library

ABI_VERSION = 2
ACOUSTIC, ELASTIC = 0, 1
MODE = {"forward": 0, "forward_save": 1, "adjoint": 2}
RUN_BORN = 0x100                     # OR into the mode: FD-Born scattering sources from pw 1 into pw 2
RUN_UNSHIFTED_RHO = 0x200            # OR into an adjoint run: g_rho without upstream's one-cell shift (exact transpose of the Born map)
FACE = {"zmin": 1, "zmax": 2, "ymin": 4, "ymax": 8, "xmin": 16, "xmax": 32}
PARAM = {"invK": 0, "rho": 1, "invlambda": 2, "invmu": 3}
FIELDS = [
    "p", "vx", "vy", "vz", "tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz",
    "dpdx", "dpdy", "dpdz", "dvxdx", "dvydy", "dvzdz",
    "dvxdy", "dvxdz", "dvydx", "dvydz", "dvzdx", "dvzdy",
    "dtauxxdx", "dtauyydy", "dtauzzdz", "dtauxydx", "dtauxydy", "dtauxzdx", "dtauxzdz", "dtauyzdy", "dtauyzdz",
]
FIELD = {name: i for i, name in enumerate(FIELDS)}
NWAVEFIELD = 10
SPRAY, INTERP = 0, 1
RESET_WAVEFIELDS, RESET_RECORDS, RESET_GRADIENTS, RESET_BOUNDARY, RESET_SNAPS = 1, 2, 4, 8, 16


class GpiConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ndims", C.c_int32), ("physics", C.c_int32), ("order", C.c_int32),
        ("n", C.c_int32 * 3), ("nt", C.c_int32), ("npml", C.c_int32), ("nbound", C.c_int32),
        ("pml_faces", C.c_int32), ("rigid_faces", C.c_int32), ("stressfree_faces", C.c_int32),
        ("npw", C.c_int32), ("nshots", C.c_int32), ("store_boundary", C.c_int32),
        ("nsnaps", C.c_int32), ("snaps_field", C.c_int32), ("device", C.c_int32), ("shot_batch", C.c_int32),
        ("slab_rank", C.c_int32), ("slab_nranks", C.c_int32),
        ("dt", C.c_double), ("dtI", C.c_double), ("d", C.c_double * 3), ("dI", C.c_double * 3),
    ]


class GpiTimers(C.Structure):
    _fields_ = [("run_ms", C.c_double), ("steps", C.c_double), ("cell_updates", C.c_double),
                ("stencil_ms", C.c_double), ("launches", C.c_double),
                ("vel_ms", C.c_double), ("vel_n", C.c_double), ("stress_ms", C.c_double), ("stress_n", C.c_double),
                ("exch_ms", C.c_double), ("exch_n", C.c_double), ("allreduce_ms", C.c_double)]


def face_mask(faces) -> int:
    m = 0
    for f in faces:
        f = str(f).lstrip(":")
        if f in FACE:
            m |= FACE[f]
    return m


EXPORTS = [
    "gpi_create", "gpi_destroy", "gpi_last_error", "gpi_abi_version", "gpi_set_medium", "gpi_set_medium_rows", "gpi_set_medium_interior", "gpi_set_medium_fields", "gpi_get_medium", "gpi_slab_range",
    "gpi_update_dmod", "gpi_set_medium_pert", "gpi_update_born", "gpi_set_pml", "gpi_set_sparse", "gpi_set_wavelets", "gpi_run", "gpi_get_records",
    "gpi_get_gradient", "gpi_get_snap", "gpi_set_snap_steps", "gpi_get_field", "gpi_set_field", "gpi_reset",
    "gpi_nccl_unique_id", "gpi_nccl_init", "gpi_allreduce_gradients", "gpi_records_device_ptr",
    "gpi_gradient_device_ptr", "gpi_set_illum", "gpi_get_illum", "gpi_set_stream", "gpi_synchronize", "gpi_get_timers", "gpi_kernel_family", "gpi_field_shape", "gpi_field_shape_order",
]

_lib = None


def load_library(path: str = LIB_PATH):
    """Load libgpifdtd.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with geophyinv.jl_b200/csrc/build.sh (or __graft_entry__.build()). "
            "The engine has no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    # "CUDA or an error": GPI_LIB may name a tuning variant of this library, never the CPU emulation tests/emu builds
    # (it exports gpi_emu_marker); the emulation tests opt in with GPI_TESTS_ALLOW_EMU=1.
    if hasattr(lib, "gpi_emu_marker") and os.environ.get("GPI_TESTS_ALLOW_EMU") != "1":
        raise RuntimeError(f"{path} is the CPU emulation of the engine (test infrastructure); the product loads the CUDA library only")
    fp, ip, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    vp = C.c_void_p
    sig = {
        "gpi_create": ([C.POINTER(GpiConfig), C.POINTER(vp)], C.c_int),
        "gpi_destroy": ([vp], C.c_int),
        "gpi_last_error": ([vp], C.c_char_p),
        "gpi_abi_version": ([], C.c_int),
        "gpi_set_medium": ([vp, C.c_int, fp], C.c_int),
        "gpi_get_medium": ([vp, C.c_int, fp], C.c_int),
        "gpi_set_medium_rows": ([vp, C.c_int, fp, C.c_int, C.c_int], C.c_int),
        "gpi_set_medium_interior": ([vp, C.c_int, fp, ip, ip], C.c_int),
        "gpi_set_medium_fields": ([vp, fp, fp, fp, ip, ip], C.c_int),
        "gpi_slab_range": ([vp, ip, ip], C.c_int),
        "gpi_update_dmod": ([vp], C.c_int),
        "gpi_set_medium_pert": ([vp, C.c_int, fp], C.c_int),
        "gpi_update_born": ([vp], C.c_int),
        "gpi_set_pml": ([vp, C.c_int, fp, fp, fp], C.c_int),
        "gpi_set_sparse": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64p, i64p, fp], C.c_int),
        "gpi_set_wavelets": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "gpi_run": ([vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "gpi_get_records": ([vp, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "gpi_get_gradient": ([vp, C.c_int, fp], C.c_int),
        "gpi_get_snap": ([vp, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "gpi_set_snap_steps": ([vp, C.c_int, ip], C.c_int),
        "gpi_set_illum": ([vp, C.c_int], C.c_int),
        "gpi_get_illum": ([vp, C.POINTER(C.c_double)], C.c_int),
        "gpi_get_field": ([vp, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "gpi_set_field": ([vp, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "gpi_reset": ([vp, C.c_int], C.c_int),
        "gpi_nccl_unique_id": ([vp], C.c_int),
        "gpi_nccl_init": ([vp, vp, C.c_int, C.c_int], C.c_int),
        "gpi_allreduce_gradients": ([vp], C.c_int),
        "gpi_records_device_ptr": ([vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)], C.c_int),
        "gpi_gradient_device_ptr": ([vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)], C.c_int),
        "gpi_set_stream": ([vp, vp], C.c_int),
        "gpi_synchronize": ([vp], C.c_int),
        "gpi_get_timers": ([vp, C.POINTER(GpiTimers)], C.c_int),
        "gpi_field_shape": ([C.c_int, C.c_int, C.c_int, ip, ip], C.c_int),
        "gpi_field_shape_order": ([C.c_int, C.c_int, C.c_int, C.c_int, ip, ip], C.c_int),
        "gpi_kernel_family": ([vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    if lib.gpi_abi_version() != ABI_VERSION:
        raise RuntimeError("libgpifdtd.so ABI version mismatch")
    _lib = lib
    return lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).ravel(order="F"))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class EngineError(RuntimeError):
    pass


class Engine:
    """One handle = one GPU's share of the supersources (the reference's per-worker state,
    src/fdtd/fdtd.jl:340-528).  Method names follow the ABI."""

    prefix = "gpi_"
    dtype = np.float32

    def __init__(self, cfg: GpiConfig):
        self.lib = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        if self.lib.gpi_create(C.byref(cfg), C.byref(self.h)) != 0:
            raise EngineError(self.lib.gpi_last_error(None).decode())

    # -- helpers ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.gpi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.gpi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_shape(self, field: str):
        out = (C.c_int32 * 3)()
        rc = self.lib.gpi_field_shape_order(self.cfg.ndims, self.cfg.physics, self.cfg.order, FIELD[field], self.cfg.n, out)
        if rc != 0:
            raise EngineError(f"field {field} does not exist for this physics / ndims")
        sh = tuple(out)
        return sh if self.cfg.ndims == 3 else (sh[0], sh[2])

    # -- seams -----------------------------------------------------------------------------------
    def set_medium(self, name: str, a):
        a = _f32(a)
        self._ck(self.lib.gpi_set_medium(self.h, PARAM[name], _fp(a)))

    def set_medium_interior(self, name: str, a, lo):
        """`a`: un-extended array [mz,(my),mx]; `lo`: padding cells on the min face of each axis.  The
        replicate padding (media.jl:260-275) runs on the device."""
        a = np.asarray(a)
        shp = a.shape if a.ndim == 3 else (a.shape[0], 1, a.shape[1])
        lo3 = tuple(lo) if len(lo) == 3 else (lo[0], 0, lo[1])
        flat = a.reshape(-1, order="F") if (a.dtype == np.float32 and a.flags.f_contiguous) else _f32(a)
        self._ck(self.lib.gpi_set_medium_interior(self.h, PARAM[name], _fp(flat), (C.c_int32 * 3)(*shp), (C.c_int32 * 3)(*lo3)))

    def set_medium_fields(self, vp, vs, rho, lo):
        """`update!(pa, medium)` in one call: the un-extended vp, vs (None for acoustic media), rho; the derived-parameter
        broadcasts (media.jl:103-130) and the replicate padding run on the device."""
        flat = lambda a: None if a is None else (a.reshape(-1, order="F") if (a.dtype == np.float32 and a.flags.f_contiguous) else _f32(a))
        vp = np.asarray(vp)
        shp = vp.shape if vp.ndim == 3 else (vp.shape[0], 1, vp.shape[1])
        lo3 = tuple(lo) if len(lo) == 3 else (lo[0], 0, lo[1])
        a, b, r = flat(vp), flat(None if vs is None else np.asarray(vs)), flat(np.asarray(rho))
        self._ck(self.lib.gpi_set_medium_fields(self.h, _fp(a), None if b is None else _fp(b), _fp(r), (C.c_int32 * 3)(*shp), (C.c_int32 * 3)(*lo3)))

    def set_medium_rows(self, name: str, rows, k_first: int):
        """rows: [nk, (ny,) nx] = global rows k_first .. k_first+nk-1 of the extended medium array."""
        rows = np.asarray(rows, np.float32)
        a = _f32(rows)
        self._ck(self.lib.gpi_set_medium_rows(self.h, PARAM[name], _fp(a), int(k_first), int(rows.shape[0])))

    def slab_range(self):
        a, b = C.c_int32(), C.c_int32()
        self._ck(self.lib.gpi_slab_range(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_medium(self, name: str):
        shp = self.field_shape("p" if self.cfg.physics == ACOUSTIC else "tauxx")
        out = np.empty(int(np.prod(shp)), np.float32)
        self._ck(self.lib.gpi_get_medium(self.h, PARAM[name], _fp(out)))
        return out.reshape(shp, order="F")

    def update_dmod(self):
        self._ck(self.lib.gpi_update_dmod(self.h))

    def set_medium_pert(self, name: str, a):
        """FD-Born: perturbation of `invK` | `rho` on the extended grid (medium.jl:103-127)."""
        a = _f32(a)
        self._ck(self.lib.gpi_set_medium_pert(self.h, PARAM[name], _fp(a)))

    def update_born(self):
        self._ck(self.lib.gpi_update_born(self.h))

    def set_pml(self, dfield: str, a, b, kI):
        a, b, kI = _f32(a), _f32(b), _f32(kI)
        assert a.size == b.size == kI.size == 2 * self.cfg.npml
        self._ck(self.lib.gpi_set_pml(self.h, FIELD[dfield], _fp(a), _fp(b), _fp(kI)))

    def set_sparse(self, kind: int, ipw: int, issp: int, field: str, colptr, rowval, nzval):
        colptr = np.ascontiguousarray(colptr, np.int64)
        rowval = np.ascontiguousarray(rowval, np.int64)
        nzval = _f32(nzval)
        i64 = C.POINTER(C.c_int64)
        self._ck(self.lib.gpi_set_sparse(self.h, kind, ipw, issp, FIELD[field], colptr.size - 1,
                                         colptr.ctypes.data_as(i64), rowval.ctypes.data_as(i64), _fp(nzval)))

    def set_wavelets(self, ipw: int, issp: int, field: str, w):
        if w is None:
            self._ck(self.lib.gpi_set_wavelets(self.h, ipw, issp, FIELD[field], 0, None))
            return
        w = np.asarray(w, np.float32)
        assert w.ndim == 2 and w.shape[0] == self.cfg.nt, "wavelets must be [nt, ns]"
        wf = _f32(w)
        self._ck(self.lib.gpi_set_wavelets(self.h, ipw, issp, FIELD[field], w.shape[1], _fp(wf)))

    def run(self, mode: str, activepw=(1,), src_flags=(True,), born: bool = False, unshifted_rho: bool = False):
        am = sum(1 << (p - 1) for p in activepw)
        sm = sum(1 << i for i, f in enumerate(src_flags) if f)
        self._ck(self.lib.gpi_run(self.h, MODE[mode] | (RUN_BORN if born else 0) | (RUN_UNSHIFTED_RHO if unshifted_rho else 0), am, sm))

    def get_records(self, ipw: int, issp: int, field: str, nr: int):
        out = np.empty(self.cfg.nt * nr, np.float32)
        self._ck(self.lib.gpi_get_records(self.h, ipw, issp, FIELD[field], _fp(out)))
        return out.reshape((self.cfg.nt, nr), order="F")

    def get_gradient(self, name: str):
        shp = self.field_shape("p" if self.cfg.physics == ACOUSTIC else "tauxx")
        out = np.empty(int(np.prod(shp)), np.float32)
        self._ck(self.lib.gpi_get_gradient(self.h, PARAM[name], _fp(out)))
        return out.reshape(shp, order="F")

    def get_field(self, ipw: int, field: str, ibatch: int = 0):
        shp = self.field_shape(field)
        out = np.empty(int(np.prod(shp)), np.float32)
        self._ck(self.lib.gpi_get_field(self.h, ipw, ibatch, FIELD[field], _fp(out)))
        return out.reshape(shp, order="F")

    def set_field(self, ipw: int, field: str, a, ibatch: int = 0):
        a = _f32(a)
        self._ck(self.lib.gpi_set_field(self.h, ipw, ibatch, FIELD[field], _fp(a)))

    def set_snap_steps(self, its):
        its = np.ascontiguousarray(its, np.int32)
        self._ck(self.lib.gpi_set_snap_steps(self.h, its.size, its.ctypes.data_as(C.POINTER(C.c_int32))))

    def get_snap(self, ipw: int, issp: int, isnap: int):
        shp = self.field_shape(FIELDS[self.cfg.snaps_field])
        out = np.empty(int(np.prod(shp)), np.float32)
        self._ck(self.lib.gpi_get_snap(self.h, ipw, issp, isnap, _fp(out)))
        return out.reshape(shp, order="F")

    def set_illum(self, on: bool):
        self._ck(self.lib.gpi_set_illum(self.h, 1 if on else 0))

    def get_illum(self):
        """Float64, extended grid of :p (fdtd.jl:556-581)."""
        shp = self.field_shape("p")
        out = np.empty(int(np.prod(shp)), np.float64)
        self._ck(self.lib.gpi_get_illum(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape(shp, order="F")

    def reset(self, what: int):
        self._ck(self.lib.gpi_reset(self.h, what))

    def kernel_family(self) -> str:
        """Which stencil kernels `run` launches: 'scalar', 'vec4' (k_*2v / k_*3v), 'tma' (t3::k_step3t), 'order4', or 'vec4-pipelined'
        (z-slab handle: x-split launches of k_*3v with the halo exchange on a side stream)."""
        return {0: "scalar", 1: "vec4", 2: "tma", 3: "vec4-pipelined", 4: "order4"}[self.lib.gpi_kernel_family(self.h)]

    def timers(self) -> dict:
        t = GpiTimers()
        self._ck(self.lib.gpi_get_timers(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in GpiTimers._fields_}

    # -- multi-GPU -------------------------------------------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.gpi_nccl_unique_id(buf) != 0:
            raise EngineError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def nccl_init(self, uid: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self.lib.gpi_nccl_init(self.h, buf, rank, nranks))

    def allreduce_gradients(self):
        self._ck(self.lib.gpi_allreduce_gradients(self.h))

    def gradient_device_ptr(self, name: str):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.gpi_gradient_device_ptr(self.h, PARAM[name], C.byref(p), C.byref(n)))
        return p.value, n.value

    def records_device_ptr(self, ipw: int, issp: int, field: str):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.gpi_records_device_ptr(self.h, ipw, issp, FIELD[field], C.byref(p), C.byref(n)))
        return p.value, n.value

    def synchronize(self):
        self._ck(self.lib.gpi_synchronize(self.h))
