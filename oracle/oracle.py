"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the CPU restatement (oracle/fdtd_oracle.c).

`OracleEngine` has the method surface of `geophyinv_jl_b200.Engine`, so the SAME host layer
(Medium/AGeom/Srcs -> C-ABI calls) can drive either the CUDA engine (product) or this checker:
`OraclePFdtd` is `PFdtd` with `_make_engine` overridden.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package never
does, and has no code path that could reach it.

Parity status: PINNED to the reference's source text (the reference ships no golden vectors for this path and
Julia is not installed here, so its kernel text is parsed and evaluated with numpy: tests/golden/from_reference.py;
tests/test_reference_pinned.py holds the oracle to those fixtures bit for bit) and to the reference's own test
invariants (tests/test_oracle_invariants.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200 import engine as E  # noqa: E402

_libs = {}


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load(dtype=np.float32, literals="f32"):
    """`literals`: typing of the float literals inside @parallel kernels -- "f32" (ParallelStencil retypes them to the kernel number
    type: the reference's behaviour under its Float32 preference, default) or "f64" (plain Julia promotion); Float64 builds: moot."""
    size = np.dtype(dtype).itemsize
    key = (size, literals if size == 4 else "f64")
    if key in _libs:
        return _libs[key]
    path = os.path.join(_HERE, "liboracle_f64.so" if size == 8 else "liboracle_f32.so" if literals == "f32" else "liboracle_f32_lit64.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "fdtd_oracle.c")):
        build()
    lib = C.CDLL(path)
    assert lib.orc_real_size() == size
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_last_error.argtypes = [C.c_void_p]
    lib.orc_last_run_seconds.restype = C.c_double
    lib.orc_last_run_seconds.argtypes = [C.c_void_p]
    _libs[key] = lib
    return lib


class OracleEngine:
    def __init__(self, cfg: E.GpiConfig, dtype=np.float32, threads: int | None = None, literals: str = "f32"):
        self.dtype = np.dtype(dtype)
        self.lib = load(dtype, literals)
        self.cfg = cfg
        self.h = C.c_void_p()
        if threads:
            self.lib.orc_set_threads(int(threads))
        self.threads = threads or self.lib.orc_max_threads()
        if self.lib.orc_create(C.byref(cfg), C.byref(self.h)) != 0:
            raise RuntimeError(self.lib.orc_last_error(None).decode())

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())

    def _a(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=self.dtype).ravel(order="F"))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def close(self):
        if self.h and self.h.value:
            self.lib.orc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_shape(self, field):
        out = (C.c_int32 * 3)()
        if self.lib.orc_field_shape_order(self.cfg.ndims, self.cfg.order, E.FIELD[field], self.cfg.n, out) != 0:
            raise RuntimeError(f"no field {field}")
        sh = tuple(out)
        return sh if self.cfg.ndims == 3 else (sh[0], sh[2])

    def set_medium(self, name, a):
        a = self._a(a); self._ck(self.lib.orc_set_medium(self.h, E.PARAM[name], self._p(a)))

    def set_medium_interior(self, name, a, lo):
        """Host-side replicate padding (media.jl:260-275) -- the independent check of the engine's device-side pad."""
        a = np.asarray(a)
        n = [self.cfg.n[0], self.cfg.n[2]] if a.ndim == 2 else list(self.cfg.n)
        idx = [np.clip(np.arange(-l, nn - l), 0, m - 1) for l, nn, m in zip(lo, n, a.shape)]
        self.set_medium(name, a[np.ix_(*idx)])

    def set_medium_fields(self, vp, vs, rho, lo):
        """Host-side restatement of the derived-parameter broadcasts (media.jl:103-130: invK / invlambda / invmu, Float32,
        inv(x) = 1/x) -- the independent check of the engine's device-side k_pad_derive."""
        f = np.float32
        vp, rho = np.asarray(vp, f), np.asarray(rho, f)
        if vs is None:
            self.set_medium_interior("invK", f(1) / (vp * vp * rho), lo)
        else:
            vs = np.asarray(vs, f)
            self.set_medium_interior("invlambda", f(1) / ((vp * vp - f(2) * (vs * vs)) * rho), lo)
            self.set_medium_interior("invmu", f(1) / (vs * vs * rho), lo)
        self.set_medium_interior("rho", rho, lo)

    def get_medium(self, name):
        shp = self.field_shape("p" if self.cfg.physics == E.ACOUSTIC else "tauxx")
        out = np.empty(int(np.prod(shp)), self.dtype)
        self._ck(self.lib.orc_get_medium(self.h, E.PARAM[name], self._p(out)))
        return out.reshape(shp, order="F")

    def update_dmod(self):
        self._ck(self.lib.orc_update_dmod(self.h))

    def set_medium_pert(self, name, a):
        a = self._a(a); self._ck(self.lib.orc_set_medium_pert(self.h, E.PARAM[name], self._p(a)))

    def update_born(self):
        self._ck(self.lib.orc_update_born(self.h))

    def set_pml(self, dfield, a, b, kI):
        a, b, kI = self._a(a), self._a(b), self._a(kI)
        self._ck(self.lib.orc_set_pml(self.h, E.FIELD[dfield], self._p(a), self._p(b), self._p(kI)))

    def set_sparse(self, kind, ipw, issp, field, colptr, rowval, nzval):
        colptr = np.ascontiguousarray(colptr, np.int64); rowval = np.ascontiguousarray(rowval, np.int64)
        nzval = self._a(nzval)
        self._ck(self.lib.orc_set_sparse(self.h, kind, ipw, issp, E.FIELD[field], colptr.size - 1,
                                         self._p(colptr), self._p(rowval), self._p(nzval)))

    def set_wavelets(self, ipw, issp, field, w):
        if w is None:
            self._ck(self.lib.orc_set_wavelets(self.h, ipw, issp, E.FIELD[field], 0, None)); return
        w = np.asarray(w)
        wf = self._a(w)
        self._ck(self.lib.orc_set_wavelets(self.h, ipw, issp, E.FIELD[field], w.shape[1], self._p(wf)))

    def run(self, mode, activepw=(1,), src_flags=(True,), born=False, unshifted_rho=False):
        am = sum(1 << (p - 1) for p in activepw)
        sm = sum(1 << i for i, f in enumerate(src_flags) if f)
        self._ck(self.lib.orc_run(self.h, E.MODE[mode] | (E.RUN_BORN if born else 0) | (E.RUN_UNSHIFTED_RHO if unshifted_rho else 0), am, sm))

    def advance(self, it0, nsteps):
        self._ck(self.lib.orc_advance(self.h, int(it0), int(nsteps)))

    def get_records(self, ipw, issp, field, nr):
        out = np.empty(self.cfg.nt * nr, self.dtype)
        self._ck(self.lib.orc_get_records(self.h, ipw, issp, E.FIELD[field], self._p(out)))
        return out.reshape((self.cfg.nt, nr), order="F")

    def get_gradient(self, name):
        shp = self.field_shape("p" if self.cfg.physics == E.ACOUSTIC else "tauxx")
        out = np.empty(int(np.prod(shp)), self.dtype)
        self._ck(self.lib.orc_get_gradient(self.h, E.PARAM[name], self._p(out)))
        return out.reshape(shp, order="F")

    def get_field(self, ipw, field, ibatch=0):
        shp = self.field_shape(field)
        out = np.empty(int(np.prod(shp)), self.dtype)
        self._ck(self.lib.orc_get_field(self.h, ipw, ibatch, E.FIELD[field], self._p(out)))
        return out.reshape(shp, order="F")

    def set_field(self, ipw, field, a, ibatch=0):
        a = self._a(a); self._ck(self.lib.orc_set_field(self.h, ipw, ibatch, E.FIELD[field], self._p(a)))

    def get_dmod(self, idx):
        shape = (C.c_int32 * 3)()
        self._ck(self.lib.orc_get_dmod(self.h, idx, None, shape))
        out = np.empty(int(np.prod(tuple(shape))), self.dtype)
        self._ck(self.lib.orc_get_dmod(self.h, idx, self._p(out), shape))
        sh = tuple(shape)
        return out.reshape(sh if self.cfg.ndims == 3 else (sh[0], sh[2]), order="F")

    def set_snap_steps(self, its):
        its = np.ascontiguousarray(its, np.int32)
        self._ck(self.lib.orc_set_snap_steps(self.h, its.size, self._p(its)))

    def get_snap(self, ipw, issp, isnap):
        shp = self.field_shape(E.FIELDS[self.cfg.snaps_field])
        out = np.empty(int(np.prod(shp)), self.dtype)
        self._ck(self.lib.orc_get_snap(self.h, ipw, issp, isnap, self._p(out)))
        return out.reshape(shp, order="F")

    def set_illum(self, on):
        self._ck(self.lib.orc_set_illum(self.h, 1 if on else 0))

    def get_illum(self):
        shp = self.field_shape("p")
        out = np.empty(int(np.prod(shp)), np.float64)
        self.lib.orc_get_illum.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._ck(self.lib.orc_get_illum(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape(shp, order="F")

    def reset(self, what):
        self._ck(self.lib.orc_reset(self.h, what))

    def timers(self):
        s = self.lib.orc_last_run_seconds(self.h)
        return {"run_ms": s * 1e3, "steps": 0.0, "cell_updates": 0.0, "stencil_ms": 0.0, "launches": 0.0}

    def allreduce_gradients(self):
        pass


class OraclePFdtd(G.PFdtd):
    """The product's host layer driving the CPU restatement instead of the GPU (tests only)."""
    oracle_dtype = np.float32
    oracle_threads = None

    def _make_engine(self, cfg):
        return OracleEngine(cfg, self.oracle_dtype, self.oracle_threads)


class OraclePFdtd64(OraclePFdtd):
    oracle_dtype = np.float64
