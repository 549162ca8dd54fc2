"""`SeisForwExpt` / `PFdtd` and the `update!` workflow, backed by the B200 engine.

Python mirror of the reference's L3 host layer (citations relative to /root/reference):
construction src/fdtd/fdtd.jl:43-295, `update!(pa)` src/fdtd/propagate.jl:74-135,
`update!(pa, medium)` src/fdtd/medium.jl:103-186, `update!(pa, srcwav, src_types)`
src/fdtd/source.jl:192-246, `update!(pa, ageom)` src/fdtd/ageom.jl:58-98, `pa[:data]`
src/fdtd/getprop.jl:10-31, `lossvalue` / `gradient!` src/fdtd/func_grad.jl:1-49.

Everything that touched device arrays in the reference is a call into the C ABI
(include/gpifdtd.h) here -- exactly the lines the Julia shim (julia/GPIFdtdB200.jl) replaces with
`ccall`.  Julia `Distributed` workers become ranks: one process per GPU, each owning the
contiguous chunk of supersources `sschunks[rank]` (fdtd.jl:246-255).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .. import engine as E
from .cpml import pml_coefficients
from .data import (AGeomss, Medium, Recs, Srcs, findfreq, findfreq_all, get_adjoint_ageom, get_source, make_recs,
                   pad_widths, padarray, padmgrid)
from .grids import NBOUND, NPML, ORDER, StepRange, dfields_of, dim_names, field_shape, npml_of, wavefields_of
from .proj import get_proj_matrix

F32 = np.float32
ALL_FACES = ["zmin", "zmax", "ymin", "ymax", "xmax", "xmin"]


# --------------------------------------------------------------------------------------------------
# attrib_mod (src/physics_types.jl:17-49)
# --------------------------------------------------------------------------------------------------
@dataclass
class _Fdtd:
    mode: str = "forward"
    npw: int = 1
    born: bool = False

    def __post_init__(self):
        self.mode = str(self.mode).lstrip(":")
        if (self.mode != "forward" or self.born) and self.npw == 1:
            self.npw = 2                       # FdtdAcoustic(:forward_save) / FdtdAcoustic{Born}() have npw = 2 (physics_types.jl:37-48)


class FdtdAcoustic(_Fdtd):
    physics = "acoustic"


class FdtdElastic(_Fdtd):
    physics = "elastic"


def medium_parameters(attrib_mod) -> List[str]:
    """Independent parameters `mod` (src/fdtd/medium.jl:81-95)."""
    return ["invK", "rho"] if attrib_mod.physics == "acoustic" else ["invlambda", "invmu", "rho"]


def sschunks(nss: int, nworker: int):
    """fdtd.jl:251-255 (0-based half-open ranges)."""
    ssi = [int(round(float(s))) for s in np.linspace(0, nss, nworker + 1)]      # round(Int, .): ties to even, like Python
    return [range(ssi[i], ssi[i + 1]) for i in range(nworker)]


def view_inner(a: np.ndarray, npml: int, faces) -> np.ndarray:
    """medium.jl:13-26"""
    fs = {str(f).lstrip(":") for f in faces}
    sl = []
    for d, n in zip(dim_names(a.ndim), a.shape):
        sl.append(slice(npml if d + "min" in fs else 0, n - npml if d + "max" in fs else n))
    return a[tuple(sl)]


class PCommon:
    """`P_common` (src/fdtd/types.jl:132-166): parameters shared by all supersources.

    `exmedium` (fdtd.jl:137) and `mod` (fdtd.jl:139-146) are materialised on the host only when something
    reads them (reference values, CPML bounds, the model-vector path): `update!(pa, medium)` itself sends the
    un-extended arrays to the engine, which pads on the device."""

    _exmedium = None
    _mod = None
    order = ORDER
    npml = NPML

    @property
    def exmedium(self) -> Medium:
        if self._exmedium is None:
            self._exmedium = padarray(self.medium, self.npml, self.pml_faces)
        return self._exmedium

    @property
    def mod(self):
        if self._mod is None:
            self._mod = {name: self.exmedium[name] for name in self.mparams}
        return self._mod


class PFdtd:
    """`PFdtd` (types.jl:181-186).  Build with `SeisForwExpt(attrib_mod; ...)`."""

    def __init__(self, attrib_mod, *, medium: Medium, tgrid: StepRange, ageom, srcwav,
                 pml_faces: Sequence[str] = tuple(ALL_FACES), rigid_faces=None, rfields: Sequence[str] = ("vz",),
                 stressfree_faces: Sequence[str] = ("dummy",), tsnaps=None, snaps_field: Optional[str] = None,
                 verbose: bool = False, nworker: Optional[int] = None, rank: int = 0, device: int = -1,
                 shot_batch: int = 0, upstream_3d_swap: bool = True, zslab=None, order: int = ORDER,
                 jobname: str = "forward_propagation", backprop_flag=None, illum_flag: bool = False):
        """`zslab=(rank, nranks)`: z-slab domain decomposition of ONE experiment over `nranks` GPUs (new
        capability): every rank builds the same experiment, owns one slab of the extended grid and
        exchanges halo planes over NVLink inside `update!`; call `init_nccl` (or `dist.attach_nccl`) first.
        `order` = `_fd_order` (2 or 4), a compile-time preference upstream (src/GeoPhyInv.jl:85-92).
        `jobname`, `backprop_flag` are accepted for call compatibility (fdtd.jl:61-78): upstream stores the first and no longer reads
        the second (the mode of `attrib_mod` replaced it).  `illum_flag` ("flag to output wavefield energy or source illumination; it can
        be used as preconditioner during inversion", fdtd.jl:59): upstream keeps `compute_illum!` / `stack_illums!` (fdtd.jl:556-581) but
        has their calls commented out (propagate.jl:114,236) and allocates a dummy; here the flag does what those functions say --
        `pa.c.illum_stack` (Float64, medium grid) = sum over supersources and time steps of abs2(p) of pw 1, refreshed by every `update!`
        (acoustic experiments)."""
        N = medium.ndims
        npml = npml_of(order)
        if order != 2 and (attrib_mod.born or zslab is not None):
            raise NotImplementedError("FD-Born and z-slabs are implemented for order 2")
        npw = attrib_mod.npw
        assert (attrib_mod.physics == "elastic") == medium.elastic, "attrib_mod / medium mismatch"
        if attrib_mod.born and not (attrib_mod.physics == "acoustic" and N == 2):
            raise NotImplementedError("FD-Born exists upstream for 2-D acoustic media only (src/fdtd/born.jl:1-12)")
        pml_faces = [str(f).lstrip(":") for f in pml_faces]
        rigid_faces = pml_faces if rigid_faces is None else [str(f).lstrip(":") for f in rigid_faces]
        rfields = [str(f).lstrip(":") for f in rfields]

        # ---- normalise ageom / srcwav to npw entries (fdtd.jl:88-102)
        if isinstance(ageom[0], AGeomss):
            if npw == 2:
                adj = get_adjoint_ageom(ageom)
                srcwav = [srcwav, [Srcs(a.ns, tgrid, rfields) for a in adj]]
                ageom = [ageom, adj]
            else:
                ageom, srcwav = [ageom] * npw, [srcwav] * npw
        assert len(ageom) == npw and len(srcwav) == npw
        fields_ok = set(wavefields_of(attrib_mod.physics, N)) | set(dfields_of(attrib_mod.physics, N))
        for rf in rfields:
            assert rf in fields_ok, f"rfield {rf} not a field of {attrib_mod.physics} {N}-D"
        for sw in srcwav:
            for s in sw:
                for f in s.fields:
                    assert f in fields_ok, f"source field {f} not a field of {attrib_mod.physics} {N}-D"
        nss = len(ageom[0])
        assert all(len(a) == nss for a in ageom), "different supersources"
        assert tgrid.last >= srcwav[0][0].grid.last - 1e-12, "modeling time is less than source time"
        for a in ageom:
            for ass in a:
                if not ass.inside(medium.grid):
                    raise ValueError("sources or receivers not inside medium")
        for a, sw in zip(ageom, srcwav):
            assert all(x.ns == y.n for x, y in zip(a, sw)), "ageom and srcwav mismatch"

        c = self.c = PCommon()
        c.order, c.npml = order, npml
        c.attrib_mod, c.medium, c.ageom, c.srcwav = attrib_mod, medium.copy(), [list(a) for a in ageom], [[s.copy() for s in sw] for sw in srcwav]
        c.pml_faces = pml_faces
        c.rigid_faces = list(dict.fromkeys(list(rigid_faces) + pml_faces))          # fdtd.jl:215
        c.stressfree_faces = [str(f).lstrip(":") for f in stressfree_faces]
        c.rfields, c.tgrid, c.verbose = rfields, tgrid, verbose
        c.upstream_3d_swap = upstream_3d_swap
        c.exgrid = padmgrid(medium.grid, npml, pml_faces)                             # grid of exmedium (fdtd.jl:137)
        c.mparams = medium_parameters(attrib_mod)
        c.ref_mod = {name: c.exmedium.ref(name) for name in c.mparams}               # fdtd.jl:175
        n = [len(g) for g in c.exgrid]
        nt = len(tgrid)
        # fc / ic (fdtd.jl:301-332): Float32 copies of Float64 host values
        ds = [g.step for g in c.exgrid]
        dt = tgrid.step
        c.fc = {"dt": F32(dt), "dtI": F32(1.0 / dt)}
        for d, s in zip(dim_names(N), ds):
            c.fc["d" + d] = F32(s)
            c.fc["d" + d + "I"] = F32(1.0 / s) if order == 2 else F32(1.0 / (s * 24.0))     # fdtd.jl:316-319
        c.ic = {**{"n" + d: nn for d, nn in zip(dim_names(N), n)}, "nt": nt, "nsls": 0, "npw": npw}
        c.gradients = {name: np.zeros(n, F32, order="F") for name in c.mparams}      # fdtd.jl:164-170
        c.data = [make_recs(tgrid, ageom[ip], rfields) for ip in range(npw)]         # fdtd.jl:194
        if snaps_field is not None:
            tsn = [0.5 * (tgrid.last + tgrid.first)] if tsnaps is None else list(tsnaps)
            c.itsnaps = [int(np.argmin(np.abs(tgrid.values - t))) + 1 for t in tsn]   # fdtd.jl:181-185
            c.snaps_field = str(snaps_field).lstrip(":")
        else:
            c.itsnaps, c.snaps_field = [], None

        # ---- shots -> workers (fdtd.jl:246-255); one rank = one GPU
        if nworker is None:
            nworker = 1
        nworker = min(nss, nworker)
        self.sschunks = sschunks(nss, nworker)
        self.rank, self.nworker = rank, nworker
        self.local = self.sschunks[rank] if rank < nworker else range(0)
        self._nccl = False
        self.zslab = None
        if zslab is not None and zslab[1] > 1:
            assert nworker == 1 and N == 3 and npw == 1, "z-slabs: one 3-D forward experiment shared by all ranks"
            self.zslab = (int(zslab[0]), int(zslab[1]))
            self.rank, self.local = self.zslab[0], self.sschunks[0]

        # ---- engine (replaces P_x_worker_x_pw / P_x_worker_x_pw_x_ss, fdtd.jl:340-528)
        cfg = E.GpiConfig()
        cfg.abi_version, cfg.ndims, cfg.order = E.ABI_VERSION, N, order
        cfg.physics = E.ACOUSTIC if attrib_mod.physics == "acoustic" else E.ELASTIC
        cfg.n[0], cfg.n[1], cfg.n[2] = n[0], (n[1] if N == 3 else 1), n[-1]
        cfg.nt, cfg.npml, cfg.nbound = nt, npml, NBOUND
        cfg.pml_faces, cfg.rigid_faces = E.face_mask(c.pml_faces), E.face_mask(c.rigid_faces)
        cfg.stressfree_faces = E.face_mask(c.stressfree_faces)
        cfg.npw, cfg.nshots = npw, max(len(self.local), 1)
        cfg.store_boundary = 1 if attrib_mod.mode == "forward_save" else 0            # fdtd.jl:445-455
        cfg.nsnaps = len(c.itsnaps)
        cfg.snaps_field = E.FIELD[c.snaps_field] if c.snaps_field else 0
        cfg.device, cfg.shot_batch = device, shot_batch
        if self.zslab:
            cfg.slab_rank, cfg.slab_nranks = self.zslab
        cfg.dt, cfg.dtI = float(c.fc["dt"]), float(c.fc["dtI"])
        names3 = ["z", "y", "x"]
        for q, d in enumerate(names3):
            cfg.d[q] = float(c.fc.get("d" + d, F32(1)))
            cfg.dI[q] = float(c.fc.get("d" + d + "I", F32(1)))
        self.cfg = cfg
        self.engine = self._make_engine(cfg)
        if c.itsnaps:
            self.engine.set_snap_steps(c.itsnaps)
        c.illum_flag = bool(illum_flag)
        c.illum_stack = np.zeros([len(g) for g in medium.grid], np.float64) if illum_flag else None     # fdtd.jl:177
        if illum_flag:
            if attrib_mod.physics != "acoustic":
                raise NotImplementedError("illum_flag: the illumination is the energy of the pressure field (fdtd.jl:570-581), acoustic experiments only")
            self.engine.set_illum(True)

        self.update_medium(medium)                                                     # fdtd.jl:240
        if attrib_mod.born:                                                            # fdtd.jl:242-244: no perturbation yet
            c.dmod_pert = {name: np.zeros(n, F32, order="F") for name in c.mparams}
            self._upload_born()
        self.update_ageom(c.ageom)                                                     # fdtd.jl:522-525
        self.update_srcwav(c.srcwav, [1] * npw)                                        # fdtd.jl:274
        self.update_pml()                                                              # fdtd.jl:279
        self.stability = check_stability(self, verbose)                                # fdtd.jl:282

    # the only place the backend is chosen: the CUDA library, or nothing
    def _make_engine(self, cfg):
        return E.Engine(cfg)

    # ---------------------------------------------------------------------------------------------
    # update!(pa, medium)  (medium.jl:131-140)
    # ---------------------------------------------------------------------------------------------
    def update_medium(self, medium: Medium):
        c = self.c
        if medium is not c.medium:
            c.medium.copy_from(medium)                                                 # copyto!(pac.medium, medium)
        c._exmedium = c._mod = None                                                    # padarray! happens on the device
        lo, _ = pad_widths(c.medium.ndims, c.npml, c.pml_faces)
        # copyto!(mod[name], exmedium, name) for every name: the getters (media.jl:103-130) are per-cell broadcasts, run on the device
        self.engine.set_medium_fields(c.medium.vp, c.medium.vs, c.medium.rho, lo)
        self.engine.update_dmod()

    # ---------------------------------------------------------------------------------------------
    # update!(pa, medium, medium_pert)  (medium.jl:103-127): FD-Born, waves propagate in `medium`
    # ---------------------------------------------------------------------------------------------
    def update_medium_pert(self, medium: Medium, medium_pert: Medium):
        c = self.c
        assert c.attrib_mod.born, "update!(pa, medium, medium_pert) needs FdtdAcoustic{Born}"
        self.update_medium(medium)
        expert = padarray(medium_pert, c.npml, c.pml_faces)
        c.dmod_pert = {name: np.asfortranarray(expert[name] - c.mod[name]) for name in c.mparams}      # δmod .= exmedium_pert .- mod
        self._upload_born()

    def _upload_born(self):
        for name in self.c.mparams:
            self.engine.set_medium_pert(name, self.c.dmod_pert[name])
        self.engine.update_born()

    # update!(pa, m, mparams): log-parameterised model vector (medium.jl:31-52)
    def update_model(self, m: np.ndarray, mparams=None):
        c = self.c
        mparams = c.mparams if mparams is None else mparams
        chunks = np.split(np.asarray(m, F32), len(mparams))
        for x, name in zip(chunks, mparams):
            inner = view_inner(c.mod[name], c.npml, c.pml_faces)
            inner[...] = (np.exp(x.reshape(inner.shape, order="F")) * c.ref_mod[name]).astype(F32)
            self.engine.set_medium(name, c.mod[name])        # the padding keeps its old values, as in the reference
        self.engine.update_dmod()

    def get_modelvector(self, mparams=None) -> np.ndarray:
        """medium.jl:3-9, 55-76"""
        c = self.c
        mparams = c.mparams if mparams is None else mparams
        out = []
        for name in mparams:
            inner = view_inner(c.mod[name], c.npml, c.pml_faces)
            out.append(np.log(inner * (F32(1) / c.ref_mod[name])).astype(F32).ravel(order="F"))
        return np.concatenate(out)

    # ---------------------------------------------------------------------------------------------
    # update!(pa, ageom[, Srcs|Recs])  (ageom.jl:33-98)
    # ---------------------------------------------------------------------------------------------
    def update_ageom(self, ageom, what: str = "both"):
        c = self.c
        if isinstance(ageom[0], AGeomss):
            ageom = [ageom]
        names = dim_names(c.medium.ndims)
        for ipw in range(c.ic["npw"]):
            c.ageom[ipw] = list(ageom[ipw])
            for issp, iss in enumerate(self.local):
                a = c.ageom[ipw][iss]
                # The reference rebuilds every spray matrix on each update!(pa, srcwav) (source.jl:225); here a matrix
                # whose points have not moved since it was uploaded is left alone (the engine keeps it).
                cache = self.__dict__.setdefault("_sparse_cache", {})
                if what in ("both", "srcs"):
                    pts = np.array([[a.s[d][i] for d in names] for i in range(a.ns)], np.float64).reshape(a.ns, len(names))
                    for sf in c.srcwav[ipw][iss].fields:
                        key = (E.SPRAY, ipw, issp, sf)
                        if key in cache and np.array_equal(cache[key], pts):
                            continue
                        cp, rv, nz, _ = get_proj_matrix(sf, c.exgrid, pts, c.upstream_3d_swap, c.order)
                        self.engine.set_sparse(E.SPRAY, ipw, issp, sf, cp, rv, nz)
                        cache[key] = pts
                if what in ("both", "recs"):
                    pts = np.array([[a.r[d][i] for d in names] for i in range(a.nr)], np.float64).reshape(a.nr, len(names))
                    for rf in c.rfields:
                        key = (E.INTERP, ipw, issp, rf)
                        if key in cache and np.array_equal(cache[key], pts):
                            continue
                        cp, rv, nz, _ = get_proj_matrix(rf, c.exgrid, pts, c.upstream_3d_swap, c.order)
                        self.engine.set_sparse(E.INTERP, ipw, issp, rf, cp, rv, nz)
                        cache[key] = pts

    # ---------------------------------------------------------------------------------------------
    # update!(pa, srcwav, src_types)  (source.jl:192-246)
    # ---------------------------------------------------------------------------------------------
    def update_srcwav(self, srcwav, src_types=None):
        c = self.c
        if isinstance(srcwav[0], Srcs):
            srcwav = [srcwav]
        npw = c.ic["npw"]
        assert len(srcwav) == npw
        src_types = [1] * npw if src_types is None else list(src_types)
        old_fields = [[list(s.fields) for s in sw] for sw in c.srcwav]
        c.srcwav = [[s.copy() for s in sw] for sw in srcwav]
        nt = c.ic["nt"]
        for ipw in range(npw):
            freqmin, freqmax, freqpeaks = 0.0, np.inf, []
            for issp, iss in enumerate(self.local):
                s = c.srcwav[ipw][iss]
                for f in old_fields[ipw][iss]:
                    if f not in s.fields:
                        self.engine.set_wavelets(ipw, issp, f, None)
                peaks = []
                for sf in s.fields:                                         # fill_wavelets! (source.jl:24-58)
                    w = np.zeros((nt, s.n), F32, order="F")
                    w[: len(s.grid), :] = s.d[sf][:nt, :]
                    w = get_source(w, sf, src_types[ipw])
                    self.engine.set_wavelets(ipw, issp, sf, w)
                    # frequency bounds (source.jl:204-216): the reference evaluates them for every pw but keeps those of pw 1 only
                    # (source.jl:229), so the spectra of the other wavefields' wavelets (the adjoint sources: nt x nr per
                    # supersource and gradient) are not computed here; one spectrum serves min, max and peak
                    if ipw == 0 and w.size and w.any():   # !all(isapprox.(w, 0.0)) (source.jl:45): Julia's default atol is 0, so only exactly-zero wavelets are empty
                        fmin, fmax, fpeak = findfreq_all(w, s.grid)
                        freqmax = min(fmax, freqmax)
                        freqmin = max(fmin, freqmin)
                        peaks.append(fpeak)
                if peaks:
                    freqpeaks.append(float(np.mean(peaks)))
            # source fields may have changed => rebuild the spray matrices (source.jl:225)
            self.update_ageom(c.ageom, "srcs")
            if ipw == 0 and freqpeaks:
                c.fc["freqmin"], c.fc["freqmax"] = F32(freqmin), F32(freqmax)
                c.fc["freqpeak"] = F32(np.mean(freqpeaks))
        if "freqpeak" not in c.fc:
            c.fc["freqmin"], c.fc["freqmax"], c.fc["freqpeak"] = F32(0), F32(np.inf), F32(0)

    # update_pml!(pac)  (cpml.jl:144-155)
    def update_pml(self):
        c = self.c
        N = c.medium.ndims
        vb = c.exmedium.bounds("vp")
        velavg = F32((vb[0] + vb[1]) / F32(2))
        c.pml = pml_coefficients(dfields_of(c.attrib_mod.physics, N), c.exgrid, c.medium.grid,
                                 c.pml_faces, float(c.fc["dt"]), float(velavg), float(c.fc["freqpeak"]), c.npml, c.order)
        for df, (a, b, kI) in c.pml.items():
            self.engine.set_pml(df, a, b, kI)

    # ---------------------------------------------------------------------------------------------
    # update!(pa)  (propagate.jl:74-135)
    # ---------------------------------------------------------------------------------------------
    def get_update_parameters(self):
        """propagate.jl:38-60"""
        c = self.c
        if c.attrib_mod.born:                                                          # propagate.jl:53-60
            if c.attrib_mod.mode == "adjoint":
                return dict(activepw=[1, 2], src_flags=[True, True], rec_flags=[False, False])
            return dict(activepw=[1, 2], src_flags=[True, False], rec_flags=[False, True])
        if c.ic["npw"] == 1:
            return dict(activepw=[1], src_flags=[True], rec_flags=[True])
        if c.attrib_mod.mode == "adjoint":
            return dict(activepw=[1, 2], src_flags=[True, True], rec_flags=[False, False])
        if c.attrib_mod.mode in ("forward", "forward_save"):
            return dict(activepw=[1], src_flags=[True, False], rec_flags=[True, False])
        raise ValueError(f"unknown mode {c.attrib_mod.mode}")

    def update(self, upa=None):
        c = self.c
        upa = self.get_update_parameters() if upa is None else upa
        mode = c.attrib_mod.mode
        # initialize!(pa.c), initialize_boundary!, initialize!(localpart)  (propagate.jl:82-94, types.jl:41-176)
        if getattr(self, "_grad_dirty", True):
            for g in c.gradients.values():
                g[...] = 0
            self._grad_dirty = False
        for dat in c.data:
            for d in dat:
                d.fill(0.0)
        what = E.RESET_RECORDS | E.RESET_GRADIENTS | E.RESET_SNAPS | E.RESET_WAVEFIELDS
        if mode == "forward_save":
            what |= E.RESET_BOUNDARY
        self.engine.reset(what)
        if len(self.local):
            # mod_x_proc! (propagate.jl:100-106)
            if c.attrib_mod.born and mode != "adjoint":
                self.engine.run(mode, upa["activepw"], upa["src_flags"], born=True)
            elif getattr(self, "unshifted_rho", False) and mode == "adjoint":
                self.engine.run(mode, upa["activepw"], upa["src_flags"], unshifted_rho=True)
            else:
                self.engine.run(mode, upa["activepw"], upa["src_flags"])
        # sum_grads! (propagate.jl:110-117, gradient.jl:2-11)
        if mode == "adjoint" and c.ic["npw"] == 2 and 2 in upa["activepw"]:
            if self._nccl:
                self.engine.allreduce_gradients()
            elif getattr(self, "nworker", 1) > 1 and not getattr(self, "partial_gradients_ok", False):
                # the reference's sum_grads! always stacks over ALL workers (gradient.jl:2-11, propagate.jl:110-117)
                raise RuntimeError("sharded adjoint run (nworker > 1) without a communicator: call init_nccl / dist.attach_nccl first, "
                                   "or set pa.partial_gradients_ok = True to read this rank's partial gradient on purpose")
            self._grad_dirty = True
            for name in c.mparams:
                c.gradients[name][...] = self.engine.get_gradient(name)
        # stack_illums! (fdtd.jl:556-565; propagate.jl:114): interior view of the stacked energy
        if getattr(c, "illum_flag", False):
            c.illum_stack[...] = view_inner(self.engine.get_illum(), c.npml, c.pml_faces) if len(self.local) else 0.0
        # update_datamat! + update_data! (propagate.jl:119-133, receiver.jl:17-46)
        for ipw in upa["activepw"]:
            if upa["rec_flags"][ipw - 1]:
                for rf in c.rfields:
                    for issp, iss in enumerate(self.local):
                        nr = c.ageom[ipw - 1][iss].nr
                        c.data[ipw - 1][iss].d[rf][...] = self.engine.get_records(ipw - 1, issp, rf, nr)
        t = self.engine.timers()
        # accumulated over the passes of one lossvalue / gradient! call (they reset it)
        self.last_run_ms = getattr(self, "last_run_ms", 0.0) + t.get("run_ms", 0.0)
        self.last_launches = getattr(self, "last_launches", 0.0) + t.get("launches", 0.0)
        return t

    # multi-GPU plumbing: the ncclUniqueId travels through whatever the host already has
    def init_nccl(self, uid: Optional[bytes], nranks: int):
        self.engine.nccl_init(uid, self.rank, nranks)
        self._nccl = True

    # pa[:data, i], pa[:snaps, i] (getprop.jl:10-31)
    def __getitem__(self, key):
        if isinstance(key, tuple):
            what, ipw = key
        else:
            what, ipw = key, 1
        what = str(what).lstrip(":")
        c = self.c
        if what == "data":
            return c.data[ipw - 1]
        if what == "snaps":
            return [[self.engine.get_snap(ipw - 1, issp, k) for k in range(len(c.itsnaps))] for issp, _ in enumerate(self.local)]
        if what == "illum":
            return c.illum_stack
        if what in ("medium", "exmedium", "ageom", "srcwav"):
            return getattr(c, what)
        raise KeyError(key)


def check_stability(pa: "PFdtd", verbose: bool = True, H: Optional[int] = None, epsilon: Optional[float] = None):
    """`check_stability(pa, verbose; H = div(20, _fd_order), epsilon = 1/sqrt(ndims))` (stability.jl:6-58): grid points per minimum
    wavelength and the Courant criterion.  Upstream only warns; this returns what it found (and warns the same way) so that
    callers can read it: {"ds_max", "ds_recommended", "dt", "dt_recommended", "warnings"}."""
    import warnings
    c = pa.c
    H = 20 // c.order if H is None else H
    epsilon = 1.0 / np.sqrt(c.medium.ndims) if epsilon is None else epsilon
    ds = [g.step for g in c.medium.grid]
    dt = float(c.fc["dt"])
    freqmax = float(c.fc.get("freqmax", np.inf))
    if c.attrib_mod.physics == "elastic":
        vs = c.medium.vs[c.medium.vs != 0]
        vmin = float(vs.min()) if vs.size else 0.0                                   # the vs condition overrides
        vmax = float(np.sqrt(float(c.medium.bounds("vp")[1]) ** 2 + float(c.medium.bounds("vs")[1]) ** 2))   # Virieux (1986)
    else:
        vmin, vmax = float(c.medium.bounds("vp")[0]), float(c.medium.bounds("vp")[1])
    out = {"ds_max": max(ds), "dt": dt, "warnings": []}
    ds_temp = round(vmin / H / freqmax, 2) if np.isfinite(freqmax) and freqmax > 0 else np.inf
    out["ds_recommended"] = ds_temp
    if f"{max(ds):0.2e}" != f"{ds_temp:0.2e}" and max(ds) > ds_temp:
        out["warnings"].append(f"decrease maximum spatial sampling ({max(ds):0.2e}) below {ds_temp:0.2e}")
    dt_temp = epsilon * min(ds) / vmax
    out["dt_recommended"] = dt_temp
    if f"{dt:0.2e}" != f"{dt_temp:0.2e}" and dt > dt_temp:
        out["warnings"].append(f"decrease time sampling ({dt:0.2e}) below {dt_temp:0.2e}")
    if verbose:
        for w in out["warnings"]:
            warnings.warn(w)
    return out


def SeisForwExpt(attrib_mod, **kw) -> PFdtd:
    """`SeisForwExpt(attrib_mod; medium, ageom, srcwav, tgrid, ...)` (fdtd.jl:43-47)."""
    return PFdtd(attrib_mod, **kw)


def update(pa: PFdtd, *args, **kw):
    """Dispatch like the reference's `update!` methods."""
    if not args:
        return pa.update(**kw)
    a = args[0]
    if isinstance(a, Medium):
        if len(args) > 1 and isinstance(args[1], Medium):
            return pa.update_medium_pert(a, args[1])
        return pa.update_medium(a)
    if isinstance(a, np.ndarray):
        return pa.update_model(a, *args[1:])
    first = a[0][0] if isinstance(a[0], (list, tuple)) else a[0]
    if isinstance(first, Srcs):
        return pa.update_srcwav(a, *args[1:])
    if isinstance(first, AGeomss):
        what = args[1] if len(args) > 1 else "both"            # update!(pa, ageom, Srcs | Recs) (ageom.jl:58-98)
        what = {"Srcs": "srcs", "Recs": "recs"}.get(getattr(what, "__name__", what), what)
        return pa.update_ageom(a, what)
    raise TypeError("no method update! for these arguments")


# --------------------------------------------------------------------------------------------------
# FWI objective and gradient (src/fdtd/func_grad.jl:1-49, src/database/database.jl:355-379)
# --------------------------------------------------------------------------------------------------
def l2_lossvalue(dobs, data) -> float:
    """`lossvalue(L2DistLoss(), dobs, data)`: sum over shots, fields, samples of (d1 - d2)^2."""
    tot = 0.0
    for a, b in zip(dobs, data):
        for f in a.fields:
            d = np.subtract(a.d[f], b.d[f], dtype=np.float64)           # exact in Float64; no Float64 copies of the records
            tot += float(np.sum(np.square(d, out=d)))
    return tot


def l2_adjoint_source(buffer, dobs, data):
    """`gradient!(buffer, loss, dobs, data)`: g = deriv(L2DistLoss(), dobs, data) element-wise
    (database.jl:369-373).  With LossFunctions 0.11's (output, target) order this is 2*(dobs - data)
    (third-party convention)."""
    for g, a, b in zip(buffer, dobs, data):
        for f in g.fields:
            g.d[f][...] = F32(2) * (a.d[f] - b.d[f])


def lossvalue(m, dobs, pa: PFdtd, mparams=None) -> float:
    """func_grad.jl:1-9"""
    pa.update_model(m, mparams)
    mode_save = pa.c.attrib_mod.mode
    pa.c.attrib_mod.mode = "forward_save"
    pa.update_srcwav(pa.c.srcwav, [1, 0])
    pa.update()
    pa.c.attrib_mod.mode = mode_save
    return l2_lossvalue(dobs, pa.c.data[0])


def gradient(g: np.ndarray, m, dobs, pa: PFdtd, mparams=None) -> float:
    """`gradient!(g, m, loss, dobs, pa, mparams)` (func_grad.jl:11-49).  Returns the loss of the
    forward pass (the reference re-evaluates it on `pa.c.data[1]` after the adjoint pass has
    zeroed that container -- an upstream slip we do not copy)."""
    c = pa.c
    mparams = c.mparams if mparams is None else mparams
    pa.last_run_ms = pa.last_launches = 0.0
    pa.update_model(m, mparams)
    mode_save = c.attrib_mod.mode
    c.attrib_mod.mode = "forward_save"
    pa.update_srcwav(c.srcwav, [1, 0])
    pa.update()
    loss = l2_lossvalue(dobs, c.data[0])
    l2_adjoint_source(c.srcwav[1], dobs, c.data[0])
    for s in c.srcwav[1]:
        s.reverse()
    pa.update_srcwav(c.srcwav, [-1, 1])
    c.attrib_mod.mode = "adjoint"
    pa.update()
    chunks = np.split(np.asarray(m, F32), len(mparams))
    gch = np.split(g, len(mparams))
    for x, gi, name in zip(chunks, gch, mparams):
        r = c.ref_mod[name]
        gm = view_inner(c.gradients[name], c.npml, c.pml_faces)
        gm[...] = gm * np.exp(x.reshape(gm.shape, order="F")) * r      # chain rule (func_grad.jl:36-39)
        gi[...] = gm.ravel(order="F")
    c.attrib_mod.mode = mode_save
    return loss


# --------------------------------------------------------------------------------------------------
# FD-Born linearised map (src/fdtd/func_grad.jl:51-120, commented upstream; test/fwi/born_map.jl)
# --------------------------------------------------------------------------------------------------
def forward_map(d: np.ndarray, m: np.ndarray, pa: PFdtd) -> np.ndarray:
    """`forward_map!(d, m, pa)` (func_grad.jl:53-68): `m` = [δinvK; δrho] on the extended grid (column-major),
    `d` = the scattered data of the first supersource, fields in `rfields` order."""
    c = pa.c
    n = c.gradients[c.mparams[0]].shape
    for x, name in zip(np.split(np.asarray(m, F32), len(c.mparams)), c.mparams):
        c.dmod_pert[name] = np.asfortranarray(x.reshape(n, order="F"))
    pa._upload_born()
    pa.update_srcwav(c.srcwav, [1, 1])
    mode_save = c.attrib_mod.mode
    c.attrib_mod.mode = "forward"
    pa.update()
    c.attrib_mod.mode = mode_save
    d[...] = np.concatenate([c.data[1][0].d[f].ravel(order="F") for f in c.rfields])
    return d


def adjoint_map(gm: np.ndarray, d: np.ndarray, pa: PFdtd, exact: bool = True) -> np.ndarray:
    """`adjoint_map!(gm, d, pa)` (func_grad.jl:70-92): `d` becomes the adjoint source at the receivers of the first
    supersource; the imaging condition of `compute_gradient!` gives `gm` = [g_invK; g_rho].  The forward_save pass
    that fills the boundary store of pw 1 runs here (upstream assumes an earlier call did it).
    `exact=True`: g_rho without upstream's one-cell shift (gradient.jl:53-56) and the sign of a transpose (the
    imaging accumulates MINUS the transpose, the convention of the FWI gradient), so that
    <y, F x> == <x, F' y> (test/fwi/born_map.jl); `exact=False` returns `pa.c.gradients` as the reference would."""
    c = pa.c
    assert c.attrib_mod.physics == "acoustic" and all(f in ("vx", "vz") for f in c.rfields), \
        "adjoint sources exist for velocity receivers only (source.jl:142-156)"
    mode_save = c.attrib_mod.mode
    c.attrib_mod.mode = "forward_save"
    pa.update_srcwav(c.srcwav, [1, 0])
    pa.update()
    s2 = c.srcwav[1][0]
    for x, f in zip(np.split(np.asarray(d, F32), len(c.rfields)), c.rfields):
        s2.d[f][...] = x.reshape(s2.d[f].shape, order="F")
    for s in c.srcwav[1]:
        s.reverse()
    pa.update_srcwav(c.srcwav, [-1, 1])
    c.attrib_mod.mode = "adjoint"
    pa.unshifted_rho = bool(exact)
    try:
        pa.update()
    finally:
        pa.unshifted_rho = False
        c.attrib_mod.mode = mode_save
    gm[...] = np.concatenate([c.gradients[name].ravel(order="F") for name in c.mparams])
    if exact:
        gm *= F32(-1)
    return gm


class LinearMap:
    """`LinearMap(pa)` for `FdtdAcoustic{Born}` (func_grad.jl:106-120): `F @ x` = forward_map, `F.T @ y` = adjoint_map.

    `F` is exactly linear in x = [δinvK; δrho].  `F.T` is its transpose to rounding (dot test 1e-8 in Float64,
    tests/test_oracle_invariants.py) for perturbations supported away from the source and receiver cells: there the
    imaging condition of gradient.jl:17-56 sees the injected wavelets inside `v1 - v1_tp` and the adjoint source
    inside `v2_tp`, which the scattering sources of born.jl do not contain (upstream behaviour, kept)."""

    def __init__(self, pa: PFdtd):
        self.pa = pa
        c = pa.c
        nd = sum(c.data[0][0].d[f].size for f in c.rfields)
        nm = sum(g.size for g in c.gradients.values())
        self.shape = (nd, nm)

    def matvec(self, x):
        return forward_map(np.zeros(self.shape[0], F32), x, self.pa)

    def rmatvec(self, y):
        return adjoint_map(np.zeros(self.shape[1], F32), y, self.pa)

    __matmul__ = matvec

    @property
    def T(self):
        outer = self

        class _T:
            shape = (outer.shape[1], outer.shape[0])

            def __matmul__(self, y):
                return outer.rmatvec(y)
        return _T()
