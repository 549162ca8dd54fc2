#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors generated FROM THE REFERENCE'S SOURCE TEXT.

    python tests/golden/from_reference.py [--ref /root/reference] [--only NAME]

For every case of tests/golden/refcases.py this script runs the reference's own time loop (`mod_x_proc!`,
src/fdtd/propagate.jl:138-261) with numpy, where every array operation is the reference's kernel text evaluated by
tests/golden/jlref.py:

  parsed from the text                                               hand-written here, citing the host code it follows
  ---------------------------------------------------------------   --------------------------------------------------
  @d_* / @av_* / @all / @inn / @within (diff2D.jl, diff3D.jl)        the order of calls inside the time loop (propagate.jl:170-258)
  compute_dp!/v!/dv!/p!/dstress!/stressii!/stressij! (advance_*.jl)  record! / add_*_source! glue: SpMV `mul!` in CSC order
  free_surface!, free_surface_mirror! (advance_elastic.jl)              (receiver.jl:3-14, source.jl:61-157)
  update_dstress!/update_v!/update_dv!/update_stress!: the sequence   memory*! / boundary_*! drivers: which slab, which offsets
    of kernel calls and their arguments (advance_*.jl)                   (cpml.jl:185-212, boundary.jl:22-52)
  memorynp* statement template (cpml.jl:175-184)                     update_dmod!'s two broadcasts (medium.jl:143-186)
  dirichlet*! statement templates + ghost index lists (dirichlet.jl) save_tp!, boundary snapshots (save_tp.jl, boundary.jl)
  boundary_half*! template (boundary.jl:17-20)
  store_invav*!, muladd_*!, compute_gmod*!, combine_gmodrho! (medium.jl, source.jl, gradient.jl)
  array shapes: get_mgrid methods (src/fields.jl)

Float literals are evaluated in both typings (`f32`: retyped to the kernel number type as ParallelStencil's @parallel does under
@init_parallel_stencil(Threads, Float32, N); `f64`: left as Julia Float64 literals).  Output: tests/golden/ref_<case>_<lit>.npz
with records, strided samples and checksums of every final wavefield and (adjoint case) gradients.  The inputs are rebuilt from
refcases.py by the tests, so the fixtures stay small.  /root/reference is needed to RUN this script, never by the tests.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jlref as J  # noqa: E402
import refcases  # noqa: E402

F32 = np.float32


def call_args(text, head):
    """argument lists (whitespace-normalised strings) of every call `head...)` in text"""
    out, pos = [], 0
    while True:
        k = text.find(head, pos)
        if k < 0:
            return out
        a0 = k + len(head) - 1
        a1 = J.balanced(text, a0)
        out.append([" ".join(a.split()) for a in J.split_top(text[a0 + 1:a1 - 1]) if a.strip()])
        pos = a1


class RefText:
    """everything parsed from the reference for one (ndims, order)"""

    def __init__(self, ref: str, ndims: int, order: int):
        fd = os.path.join(ref, "src", "fdtd")
        rd = lambda f: open(os.path.join(fd, f)).read()
        self.ndims, self.order = ndims, order
        self.macros = J.parse_macros(rd("diff2D.jl" if ndims == 2 else "diff3D.jl"), order, ndims)
        self.kernels = {}
        for f in ("advance_acou.jl", "advance_elastic.jl", "medium.jl", "source.jl", "gradient.jl"):
            for k in J.parse_parallel_kernels(rd(f)):
                self.kernels.setdefault(k.name, []).append(k)
        self.shapes = J.parse_field_shapes(open(os.path.join(ref, "src", "fields.jl")).read(), order)
        self.host = {}
        for f in ("advance_acou.jl", "advance_elastic.jl"):
            txt = J.strip_comments(rd(f))
            for m in re.finditer(r"^function (update_\w+!)\(pap, pac::T\) where \{T<:P_common\{<:(Fdtd\w+),\s*(\d)\}\}\n(.*?)\n^end", txt, re.S | re.M):
                self.host[(m.group(1), m.group(2), int(m.group(3)))] = J.logical_statements(m.group(4))
        g = J.strip_comments(rd("gradient.jl"))
        for m in re.finditer(r"^function (grad\w+!)\(issp, pap, pac::T\) where \{T<:P_common\{<:(Fdtd\w+)(?:,\s*(\d))?\}\}\n(.*?)\n^end", g, re.S | re.M):
            self.host[(m.group(1), m.group(2), int(m.group(3)) if m.group(3) else None)] = J.logical_statements(m.group(4))
        self._cpml(J.strip_comments(rd("cpml.jl")))
        self._dirichlet(J.strip_comments(rd("dirichlet.jl")))
        self._boundary(J.strip_comments(rd("boundary.jl")))
        consts = open(os.path.join(ref, "src", "GeoPhyInv.jl")).read()
        m = re.search(r"const _fd_npml = (\d+) \+ \(_fd_order - 1\)", consts)
        self.npml = int(m.group(1)) + order - 1
        self.nbound = int(re.search(r"const _fd_nbound = (\d+)", consts).group(1))

    def dims(self):
        return ["z", "y", "x"] if self.ndims == 3 else ["z", "x"]

    # ---- cpml.jl:175-184: `memory[ismoff...] = b[imoff] * memory[ismoff...] + a[imoff] * d[isdoff...]` etc.
    def _cpml(self, txt):
        m = re.search(r"function \$fnamenp\(memory::Data\.Array\{\$N\}, d, a, b, kI, moff, doff\)\n(.*?)\n\s*return", txt, re.S)
        tmpl = J.logical_statements(m.group(1))
        assert len(tmpl) == 2 and "$(ismoff...)" in tmpl[0] and "$(isdoff...)" in tmpl[1], tmpl
        # the generating loop's definitions (cpml.jl:166-173) and the two launches of the driver (cpml.jl:185-212)
        assert "ismoff = replace(is, i => :($i + moff))" in txt and "isdoff = replace(is, i => :(doff + $i))" in txt
        assert re.search(r"\[:\(\$i \+ moff\), :\(\$i \+ moff \+ 1\)\]", txt)
        drv = re.search(r"@eval function \$fname\(memory::Data\.Array\{\$N\}, d, a, b, kI, pml_faces\)(.*?)\n            end", txt, re.S).group(1)
        calls = [tuple(a[5:]) for a in call_args(drv, "$fnamenp(") if a[:5] == ["memory", "d", "a", "b", "kI"]]
        assert calls == [("0", "0"), ("_fd_npml", "getindex(size(d), $idim) - _fd_npml")], calls
        assert "setindex!(sm, _fd_npml, $idim)" in drv
        self.memory_kernels = {}
        iv = ["i" + d for d in self.dims()]
        for q, d in enumerate(self.dims()):
            ismoff = ", ".join(f"{v} + moff" if v == "i" + d else v for v in iv)
            isdoff = ", ".join(f"doff + {v}" if v == "i" + d else v for v in iv)
            st = [s.replace("$(ismoff...)", ismoff).replace("$(isdoff...)", isdoff).replace("$imoff", f"i{d} + moff") for s in tmpl]
            self.memory_kernels[d] = (q, J.Kernel(f"memorynp{d}!", ["memory", "d", "a", "b", "kI", "moff", "doff"], None,
                                                  [J.parse_assign(s) for s in st], iv, "\n".join(st)))

    # ---- dirichlet.jl: `$vv[is1...] = 0` for the tangential velocities, `$v[ig1...] = -$v[ig2...]` for each ghost pair
    def _dirichlet(self, txt):
        assert "fdh = div(_fd_order, 2)" in txt
        fdh = self.order // 2
        gm = re.search(r"ighostmin = \[\s*\[replace\(is, i => :\(\$(\w+)\)\), replace\(is, i => :\(\$\((.*?)\)\)\)\] for\s*ifd = 1:fdh", txt, re.S)
        gx = re.search(r"ighostmax = \[\s*\[\s*replace\(is, i => :\(n \+ \$\((.*?)\)\)\),\s*replace\(is, i => :\(n \+ \$\((.*?)\)\)\),\s*\] for ifd = 1:fdh", txt, re.S)
        assert gm and gx, "dirichlet.jl ghost index lists not recognised"
        ev = lambda e, ifd: int(eval(e, {"_fd_order": self.order, "ifd": ifd}))
        ghost_min = [(str(ev(gm.group(1), k)), str(ev(gm.group(2), k))) for k in range(1, fdh + 1)]
        ghost_max = [(f"n + {ev(gx.group(1), k)}", f"n + {ev(gx.group(2), k)}") for k in range(1, fdh + 1)]
        assert "isn = replace(is, i => :n)" in txt and "is1 = replace(is, i => :1)" in txt
        bodies = re.findall(r"function \$fname\(\$v::Data\.Array\{\$N\}, \$\(vrest\.\.\.\), n\)(.*?)return", txt, re.S)
        assert len(bodies) == 2
        want = [("$vv[$(is1...)] = 0", "$v[$(ig[1]...)] = -$v[$(ig[2]...)]", "ighostmin"), ("$vv[$(isn...)] = 0", "$v[$(ig[1]...)] = -$v[$(ig[2]...)]", "ighostmax")]
        for b, (t0, t1, lst) in zip(bodies, want):
            assert t0 in b and t1 in b and f"for ig in {lst}" in b and "for vv in vrest" in b, b
            assert b.index(t0) < b.index(t1)          # tangential zeros first, then the ghost pairs
        self.dirichlet = {}
        dims = self.dims()
        iv = ["i" + d for d in dims]
        for d in dims:
            rest = [x for x in dims if x != d]
            for side, fixed, ghosts in (("min", "1", ghost_min), ("max", "n", ghost_max)):
                st = []
                for vv in rest:
                    st.append(f"v{vv}[{', '.join(fixed if v == 'i' + d else v for v in iv)}] = 0")
                for g1, g2 in ghosts:
                    st.append(f"v{d}[{', '.join(g1 if v == 'i' + d else v for v in iv)}] = -v{d}[{', '.join(g2 if v == 'i' + d else v for v in iv)}]")
                params = [f"v{d}"] + [f"v{x}" for x in rest] + ["n"]
                self.dirichlet[f"dirichlet{d}{side}!"] = J.Kernel(f"dirichlet{d}{side}!", params, None, [J.parse_assign(s) for s in st],
                                                                 [v for v in iv if v != "i" + d], "\n".join(st))

    # ---- boundary.jl:17-52
    def _boundary(self, txt):
        assert re.search(r"function \$fnamehalf\(d::Data\.Array\{\$N\}, b, doff, boff\)\s*d\[\$\(isdoff\.\.\.\)\] = b\[\$\(isboff\.\.\.\)\]", txt)
        assert "isboff = replace(is, i => :($i + boff))" in txt and "isdoff = replace(is, i => :(doff + $i))" in txt
        force = re.search(r"@eval function \$fname\(d::Data\.Array\{\$N\}, b, pml_faces\)(.*?)\n        end", txt, re.S).group(1)
        save = re.search(r"@eval function \$fname\(b::Data\.Array\{\$N\}, d, pml_faces\)(.*?)\n        end", txt, re.S).group(1)
        norm = lambda s: [tuple(a) for a in call_args(s, "$fnamehalf(")]
        assert norm(force) == [("d", "b", "np", "0"), ("d", "b", "getindex(size(d), $idim) - np - _fd_nbound", "_fd_nbound")], norm(force)
        assert norm(save) == [("b", "d", "0", "np"), ("b", "d", "_fd_nbound", "getindex(size(d), $idim) - np - _fd_nbound")], norm(save)
        assert "np = ($(Meta.quot(dimmin)) ∈ pml_faces) ? _fd_npml : 0" in force and "setindex!(sb, _fd_nbound, $idim)" in force
        self.boundary_half = {}
        iv = ["i" + d for d in self.dims()]
        for q, d in enumerate(self.dims()):
            isdoff = ", ".join(f"doff + {v}" if v == "i" + d else v for v in iv)
            isboff = ", ".join(f"{v} + boff" if v == "i" + d else v for v in iv)
            st = f"d[{isdoff}] = b[{isboff}]"
            self.boundary_half[d] = (q, J.Kernel(f"boundary_half{d}!", ["d", "b", "doff", "boff"], None, [J.parse_assign(st)], iv, st))

    def kernel(self, name, nargs, first_ndim):
        c = [k for k in self.kernels[name] if len(k.params) == nargs and (k.nd_annot is None or k.nd_annot == first_ndim)]
        assert len(c) == 1, (name, nargs, first_ndim, [(len(k.params), k.nd_annot) for k in self.kernels[name]])
        return c[0]


class RefSim:
    """`P_common` + `P_x_worker_x_pw` state and the time loop, for one supersource"""

    def __init__(self, T: RefText, case: dict, literals: str):
        self.T, self.case = T, case
        self.ev = J.Evaluator(T.macros, literals)
        self.nd, self.order = T.ndims, T.order
        self.attrib = "FdtdAcoustic" if case["physics"] == "acoustic" else "FdtdElastic"
        self.dims = T.dims()
        n = case["n"]
        self.gn = dict(zip(["m" + d for d in self.dims], n))
        self.npw = case.get("npw", 1)
        fields = [f for (f, a, nd) in T.shapes if a == self.attrib and nd == self.nd]
        self.shape = {f: tuple(J.eval_length(e, self.gn, self.order) for e in T.shapes[(f, self.attrib, self.nd)]) for f in fields}
        self.wavefields = [f for f in fields if not (f.startswith("d") and len(f) > 3)]
        self.dfields = [f for f in fields if f not in self.wavefields]
        z = lambda f: np.zeros(self.shape[f], F32)
        self.pap = []
        for _ in range(self.npw):
            pw = {"w1": {"t": {f: z(f) for f in fields}, "tp": {f: z(f) for f in self.wavefields}},
                  "memory_pml": {}, "velocity_buffer": {f: z(f) for f in self.wavefields if f.startswith("v")},
                  "tauii_buffer": z("p" if self.attrib == "FdtdAcoustic" else "tauxx")}
            for f in self.dfields:          # zeros(::f, ...; pml=true): the axis named by the field's LAST letter has 2 npml entries (fields.jl:75-79)
                sh = list(self.shape[f])
                sh[self.dims.index(f[-1])] = 2 * T.npml
                pw["memory_pml"][f] = np.zeros(sh, F32)
            self.pap.append(pw)
        inputs = refcases.build_inputs(case, self.shape)
        self.inputs = inputs
        self.pac = {"fc": {k: F32(v) for k, v in inputs["fc"].items()}, "ic": {"n" + d: int(v) for d, v in zip(self.dims, n)},
                    "mod": {k: np.asarray(v, F32) for k, v in inputs["mod"].items()}, "dmod": {}, "pml": inputs["pml"],
                    "pml_faces": set(case["pml_faces"]), "rigid_faces": set(case["pml_faces"]) | set(case.get("rigid_faces", [])),
                    "stressfree_faces": set(case.get("stressfree_faces", []))}
        self.pac["ic"]["nt"] = case["nt"]
        self.update_dmod()

    # ---------------------------------------------------------------------------------------------
    # generic interpreter of the reference's host statements: aliases, @parallel calls, memory*!, guards
    # ---------------------------------------------------------------------------------------------
    def resolve(self, s, env):
        s = s.strip()
        if re.fullmatch(r"\d+", s):
            return int(s)
        m = re.fullmatch(r"size\((.*), (\d)\)", s)
        if m:
            return self.resolve(m.group(1), env).shape[int(m.group(2)) - 1]
        m = re.match(r"^(\w+)", s)
        obj = env[m.group(1)]
        rest = s[m.end():]
        while rest:
            m = re.match(r"^\.(\w+)", rest)
            if m:
                obj = obj[m.group(1)]
            else:
                m = re.match(r"^\[:(\w+)\]", rest)
                if m:
                    obj = obj[m.group(1)]
                else:
                    m = re.match(r"^\[(\w+)\]", rest)
                    assert m, (s, rest)
                    k = m.group(1)
                    obj = obj[(int(k) if k.isdigit() else int(env[k])) - 1]        # 1-based
            rest = rest[m.end():]
        return obj

    def ranges(self, s, env):
        out = []
        for r in J.split_top(s.strip()[1:-1]):
            lo, hi = r.split(":", 1)
            out.append((int(lo), int(self.resolve(hi, env))))
        return out

    def call_parallel(self, stmt, env):
        """`@parallel [ranges] name(args...)`"""
        rest = stmt[len("@parallel"):].strip()
        rng = None
        if rest.startswith("("):
            e = J.balanced(rest, 0)
            rng = self.ranges(rest[:e], env)
            rest = rest[e:].strip()
        m = re.match(r"^([\w!]+)\(", rest)
        name = m.group(1)
        args = [self.resolve(a, env) for a in J.split_top(rest[m.end():J.balanced(rest, m.end() - 1) - 1]) if a]
        if name in self.T.dirichlet:
            self.ev.run(self.T.dirichlet[name], args, rng)
            return
        first = args[0]
        self.ev.run(self.T.kernel(name, len(args), first.ndim), args, rng)

    def memory(self, d, memory, darr, a, b, kI, pml_faces):
        """memory{d}! driver, cpml.jl:185-212"""
        q, kern = self.T.memory_kernels[d]
        sm = list(memory.shape)
        sm[q] = self.T.npml
        rng = [(1, x) for x in sm]
        if d + "min" in pml_faces:
            self.ev.run(kern, [memory, darr, a, b, kI, 0, 0], rng)
        if d + "max" in pml_faces:
            self.ev.run(kern, [memory, darr, a, b, kI, self.T.npml, darr.shape[q] - self.T.npml], rng)

    def exec_host(self, stmts, env):
        env = dict(env)
        i, skip = 0, 0
        while i < len(stmts):
            s = stmts[i]
            i += 1
            m = re.fullmatch(r"if \(:(\w+) ∈ (.*)\)", s)
            if m:
                if m.group(1) not in self.resolve(m.group(2), env):
                    while stmts[i] != "end":
                        i += 1
                continue
            if s == "end":
                continue
            m = re.match(r"^\(:(\w+) ∈ (.*?)\) &&\s*(.*)$", s)
            if m:
                if m.group(1) not in self.resolve(m.group(2), env):
                    continue
                s = m.group(3)
            if s.startswith("@parallel"):
                self.call_parallel(s, env)
                continue
            m = re.match(r"^memory(\w)!\((.*)\)$", s, re.S)
            if m:
                self.memory(m.group(1), *[self.resolve(a, env) for a in J.split_top(m.group(2)) if a])
                continue
            m = re.fullmatch(r"(\w+) = (.*)", s)
            if m and "(" not in m.group(2):
                env[m.group(1)] = self.resolve(m.group(2), env)
                continue
            raise SyntaxError(f"host statement not understood: {s!r}")

    def host(self, fname, ipw):
        key = (fname, self.attrib, self.nd)
        self.exec_host(self.T.host[key], {"pap": self.pap[ipw], "pac": self.pac})

    # ---------------------------------------------------------------------------------------------
    # update_dmod! (medium.jl:143-186): @parallel store_* from the text; the two broadcasts restated
    # ---------------------------------------------------------------------------------------------
    def update_dmod(self):
        mod, dmod, dt = self.pac["mod"], self.pac["dmod"], self.pac["fc"]["dt"]
        el = self.attrib == "FdtdElastic"

        def store(kname, dname, src, like):
            dmod[dname] = np.zeros(self.shape[like], F32)
            self.ev.run(self.T.kernel(kname, 3, None), [dmod[dname], mod[src], dt])

        pre = "dtau" if el else "dp"
        like = {"x": ("dtauxxdx" if el else "dpdx"), "y": ("dtauyydy" if el else "dpdy"), "z": ("dtauzzdz" if el else "dpdz")}
        for d in self.dims:
            store(f"store_invav{d}i!", f"dtinvav{d}irho", "rho", like[d])
        if not el:
            # broadcast!(inv, dmod[:dtK], mod[:invK]); rmul!(dmod[:dtK], dt)   (medium.jl:147-148)
            dmod["dtK"] = ((F32(1) / mod["invK"]).astype(F32) * dt).astype(F32)
        else:
            dmod["dtlambda"] = ((F32(1) / mod["invlambda"]).astype(F32) * dt).astype(F32)
            # broadcast!(dmod[:dtM], invλ, invμ) do inv(invλ) + 2.0 * inv(invμ) end: plain Julia, the 2.0 stays Float64 (medium.jl:164-166)
            M = ((F32(1) / mod["invlambda"]).astype(np.float64) + np.float64(2.0) * (F32(1) / mod["invmu"]).astype(np.float64)).astype(F32)
            dmod["dtM"] = (M * dt).astype(F32)
            if self.nd == 2:
                store("store_invav!", "dtavmu", "invmu", "tauxz")
            else:
                store("store_invavxzi!", "dtavxzimu", "invmu", "tauxz")
                store("store_invavxyi!", "dtavxyimu", "invmu", "tauxy")
                store("store_invavyzi!", "dtavyzimu", "invmu", "tauyz")

    # ---------------------------------------------------------------------------------------------
    # sources / receivers (receiver.jl:3-14, source.jl:61-157): SparseArrays mul! in CSC order, Float32 accumulation
    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def spmv(buf, csc, w):
        """mul!(vec(buf), S, w): y = 0; for each column, for each stored entry: y[row] += val * w[col]"""
        colptr, rowval, nzval = csc
        y = buf.reshape(-1, order="F")
        y[...] = 0
        for c in range(len(colptr) - 1):
            for e in range(colptr[c] - 1, colptr[c + 1] - 1):
                y[rowval[e] - 1] = F32(y[rowval[e] - 1] + F32(nzval[e] * F32(w[c])))
        buf[...] = y.reshape(buf.shape, order="F")

    @staticmethod
    def spmv_t(csc, field):
        """mul!(recs, transpose(R), vec(field)): one Float32 accumulator per column, entries in stored order"""
        colptr, rowval, nzval = csc
        x = field.reshape(-1, order="F")
        out = np.zeros(len(colptr) - 1, F32)
        for c in range(len(colptr) - 1):
            acc = F32(0)
            for e in range(colptr[c] - 1, colptr[c + 1] - 1):
                acc = F32(acc + F32(nzval[e] * x[rowval[e] - 1]))
            out[c] = acc
        return out

    def record(self, it, fields, activepw):
        for ipw in activepw:
            for f in fields:
                if f in self.inputs["recv"]:
                    self.records[ipw][f][it - 1, :] = self.spmv_t(self.inputs["recv"][f], self.pap[ipw]["w1"]["t"][f])

    def add_velocity_source(self, it, activepw, src_flags):
        for ipw in activepw:
            if not src_flags[ipw]:
                continue
            for f, wav in self.wavelets[ipw].items():
                if not f.startswith("v"):
                    continue
                pv = self.pap[ipw]["velocity_buffer"][f]
                csc = self.inputs["spray"][f] if ipw == 0 else self.inputs["recv"][f]        # adjoint sources through pw 1's rinterpolatew
                self.spmv(pv, csc, wav[it - 1])
                k = self.T.kernel(f"muladd_with_density_{f}!", 4, None)
                self.ev.run(k, [self.pap[ipw]["w1"]["t"][f], pv, self.pac["mod"]["rho"], self.pac["fc"]["dt"]])

    def add_stress_source(self, it, activepw, src_flags):
        if 0 not in activepw or not src_flags[0]:
            return
        for f, wav in self.wavelets[0].items():
            if f.startswith("v"):
                continue
            pv = self.pap[0]["tauii_buffer"]
            self.spmv(pv, self.inputs["spray"][f], wav[it - 1])
            k = self.T.kernel("muladd_tauii!", 3, None)
            w1t = self.pap[0]["w1"]["t"]
            if self.attrib == "FdtdAcoustic":
                self.ev.run(k, [w1t["p"], pv, self.pac["dmod"]["dtK"]])
            else:
                for t in (["tauxx", "tauyy", "tauzz"] if self.nd == 3 else ["tauxx", "tauzz"]):
                    self.ev.run(k, [w1t[t], pv, self.pac["dmod"]["dtM"]])

    # ---------------------------------------------------------------------------------------------
    # boundary store (boundary.jl:22-52, 113-306)
    # ---------------------------------------------------------------------------------------------
    def bnd_fields(self):
        return ["p"] if self.attrib == "FdtdAcoustic" else ["tauxx", "tauxz", "tauzz"]

    def boundary_half(self, d, dst, src, doff, boff, sb):
        q, kern = self.T.boundary_half[d]
        self.ev.run(kern, [dst, src, doff, boff], [(1, x) for x in sb])

    def boundary_save(self, it):
        nb = self.T.nbound
        for f in self.bnd_fields():
            fld = self.pap[0]["w1"]["t"][f]
            for d in reversed(self.dims):                          # x, (y,) z
                q = self.dims.index(d)
                b = self.boundary[f][d][it - 1]
                npm = self.T.npml if d + "min" in self.pac["pml_faces"] else 0
                sb = list(b.shape)
                sb[q] = nb
                self.boundary_half(d, b, fld, 0, npm, sb)
                self.boundary_half(d, b, fld, nb, fld.shape[q] - npm - nb, sb)
                b *= -F32(1)                                       # rmul!(b, -one(Data.Number))

    def boundary_force(self, it):
        nb = self.T.nbound
        for f in self.bnd_fields():
            fld = self.pap[0]["w1"]["t"][f]
            for d in reversed(self.dims):
                q = self.dims.index(d)
                b = self.boundary[f][d][it - 1]
                npm = self.T.npml if d + "min" in self.pac["pml_faces"] else 0
                sb = list(b.shape)
                sb[q] = nb
                self.boundary_half(d, fld, b, npm, 0, sb)
                self.boundary_half(d, fld, b, fld.shape[q] - npm - nb, nb, sb)

    # ---------------------------------------------------------------------------------------------
    # mod_x_proc! (propagate.jl:138-261)
    # ---------------------------------------------------------------------------------------------
    def run(self, mode, activepw, src_flags, wavelets):
        nt = self.case["nt"]
        self.wavelets = wavelets
        if mode != "adjoint":
            self.records = [{f: np.zeros((nt, len(self.inputs["recv"][f][0]) - 1), F32) for f in self.inputs["recv"]} for _ in range(self.npw)]
        for pw in self.pap:                                          # reset_w2! (types.jl:100-113)
            for grp in (pw["w1"]["t"], pw["w1"]["tp"], pw["memory_pml"], pw["velocity_buffer"]):
                for a in grp.values():
                    a[...] = 0
            pw["tauii_buffer"][...] = 0
        if mode == "forward_save":
            self.boundary = {f: {d: [np.zeros([2 * self.T.nbound if x == d else s for x, s in zip(self.dims, self.shape[f])], F32) for _ in range(nt)]
                                 for d in self.dims} for f in self.bnd_fields()}
            self.snap = {}
        if mode == "adjoint":
            self.gradients = {k: np.zeros(self.shape["p"], F32) for k in ("invK", "rho")}
            w1t = self.pap[0]["w1"]["t"]                             # boundary_force_snap_tau! / _v! (propagate.jl:150-151)
            for f in self.bnd_fields() + [f for f in self.wavefields if f.startswith("v")]:
                w1t[f][...] = self.snap[f]
        for it in range(1, nt + 1):
            self.record(it, [f for f in ("p",) if f in self.wavefields], activepw if mode != "adjoint" else [])
            if mode == "adjoint":
                for ipw in activepw:                                 # save_tp! (save_tp.jl:5-12)
                    for f in self.wavefields:
                        self.pap[ipw]["w1"]["tp"][f][...] = self.pap[ipw]["w1"]["t"][f]
                self.boundary_force(nt - it + 1)
            for ipw in activepw:
                self.host("update_dstress!", ipw)
            for ipw in activepw:
                self.host("update_v!", ipw)
            self.add_velocity_source(it, activepw, src_flags)
            self.record(it, ["vx", "vy", "vz"], activepw if mode != "adjoint" else [])
            for ipw in activepw:
                self.host("update_dv!", ipw)
            for ipw in activepw:
                self.host("update_stress!", ipw)
            self.add_stress_source(it, activepw, src_flags)
            if mode == "forward_save":
                self.boundary_save(it)
            if mode == "adjoint" and self.npw == 2:
                self.compute_gradient()
        if mode == "forward_save":                                   # propagate.jl:251-258
            w1t = self.pap[0]["w1"]["t"]
            for f in self.bnd_fields():
                self.snap[f] = (w1t[f] * -F32(1)).astype(F32)
        if mode != "adjoint" or True:
            self.host("update_dstress!", 0)
            self.host("update_v!", 0)
        if mode == "forward_save":
            for f in self.wavefields:
                if f.startswith("v"):
                    self.snap[f] = self.pap[0]["w1"]["t"][f].copy()

    def compute_gradient(self):
        """compute_gradient!(::Val{:adjoint}, ::Val{2}, ...) = gradlame! + gradrho! (gradient.jl:17-61); 2-D acoustic"""
        grads = {"gradients": self.gradients}
        pap = [dict(pw, ss=[grads]) for pw in self.pap]
        env = {"pap": pap, "pac": self.pac, "issp": 1}
        self.exec_host(self.T.host[("gradlame!", "FdtdAcoustic", None)], env)
        self.exec_host(self.T.host[("gradrho!", "FdtdAcoustic", 2)], env)


def sample(a, stride=None):
    stride = stride or (3 if a.ndim == 2 else 5)
    return np.ascontiguousarray(a[tuple(slice(None, None, stride) for _ in range(a.ndim))])


def checksum(a):
    """order-independent and exact: sum (mod 2^64) and xor of the Float32 bit patterns of EVERY entry.  `+ 0` first: -0.0 and +0.0 are
    the same number (IEEE ==, Julia ==) and count as the same entry -- the reference's full-grid `muladd_tauii!` (source.jl:160-163) adds an
    exact +0.0 to every cell without a source and so turns the -0.0 of `free_surface_mirror!` (tau[1] = -tau[2] over a still-quiet
    surface) into +0.0, where an engine that injects at the source cells only leaves -0.0; no later operation can tell them apart."""
    bits = (np.ascontiguousarray(a, F32) + F32(0)).view(np.uint32).astype(np.uint64).ravel()
    return np.array([np.add.reduce(bits), np.bitwise_xor.reduce(bits)], np.uint64)


def generate(ref, name, case, literals):
    T = RefText(ref, case["ndims"], case["order"])
    sim = RefSim(T, case, literals)
    out = {}
    nt = case["nt"]
    fwd_w = {f: np.asarray(w, F32) for f, w in sim.inputs["wavelets"].items()}
    if case.get("gradient"):
        sim.run("forward_save", [0], [True, False], [fwd_w, {}])
        for f, r in sim.records[0].items():
            out[f"rec_{f}"] = r
        adj = refcases.adjoint_wavelets(case, sim.records[0])
        # get_source(Val{-1}) for pw 1 (source.jl:3-19) is applied by refcases (host side, shared with the test)
        back = refcases.backward_wavelets(fwd_w)
        sim.run("adjoint", [0, 1], [True, True], [back, adj])
        for k, g in sim.gradients.items():
            out[f"grad_{k}"] = g
    else:
        sim.run("forward", [0], [True], [fwd_w])
        for f, r in sim.records[0].items():
            out[f"rec_{f}"] = r
    for ipw in range(sim.npw):
        for f in sim.wavefields:
            a = sim.pap[ipw]["w1"]["t"][f]
            out[f"fld{ipw}_{f}"] = sample(a)
            out[f"sum{ipw}_{f}"] = checksum(a)
    for k, a in sim.pac["dmod"].items():
        out[f"dmod_{k}"] = sample(a)
        out[f"dmodsum_{k}"] = checksum(a)
    out["shapes"] = np.array([f"{f}:{','.join(map(str, s))}" for f, s in sim.shape.items()])
    out["npml"] = np.array([T.npml, T.nbound])
    nz = sum(float(np.abs(v).max()) > 0 for k, v in out.items() if k.startswith("rec_"))
    assert nz > 0, "all records are zero: the case tests nothing"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default=None)
    ap.add_argument("--literals", default="f32,f64")
    args = ap.parse_args()
    srcs = ["src/fdtd/" + f for f in ("diff2D.jl", "diff3D.jl", "advance_acou.jl", "advance_elastic.jl", "cpml.jl", "dirichlet.jl", "boundary.jl",
                                      "medium.jl", "source.jl", "gradient.jl")] + ["src/fields.jl"]
    digest = hashlib.sha256(b"".join(open(os.path.join(args.ref, f), "rb").read() for f in srcs)).hexdigest()
    for name, case in refcases.CASES.items():
        if args.only and args.only != name:
            continue
        for lit in args.literals.split(","):
            out = generate(args.ref, name, case, lit)
            out["reference_sha256"] = np.array(digest)
            path = os.path.join(HERE, f"ref_{name}_{lit}.npz")
            np.savez_compressed(path, **out)
            recs = {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("rec_")}
            print(f"{name} [{lit}]: {os.path.getsize(path) / 1024:.0f} KiB, max |record| {recs}", flush=True)


if __name__ == "__main__":
    main()
