#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 FDTD engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5|c3small|c5small]

Workload at N=1 (default, `c3`): BASELINE config 3 -- 3-D isotropic elastic 256^3 (+41-cell CPML on six
faces = 338^3 extended cells), one :vz source, 64 receivers, 2000 time steps.  One bench "step" is one
pass of the hot path (`mod_x_proc!`, reference src/fdtd/propagate.jl:138-261) over one supersource,
i.e. 2000 time steps.  With N ranks (torchrun, one rank per GPU) every rank propagates its own
supersource of the same shape (weak scaling: "one supersource per GPU", no data-path collective).

value     = extended-grid cell updates per second of the whole job, inputs resident in HBM, timed
            with CUDA events on the engine's stream inside gpi_run, max over ranks.
e2e       = same metric through the host API a user calls (`update!(pa, medium)`; `update!(pa)`;
            records to host) with HOST buffers: medium H2D, dmod rebuild, run, records D2H all timed.
roofline  = the dominant kernel (fused stress kernel) timed live with CUDA events inside the run
            (every 16th launch), algorithmic bytes per DESIGN.md.
cpu_baseline = the reference-structured CPU restatement (oracle/, all host threads) on a bounded
            sample (same grid, a few time steps).

`--impl reference` times that CPU restatement alone (Julia is not installed here or on the GPU box,
so the reference's own Base.Threads path cannot run).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback
NPML = 41


def measured_peak():
    """HBM peak in GB/s: the driver-written MEASURED_PEAKS.json when present (the sustained figure if it distinguishes burst and
    sustained -- the kernels are timed inside a long step), else the fallback of B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        found = []

        def walk(x, path):
            if isinstance(x, dict):
                for k, v in x.items():
                    walk(v, path + [str(k).lower()])
            elif isinstance(x, (int, float)) and not isinstance(x, bool):
                joined = "/".join(path)
                if "hbm" in joined or "dram" in joined or "copy" in joined:
                    found.append((joined, float(x)))
        walk(d, [])
        for pref in ("sustain", "hbm_gbs", "hbm_gb_s", "hbm_copy_gbs", "hbm", ""):
            for path, v in found:
                if pref in path and v > 0:
                    if v < 50:            # TB/s
                        v *= 1000.0
                    elif v > 1e6:         # B/s
                        v /= 1e9
                    if 1000.0 < v < 20000.0:
                        return v, "measured"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback"


# ---- algorithmic bytes: Float32, order 2 -----------------------
def algorithmic_bytes_per_step(ndims, elastic, n_ex, npml_faces_per_axis):
    """(velocity kernel bytes, stress kernel bytes) per time step for one wavefield."""
    N = float(np.prod(n_ex))
    if ndims == 3 and elastic:
        vel_f, str_f, vel_m, str_m = 15, 20, 3, 3     # floats per cell; CPML memory variables per slab cell per axis
    elif ndims == 3:
        vel_f, str_f, vel_m, str_m = 10, 6, 1, 1
    elif elastic:
        vel_f, str_f, vel_m, str_m = 9, 11, 2, 2
    else:
        vel_f, str_f, vel_m, str_m = 7, 5, 1, 1
    slab = 0.0
    for q, n in enumerate(n_ex):
        slab += npml_faces_per_axis[q] * NPML * N / n
    return 4 * vel_f * N + 8 * vel_m * slab, 4 * str_f * N + 8 * str_m * slab


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(name, nt_override=None, nss=None):
    from geophyinv_jl_b200.host import gallery
    import geophyinv_jl_b200 as G
    if name == "c3":
        kw = gallery.c3_elastic3d(n=256, nt=nt_override or 2000)
        return dict(kw=kw, attrib=G.FdtdElastic, label="C3: 3-D elastic 256^3 + CPML(41) = 338^3, 1 source, 64 receivers", ndims=3, elastic=True)
    if name == "c3small":
        kw = gallery.c3_elastic3d(n=64, nt=nt_override or 200, fq=20.0)
        return dict(kw=kw, attrib=G.FdtdElastic, label="3-D elastic 64^3 + CPML(41) = 146^3 (debug size)", ndims=3, elastic=True)
    if name == "c2":
        kw = gallery.c2_acou2d_layered(nt=nt_override or 4000, nss=nss or 8)
        return dict(kw=kw, attrib=G.FdtdAcoustic, label=f"C2: 2-D acoustic layered 350x1700 + CPML, {nss or 8} supersources per GPU", ndims=2, elastic=False)
    raise SystemExit(f"unknown workload {name}")


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ====================================================================================================
def run_reference(args):
    """CPU arm: the reference-structured restatement (oracle/) with all host threads, bounded sample."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import geophyinv_jl_b200 as G
    O.build()
    nt_s = args.cpu_steps
    wl = workload(args.workload, nt_override=nt_s)
    po = O.OraclePFdtd(wl["attrib"](), **wl["kw"])
    n_ex = [len(g) for g in po.c.exgrid]
    nss = len(po.local)
    cells = float(np.prod(n_ex)) * nt_s * nss
    for _ in range(args.warmup):
        po.update()
    ts = []
    for _ in range(args.steps):
        t = po.update()
        ts.append(t["run_ms"] * 1e-3)
    sec = float(np.sum(ts))
    val = cells * args.steps / sec / 1e9
    cores = int(po.engine.threads)
    sample = f"{nt_s} time steps of the same extended grid {n_ex} per step ({args.steps} steps timed, {args.warmup} warm-up)"
    out = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": val, "unit": "Gcell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "time_steps_per_step": nt_s, "note": "CPU restatement of src/fdtd (Julia absent); unfused reference sweep structure, OpenMP static"},
        "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_sample(wl_name, nt_s=8):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    wl = workload(wl_name, nt_override=nt_s)
    po = O.OraclePFdtd(wl["attrib"](), **wl["kw"])
    n_ex = [len(g) for g in po.c.exgrid]
    po.update()                                        # warm-up (page faults, thread pool)
    reps, sec = 0, 0.0
    while sec < 10.0 and reps < 6:
        sec += po.update()["run_ms"] * 1e-3
        reps += 1
    cells = float(np.prod(n_ex)) * nt_s * len(po.local) * reps
    cores = int(po.engine.threads)
    po.engine.close()
    return {"value": cells / sec / 1e9, "unit": "Gcell-updates/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {nt_s} time steps of the full extended grid {n_ex} (same medium, source and receivers), {sec:.1f} s of CPU work"}


def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import geophyinv_jl_b200 as G
    wl = workload(args.workload, nt_override=args.nt, nss=args.nss)
    kw = wl["kw"]
    t0 = time.time()
    pa = G.SeisForwExpt(wl["attrib"](), **kw, device=local_rank, shot_batch=args.shot_batch, order=args.order)
    t_build = time.time() - t0
    c = pa.c
    n_ex = [len(g) for g in c.exgrid]
    nt, nss = c.ic["nt"], len(pa.local)
    cells_per_step = float(np.prod(n_ex)) * nt * nss

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        pa.update()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, launches = 0.0, 0.0
    vel_ms = vel_n = str_ms = str_n = 0.0
    for _ in range(args.steps):
        pa.engine.reset(G.engine.RESET_WAVEFIELDS | G.engine.RESET_RECORDS)
        pa.engine.run("forward", [1], [True])
        t = pa.engine.timers()
        dev_ms += t["run_ms"]; launches += t["launches"]
        vel_ms += t["vel_ms"]; vel_n += t["vel_n"]; str_ms += t["stress_ms"]; str_n += t["stress_n"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = max_over_ranks(dev_ms)
    value = cells_per_step * args.steps * world / (dev_ms * 1e-3) / 1e9

    # ---- end to end through the host API with host buffers ----------------------------------------
    medium = kw["medium"]
    nmod = len(c.mparams)
    # update!(pa, medium) copies the UN-extended vp, (vs,) rho (gpi_set_medium_fields: derived parameters and padding on the device)
    h2d = int(nmod * np.prod(medium.vp.shape) * 4)
    d2h = int(sum(nt * c.ageom[0][iss].nr * 4 * len(c.rfields) for iss in pa.local))
    for _ in range(1):
        pa.update_medium(medium); pa.update()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        pa.update_medium(medium)          # update!(pa, medium): pad, H2D of mod arrays, dmod kernel
        pa.update()                       # update!(pa): reset, mod_x_proc!, records D2H into Recs
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    e2e_val = cells_per_step * args.steps * world / e2e_s / 1e9

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    faces = [sum(1 for f in c.pml_faces if f.startswith(ax)) for ax in (("z", "y", "x") if wl["ndims"] == 3 else ("z", "x"))]
    bv, bs = algorithmic_bytes_per_step(wl["ndims"], wl["elastic"], n_ex, faces)
    B = max(1, min(nss, pa.engine.cfg.shot_batch or (16 if wl["ndims"] == 2 else 1)))
    peak, peak_kind = measured_peak()
    kern = {}
    # kernel names as ncu lists them: 3-D elastic runs the TMA-pipelined k_step3t<0|1> unless GPI_TMA3=0
    if pa.engine.kernel_family() == "tma":
        KV, KS = "k_step3t<0> (velocity)", "k_step3t<1> (stress)"
    else:
        KV, KS = ("k_vel3v", "k_stress3v") if wl["ndims"] == 3 else ("k_vel2v", "k_stress2v")
        if args.order == 4:
            KV, KS = ("k_vel4v + k_dirichlet4", "k_stress4v") if os.environ.get("GPI_O4VEC", "1") != "0" else ("k_vel4 + k_dirichlet4", "k_stress4")
    if vel_n > 0 and str_n > 0:
        kern[KV] = {"ms": vel_ms / vel_n, "bytes": bv * B}
        kern[KS] = {"ms": str_ms / str_n, "bytes": bs * B}
    roof = None
    if kern:
        dom = max(kern, key=lambda k: kern[k]["ms"])
        ach = kern[dom]["bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
        step_share = (kern[KV]["ms"] + kern[KS]["ms"]) * nt * (nss / B) * args.steps / max(dev_ms, 1e-9) if world == 1 else None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_kind, "traffic": None, "avg_launch_ms": kern[dom]["ms"],
                "algorithmic_bytes_per_launch": kern[dom]["bytes"],
                "other": {k: {"avg_launch_ms": v["ms"], "achieved": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak} for k, v in kern.items() if k != dom},
                "both_kernels_frac": (bv + bs) * B / ((kern[KV]["ms"] + kern[KS]["ms"]) * 1e-3) / 1e9 / peak,
                "stencil_share_of_step": step_share}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get(args.workload, {}).get(dom)
            except Exception:
                pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline_sample(args.workload, nt_s=args.cpu_steps)
        except Exception as e:      # the baseline is reported, never required
            cpu = {"value": None, "unit": "Gcell-updates/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        out = {
            "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "fd_order": args.order, "shot_batch": int(B), "extended_grid": n_ex, "time_steps_per_step": nt, "supersources_per_gpu": nss,
                       "parallelism": f"one supersource stream per GPU x{world}", "l2": "working set (3.3 GB at C3) far exceeds the 126 MB L2; no flush needed",
                       "interior_equivalent_value": value * float(np.prod([len(g) for g in c.medium.grid])) / float(np.prod(n_ex))},
            "e2e": {"value": e2e_val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "wall_s_timed_region": wall, "build_s": t_build,
            "shots_per_hour": 3600.0 * nss * world * args.steps / (dev_ms * 1e-3),
        }
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE config 5: ONE 3-D elastic shot over z-slabs (strong scaling: the grid is fixed, each GPU owns
    1/N of its z planes and exchanges halo planes over NVLink every half step).  `c5` = [512, 1024, 1024]
    interior cells (594 x 1106 x 1106 extended); `c5small` = 96 x 128 x 128 for debugging."""
    import torch
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    from geophyinv_jl_b200.host import dist as D, slab
    from geophyinv_jl_b200.host.data import AGeomss, ricker
    from geophyinv_jl_b200.host.grids import StepRange
    dist = D.init_process_group("nccl")
    ni = (512, 1024, 1024) if args.workload == "c5" else (96, 128, 128)
    d = 10.0
    grid = [StepRange(0.0, d, m) for m in ni]
    nt = args.nt or 200
    tgrid = StepRange(0.0, 1e-3, nt)
    exn = [m + 82 for m in ni]
    L = [g.last for g in grid]
    src = {"z": [0.5 * L[0] + 0.3 * d], "y": [0.5 * L[1] + 0.2 * d], "x": [0.5 * L[2] + 0.1 * d]}
    nr = 64
    rec = {"z": np.linspace(0.1 * L[0], 0.9 * L[0], nr), "y": np.full(nr, 0.5 * L[1] + 0.2 * d), "x": np.linspace(0.1 * L[2], 0.9 * L[2], nr)}
    need = int(np.ceil((0.16 + 0.15) / 1e-3)) + 2
    wav = ricker(10.0, StepRange(0.0, 1e-3, max(nt, need)), tpeak=0.16)[:nt] * 1e6
    t0 = time.time()
    ex = slab.SlabExpt("elastic", grid, tgrid, AGeomss(src, rec), "vz", wav, ["vz"], slab.synthetic_rows(ni, exn),
                       vp_bounds=(3000.0 * 0.8, 3000.0 * 1.2), rank=rank, nranks=world, device=local_rank)
    ex.attach_nccl(dist)
    t_build = time.time() - t0
    N_ex = float(np.prod(exn))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        ex.update()
    sampler = ClockSampler(local_rank)
    barrier(); sampler.start()
    w0 = time.perf_counter()
    dev_ms = launches = vel_ms = vel_n = str_ms = str_n = 0.0
    for _ in range(args.steps):
        t = ex.update()
        dev_ms += t["run_ms"]; launches += t["launches"]
        vel_ms += t["vel_ms"]; vel_n += t["vel_n"]; str_ms += t["stress_ms"]; str_n += t["stress_n"]
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    dev_ms = allmax(dev_ms)
    value = N_ex * nt * args.steps / (dev_ms * 1e-3) / 1e9
    # end to end: the public call with the records brought to the host
    barrier(); e0 = time.perf_counter()
    for _ in range(args.steps):
        ex.update(); rec_h = ex.records("vz")
    barrier()
    e2e_s = allmax(time.perf_counter() - e0)
    # roofline of the slowest rank's stencil kernels against its share of the algorithmic bytes
    own = (ex.kb - ex.ka) / float(exn[0] + 1)
    bv, bs = algorithmic_bytes_per_step(3, True, exn, [2, 2, 2])
    peak, peak_kind = measured_peak()
    kv, ks = allmax(vel_ms / max(vel_n, 1)), allmax(str_ms / max(str_n, 1))
    fam = ex.engine.kernel_family()                      # 'tma' or 'vec4': narrow slabs run the register-staged kernels
    KV, KS = ("k_step3t<0> (velocity)", "k_step3t<1> (stress)") if fam == "tma" else ("k_vel3v", "k_stress3v")
    roof = {"bound": "hbm", "kernel": KS, "achieved": bs / world / (ks * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": bs / world / (ks * 1e-3) / 1e9 / peak, "peak_source": peak_kind, "traffic": None, "avg_launch_ms": ks,
            "other": {KV: {"avg_launch_ms": kv, "frac": bv / world / (kv * 1e-3) / 1e9 / peak}},
            "both_kernels_frac": (bv + bs) / world / ((kv + ks) * 1e-3) / 1e9 / peak,
            "stencil_share_of_step": (kv + ks) * nt * args.steps / max(dev_ms, 1e-9),
            "note": "per-GPU figures (max over ranks of the kernel times, 1/N of the whole-grid algorithmic bytes)"}
    if rank == 0:
        print(json.dumps({
            "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5: 3-D elastic {list(ni)} + CPML(41) = {exn}, z-slabs over {world} GPU(s), NCCL halo exchange over NVLink",
                       "extended_grid": exn, "time_steps_per_step": nt, "parallelism": f"zslab{world}", "owned_fraction_rank0": own,
                       "l2": "working set exceeds L2 by orders of magnitude; no flush needed"},
            "e2e": {"value": N_ex * nt * args.steps / e2e_s / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(rec_h.nbytes), "note": "inputs of a repeated shot stay resident; records D2H per step"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None, "clocks": clocks,
            "wall_s_timed_region": wall, "build_s": t_build, "ms_per_time_step": dev_ms / args.steps / nt}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_c4(args):
    """BASELINE config 4: 2-D acoustic FWI gradient (forward_save + adjoint + imaging) over the supersources of
    every rank (32 in total at N = 1; `--nss` per GPU otherwise), one NCCL sum all-reduce of the gradient over NVLink.
    A bench step = one `gradient!` call (func_grad.jl:11-49) through the host API: model vector in, gradient out."""
    import torch
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    import geophyinv_jl_b200 as G
    from geophyinv_jl_b200.host import dist as D, gallery
    dist = D.init_process_group("nccl")
    nss_per = args.nss or 32
    nt = args.nt or 3000
    kw, true = gallery.c4_fwi2d(nt=nt, nss=nss_per * world)
    t0 = time.time()
    pa = G.SeisForwExpt(G.FdtdAcoustic("forward_save"), **kw, nworker=world, rank=rank, device=local_rank)
    D.attach_nccl(pa, dist)
    t_build = time.time() - t0
    n_ex = [len(g) for g in pa.c.exgrid]
    # "observed" data: the +5 % box, modelled once (not timed)
    pt = G.SeisForwExpt(G.FdtdAcoustic(), **{**kw, "medium": true}, nworker=world, rank=rank, device=local_rank)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    del pt
    m = pa.get_modelvector()
    g = np.zeros_like(m)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        G.gradient(g, m, dobs, pa)
    sampler = ClockSampler(local_rank)
    barrier(); sampler.start()
    e0 = time.perf_counter()
    dev_ms = launches = 0.0
    for _ in range(args.steps):
        G.gradient(g, m, dobs, pa)
        # device time of the two passes of this call (forward_save, adjoint), from the engine's CUDA events
        dev_ms += pa.last_run_ms
        launches += pa.last_launches
    barrier()
    e2e_s = allmax(time.perf_counter() - e0)
    clocks = sampler.stop()
    dev_ms = allmax(dev_ms)
    cells = float(np.prod(n_ex)) * nt * nss_per * world * 3          # forward_save: 1 wavefield; adjoint: 2
    if rank == 0:
        peak, peak_kind = measured_peak()
        gb = cells / 3 * (48 + 112) / 1e9                             # 48 B forward, 112 B adjoint step per cell
        print(json.dumps({
            "metric": "Gcell-updates/s", "value": cells * args.steps / (dev_ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4: 2-D acoustic FWI gradient, {n_ex} extended cells, {nt} steps, {nss_per} supersources per GPU, NCCL gradient all-reduce",
                       "extended_grid": n_ex, "time_steps_per_step": nt, "supersources_per_gpu": nss_per, "parallelism": f"shots sharded x{world}",
                       "l2": "16 resident shots x 3 wavefields exceed the 126 MB L2",
                       "adjoint_time_levels": "ping-pong (GPI_PINGPONG=1)" if os.environ.get("GPI_PINGPONG", "0") not in ("", "0") else "save_tp copy"},
            "e2e": {"value": cells * args.steps / e2e_s / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(m.nbytes), "d2h_bytes_per_step": int(g.nbytes),
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "forward_save + adjoint passes (k_vel2v, k_stress2v, k_grad2d, boundary, tp copy)",
                         "achieved": gb * args.steps / (dev_ms * 1e-3), "peak": peak, "unit": "GB/s", "frac": gb * args.steps / (dev_ms * 1e-3) / peak / world,
                         "peak_source": peak_kind, "traffic": None, "note": "whole-pass figure: algorithmic bytes of SURVEY 8d (48 + 112 B per cell-step) / device time, per GPU"},
            "cpu_baseline": None, "clocks": clocks, "build_s": t_build,
            "gradients_per_hour": 3600.0 * args.steps / e2e_s, "shots_per_hour": 3600.0 * nss_per * world * args.steps / e2e_s}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--nt", type=int, default=None, help="override the number of time steps per bench step (debug)")
    ap.add_argument("--cpu-steps", type=int, default=8, help="time steps per CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nss", type=int, default=None, help="supersources per GPU (c2, c4)")
    ap.add_argument("--shot-batch", type=int, default=0, help="supersources resident per launch (0 = engine default: 16 in 2-D, 1 in 3-D)")
    ap.add_argument("--order", type=int, default=2, help="_fd_order: 2 (default, the reference's default) or 4")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c5", "c5small"):
        run_c5(args)
    elif args.workload == "c4":
        run_c4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
