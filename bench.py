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
so the reference's own Base.Threads path cannot run).  It uses every host core it is allowed
to (the affinity mask, NOT the OMP_NUM_THREADS=1 that torchrun exports) and prints the same `config`.

extra     = after the headline leg, the two paths that COMMUNICATE run in the same process group and land
            in the same JSON line (skip with --no-extra):
            extra.c4: BASELINE config 4, 32 supersources in total sharded over the N ranks (strong scaling),
                      one NCCL sum all-reduce of the gradient per `gradient!`, `parity_ok` = the N-rank gradient of
                      a reduced grid against rank 0 running the same shots alone;
            extra.c5: BASELINE config 5, one 3-D elastic shot over N z-slabs (strong scaling), halo planes over
                      NVLink every half step, `slab_parity_ok` = records and owned wavefield rows of a reduced grid
                      against the single-GPU run, `exchange_share` = sampled exchange time / step time.
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


def host_threads():
    """Host cores this process may use: the affinity mask (torchrun exports OMP_NUM_THREADS=1, which says nothing about the box)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


_DIST = None


def get_dist():
    """One torch.distributed process group (NCCL) for every leg of the run; None for a single process."""
    global _DIST
    import torch
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1 and _DIST is None:
        from geophyinv_jl_b200.host import dist as D
        _DIST = D.init_process_group("nccl")
    return _DIST


def barrier():
    import torch
    torch.cuda.synchronize()
    if _DIST is not None:
        _DIST.barrier()
    torch.cuda.synchronize()


def allmax(x):
    import torch
    if _DIST is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    _DIST.all_reduce(t, op=_DIST.ReduceOp.MAX)
    return float(t.item())


def bench_config(wl, order, B, n_ex, nt, nss, world):
    """`config` of the headline leg -- the SAME dict for our arm and for the reference arm (the CPU arm times a bounded sample of it:
    that is said in its cpu_baseline.sample, not here)."""
    return {"workload": wl["label"], "fd_order": order, "shot_batch": int(B), "extended_grid": [int(x) for x in n_ex], "time_steps_per_step": int(nt),
            "supersources_per_gpu": int(nss), "parallelism": f"one supersource stream per GPU x{world}",
            "l2": "working set (3.3 GB at C3) far exceeds the 126 MB L2; no flush needed"}


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
    O.OraclePFdtd.oracle_threads = host_threads()            # explicit: torchrun exports OMP_NUM_THREADS=1
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
    nt_full = args.nt or {"c3": 2000, "c3small": 200, "c2": 4000}.get(args.workload, nt_s)
    B = args.shot_batch or (16 if wl["ndims"] == 2 else 1)
    sample = (f"{nt_s} of the {nt_full} time steps of the same extended grid {n_ex} per bench step ({args.steps} steps timed, {args.warmup} warm-up); "
              f"the metric is per cell update, so the sample length does not enter it; {cores} OpenMP threads")
    print(f"bench.py --impl reference: {cores} host threads (affinity mask; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')})", file=sys.stderr, flush=True)
    out = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": val, "unit": "Gcell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, args.order, max(1, min(nss, B)), n_ex, nt_full, nss, max(1, world)),
        "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "sample": sample, "sample_time_steps": nt_s},
        "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of src/fdtd (Julia absent); unfused reference sweep structure, OpenMP static; rank 0 only",
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_sample(wl_name, nt_s=8):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    wl = workload(wl_name, nt_override=nt_s)
    O.OraclePFdtd.oracle_threads = host_threads()
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
    dist = get_dist()
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

    max_over_ranks = allmax

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

    out = None
    if rank == 0:
        out = {
            "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(wl, args.order, B, n_ex, nt, nss, world),
            "interior_equivalent_value": value * float(np.prod([len(g) for g in c.medium.grid])) / float(np.prod(n_ex)),
            "e2e": {"value": e2e_val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "wall_s_timed_region": wall, "build_s": t_build,
            "shots_per_hour": 3600.0 * nss * world * args.steps / (dev_ms * 1e-3),
        }
    del pa
    return out


def slab_parity(dist, rank, local_rank, world):
    """A reduced 3-D elastic grid over `world` z-slabs against rank 0 propagating the same shot alone: every wavefield's owned rows
    and the records, bit for bit (tests/_slab_worker.py is the pytest form).  Returns a dict on rank 0."""
    import torch
    import geophyinv_jl_b200 as G
    from geophyinv_jl_b200.host import dist as D, gallery
    if world == 1:
        return {"slab_parity_ok": None, "note": "one rank: the slab run IS the single-GPU run"}
    n = max(38, 12 * world + 8)
    kw = gallery.c3_elastic3d(n=n, nt=120, nr=12, fq=30.0, rfields=("vz", "vx"))
    L = kw["medium"].grid[0].last
    a = kw["ageom"][0]
    a.r["z"][...] = np.linspace(0.1 * L, 0.9 * L, a.nr)          # receivers on a vertical line: taps straddle the cuts
    ps = G.SeisForwExpt(G.FdtdElastic(), **kw, device=local_rank, zslab=(rank, world))
    D.attach_nccl(ps, dist)
    ps.update()
    ref = None
    if rank == 0:
        ref = G.SeisForwExpt(G.FdtdElastic(), **kw, device=local_rank)
        ref.update()
    fields_equal, worst = True, 0.0
    for f in ("tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"):
        t = torch.from_numpy(np.ascontiguousarray(ps.engine.get_field(0, f))).cuda()
        dist.all_reduce(t)                                          # slabs are disjoint: the sum is the whole field
        if rank == 0:
            want = ref.engine.get_field(0, f)
            fields_equal = fields_equal and bool(np.abs(want).max() > 0) and bool(np.array_equal(t.cpu().numpy(), want))
    rec_equal = True
    if rank == 0:
        for f in ps.c.rfields:
            x, y = ps.c.data[0][0].d[f], ref.c.data[0][0].d[f]
            rec_equal = rec_equal and bool(np.array_equal(x, y))
            worst = max(worst, float(np.linalg.norm(x.astype(np.float64) - y) / np.linalg.norm(y)))
    del ps, ref
    if rank != 0:
        return None
    return {"slab_parity_ok": bool(fields_equal and worst <= 1e-6), "wavefields_bit_identical": bool(fields_equal), "records_bit_identical": bool(rec_equal),
            "records_rel_l2": worst, "grid": f"3-D elastic {n}^3 + CPML, 120 steps, 12 receivers on a vertical line, {world} z-slabs vs 1 GPU",
            "note": "a receiver whose taps straddle a cut is summed over two ranks (different association): records gate 1e-6, wavefields bit for bit"}


def run_c5(args, steps=None, warmup=None):
    """BASELINE config 5: ONE 3-D elastic shot over z-slabs (strong scaling: the grid is fixed, each GPU owns
    1/N of its z planes and exchanges halo planes over NVLink every half step).  `c5` = [512, 1024, 1024]
    interior cells (594 x 1106 x 1106 extended); `c5small` = 96 x 128 x 128 for debugging."""
    import torch
    rank, local_rank, world = dist_env()
    dist = get_dist()
    from geophyinv_jl_b200.host import slab
    from geophyinv_jl_b200.host.data import AGeomss, ricker
    from geophyinv_jl_b200.host.grids import StepRange
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    small = args.workload == "c5small" or args.c5_small
    ni = (96, 128, 128) if small else (512, 1024, 1024)
    if getattr(args, "c5_ni", None):           # e.g. 68,1024,1024 on 2 GPUs: the 75-plane slab windows C5 has on 8 GPUs
        ni = tuple(int(v) for v in args.c5_ni.split(","))
    d = 10.0
    grid = [StepRange(0.0, d, m) for m in ni]
    nt = (args.nt if args.workload in ("c5", "c5small") else None) or 200
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

    for _ in range(warmup):
        ex.update()
    sampler = ClockSampler(local_rank)
    barrier(); sampler.start()
    w0 = time.perf_counter()
    dev_ms = launches = vel_ms = vel_n = str_ms = str_n = exch_ms = exch_n = 0.0
    for _ in range(steps):
        t = ex.update()
        dev_ms += t["run_ms"]; launches += t["launches"]
        vel_ms += t["vel_ms"]; vel_n += t["vel_n"]; str_ms += t["stress_ms"]; str_n += t["stress_n"]
        exch_ms += t.get("exch_ms", 0.0); exch_n += t.get("exch_n", 0.0)
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    dev_ms = allmax(dev_ms)
    value = N_ex * nt * steps / (dev_ms * 1e-3) / 1e9
    # end to end: the public call with the records brought to the host
    barrier(); e0 = time.perf_counter()
    for _ in range(steps):
        ex.update(); rec_h = ex.records("vz")
    barrier()
    e2e_s = allmax(time.perf_counter() - e0)
    # roofline of the slowest rank's stencil kernels against its share of the algorithmic bytes
    own = (ex.kb - ex.ka) / float(exn[0] + 1)
    bv, bs = algorithmic_bytes_per_step(3, True, exn, [2, 2, 2])
    peak, peak_kind = measured_peak()
    kv, ks = allmax(vel_ms / max(vel_n, 1)), allmax(str_ms / max(str_n, 1))
    # two exchanges per time step; exch_* are sampled on the same steps as the kernels
    exch_per_step = allmax(2.0 * exch_ms / max(exch_n, 1)) if world > 1 else 0.0
    fam = ex.engine.kernel_family()                      # 'tma', 'vec4' or 'vec4-pipelined': narrow slabs run the register-staged kernels
    piped = fam == "vec4-pipelined"                      # halo exchange on a side stream beside the other x half of each launch
    KV, KS = ("k_step3t<0> (velocity)", "k_step3t<1> (stress)") if fam == "tma" else ("k_vel3v", "k_stress3v")
    roof = {"bound": "hbm", "kernel": KS, "achieved": bs / world / (ks * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": bs / world / (ks * 1e-3) / 1e9 / peak, "peak_source": peak_kind, "traffic": None, "avg_launch_ms": ks,
            "other": {KV: {"avg_launch_ms": kv, "frac": bv / world / (kv * 1e-3) / 1e9 / peak}},
            "both_kernels_frac": (bv + bs) / world / ((kv + ks) * 1e-3) / 1e9 / peak,
            "stencil_share_of_step": (kv + ks) * nt * steps / max(dev_ms, 1e-9),
            "whole_step_frac": (bv + bs) / world * nt * steps / (dev_ms * 1e-3) / 1e9 / peak,
            "note": "per-GPU figures (max over ranks of the kernel times, 1/N of the whole-grid algorithmic bytes)"}
    del ex
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {
        "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5: 3-D elastic {list(ni)} + CPML(41) = {exn}, z-slabs over {world} GPU(s), halo planes over NVLink",
                   "extended_grid": exn, "time_steps_per_step": nt, "parallelism": f"zslab{world}", "owned_fraction_rank0": own,
                   "l2": "working set exceeds L2 by orders of magnitude; no flush needed"},
        "e2e": {"value": N_ex * nt * steps / e2e_s / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(rec_h.nbytes), "note": "inputs of a repeated shot stay resident; records D2H per step"},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None, "clocks": clocks,
        # serial exchange: its time on the compute stream.  Pipelined: exch_* is the side stream's time from the first pack to the last
        # unpack, most of it beside a stencil kernel; what the exchange still costs the step is the step minus the stencil kernels
        "exchange_ms_per_time_step": exch_per_step, "exchange_overlapped": bool(piped),
        "exchange_share": (max(0.0, 1.0 - roof["stencil_share_of_step"]) if piped else exch_per_step * nt * steps / max(dev_ms, 1e-9)),
        "wall_s_timed_region": wall, "build_s": t_build, "ms_per_time_step": dev_ms / steps / nt}


def c4_parity(dist, rank, local_rank, world):
    """The FWI gradient of a reduced grid, 8 supersources sharded over `world` ranks + NCCL all-reduce, against rank 0 running all 8 alone.
    Per-rank partial sums are stacked in shot order and NCCL sums the partials in its own order, so the association differs from the
    one-rank stack: the gate is the north-star's 1e-4 (measured ~1e-7); `bit_identical` is reported as found."""
    import geophyinv_jl_b200 as G
    from geophyinv_jl_b200.host import dist as D, gallery
    if world == 1:
        return {"parity_ok": None, "note": "one rank: no all-reduce"}
    kwg, true = gallery.c4_fwi2d(nz=60, nx=90, nt=300, nss=8, nr=12, fq=12.0)
    pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kwg, nworker=world, rank=rank, device=local_rank)
    D.attach_nccl(pg, dist)
    box = [None]
    p1 = None
    if rank == 0:
        pt = G.SeisForwExpt(G.FdtdAcoustic(), **{**kwg, "medium": true}, device=local_rank)
        pt.update()
        box[0] = [d.copy() for d in pt.c.data[0]]
        del pt
    dist.broadcast_object_list(box, src=0)
    dobs = box[0]
    m = pg.get_modelvector()
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pg)                    # every rank ends with the all-reduced gradient
    out = None
    if rank == 0:
        p1 = G.PFdtd(G.FdtdAcoustic("forward_save"), **kwg, device=local_rank)
        g1 = np.zeros_like(m)
        G.gradient(g1, m, dobs, p1)
        err = float(np.linalg.norm(g.astype(np.float64) - g1) / np.linalg.norm(g1))
        out = {"parity_ok": bool(np.isfinite(err) and np.abs(g1).max() > 0 and err <= 1e-4), "gradient_rel_l2": err, "bit_identical": bool(np.array_equal(g, g1)),
               "grid": f"2-D acoustic 60x90 + CPML, 300 steps, 8 supersources over {world} ranks vs 1 GPU, parameters invK and rho"}
        del p1
    del pg
    return out


def run_c4(args, steps=None, warmup=None, nss_total=None):
    """BASELINE config 4: 2-D acoustic FWI gradient (forward_save + adjoint + imaging), one NCCL sum all-reduce of the gradient over NVLink.
    A bench step = one `gradient!` call (func_grad.jl:11-49) through the host API: model vector in, gradient out.
    `nss_total` supersources in total sharded over the ranks (strong scaling, the extra leg); else `--nss` per GPU (weak, default 32)."""
    import torch
    rank, local_rank, world = dist_env()
    dist = get_dist()
    import geophyinv_jl_b200 as G
    from geophyinv_jl_b200.host import dist as D, gallery
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    strong = nss_total is not None
    nss_all = nss_total if strong else (args.nss or 32) * world
    nt = (args.nt if args.workload == "c4" else None) or 3000
    kw, true = gallery.c4_fwi2d(nt=nt, nss=nss_all)
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

    for _ in range(warmup):
        G.gradient(g, m, dobs, pa)
    sampler = ClockSampler(local_rank)
    barrier(); sampler.start()
    e0 = time.perf_counter()
    dev_ms = launches = ar_ms = 0.0
    for _ in range(steps):
        G.gradient(g, m, dobs, pa)
        # device time of the two passes of this call (forward_save, adjoint) plus the all-reduce, from the engine's CUDA events
        ar = pa.engine.timers().get("allreduce_ms", 0.0) if world > 1 else 0.0
        dev_ms += pa.last_run_ms + ar
        ar_ms += ar
        launches += pa.last_launches
    barrier()
    e2e_s = allmax(time.perf_counter() - e0)
    clocks = sampler.stop()
    dev_ms = allmax(dev_ms)
    ar_ms = allmax(ar_ms)
    cells = float(np.prod(n_ex)) * nt * nss_all * 3          # forward_save: 1 wavefield; adjoint: 2
    nss_local = len(pa.local)
    B = int(pa.engine.cfg.shot_batch or min(16, max(nss_local, 1)))
    pp = os.environ.get("GPI_PINGPONG", "1") not in ("", "0")
    del pa
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_kind = measured_peak()
    gb = cells / 3 * (48 + 112) / 1e9                             # 48 B forward, 112 B adjoint step per cell
    return {
        "metric": "Gcell-updates/s", "value": cells * steps / (dev_ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4: 2-D acoustic FWI gradient, {n_ex} extended cells, {nt} steps, {nss_all} supersources in total ({nss_local} on rank 0), NCCL gradient all-reduce",
                   "extended_grid": n_ex, "time_steps_per_step": nt, "supersources_total": nss_all, "supersources_rank0": nss_local, "shot_batch": B,
                   "parallelism": f"shots sharded x{world}", "l2": "the resident shots x 3 wavefields exceed the 126 MB L2 from 6 shots on",
                   "adjoint_time_levels": "ping-pong (default)" if pp else "save_tp copy (GPI_PINGPONG=0)"},
        "e2e": {"value": cells * steps / e2e_s / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(m.nbytes), "d2h_bytes_per_step": int(g.nbytes),
                "ms_per_step": e2e_s / steps * 1e3},
        "gpu_launches": int(launches), "allreduce_ms_per_gradient": ar_ms / steps,
        "roofline": {"bound": "hbm", "kernel": "forward_save + adjoint passes (k_vel2v, k_stress2v, imaging, boundary)",
                     "achieved": gb * steps / (dev_ms * 1e-3) / world, "peak": peak, "unit": "GB/s", "frac": gb * steps / (dev_ms * 1e-3) / peak / world,
                     "peak_source": peak_kind, "traffic": None, "note": "whole-pass figure: algorithmic bytes of SURVEY 8d (48 + 112 B per cell-step) / device time, per GPU"},
        "cpu_baseline": None, "clocks": clocks, "build_s": t_build,
        "gradients_per_hour": 3600.0 * steps / e2e_s, "shots_per_hour": 3600.0 * nss_all * steps / e2e_s}


def extra_legs(args):
    """The two communicating paths (BASELINE configs 4 and 5) in the same process group, after the headline leg."""
    rank, local_rank, world = dist_env()
    dist = get_dist()
    extra = {}
    for name, fn in (("c4", lambda: {**(run_c4(args, steps=2, warmup=3, nss_total=32) or {}), **(c4_parity(dist, rank, local_rank, world) or {})}),
                     ("c5", lambda: {**(run_c5(args, steps=2, warmup=3) or {}), **(slab_parity(dist, rank, local_rank, world) or {})})):
        t0 = time.time()
        try:
            leg = fn()
        except Exception as e:                     # an extra leg must not take the headline line with it; the failure is reported
            leg = {"failed": f"{type(e).__name__}: {e}"}
        if rank == 0:
            leg["leg_wall_s"] = time.time() - t0
            extra[name] = leg
    return extra if rank == 0 else None


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
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 legs that follow the headline C3 leg")
    ap.add_argument("--c5-small", action="store_true", help="extra C5 leg on the 96 x 128 x 128 debug grid")
    ap.add_argument("--c5-ni", default=None, help="interior cells nz,ny,nx of the c5 workload (debug / slab-shape studies)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload in ("c5", "c5small"):
        out = run_c5(args)
    elif args.workload == "c4":
        out = run_c4(args)
    else:
        out = run_ours(args)
        if args.workload == "c3" and not args.no_extra:
            extra = extra_legs(args)
            if out is not None:
                out["extra"] = extra
    if out is not None:
        print(json.dumps(out), flush=True)
    if _DIST is not None:
        barrier()
        _DIST.destroy_process_group()


if __name__ == "__main__":
    main()
