"""Worker of tests/test_multigpu.py::test_zslab_decomposition_matches_single_gpu (one rank per GPU, NCCL).
One 3-D experiment split into z-slabs over the ranks; rank 0 also runs the same experiment on its GPU alone.
Owned rows of every wavefield must equal the single-GPU run bit for bit, records to 1e-6 (a receiver whose
taps straddle a cut is summed in a different order)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import dist as D, gallery  # noqa: E402


def run_case(name, attrib, kw, fields, dist, rank, world, local_rank):
    import torch
    ps = G.SeisForwExpt(attrib(), **kw, device=local_rank, zslab=(rank, world))
    D.attach_nccl(ps, dist)
    ps.update()
    ka, kb = ps.engine.slab_range()
    worst = 0.0
    ref = None
    if rank == 0:
        ref = G.SeisForwExpt(attrib(), **kw, device=local_rank)
        ref.update()
    for f in fields:
        mine = ps.engine.get_field(0, f)                       # owned rows, zeros elsewhere
        t = torch.from_numpy(np.ascontiguousarray(mine)).cuda()
        dist.all_reduce(t)                                      # slabs are disjoint: the sum is the whole field
        if rank == 0:
            whole, want = t.cpu().numpy(), ref.engine.get_field(0, f)
            assert np.abs(want).max() > 0, f"{name}: field {f} is empty, the case tests nothing"
            assert np.array_equal(whole, want), f"{name}: field {f} differs from the single-GPU run (max abs {np.abs(whole - want).max():.3e})"
    if rank == 0:
        for f in ps.c.rfields:
            a, b = ps.c.data[0][0].d[f], ref.c.data[0][0].d[f]
            assert np.abs(b).max() > 0
            worst = max(worst, float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b)))
        assert worst <= 1e-6, f"{name}: records differ, rel-L2 {worst:.3e}"
    # every rank holds the same all-reduced records
    r = torch.from_numpy(np.ascontiguousarray(ps.c.data[0][0].d[ps.c.rfields[0]])).cuda()
    rmax = r.clone(); dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
    assert torch.equal(r, rmax)
    return worst, (ka, kb)


def main():
    dist = D.init_process_group("nccl")
    rank, local_rank, world = D.env_ranks()
    assert dist is not None and world >= 2
    out = []
    # elastic, PML on all faces, receivers on a vertical line so that taps straddle the cuts
    kw = gallery.c3_elastic3d(n=38, nt=140, nr=12, fq=30.0, rfields=("vz", "vx"))
    L = kw["medium"].grid[0].last
    a = kw["ageom"][0]
    a.r["z"][...] = np.linspace(0.1 * L, 0.9 * L, a.nr)
    out.append(run_case("elastic", G.FdtdElastic, kw, ["tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"], dist, rank, world, local_rank))
    # elastic with a free surface on zmin and no PML there
    kw = gallery.c3_elastic3d(n=44, nt=140, nr=10, fq=30.0, rfields=("vz",), stressfree=True)
    out.append(run_case("elastic-freesurface", G.FdtdElastic, kw, ["tauzz", "tauxz", "vz", "vx"], dist, rank, world, local_rank))
    # acoustic
    kw = gallery.acou3d(n=45, nt=200)
    out.append(run_case("acoustic", G.FdtdAcoustic, kw, ["p", "vx", "vy", "vz"], dist, rank, world, local_rank))
    if rank == 0:
        print("SLAB_OK world=%d " % world + " ".join(f"records {w:.1e} slab0={r}" for w, r in out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
