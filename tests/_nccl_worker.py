"""Worker of tests/test_multigpu.py (torch.distributed.run, one rank per GPU, NCCL).
Shot-sharded forward modelling and the FWI gradient with the engine's own NCCL all-reduce
(gpi_nccl_init / gpi_allreduce_gradients), checked on rank 0 against the CPU oracle run over all shots."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import dist as D, gallery  # noqa: E402


def main():
    dist = D.init_process_group("nccl")
    rank, local_rank, world = D.env_ranks()
    assert dist is not None and world >= 2

    # forward: 5 supersources over the ranks
    kw = gallery.c2_acou2d_layered(nz=60, nx=90, nt=220, nss=5, nr=12, fq=15.0)
    pa = G.SeisForwExpt(G.FdtdAcoustic(), **kw, nworker=world, rank=rank, device=local_rank)
    pa.update()
    D.gather_records(pa, dist, dst=0)

    # FWI gradient: local shots, on-device stack, one NCCL all-reduce inside the engine
    kwg, true = gallery.c4_fwi2d(nz=50, nx=70, nt=300, nss=world + 1, nr=10, fq=12.0)
    pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kwg, nworker=world, rank=rank, device=local_rank)
    D.attach_nccl(pg, dist)
    if rank == 0:
        import oracle as O
        pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kwg, "medium": true}); pt.update()
        dobs = [d.copy() for d in pt.c.data[0]]
    else:
        dobs = None
    box = [dobs]
    dist.broadcast_object_list(box, src=0)
    dobs = box[0]
    m = pg.get_modelvector()
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pg)            # every rank ends up with the all-reduced gradient

    if rank == 0:
        ref = O.OraclePFdtd(G.FdtdAcoustic(), **kw); ref.update()
        for iss in range(5):
            a, b = pa.c.data[0][iss].d["p"], ref.c.data[0][iss].d["p"]
            assert np.abs(b).max() > 0 and np.array_equal(a, b), f"records of supersource {iss} differ"
        pr = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kwg)
        gr = np.zeros_like(m)
        G.gradient(gr, m, dobs, pr)
        err = np.linalg.norm(g - gr) / np.linalg.norm(gr)
        assert err < 1e-4, err
        print(f"NCCL_SHOTS_OK world={world} gradient rel-L2 {err:.2e}")
    # all ranks hold the same reduced gradient
    import torch
    t = torch.from_numpy(g.copy()).cuda()
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    assert torch.equal(t, tmax), "ranks disagree on the all-reduced gradient"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
