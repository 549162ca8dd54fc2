"""Worker of tests/test_multirank_gloo.py (launched by torch.distributed.run, world_size 2, gloo, CPU).
Exercises the N>1 host path -- shot chunks per rank, record gather, gradient sum, unique-id broadcast --
with the CPU oracle standing in for the GPU engine (test infrastructure only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import dist as D, gallery  # noqa: E402
import oracle as O  # noqa: E402


def main():
    dist = D.init_process_group("gloo")
    rank, _, world = D.env_ranks()
    assert world == 2 and dist is not None
    O.OraclePFdtd.oracle_threads = 2

    # --- forward modelling, 5 supersources over 2 ranks (chunks 0:2 and 2:5, fdtd.jl:251-255)
    kw = gallery.c2_acou2d_layered(nz=60, nx=90, nt=220, nss=5, nr=12, fq=15.0)
    pa = O.OraclePFdtd(G.FdtdAcoustic(), **kw, nworker=world, rank=rank, illum_flag=True)
    assert [list(c) for c in pa.sschunks] == [[0, 1], [2, 3, 4]]
    assert list(pa.local) == [[0, 1], [2, 3, 4]][rank]
    pa.update()
    D.gather_records(pa, dist, dst=0)
    D.stack_illum(pa, dist)              # stack_illums! over the workers (fdtd.jl:556-565)

    # --- the 128-byte id travels from rank 0 to everyone
    uid = D.share_unique_id(lambda: bytes(range(128)), dist)
    assert uid == bytes(range(128))

    # --- FWI gradient: local shots per rank, then one sum over ranks
    kwg, true = gallery.c4_fwi2d(nz=50, nx=70, nt=300, nss=3, nr=10, fq=12.0)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kwg, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pg = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kwg, nworker=world, rank=rank)
    m = pg.get_modelvector()
    g = np.zeros_like(m)
    try:                                  # a sharded adjoint run without a communicator must not hand out a partial gradient silently
        G.gradient(g.copy(), m, dobs, pg)
        raise AssertionError("expected the partial-gradient guard to fire")
    except RuntimeError as e:
        assert "partial" in str(e)
    pg.partial_gradients_ok = True        # the CPU test sums the partial gradients itself (control-plane all-reduce)
    G.gradient(g, m, dobs, pg)
    gsum = D.allreduce_host([g], dist)[0]

    # --- the same for an elastic experiment at order 4 (three parameters, six-point-wide stencils): the host path is generic
    kwe, truee = gallery.fwi2d_elastic(nz=30, nx=36, nt=160, nr=8, nss=3)
    pte = O.OraclePFdtd(G.FdtdElastic(), **{**kwe, "medium": truee}, order=4)
    pte.update()
    dobse = [d.copy() for d in pte.c.data[0]]
    pge = O.OraclePFdtd(G.FdtdElastic("forward_save"), **kwe, nworker=world, rank=rank, order=4)
    me = pge.get_modelvector()
    ge = np.zeros_like(me)
    pge.partial_gradients_ok = True
    G.gradient(ge, me, dobse, pge)
    gesum = D.allreduce_host([ge], dist)[0]

    if rank == 0:
        pre = O.OraclePFdtd(G.FdtdElastic("forward_save"), **kwe, order=4)
        gre = np.zeros_like(me)
        G.gradient(gre, me, dobse, pre)
        erre = np.linalg.norm(gesum - gre) / np.linalg.norm(gre)
        assert np.abs(gre).max() > 0 and erre < 1e-5, erre
        ref = O.OraclePFdtd(G.FdtdAcoustic(), **kw, illum_flag=True)
        ref.update()
        assert ref.c.illum_stack.max() > 0 and np.allclose(pa.c.illum_stack, ref.c.illum_stack, rtol=1e-13, atol=0), "illumination stacked over the ranks differs"
        for iss in range(5):
            a, b = pa.c.data[0][iss].d["p"], ref.c.data[0][iss].d["p"]
            assert np.abs(b).max() > 0 and np.array_equal(a, b), f"records of supersource {iss} differ"
        pr = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kwg)
        gr = np.zeros_like(m)
        G.gradient(gr, m, dobs, pr)
        err = np.linalg.norm(gsum - gr) / np.linalg.norm(gr)
        assert err < 1e-5, err
        print(f"MULTIRANK_OK gradient rel-L2 {err:.2e}, elastic order-4 gradient rel-L2 {erre:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
