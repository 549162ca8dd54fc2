"""Programmatic dependent launch of the 2-D chain (GPI_PDL, kernels.cuh: pdl_wait / pdl_release; engine.cu: launch_pdl).

With the attribute a kernel's CTAs are scheduled while the kernel ahead of it in the stream is still draining; nothing a kernel reads may
come from before the preceding grid has completed.  A missing or misplaced wait shows up as a wavefield that differs from the serialised
run, so every case here is bit-for-bit: the same experiment on a handle created with GPI_PDL=0 and on one created with GPI_PDL=1, launch
by launch and as captured / replayed CUDA graphs (programmatic edges), on grids of several waves of CTAs -- plus the oracle on a small one.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _records(p):
    return [[r.d[f].copy() for f in p.c.rfields] for r in p.c.data[0]]


@pytest.mark.parametrize("physics", ["acoustic", "elastic"])
@pytest.mark.parametrize("graph", ["0", "2"])
def test_pdl_forward_runs_are_bit_identical(G, monkeypatch, physics, graph):
    """Forward runs (k_vel2v, k_stress2v, k_post per time step), one resident shot and a batch of three, 400 x 520 cells + CPML:
    records and final fields equal those of the serialised launches; the launch count is the same."""
    from geophyinv_jl_b200.host import gallery
    if physics == "acoustic":
        kw = gallery.c2_acou2d_layered(nz=400, nx=520, nt=260, nss=3, nr=40, fq=15.0, rfields=("p", "vz"))
        attrib, fields = G.FdtdAcoustic, ("p", "vx", "vz")
    else:
        kw = gallery.elastic2d(nz=300, nx=420, nt=260)
        attrib, fields = G.FdtdElastic, ("vx", "vz", "tauxx", "tauzz", "tauxz")
    monkeypatch.setenv("GPI_GRAPH", graph)
    for batch in (1, 3):
        out = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("GPI_PDL", flag)             # read by gpi_create
            p = G.SeisForwExpt(attrib(), **kw, shot_batch=batch)
            runs = []
            for rep in range(3):                            # graph mode 2: captured at the first run, replayed afterwards
                n = p.update()["launches"]
                runs.append((n, _records(p)))
            out[flag] = (runs, [p.engine.get_field(0, f) for f in fields])
        for rep in range(3):
            n0, r0 = out["0"][0][rep]; n1, r1 = out["1"][0][rep]
            assert n0 == n1
            for a, b in zip(r0, r1):
                for x, y in zip(a, b):
                    assert np.abs(x).max() > 0 and np.array_equal(x, y), f"PDL run {rep} (batch {batch}) differs from the serialised launches"
        for x, y in zip(out["0"][1], out["1"][1]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("physics", ["acoustic", "elastic"])
def test_pdl_gradient_is_bit_identical(G, O, monkeypatch, physics):
    """forward_save + adjoint with ping-pong time levels (k_boundary save / stash / force, out-of-place kernels, k_stress2a for acoustic
    media, the separate imaging kernel for elastic ones): four gradients per handle (launch by launch, captured, replayed twice; odd nt:
    the levels swap per run) equal those of the serialised launches bit for bit, and the oracle's within the gate."""
    from geophyinv_jl_b200.host import gallery
    if physics == "acoustic":
        kw, true = gallery.c4_fwi2d(nz=60, nx=90, nt=301, nss=3, nr=16, fq=10.0)
        attrib = G.FdtdAcoustic
    else:
        kw, true = gallery.fwi2d_elastic(nt=301)
        attrib = G.FdtdElastic
    pt = O.OraclePFdtd(attrib(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    po = O.OraclePFdtd(attrib("forward_save"), **kw)
    m = po.get_modelvector()
    go = np.zeros_like(m)
    G.gradient(go, m, dobs, po)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("GPI_PDL", flag)
        pg = G.PFdtd(attrib("forward_save"), **kw, shot_batch=2)
        for rep in range(4):
            g = np.zeros_like(m)
            loss = G.gradient(g, m, dobs, pg)
            res[flag, rep] = (g, loss, pg.last_launches)
    for rep in range(4):
        g0, l0, n0 = res["0", rep]; g1, l1, n1 = res["1", rep]
        assert np.array_equal(g0, g1) and l0 == l1 and n0 == n1, f"PDL gradient differs from the serialised launches (run {rep})"
        assert rel_l2(g1, go) <= 1e-4


def test_pdl_gradient_on_a_grid_of_several_waves(G, monkeypatch):
    """The same comparison without the oracle on a grid large enough for every kernel to need several waves of CTAs per SM
    (200 x 500 cells + CPML, four resident shots: 5000 CTAs per stencil launch)."""
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.c4_fwi2d(nz=200, nx=500, nt=701, nss=4, nr=60, fq=10.0)
    monkeypatch.setenv("GPI_PDL", "0")
    pt = G.SeisForwExpt(G.FdtdAcoustic(), **{**kw, "medium": true}, shot_batch=4)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("GPI_PDL", flag)
        pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, shot_batch=4)
        m = pg.get_modelvector()
        for rep in range(3):
            g = np.zeros_like(m)
            loss = G.gradient(g, m, dobs, pg)
            res[flag, rep] = (g, loss)
    for rep in range(3):
        assert np.abs(res["0", rep][0]).max() > 0
        assert np.array_equal(res["0", rep][0], res["1", rep][0]) and res["0", rep][1] == res["1", rep][1], f"run {rep}"
