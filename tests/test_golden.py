"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the Float32 oracle).

CPU (`-m "not gpu"`): the oracle still reproduces them bit for bit (regression pin of the restatement; a
deterministic OpenMP-static C program must not drift).  GPU (`-m gpu`): the CUDA engine, through the C ABI,
matches the committed vectors to the north-star tolerances without the oracle in the loop."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)

REC_TOL, GRAD_TOL = 1e-5, 1e-4


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("name", ["c1_acou2d_p", "elastic2d_freesurface", "acou3d", "o4_acou2d", "o4_elastic2d_freesurface"])
def test_oracle_reproduces_golden(O, name):
    got, want = MG.records_case(name), load(name)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]) if k.endswith("_dec") else np.allclose(got[k], want[k], rtol=1e-12, atol=0), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MG.CASES))
def test_engine_matches_golden_records(G, name):
    attrib, build, fields = MG.CASES[name]
    pg = G.SeisForwExpt(attrib(), **build())
    pg.update()
    want = load(name)
    for iss, rec in enumerate(pg.c.data[0]):
        for f in fields:
            d = rec.d[f]
            assert rel_l2(d[::8, ::3], want[f"s{iss}_{f}_dec"]) <= REC_TOL, (name, iss, f)
            s, w = MG.summary(d), want[f"s{iss}_{f}_sum"]
            assert np.all(np.abs(s - w) <= 1e-5 * np.abs(w).max()), (name, iss, f, s, w)


@pytest.mark.gpu
def test_engine_matches_golden_gradient(G, O):
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.c4_fwi2d(nz=70, nx=110, nt=500, nss=3, nr=24, fq=10.0)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true}); pt.update()      # observed data (input, not the thing checked)
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw)
    m = pa.get_modelvector()
    g = np.zeros_like(m)
    loss = G.gradient(g, m, dobs, pa)
    want = load("c4_fwi2d_gradient")
    half = g.size // 2
    assert abs(loss - want["loss"][0]) <= 1e-5 * abs(want["loss"][0])
    assert rel_l2(g[:half:7], want["gK_dec"]) <= GRAD_TOL
    assert rel_l2(g[half::7], want["gR_dec"]) <= GRAD_TOL
