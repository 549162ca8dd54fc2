"""SeisInvExpt's inversion grid (reference: src/fwi/fwi.jl:11-38, src/fwi/func_grad.jl:1-68, src/proj_mat.jl:18-26, 208-247):
the separable interpolation between the optimiser's coarse grid and the modelling grid, its transpose for the gradient, and the
gradient on the inversion grid against finite differences of the loss (the check of test/fwi/gradient_accuracy.jl:57-109, here with the
model vector living on the coarse grid)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import geophyinv_jl_b200 as G  # noqa: E402
from geophyinv_jl_b200.host import gallery  # noqa: E402

F32 = np.float32


def test_interpolation_weights_sum_to_one_and_identity():
    """proj_mat.jl:298-301 (`sum(bilinear_interp(...)) ≈ 1`); a grid interpolated on itself is the identity."""
    mm = G.StepRange(0.0, 10.0, 37)
    mi = G.StepRange.from_stop(40.0, 320.0, 9)
    P = G.proj_matrix_1d(mi, mm.values)                   # modelling nodes on the inversion grid
    assert P.shape == (9, 37) and np.allclose(P.sum(axis=0), 1.0, atol=1e-6) and (P >= 0).all()
    assert np.count_nonzero(P, axis=0).max() <= 2
    # nodes outside the inversion grid take its end node (the reference clamps the fraction)
    assert P[0, 0] == 1 and P[-1, -1] == 1
    assert np.array_equal(G.proj_matrix_1d(mm, mm.values), np.eye(37, dtype=F32))


@pytest.mark.parametrize("nd", [2, 3])
def test_projection_is_linear_exact_and_its_transpose_is_the_adjoint(nd):
    rng = np.random.default_rng(3)
    mm = [G.StepRange(0.0, 10.0, n) for n in (31, 23, 27)[:nd]]
    mi = [G.StepRange.from_stop(g.first + 4 * g.step, g.last - 4 * g.step, n) for g, n in zip(mm, (7, 5, 6))]
    P = [G.proj_matrix_1d(a, b.values) for a, b in zip(mi, mm)]
    m = rng.standard_normal([len(g) for g in mi]).astype(F32)
    y = rng.standard_normal([len(g) for g in mm]).astype(F32)
    Pm = G.apply_proj_matrix(m, P)
    Pty = G.apply_proj_matrix(y, [p.T for p in P])
    assert Pm.shape == y.shape and Pty.shape == m.shape
    a, b = float(np.vdot(Pm.astype(np.float64), y)), float(np.vdot(m.astype(np.float64), Pty))
    assert abs(a - b) <= 1e-5 * max(abs(a), abs(b))
    # a (multi)linear function on the coarse grid is reproduced exactly inside it
    co = np.meshgrid(*[g.values for g in mi], indexing="ij")
    lin = sum((q + 1) * c for q, c in enumerate(co)).astype(F32)
    cf = np.meshgrid(*[g.values for g in mm], indexing="ij")
    want = sum((q + 1) * c for q, c in enumerate(cf))
    inside = np.ones(want.shape, bool)
    for q, (gi, gm) in enumerate(zip(mi, mm)):
        sel = (gm.values >= gi.first) & (gm.values <= gi.last)
        inside &= np.moveaxis(np.broadcast_to(sel, want.shape[:q] + want.shape[q + 1:] + (sel.size,)), -1, q)
    got = G.apply_proj_matrix(lin, P)
    assert np.allclose(got[inside], want[inside], rtol=2e-6)


def _inv_case(P_cls, **extra):
    kw, true = gallery.c4_fwi2d(nz=50, nx=70, nt=260, nss=2, nr=12, fq=12.0)
    return P_cls(G.FdtdAcoustic("forward_save"), **kw, **extra), kw, true


def test_inversion_grid_gradient_matches_finite_differences_of_the_loss():
    """CPU (oracle-backed host layer, test infrastructure): g on the coarse grid = P' g_full, checked as a directional derivative."""
    import oracle as O
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**gallery.c4_fwi2d(nz=50, nx=70, nt=260, nss=2, nr=12, fq=12.0)[0],
                                            "medium": gallery.c4_fwi2d(nz=50, nx=70, nt=260, nss=2, nr=12, fq=12.0)[1]})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa, kw, true = _inv_case(O.OraclePFdtd64)
    inv = G.SeisInvExpt(pa, dobs, [9, 12])
    m = inv.get_modelvector()
    assert m.size == 2 * 9 * 12
    g = np.zeros_like(m)
    loss0 = inv.gradient(g, m)
    assert loss0 > 0 and np.abs(g).max() > 0
    rng = np.random.default_rng(0)
    dm = np.zeros_like(m); dm[: m.size // 2] = rng.standard_normal(m.size // 2).astype(F32)       # invK block
    eps = 2e-3
    lp, lm = inv.lossvalue(m + F32(eps) * dm), inv.lossvalue(m - F32(eps) * dm)
    fd = (lp - lm) / (2 * eps)
    ad = float(np.vdot(g.astype(np.float64), dm))
    assert abs(ad / fd - 1) < 2e-2, (ad, fd)
    # migrid == mmgrid: the projection is the identity and the gradient is the modelling-grid gradient
    inv1 = G.SeisInvExpt(pa, dobs)
    m1 = inv1.get_modelvector()
    g1 = np.zeros_like(m1)
    inv1.gradient(g1, m1)
    assert np.array_equal(g1, inv1.gmfull)


@pytest.mark.gpu
def test_inversion_grid_gradient_engine_vs_oracle(G, O):
    kw, true = gallery.c4_fwi2d(nz=50, nx=70, nt=260, nss=2, nr=12, fq=12.0)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    res = []
    for cls in (G.PFdtd, O.OraclePFdtd):
        pa = cls(G.FdtdAcoustic("forward_save"), **kw)
        inv = G.SeisInvExpt(pa, dobs, [9, 12])
        m = inv.get_modelvector()
        g = np.zeros_like(m)
        loss = inv.gradient(g, m)
        res.append((m, g, loss))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][2] == res[1][2]
    assert np.array_equal(res[0][1], res[1][1]), "inversion-grid gradient: engine and oracle differ"
