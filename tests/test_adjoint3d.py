"""3-D acoustic adjoint + gradient imaging.  Upstream has `gradlame!` for any dimension but `gradrho!` and
`compute_gradient!` for 2-D only (src/fdtd/gradient.jl:17-61); the 3-D method is the same construction with a y term.
No upstream counterpart exists to compare with, so the check is the one upstream's own gradient test uses
(test/fwi/gradient_accuracy.jl:57-109): the adjoint-state gradient against finite differences of the loss."""
import numpy as np
import pytest

from conftest import rel_l2

GRAD_TOL = 1e-4


def test_gradient3d_vs_finite_differences(G, O):
    from geophyinv_jl_b200.host import gallery
    from scipy.ndimage import gaussian_filter
    n = 16
    kw, true = gallery.fwi3d(n=n, nt=150, nr=10, nss=1)
    # CPML on one face per axis: the adjoint-state identity does not depend on what the other three faces do, and the eleven Float64
    # passes run over (16 + 41)^3 instead of (16 + 82)^3 cells (the six-face adjoint path is compared with the engine in the GPU tests)
    kw = {**kw, "pml_faces": ["zmax", "ymax", "xmax"]}
    pt = O.OraclePFdtd64(G.FdtdAcoustic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd64(G.FdtdAcoustic("forward_save"), **kw)
    m = pa.get_modelvector().astype(np.float64)
    rng = np.random.default_rng(5)
    dK = gaussian_filter(rng.standard_normal((n, n, n)), 3.0).ravel(order="F")
    dR = gaussian_filter(rng.standard_normal((n, n, n)), 3.0).ravel(order="F")
    fds = {}
    for name, dm in (("invK", np.concatenate([dK, 0 * dR])), ("rho", np.concatenate([0 * dK, dR]))):
        dm = dm / np.abs(dm).max()
        eps = 2e-3
        fds[name] = ((G.lossvalue(m + eps * dm, dobs, pa) - G.lossvalue(m - eps * dm, dobs, pa)) / (2 * eps), dm)
    for unshifted, tol_rho in ((False, 0.12), (True, 0.02)):
        pa.unshifted_rho = unshifted
        g = np.zeros_like(m)
        G.gradient(g, m, dobs, pa)
        for name, (fd, dm) in fds.items():
            ratio = float(np.dot(g, dm)) / fd
            print(f"3-D acoustic, unshifted={unshifted}: d loss / d {name} adjoint / finite-difference = {ratio:.4f}")
            assert abs(ratio - 1) < (0.02 if name == "invK" else tol_rho)


@pytest.mark.gpu
@pytest.mark.parametrize("unshifted", [False, True])
def test_gradient3d_engine_matches_oracle(G, O, unshifted):
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.fwi3d(n=26, nt=200, nr=12)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw)
    po = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw)
    pg.unshifted_rho = po.unshifted_rho = unshifted
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    lg, lo = G.gradient(gg, m, dobs, pg), G.gradient(go, m, dobs, po)
    half = m.size // 2
    eK, eR = rel_l2(gg[:half], go[:half]), rel_l2(gg[half:], go[half:])
    print(f"3-D FWI gradient unshifted={unshifted}: invK rel-L2 {eK:.3e}, rho rel-L2 {eR:.3e}, loss {lg:.6e} vs {lo:.6e}")
    assert abs(lg - lo) <= 1e-5 * abs(lo)
    assert np.abs(go[:half]).max() > 0 and np.abs(go[half:]).max() > 0
    assert eK <= GRAD_TOL and eR <= GRAD_TOL


@pytest.mark.gpu
def test_gradient3d_order4_engine_matches_oracle(G, O):
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.fwi3d(n=22, nt=150, nr=8, nss=1)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true}, order=4)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, order=4)
    po = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw, order=4)
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    G.gradient(gg, m, dobs, pg); G.gradient(go, m, dobs, po)
    half = m.size // 2
    eK, eR = rel_l2(gg[:half], go[:half]), rel_l2(gg[half:], go[half:])
    print(f"3-D FWI gradient order 4: invK rel-L2 {eK:.3e}, rho rel-L2 {eR:.3e}")
    assert eK <= GRAD_TOL and eR <= GRAD_TOL


# --------------------------------------------------------------------------------------------------
# 2-D elastic gradient (invlambda, invmu, rho): nothing upstream; finite differences are the check
# --------------------------------------------------------------------------------------------------
def test_elastic2d_gradient_vs_finite_differences(G, O):
    from geophyinv_jl_b200.host import gallery
    from scipy.ndimage import gaussian_filter
    kw, true = gallery.fwi2d_elastic()
    pt = O.OraclePFdtd64(G.FdtdElastic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd64(G.FdtdElastic("forward_save"), **kw)
    assert pa.c.mparams == ["invlambda", "invmu", "rho"]
    m = pa.get_modelvector().astype(np.float64)
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
    rng = np.random.default_rng(5)
    n3 = m.size // 3
    for k, name in enumerate(pa.c.mparams):
        dm = np.zeros_like(m)
        dm[k * n3:(k + 1) * n3] = gaussian_filter(rng.standard_normal((44, 56)), 4.0).ravel(order="F")
        dm /= np.abs(dm).max()
        eps = 2e-3
        fd = (G.lossvalue(m + eps * dm, dobs, pa) - G.lossvalue(m - eps * dm, dobs, pa)) / (2 * eps)
        ratio = float(np.dot(g, dm)) / fd
        print(f"2-D elastic: d loss / d {name} adjoint / finite-difference = {ratio:.4f}")
        assert abs(ratio - 1) < 0.06


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_elastic2d_gradient_engine_matches_oracle(G, O, order):
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.fwi2d_elastic(nz=52, nx=66, nt=380, nss=3)
    pt = O.OraclePFdtd(G.FdtdElastic(), **{**kw, "medium": true}, order=order)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pg = G.PFdtd(G.FdtdElastic("forward_save"), **kw, shot_batch=2, order=order)
    po = O.OraclePFdtd(G.FdtdElastic("forward_save"), **kw, order=order)
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    lg, lo = G.gradient(gg, m, dobs, pg), G.gradient(go, m, dobs, po)
    assert abs(lg - lo) <= 1e-5 * abs(lo)
    n3 = m.size // 3
    for k, name in enumerate(pg.c.mparams):
        e = rel_l2(gg[k * n3:(k + 1) * n3], go[k * n3:(k + 1) * n3])
        print(f"2-D elastic gradient order {order}, {name}: rel-L2 {e:.3e}")
        assert np.abs(go[k * n3:(k + 1) * n3]).max() > 0
        assert e <= GRAD_TOL


# --------------------------------------------------------------------------------------------------
# 3-D elastic: boundary store of all six stresses (no upstream method, boundary.jl:215-264) + gradient
# --------------------------------------------------------------------------------------------------
def test_elastic3d_gradient_vs_finite_differences(G, O):
    """Perturbation = a Gaussian bump in the middle of the model, away from the source and receiver cells (the loss also depends on
    rho at those cells through the source scaling dt/rho, source.jl:166-177, which no adjoint-state gradient -- upstream's
    included -- accounts for).  invlambda / invmu agree with finite differences to 1e-3; rho carries the one-cell shift of
    upstream's combine_gmodrho! construction."""
    from geophyinv_jl_b200.host import gallery
    n = 14
    kw, true = gallery.fwi3d_elastic(n=n, nt=100)
    kw = {**kw, "pml_faces": ["zmax", "ymax", "xmax"]}          # as above: (14 + 41)^3 cells per pass instead of (14 + 82)^3
    pt = O.OraclePFdtd64(G.FdtdElastic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd64(G.FdtdElastic("forward_save"), **kw)
    m = pa.get_modelvector().astype(np.float64)
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
    zz, yy, xx = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    c0 = 0.5 * (n - 1)
    bump = np.exp(-((zz - c0) ** 2 + (yy - c0) ** 2 + (xx - c0) ** 2) / (2 * 2.0 ** 2)).ravel(order="F")
    n3 = m.size // 3
    for k, (name, tol) in enumerate(zip(pa.c.mparams, (0.01, 0.01, 0.3))):
        dm = np.zeros_like(m)
        dm[k * n3:(k + 1) * n3] = bump
        eps = 2e-3
        fd = (G.lossvalue(m + eps * dm, dobs, pa) - G.lossvalue(m - eps * dm, dobs, pa)) / (2 * eps)
        ratio = float(np.dot(g, dm)) / fd
        print(f"3-D elastic: d loss / d {name} adjoint / finite-difference = {ratio:.5f}")
        assert abs(ratio - 1) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_elastic3d_adjoint_engine_matches_oracle(G, O, order):
    """forward_save (six stresses on 3+3 planes per axis) + adjoint + imaging through the TMA-pipelined / order-4 kernels."""
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.fwi3d_elastic(n=22, nt=160, nr=8)
    pt = O.OraclePFdtd(G.FdtdElastic(), **{**kw, "medium": true}, order=order)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pg = G.PFdtd(G.FdtdElastic("forward_save"), **kw, order=order)
    po = O.OraclePFdtd(G.FdtdElastic("forward_save"), **kw, order=order)
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    lg, lo = G.gradient(gg, m, dobs, pg), G.gradient(go, m, dobs, po)
    assert abs(lg - lo) <= 1e-5 * abs(lo)
    n3 = m.size // 3
    for k, name in enumerate(pg.c.mparams):
        e = rel_l2(gg[k * n3:(k + 1) * n3], go[k * n3:(k + 1) * n3])
        print(f"3-D elastic gradient order {order}, {name}: rel-L2 {e:.3e}")
        assert np.abs(go[k * n3:(k + 1) * n3]).max() > 0
        assert e <= GRAD_TOL
    # the back-propagated forward field (pw 1) ends where the forward run started: compare the final fields of both engines
    for f in ("tauxx", "tauxy", "tauyz", "vx", "vz"):
        assert np.array_equal(pg.engine.get_field(0, f), po.engine.get_field(0, f)), f
