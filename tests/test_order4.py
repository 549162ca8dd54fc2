"""Fourth-order stencils (`_fd_order = 4`): reference src/fdtd/diff2D.jl:47-97, diff3D.jl:60-134,
src/fields.jl:92-671 (array shapes), dirichlet.jl:12-24 (two ghost pairs), GeoPhyInv.jl:90 (npml = 43),
fdtd.jl:316-319 (1/24 in d?I).

CPU part: the oracle's order-4 restatement is pinned by the reference's own accuracy test, which runs at order 4
(test/fdtd/accuracy2D.jl:6,30-82: < 1e-2 against the analytic homogeneous solution), by time reversal and by the
gradient check; shapes are checked against fields.jl.  GPU part: engine (kernels4.cuh) vs oracle, bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_l2

REC_TOL = 1e-5
GRAD_TOL = 1e-4


# --------------------------------------------------------------------------------------------------
# CPU: shapes, oracle invariants
# --------------------------------------------------------------------------------------------------
def test_order4_field_shapes_follow_fields_jl(G, O):
    """fields.jl:92-671 with _fd_order = 4: velocity axes n+3, half-node axes n-3, inner axes n-6."""
    from geophyinv_jl_b200 import engine as E
    lib = E.load_library()
    n = (C.c_int32 * 3)(100, 110, 120)
    out = (C.c_int32 * 3)()
    want3 = {"tauxx": (100, 110, 120), "vx": (100, 110, 123), "vy": (100, 113, 120), "vz": (103, 110, 120),
             "tauxy": (94, 107, 117), "tauxz": (97, 104, 117), "tauyz": (97, 107, 114),
             "dtauxxdx": (94, 104, 117), "dtauyydy": (94, 107, 114), "dtauzzdz": (97, 104, 114), "dvxdx": (100, 110, 120)}
    olib = O.load(np.float32)
    for f, shp in want3.items():
        assert lib.gpi_field_shape_order(3, E.ELASTIC, 4, E.FIELD[f], n, out) == 0
        assert tuple(out) == shp, f
        assert olib.orc_field_shape_order(3, 4, E.FIELD[f], n, out) == 0
        assert tuple(out) == shp, f
        assert G.field_shape(f, (100, 110, 120), 4) == shp
    n2 = (C.c_int32 * 3)(100, 1, 120)
    want2 = {"p": (100, 1, 120), "vx": (100, 1, 123), "vz": (103, 1, 120), "dpdx": (94, 1, 117), "dpdz": (97, 1, 114)}
    for f, shp in want2.items():
        assert lib.gpi_field_shape_order(2, E.ACOUSTIC, 4, E.FIELD[f], n2, out) == 0
        assert tuple(out) == shp, f
    assert lib.gpi_field_shape_order(2, E.ACOUSTIC, 6, E.FIELD["p"], n2, out) != 0      # 6 / 8 are broken upstream
    # grids: vx starts 1.5 cells before the tauii grid, dpdx 1.5 cells after (fields.jl:92-104,169-176)
    gz, gx = G.StepRange(0.0, 10.0, 100), G.StepRange(-50.0, 10.0, 120)
    mz, mx = G.get_mgrid("vx", [gz, gx], 4)
    assert (mz.start, len(mz), mx.start, len(mx)) == (0.0, 100, -65.0, 123)
    mz, mx = G.get_mgrid("dpdx", [gz, gx], 4)
    assert (mz.start, len(mz), mx.start, len(mx)) == (30.0, 94, -35.0, 117)


def test_order4_beats_order2_against_the_analytic_solution(G, O):
    """16 Hz Ricker on the 10 m grid (15 points per wavelength) with a small time step: the second-order scheme is
    dispersive (misfit ~0.7), the fourth-order one passes the reference's own gate of 1e-2 (accuracy2D.jl:39)."""
    from geophyinv_jl_b200.host import gallery
    from test_oracle_invariants import analytic_p_record
    kw = gallery.c1_acou2d_homo(nr=8, nt=1800, dt=0.5e-3, fq=16.0)
    a = analytic_p_record(kw, 2500.0, 2500.0)
    err = {}
    for order in (2, 4):
        po = O.OraclePFdtd64(G.FdtdAcoustic(), **kw, order=order)
        assert po.cfg.npml == 40 + order - 1
        po.update()
        d = po.c.data[0][0].d["p"].astype(np.float64)
        err[order] = np.sum((d[2:] - a[:-2]) ** 2) / np.sum(a[:-2] ** 2)
    print(f"analytic 2-D acoustic at 16 Hz: order 2 misfit {err[2]:.3e}, order 4 misfit {err[4]:.3e}")
    assert err[4] < 1e-2
    assert err[4] < err[2] / 20


def test_order4_time_reversal_and_gradient(G, O):
    """Boundary save / force still closes the time reversal at order 4 (three stored planes cover the four-point
    stencil's reach into the interior, boundary.jl:17-52), and the invK gradient matches finite differences; the rho
    gradient carries upstream's `combine_gmodrho!` shift, which grows to 3-4 cells at order 4 (gradient.jl:53-56)."""
    from geophyinv_jl_b200.host import gallery
    nt, its = 260, [60, 130, 200]
    kw = gallery.c1_acou2d_homo(nz=81, nx=91, nr=8, nt=nt, dt=1.5e-3, fq=14.0, sfield="vz", rfields=("vz",))
    tg = kw["tgrid"]
    pa = O.OraclePFdtd64(G.FdtdAcoustic("forward_save"), **kw, snaps_field="p", tsnaps=[tg.values[i - 1] for i in its], order=4)
    pa.update()
    forw = [s.copy() for s in pa["snaps", 1][0]]
    pa.update_srcwav(pa.c.srcwav, [-1, 0])
    pa.c.attrib_mod.mode = "adjoint"
    pa.c.itsnaps = [nt - i for i in its]
    pa.engine.set_snap_steps(pa.c.itsnaps)
    pa.update(dict(activepw=[1], src_flags=[True, False], rec_flags=[False, False]))
    n = pa.c.npml + 6
    for f, b in zip(forw, pa["snaps", 1][0]):
        assert rel_l2(-b[n:-n, n:-n], f[n:-n, n:-n]) < 1e-9

    kw, true = gallery.c4_fwi2d(nz=40, nx=50, nt=420, nss=2, nr=16, fq=12.0, dt=1.2e-3)
    pt = O.OraclePFdtd64(G.FdtdAcoustic(), **{**kw, "medium": true}, order=4)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd64(G.FdtdAcoustic("forward_save"), **kw, order=4)
    m = pa.get_modelvector().astype(np.float64)
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
    rng = np.random.default_rng(5)
    from scipy.ndimage import gaussian_filter
    dK = gaussian_filter(rng.standard_normal((40, 50)), 4.0).ravel(order="F")
    dR = gaussian_filter(rng.standard_normal((40, 50)), 4.0).ravel(order="F")
    for name, dm, tol in (("invK", np.concatenate([dK, 0 * dR]), 0.05), ("rho", np.concatenate([0 * dK, dR]), 0.4)):
        dm = dm / np.abs(dm).max()
        eps = 2e-3
        fd = (G.lossvalue(m + eps * dm, dobs, pa) - G.lossvalue(m - eps * dm, dobs, pa)) / (2 * eps)
        ad = float(np.dot(g, dm))
        print(f"order 4, d loss / d {name}: adjoint {ad:.6e}  finite-difference {fd:.6e}  ratio {ad / fd:.4f}")
        assert np.sign(ad) == np.sign(fd) and abs(ad / fd - 1) < tol


# --------------------------------------------------------------------------------------------------
# GPU: engine vs oracle
# --------------------------------------------------------------------------------------------------
def _pair(G, O, attrib, kw, **extra):
    return G.SeisForwExpt(attrib(), **kw, order=4, **extra), O.OraclePFdtd(attrib(), **kw, order=4)


def _check(pg, po, fields, label):
    worst, exact = 0.0, True
    for iss in range(len(pg.c.data[0])):
        for f in pg.c.rfields:
            a, b = pg.c.data[0][iss].d[f], po.c.data[0][iss].d[f]
            assert np.isfinite(a).all() and np.abs(b).max() > 0
            worst = max(worst, rel_l2(a, b)); exact &= bool(np.array_equal(a, b))
    nss = len(pg.local)
    B = max(1, min(nss, pg.cfg.shot_batch or (16 if pg.cfg.ndims == 2 else 1)))
    for f in fields:
        a, b = pg.engine.get_field(0, f, (nss - 1) % B), po.engine.get_field(0, f)
        assert a.shape == b.shape
        e = rel_l2(a, b)
        worst = max(worst, e); exact &= bool(np.array_equal(a, b))
    print(f"order 4 {label}: rel-L2 {worst:.3e}, bit-exact {exact}")
    assert worst <= REC_TOL
    return exact


@pytest.mark.gpu
@pytest.mark.parametrize("sfield,rfields", [("p", ("p", "vx")), ("vz", ("vz", "vx", "p"))])
def test_order4_acoustic2d(G, O, sfield, rfields):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c2_acou2d_layered(nz=90, nx=140, nt=400, nss=3, nr=20, fq=15.0, sfield=sfield, rfields=rfields)
    pg, po = _pair(G, O, G.FdtdAcoustic, kw, shot_batch=2)
    pg.update(); po.update()
    _check(pg, po, ["p", "vx", "vz"], f"2-D acoustic {sfield}")


@pytest.mark.gpu
@pytest.mark.parametrize("stressfree,sfield", [(False, "vz"), (True, "vz"), (False, "tauxx")])
def test_order4_elastic2d(G, O, stressfree, sfield):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(nt=350, stressfree=stressfree, sfield=sfield)
    pg, po = _pair(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    _check(pg, po, ["tauxx", "tauzz", "tauxz", "vx", "vz"], f"2-D elastic stressfree={stressfree} source {sfield}")


@pytest.mark.gpu
def test_order4_acoustic3d(G, O):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.acou3d(n=36, nt=150)
    pg, po = _pair(G, O, G.FdtdAcoustic, kw)
    pg.update(); po.update()
    _check(pg, po, ["p", "vx", "vy", "vz"], "3-D acoustic")


@pytest.mark.gpu
@pytest.mark.parametrize("stressfree,faces", [(False, None), (True, None), (False, ("zmax", "xmin", "ymax"))])
def test_order4_elastic3d(G, O, stressfree, faces):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c3_elastic3d(n=30, nt=110, nr=10, fq=25.0, rfields=("vz", "vx", "vy"), stressfree=stressfree)
    if faces is not None:
        kw["pml_faces"] = list(faces)
    pg, po = _pair(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    _check(pg, po, ["tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"], f"3-D elastic stressfree={stressfree} faces={faces}")


@pytest.mark.gpu
def test_order4_fwi_gradient(G, O):
    """forward_save + adjoint + imaging at order 4 (boundary store, save_tp, k_grad2d with the order-4 offsets)."""
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.c4_fwi2d(nz=60, nx=90, nt=400, nss=3, nr=20, fq=10.0)
    pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, shot_batch=2, order=4)
    po = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw, order=4)
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true}, order=4)
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    lg, lo = G.gradient(gg, m, dobs, pg), G.gradient(go, m, dobs, po)
    half = m.size // 2
    eK, eR = rel_l2(gg[:half], go[:half]), rel_l2(gg[half:], go[half:])
    print(f"order 4 FWI gradient: invK rel-L2 {eK:.3e}, rho rel-L2 {eR:.3e}, loss {lg:.6e} vs {lo:.6e}")
    assert abs(lg - lo) <= 1e-5 * abs(lo)
    assert np.abs(go[:half]).max() > 0 and np.abs(go[half:]).max() > 0
    assert eK <= GRAD_TOL and eR <= GRAD_TOL
    for name in ("invK", "rho"):
        assert rel_l2(pg.engine.get_gradient(name), po.engine.get_gradient(name)) <= GRAD_TOL


@pytest.mark.gpu
def test_order4_medium_and_rejections(G, O):
    """Device-side replicate padding lands on the shifted box; FD-Born and z-slabs refuse order 4."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(nz=33, nx=47, nt=4, stressfree=True)
    pg = G.SeisForwExpt(G.FdtdElastic(), **kw, order=4)
    ex = G.padarray(kw["medium"], 43, pg.c.pml_faces)
    for name in pg.c.mparams:
        assert np.array_equal(pg.engine.get_medium(name), ex[name]), name
    with pytest.raises(NotImplementedError):
        G.SeisForwExpt(G.FdtdAcoustic(born=True), **gallery.c1_acou2d_homo(nz=41, nx=41, nt=4, nr=4), order=4)
