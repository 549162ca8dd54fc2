"""Pins the CPU oracle (oracle/fdtd_oracle.c) with the invariants the reference's own tests state
 -- the reference ships no golden vectors for this path and Julia is absent, so
these physics / consistency checks are what stands between the restatement and "unpinned":

  * analytic homogeneous 2-D acoustic solution (test/fdtd/accuracy2D.jl:30-82, src/born/homo.jl:27-31)
  * time reversal with boundary save / force (test/fdtd/backprop.jl:10-47,
    notebooks/boundary-value-problems.jl `test_backprop`)
  * adjoint-state gradient vs finite differences (test/fwi/gradient_accuracy.jl:57-109,
    notebooks/gradient_finitediff_testing.jl)
  * interpolation weights sum to one (src/proj_mat.jl:298-301)
  * Float32 build agrees with the Float64 build to rounding level

CPU only (no GPU, no CUDA library calls).
"""
import numpy as np
import pytest
from scipy.special import hankel2

from conftest import rel_l2


def test_weights_sum_to_one(G):
    """src/proj_mat.jl:298-301"""
    from geophyinv_jl_b200.host.proj import get_proj_matrix
    from geophyinv_jl_b200.host.grids import StepRange
    rng = np.random.default_rng(0)
    grid2 = [StepRange(0.0, 10.0, 30), StepRange(-50.0, 5.0, 40)]
    pts = [[rng.uniform(0, 290), rng.uniform(-50, 145)] for _ in range(20)]
    for f in ("p", "vx", "vz"):
        cp, rv, nz, shp = get_proj_matrix(f, grid2, pts, True)
        for j in range(len(pts)):
            assert abs(nz[cp[j] - 1:cp[j + 1] - 1].sum() - 1) < 1e-5
            assert cp[j + 1] - cp[j] == 4
    grid3 = [StepRange(0.0, 10.0, 12)] * 3
    pts = [list(rng.uniform(0, 110, 3)) for _ in range(10)]
    for f in ("tauxx", "vx", "vy", "vz"):
        cp, rv, nz, shp = get_proj_matrix(f, grid3, pts, True)
        for j in range(len(pts)):
            assert abs(nz[cp[j] - 1:cp[j + 1] - 1].sum() - 1) < 1e-5
            assert cp[j + 1] - cp[j] == 8
            assert rv[cp[j] - 1:cp[j + 1] - 1].max() <= np.prod(shp)


def test_batched_proj_matrix_equals_pointwise():
    """`get_proj_matrix` builds all columns at once; the reference (and `get_proj_matrix_pointwise`) one point at a time
    (proj_mat.jl:229-247).  Same bits, including points outside the grid (clamped fractions) and on nodes."""
    import geophyinv_jl_b200 as G
    from geophyinv_jl_b200.host.proj import get_proj_matrix, get_proj_matrix_pointwise
    rng = np.random.default_rng(0)
    g2 = [G.StepRange(0.0, 10.0, 50), G.StepRange(-100.0, 7.5, 60)]
    pts = [list(p) for p in np.column_stack([rng.uniform(-30, 520, 300), rng.uniform(-130, 380, 300)])]
    pts += [[0.0, -100.0], [490.0, 342.5], [10.0, -92.5], [5.0, -96.25]]
    for f in ("p", "vx", "vz", "dpdx"):
        for order in (2, 4):
            A, B = get_proj_matrix(f, g2, pts, True, order), get_proj_matrix_pointwise(f, g2, pts, True, order)
            assert all(np.array_equal(a, b) for a, b in zip(A[:3], B[:3])) and A[3] == B[3], (f, order)
    g3 = [G.StepRange(0.0, 10.0, 20), G.StepRange(5.0, 10.0, 22), G.StepRange(-10.0, 10.0, 24)]
    pts = [list(p) for p in np.column_stack([rng.uniform(-20, 220, 200), rng.uniform(-20, 240, 200), rng.uniform(-30, 250, 200)])]
    for f in ("tauxx", "vx", "vy", "vz", "tauxy"):
        for swap in (True, False):
            A, B = get_proj_matrix(f, g3, pts, swap, 2), get_proj_matrix_pointwise(f, g3, pts, swap, 2)
            assert all(np.array_equal(a, b) for a, b in zip(A[:3], B[:3])), (f, swap)


def analytic_p_record(kw, medium_vp, medium_rho):
    """Frequency-domain homogeneous solution of  p_tt/c^2 - lap p = rho q'(t) delta(x):
    P(w) = rho (i w) Q(w) (-i/4) H0^(2)(k r), with point-source strength Q = wavelet * dz * dx because the
    engine adds wavelet*dt*K to the pressure of one cell (area dz*dx) per step (source.jl:61-75,160-163).
    Same construction as the reference's analytic modeller (src/born/core.jl:77-174) with
    G0 of src/born/homo.jl:27-31 (which omits the i w factor because its old FD source was time-integrated)."""
    tgrid, ageom, srcwav = kw["tgrid"], kw["ageom"][0], kw["srcwav"][0]
    nt, dt = len(tgrid), tgrid.step
    np2 = int(2 ** np.ceil(np.log2(2 * nt)))
    f = np.fft.fftfreq(np2, dt)
    dz, dx = kw["medium"].grid[0].step, kw["medium"].grid[1].step
    out = np.zeros((nt, ageom.nr))
    s = np.zeros(np2); s[:nt] = srcwav.d["p"][:, 0]
    S = np.fft.fft(s)
    for ir in range(ageom.nr):
        r = np.hypot(ageom.r["z"][ir] - ageom.s["z"][0], ageom.r["x"][ir] - ageom.s["x"][0])
        w = 2 * np.pi * np.abs(f)
        Gf = np.zeros(np2, complex)
        pos = f > 0
        k = w / medium_vp
        term = np.zeros(np2, complex)
        nzf = w > 0
        term[nzf] = medium_rho * (1j * w[nzf]) * (-0.25j) * hankel2(0, k[nzf] * r)
        Gf[pos] = term[pos]
        neg = f < 0
        Gf[neg] = np.conj(term[neg])
        out[:, ir] = np.real(np.fft.ifft(Gf * S))[:nt] * (dz * dx)
    return out


def test_analytic_acoustic2d(G, O):
    """FD records of a :p source recorded as :p against the analytic homogeneous solution.
    Gate: the reference's own < 1e-2 normalised least-squares misfit (accuracy2D.jl:39), with NO free
    scale factor (amplitude is part of the check)."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nr=16, nt=900, dt=1e-3, fq=8.0)
    po = O.OraclePFdtd64(G.FdtdAcoustic(), **kw)
    po.update()
    d = po.c.data[0][0].d["p"].astype(np.float64)
    # alignment: half a sample by construction (record!(p) samples the field at the START of step it, propagate.jl:177, and wavelet
    # sample it acts over the step that ends there, propagate.jl:223; see test_analytic_acoustic3d) plus the delay that the
    # second-order numerical dispersion accumulates over the 160 cells between source and receivers: two samples in total
    a = analytic_p_record(kw, 2500.0, 2500.0)
    d_al, a_al = d[2:, :], a[:-2, :]
    err = np.sum((d_al - a_al) ** 2) / np.sum(a_al ** 2)
    print(f"analytic 2-D acoustic: normalised squared misfit {err:.3e}")
    assert err < 1e-2


def _interior(a, n=41 + 6):
    return a[n:-n, n:-n]


@pytest.mark.parametrize("cls_name,tol", [("OraclePFdtd64", 1e-9), ("OraclePFdtd", 2e-4)])
def test_time_reversal_boundary_save_force(G, O, cls_name, tol):
    """forward_save stores 3+3 boundary planes per step and the final state; the adjoint-mode run of
    pw 1 alone (source type -1) must retrace the forward field backwards inside the PML-free interior:
    p_back(it') = -p_forw(nt - it')  (boundary.jl:17-306, propagate.jl:150-151,188,229,251-258)."""
    from geophyinv_jl_b200.host import gallery
    nt = 260
    its = [60, 130, 200]
    kw = gallery.c1_acou2d_homo(nz=81, nx=91, nr=8, nt=nt, dt=1.5e-3, fq=14.0, sfield="vz", rfields=("vz",))
    tg = kw["tgrid"]
    cls = getattr(O, cls_name)
    pa = cls(G.FdtdAcoustic("forward_save"), **kw, snaps_field="p", tsnaps=[tg.values[i - 1] for i in its])
    assert pa.c.itsnaps == its
    pa.update()
    forw = [s.copy() for s in pa["snaps", 1][0]]
    assert max(np.abs(f).max() for f in forw) > 0
    # back-propagate pw 1 only
    pa.update_srcwav(pa.c.srcwav, [-1, 0])
    pa.c.attrib_mod.mode = "adjoint"
    # snapshots in the adjoint run at it' = nt - it
    pa.c.itsnaps = [nt - i for i in its]
    pa.engine.set_snap_steps(pa.c.itsnaps)
    pa.update(dict(activepw=[1], src_flags=[True, False], rec_flags=[False, False]))
    back = pa["snaps", 1][0]
    for f, b, it in zip(forw, back, its):
        e = rel_l2(_interior(-b), _interior(f))
        print(f"{cls_name}: time reversal at step {it}: rel-L2 {e:.3e}")
        assert e < tol


def test_f32_matches_f64_to_rounding(G, O):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(nt=300)
    p32 = O.OraclePFdtd(G.FdtdElastic(), **kw); p32.update()
    p64 = O.OraclePFdtd64(G.FdtdElastic(), **kw); p64.update()
    for f in ("vz", "vx"):
        e = rel_l2(p32.c.data[0][0].d[f], p64.c.data[0][0].d[f])
        print(f"f32 vs f64 oracle, 2-D elastic {f}: {e:.3e}")
        assert e < 5e-5


def test_gradient_vs_finite_differences(G, O):
    """Adjoint-state gradient of the L2 misfit w.r.t. the log-parameterised model (func_grad.jl:11-49)
    against central finite differences of `lossvalue` (notebooks/gradient_finitediff_testing.jl), Float64
    oracle.  The FD imaging condition is a discretise-then-approximate gradient (gradient.jl:17-56), so
    the agreement is to a few per cent in direction, not to rounding."""
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.c4_fwi2d(nz=40, nx=50, nt=420, nss=2, nr=16, fq=12.0, dt=1.2e-3)
    pt = O.OraclePFdtd64(G.FdtdAcoustic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    pa = O.OraclePFdtd64(G.FdtdAcoustic("forward_save"), **kw)
    m = pa.get_modelvector().astype(np.float64)
    g = np.zeros_like(m)
    G.gradient(g, m, dobs, pa)
    assert np.abs(g).max() > 0
    # directional derivative along a smooth random direction, both parameters
    rng = np.random.default_rng(5)
    nzx = (40, 50)
    half = m.size // 2
    from scipy.ndimage import gaussian_filter
    dK = gaussian_filter(rng.standard_normal(nzx), 4.0).ravel(order="F")
    dR = gaussian_filter(rng.standard_normal(nzx), 4.0).ravel(order="F")
    for name, dm in (("invK", np.concatenate([dK, 0 * dR])), ("rho", np.concatenate([0 * dK, dR]))):
        dm = dm / np.abs(dm).max()
        eps = 2e-3
        lp = G.lossvalue(m + eps * dm, dobs, pa)
        lm = G.lossvalue(m - eps * dm, dobs, pa)
        fd = (lp - lm) / (2 * eps)
        ad = float(np.dot(g.astype(np.float64), dm))
        print(f"d loss / d {name}: adjoint {ad:.6e}  finite-difference {fd:.6e}  ratio {ad / fd:.4f}")
        assert np.sign(ad) == np.sign(fd)
        assert abs(ad / fd - 1) < 0.1


# --------------------------------------------------------------------------------------------------
# FD-Born (src/fdtd/born.jl, test/fwi/born_map.jl): first-order accuracy, linearity, dot test
# --------------------------------------------------------------------------------------------------
def _born_kw(G, sfield="p", rfields=("p", "vz"), nt=700):
    from geophyinv_jl_b200.host import gallery
    return gallery.c1_acou2d_homo(nz=51, nx=57, nt=nt, nr=10, sfield=sfield, rfields=rfields, fq=8.0, dt=1.8e-3)


def test_born_is_first_order_in_the_perturbation(G, O):
    """The scattered data of `update!(pa, medium, medium_pert)` equal full modelling in the perturbed medium minus the
    background up to O(eps^2): the relative error halves with eps (pins sign and scaling of the scattering sources)."""
    errs = []
    for eps in (0.01, 0.005):
        kw = _born_kw(G)
        m0 = kw["medium"]
        mp = m0.copy()
        mp.vp[20:30, 24:34] *= np.float32(1 + eps)
        mp.rho[20:30, 24:34] *= np.float32(1 + eps)
        pb = O.OraclePFdtd64(G.FdtdAcoustic(born=True), **kw)
        G.update(pb, m0, mp)
        pb.update()
        p0 = O.OraclePFdtd64(G.FdtdAcoustic(), **kw); p0.update()
        p1 = O.OraclePFdtd64(G.FdtdAcoustic(), **{**kw, "medium": mp}); p1.update()
        for f in ("p", "vz"):
            dd = p1.c.data[0][0].d[f].astype(np.float64) - p0.c.data[0][0].d[f].astype(np.float64)
            db = pb.c.data[1][0].d[f].astype(np.float64)
            assert np.linalg.norm(dd) > 0
            errs.append(np.linalg.norm(db - dd) / np.linalg.norm(dd))
    assert max(errs[:2]) < 0.1 and max(errs[2:]) < 0.05
    assert errs[2] < 0.6 * errs[0] and errs[3] < 0.6 * errs[1]


def test_born_linear_map_linearity_and_dot_test(G, O):
    """test/fwi/born_map.jl: F(x1 + x2) == F x1 + F x2 and <y, F x> == <x, F' y> (rtol 1e-5 upstream)."""
    kw = _born_kw(G, sfield="vz", rfields=("vz",), nt=500)
    rng = np.random.default_rng(5)
    m = kw["medium"]
    m.vp *= (1 + 0.03 * rng.standard_normal(m.vp.shape)).astype(np.float32)
    m.rho *= (1 + 0.03 * rng.standard_normal(m.rho.shape)).astype(np.float32)
    pa = O.OraclePFdtd64(G.FdtdAcoustic("forward_save", born=True), **kw)
    F = G.LinearMap(pa)
    n = pa.c.gradients["invK"].shape

    def randm():
        a = np.zeros(n, np.float32, order="F"); b = np.zeros(n, np.float32, order="F")
        inner = (slice(G.NPML + 9, n[0] - G.NPML - 9), slice(G.NPML + 9, n[1] - G.NPML - 9))     # away from sources / receivers, see LinearMap
        a[inner] = rng.standard_normal(a[inner].shape) * 1e-12          # invK ~ 6e-11
        b[inner] = rng.standard_normal(b[inner].shape) * 25.0           # rho ~ 2500
        return np.concatenate([a.ravel(order="F"), b.ravel(order="F")])

    x1, x2 = randm(), randm()
    d1, d2, d12 = (F @ x1).astype(np.float64), (F @ x2).astype(np.float64), (F @ (x1 + x2)).astype(np.float64)
    assert np.sum(d12 ** 2) > 0
    assert np.sum((d12 - (d1 + d2)) ** 2) / np.sum(d12 ** 2) < 1e-12      # records are stored in Float32
    y = rng.standard_normal(F.shape[0]).astype(np.float32)
    a = float(np.dot(y.astype(np.float64), d1))
    b = float(np.dot(x1.astype(np.float64), (F.T @ y).astype(np.float64)))
    assert abs(a - b) <= 1e-5 * abs(a), (a, b)
    # the reference's own imaging (one-cell shift in combine_gmodrho!, opposite sign) is not the transpose
    g_ref = G.adjoint_map(np.zeros(F.shape[1], np.float32), y, pa, exact=False)
    b_ref = -float(np.dot(x1.astype(np.float64), g_ref.astype(np.float64)))
    assert abs(a - b_ref) > 1e-3 * abs(a)


def test_check_stability_mirrors_the_courant_criterion(G, O):
    """`check_stability` (stability.jl:6-58): dt above epsilon * min(ds) / vmax is flagged, elastic media use
    sqrt(vp^2 + vs^2) (Virieux 1986) and the vs-based wavelength; nothing is flagged for a time step below the limit."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nz=41, nx=41, nt=20, nr=4, dt=2e-3)
    ds = kw["medium"].grid[0].step
    ok = O.OraclePFdtd(G.FdtdAcoustic(), **kw).stability
    bad = O.OraclePFdtd(G.FdtdAcoustic(), **gallery.c1_acou2d_homo(nz=41, nx=41, nt=20, nr=4, dt=2e-2)).stability
    assert abs(ok["dt_recommended"] - (1 / np.sqrt(2)) * ds / 2750.0) < 1e-9            # bounds = max + 0.1 * mean (media.jl:24-26)
    assert not any("time sampling" in w for w in ok["warnings"])
    assert any("time sampling" in w for w in bad["warnings"])
    el = O.OraclePFdtd(G.FdtdElastic(), **gallery.elastic2d(nz=40, nx=50, nt=10, nr=4)).stability
    assert el["dt_recommended"] < (1 / np.sqrt(2)) * 10.0 / 3600.0                       # vmax = sqrt(vp^2 + vs^2) > the largest vp


@pytest.mark.parametrize("order", [2, 4])
def test_reciprocity_of_the_pressure_green_function(G, O, order):
    """Source-receiver reciprocity: in a heterogeneous medium with CPML, a :p source at node A recorded as :p at node B equals
    the swapped experiment.  The source term enters as wavelet * dt * K (source.jl:61-75,160-163), i.e. as a volume-injection rate
    in (1/K) dp/dt = div v + s, whose pressure response is symmetric.  A property of the discrete operator, independent of any
    reference data: it would break with a misplaced staggered node, a wrong averaging of rho or an asymmetric CPML term."""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.data import AGeomss, make_srcwav
    kw = gallery.c2_acou2d_layered(nz=70, nx=90, nt=420, nss=1, nr=4, fq=12.0)
    med, grid, tg = kw["medium"], kw["medium"].grid, kw["tgrid"]
    A, B = (12, 20), (55, 71)                                   # (iz, ix) nodes in different layers
    pos = lambda n: {"z": [grid[0][n[0]]], "x": [grid[1][n[1]]]}
    wav = kw["srcwav"][0].d["p"][:, 0]
    recs = []
    for s, r in ((A, B), (B, A)):
        ag = [AGeomss(pos(s), pos(r))]
        sw = make_srcwav(tg, ag, ["p"], wav)
        pa = O.OraclePFdtd64(G.FdtdAcoustic(), **{**kw, "ageom": ag, "srcwav": sw, "rfields": ["p"]}, order=order)
        pa.update()
        recs.append(pa.c.data[0][0].d["p"][:, 0].astype(np.float64))
    assert np.abs(recs[0]).max() > 0
    e = rel_l2(recs[0], recs[1])
    print(f"reciprocity p_AB vs p_BA (K_A / K_B = {float(med.vp[A]) ** 2 * float(med.rho[A]) / (float(med.vp[B]) ** 2 * float(med.rho[B])):.3f}): rel-L2 {e:.3e}")
    assert e < 1e-6


def test_reciprocity_of_the_elastic_velocity_response(G, O):
    """Elastic reciprocity with a free surface: a :vz force at A recorded as :vx at B equals a :vx force at B recorded as :vz at A
    (and vz-vz likewise).  Sources sit on nodes of their own staggered grids, so that spraying and sampling use one tap; the
    velocity source enters as wavelet * dt / av(rho) (source.jl:166-177), i.e. as a force density in rho dv/dt = ... + f.
    Exercises every elastic operator at once: the two shear-modulus averages, lambda / M, the rho averages, the free surface."""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.data import AGeomss, make_srcwav
    kw = gallery.elastic2d(nz=60, nx=76, nt=360, nr=4, nss=1, stressfree=True)
    grid, tg = kw["medium"].grid, kw["tgrid"]
    gvz, gvx = G.get_mgrid("vz", grid), G.get_mgrid("vx", grid)
    wav = kw["srcwav"][0].d["vz"][:, 0]
    A, B = (14, 22), (41, 57)
    node = lambda g, n: {"z": [g[0][n[0]]], "x": [g[1][n[1]]]}

    def run(sfield, spos, rfield, rpos):
        ag = [AGeomss(spos, rpos)]
        sw = make_srcwav(tg, ag, [sfield], wav)
        pa = O.OraclePFdtd64(G.FdtdElastic(), **{**kw, "ageom": ag, "srcwav": sw, "rfields": [rfield]})
        pa.update()
        return pa.c.data[0][0].d[rfield][:, 0].astype(np.float64)

    zx_ab = run("vz", node(gvz, A), "vx", node(gvx, B))
    zx_ba = run("vx", node(gvx, B), "vz", node(gvz, A))
    zz_ab = run("vz", node(gvz, A), "vz", node(gvz, B))
    zz_ba = run("vz", node(gvz, B), "vz", node(gvz, A))
    assert np.abs(zx_ab).max() > 0 and np.abs(zz_ab).max() > 0
    e1, e2 = rel_l2(zx_ab, zx_ba), rel_l2(zz_ab, zz_ba)
    print(f"elastic reciprocity: vz->vx vs vx->vz rel-L2 {e1:.3e}, vz->vz swapped {e2:.3e}")
    assert e1 < 1e-6 and e2 < 1e-6


def test_reciprocity_3d_elastic(G, O):
    """The same property for the 3-D elastic operators behind the roofline case (no analytic solution is at hand for them):
    :vz force at A recorded as :vy at B  ==  :vy force at B recorded as :vz at A, heterogeneous medium, CPML on all faces.
    (`upstream_3d_swap=False`: the z/x swap of the 3-D weights, SURVEY App. C.1, is host-side geometry and not under test here.)"""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.data import AGeomss, make_srcwav
    kw = gallery.c3_elastic3d(n=20, nt=170, nr=4, fq=25.0)
    grid, tg = kw["medium"].grid, kw["tgrid"]
    gvz, gvy = G.get_mgrid("vz", grid), G.get_mgrid("vy", grid)
    wav = kw["srcwav"][0].d["vz"][:, 0]
    A, B = (5, 6, 4), (13, 12, 15)
    node = lambda g, n: {"z": [g[0][n[0]]], "y": [g[1][n[1]]], "x": [g[2][n[2]]]}

    def run(sfield, spos, rfield, rpos):
        ag = [AGeomss(spos, rpos)]
        sw = make_srcwav(tg, ag, [sfield], wav)
        pa = O.OraclePFdtd64(G.FdtdElastic(), **{**kw, "ageom": ag, "srcwav": sw, "rfields": [rfield]}, upstream_3d_swap=False)
        pa.update()
        return pa.c.data[0][0].d[rfield][:, 0].astype(np.float64)

    ab = run("vz", node(gvz, A), "vy", node(gvy, B))
    ba = run("vy", node(gvy, B), "vz", node(gvz, A))
    assert np.abs(ab).max() > 0
    e = rel_l2(ab, ba)
    print(f"3-D elastic reciprocity vz->vy vs vy->vz: rel-L2 {e:.3e}")
    assert e < 1e-6


def test_update_ageom_recs_round_trip_through_the_upload_cache(G, O):
    """`update!(pa, ageom, Recs)` (ageom.jl:58-98): moved receivers change the records, restoring them restores the records bit for
    bit -- the host keeps the uploaded interpolation matrices in a cache keyed by the points, which must notice both moves."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nz=41, nx=41, nt=400, nr=4)
    pa = O.OraclePFdtd(G.FdtdAcoustic(), **kw)
    pa.update(); d0 = pa.c.data[0][0].d["p"].copy()
    ag = kw["ageom"]
    moved = [G.AGeomss(ag[0].s, {k: v + 30.0 for k, v in ag[0].r.items()})]
    G.update(pa, moved, G.Recs); pa.update(); d1 = pa.c.data[0][0].d["p"].copy()
    G.update(pa, ag, "recs"); pa.update(); d2 = pa.c.data[0][0].d["p"].copy()
    assert np.abs(d0).max() > 0 and not np.array_equal(d0, d1) and np.array_equal(d0, d2)


@pytest.mark.parametrize("order", [2, 4])
def test_analytic_acoustic3d(G, O, order):
    """3-D homogeneous Green's function: a :p source (wavelet * dt * K added to one cell of volume dV per step, source.jl:61-75) is a
    volume-injection rate q(t) dV, so  p(r, t) = rho * dV * q'(t - r/c) / (4 pi r).  Amplitude and timing with NO free parameter:
    the only shift is the half sample of the leapfrog (sample `it` of the wavelet acts over the step that ends at record `it`).
    Receivers 6 to 32 cells away, so numerical dispersion is small and the gate can be ten times tighter than the reference's
    2-D accuracy test (1e-2, test/fdtd/accuracy2D.jl:39): measured 1.5e-4.  Pins the 3-D derivative operators, dt*K and dt/rho
    scaling, the CPML (nothing comes back) and the record weights."""
    from geophyinv_jl_b200.host.data import AGeomss, Medium, make_srcwav, ricker
    n, d, dt, nt, fq, c0, rho0 = 56, 10.0, 1e-3, 400, 8.0, 2500.0, 2000.0
    grid = [G.StepRange(0.0, d, n)] * 3
    medium = Medium(grid, np.full((n, n, n), c0, np.float32), np.full((n, n, n), rho0, np.float32))
    tgrid = G.StepRange(0.0, dt, nt)
    s = [grid[0][20], grid[1][22], grid[2][21]]                                    # on a node of the p grid
    nr = 6
    rec = {"z": np.array([grid[0][20 + 3 * k] for k in range(1, nr + 1)]), "y": np.array([grid[1][22 + 2 * k] for k in range(1, nr + 1)]),
           "x": np.array([grid[2][21 + 4 * k] for k in range(1, nr + 1)])}
    ageom = [AGeomss({"z": [s[0]], "y": [s[1]], "x": [s[2]]}, rec)]
    wav = ricker(fq, tgrid, tpeak=1.5 / fq + 0.01)
    srcwav = make_srcwav(tgrid, ageom, ["p"], wav)
    po = O.OraclePFdtd64(G.FdtdAcoustic(), medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=["p"], upstream_3d_swap=False, order=order)
    po.update()
    dat = po.c.data[0][0].d["p"].astype(np.float64)
    np2 = int(2 ** np.ceil(np.log2(2 * nt)))
    W = np.fft.rfft(np.asarray(wav, np.float64), np2)
    f = np.fft.rfftfreq(np2, dt)
    ana = np.zeros_like(dat)
    for ir in range(nr):
        r = np.sqrt((rec["z"][ir] - s[0]) ** 2 + (rec["y"][ir] - s[1]) ** 2 + (rec["x"][ir] - s[2]) ** 2)
        spec = W * (2j * np.pi * f) * np.exp(-2j * np.pi * f * (r / c0 + 0.5 * dt)) * rho0 * d ** 3 / (4 * np.pi * r)
        ana[:, ir] = np.fft.irfft(spec, np2)[:nt]
    err = np.sum((dat - ana) ** 2) / np.sum(ana ** 2)
    print(f"analytic 3-D acoustic, order {order}: normalised squared misfit {err:.3e}")
    assert err < 1e-3


def stokes_misfits(G, make):
    """3-D elastic full space, point force along z: the Stokes solution (Aki & Richards eq. 4.23 -- near-field term between the P
    and S arrivals, far-field P and S), no free parameter.  The velocity source adds wavelet * dt / rho to one vz node
    (source.jl:166-177), i.e. a force F(t) = wavelet * dV; records are velocities, v = du/dt, taken after the velocity update of
    step `it`, half a sample BEFORE the wavelet's time axis (leapfrog).  Receivers on nodes of their own staggered grids, 9-20 cells
    away in all three directions, :vz and :vx.  Pins amplitude, timing and radiation pattern of the 3-D elastic operators behind the
    roofline case (lambda + 2 mu, lambda, the three shear-modulus averages, dt / rho) against something that is not this code."""
    from geophyinv_jl_b200.host.data import AGeomss, Medium, make_srcwav, ricker
    n, d, dt, nt, fq = 46, 10.0, 1.5e-3, 380, 6.0
    al, be, rho = 3000.0, 1700.0, 2300.0
    grid = [G.StepRange(0.0, d, n)] * 3
    medium = Medium(grid, np.full((n, n, n), al, np.float32), np.full((n, n, n), rho, np.float32), np.full((n, n, n), be, np.float32))
    tgrid = G.StepRange(0.0, dt, nt)
    gvz, gvx = G.get_mgrid("vz", grid), G.get_mgrid("vx", grid)
    S = (16, 18, 17)
    spos = [gvz[0][S[0]], gvz[1][S[1]], gvz[2][S[2]]]
    offs = [(6, 4, 8), (10, 9, 5), (14, 6, 12), (3, 12, 10)]
    wav = ricker(fq, tgrid, tpeak=1.5 / fq + 0.01) * 1e6
    np2 = int(2 ** np.ceil(np.log2(2 * nt)))
    F = np.fft.rfft(np.asarray(wav, np.float64), np2) * d ** 3                    # point force = wavelet * dV
    w = 2 * np.pi * np.fft.rfftfreq(np2, dt)

    def stokes_velocity(ri, comp):
        r = np.linalg.norm(ri); gam = ri / r
        gi, gj, dij = gam[comp], gam[0], 1.0 if comp == 0 else 0.0               # force along z = axis 0 of (z, y, x)
        ww = w[1:]
        anti = lambda t: np.exp(-1j * ww * t) * (1j * t / ww + 1.0 / ww ** 2)     # antiderivative of tau * exp(-i w tau)
        Gw = np.zeros(w.size, complex)
        Gw[1:] = ((3 * gi * gj - dij) / r ** 3 * (anti(r / be) - anti(r / al)) + gi * gj * np.exp(-1j * ww * r / al) / (al ** 2 * r)
                  - (gi * gj - dij) * np.exp(-1j * ww * r / be) / (be ** 2 * r)) / (4 * np.pi * rho)
        return np.fft.irfft(1j * w * Gw * F * np.exp(1j * w * 0.5 * dt), np2)[:nt]

    # one run records both fields at the union of the two receiver sets (nodes of the :vz grid, then nodes of the :vx grid); each field is
    # compared at its own nodes only, where the interpolation weights are a single 1
    sets = {rf: {k: np.array([g[q][S[q] + o[q]] for o in offs]) for q, k in enumerate(("z", "y", "x"))} for rf, g in (("vz", gvz), ("vx", gvx))}
    rec = {k: np.concatenate([sets["vz"][k], sets["vx"][k]]) for k in ("z", "y", "x")}
    ageom = [AGeomss({"z": [spos[0]], "y": [spos[1]], "x": [spos[2]]}, rec)]
    srcwav = make_srcwav(tgrid, ageom, ["vz"], wav)
    po = make(G.FdtdElastic(), medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=["vz", "vx"], upstream_3d_swap=False)
    po.update()
    out = {}
    for rf, comp, first in (("vz", 0, 0), ("vx", 2, len(offs))):
        dat = po.c.data[0][0].d[rf].astype(np.float64)[:, first:first + len(offs)]
        r = sets[rf]
        ana = np.stack([stokes_velocity(np.array([r["z"][ir] - spos[0], r["y"][ir] - spos[1], r["x"][ir] - spos[2]]), comp)
                        for ir in range(len(offs))], axis=1)
        out[rf] = float(np.sum((dat - ana) ** 2) / np.sum(ana ** 2))
    return out


def test_analytic_elastic3d_stokes(G, O):
    for rf, err in stokes_misfits(G, O.OraclePFdtd64).items():
        print(f"Stokes solution, :vz force recorded as :{rf}: normalised squared misfit {err:.3e}")
        assert err < 1e-3




@pytest.mark.parametrize("order", [2, 4])
def test_analytic_elastic2d_line_force(G, O, order):
    """2-D elastic (plane strain) full space, line force along z:  G_ij = g_s delta_ij / (rho beta^2) + d_i d_j (g_s - g_p) / (rho w^2)
    with the 2-D Helmholtz Green's functions g_c = (-i/4) H0^(2)(w r / c).  No free parameter; same conventions as the Stokes test
    (force = wavelet * dA, velocities half a sample ahead).  Pins the 2-D elastic operators (tauxz on the (H, H) grid, @av(invmu))."""
    from scipy.special import hankel2
    from geophyinv_jl_b200.host.data import AGeomss, Medium, make_srcwav, ricker
    nz, nx, d, dt, nt, fq = 90, 100, 10.0, 1.5e-3, 520, 6.0
    al, be, rho = 3000.0, 1700.0, 2300.0
    grid = [G.StepRange(0.0, d, nz), G.StepRange(0.0, d, nx)]
    medium = Medium(grid, np.full((nz, nx), al, np.float32), np.full((nz, nx), rho, np.float32), np.full((nz, nx), be, np.float32))
    tgrid = G.StepRange(0.0, dt, nt)
    gvz, gvx = G.get_mgrid("vz", grid), G.get_mgrid("vx", grid)
    S = (30, 32)
    spos = [gvz[0][S[0]], gvz[1][S[1]]]
    offs = [(12, 9), (20, 14), (7, 25), (28, 30)]
    wav = ricker(fq, tgrid, tpeak=1.5 / fq + 0.01) * 1e6
    np2 = int(2 ** np.ceil(np.log2(2 * nt)))
    F = np.fft.rfft(np.asarray(wav, np.float64), np2) * d ** 2
    w = 2 * np.pi * np.fft.rfftfreq(np2, dt)

    def velocity(ri, comp):
        r = np.linalg.norm(ri); gam = ri / r
        gi, gj, dij = gam[comp], gam[0], 1.0 if comp == 0 else 0.0
        ww = w[1:]

        def hess(c):            # d_i d_j of g_c(r) = (-i/4) H0^(2)(k r):  g'' gi gj + g' (dij - gi gj) / r
            k = ww / c
            h0, h1 = hankel2(0, k * r), hankel2(1, k * r)
            g1 = (-0.25j) * (-k * h1)
            g2 = (-0.25j) * (-k * k) * (h0 - h1 / (k * r))
            return g2 * gi * gj + g1 * (dij - gi * gj) / r
        Gw = np.zeros(w.size, complex)
        Gw[1:] = (-0.25j) * hankel2(0, ww / be * r) * dij / (rho * be ** 2) + (hess(be) - hess(al)) / (rho * ww ** 2)
        return np.fft.irfft(1j * w * Gw * F * np.exp(1j * w * 0.5 * dt), np2)[:nt]

    for rf, g, comp in (("vz", gvz, 0), ("vx", gvx, 1)):
        rec = {k: np.array([g[q][S[q] + o[q]] for o in offs]) for q, k in enumerate(("z", "x"))}
        ageom = [AGeomss({"z": [spos[0]], "x": [spos[1]]}, rec)]
        srcwav = make_srcwav(tgrid, ageom, ["vz"], wav)
        po = O.OraclePFdtd64(G.FdtdElastic(), medium=medium, tgrid=tgrid, ageom=ageom, srcwav=srcwav, rfields=[rf], order=order)
        po.update()
        dat = po.c.data[0][0].d[rf].astype(np.float64)
        ana = np.stack([velocity(np.array([rec["z"][ir] - spos[0], rec["x"][ir] - spos[1]]), comp) for ir in range(len(offs))], axis=1)
        err = np.sum((dat - ana) ** 2) / np.sum(ana ** 2)
        print(f"2-D elastic line force, order {order}, :vz force recorded as :{rf}: normalised squared misfit {err:.3e}")
        assert err < 2e-3


def test_illumination_is_the_stacked_energy_of_the_pressure_snapshots(G, O):
    """`illum_flag` (fdtd.jl:59,73): compute_illum! adds abs2(p) of pw 1 at every time step into a Float64 array, stack_illums! sums the
    interior views over the supersources (fdtd.jl:556-581; the calls are commented out upstream, propagate.jl:114,236).  Checked against
    the same sum formed from a snapshot of :p at EVERY step of the same run (Float32 square, Float64 sum, shot order), bit for bit;
    a second update! starts from zero again (types.jl:171)."""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.fdtd import view_inner
    kw = gallery.c2_acou2d_layered(nz=40, nx=56, nt=120, nss=2, nr=6, fq=20.0)
    tg = kw["tgrid"]
    pa = O.OraclePFdtd(G.FdtdAcoustic(), **kw, illum_flag=True, snaps_field="p", tsnaps=list(tg.values))
    assert pa.c.itsnaps == list(range(1, 121))
    for rep in range(2):
        pa.update()
        want = np.zeros(pa.c.illum_stack.shape, np.float64)
        for shot in pa["snaps", 1]:
            e = np.zeros_like(want)
            for snap in shot:
                inner = view_inner(snap, pa.c.npml, pa.c.pml_faces)
                e += (inner * inner).astype(np.float64)               # Float32 square, Float64 sum
            want += e
        assert want.max() > 0 and pa.c.illum_stack.shape == tuple(len(g) for g in kw["medium"].grid)
        assert np.array_equal(pa["illum"], want), f"update {rep}"
    with pytest.raises(Exception):
        O.OraclePFdtd(G.FdtdElastic(), **gallery.elastic2d(nt=5), illum_flag=True)
