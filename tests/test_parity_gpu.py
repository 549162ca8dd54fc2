"""Parity of the CUDA engine against the CPU restatement (oracle/) through the C ABI.

Gates (BASELINE.json north_star): receiver records rel-L2 <= 1e-5, FWI gradients <= 1e-4 in Float32.
Because the kernels reproduce the reference's Float32 operation order without FMA contraction, the
expected result is bit-identical records; the tests assert the gate and report exactness.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

REC_TOL = 1e-5      # north_star: relative L2 misfit on receiver records
GRAD_TOL = 1e-4     # north_star: relative L2 misfit on FWI gradients


def both(G, O, attrib_factory, kw, **extra):
    pg = G.SeisForwExpt(attrib_factory(), **kw, **extra)
    po = O.OraclePFdtd(attrib_factory(), **kw)
    return pg, po


def compare_records(pg, po, ipw=0):
    worst = 0.0
    exact = True
    for iss in range(len(pg.c.data[ipw])):
        for f in pg.c.rfields:
            a, b = pg.c.data[ipw][iss].d[f], po.c.data[ipw][iss].d[f]
            assert np.isfinite(a).all()
            assert np.abs(b).max() > 0, "oracle record is empty; the case does not test anything"
            worst = max(worst, rel_l2(a, b))
            exact &= bool(np.array_equal(a, b))
    return worst, exact


def compare_fields(pg, po, fields, ipw=0):
    """Final wavefields of the LAST supersource (the oracle propagates shots serially in one slot; the
    engine keeps a batch of shots resident, so the last shot sits in the last used batch slot)."""
    worst = 0.0
    nss = len(pg.local)
    B = max(1, min(nss, pg.cfg.shot_batch or (16 if pg.cfg.ndims == 2 else 1)))
    slot = (nss - 1) % B
    for f in fields:
        a, b = pg.engine.get_field(ipw, f, slot), po.engine.get_field(ipw, f)
        if np.abs(b).max() == 0:
            assert np.abs(a).max() == 0
            continue
        worst = max(worst, rel_l2(a, b))
    return worst


@pytest.mark.parametrize("sfield,rfields", [("p", ("p",)), ("vz", ("vz", "vx", "p")), ("vx", ("vx",))])
def test_c1_acoustic2d_records(G, O, sfield, rfields):
    """BASELINE config 1: 2-D acoustic 201x201, Ricker, 64 receivers, 1000 steps, CPML."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(sfield=sfield, rfields=rfields)
    pg, po = both(G, O, G.FdtdAcoustic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"C1 {sfield}->{rfields}: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["p", "vx", "vz"]) <= REC_TOL


def test_acoustic2d_multishot_batches(G, O):
    """Several supersources: batched on the GPU (shot_batch=3 -> ragged last batch), serial in the oracle."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c2_acou2d_layered(nz=90, nx=140, nt=300, nss=5, nr=20, fq=15.0)
    pg, po = both(G, O, G.FdtdAcoustic, kw, shot_batch=3)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"2-D acoustic 5 shots: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL


@pytest.mark.parametrize("stressfree", [False, True])
def test_elastic2d_records(G, O, stressfree):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(stressfree=stressfree)
    pg, po = both(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"2-D elastic stressfree={stressfree}: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["tauxx", "tauzz", "tauxz", "vx", "vz"]) <= REC_TOL


def test_elastic2d_stress_source(G, O):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(sfield="tauxx", rfields=("vz", "vx"))
    pg, po = both(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"2-D elastic explosive source: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL


def test_acoustic3d_records(G, O):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.acou3d()
    pg, po = both(G, O, G.FdtdAcoustic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"3-D acoustic: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["p", "vx", "vy", "vz"]) <= REC_TOL


@pytest.mark.parametrize("stressfree", [False, True])
def test_c3_elastic3d_reduced(G, O, stressfree):
    """BASELINE config 3 down-sized (40^3 + CPML = 122^3, 150 steps): the roofline kernel's parity case."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c3_elastic3d(n=40, nt=150, nr=12, fq=25.0, rfields=("vz", "vx", "vy"), stressfree=stressfree)
    pg, po = both(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"3-D elastic stressfree={stressfree}: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"]) <= REC_TOL


@pytest.mark.parametrize("n", [37, 38, 39])
def test_elastic3d_ragged_z(G, O, n):
    """Every residue of the extended z extent modulo the four-cell thread group (tail / ghost handling of
    the vector kernels), three recorded components."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c3_elastic3d(n=n, nt=130, nr=10, fq=30.0, rfields=("vz", "vx", "vy"))
    pg, po = both(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"3-D elastic n={n}: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"]) <= REC_TOL


@pytest.mark.parametrize("faces", [("zmax", "xmin", "ymax"), ("zmin", "ymin", "ymax"), ()])
def test_elastic3d_partial_pml_faces(G, O, faces, monkeypatch):
    """CPML on a subset of faces (or none): slab logic per face, untouched faces stay untouched.  GPI_TMA3=2 keeps the
    TMA-pipelined kernels on these short z extents (by default a grid that fills < 75 % of its 128-cell z tiles runs
    the register-staged kernels), so that two-chunk tiles with and without z slabs stay covered."""
    from geophyinv_jl_b200.host import gallery
    monkeypatch.setenv("GPI_TMA3", "2")
    kw = gallery.c3_elastic3d(n=60, nt=230, nr=10, fq=30.0, rfields=("vz", "vy"))
    kw["pml_faces"] = list(faces)
    kw["rigid_faces"] = list(faces)
    pg, po = both(G, O, G.FdtdElastic, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"3-D elastic faces={faces}: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, ["tauxx", "tauyy", "tauzz", "tauxy", "tauxz", "tauyz", "vx", "vy", "vz"]) <= REC_TOL


def test_vector_and_scalar_3d_kernels_agree(G, monkeypatch):
    """The float4-per-thread kernels (kernels3d.cuh) and the scalar reference-order kernels produce the
    same bits: same association order, no FMA contraction."""
    from geophyinv_jl_b200.host import gallery
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("GPI_SCALAR3D", flag)
        for name, attrib, kw in (("el", G.FdtdElastic, gallery.c3_elastic3d(n=38, nt=120, nr=8, fq=30.0, rfields=("vz", "vx"))),
                                 ("ac", G.FdtdAcoustic, gallery.acou3d(n=45, nt=200))):
            pg = G.SeisForwExpt(attrib(), **kw)
            pg.update()
            fields = ["vx", "vy", "vz"] + (["tauxx", "tauxy", "tauyz"] if name == "el" else ["p"])
            out[(name, flag)] = [pg.c.data[0][0].d[f].copy() for f in pg.c.rfields] + [pg.engine.get_field(0, f) for f in fields]
    for name in ("el", "ac"):
        for a, b in zip(out[(name, "0")], out[(name, "1")]):
            assert np.abs(b).max() > 0
            assert np.array_equal(a, b), f"{name}: vector and scalar kernels differ"


def test_tma_and_register_staged_kernels_agree_at_full_c3_size(G, monkeypatch):
    """BASELINE config 3 at its full size (338^3 extended cells, 60 time steps): the TMA-pipelined kernels
    (kernels3t.cuh), the register-staged float4 kernels (kernels3d.cuh) produce the same bits in every wavefield
    and record; and doubling the wavelet doubles the wavefield exactly wherever it is above the subnormal range
    (scaling by two is exact in binary floating point, so linearity in the source is a bit-level property of the path)."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c3_elastic3d(n=256, nt=60, nr=16, rfields=("vz", "vx"))
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("GPI_TMA3", flag)
        pg = G.SeisForwExpt(G.FdtdElastic(), **kw)
        pg.update()
        out[flag] = [pg.c.data[0][0].d[f].copy() for f in pg.c.rfields] + [pg.engine.get_field(0, f) for f in ("vx", "vy", "vz", "tauxx", "tauzz", "tauxy", "tauxz", "tauyz")]
        if flag == "1":
            for s in pg.c.srcwav[0]:
                for f in s.fields:
                    s.d[f] *= np.float32(2)
            pg.update_srcwav(pg.c.srcwav)
            pg.update()
            doubled = [pg.c.data[0][0].d[f].copy() for f in pg.c.rfields] + [pg.engine.get_field(0, f) for f in ("vz", "tauxy")]
        del pg
    for a, b in zip(out["1"], out["0"]):
        assert np.isfinite(a).all()
        assert np.array_equal(a, b), "TMA and register-staged kernels differ"
    assert all(np.abs(a).max() > 0 for a in out["1"][2:])             # the wave has not reached the receivers after 60 steps
    for a, d in zip(out["1"][:2] + [out["1"][4], out["1"][7]], doubled):
        big = np.abs(a) > 1e-20            # the numerical tail far from the source lives in the subnormal range, where rounding is absolute
        assert np.array_equal(d[big], a[big] * np.float32(2)), "the wavefield is not linear in the wavelet"
    assert (np.abs(out["1"][4]) > 1e-20).sum() > 50000       # 79155 on B200: the sphere the wave has reached after 60 steps


def test_tiles_on_a_ragged_z_pitch_match_the_oracle(G, O, monkeypatch):
    """The z pitch of the unified box is a multiple of 8 floats (32-byte sectors), not of whole 128-byte lines: rows of a 12 + 2 x 41 =
    94-node grid are 96 floats apart, those of the 10 + 82 = 92-node one 96 too, 13 + 82 -> 104.  The TMA tiles (forced: GPI_TMA3=2)
    take one ragged 128-cell chunk per row; the last active lane takes its z + 1 neighbour from the staged halo, out-of-range box
    columns read zeros.  Records and final fields bit-identical to the oracle."""
    from geophyinv_jl_b200.host import gallery
    monkeypatch.setenv("GPI_TMA3", "2")
    for n in (12, 13):
        kw = gallery.c3_elastic3d(n=n, nt=90, nr=8, fq=60.0, rfields=("vz", "vx"))
        pg, po = both(G, O, G.FdtdElastic, kw)
        pg.update(); po.update()
        assert pg.engine.kernel_family() == "tma"
        for f in pg.c.rfields:
            a, b = pg.c.data[0][0].d[f], po.c.data[0][0].d[f]
            assert np.abs(b).max() > 0 and np.array_equal(a, b), f
        for f in ("vx", "vy", "vz", "tauxx", "tauzz", "tauxy", "tauxz", "tauyz"):
            assert np.array_equal(pg.engine.get_field(0, f), po.engine.get_field(0, f)), f


@pytest.mark.parametrize("graph", ["1", "2"])
def test_graph_replay_of_the_time_loop_is_bit_identical(G, monkeypatch, graph):
    """2-D forward runs: the launches of a batch's time loop are captured into a CUDA graph and replayed by later runs of the same
    configuration (one graph launch instead of 3 - 4 host launches per time step).  GPI_GRAPH=1 (default): first run launch by launch,
    second captured, third replayed; 2: captured at the first run.  Records and final fields of every run, and of a run after the
    wavelets changed (same pointers, new contents: the descriptors are uploaded outside the graph), equal those of a handle that
    launches kernel by kernel (GPI_GRAPH=0); the launch count it reports is the same."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c2_acou2d_layered(nz=70, nx=90, nt=300, nss=3, nr=10, fq=15.0, rfields=("p", "vz"))
    monkeypatch.setenv("GPI_GRAPH", "0")
    ref = G.SeisForwExpt(G.FdtdAcoustic(), **kw, shot_batch=2)
    n_ref = ref.update()["launches"]
    want = [[r.d[f].copy() for f in ref.c.rfields] for r in ref.c.data[0]]
    monkeypatch.setenv("GPI_GRAPH", graph)
    pg = G.SeisForwExpt(G.FdtdAcoustic(), **kw, shot_batch=2)
    for rep in range(4):
        n = pg.update()["launches"]
        assert n == n_ref, (rep, n, n_ref)
        for a, b in zip(want, [[r.d[f] for f in pg.c.rfields] for r in pg.c.data[0]]):
            for x, y in zip(a, b):
                assert np.abs(x).max() > 0 and np.array_equal(x, y), f"graph run {rep} differs"
    for f in ("p", "vx", "vz"):
        assert np.array_equal(pg.engine.get_field(0, f), ref.engine.get_field(0, f))
    for p_ in (ref, pg):
        for s in p_.c.srcwav[0]:
            for f in s.fields:
                s.d[f] *= np.float32(0.5)
        p_.update_srcwav(p_.c.srcwav)
        p_.update()
    for r0, r1 in zip(ref.c.data[0], pg.c.data[0]):
        for f in ref.c.rfields:
            assert np.array_equal(r0.d[f], r1.d[f]), "graph replay after a wavelet change differs"


def test_dmod_matches_oracle(G, O):
    """update_dmod! (medium.jl:143-221): coefficient arrays agree bit for bit (checked through the
    wavefield after one step with unit fields is overkill; compare the medium round trip instead)."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.elastic2d(nt=5)
    pg, po = both(G, O, G.FdtdElastic, kw)
    for name in ("invlambda", "invmu", "rho"):
        assert np.array_equal(pg.engine.get_medium(name), po.engine.get_medium(name))


def test_medium_padded_on_device(G, O):
    """update!(pa, medium) sends the un-extended arrays; the replicate padding of padarray! (media.jl:260-275)
    runs on the device.  Compare with the host-side padarray for full and partial PML faces, 2-D and 3-D."""
    from geophyinv_jl_b200.host import gallery
    cases = [(G.FdtdElastic, gallery.c3_elastic3d(n=22, nt=4, nr=4, fq=60.0, stressfree=True)),
             (G.FdtdElastic, gallery.c3_elastic3d(n=21, nt=4, nr=4, fq=60.0)),
             (G.FdtdAcoustic, gallery.c1_acou2d_homo(nz=37, nx=52, nt=4, nr=4)),
             (G.FdtdElastic, gallery.elastic2d(nz=33, nx=47, nt=4, stressfree=True))]
    for cls, kw in cases:
        pg = G.SeisForwExpt(cls(), **kw)
        ex = G.padarray(kw["medium"], G.NPML, pg.c.pml_faces)
        for name in pg.c.mparams:
            assert np.array_equal(pg.engine.get_medium(name), ex[name]), name
        # a second update with a different medium overwrites every padded cell
        m2 = kw["medium"].copy(); m2.vp *= np.float32(1.01); m2.rho *= np.float32(0.99)
        pg.update_medium(m2)
        ex2 = G.padarray(m2, G.NPML, pg.c.pml_faces)
        for name in pg.c.mparams:
            assert np.array_equal(pg.engine.get_medium(name), ex2[name]), name
        assert np.array_equal(pg.c.mod[pg.c.mparams[0]], ex2[pg.c.mparams[0]])


def test_fwi_gradient_acoustic2d(G, O):
    """BASELINE config 4 down-sized: forward_save + adjoint + imaging; gradients w.r.t. invK and rho."""
    from geophyinv_jl_b200.host import gallery
    kw, true = gallery.c4_fwi2d(nz=70, nx=110, nt=500, nss=3, nr=24, fq=10.0)
    mk = lambda cls: cls(G.FdtdAcoustic("forward_save"), **kw)
    pg, po = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, shot_batch=2), O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw)
    # observed data from the true medium (oracle), then gradient at the model medium with both
    pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    m = pg.get_modelvector()
    gg, go = np.zeros_like(m), np.zeros_like(m)
    lg = G.gradient(gg, m, dobs, pg)
    lo = G.gradient(go, m, dobs, po)
    assert abs(lg - lo) <= 1e-5 * abs(lo)
    half = m.size // 2
    eK, eR = rel_l2(gg[:half], go[:half]), rel_l2(gg[half:], go[half:])
    print(f"FWI gradient: invK rel-L2 {eK:.3e}, rho rel-L2 {eR:.3e}, loss {lg:.6e} vs {lo:.6e}")
    assert np.abs(go[:half]).max() > 0 and np.abs(go[half:]).max() > 0
    assert eK <= GRAD_TOL and eR <= GRAD_TOL
    # stacked raw gradients on the extended grid too
    for name in ("invK", "rho"):
        assert rel_l2(pg.engine.get_gradient(name), po.engine.get_gradient(name)) <= GRAD_TOL


@pytest.mark.parametrize("case", ["acou2d", "elastic3d"])
def test_simultaneous_and_coincident_sources(G, O, case):
    """Supersources of several simultaneous sources (the reference's `ns` columns of the spray matrix, fdtd.jl:469-493) with
    different wavelets: two sources at the same point and one within a cell of it (rows of S hit by several columns -- the
    per-cell summation order of `mul!(buf, S, w)` matters), one source on the first interior node of the medium (its taps reach
    the PML side), receivers that coincide, and two source fields injected in the same step."""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.data import AGeomss, make_srcwav
    rng = np.random.default_rng(3)
    if case == "acou2d":
        kw = gallery.c2_acou2d_layered(nz=70, nx=100, nt=300, nss=2, nr=8, fq=15.0, sfield="p", rfields=("p", "vz"))
        attrib, fields, sfields = G.FdtdAcoustic, ["p", "vx", "vz"], ["p", "vz"]
    else:
        kw = gallery.c3_elastic3d(n=30, nt=110, nr=6, fq=25.0, rfields=("vz", "vx"))
        attrib, fields, sfields = G.FdtdElastic, ["tauxx", "tauzz", "tauxz", "vx", "vy", "vz"], ["vz", "tauxx"]
    grid, tgrid = kw["medium"].grid, kw["tgrid"]
    names = ["z", "x"] if len(grid) == 2 else ["z", "y", "x"]
    ageom = []
    for a in kw["ageom"]:
        s0 = [float(a.s[d][0]) for d in names]
        pts = [s0, s0, [c + 0.4 * g.step for c, g in zip(s0, grid)], [g.first + 0.25 * g.step for g in grid]]
        src = {d: np.array([p[i] for p in pts]) for i, d in enumerate(names)}
        rec = {d: np.concatenate([a.r[d], a.r[d][:2]]) for d in names}            # the first two receivers twice
        ageom.append(AGeomss(src, rec))
    srcwav = make_srcwav(tgrid, ageom, sfields)
    base = kw["srcwav"][0].d[list(kw["srcwav"][0].d)[0]][:, 0]
    for s in srcwav:
        for f in sfields:
            amp = (1e6 if f.startswith("v") else 1.0) / (1e6 if list(kw["srcwav"][0].d)[0].startswith("v") else 1.0)
            s.d[f][...] = (base[:, None] * amp * rng.uniform(0.5, 1.5, (1, s.n))).astype(np.float32)
    kw = {**kw, "ageom": ageom, "srcwav": srcwav}
    pg, po = both(G, O, attrib, kw)
    pg.update(); po.update()
    err, exact = compare_records(pg, po)
    print(f"{case}: 4 simultaneous sources x 2 fields, coincident taps: rel-L2 {err:.3e}, bit-exact {exact}")
    assert err <= REC_TOL
    assert compare_fields(pg, po, fields) <= REC_TOL
    # the duplicated receivers record the same samples
    for f in pg.c.rfields:
        d = pg.c.data[0][0].d[f]
        assert np.array_equal(d[:, 0], d[:, -2]) and np.array_equal(d[:, 1], d[:, -1])


@pytest.mark.parametrize("case", ["acou2d_batched", "elastic2d", "acou3d"])
def test_boundary_save_and_force_match_oracle(G, O, case):
    """:forward_save stores 3+3 planes per axis per stored field (p | tauxx, tauxz, tauzz) and the final state;
    the :adjoint run of pw 1 forces them back (boundary.jl:17-306).  One batched launch covers all shots, fields,
    axes and planes in the engine; the back-propagated snapshots and final fields must equal the oracle's bits."""
    from geophyinv_jl_b200.host import gallery
    if case == "acou2d_batched":
        kw = gallery.c1_acou2d_homo(nz=81, nx=91, nr=8, nt=260, dt=1.5e-3, fq=14.0, sfield="vz", rfields=("vz",), nss=3)
        attrib, snapf, fields = (lambda: G.FdtdAcoustic("forward_save")), "p", ("p", "vx", "vz")
    elif case == "elastic2d":
        kw = gallery.elastic2d(nz=70, nx=84, nt=220, nr=8, nss=2)
        attrib, snapf, fields = (lambda: G.FdtdElastic("forward_save")), "tauxx", ("tauxx", "tauzz", "tauxz", "vx", "vz")
    else:
        kw = gallery.acou3d(n=30, nt=120, nr=6, sfield="vz", rfields=("vz",))
        attrib, snapf, fields = (lambda: G.FdtdAcoustic("forward_save")), "p", ("p", "vx", "vy", "vz")
    nt = len(kw["tgrid"])
    its = [nt // 4, nt // 2, 3 * nt // 4]
    tg = kw["tgrid"]
    out = []
    for cls in (G.SeisForwExpt, O.OraclePFdtd):
        pa = cls(attrib(), **kw, snaps_field=snapf, tsnaps=[tg.values[i - 1] for i in its])
        pa.update()
        forw = [[s.copy() for s in shot] for shot in pa["snaps", 1]]
        pa.update_srcwav(pa.c.srcwav, [-1, 0])
        pa.c.attrib_mod.mode = "adjoint"
        pa.c.itsnaps = [nt - i for i in its]
        pa.engine.set_snap_steps(pa.c.itsnaps)
        pa.update(dict(activepw=[1], src_flags=[True, False], rec_flags=[False, False]))
        back = [[s.copy() for s in shot] for shot in pa["snaps", 1]]
        out.append((forw, back, pa))
    (fg, bg, pg), (fo, bo, po) = out
    for iss in range(len(fg)):
        for a, b in zip(fg[iss] + bg[iss], fo[iss] + bo[iss]):
            assert np.abs(b).max() > 0
            assert np.array_equal(a, b), f"{case}: snapshot of shot {iss} differs from the oracle"
    assert compare_fields(pg, po, fields) == 0.0
    print(f"{case}: forward and back-propagated snapshots + final fields bit-identical to the oracle")


# --------------------------------------------------------------------------------------------------
# FD-Born: scattering sources pw 1 -> pw 2, LinearMap
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sfield,rfields", [("p", ["p", "vx"]), ("vz", ["vz"])])
def test_born_records_match_oracle(G, O, sfield, rfields):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nz=61, nx=73, nt=400, nr=12, sfield=sfield, rfields=rfields, fq=10.0, dt=1.8e-3, nss=3)
    m0 = kw["medium"]
    rng = np.random.default_rng(2)
    m0.vp *= (1 + 0.02 * rng.standard_normal(m0.vp.shape)).astype(np.float32)
    mp = m0.copy()
    mp.vp[24:36, 30:44] *= np.float32(1.02)
    mp.rho[20:30, 28:40] *= np.float32(0.97)
    pg = G.SeisForwExpt(G.FdtdAcoustic(born=True), **kw, shot_batch=2)
    po = O.OraclePFdtd(G.FdtdAcoustic(born=True), **kw)
    for p in (pg, po):
        G.update(p, m0, mp)
        p.update()
    worst, exact = compare_records(pg, po, ipw=1)
    print(f"FD-Born {sfield}->{rfields}: rel-L2 {worst:.3e}, bit-exact {exact}")
    assert worst <= REC_TOL
    assert max(np.abs(pg.c.data[1][0].d[f]).max() for f in rfields) > 0


def test_born_linear_map_dot_test(G, O):
    """<y, F x> == <x, F' y> on the GPU (Float32 arithmetic; upstream's gate is rtol 1e-5 in Float64) and the
    unshifted imaging matches the oracle."""
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nz=51, nx=57, nt=500, nr=10, sfield="vz", rfields=("vz",), fq=8.0, dt=1.8e-3)
    pg = G.SeisForwExpt(G.FdtdAcoustic("forward_save", born=True), **kw)
    po = O.OraclePFdtd(G.FdtdAcoustic("forward_save", born=True), **kw)
    Fg, Fo = G.LinearMap(pg), G.LinearMap(po)
    n = pg.c.gradients["invK"].shape
    rng = np.random.default_rng(9)
    a0 = np.zeros(n, np.float32, order="F"); b0 = np.zeros(n, np.float32, order="F")
    inner = (slice(G.NPML + 9, n[0] - G.NPML - 9), slice(G.NPML + 9, n[1] - G.NPML - 9))     # away from sources / receivers, see LinearMap
    a0[inner] = rng.standard_normal(a0[inner].shape) * 1e-12
    b0[inner] = rng.standard_normal(b0[inner].shape) * 25.0
    x = np.concatenate([a0.ravel(order="F"), b0.ravel(order="F")])
    y = rng.standard_normal(Fg.shape[0]).astype(np.float32)
    dg, do = Fg @ x, Fo @ x
    assert rel_l2(dg, do) <= REC_TOL
    gg, go = Fg.T @ y, Fo.T @ y
    assert rel_l2(gg, go) <= GRAD_TOL
    a = float(np.dot(y.astype(np.float64), dg.astype(np.float64)))
    b = float(np.dot(x.astype(np.float64), gg.astype(np.float64)))
    print(f"Born dot test (GPU, Float32): <y,Fx> = {a:.8e}, <x,F'y> = {b:.8e}, rel {abs(a - b) / abs(a):.2e}; "
          f"F rel-L2 vs oracle {rel_l2(dg, do):.1e}, F' {rel_l2(gg, go):.1e}")
    assert abs(a - b) <= 1e-4 * abs(a)


# --------------------------------------------------------------------------------------------------
# the CUDA engine against closed-form solutions, no oracle in the loop
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tma", ["1", "2"])
def test_engine_matches_the_stokes_solution(G, monkeypatch, tma):
    """3-D elastic full space, point force: the Float32 CUDA records against the Stokes solution (near-field term, far-field P and
    S; tests/test_oracle_invariants.py::stokes_misfits).  No free parameter and no oracle: normalised squared misfit 2.6e-4 (:vz),
    2.4e-4 (:vx), the same as the CPU restatement.  GPI_TMA3=2 sends the same case through the TMA-pipelined kernels (the default
    picks the register-staged ones for this 128-node z extent)."""
    from test_oracle_invariants import stokes_misfits
    monkeypatch.setenv("GPI_TMA3", tma)
    for rf, err in stokes_misfits(G, G.SeisForwExpt).items():
        print(f"engine vs Stokes solution (GPI_TMA3={tma}), :vz force recorded as :{rf}: normalised squared misfit {err:.3e}")
        assert err < 1e-3


# --------------------------------------------------------------------------------------------------
# illum_flag: source illumination (compute_illum! / stack_illums!, fdtd.jl:556-581)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["acou2d", "acou3d", "fwi2d"])
def test_illumination_matches_oracle(G, O, case):
    """Sum over time steps and supersources of abs2(p) of pw 1 (Float32 square, Float64 accumulation in shot order): the engine's k_illum /
    k_axpy1d against the oracle bit for bit -- three supersources in batches of two (2-D; launch by launch, captured, replayed), 3-D, and
    inside gradient! where every update! (forward_save, then adjoint with the ping-pong levels) refreshes it."""
    from geophyinv_jl_b200.host import gallery
    if case == "fwi2d":
        kw, true = gallery.c4_fwi2d(nz=44, nx=60, nt=151, nss=3, nr=10, fq=12.0)
        pt = O.OraclePFdtd(G.FdtdAcoustic(), **{**kw, "medium": true})
        pt.update()
        dobs = [d.copy() for d in pt.c.data[0]]
        pg = G.PFdtd(G.FdtdAcoustic("forward_save"), **kw, shot_batch=2, illum_flag=True)
        po = O.OraclePFdtd(G.FdtdAcoustic("forward_save"), **kw, illum_flag=True)
        m = pg.get_modelvector()
        for rep in range(3):
            gg, go = np.zeros_like(m), np.zeros_like(m)
            G.gradient(gg, m, dobs, pg); G.gradient(go, m, dobs, po)
            assert np.array_equal(gg, go)
            assert po.c.illum_stack.max() > 0 and np.array_equal(pg.c.illum_stack, po.c.illum_stack), f"run {rep}"
        return
    if case == "acou2d":
        kw, extra = gallery.c2_acou2d_layered(nz=60, nx=80, nt=200, nss=3, nr=8, fq=15.0), {"shot_batch": 2}
    else:
        kw, extra = gallery.acou3d(n=20, nt=90, nr=6), {}
    pg = G.SeisForwExpt(G.FdtdAcoustic(), **kw, **extra, illum_flag=True)
    po = O.OraclePFdtd(G.FdtdAcoustic(), **kw, illum_flag=True)
    po.update()
    assert po.c.illum_stack.max() > 0 and po.c.illum_stack.dtype == np.float64
    for rep in range(3):
        pg.update()
        assert np.array_equal(pg["illum"], po["illum"]), f"run {rep}"
    worst, exact = compare_records(pg, po)
    assert exact
    with pytest.raises(Exception):
        G.SeisForwExpt(G.FdtdElastic(), **gallery.elastic2d(nt=5), illum_flag=True)


# --------------------------------------------------------------------------------------------------
# degenerate sizes and placements
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["nt1", "nt2", "nr1", "corners", "rigid_box_no_cpml", "three_shots_batch2"])
def test_small_and_degenerate_cases(G, O, case):
    """One and two time steps (the pressure record of step it+1 is written by step it: the last row must not be), a single
    receiver, source and receivers on the corner nodes of the medium (taps next to the CPML interface), a rigid box without any
    CPML face (no slab anywhere; every axis still >= npml nodes, which update_pml! needs, cpml.jl:100-105), three shots in batches
    of two.  Records bit-identical to the oracle."""
    from geophyinv_jl_b200.host import gallery
    from geophyinv_jl_b200.host.data import AGeomss, make_srcwav
    attrib, extra = G.FdtdAcoustic, {}
    if case in ("nt1", "nt2"):
        kw = gallery.c1_acou2d_homo(nz=30, nx=34, nt=int(case[2:]), nr=5)
    elif case == "nr1":
        kw = gallery.c1_acou2d_homo(nz=30, nx=34, nt=60, nr=1)
    elif case == "corners":
        kw = gallery.c1_acou2d_homo(nz=30, nx=34, nt=60, nr=1)
        g = kw["medium"].grid
        kw["ageom"] = [AGeomss({"z": [g[0][0]], "x": [g[1][0]]}, {"z": [g[0][0], g[0][len(g[0]) - 1]], "x": [g[1][0], g[1][len(g[1]) - 1]]})]
        kw["srcwav"] = make_srcwav(kw["tgrid"], kw["ageom"], ["p"], kw["srcwav"][0].d["p"][:, 0])
    elif case == "rigid_box_no_cpml":
        kw = gallery.c1_acou2d_homo(nz=44, nx=47, nt=120, nr=3)
        kw["pml_faces"] = []; kw["rigid_faces"] = ["zmin", "zmax", "xmin", "xmax"]
    else:
        kw = gallery.elastic2d(nz=24, nx=28, nt=50, nr=4, nss=3)
        attrib, extra = G.FdtdElastic, {"shot_batch": 2}
    pg, po = both(G, O, attrib, kw, **extra)
    pg.update(); po.update()
    for iss in range(len(pg.c.data[0])):
        for f in pg.c.rfields:
            a, b = pg.c.data[0][iss].d[f], po.c.data[0][iss].d[f]
            assert a.shape == b.shape and np.array_equal(a, b), (case, iss, f)
    if case not in ("nt1", "nt2"):
        assert max(np.abs(po.c.data[0][0].d[f]).max() for f in po.c.rfields) > 0


def test_axis_shorter_than_npml_is_rejected_with_a_message(G):
    from geophyinv_jl_b200.host import gallery
    kw = gallery.c1_acou2d_homo(nz=8, nx=9, nt=4, nr=2)
    kw["pml_faces"] = []
    with pytest.raises(ValueError, match="fewer than npml"):
        G.SeisForwExpt(G.FdtdAcoustic(), **kw)


# --------------------------------------------------------------------------------------------------
# GPI_PINGPONG=1: adjoint runs without save_tp!'s copy (the two wavefield sets alternate as time levels, out-of-place kernels)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("physics", ["acoustic", "elastic", "acoustic3d", "elastic3d"])
def test_pingpong_adjoint_equals_the_copy_path(G, O, monkeypatch, physics):
    """The FWI gradient (forward_save + adjoint + imaging, order 2) with the time levels ping-ponged between W and TP must be
    bit-identical to the run that copies W -> TP every step (save_tp.jl:5-12) and to the oracle; nt is odd so that the run ends on
    the other set (the handle swaps them), and the experiment is run twice to cover the swapped start.  The launch count proves the
    ping-pong path ran: it replaces the copy by two more boundary launches per step and batch.  In 3-D the out-of-place kernels are
    the register-staged float4 ones (for elastic media the copy path runs the TMA tiles, so the two families are compared as well).
    2-D acoustic media take the fused pass of kernels2a.cuh (stress update of both wavefields + imaging; GPI_FUSE2A=0 opts out): the
    sources sit inside the medium and the boundary store's planes cross the model, so the pre-force stash and the re-imaging of the
    source cells are both exercised."""
    from geophyinv_jl_b200.host import gallery
    extra = {"shot_batch": 2}
    if physics == "acoustic":
        kw, true = gallery.c4_fwi2d(nz=60, nx=90, nt=301, nss=3, nr=16, fq=10.0)
        attrib = G.FdtdAcoustic
    elif physics == "elastic":
        kw, true = gallery.fwi2d_elastic(nt=301)
        attrib = G.FdtdElastic
    elif physics == "acoustic3d":
        kw, true = gallery.fwi3d(nt=151)
        attrib, extra = G.FdtdAcoustic, {}
    else:
        kw, true = gallery.fwi3d_elastic(nt=101)
        attrib, extra = G.FdtdElastic, {}
    pt = O.OraclePFdtd(attrib(), **{**kw, "medium": true})
    pt.update()
    dobs = [d.copy() for d in pt.c.data[0]]
    po = O.OraclePFdtd(attrib("forward_save"), **kw)
    m = po.get_modelvector()
    go = np.zeros_like(m)
    G.gradient(go, m, dobs, po)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("GPI_PINGPONG", flag)            # read by gpi_create
        pg = G.PFdtd(attrib("forward_save"), **kw, **extra)
        for rep in range(4):                    # 2-D: launch by launch, captured into CUDA graphs, replayed twice (odd nt: the levels swap per run)
            g = np.zeros_like(m)
            loss = G.gradient(g, m, dobs, pg)
            res[flag, rep] = (g, loss, pg.last_launches)
    nss = len(kw["ageom"])
    nbatch = -(-nss // extra["shot_batch"]) if extra else nss
    nt = len(kw["tgrid"])
    for rep in range(4):
        g0, l0, n0 = res["0", rep]; g1, l1, n1 = res["1", rep]
        assert np.array_equal(g0, g1) and l0 == l1, f"ping-pong differs from the copy path (run {rep})"
        # + two boundary launches per step (the shell of the TMA tiles is walked by warps of the tile kernel itself: no launch of its own);
        # 2-D acoustic: the imaging is fused into the stress pass of both wavefields (k_stress2a) -- the stash launch comes, the restore
        # and k_grad2d go
        assert n1 - n0 == (0 if physics == "acoustic" else 2 * nt * nbatch), (n0, n1)
        assert rel_l2(g1, go) <= GRAD_TOL
    print(f"ping-pong adjoint ({physics}): gradient bit-identical to the copy path, rel-L2 vs oracle {rel_l2(res['1', 0][0], go):.1e}, "
          f"launches {res['0', 0][2]:.0f} -> {res['1', 0][2]:.0f}")
