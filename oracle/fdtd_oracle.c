/*
 * fdtd_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of GeoPhyInv.jl's src/fdtd hot path.
 *
 * This file is the parity oracle and the CPU baseline for the B200 engine.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  It is a restatement, not the Julia code: Julia is not installed in this image, so
 * the reference cannot be run here, and the reference ships no golden vectors for this path
 * (SURVEY.md section 8c).  What pins it: (1) fixtures evaluated from the reference's own kernel text
 * (see "Parity status" below), (2) the reference's test invariants (analytic homogeneous solution, time
 * reversal, gradient vs finite differences, dot test) and closed-form solutions with no free parameter
 * (3-D acoustic Green's function, 3-D elastic Stokes solution, reciprocity), tests/test_oracle_invariants.py.
 *
 * Structure follows the reference literally (all citations relative to /root/reference):
 *   - same arrays with the same shapes as src/fields.jl:92-671 (column-major, [z,(y),x]);
 *   - one loop nest per ParallelStencil `@parallel` kernel, each assignment guarded like
 *     `@within` (src/fdtd/diff2D.jl:17-28, diff3D.jl:17-34): unfused derivative sweep ->
 *     CPML sweeps -> update sweep, in the order of src/fdtd/propagate.jl:170-247;
 *   - Float32 arithmetic without FMA contraction (compile with -ffp-contract=off); float literals inside @parallel
 *     kernels typed as ParallelStencil types them (see WIDE below), the Float64 promotion of the one literal outside
 *     a kernel (medium.jl:165 `2.0 *`) restated explicitly;
 *   - OpenMP static schedule over the outermost index = what ParallelStencil's Threads backend does.
 *
 * Build: see oracle/Makefile (REAL=float -> liboracle_f32.so, REAL=double -> liboracle_f64.so,
 *        REAL=float -DORC_LITERALS_F64 -> liboracle_f32_lit64.so).
 * Parity status: PINNED to the reference's source text -- tests/golden/from_reference.py evaluates the @parallel bodies,
 * macros, templated kernels and call sequences parsed out of /root/reference/src with numpy, and
 * tests/test_reference_pinned.py checks that this file reproduces those fixtures bit for bit (orders 2 and 4, 2-D / 3-D,
 * acoustic / elastic, forward and forward_save + adjoint + imaging).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/gpifdtd.h"

#ifndef REAL
#define REAL float
#endif
/* Float literals inside `@parallel` kernels (the 0.5 / 0.25 of the @av_* macros, the 27.0 of the order-4 differences).
 * ParallelStencil's @parallel retypes untyped float literals of a kernel to the number type the package was initialised
 * with (`literaltypes(numbertype, kernel)`; GeoPhyInv: @init_parallel_stencil(Threads, Float32, N), src/GeoPhyInv.jl:95-100),
 * so under the shipped Float32 preference these expressions are pure Float32 -- the DEFAULT here (WIDE = REAL).
 * -DORC_LITERALS_F64 keeps them Float64 (plain Julia promotion: expression in Float64, one rounding at the store), the
 * reading round 1 took.  Both typings are pinned bit for bit against the reference's source text evaluated in the same
 * typing (tests/test_reference_pinned.py); they differ by <= 1 ulp per operation (1e-7 relative).
 * NOT affected: `inv(invl) + 2.0 * inv(invmu)` of update_dmod! (medium.jl:164-166) is a plain Julia broadcast, outside any
 * @parallel kernel: its 2.0 stays Float64 in both typings. */
#ifdef ORC_LITERALS_F64
typedef double WIDE;
#else
typedef REAL WIDE;
#endif

/* ------------------------------------------------------------------------------------------------
 * arrays: 1-based accessors so the Julia index algebra can be restated verbatim
 * ---------------------------------------------------------------------------------------------- */
typedef struct { REAL* d; int n[3]; size_t len; } arr;   /* n = (nz, ny, nx); ny == 1 in 2-D */

#define A2(a, iz, ix)      ((a).d[((size_t)(iz) - 1) + (size_t)(a).n[0] * ((size_t)(ix) - 1)])
#define A3(a, iz, iy, ix)  ((a).d[((size_t)(iz) - 1) + (size_t)(a).n[0] * (((size_t)(iy) - 1) + (size_t)(a).n[1] * ((size_t)(ix) - 1))])

static int arr_alloc(arr* a, const int n[3]) {
    a->n[0] = n[0]; a->n[1] = n[1]; a->n[2] = n[2];
    a->len = (size_t)n[0] * n[1] * n[2];
    a->d = (REAL*)calloc(a->len ? a->len : 1, sizeof(REAL));
    return a->d ? 0 : 1;
}
static void arr_free(arr* a) { free(a->d); a->d = NULL; a->len = 0; }
static void arr_zero(arr* a) { if (a->d) memset(a->d, 0, a->len * sizeof(REAL)); }

/* ------------------------------------------------------------------------------------------------
 * field shapes: src/fields.jl:92-671 (get_mgrid) reduced to per-axis node types, O = order - 1
 *   I: tauii nodes, length n              V: velocity nodes (-O/2), length n+O
 *   H: half nodes (+O/2), length n-O      J: inner integer nodes (+O), length n-2O
 * ---------------------------------------------------------------------------------------------- */
static const char* field_types3(int f) {           /* (z,y,x) */
    switch (f) {
    case GPI_P: case GPI_TAUXX: case GPI_TAUYY: case GPI_TAUZZ:
    case GPI_DVXDX: case GPI_DVYDY: case GPI_DVZDZ: return "III";
    case GPI_VX: return "IIV"; case GPI_VY: return "IVI"; case GPI_VZ: return "VII";
    case GPI_DPDX: case GPI_DTAUXXDX: case GPI_DTAUXYDY: case GPI_DTAUXZDZ: return "JJH";
    case GPI_DPDY: case GPI_DTAUYYDY: case GPI_DTAUXYDX: case GPI_DTAUYZDZ: return "JHJ";
    case GPI_DPDZ: case GPI_DTAUZZDZ: case GPI_DTAUXZDX: case GPI_DTAUYZDY: return "HJJ";
    case GPI_TAUXY: case GPI_DVXDY: case GPI_DVYDX: return "JHH";
    case GPI_TAUXZ: case GPI_DVXDZ: case GPI_DVZDX: return "HJH";
    case GPI_TAUYZ: case GPI_DVYDZ: case GPI_DVZDY: return "HHJ";
    }
    return NULL;
}
/* O = _fd_order - 1 (1 or 3): V nodes n+O, H nodes n-O, J nodes n-2O (fields.jl:92-671).  File-scope: the entry
 * points set it from the handle's order before any shape or sweep is evaluated. */
static int g_O = 1;
static int type_len(char t, int n) { return t == 'I' ? n : t == 'V' ? n + g_O : t == 'H' ? n - g_O : n - 2 * g_O; }

/* also exported (orc_field_shape) so the tests can check the engine's gpi_field_shape against it */
static int field_shape(int ndims, int f, const int n[3], int out[3]) {
    const char* t = field_types3(f);
    if (!t) return 1;
    if (ndims == 3) {
        for (int q = 0; q < 3; q++) out[q] = type_len(t[q], n[q]);
    } else {
        /* 2-D: no y axis; tauxz / dvxdz / dvzdx live on (H,H) (fields.jl 2-D methods) */
        if ((f == GPI_VY || f == GPI_TAUYY || f == GPI_TAUXY || f == GPI_TAUYZ ||
            f == GPI_DPDY || f == GPI_DVYDY || f == GPI_DVXDY || f == GPI_DVYDX || f == GPI_DVYDZ ||
            f == GPI_DVZDY || f == GPI_DTAUYYDY || f == GPI_DTAUXYDX || f == GPI_DTAUXYDY ||
            f == GPI_DTAUYZDY || f == GPI_DTAUYZDZ)) return 1;
        out[0] = type_len(t[0], n[0]); out[1] = 1; out[2] = type_len(t[2], n[2]);
    }
    return 0;
}
/* axis (0=z,1=y,2=x) a derivative field is taken along = its last letter (cpml.jl:131-132) */
static int dfield_axis(int f) {
    switch (f) {
    case GPI_DPDX: case GPI_DVXDX: case GPI_DVYDX: case GPI_DVZDX:
    case GPI_DTAUXXDX: case GPI_DTAUXYDX: case GPI_DTAUXZDX: return 2;
    case GPI_DPDY: case GPI_DVYDY: case GPI_DVXDY: case GPI_DVZDY:
    case GPI_DTAUYYDY: case GPI_DTAUXYDY: case GPI_DTAUYZDY: return 1;
    case GPI_DPDZ: case GPI_DVZDZ: case GPI_DVXDZ: case GPI_DVYDZ:
    case GPI_DTAUZZDZ: case GPI_DTAUXZDZ: case GPI_DTAUYZDZ: return 0;
    }
    return -1;
}

/* which fields exist for (physics, ndims): Fields(attrib_mod) (fields.jl:12-26) */
static int field_exists(int ndims, int physics, int f) {
    int tmp[3]; const int n[3] = {8, ndims == 3 ? 8 : 1, 8};
    if (field_shape(ndims, f, n, tmp)) return 0;
    if (physics == GPI_ACOUSTIC) {
        if (f >= GPI_TAUXX && f <= GPI_TAUYZ) return 0;
        if (f >= GPI_DVXDY && f <= GPI_DVZDY) return 0;
        if (f >= GPI_DTAUXXDX) return 0;
        return 1;
    }
    if (f == GPI_P || f == GPI_DPDX || f == GPI_DPDY || f == GPI_DPDZ) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * state
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int ncol; int64_t* colptr; int64_t* rowval; REAL* nzval; } csc;

typedef struct {               /* P_x_worker_x_pw_x_ss (types.jl:8-23) */
    csc   spray[GPI_NWAVEFIELD], interp[GPI_NWAVEFIELD];
    REAL* wavelets[GPI_NWAVEFIELD];  int ns[GPI_NWAVEFIELD];   /* [nt, ns] */
    REAL* records[GPI_NWAVEFIELD];                              /* [nt, nr] */
    /* boundary store (boundary.jl:59-108): per wavefield, per axis [nt][field shape with axis=2*nbound] + snap */
    REAL* bnd[GPI_NWAVEFIELD][3];  size_t bnd_slot[GPI_NWAVEFIELD][3];
    REAL* snap[GPI_NWAVEFIELD];
    arr   grad[GPI_NPARAM];
    REAL** usnaps;                                              /* user snapshots [nsnaps] */
} shot_t;

typedef struct {               /* P_x_worker_x_pw (types.jl:83-90) */
    arr w[GPI_NFIELD];         /* w1[:t]  */
    arr wtp[GPI_NWAVEFIELD];   /* w1[:tp] */
    arr mem[GPI_NFIELD];       /* memory_pml */
    arr vbuf[GPI_NWAVEFIELD];  /* velocity_buffer (VX,VY,VZ slots) */
    arr taubuf;                /* tauii_buffer */
    shot_t* ss;
} pw_t;

/* dmod ids (medium.jl:81-95) */
enum { DM_BX = 0, DM_BY, DM_BZ, DM_DTK, DM_DTLAMBDA, DM_DTM, DM_MUXZ, DM_MUXY, DM_MUYZ, DM_N };

typedef struct orc_handle {
    gpi_config c;
    int nz, ny, nx, nd;
    REAL dt, dtI, dzI, dyI, dxI;
    arr mod[GPI_NPARAM], dmod[DM_N];
    arr modp[GPI_NPARAM], bornc[3];   /* FD-Born: medium perturbation; d(dtK), d(bx), d(bz) */
    int born_ready;
    REAL *pa[GPI_NFIELD], *pb[GPI_NFIELD], *pk[GPI_NFIELD];   /* CPML a, b, kI (2*npml each) */
    arr gradients[GPI_NPARAM];
    pw_t pw[2];
    int32_t* itsnaps;
    char err[256];
    double last_run_s; double last_steps;
    int illum_on; double* illum_shot; double* illum_stack;     /* compute_illum!, stack_illums! (fdtd.jl:556-581) */
} orc_handle;

static char g_err[256];

/* ------------------------------------------------------------------------------------------------
 * kernels, 2-D acoustic: src/fdtd/advance_acou.jl:258-306
 * ---------------------------------------------------------------------------------------------- */
static int imax2(int a, int b) { return a > b ? a : b; }
static int imax3(int a, int b, int c) { return imax2(a, imax2(b, c)); }

#define OMP_FOR _Pragma("omp parallel for schedule(static)")

/* ---- finite-difference macros (diff2D.jl:15-97, diff3D.jl:15-134) -------------------------------------------
 * O = _fd_order - 1 is the shift of `@inn` and of the `_i` macros (izi = iz + O, diff2D.jl:5-13).
 * order 2: A[i+1] - A[i] in the array's own precision.
 * order 4: A[i+2]*27.0 - A[i+1]*27.0 + A[i] - A[i+3]: the Float64 literals promote the whole expression, the
 *          product with d?I (which carries the 1/24, fdtd.jl:318-319) stays Float64 and the store rounds once.
 * The order is per handle; the entry points copy it into the file-scope g_O before any sweep runs. */
static inline REAL fd(const REAL* p, size_t s, REAL sI) {
    if (g_O == 1) return (p[s] - p[0]) * sI;
    return (REAL)(((WIDE)p[2 * s] * (WIDE)27.0 - (WIDE)p[s] * (WIDE)27.0 + (WIDE)p[0] - (WIDE)p[3 * s]) * (WIDE)sI);
}
#define DZ2(a, iz, ix, sI)     fd(&A2(a, iz, ix), 1, sI)
#define DX2(a, iz, ix, sI)     fd(&A2(a, iz, ix), (size_t)(a).n[0], sI)
#define DZ3(a, iz, iy, ix, sI) fd(&A3(a, iz, iy, ix), 1, sI)
#define DY3(a, iz, iy, ix, sI) fd(&A3(a, iz, iy, ix), (size_t)(a).n[0], sI)
#define DX3(a, iz, iy, ix, sI) fd(&A3(a, iz, iy, ix), (size_t)(a).n[0] * (size_t)(a).n[1], sI)

/* compute_dp! (advance_acou.jl:258-262) */
static void compute_dp_2d(arr p, arr dpdx, arr dpdz, REAL dzI, REAL dxI) {
    const int O = g_O;
    int nz = imax3(p.n[0], dpdx.n[0], dpdz.n[0]), nx = imax3(p.n[2], dpdx.n[2], dpdz.n[2]);
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dpdx.n[0] && ix <= dpdx.n[2]) A2(dpdx, iz, ix) = DX2(p, iz + O, ix, dxI);   /* @d_xi */
        if (iz <= dpdz.n[0] && ix <= dpdz.n[2]) A2(dpdz, iz, ix) = DZ2(p, iz, ix + O, dzI);   /* @d_zi */
    }
}
/* compute_v! (advance_acou.jl:273-277) */
static void compute_v_acou_2d(arr vx, arr vz, arr bx, arr bz, arr dpdx, arr dpdz) {
    const int O = g_O;
    int nz = imax2(vx.n[0], vz.n[0]), nx = imax2(vx.n[2], vz.n[2]);
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= vx.n[0] - 2 * O && ix <= vx.n[2] - 2 * O) A2(vx, iz + O, ix + O) = A2(vx, iz + O, ix + O) + A2(bx, iz, ix) * A2(dpdx, iz, ix);
        if (iz <= vz.n[0] - 2 * O && ix <= vz.n[2] - 2 * O) A2(vz, iz + O, ix + O) = A2(vz, iz + O, ix + O) + A2(bz, iz, ix) * A2(dpdz, iz, ix);
    }
}
/* compute_dv! (advance_acou.jl:288-292) */
static void compute_dv_acou_2d(arr vx, arr vz, arr dvxdx, arr dvzdz, REAL dxI, REAL dzI) {
    int nz = imax2(vz.n[0], dvxdx.n[0]), nx = imax2(vx.n[2], dvxdx.n[2]);
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dvxdx.n[0] && ix <= dvxdx.n[2]) A2(dvxdx, iz, ix) = DX2(vx, iz, ix, dxI);   /* @d_xa */
        if (iz <= dvzdz.n[0] && ix <= dvzdz.n[2]) A2(dvzdz, iz, ix) = DZ2(vz, iz, ix, dzI);   /* @d_za */
    }
}
/* compute_p! (advance_acou.jl:303-306) */
static void compute_p_2d(arr p, arr dvxdx, arr dvzdz, arr dtK) {
    OMP_FOR
    for (int ix = 1; ix <= p.n[2]; ix++) for (int iz = 1; iz <= p.n[0]; iz++)
        A2(p, iz, ix) = A2(p, iz, ix) + (A2(dvxdx, iz, ix) + A2(dvzdz, iz, ix)) * A2(dtK, iz, ix);
}

/* ------------------------------------------------------------------------------------------------
 * kernels, 3-D acoustic: src/fdtd/advance_acou.jl:265-313
 * ---------------------------------------------------------------------------------------------- */
static void compute_dp_3d(arr p, arr dpdx, arr dpdy, arr dpdz, REAL dzI, REAL dyI, REAL dxI) {
    const int O = g_O;
    int nz = p.n[0], ny = p.n[1], nx = p.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dpdx.n[0] && iy <= dpdx.n[1] && ix <= dpdx.n[2]) A3(dpdx, iz, iy, ix) = DX3(p, iz + O, iy + O, ix, dxI);
        if (iz <= dpdy.n[0] && iy <= dpdy.n[1] && ix <= dpdy.n[2]) A3(dpdy, iz, iy, ix) = DY3(p, iz + O, iy, ix + O, dyI);
        if (iz <= dpdz.n[0] && iy <= dpdz.n[1] && ix <= dpdz.n[2]) A3(dpdz, iz, iy, ix) = DZ3(p, iz, iy + O, ix + O, dzI);
    }
}
#define INN3(a) (iz <= (a).n[0] - 2 * O && iy <= (a).n[1] - 2 * O && ix <= (a).n[2] - 2 * O)
#define I3(a)   A3(a, iz + O, iy + O, ix + O)
static void compute_v_acou_3d(arr vx, arr vy, arr vz, arr bx, arr by, arr bz, arr dpdx, arr dpdy, arr dpdz) {
    const int O = g_O;
    int nz = vz.n[0], ny = vy.n[1], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (INN3(vx)) I3(vx) = I3(vx) + A3(bx, iz, iy, ix) * A3(dpdx, iz, iy, ix);
        if (INN3(vy)) I3(vy) = I3(vy) + A3(by, iz, iy, ix) * A3(dpdy, iz, iy, ix);
        if (INN3(vz)) I3(vz) = I3(vz) + A3(bz, iz, iy, ix) * A3(dpdz, iz, iy, ix);
    }
}
static void compute_dv_acou_3d(arr vx, arr vy, arr vz, arr dvxdx, arr dvydy, arr dvzdz, REAL dxI, REAL dyI, REAL dzI) {
    int nz = vz.n[0], ny = vy.n[1], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dvxdx.n[0] && iy <= dvxdx.n[1] && ix <= dvxdx.n[2]) {
            A3(dvxdx, iz, iy, ix) = DX3(vx, iz, iy, ix, dxI);
            A3(dvydy, iz, iy, ix) = DY3(vy, iz, iy, ix, dyI);
            A3(dvzdz, iz, iy, ix) = DZ3(vz, iz, iy, ix, dzI);
        }
    }
}
/* compute_p! 3-D (advance_acou.jl:310-313): note the summation order (dvxdx + dvzdz + dvydy) */
static void compute_p_3d(arr p, arr dvxdx, arr dvydy, arr dvzdz, arr dtK) {
    OMP_FOR
    for (int ix = 1; ix <= p.n[2]; ix++) for (int iy = 1; iy <= p.n[1]; iy++) for (int iz = 1; iz <= p.n[0]; iz++)
        A3(p, iz, iy, ix) = A3(p, iz, iy, ix) + (A3(dvxdx, iz, iy, ix) + A3(dvzdz, iz, iy, ix) + A3(dvydy, iz, iy, ix)) * A3(dtK, iz, iy, ix);
}

/* ------------------------------------------------------------------------------------------------
 * kernels, 2-D elastic: src/fdtd/advance_elastic.jl:48-67,96-110,145-153,179-210
 * ---------------------------------------------------------------------------------------------- */
static void compute_dstress_2d(arr tauxx, arr tauzz, arr tauxz, arr dtauxxdx, arr dtauxzdx, arr dtauzzdz, arr dtauxzdz, REAL dxI, REAL dzI) {
    const int O = g_O;
    int nz = tauxx.n[0], nx = tauxx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dtauxxdx.n[0] && ix <= dtauxxdx.n[2]) A2(dtauxxdx, iz, ix) = DX2(tauxx, iz + O, ix, dxI);  /* @d_xi */
        if (iz <= dtauxzdx.n[0] && ix <= dtauxzdx.n[2]) A2(dtauxzdx, iz, ix) = DX2(tauxz, iz, ix, dxI);      /* @d_xa */
        if (iz <= dtauzzdz.n[0] && ix <= dtauzzdz.n[2]) A2(dtauzzdz, iz, ix) = DZ2(tauzz, iz, ix + O, dzI);  /* @d_zi */
        if (iz <= dtauxzdz.n[0] && ix <= dtauxzdz.n[2]) A2(dtauxzdz, iz, ix) = DZ2(tauxz, iz, ix, dzI);      /* @d_za */
    }
}
static void compute_v_el_2d(arr vx, arr vz, arr dtauxxdx, arr dtauxzdx, arr dtauzzdz, arr dtauxzdz, arr bx, arr bz) {
    const int O = g_O;
    int nz = vz.n[0], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= vx.n[0] - 2 * O && ix <= vx.n[2] - 2 * O)
            A2(vx, iz + O, ix + O) = A2(vx, iz + O, ix + O) - A2(bx, iz, ix) * (A2(dtauxxdx, iz, ix) + A2(dtauxzdz, iz, ix));
        if (iz <= vz.n[0] - 2 * O && ix <= vz.n[2] - 2 * O)
            A2(vz, iz + O, ix + O) = A2(vz, iz + O, ix + O) - A2(bz, iz, ix) * (A2(dtauxzdx, iz, ix) + A2(dtauzzdz, iz, ix));
    }
}
static void compute_dv_el_2d(arr vx, arr vz, arr dvxdx, arr dvzdz, arr dvxdz, arr dvzdx, REAL dxI, REAL dzI) {
    const int O = g_O;
    int nz = vz.n[0], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iz = 1; iz <= nz; iz++) {
        if (iz <= dvxdx.n[0] && ix <= dvxdx.n[2]) {
            A2(dvxdx, iz, ix) = DX2(vx, iz, ix, dxI);          /* @d_xa */
            A2(dvzdz, iz, ix) = DZ2(vz, iz, ix, dzI);          /* @d_za */
        }
        if (iz <= dvxdz.n[0] && ix <= dvxdz.n[2]) {
            A2(dvxdz, iz, ix) = DZ2(vx, iz, ix + O, dzI);      /* @d_zi */
            A2(dvzdx, iz, ix) = DX2(vz, iz + O, ix, dxI);      /* @d_xi */
        }
    }
}
static void compute_stressii_2d(arr tauxx, arr tauzz, arr dvxdx, arr dvzdz, arr dtM, arr dtlambda) {
    OMP_FOR
    for (int ix = 1; ix <= tauxx.n[2]; ix++) for (int iz = 1; iz <= tauxx.n[0]; iz++) {
        A2(tauxx, iz, ix) = A2(tauxx, iz, ix) - (A2(dtM, iz, ix) * A2(dvxdx, iz, ix)) - (A2(dtlambda, iz, ix) * (A2(dvzdz, iz, ix)));
        A2(tauzz, iz, ix) = A2(tauzz, iz, ix) - (A2(dtM, iz, ix) * A2(dvzdz, iz, ix)) - (A2(dtlambda, iz, ix) * (A2(dvxdx, iz, ix)));
    }
}
static void compute_stressij_2d(arr tauxz, arr dvxdz, arr dvzdx, arr dtavmu) {
    OMP_FOR
    for (int ix = 1; ix <= tauxz.n[2]; ix++) for (int iz = 1; iz <= tauxz.n[0]; iz++)
        A2(tauxz, iz, ix) = A2(tauxz, iz, ix) - A2(dtavmu, iz, ix) * (A2(dvxdz, iz, ix) + A2(dvzdx, iz, ix));
}

/* ------------------------------------------------------------------------------------------------
 * kernels, 3-D elastic: src/fdtd/advance_elastic.jl:10-46,69-94,112-143,155-206
 * ---------------------------------------------------------------------------------------------- */
#define IN3(a) (iz <= (a).n[0] && iy <= (a).n[1] && ix <= (a).n[2])

static void compute_dstress_3d(arr* w, REAL dxI, REAL dyI, REAL dzI) {
    const int O = g_O;
    arr tauxx = w[GPI_TAUXX], tauyy = w[GPI_TAUYY], tauzz = w[GPI_TAUZZ], tauxy = w[GPI_TAUXY], tauxz = w[GPI_TAUXZ], tauyz = w[GPI_TAUYZ];
    arr dtauxxdx = w[GPI_DTAUXXDX], dtauxydx = w[GPI_DTAUXYDX], dtauxzdx = w[GPI_DTAUXZDX];
    arr dtauyydy = w[GPI_DTAUYYDY], dtauxydy = w[GPI_DTAUXYDY], dtauyzdy = w[GPI_DTAUYZDY];
    arr dtauzzdz = w[GPI_DTAUZZDZ], dtauyzdz = w[GPI_DTAUYZDZ], dtauxzdz = w[GPI_DTAUXZDZ];
    int nz = tauxx.n[0], ny = tauxx.n[1], nx = tauxx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (IN3(dtauxxdx)) A3(dtauxxdx, iz, iy, ix) = DX3(tauxx, iz + O, iy + O, ix, dxI);  /* @d_xi */
        if (IN3(dtauxydx)) A3(dtauxydx, iz, iy, ix) = DX3(tauxy, iz, iy, ix, dxI);          /* @d_xa */
        if (IN3(dtauxzdx)) A3(dtauxzdx, iz, iy, ix) = DX3(tauxz, iz, iy, ix, dxI);          /* @d_xa */
        if (IN3(dtauyydy)) A3(dtauyydy, iz, iy, ix) = DY3(tauyy, iz + O, iy, ix + O, dyI);  /* @d_yi */
        if (IN3(dtauxydy)) A3(dtauxydy, iz, iy, ix) = DY3(tauxy, iz, iy, ix, dyI);          /* @d_ya */
        if (IN3(dtauyzdy)) A3(dtauyzdy, iz, iy, ix) = DY3(tauyz, iz, iy, ix, dyI);          /* @d_ya */
        if (IN3(dtauzzdz)) A3(dtauzzdz, iz, iy, ix) = DZ3(tauzz, iz, iy + O, ix + O, dzI);  /* @d_zi */
        if (IN3(dtauxzdz)) A3(dtauxzdz, iz, iy, ix) = DZ3(tauxz, iz, iy, ix, dzI);          /* @d_za */
        if (IN3(dtauyzdz)) A3(dtauyzdz, iz, iy, ix) = DZ3(tauyz, iz, iy, ix, dzI);          /* @d_za */
    }
}
static void compute_v_el_3d(arr* w, arr bx, arr by, arr bz) {
    const int O = g_O;
    arr vx = w[GPI_VX], vy = w[GPI_VY], vz = w[GPI_VZ];
    arr dtauxxdx = w[GPI_DTAUXXDX], dtauxydx = w[GPI_DTAUXYDX], dtauxzdx = w[GPI_DTAUXZDX];
    arr dtauyydy = w[GPI_DTAUYYDY], dtauxydy = w[GPI_DTAUXYDY], dtauyzdy = w[GPI_DTAUYZDY];
    arr dtauzzdz = w[GPI_DTAUZZDZ], dtauyzdz = w[GPI_DTAUYZDZ], dtauxzdz = w[GPI_DTAUXZDZ];
    int nz = vz.n[0], ny = vy.n[1], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (INN3(vx)) I3(vx) = I3(vx) - A3(bx, iz, iy, ix) * (A3(dtauxxdx, iz, iy, ix) + A3(dtauxydy, iz, iy, ix) + A3(dtauxzdz, iz, iy, ix));
        if (INN3(vy)) I3(vy) = I3(vy) - A3(by, iz, iy, ix) * (A3(dtauxydx, iz, iy, ix) + A3(dtauyydy, iz, iy, ix) + A3(dtauyzdz, iz, iy, ix));
        if (INN3(vz)) I3(vz) = I3(vz) - A3(bz, iz, iy, ix) * (A3(dtauxzdx, iz, iy, ix) + A3(dtauyzdy, iz, iy, ix) + A3(dtauzzdz, iz, iy, ix));
    }
}
static void compute_dv_el_3d(arr* w, REAL dxI, REAL dyI, REAL dzI) {
    const int O = g_O;
    arr vx = w[GPI_VX], vy = w[GPI_VY], vz = w[GPI_VZ];
    arr dvxdx = w[GPI_DVXDX], dvydy = w[GPI_DVYDY], dvzdz = w[GPI_DVZDZ];
    arr dvxdy = w[GPI_DVXDY], dvxdz = w[GPI_DVXDZ], dvydx = w[GPI_DVYDX], dvydz = w[GPI_DVYDZ], dvzdx = w[GPI_DVZDX], dvzdy = w[GPI_DVZDY];
    int nz = vz.n[0], ny = vy.n[1], nx = vx.n[2];
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (IN3(dvxdx)) A3(dvxdx, iz, iy, ix) = DX3(vx, iz, iy, ix, dxI);                 /* @d_xa */
        if (IN3(dvydy)) A3(dvydy, iz, iy, ix) = DY3(vy, iz, iy, ix, dyI);                 /* @d_ya */
        if (IN3(dvzdz)) A3(dvzdz, iz, iy, ix) = DZ3(vz, iz, iy, ix, dzI);                 /* @d_za */
        if (IN3(dvxdy)) A3(dvxdy, iz, iy, ix) = DY3(vx, iz + O, iy, ix + O, dyI);         /* @d_yi */
        if (IN3(dvxdz)) A3(dvxdz, iz, iy, ix) = DZ3(vx, iz, iy + O, ix + O, dzI);         /* @d_zi */
        if (IN3(dvydz)) A3(dvydz, iz, iy, ix) = DZ3(vy, iz, iy + O, ix + O, dzI);         /* @d_zi */
        if (IN3(dvydx)) A3(dvydx, iz, iy, ix) = DX3(vy, iz + O, iy + O, ix, dxI);         /* @d_xi */
        if (IN3(dvzdx)) A3(dvzdx, iz, iy, ix) = DX3(vz, iz + O, iy + O, ix, dxI);         /* @d_xi */
        if (IN3(dvzdy)) A3(dvzdy, iz, iy, ix) = DY3(vz, iz + O, iy, ix + O, dyI);         /* @d_yi */
    }
}
static void compute_stressii_3d(arr* w, arr dtM, arr dtlambda) {
    arr tauxx = w[GPI_TAUXX], tauyy = w[GPI_TAUYY], tauzz = w[GPI_TAUZZ], dvxdx = w[GPI_DVXDX], dvydy = w[GPI_DVYDY], dvzdz = w[GPI_DVZDZ];
    OMP_FOR
    for (int ix = 1; ix <= tauxx.n[2]; ix++) for (int iy = 1; iy <= tauxx.n[1]; iy++) for (int iz = 1; iz <= tauxx.n[0]; iz++) {
        REAL M = A3(dtM, iz, iy, ix), L = A3(dtlambda, iz, iy, ix);
        REAL dxx = A3(dvxdx, iz, iy, ix), dyy = A3(dvydy, iz, iy, ix), dzz = A3(dvzdz, iz, iy, ix);
        A3(tauxx, iz, iy, ix) = A3(tauxx, iz, iy, ix) - (M * dxx) - (L * (dyy + dzz));
        A3(tauyy, iz, iy, ix) = A3(tauyy, iz, iy, ix) - (M * dyy) - (L * (dxx + dzz));
        A3(tauzz, iz, iy, ix) = A3(tauzz, iz, iy, ix) - (M * dzz) - (L * (dyy + dxx));
    }
}
static void compute_stressij_3d(arr* w, arr muxz, arr muxy, arr muyz) {
    arr tauxy = w[GPI_TAUXY], tauxz = w[GPI_TAUXZ], tauyz = w[GPI_TAUYZ];
    arr dvxdy = w[GPI_DVXDY], dvxdz = w[GPI_DVXDZ], dvydx = w[GPI_DVYDX], dvydz = w[GPI_DVYDZ], dvzdx = w[GPI_DVZDX], dvzdy = w[GPI_DVZDY];
    int nz = imax3(tauxy.n[0], tauxz.n[0], tauyz.n[0]), ny = imax3(tauxy.n[1], tauxz.n[1], tauyz.n[1]), nx = imax3(tauxy.n[2], tauxz.n[2], tauyz.n[2]);
    OMP_FOR
    for (int ix = 1; ix <= nx; ix++) for (int iy = 1; iy <= ny; iy++) for (int iz = 1; iz <= nz; iz++) {
        if (IN3(tauxz)) A3(tauxz, iz, iy, ix) = A3(tauxz, iz, iy, ix) - A3(muxz, iz, iy, ix) * (A3(dvxdz, iz, iy, ix) + A3(dvzdx, iz, iy, ix));
        if (IN3(tauxy)) A3(tauxy, iz, iy, ix) = A3(tauxy, iz, iy, ix) - A3(muxy, iz, iy, ix) * (A3(dvxdy, iz, iy, ix) + A3(dvydx, iz, iy, ix));
        if (IN3(tauyz)) A3(tauyz, iz, iy, ix) = A3(tauyz, iz, iy, ix) - A3(muyz, iz, iy, ix) * (A3(dvydz, iz, iy, ix) + A3(dvzdy, iz, iy, ix));
    }
}

/* ------------------------------------------------------------------------------------------------
 * CPML memory variables: src/fdtd/cpml.jl:158-215 (memory{z,y,x}! -> memorynp*)
 *   m[.., i+moff, ..] = b[i+moff]*m + a[i+moff]*d[.., doff+i, ..];  d = d*kI[i+moff] + m
 * ---------------------------------------------------------------------------------------------- */
static void memory_np(arr m, arr d, const REAL* a, const REAL* b, const REAL* kI, int axis, int npml, int moff, int doff) {
    int sm[3] = {m.n[0], m.n[1], m.n[2]};
    sm[axis] = npml;
    OMP_FOR
    for (int ix = 1; ix <= sm[2]; ix++) for (int iy = 1; iy <= sm[1]; iy++) for (int iz = 1; iz <= sm[0]; iz++) {
        int im[3] = {iz, iy, ix}, id[3] = {iz, iy, ix};
        int i = im[axis];
        im[axis] = i + moff; id[axis] = doff + i;
        REAL* mm = &A3(m, im[0], im[1], im[2]);
        REAL* dd = &A3(d, id[0], id[1], id[2]);
        *mm = b[i + moff - 1] * *mm + a[i + moff - 1] * *dd;
        *dd = *dd * kI[i + moff - 1] + *mm;
    }
}
static void memory_pml(orc_handle* h, pw_t* pw, int df) {
    int axis = dfield_axis(df), npml = h->c.npml;
    int minbit = axis == 0 ? GPI_ZMIN : axis == 1 ? GPI_YMIN : GPI_XMIN;
    int maxbit = axis == 0 ? GPI_ZMAX : axis == 1 ? GPI_YMAX : GPI_XMAX;
    if (h->c.pml_faces & minbit) memory_np(pw->mem[df], pw->w[df], h->pa[df], h->pb[df], h->pk[df], axis, npml, 0, 0);
    if (h->c.pml_faces & maxbit) memory_np(pw->mem[df], pw->w[df], h->pa[df], h->pb[df], h->pk[df], axis, npml, npml, pw->w[df].n[axis] - npml);
}

/* ------------------------------------------------------------------------------------------------
 * rigid faces: src/fdtd/dirichlet.jl:3-78 (order/2 ghost pairs)
 *   dirichlet{q}min!(vq, vrest..., n): vrest[.., 1, ..] = 0; vq[.., 1, ..] = -vq[.., 2, ..]
 *   dirichlet{q}max!(vq, vrest..., n): vrest[.., n, ..] = 0; vq[.., n+1, ..] = -vq[.., n, ..]
 *   launched over (1:n_other...) of the tauii grid (advance_acou.jl:51-58,161-193)
 * ---------------------------------------------------------------------------------------------- */
static void dirichlet(orc_handle* h, pw_t* pw) {
    int nn[3] = {h->nz, h->ny, h->nx};
    int vq[3] = {GPI_VZ, GPI_VY, GPI_VX};
    /* call order of the reference: xmin, xmax, (ymin, ymax,) zmin, zmax */
    const int axes[3] = {2, 1, 0};
    for (int ia = 0; ia < 3; ia++) {
        int q = axes[ia];
        if (q == 1 && h->nd == 2) continue;
        int minbit = q == 0 ? GPI_ZMIN : q == 1 ? GPI_YMIN : GPI_XMIN;
        int maxbit = q == 0 ? GPI_ZMAX : q == 1 ? GPI_YMAX : GPI_XMAX;
        for (int side = 0; side < 2; side++) {
            if (!(h->c.rigid_faces & (side ? maxbit : minbit))) continue;
            int n = nn[q];
            int o1 = (q + 1) % 3, o2 = (q + 2) % 3;
            for (int i2 = 1; i2 <= nn[o2]; i2++) for (int i1 = 1; i1 <= nn[o1]; i1++) {
                int id[3]; id[o1] = i1; id[o2] = i2;
                for (int r = 0; r < 3; r++) {              /* tangential components */
                    if (r == q || (r == 1 && h->nd == 2)) continue;
                    id[q] = side ? n : 1;
                    A3(pw->w[vq[r]], id[0], id[1], id[2]) = 0;
                }
                const int order = g_O + 1, fdh = order / 2;       /* dirichlet.jl:12-24: ghost <- -mirror, ifd = 1..fdh */
                for (int ifd = 1; ifd <= fdh; ifd++) {
                    int ig = side ? n + order - ifd : ifd, is = side ? n + ifd - 1 : order + 1 - ifd;
                    id[q] = is; REAL v = A3(pw->w[vq[q]], id[0], id[1], id[2]);
                    id[q] = ig; A3(pw->w[vq[q]], id[0], id[1], id[2]) = -v;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * the four phases of a step for one wavefield (update_dstress!/update_v!/update_dv!/update_stress!)
 * ---------------------------------------------------------------------------------------------- */
static void update_dstress(orc_handle* h, pw_t* pw) {
    arr* w = pw->w;
    if (h->c.physics == GPI_ACOUSTIC) {
        if (h->nd == 2) {   /* advance_acou.jl:11-35 */
            compute_dp_2d(w[GPI_P], w[GPI_DPDX], w[GPI_DPDZ], h->dzI, h->dxI);
            memory_pml(h, pw, GPI_DPDX); memory_pml(h, pw, GPI_DPDZ);
        } else {            /* advance_acou.jl:98-138 */
            compute_dp_3d(w[GPI_P], w[GPI_DPDX], w[GPI_DPDY], w[GPI_DPDZ], h->dzI, h->dyI, h->dxI);
            memory_pml(h, pw, GPI_DPDX); memory_pml(h, pw, GPI_DPDY); memory_pml(h, pw, GPI_DPDZ);
        }
    } else {
        if (h->nd == 2) {   /* advance_elastic.jl:534-585 */
            compute_dstress_2d(w[GPI_TAUXX], w[GPI_TAUZZ], w[GPI_TAUXZ], w[GPI_DTAUXXDX], w[GPI_DTAUXZDX], w[GPI_DTAUZZDZ], w[GPI_DTAUXZDZ], h->dxI, h->dzI);
            memory_pml(h, pw, GPI_DTAUXXDX); memory_pml(h, pw, GPI_DTAUXZDX);
            memory_pml(h, pw, GPI_DTAUZZDZ); memory_pml(h, pw, GPI_DTAUXZDZ);
        } else {            /* advance_elastic.jl:233-334 */
            compute_dstress_3d(w, h->dxI, h->dyI, h->dzI);
            memory_pml(h, pw, GPI_DTAUXXDX); memory_pml(h, pw, GPI_DTAUXYDX); memory_pml(h, pw, GPI_DTAUXZDX);
            memory_pml(h, pw, GPI_DTAUYYDY); memory_pml(h, pw, GPI_DTAUXYDY); memory_pml(h, pw, GPI_DTAUYZDY);
            memory_pml(h, pw, GPI_DTAUZZDZ); memory_pml(h, pw, GPI_DTAUYZDZ); memory_pml(h, pw, GPI_DTAUXZDZ);
        }
    }
}
static void update_v(orc_handle* h, pw_t* pw) {
    arr* w = pw->w;
    if (h->c.physics == GPI_ACOUSTIC) {
        if (h->nd == 2) compute_v_acou_2d(w[GPI_VX], w[GPI_VZ], h->dmod[DM_BX], h->dmod[DM_BZ], w[GPI_DPDX], w[GPI_DPDZ]);
        else compute_v_acou_3d(w[GPI_VX], w[GPI_VY], w[GPI_VZ], h->dmod[DM_BX], h->dmod[DM_BY], h->dmod[DM_BZ], w[GPI_DPDX], w[GPI_DPDY], w[GPI_DPDZ]);
    } else {
        if (h->nd == 2) compute_v_el_2d(w[GPI_VX], w[GPI_VZ], w[GPI_DTAUXXDX], w[GPI_DTAUXZDX], w[GPI_DTAUZZDZ], w[GPI_DTAUXZDZ], h->dmod[DM_BX], h->dmod[DM_BZ]);
        else compute_v_el_3d(w, h->dmod[DM_BX], h->dmod[DM_BY], h->dmod[DM_BZ]);
    }
    dirichlet(h, pw);
}
static void update_dv(orc_handle* h, pw_t* pw) {
    arr* w = pw->w;
    if (h->c.physics == GPI_ACOUSTIC) {
        if (h->nd == 2) {   /* advance_acou.jl:59-89 */
            compute_dv_acou_2d(w[GPI_VX], w[GPI_VZ], w[GPI_DVXDX], w[GPI_DVZDZ], h->dxI, h->dzI);
            memory_pml(h, pw, GPI_DVXDX); memory_pml(h, pw, GPI_DVZDZ);
        } else {            /* advance_acou.jl:196-236 */
            compute_dv_acou_3d(w[GPI_VX], w[GPI_VY], w[GPI_VZ], w[GPI_DVXDX], w[GPI_DVYDY], w[GPI_DVZDZ], h->dxI, h->dyI, h->dzI);
            memory_pml(h, pw, GPI_DVXDX); memory_pml(h, pw, GPI_DVYDY); memory_pml(h, pw, GPI_DVZDZ);
        }
    } else {
        if (h->nd == 2) {   /* advance_elastic.jl:609-657 */
            compute_dv_el_2d(w[GPI_VX], w[GPI_VZ], w[GPI_DVXDX], w[GPI_DVZDZ], w[GPI_DVXDZ], w[GPI_DVZDX], h->dxI, h->dzI);
            memory_pml(h, pw, GPI_DVXDX); memory_pml(h, pw, GPI_DVZDZ); memory_pml(h, pw, GPI_DVXDZ); memory_pml(h, pw, GPI_DVZDX);
        } else {            /* advance_elastic.jl:395-492 */
            compute_dv_el_3d(w, h->dxI, h->dyI, h->dzI);
            memory_pml(h, pw, GPI_DVXDX); memory_pml(h, pw, GPI_DVYDY); memory_pml(h, pw, GPI_DVZDZ);
            memory_pml(h, pw, GPI_DVXDY); memory_pml(h, pw, GPI_DVXDZ); memory_pml(h, pw, GPI_DVYDX);
            memory_pml(h, pw, GPI_DVYDZ); memory_pml(h, pw, GPI_DVZDX); memory_pml(h, pw, GPI_DVZDY);
        }
    }
}
static void update_stress(orc_handle* h, pw_t* pw) {
    arr* w = pw->w;
    if (h->c.physics == GPI_ACOUSTIC) {
        if (h->nd == 2) compute_p_2d(w[GPI_P], w[GPI_DVXDX], w[GPI_DVZDZ], h->dmod[DM_DTK]);
        else compute_p_3d(w[GPI_P], w[GPI_DVXDX], w[GPI_DVYDY], w[GPI_DVZDZ], h->dmod[DM_DTK]);
        return;   /* acoustic ignores stressfree_faces (advance_acou.jl:90-96,237-250) */
    }
    if (h->nd == 2) {
        compute_stressii_2d(w[GPI_TAUXX], w[GPI_TAUZZ], w[GPI_DVXDX], w[GPI_DVZDZ], h->dmod[DM_DTM], h->dmod[DM_DTLAMBDA]);
        compute_stressij_2d(w[GPI_TAUXZ], w[GPI_DVXDZ], w[GPI_DVZDX], h->dmod[DM_MUXZ]);
        if (h->c.stressfree_faces & GPI_ZMIN) {    /* advance_elastic.jl:677-680 */
            for (int ix = 1; ix <= w[GPI_TAUZZ].n[2]; ix++) A2(w[GPI_TAUZZ], 1, ix) = -A2(w[GPI_TAUZZ], 2, ix);
            for (int ix = 1; ix <= w[GPI_TAUXZ].n[2]; ix++) A2(w[GPI_TAUXZ], 1, ix) = 0;
        }
    } else {
        compute_stressii_3d(w, h->dmod[DM_DTM], h->dmod[DM_DTLAMBDA]);
        compute_stressij_3d(w, h->dmod[DM_MUXZ], h->dmod[DM_MUXY], h->dmod[DM_MUYZ]);
        if (h->c.stressfree_faces & GPI_ZMIN) {    /* advance_elastic.jl:519-529 */
            arr t = w[GPI_TAUZZ];
            for (int ix = 1; ix <= t.n[2]; ix++) for (int iy = 1; iy <= t.n[1]; iy++) A3(t, 1, iy, ix) = -A3(t, 2, iy, ix);
            t = w[GPI_TAUXZ];
            for (int ix = 1; ix <= t.n[2]; ix++) for (int iy = 1; iy <= t.n[1]; iy++) A3(t, 1, iy, ix) = 0;
            t = w[GPI_TAUYZ];
            for (int ix = 1; ix <= t.n[2]; ix++) for (int iy = 1; iy <= t.n[1]; iy++) A3(t, 1, iy, ix) = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * medium: update_dmod! + store_invav*! (src/fdtd/medium.jl:143-221)
 * `dt / @av_*(b)`: the average is a Float32 sum times a Float64 literal (0.5/0.25), so the
 * quotient is evaluated in Float64 and rounded on store.
 * ---------------------------------------------------------------------------------------------- */
static void update_dmod(orc_handle* h) {
    WIDE dt = (WIDE)h->dt;
    const int O = g_O;                      /* izi = iz + O (diff2D.jl:5-13); the `+ 1` neighbours of the @av macros do not scale with the order */
    arr rho = h->mod[GPI_RHO];
    if (h->nd == 2) {
        arr bx = h->dmod[DM_BX], bz = h->dmod[DM_BZ];
        for (int ix = 1; ix <= bx.n[2]; ix++) for (int iz = 1; iz <= bx.n[0]; iz++)     /* store_invavxi!: @av_xi */
            A2(bx, iz, ix) = (REAL)(dt / ((WIDE)(REAL)(A2(rho, iz + O, ix) + A2(rho, iz + O, ix + 1)) * (WIDE)0.5));
        for (int ix = 1; ix <= bz.n[2]; ix++) for (int iz = 1; iz <= bz.n[0]; iz++)     /* store_invavzi!: @av_zi */
            A2(bz, iz, ix) = (REAL)(dt / ((WIDE)(REAL)(A2(rho, iz, ix + O) + A2(rho, iz + 1, ix + O)) * (WIDE)0.5));
    } else {
        arr bx = h->dmod[DM_BX], by = h->dmod[DM_BY], bz = h->dmod[DM_BZ];
        for (int ix = 1; ix <= bx.n[2]; ix++) for (int iy = 1; iy <= bx.n[1]; iy++) for (int iz = 1; iz <= bx.n[0]; iz++)
            A3(bx, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(rho, iz + O, iy + O, ix) + A3(rho, iz + O, iy + O, ix + 1)) * (WIDE)0.5));
        for (int ix = 1; ix <= by.n[2]; ix++) for (int iy = 1; iy <= by.n[1]; iy++) for (int iz = 1; iz <= by.n[0]; iz++)
            A3(by, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(rho, iz + O, iy, ix + O) + A3(rho, iz + O, iy + 1, ix + O)) * (WIDE)0.5));
        for (int ix = 1; ix <= bz.n[2]; ix++) for (int iy = 1; iy <= bz.n[1]; iy++) for (int iz = 1; iz <= bz.n[0]; iz++)
            A3(bz, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(rho, iz, iy + O, ix + O) + A3(rho, iz + 1, iy + O, ix + O)) * (WIDE)0.5));
    }
    if (h->c.physics == GPI_ACOUSTIC) {
        arr K = h->dmod[DM_DTK], iK = h->mod[GPI_INVK];    /* broadcast!(inv, dtK, invK); rmul!(dtK, dt) */
        for (size_t i = 0; i < K.len; i++) K.d[i] = ((REAL)1 / iK.d[i]) * h->dt;
        return;
    }
    arr il = h->mod[GPI_INVLAMBDA], im = h->mod[GPI_INVMU], L = h->dmod[DM_DTLAMBDA], M = h->dmod[DM_DTM];
    for (size_t i = 0; i < L.len; i++) L.d[i] = ((REAL)1 / il.d[i]) * h->dt;
    for (size_t i = 0; i < M.len; i++)   /* inv(invl) + 2.0 * inv(invmu): Float64 sum, rounded on store; then rmul!(dt) */
        M.d[i] = (REAL)((double)((REAL)1 / il.d[i]) + 2.0 * (double)((REAL)1 / im.d[i])) * h->dt;
    if (h->nd == 2) {          /* store_invav!: @av = 4-point mean * (WIDE)0.25 (diff2D.jl:222-225) */
        arr mu = h->dmod[DM_MUXZ];
        for (int ix = 1; ix <= mu.n[2]; ix++) for (int iz = 1; iz <= mu.n[0]; iz++)
            A2(mu, iz, ix) = (REAL)(dt / ((WIDE)(REAL)(A2(im, iz, ix) + A2(im, iz + 1, ix) + A2(im, iz, ix + 1) + A2(im, iz + 1, ix + 1)) * (WIDE)0.25));
    } else {                   /* @av_xzi / @av_xyi / @av_yzi (diff3D.jl:339-378) */
        arr m1 = h->dmod[DM_MUXZ], m2 = h->dmod[DM_MUXY], m3 = h->dmod[DM_MUYZ];
        for (int ix = 1; ix <= m1.n[2]; ix++) for (int iy = 1; iy <= m1.n[1]; iy++) for (int iz = 1; iz <= m1.n[0]; iz++)
            A3(m1, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(im, iz, iy + O, ix) + A3(im, iz + 1, iy + O, ix) + A3(im, iz, iy + O, ix + 1) + A3(im, iz + 1, iy + O, ix + 1)) * (WIDE)0.25));
        for (int ix = 1; ix <= m2.n[2]; ix++) for (int iy = 1; iy <= m2.n[1]; iy++) for (int iz = 1; iz <= m2.n[0]; iz++)
            A3(m2, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(im, iz + O, iy, ix) + A3(im, iz + O, iy + 1, ix) + A3(im, iz + O, iy, ix + 1) + A3(im, iz + O, iy + 1, ix + 1)) * (WIDE)0.25));
        for (int ix = 1; ix <= m3.n[2]; ix++) for (int iy = 1; iy <= m3.n[1]; iy++) for (int iz = 1; iz <= m3.n[0]; iz++)
            A3(m3, iz, iy, ix) = (REAL)(dt / ((WIDE)(REAL)(A3(im, iz, iy, ix + O) + A3(im, iz + 1, iy, ix + O) + A3(im, iz, iy + 1, ix + O) + A3(im, iz + 1, iy + 1, ix + O)) * (WIDE)0.25));
    }
}

/* ------------------------------------------------------------------------------------------------
 * sources: src/fdtd/source.jl:61-177.  buf = S*w (dense overwrite, SparseArrays CSC mul!),
 * then a full-grid muladd sweep.
 * ---------------------------------------------------------------------------------------------- */
static void spmv(arr buf, const csc* S, const REAL* w, int nt, int it) {
    arr_zero(&buf);
    for (int j = 0; j < S->ncol; j++) {
        REAL xj = w[(size_t)(it - 1) + (size_t)nt * j];
        for (int64_t k = S->colptr[j] - 1; k < S->colptr[j + 1] - 1; k++)
            buf.d[S->rowval[k] - 1] += S->nzval[k] * xj;
    }
}
/* muladd_with_density_v{x,y,z}! (source.jl:166-177): @inn(pw) += @inn(pv) / @av_?i(rho) * dt,
 * evaluated in Float64 because @av_?i carries a Float64 literal. */
static void muladd_with_density(orc_handle* h, arr pw, arr pv, int vfield) {
    arr rho = h->mod[GPI_RHO]; WIDE dt = (WIDE)h->dt;
    const int O = g_O;
    if (h->nd == 2) {
        OMP_FOR
        for (int ix = 1; ix <= pw.n[2] - 2 * O; ix++) for (int iz = 1; iz <= pw.n[0] - 2 * O; iz++) {
            REAL s = vfield == GPI_VX ? (REAL)(A2(rho, iz + O, ix) + A2(rho, iz + O, ix + 1)) : (REAL)(A2(rho, iz, ix + O) + A2(rho, iz + 1, ix + O));
            A2(pw, iz + O, ix + O) = (REAL)((WIDE)A2(pw, iz + O, ix + O) + ((WIDE)A2(pv, iz + O, ix + O) / ((WIDE)s * (WIDE)0.5) * dt));
        }
    } else {
        OMP_FOR
        for (int ix = 1; ix <= pw.n[2] - 2 * O; ix++) for (int iy = 1; iy <= pw.n[1] - 2 * O; iy++) for (int iz = 1; iz <= pw.n[0] - 2 * O; iz++) {
            REAL s = vfield == GPI_VX ? (REAL)(A3(rho, iz + O, iy + O, ix) + A3(rho, iz + O, iy + O, ix + 1))
                   : vfield == GPI_VY ? (REAL)(A3(rho, iz + O, iy, ix + O) + A3(rho, iz + O, iy + 1, ix + O))
                                      : (REAL)(A3(rho, iz, iy + O, ix + O) + A3(rho, iz + 1, iy + O, ix + O));
            I3(pw) = (REAL)((WIDE)I3(pw) + ((WIDE)I3(pv) / ((WIDE)s * (WIDE)0.5) * dt));
        }
    }
}
static void muladd_tauii(arr pw, arr pv, arr dtK) {      /* source.jl:160-163 */
    OMP_FOR
    for (size_t i = 0; i < pw.len; i++) pw.d[i] = pw.d[i] + (pv.d[i] * dtK.d[i]);
}
/* add_velocity_source! (source.jl:124-157) */
static void add_velocity_source(orc_handle* h, int it, int issp, int activepw, int src_flags) {
    const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ};
    if (activepw & 1) for (int i = 0; i < 3; i++) {
        int f = vf[i]; shot_t* s = &h->pw[0].ss[issp];
        if (!s->wavelets[f] || !(src_flags & 1)) continue;
        spmv(h->pw[0].vbuf[f], &s->spray[f], s->wavelets[f], h->c.nt, it);
        muladd_with_density(h, h->pw[0].w[f], h->pw[0].vbuf[f], f);
    }
    if (activepw & 2) for (int i = 0; i < 3; i++) {      /* adjoint sources through pw 1's receiver matrix */
        int f = vf[i]; shot_t* s = &h->pw[1].ss[issp];
        if (!s->wavelets[f] || !(src_flags & 2)) continue;
        spmv(h->pw[1].vbuf[f], &h->pw[0].ss[issp].interp[f], s->wavelets[f], h->c.nt, it);
        muladd_with_density(h, h->pw[1].w[f], h->pw[1].vbuf[f], f);
    }
}
/* add_stress_source! (source.jl:61-119): pw 1 only */
static void add_stress_source(orc_handle* h, int it, int issp, int activepw, int src_flags) {
    if (!(activepw & 1) || !(src_flags & 1)) return;
    shot_t* s = &h->pw[0].ss[issp]; pw_t* pw = &h->pw[0];
    const int sf[4] = {GPI_P, GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ};
    for (int i = 0; i < 4; i++) {
        int f = sf[i];
        if (!s->wavelets[f] || !field_exists(h->nd, h->c.physics, f)) continue;
        spmv(pw->taubuf, &s->spray[f], s->wavelets[f], h->c.nt, it);
        if (h->c.physics == GPI_ACOUSTIC) muladd_tauii(pw->w[GPI_P], pw->taubuf, h->dmod[DM_DTK]);
        else {
            muladd_tauii(pw->w[GPI_TAUXX], pw->taubuf, h->dmod[DM_DTM]);
            if (h->nd == 3) muladd_tauii(pw->w[GPI_TAUYY], pw->taubuf, h->dmod[DM_DTM]);
            muladd_tauii(pw->w[GPI_TAUZZ], pw->taubuf, h->dmod[DM_DTM]);
        }
    }
}

/* record! (receiver.jl:3-14): rec[it][ir] = transpose(R)[ir,:] * field, SparseArrays order */
static void record(orc_handle* h, int it, int issp, int activepw, const int* fields, int nf) {
    for (int ipw = 0; ipw < h->c.npw; ipw++) {
        if (!(activepw & (1 << ipw))) continue;
        shot_t* s = &h->pw[ipw].ss[issp];
        for (int i = 0; i < nf; i++) {
            int f = fields[i];
            if (!s->records[f]) continue;
            const csc* R = &s->interp[f]; arr fld = h->pw[ipw].w[f];
            for (int ir = 0; ir < R->ncol; ir++) {
                REAL tmp = 0;
                for (int64_t k = R->colptr[ir] - 1; k < R->colptr[ir + 1] - 1; k++) tmp += R->nzval[k] * fld.d[R->rowval[k] - 1];
                s->records[f][(size_t)(it - 1) + (size_t)h->c.nt * ir] = tmp;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * boundary store for time reversal: src/fdtd/boundary.jl:17-306
 * ---------------------------------------------------------------------------------------------- */
static const int* boundary_fields(orc_handle* h, int* nf) {
    static const int ac[1] = {GPI_P}, el[3] = {GPI_TAUXX, GPI_TAUXZ, GPI_TAUZZ};
    /* 3-D elastic: upstream has no boundary_save!/force! method (boundary.jl:215-264); the 2-D construction applied to all six
     * stresses */
    static const int el3[6] = {GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ, GPI_TAUXY, GPI_TAUXZ, GPI_TAUYZ};
    if (h->c.physics == GPI_ACOUSTIC) { *nf = 1; return ac; }
    if (h->nd == 3) { *nf = 6; return el3; }
    *nf = 3; return el;
}
/* boundary_half{q}!(d, b, doff, boff): d[.., doff+i, ..] = b[.., i+boff, ..] over size(b) with axis=nbound */
static void boundary_half(REAL* dst, const int dn[3], const REAL* src, const int sn[3], int axis, int nb, int doff, int boff, const int rn[3]) {
    for (int ix = 1; ix <= rn[2]; ix++) for (int iy = 1; iy <= rn[1]; iy++) for (int iz = 1; iz <= rn[0]; iz++) {
        int id[3] = {iz, iy, ix}, is[3] = {iz, iy, ix};
        id[axis] = doff + id[axis]; is[axis] = is[axis] + boff;
        dst[((size_t)id[0] - 1) + (size_t)dn[0] * (((size_t)id[1] - 1) + (size_t)dn[1] * ((size_t)id[2] - 1))] =
            src[((size_t)is[0] - 1) + (size_t)sn[0] * (((size_t)is[1] - 1) + (size_t)sn[1] * ((size_t)is[2] - 1))];
    }
    (void)nb;
}
static void boundary_save(orc_handle* h, int it, int issp) {    /* boundary.jl:217-264 */
    int nf; const int* bf = boundary_fields(h, &nf);
    shot_t* s = &h->pw[0].ss[issp]; int nb = h->c.nbound;
    for (int i = 0; i < nf; i++) {
        int f = bf[i]; arr d = h->pw[0].w[f];
        for (int q = 0; q < 3; q++) {
            if (q == 1 && h->nd == 2) continue;
            int minbit = q == 0 ? GPI_ZMIN : q == 1 ? GPI_YMIN : GPI_XMIN;
            int np = (h->c.pml_faces & minbit) ? h->c.npml : 0;   /* min-face flag also used for the max side */
            int bn[3] = {d.n[0], d.n[1], d.n[2]}; bn[q] = 2 * nb;
            int rn[3] = {bn[0], bn[1], bn[2]}; rn[q] = nb;
            REAL* b = s->bnd[f][q] + (size_t)(it - 1) * s->bnd_slot[f][q];
            boundary_half(b, bn, d.d, d.n, q, nb, 0, np, rn);
            boundary_half(b, bn, d.d, d.n, q, nb, nb, d.n[q] - np - nb, rn);
            for (size_t k = 0; k < s->bnd_slot[f][q]; k++) b[k] = b[k] * (REAL)-1;   /* rmul!(-1) */
        }
    }
}
static void boundary_force(orc_handle* h, int it, int issp) {   /* boundary.jl:113-171 */
    int nf; const int* bf = boundary_fields(h, &nf);
    shot_t* s = &h->pw[0].ss[issp]; int nb = h->c.nbound;
    for (int i = 0; i < nf; i++) {
        int f = bf[i]; arr d = h->pw[0].w[f];
        const int axes[3] = {2, 1, 0};                           /* x, (y,) z */
        for (int ia = 0; ia < 3; ia++) {
            int q = axes[ia];
            if (q == 1 && h->nd == 2) continue;
            int minbit = q == 0 ? GPI_ZMIN : q == 1 ? GPI_YMIN : GPI_XMIN;
            int np = (h->c.pml_faces & minbit) ? h->c.npml : 0;
            int bn[3] = {d.n[0], d.n[1], d.n[2]}; bn[q] = 2 * nb;
            int rn[3] = {bn[0], bn[1], bn[2]}; rn[q] = nb;
            const REAL* b = s->bnd[f][q] + (size_t)(it - 1) * s->bnd_slot[f][q];
            boundary_half(d.d, d.n, b, bn, q, nb, np, 0, rn);
            boundary_half(d.d, d.n, b, bn, q, nb, d.n[q] - np - nb, nb, rn);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * gradient imaging: src/fdtd/gradient.jl:17-61 (2-D acoustic; gradlame! + gradrho!)
 * ---------------------------------------------------------------------------------------------- */
/* 3-D acoustic imaging.  gradlame! (gradient.jl:17-29) is dimension-free; gradrho! exists upstream for 2-D only
 * (gradient.jl:31, compute_gradient!(::Val{:adjoint}, ::Val{2}, ...) :58-61).  The 3-D method here is the same construction with the
 * y term added between x and z:  @inn(g) = @inn(g) - @av_xi(vxbuffer) - @av_yi(vybuffer) - @av_zi(vzbuffer)
 * (new math upstream lacks; checked against finite differences in tests/test_adjoint3d.py). */
static void compute_gradient_3d(orc_handle* h, int issp, int unshifted) {
    pw_t *p1 = &h->pw[0], *p2 = &h->pw[1]; shot_t* s = &p1->ss[issp];
    arr g = s->grad[GPI_INVK], pf = p1->w[GPI_P], pfp = p1->wtp[GPI_P], pap = p2->wtp[GPI_P];
    REAL dtI = h->dtI;
    OMP_FOR
    for (size_t i = 0; i < g.len; i++) g.d[i] = g.d[i] + pap.d[i] * (pfp.d[i] - pf.d[i]) * dtI;
    const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ};
    for (int k = 0; k < 3; k++) {
        arr b = p1->vbuf[vf[k]], v = p1->w[vf[k]], vp = p1->wtp[vf[k]], va = p2->wtp[vf[k]];
        OMP_FOR
        for (size_t i = 0; i < b.len; i++) b.d[i] = va.d[i] * (v.d[i] - vp.d[i]) * dtI;
    }
    arr gr = s->grad[GPI_RHO], bx = p1->vbuf[GPI_VX], by = p1->vbuf[GPI_VY], bz = p1->vbuf[GPI_VZ];
    const int O = g_O;
    if (unshifted) {     /* GPI_RUN_UNSHIFTED_RHO: cell (iz, iy, ix) <- the interior velocity nodes that bound it (vx ix, ix+1; vy iy, iy+1; vz iz, iz+1) */
        #define INT3(b, z, y, x) (((z) >= 2 && (z) <= (b).n[0] - 1 && (y) >= 2 && (y) <= (b).n[1] - 1 && (x) >= 2 && (x) <= (b).n[2] - 1) ? A3(b, z, y, x) : (REAL)0)
        OMP_FOR
        for (int ix = 1; ix <= gr.n[2]; ix++) for (int iy = 1; iy <= gr.n[1]; iy++) for (int iz = 1; iz <= gr.n[0]; iz++)
            A3(gr, iz, iy, ix) = (REAL)((WIDE)A3(gr, iz, iy, ix)
                - (WIDE)(REAL)(INT3(bx, iz, iy, ix) + INT3(bx, iz, iy, ix + 1)) * (WIDE)0.5
                - (WIDE)(REAL)(INT3(by, iz, iy, ix) + INT3(by, iz, iy + 1, ix)) * (WIDE)0.5
                - (WIDE)(REAL)(INT3(bz, iz, iy, ix) + INT3(bz, iz + 1, iy, ix)) * (WIDE)0.5);
        return;
    }
    OMP_FOR
    for (int ix = 1; ix <= gr.n[2] - 2 * O; ix++) for (int iy = 1; iy <= gr.n[1] - 2 * O; iy++) for (int iz = 1; iz <= gr.n[0] - 2 * O; iz++)
        I3(gr) = (REAL)((WIDE)I3(gr)
            - (WIDE)(REAL)(A3(bx, iz + O, iy + O, ix) + A3(bx, iz + O, iy + O, ix + 1)) * (WIDE)0.5
            - (WIDE)(REAL)(A3(by, iz + O, iy, ix + O) + A3(by, iz + O, iy + 1, ix + O)) * (WIDE)0.5
            - (WIDE)(REAL)(A3(bz, iz, iy + O, ix + O) + A3(bz, iz + 1, iy + O, ix + O)) * (WIDE)0.5);
}
/* 2-D elastic imaging (nothing upstream: gradient.jl has acoustic methods only).  Same adjoint-state
 * construction as gradlame!/gradrho! -- "adjoint field at the previous step times the change of the forward field over the
 * step, times dtI" -- written for the compliance form of the stress update, S dtau/dt = -eps(v):
 *     e   = (txx2 + tzz2)_tp * ((txx1 + tzz1)_tp - (txx1 + tzz1)) * dtI        (isotropic part,  dS = dc/4,  c = 1/(lambda+mu))
 *     d   = (txx2 - tzz2)_tp * ((txx1 - tzz1)_tp - (txx1 - tzz1)) * dtI        (deviatoric part, dS = dinvmu/4)
 *     s   =  txz2_tp * (txz1_tp - txz1) * dtI   on the shear grid, spread back to the four cells its @av(invmu) averages
 *     g_invlambda += e/4 * dc/dinvlambda,   g_invmu += e/4 * dc/dinvmu + d/4 + sum_4 s/4
 * with c = invlambda*invmu/(invlambda+invmu).  g_rho is gradrho! unchanged (it only involves the velocities).
 * Checked against finite differences of the loss in tests/test_adjoint3d.py. */
static void compute_gradient_el2d(orc_handle* h, int issp) {
    pw_t *p1 = &h->pw[0], *p2 = &h->pw[1]; shot_t* s = &p1->ss[issp];
    arr gl = s->grad[GPI_INVLAMBDA], gm = s->grad[GPI_INVMU], gr = s->grad[GPI_RHO];
    arr il = h->mod[GPI_INVLAMBDA], im = h->mod[GPI_INVMU];
    arr xx1 = p1->w[GPI_TAUXX], zz1 = p1->w[GPI_TAUZZ], xz1 = p1->w[GPI_TAUXZ];
    arr xx1p = p1->wtp[GPI_TAUXX], zz1p = p1->wtp[GPI_TAUZZ], xz1p = p1->wtp[GPI_TAUXZ];
    arr xx2p = p2->wtp[GPI_TAUXX], zz2p = p2->wtp[GPI_TAUZZ], xz2p = p2->wtp[GPI_TAUXZ];
    REAL dtI = h->dtI;
    const int O = g_O;
    OMP_FOR
    for (int ix = 1; ix <= gl.n[2]; ix++) for (int iz = 1; iz <= gl.n[0]; iz++) {
        const REAL e = (A2(xx2p, iz, ix) + A2(zz2p, iz, ix)) * ((A2(xx1p, iz, ix) + A2(zz1p, iz, ix)) - (A2(xx1, iz, ix) + A2(zz1, iz, ix))) * dtI;
        const REAL d = (A2(xx2p, iz, ix) - A2(zz2p, iz, ix)) * ((A2(xx1p, iz, ix) - A2(zz1p, iz, ix)) - (A2(xx1, iz, ix) - A2(zz1, iz, ix))) * dtI;
        const REAL a = A2(il, iz, ix), b = A2(im, iz, ix), ab = a + b;
        const REAL dca = (b * b) / (ab * ab), dcb = (a * a) / (ab * ab);
        A2(gl, iz, ix) = A2(gl, iz, ix) + (REAL)0.25 * e * dca;
        A2(gm, iz, ix) = A2(gm, iz, ix) + (REAL)0.25 * e * dcb + (REAL)0.25 * d;
    }
    /* shear part: tauxz[jz, jx] uses @av(invmu) of the cells (jz, jx), (jz+1, jx), (jz, jx+1), (jz+1, jx+1) in array indices, whatever
     * the order (diff2D.jl:222-225): every cell collects a quarter of the (up to) four shear nodes whose average contains it */
    const int hh = 0;
    OMP_FOR
    for (int ix = 1; ix <= gm.n[2]; ix++) for (int iz = 1; iz <= gm.n[0]; iz++) {
        REAL acc = 0;
        for (int dx = 0; dx < 2; dx++) for (int dz = 0; dz < 2; dz++) {
            const int jz = iz - dz - hh, jx = ix - dx - hh;          /* shear node whose average contains cell (iz, ix) */
            if (jz < 1 || jz > xz1.n[0] || jx < 1 || jx > xz1.n[2]) continue;
            acc = acc + A2(xz2p, jz, jx) * (A2(xz1p, jz, jx) - A2(xz1, jz, jx)) * dtI;
        }
        A2(gm, iz, ix) = A2(gm, iz, ix) + (REAL)0.25 * acc;
    }
    const int vf[2] = {GPI_VX, GPI_VZ};
    for (int k = 0; k < 2; k++) {
        arr b = p1->vbuf[vf[k]], v = p1->w[vf[k]], vp = p1->wtp[vf[k]], va = p2->wtp[vf[k]];
        OMP_FOR
        for (size_t i = 0; i < b.len; i++) b.d[i] = va.d[i] * (v.d[i] - vp.d[i]) * dtI;
    }
    arr bx = p1->vbuf[GPI_VX], bz = p1->vbuf[GPI_VZ];
    OMP_FOR
    for (int ix = 1; ix <= gr.n[2] - 2 * O; ix++) for (int iz = 1; iz <= gr.n[0] - 2 * O; iz++)
        A2(gr, iz + O, ix + O) = (REAL)((WIDE)A2(gr, iz + O, ix + O)
            - (WIDE)(REAL)(A2(bx, iz + O, ix) + A2(bx, iz + O, ix + 1)) * (WIDE)0.5
            - (WIDE)(REAL)(A2(bz, iz, ix + O) + A2(bz, iz + 1, ix + O)) * (WIDE)0.5);
}
/* 3-D elastic imaging: the construction of compute_gradient_el2d with the 3-D isotropic compliance
 *     eps = dev(tau) / (2 mu) + tr(tau) / (3 (3 lambda + 2 mu)) I,      c3 = 1 / (3 lambda + 2 mu) = invlambda invmu / (3 invmu + 2 invlambda)
 *     eT  = T2_tp * (T1_tp - T1) * dtI,  T = txx + tyy + tzz                      g_invlambda += eT/3 * dc3/dinvlambda
 *     eD  = sum_i dev2_ii_tp * (dev1_ii_tp - dev1_ii) * dtI,  dev_ii = t_ii - T/3   g_invmu     += eT/3 * dc3/dinvmu + eD/2 + shear
 *     shear: t_ij2_tp * (t_ij1_tp - t_ij1) * dtI on the three shear grids, a quarter to each cell of the node's @av_??i(invmu) */
static void compute_gradient_el3d(orc_handle* h, int issp) {
    pw_t *p1 = &h->pw[0], *p2 = &h->pw[1]; shot_t* s = &p1->ss[issp];
    arr gl = s->grad[GPI_INVLAMBDA], gm = s->grad[GPI_INVMU], gr = s->grad[GPI_RHO];
    arr il = h->mod[GPI_INVLAMBDA], im = h->mod[GPI_INVMU];
    REAL dtI = h->dtI;
    const int O = g_O;
    const int nf[3] = {GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ};
    OMP_FOR
    for (size_t i = 0; i < gl.len; i++) {
        REAL t1 = 0, t1p = 0, t2p = 0;
        for (int q = 0; q < 3; q++) { t1 = t1 + p1->w[nf[q]].d[i]; t1p = t1p + p1->wtp[nf[q]].d[i]; t2p = t2p + p2->wtp[nf[q]].d[i]; }
        const REAL third = (REAL)1 / (REAL)3;
        const REAL eT = t2p * (t1p - t1) * dtI;
        REAL eD = 0;
        for (int q = 0; q < 3; q++) {
            const REAL d1 = p1->w[nf[q]].d[i] - t1 * third, d1p = p1->wtp[nf[q]].d[i] - t1p * third, d2p = p2->wtp[nf[q]].d[i] - t2p * third;
            eD = eD + d2p * (d1p - d1) * dtI;
        }
        const REAL a = il.d[i], b = im.d[i], den = (REAL)3 * b + (REAL)2 * a;
        const REAL dca = ((REAL)3 * (b * b)) / (den * den), dcb = ((REAL)2 * (a * a)) / (den * den);
        gl.d[i] = gl.d[i] + third * eT * dca;
        gm.d[i] = gm.d[i] + third * eT * dcb + (REAL)0.5 * eD;
    }
    /* shear nodes -> cells: tauxz[jz, jy, jx] averages invmu[jz..jz+1, jy+O, jx..jx+1] (@av_xzi), tauxy[..] invmu[jz+O, jy..jy+1, jx..jx+1]
     * (@av_xyi), tauyz[..] invmu[jz..jz+1, jy..jy+1, jx+O] (@av_yzi)  (diff3D.jl:339-378) */
    arr xz1 = p1->w[GPI_TAUXZ], xz1p = p1->wtp[GPI_TAUXZ], xz2p = p2->wtp[GPI_TAUXZ];
    arr xy1 = p1->w[GPI_TAUXY], xy1p = p1->wtp[GPI_TAUXY], xy2p = p2->wtp[GPI_TAUXY];
    arr yz1 = p1->w[GPI_TAUYZ], yz1p = p1->wtp[GPI_TAUYZ], yz2p = p2->wtp[GPI_TAUYZ];
    #define INA(a, z, y, x) ((z) >= 1 && (z) <= (a).n[0] && (y) >= 1 && (y) <= (a).n[1] && (x) >= 1 && (x) <= (a).n[2])
    OMP_FOR
    for (int ix = 1; ix <= gm.n[2]; ix++) for (int iy = 1; iy <= gm.n[1]; iy++) for (int iz = 1; iz <= gm.n[0]; iz++) {
        REAL acc = 0;
        for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {      /* d1: first averaged axis, d2: second */
            int jz = iz - d1, jy = iy - O, jx = ix - d2;                      /* tauxz: (z, x) averaged */
            if (INA(xz1, jz, jy, jx)) acc = acc + A3(xz2p, jz, jy, jx) * (A3(xz1p, jz, jy, jx) - A3(xz1, jz, jy, jx)) * dtI;
        }
        for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {
            int jz = iz - O, jy = iy - d1, jx = ix - d2;                      /* tauxy: (y, x) averaged */
            if (INA(xy1, jz, jy, jx)) acc = acc + A3(xy2p, jz, jy, jx) * (A3(xy1p, jz, jy, jx) - A3(xy1, jz, jy, jx)) * dtI;
        }
        for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {
            int jz = iz - d1, jy = iy - d2, jx = ix - O;                      /* tauyz: (z, y) averaged */
            if (INA(yz1, jz, jy, jx)) acc = acc + A3(yz2p, jz, jy, jx) * (A3(yz1p, jz, jy, jx) - A3(yz1, jz, jy, jx)) * dtI;
        }
        A3(gm, iz, iy, ix) = A3(gm, iz, iy, ix) + (REAL)0.25 * acc;
    }
    const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ};
    for (int k = 0; k < 3; k++) {
        arr b = p1->vbuf[vf[k]], v = p1->w[vf[k]], vp = p1->wtp[vf[k]], va = p2->wtp[vf[k]];
        OMP_FOR
        for (size_t i = 0; i < b.len; i++) b.d[i] = va.d[i] * (v.d[i] - vp.d[i]) * dtI;
    }
    arr bx = p1->vbuf[GPI_VX], by = p1->vbuf[GPI_VY], bz = p1->vbuf[GPI_VZ];
    OMP_FOR
    for (int ix = 1; ix <= gr.n[2] - 2 * O; ix++) for (int iy = 1; iy <= gr.n[1] - 2 * O; iy++) for (int iz = 1; iz <= gr.n[0] - 2 * O; iz++)
        I3(gr) = (REAL)((WIDE)I3(gr)
            - (WIDE)(REAL)(A3(bx, iz + O, iy + O, ix) + A3(bx, iz + O, iy + O, ix + 1)) * (WIDE)0.5
            - (WIDE)(REAL)(A3(by, iz + O, iy, ix + O) + A3(by, iz + O, iy + 1, ix + O)) * (WIDE)0.5
            - (WIDE)(REAL)(A3(bz, iz, iy + O, ix + O) + A3(bz, iz + 1, iy + O, ix + O)) * (WIDE)0.5);
}
static void compute_gradient(orc_handle* h, int issp, int unshifted) {
    if (h->c.physics == GPI_ELASTIC) { if (h->nd == 3) compute_gradient_el3d(h, issp); else compute_gradient_el2d(h, issp); return; }
    if (h->nd == 3) { compute_gradient_3d(h, issp, unshifted); return; }
    pw_t *p1 = &h->pw[0], *p2 = &h->pw[1]; shot_t* s = &p1->ss[issp];
    arr g = s->grad[GPI_INVK], pf = p1->w[GPI_P], pfp = p1->wtp[GPI_P], pap = p2->wtp[GPI_P];
    REAL dtI = h->dtI;
    OMP_FOR
    for (size_t i = 0; i < g.len; i++) g.d[i] = g.d[i] + pap.d[i] * (pfp.d[i] - pf.d[i]) * dtI;           /* compute_gmodKI! */
    const int vf[2] = {GPI_VX, GPI_VZ};
    for (int k = 0; k < 2; k++) {                                                                             /* compute_gmodrho! */
        arr b = p1->vbuf[vf[k]], v = p1->w[vf[k]], vp = p1->wtp[vf[k]], va = p2->wtp[vf[k]];
        OMP_FOR
        for (size_t i = 0; i < b.len; i++) b.d[i] = va.d[i] * (v.d[i] - vp.d[i]) * dtI;
    }
    arr gr = s->grad[GPI_RHO], bx = p1->vbuf[GPI_VX], bz = p1->vbuf[GPI_VZ];
    if (unshifted) {     /* GPI_RUN_UNSHIFTED_RHO: cell (iz, ix) <- vx nodes ix, ix+1 and vz nodes iz, iz+1 (interior nodes only) */
        OMP_FOR
        for (int ix = 1; ix <= gr.n[2]; ix++) for (int iz = 1; iz <= gr.n[0]; iz++) {
            #define BXI(z, x) (((z) >= 2 && (z) <= bx.n[0] - 1 && (x) >= 2 && (x) <= bx.n[2] - 1) ? A2(bx, z, x) : (REAL)0)
            #define BZI(z, x) (((z) >= 2 && (z) <= bz.n[0] - 1 && (x) >= 2 && (x) <= bz.n[2] - 1) ? A2(bz, z, x) : (REAL)0)
            A2(gr, iz, ix) = (REAL)((WIDE)A2(gr, iz, ix) - (WIDE)(REAL)(BXI(iz, ix) + BXI(iz, ix + 1)) * (WIDE)0.5
                                                             - (WIDE)(REAL)(BZI(iz, ix) + BZI(iz + 1, ix)) * (WIDE)0.5);
        }
        return;
    }
    const int O = g_O;
    OMP_FOR
    for (int ix = 1; ix <= gr.n[2] - 2 * O; ix++) for (int iz = 1; iz <= gr.n[0] - 2 * O; iz++)             /* combine_gmodrho!: Float64 via 0.5 literals */
        A2(gr, iz + O, ix + O) = (REAL)((WIDE)A2(gr, iz + O, ix + O)
            - (WIDE)(REAL)(A2(bx, iz + O, ix) + A2(bx, iz + O, ix + 1)) * (WIDE)0.5
            - (WIDE)(REAL)(A2(bz, iz, ix + O) + A2(bz, iz + 1, ix + O)) * (WIDE)0.5);
}

/* ------------------------------------------------------------------------------------------------
 * mod_x_proc! (src/fdtd/propagate.jl:138-261)
 * ---------------------------------------------------------------------------------------------- */
static void reset_w2(orc_handle* h) {       /* types.jl:100-113 */
    for (int ipw = 0; ipw < h->c.npw; ipw++) {
        pw_t* pw = &h->pw[ipw];
        for (int f = 0; f < GPI_NFIELD; f++) { arr_zero(&pw->w[f]); arr_zero(&pw->mem[f]); }
        for (int f = 0; f < GPI_NWAVEFIELD; f++) { arr_zero(&pw->wtp[f]); arr_zero(&pw->vbuf[f]); }
        arr_zero(&pw->taubuf);
    }
}
static const int WAVEF[GPI_NWAVEFIELD] = {GPI_P, GPI_VX, GPI_VY, GPI_VZ, GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ, GPI_TAUXY, GPI_TAUXZ, GPI_TAUYZ};

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * FD-Born scattering sources, 2-D acoustic (born.jl:1-12 names the intent -- compute_v!(v2, d invrho, dp1, dt),
 * compute_p!(p2, dv1, d K, dt) -- with argument lists that match no kernel upstream; commented legacy code
 * born.jl:32-99).  The coefficient perturbations are linearised in (d invK, d rho):
 *     d(dt K) = -((K K) d invK) dt;   d(dt / av rho) = -(dt av(d rho)) / (av rho)^2  (Float64, rounded once)
 * ---------------------------------------------------------------------------------------------- */
static void update_born(orc_handle* h) {
    double dt = (double)h->dt;
    arr rho = h->mod[GPI_RHO], drho = h->modp[GPI_RHO], iK = h->mod[GPI_INVK], diK = h->modp[GPI_INVK];
    arr dK = h->bornc[0], dbx = h->bornc[1], dbz = h->bornc[2];
    for (size_t i = 0; i < dK.len; i++) { REAL K = (REAL)1 / iK.d[i]; dK.d[i] = -((REAL)((REAL)(K * K) * diK.d[i]) * h->dt); }
    for (int ix = 1; ix <= dbx.n[2]; ix++) for (int iz = 1; iz <= dbx.n[0]; iz++) {
        double av = (double)(REAL)(A2(rho, iz + 1, ix) + A2(rho, iz + 1, ix + 1)) * 0.5, dav = (double)(REAL)(A2(drho, iz + 1, ix) + A2(drho, iz + 1, ix + 1)) * 0.5;
        A2(dbx, iz, ix) = (REAL)(-(dt * dav) / (av * av));
    }
    for (int ix = 1; ix <= dbz.n[2]; ix++) for (int iz = 1; iz <= dbz.n[0]; iz++) {
        double av = (double)(REAL)(A2(rho, iz, ix + 1) + A2(rho, iz + 1, ix + 1)) * 0.5, dav = (double)(REAL)(A2(drho, iz, ix + 1) + A2(drho, iz + 1, ix + 1)) * 0.5;
        A2(dbz, iz, ix) = (REAL)(-(dt * dav) / (av * av));
    }
}
/* add_born_sources_velocity! (propagate.jl:205): @inn(v?2) = @inn(v?2) + @all(d b?) * @all(dpd?1) */
static void born_velocity(orc_handle* h) {
    arr vx = h->pw[1].w[GPI_VX], vz = h->pw[1].w[GPI_VZ], dpdx = h->pw[0].w[GPI_DPDX], dpdz = h->pw[0].w[GPI_DPDZ];
    arr dbx = h->bornc[1], dbz = h->bornc[2];
    OMP_FOR
    for (int ix = 1; ix <= dpdx.n[2]; ix++) for (int iz = 1; iz <= dpdx.n[0]; iz++)
        A2(vx, iz + 1, ix + 1) = A2(vx, iz + 1, ix + 1) + A2(dbx, iz, ix) * A2(dpdx, iz, ix);
    OMP_FOR
    for (int ix = 1; ix <= dpdz.n[2]; ix++) for (int iz = 1; iz <= dpdz.n[0]; iz++)
        A2(vz, iz + 1, ix + 1) = A2(vz, iz + 1, ix + 1) + A2(dbz, iz, ix) * A2(dpdz, iz, ix);
}
/* add_born_sources_stress! (propagate.jl:226): @all(p2) = @all(p2) + (@all(dvxdx1) + @all(dvzdz1)) * @all(d dtK) */
static void born_stress(orc_handle* h) {
    arr p = h->pw[1].w[GPI_P], dvxdx = h->pw[0].w[GPI_DVXDX], dvzdz = h->pw[0].w[GPI_DVZDZ], dK = h->bornc[0];
    OMP_FOR
    for (int ix = 1; ix <= p.n[2]; ix++) for (int iz = 1; iz <= p.n[0]; iz++)
        A2(p, iz, ix) = A2(p, iz, ix) + (A2(dvxdx, iz, ix) + A2(dvzdz, iz, ix)) * A2(dK, iz, ix);
}

int orc_run(orc_handle* h, int mode, int activepw, int src_flags) {
    double t0 = now_s();
    const int born = (mode & GPI_RUN_BORN) != 0, unshifted = (mode & GPI_RUN_UNSHIFTED_RHO) != 0;
    mode &= ~(GPI_RUN_BORN | GPI_RUN_UNSHIFTED_RHO);
    g_O = h->c.order - 1;
    if (unshifted && h->c.physics == GPI_ELASTIC) { snprintf(h->err, sizeof h->err, "the exact-transpose rho imaging is defined for acoustic media"); return 1; }
    if ((born || unshifted) && h->c.order != 2) { snprintf(h->err, sizeof h->err, "FD-Born and its exact-transpose imaging are defined for order 2 only"); return 1; }
    if (born && (h->nd != 2 || h->c.physics != GPI_ACOUSTIC || h->c.npw != 2 || (activepw & 3) != 3 || mode == GPI_MODE_ADJOINT || !h->born_ready)) {
        snprintf(h->err, sizeof h->err, "FD-Born needs a 2-D acoustic experiment, both wavefields active, a forward mode and orc_update_born"); return 1;
    }
    int nt = h->c.nt;
    const int recp[1] = {GPI_P}, recv[3] = {GPI_VX, GPI_VY, GPI_VZ};
    if (mode == GPI_MODE_FORWARD_SAVE && !h->c.store_boundary) {
        snprintf(h->err, sizeof h->err, "forward_save needs store_boundary=1 at construction (fdtd.jl:445-455)"); return 1;
    }
    const size_t nill = h->illum_on ? h->pw[0].w[GPI_P].len : 0;
    if (nill) memset(h->illum_stack, 0, nill * sizeof(double));           /* initialize!(pac): fill!(illum_stack, 0.0) (types.jl:171) */
    for (int issp = 0; issp < h->c.nshots; issp++) {
        reset_w2(h);
        shot_t* s1 = &h->pw[0].ss[issp];
        if (nill) memset(h->illum_shot, 0, nill * sizeof(double));
        if (mode == GPI_MODE_ADJOINT) {      /* boundary_force_snap_tau!/v! (boundary.jl:173-212) */
            int nf; const int* bf = boundary_fields(h, &nf);
            for (int i = 0; i < nf; i++) memcpy(h->pw[0].w[bf[i]].d, s1->snap[bf[i]], h->pw[0].w[bf[i]].len * sizeof(REAL));
            for (int i = 0; i < 3; i++) if (h->pw[0].w[recv[i]].d && s1->snap[recv[i]])
                memcpy(h->pw[0].w[recv[i]].d, s1->snap[recv[i]], h->pw[0].w[recv[i]].len * sizeof(REAL));
        }
        for (int it = 1; it <= nt; it++) {
            record(h, it, issp, activepw, recp, 1);
            if (mode == GPI_MODE_ADJOINT) {  /* save_tp! (save_tp.jl:5-12) */
                for (int ipw = 0; ipw < h->c.npw; ipw++) if (activepw & (1 << ipw))
                    for (int k = 0; k < GPI_NWAVEFIELD; k++) { int f = WAVEF[k]; if (h->pw[ipw].w[f].d) memcpy(h->pw[ipw].wtp[f].d, h->pw[ipw].w[f].d, h->pw[ipw].w[f].len * sizeof(REAL)); }
                boundary_force(h, nt - it + 1, issp);
            }
            for (int ipw = 0; ipw < h->c.npw; ipw++) if (activepw & (1 << ipw)) update_dstress(h, &h->pw[ipw]);
            for (int ipw = 0; ipw < h->c.npw; ipw++) if (activepw & (1 << ipw)) update_v(h, &h->pw[ipw]);
            add_velocity_source(h, it, issp, activepw, src_flags);
            if (born) born_velocity(h);
            record(h, it, issp, activepw, recv, 3);
            for (int ipw = 0; ipw < h->c.npw; ipw++) if (activepw & (1 << ipw)) update_dv(h, &h->pw[ipw]);
            for (int ipw = 0; ipw < h->c.npw; ipw++) if (activepw & (1 << ipw)) update_stress(h, &h->pw[ipw]);
            add_stress_source(h, it, issp, activepw, src_flags);
            if (born) born_stress(h);
            if (mode == GPI_MODE_FORWARD_SAVE) boundary_save(h, it, issp);
            if (mode == GPI_MODE_ADJOINT && h->c.npw == 2 && (activepw & 2)) compute_gradient(h, issp, unshifted);
            if (nill) {      /* compute_illum! (fdtd.jl:570-581; commented out at propagate.jl:236): illum[i] += abs2(p[i]), Float32 square, Float64 sum */
                const REAL* p = h->pw[0].w[GPI_P].d;
                OMP_FOR
                for (long long i = 0; i < (long long)nill; i++) { const REAL sq = p[i] * p[i]; h->illum_shot[i] += (double)sq; }
            }
            if (h->c.nsnaps > 0 && h->itsnaps)
                for (int k = 0; k < h->c.nsnaps; k++) if (h->itsnaps[k] == it)
                    for (int ipw = 0; ipw < h->c.npw; ipw++) if ((activepw & (1 << ipw)) && h->pw[ipw].ss[issp].usnaps)
                        memcpy(h->pw[ipw].ss[issp].usnaps[k], h->pw[ipw].w[h->c.snaps_field].d, h->pw[ipw].w[h->c.snaps_field].len * sizeof(REAL));
        }
        if (mode == GPI_MODE_FORWARD_SAVE) {     /* propagate.jl:251-258 */
            int nf; const int* bf = boundary_fields(h, &nf);
            for (int i = 0; i < nf; i++) { arr a = h->pw[0].w[bf[i]]; for (size_t k = 0; k < a.len; k++) s1->snap[bf[i]][k] = a.d[k] * (REAL)-1; }
        }
        if (nill) for (size_t i = 0; i < nill; i++) h->illum_stack[i] += h->illum_shot[i];      /* stack_illums! (fdtd.jl:556-565): shot order */
        update_dstress(h, &h->pw[0]);
        update_v(h, &h->pw[0]);
        if (mode == GPI_MODE_FORWARD_SAVE)
            for (int i = 0; i < 3; i++) if (h->pw[0].w[recv[i]].d) memcpy(s1->snap[recv[i]], h->pw[0].w[recv[i]].d, h->pw[0].w[recv[i]].len * sizeof(REAL));
    }
    /* sum_grads! (gradient.jl:2-11) */
    if (mode == GPI_MODE_ADJOINT && h->c.npw == 2)
        for (int issp = 0; issp < h->c.nshots; issp++) for (int p = 0; p < GPI_NPARAM; p++) {
            arr g1 = h->gradients[p], g = h->pw[0].ss[issp].grad[p];
            if (!g1.d || !g.d) continue;
            for (size_t i = 0; i < g1.len; i++) g1.d[i] = g1.d[i] + g.d[i];
        }
    h->last_run_s = now_s() - t0;
    h->last_steps = (double)nt * h->c.nshots;
    return 0;
}

/* steps-only entry for the CPU baseline: advance nsteps of shot 0 in forward mode, no reset */
int orc_advance(orc_handle* h, int it0, int nsteps) {
    g_O = h->c.order - 1;
    const int recp[1] = {GPI_P}, recv[3] = {GPI_VX, GPI_VY, GPI_VZ};
    for (int it = it0; it < it0 + nsteps && it <= h->c.nt; it++) {
        record(h, it, 0, 1, recp, 1);
        update_dstress(h, &h->pw[0]); update_v(h, &h->pw[0]);
        add_velocity_source(h, it, 0, 1, 1);
        record(h, it, 0, 1, recv, 3);
        update_dv(h, &h->pw[0]); update_stress(h, &h->pw[0]);
        add_stress_source(h, it, 0, 1, 1);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * lifecycle + setters/getters (same seams as include/gpifdtd.h, prefix orc_)
 * ---------------------------------------------------------------------------------------------- */
const char* orc_last_error(const orc_handle* h) { return h ? h->err : g_err; }
int orc_real_size(void) { return (int)sizeof(REAL); }
int orc_field_shape_order(int ndims, int order, int f, const int32_t n[3], int32_t out[3]) {
    const int keep = g_O; g_O = order - 1;
    int nn[3] = {n[0], n[1], n[2]}, o[3]; int r = field_shape(ndims, f, nn, o); if (!r) { out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; }
    g_O = keep;
    return r;
}
int orc_field_shape(int ndims, int f, const int32_t n[3], int32_t out[3]) { return orc_field_shape_order(ndims, 2, f, n, out); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int is_velocity(int f) { return f == GPI_VX || f == GPI_VY || f == GPI_VZ; }

int orc_create(const gpi_config* cfg, orc_handle** out) {
    if (!cfg || !out) { snprintf(g_err, sizeof g_err, "null argument"); return 1; }
    if (cfg->order != 2 && cfg->order != 4) { snprintf(g_err, sizeof g_err, "orders 2 and 4 are restated (6 and 8 are broken upstream)"); return 1; }
    if (cfg->npml != 40 + (cfg->order - 1)) { snprintf(g_err, sizeof g_err, "npml must be 40 + (order - 1) (GeoPhyInv.jl:90)"); return 1; }
    g_O = cfg->order - 1;
    if (cfg->ndims != 2 && cfg->ndims != 3) { snprintf(g_err, sizeof g_err, "ndims must be 2 or 3"); return 1; }
    orc_handle* h = (orc_handle*)calloc(1, sizeof *h);
    h->c = *cfg; h->nd = cfg->ndims;
    h->nz = cfg->n[0]; h->ny = cfg->ndims == 3 ? cfg->n[1] : 1; h->nx = cfg->n[2];
    h->dt = (REAL)cfg->dt; h->dtI = (REAL)cfg->dtI;
    h->dzI = (REAL)cfg->dI[0]; h->dyI = (REAL)cfg->dI[1]; h->dxI = (REAL)cfg->dI[2];
    int n[3] = {h->nz, h->ny, h->nx}, sh[3];
    int ac = cfg->physics == GPI_ACOUSTIC;
    /* mod / dmod (fdtd.jl:139-154, medium.jl:81-95) */
    arr_alloc(&h->mod[GPI_RHO], n);
    if (ac) arr_alloc(&h->mod[GPI_INVK], n); else { arr_alloc(&h->mod[GPI_INVLAMBDA], n); arr_alloc(&h->mod[GPI_INVMU], n); }
    field_shape(h->nd, ac ? GPI_DPDX : GPI_DTAUXXDX, n, sh); arr_alloc(&h->dmod[DM_BX], sh);
    field_shape(h->nd, ac ? GPI_DPDZ : GPI_DTAUZZDZ, n, sh); arr_alloc(&h->dmod[DM_BZ], sh);
    if (h->nd == 3) { field_shape(3, ac ? GPI_DPDY : GPI_DTAUYYDY, n, sh); arr_alloc(&h->dmod[DM_BY], sh); }
    if (ac) arr_alloc(&h->dmod[DM_DTK], n);
    else {
        arr_alloc(&h->dmod[DM_DTM], n); arr_alloc(&h->dmod[DM_DTLAMBDA], n);
        field_shape(h->nd, GPI_DVXDZ, n, sh); arr_alloc(&h->dmod[DM_MUXZ], sh);
        if (h->nd == 3) { field_shape(3, GPI_DVXDY, n, sh); arr_alloc(&h->dmod[DM_MUXY], sh); field_shape(3, GPI_DVYDZ, n, sh); arr_alloc(&h->dmod[DM_MUYZ], sh); }
    }
    if (ac) { arr_alloc(&h->gradients[GPI_INVK], n); arr_alloc(&h->gradients[GPI_RHO], n); }
    else { arr_alloc(&h->gradients[GPI_INVLAMBDA], n); arr_alloc(&h->gradients[GPI_INVMU], n); arr_alloc(&h->gradients[GPI_RHO], n); }
    for (int f = GPI_NWAVEFIELD; f < GPI_NFIELD; f++) if (field_exists(h->nd, cfg->physics, f)) {
        h->pa[f] = (REAL*)calloc(2 * cfg->npml, sizeof(REAL)); h->pb[f] = (REAL*)calloc(2 * cfg->npml, sizeof(REAL));
        h->pk[f] = (REAL*)malloc(2 * cfg->npml * sizeof(REAL)); for (int i = 0; i < 2 * cfg->npml; i++) h->pk[f][i] = 1;   /* cpml.jl:114 */
    }
    for (int ipw = 0; ipw < cfg->npw; ipw++) {
        pw_t* pw = &h->pw[ipw];
        for (int f = 0; f < GPI_NFIELD; f++) if (field_exists(h->nd, cfg->physics, f)) {
            field_shape(h->nd, f, n, sh); arr_alloc(&pw->w[f], sh);
            if (f < GPI_NWAVEFIELD) { arr_alloc(&pw->wtp[f], sh); if (is_velocity(f)) arr_alloc(&pw->vbuf[f], sh); }
            else { sh[dfield_axis(f)] = 2 * cfg->npml; arr_alloc(&pw->mem[f], sh); }
        }
        arr_alloc(&pw->taubuf, n);
        pw->ss = (shot_t*)calloc(cfg->nshots, sizeof(shot_t));
        for (int is = 0; is < cfg->nshots; is++) {
            shot_t* s = &pw->ss[is];
            if (ac) { arr_alloc(&s->grad[GPI_INVK], n); arr_alloc(&s->grad[GPI_RHO], n); }
            else if (cfg->npw == 2) { arr_alloc(&s->grad[GPI_INVLAMBDA], n); arr_alloc(&s->grad[GPI_INVMU], n); arr_alloc(&s->grad[GPI_RHO], n); }
            if (ipw == 0) for (int k = 0; k < GPI_NWAVEFIELD; k++) {
                int f = WAVEF[k]; if (!pw->w[f].d) continue;
                s->snap[f] = (REAL*)calloc(pw->w[f].len, sizeof(REAL));
                int isb = ac ? (f == GPI_P) : (h->nd == 3 ? (f >= GPI_TAUXX && f <= GPI_TAUYZ) : (f == GPI_TAUXX || f == GPI_TAUXZ || f == GPI_TAUZZ));
                if (!isb) continue;
                for (int q = 0; q < 3; q++) {
                    if (q == 1 && h->nd == 2) continue;
                    int bn[3] = {pw->w[f].n[0], pw->w[f].n[1], pw->w[f].n[2]}; bn[q] = 2 * cfg->nbound;
                    s->bnd_slot[f][q] = (size_t)bn[0] * bn[1] * bn[2];
                    s->bnd[f][q] = (REAL*)calloc(s->bnd_slot[f][q] * (cfg->store_boundary ? cfg->nt : 1), sizeof(REAL));
                }
            }
            if (cfg->nsnaps > 0 && field_exists(h->nd, cfg->physics, cfg->snaps_field)) {
                s->usnaps = (REAL**)calloc(cfg->nsnaps, sizeof(REAL*));
                for (int k = 0; k < cfg->nsnaps; k++) s->usnaps[k] = (REAL*)calloc(pw->w[cfg->snaps_field].len, sizeof(REAL));
            }
        }
    }
    *out = h;
    return 0;
}

static void csc_free(csc* m) { free(m->colptr); free(m->rowval); free(m->nzval); memset(m, 0, sizeof *m); }

int orc_destroy(orc_handle* h) {
    if (!h) return 0;
    for (int p = 0; p < GPI_NPARAM; p++) { arr_free(&h->mod[p]); arr_free(&h->gradients[p]); }
    for (int d = 0; d < DM_N; d++) arr_free(&h->dmod[d]);
    for (int p = 0; p < GPI_NPARAM; p++) arr_free(&h->modp[p]);
    for (int q = 0; q < 3; q++) arr_free(&h->bornc[q]);
    for (int f = 0; f < GPI_NFIELD; f++) { free(h->pa[f]); free(h->pb[f]); free(h->pk[f]); }
    for (int ipw = 0; ipw < h->c.npw; ipw++) {
        pw_t* pw = &h->pw[ipw];
        for (int f = 0; f < GPI_NFIELD; f++) { arr_free(&pw->w[f]); arr_free(&pw->mem[f]); }
        for (int f = 0; f < GPI_NWAVEFIELD; f++) { arr_free(&pw->wtp[f]); arr_free(&pw->vbuf[f]); }
        arr_free(&pw->taubuf);
        for (int is = 0; is < h->c.nshots; is++) {
            shot_t* s = &pw->ss[is];
            for (int f = 0; f < GPI_NWAVEFIELD; f++) {
                csc_free(&s->spray[f]); csc_free(&s->interp[f]); free(s->wavelets[f]); free(s->records[f]); free(s->snap[f]);
                for (int q = 0; q < 3; q++) free(s->bnd[f][q]);
            }
            for (int p = 0; p < GPI_NPARAM; p++) arr_free(&s->grad[p]);
            if (s->usnaps) { for (int k = 0; k < h->c.nsnaps; k++) free(s->usnaps[k]); free(s->usnaps); }
        }
        free(pw->ss);
    }
    free(h->itsnaps);
    free(h->illum_shot); free(h->illum_stack);
    free(h);
    return 0;
}

#define CHECK(cond, msg) do { if (!(cond)) { snprintf(h->err, sizeof h->err, "%s", msg); return 1; } } while (0)

int orc_set_medium(orc_handle* h, int p, const REAL* a) {
    CHECK(p >= 0 && p < GPI_NPARAM && h->mod[p].d, "medium parameter not part of this physics");
    memcpy(h->mod[p].d, a, h->mod[p].len * sizeof(REAL)); return 0;
}
int orc_get_medium(orc_handle* h, int p, REAL* out) {
    CHECK(p >= 0 && p < GPI_NPARAM && h->mod[p].d, "medium parameter not part of this physics");
    memcpy(out, h->mod[p].d, h->mod[p].len * sizeof(REAL)); return 0;
}
int orc_update_dmod(orc_handle* h) { g_O = h->c.order - 1; update_dmod(h); return 0; }
int orc_set_medium_pert(orc_handle* h, int p, const REAL* a) {
    CHECK(h->nd == 2 && h->c.physics == GPI_ACOUSTIC && (p == GPI_INVK || p == GPI_RHO), "FD-Born: 2-D acoustic invK | rho only");
    if (!h->modp[p].d) { int n[3] = {h->nz, 1, h->nx}; CHECK(!arr_alloc(&h->modp[p], n), "out of memory"); }
    memcpy(h->modp[p].d, a, h->modp[p].len * sizeof(REAL)); h->born_ready = 0; return 0;
}
int orc_update_born(orc_handle* h) {
    CHECK(h->c.order == 2, "FD-Born is defined for order 2 only");
    CHECK(h->nd == 2 && h->c.physics == GPI_ACOUSTIC && h->c.npw == 2 && h->modp[GPI_INVK].d && h->modp[GPI_RHO].d, "FD-Born: set both perturbations first");
    const int dm[3] = {DM_DTK, DM_BX, DM_BZ};
    for (int q = 0; q < 3; q++) if (!h->bornc[q].d) { CHECK(!arr_alloc(&h->bornc[q], h->dmod[dm[q]].n), "out of memory"); }
    update_born(h); h->born_ready = 1; return 0;
}
int orc_set_pml(orc_handle* h, int f, const REAL* a, const REAL* b, const REAL* kI) {
    CHECK(f >= GPI_NWAVEFIELD && f < GPI_NFIELD && h->pa[f], "not a derivative field of this physics");
    size_t nb = 2 * h->c.npml * sizeof(REAL);
    memcpy(h->pa[f], a, nb); memcpy(h->pb[f], b, nb); memcpy(h->pk[f], kI, nb); return 0;
}
int orc_set_sparse(orc_handle* h, int kind, int ipw, int issp, int f, int ncol, const int64_t* colptr, const int64_t* rowval, const REAL* nzval) {
    CHECK(ipw >= 0 && ipw < h->c.npw && issp >= 0 && issp < h->c.nshots, "bad pw/shot index");
    CHECK(f >= 0 && f < GPI_NWAVEFIELD && h->pw[ipw].w[f].d, "field not part of this physics");
    shot_t* s = &h->pw[ipw].ss[issp];
    csc* m = kind == GPI_SPRAY ? &s->spray[f] : &s->interp[f];
    csc_free(m);
    int64_t nnz = colptr[ncol] - 1;
    for (int64_t k = 0; k < nnz; k++) CHECK(rowval[k] >= 1 && (size_t)rowval[k] <= h->pw[ipw].w[f].len, "row index outside the field array");
    m->ncol = ncol;
    m->colptr = (int64_t*)malloc((ncol + 1) * sizeof(int64_t)); memcpy(m->colptr, colptr, (ncol + 1) * sizeof(int64_t));
    m->rowval = (int64_t*)malloc((nnz ? nnz : 1) * sizeof(int64_t)); memcpy(m->rowval, rowval, nnz * sizeof(int64_t));
    m->nzval = (REAL*)malloc((nnz ? nnz : 1) * sizeof(REAL)); memcpy(m->nzval, nzval, nnz * sizeof(REAL));
    if (kind == GPI_INTERP) { free(s->records[f]); s->records[f] = (REAL*)calloc((size_t)h->c.nt * (ncol ? ncol : 1), sizeof(REAL)); }
    return 0;
}
int orc_set_wavelets(orc_handle* h, int ipw, int issp, int f, int ns, const REAL* w) {
    CHECK(ipw >= 0 && ipw < h->c.npw && issp >= 0 && issp < h->c.nshots, "bad pw/shot index");
    CHECK(f >= 0 && f < GPI_NWAVEFIELD && h->pw[ipw].w[f].d, "field not part of this physics");
    shot_t* s = &h->pw[ipw].ss[issp];
    free(s->wavelets[f]); s->wavelets[f] = NULL; s->ns[f] = ns;
    if (!w) return 0;
    s->wavelets[f] = (REAL*)malloc((size_t)h->c.nt * ns * sizeof(REAL));
    memcpy(s->wavelets[f], w, (size_t)h->c.nt * ns * sizeof(REAL));
    return 0;
}
int orc_get_records(orc_handle* h, int ipw, int issp, int f, REAL* out) {
    CHECK(ipw >= 0 && ipw < h->c.npw && issp >= 0 && issp < h->c.nshots, "bad pw/shot index");
    shot_t* s = &h->pw[ipw].ss[issp];
    CHECK(f >= 0 && f < GPI_NWAVEFIELD && s->records[f], "no receivers set for this field");
    memcpy(out, s->records[f], (size_t)h->c.nt * s->interp[f].ncol * sizeof(REAL)); return 0;
}
int orc_get_gradient(orc_handle* h, int p, REAL* out) {
    CHECK(p >= 0 && p < GPI_NPARAM && h->gradients[p].d, "no gradient for this parameter");
    memcpy(out, h->gradients[p].d, h->gradients[p].len * sizeof(REAL)); return 0;
}
int orc_get_field(orc_handle* h, int ipw, int ibatch, int f, REAL* out) {
    (void)ibatch;
    CHECK(ipw >= 0 && ipw < h->c.npw && f >= 0 && f < GPI_NFIELD && h->pw[ipw].w[f].d, "no such field");
    memcpy(out, h->pw[ipw].w[f].d, h->pw[ipw].w[f].len * sizeof(REAL)); return 0;
}
int orc_set_field(orc_handle* h, int ipw, int ibatch, int f, const REAL* in) {
    (void)ibatch;
    CHECK(ipw >= 0 && ipw < h->c.npw && f >= 0 && f < GPI_NFIELD && h->pw[ipw].w[f].d, "no such field");
    memcpy(h->pw[ipw].w[f].d, in, h->pw[ipw].w[f].len * sizeof(REAL)); return 0;
}
int orc_get_dmod(orc_handle* h, int d, REAL* out, int32_t shape[3]) {
    CHECK(d >= 0 && d < DM_N && h->dmod[d].d, "no such dmod");
    if (out) memcpy(out, h->dmod[d].d, h->dmod[d].len * sizeof(REAL));
    shape[0] = h->dmod[d].n[0]; shape[1] = h->dmod[d].n[1]; shape[2] = h->dmod[d].n[2]; return 0;
}
int orc_get_memory(orc_handle* h, int ipw, int f, REAL* out) {
    CHECK(ipw >= 0 && ipw < h->c.npw && f >= GPI_NWAVEFIELD && f < GPI_NFIELD && h->pw[ipw].mem[f].d, "no such memory field");
    memcpy(out, h->pw[ipw].mem[f].d, h->pw[ipw].mem[f].len * sizeof(REAL)); return 0;
}
int orc_set_snap_steps(orc_handle* h, int nsnaps, const int32_t* its) {
    CHECK(nsnaps == h->c.nsnaps, "nsnaps differs from the configuration");
    free(h->itsnaps); h->itsnaps = (int32_t*)malloc((nsnaps ? nsnaps : 1) * sizeof(int32_t));
    memcpy(h->itsnaps, its, nsnaps * sizeof(int32_t)); return 0;
}
int orc_get_snap(orc_handle* h, int ipw, int issp, int isnap, REAL* out) {
    CHECK(ipw >= 0 && ipw < h->c.npw && issp >= 0 && issp < h->c.nshots && isnap >= 0 && isnap < h->c.nsnaps, "bad snapshot index");
    shot_t* s = &h->pw[ipw].ss[issp]; CHECK(s->usnaps, "snapshots not configured");
    memcpy(out, s->usnaps[isnap], h->pw[ipw].w[h->c.snaps_field].len * sizeof(REAL)); return 0;
}
int orc_get_boundary(orc_handle* h, int issp, int f, int axis, int it, REAL* out, int64_t* len) {
    shot_t* s = &h->pw[0].ss[issp];
    CHECK(f >= 0 && f < GPI_NWAVEFIELD && axis >= 0 && axis < 3 && s->bnd[f][axis], "no boundary store");
    *len = (int64_t)s->bnd_slot[f][axis];
    if (out) memcpy(out, s->bnd[f][axis] + (size_t)(it - 1) * s->bnd_slot[f][axis], s->bnd_slot[f][axis] * sizeof(REAL));
    return 0;
}
int orc_reset(orc_handle* h, int what) {
    if (what & GPI_RESET_WAVEFIELDS) reset_w2(h);
    for (int ipw = 0; ipw < h->c.npw; ipw++) for (int is = 0; is < h->c.nshots; is++) {
        shot_t* s = &h->pw[ipw].ss[is];
        if (what & GPI_RESET_RECORDS) for (int f = 0; f < GPI_NWAVEFIELD; f++) if (s->records[f]) memset(s->records[f], 0, (size_t)h->c.nt * s->interp[f].ncol * sizeof(REAL));
        if (what & GPI_RESET_GRADIENTS) for (int p = 0; p < GPI_NPARAM; p++) arr_zero(&s->grad[p]);
        if (what & GPI_RESET_BOUNDARY) for (int f = 0; f < GPI_NWAVEFIELD; f++) {
            if (s->snap[f]) memset(s->snap[f], 0, h->pw[ipw].w[f].len * sizeof(REAL));
            for (int q = 0; q < 3; q++) if (s->bnd[f][q]) memset(s->bnd[f][q], 0, s->bnd_slot[f][q] * (h->c.store_boundary ? h->c.nt : 1) * sizeof(REAL));
        }
        if ((what & GPI_RESET_SNAPS) && s->usnaps) for (int k = 0; k < h->c.nsnaps; k++) memset(s->usnaps[k], 0, h->pw[ipw].w[h->c.snaps_field].len * sizeof(REAL));
    }
    if (what & GPI_RESET_GRADIENTS) for (int p = 0; p < GPI_NPARAM; p++) arr_zero(&h->gradients[p]);
    return 0;
}
int orc_set_illum(orc_handle* h, int on) {
    if (on && h->c.physics != GPI_ACOUSTIC) { snprintf(h->err, sizeof h->err, "the illumination is defined on the pressure field: acoustic experiments only"); return 1; }
    if (on && !h->illum_shot) {
        h->illum_shot = (double*)calloc(h->pw[0].w[GPI_P].len, sizeof(double));
        h->illum_stack = (double*)calloc(h->pw[0].w[GPI_P].len, sizeof(double));
    }
    h->illum_on = on != 0;
    return 0;
}
int orc_get_illum(orc_handle* h, double* out) {
    if (!h->illum_stack) { snprintf(h->err, sizeof h->err, "no illumination: call orc_set_illum(h, 1) before orc_run"); return 1; }
    memcpy(out, h->illum_stack, h->pw[0].w[GPI_P].len * sizeof(double));
    return 0;
}
double orc_last_run_seconds(orc_handle* h) { return h->last_run_s; }
