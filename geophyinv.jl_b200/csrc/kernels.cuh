// kernels.cuh -- sm_100a stencil kernels of the B200 FDTD engine.
//
// What they compute follows GeoPhyInv.jl's src/fdtd (citations relative to /root/reference);
// how they compute it does not: the reference issues ~20 (2-D acoustic) to ~55 (3-D elastic)
// ParallelStencil kernels per time step that round-trip every derivative array through memory
// (advance_acou.jl:11-96, advance_elastic.jl:233-531, cpml.jl:158-215, dirichlet.jl:35-74).
// Here one step is two fused sweeps:
//
//   k_vel    : stress derivatives -> CPML memory update -> velocity update -> rigid faces
//              (update_dstress! + update_v!,  propagate.jl:191-198)
//   k_stress : velocity derivatives -> CPML -> stress / pressure update -> free surface
//              (update_dv! + update_stress!, propagate.jl:212-219)
//
// Derivatives never touch HBM.  Every field lives in ONE common "unified box"
// (nz+1) x (ny+1) x (nx+1), z fastest with a 128-byte-multiple pitch, so that all arrays of a
// cell share the same linear index and alignment:
//     integer nodes  (tauii/p)          : reference index iz      -> k = iz-1
//     velocity nodes (-1/2, length n+1) : reference index iz      -> k = iz-1   (node k sits at k-1/2)
//     half nodes     (+1/2, length n-1) : reference index iz      -> k = iz     (node k sits at k-1/2)
//     inner nodes    (+1,   length n-2) : reference index iz      -> k = iz
// (src/fields.jl:92-671).  A forward difference onto a half node is f(k)-f(k-1); onto an integer
// node from half nodes it is f(k+1)-f(k).
//
// Arithmetic: the reference's Base.Threads path evaluates Float32 mul/add separately (Julia does
// not contract to FMA outside @fastmath/muladd).  All updates therefore use __fmul_rn/__fadd_rn/
// __fsub_rn, which nvcc never contracts, in exactly the reference's association order, so the
// engine reproduces the CPU path bit for bit rather than "to tolerance".  The kernels are HBM-bound
// (0.3-0.5 flop/B); giving up FMA costs nothing measurable.
#pragma once
#ifndef GPI_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace gpi {

// ---- float literals of the reference's @parallel kernels ------------------------------------------------------------
// The @av_* macros carry `* 0.5` / `* 0.25` and the order-4 differences `* 27.0` (diff2D.jl, diff3D.jl).  ParallelStencil's
// @parallel retypes the float literals of a kernel to the number type the package was initialised with
// (@init_parallel_stencil(Threads, Float32, N), src/GeoPhyInv.jl:95-100), so in the reference's Float32 build these
// expressions are pure Float32: wide_t = float (default).  -DGPI_LITERALS_F64 evaluates them in Float64 with one rounding at
// the store (plain Julia promotion of a Float64 literal; what round 1 shipped).  tests/test_reference_pinned.py holds the
// engine to fixtures evaluated from the reference's own kernel text in the default typing, bit for bit.  The `2.0` of
// update_dmod!'s dtM broadcast (medium.jl:164-166) is outside any @parallel kernel and stays Float64 in both.
#ifdef GPI_LITERALS_F64
typedef double wide_t;
__device__ __forceinline__ wide_t wmul(wide_t a, wide_t b) { return __dmul_rn(a, b); }
__device__ __forceinline__ wide_t wadd(wide_t a, wide_t b) { return __dadd_rn(a, b); }
__device__ __forceinline__ wide_t wsub(wide_t a, wide_t b) { return __dsub_rn(a, b); }
__device__ __forceinline__ wide_t wdiv(wide_t a, wide_t b) { return __ddiv_rn(a, b); }
#else
typedef float wide_t;
__device__ __forceinline__ wide_t wmul(wide_t a, wide_t b) { return __fmul_rn(a, b); }
__device__ __forceinline__ wide_t wadd(wide_t a, wide_t b) { return __fadd_rn(a, b); }
__device__ __forceinline__ wide_t wsub(wide_t a, wide_t b) { return __fsub_rn(a, b); }
__device__ __forceinline__ wide_t wdiv(wide_t a, wide_t b) { return __fdiv_rn(a, b); }
#endif
// ---- programmatic dependent launch (2-D time loops, gpi_handle::pdl) ---------------------------------------------------
// A 2-D one-shot kernel lasts ~3 us of memory time but costs ~6 us as a stream / graph node: the rest is the launch of the next grid behind
// the drained one.  The kernels of the 2-D chain therefore (1) touch nothing an earlier kernel writes before pdl_wait returns (the whole
// preceding grid has completed and its stores are visible) -- k_post reads its static tables ahead of it -- and (2) then release their
// dependents (pdl_release): once every CTA of the grid has got that far, the next grid's CTAs are scheduled into the slots that free up
// and run their index preamble.  Release after wait: at most two grids overlap, and every kernel of the chain waits, so completion is
// transitive along the stream.  Both are no-ops for a launch without the attribute (and on the host).
__device__ __forceinline__ void pdl_release() {
#ifndef GPI_HOST_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifndef GPI_HOST_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// dt / (s * lit): store_invav*! (medium.jl:191-221) with s the Float32 sum of the macro
__device__ __forceinline__ float inv_av(float dt, float s, float lit) { return (float)wdiv((wide_t)dt, wmul((wide_t)s, (wide_t)lit)); }

enum { ZMIN = 1, ZMAX = 2, YMIN = 4, YMAX = 8, XMIN = 16, XMAX = 32 };

// tau slots inside a wavefield set
enum { T_XX = 0, T_YY = 1, T_ZZ = 2, T_XY = 3, T_XZ = 4, T_YZ = 5 };   // acoustic: T_XX slot holds p
enum { V_X = 0, V_Y = 1, V_Z = 2 };
// dmod slots
enum { C_BX = 0, C_BY, C_BZ, C_K /*dtK or dtM*/, C_L /*dtlambda*/, C_MUXZ, C_MUXY, C_MUYZ, C_N };

struct Geom {
    int nz, ny, nx;          // extended GLOBAL grid (tauii nodes); ny = 1 in 2-D
    // z-slab window of this handle (multi-GPU domain decomposition; the full domain has koff = 0,
    // klo = 0, khi = nz): local z index kl <-> global k = kl + koff (koff % 4 == 0); this handle
    // updates the nodes kl in [klo, khi]; kl = klo-1 and khi+1 are halo planes filled by the exchange.
    int koff, klo, khi;
    int h;                   // order 4: storage index = unified coordinate + h on every axis (kernels4.cuh); 0 at order 2
    int ioff;                // first x plane of this launch (k_*3v: plane = blockIdx.y + ioff; pipelined z-slab runs launch the grid in two x halves)
    int pz;                  // z pitch in floats (multiple of 32)
    int ny1;                 // rows per x-plane: ny+1 (3-D) or 1 (2-D)
    int nx1;                 // nx+1
    int npml;
    int pzm;                 // pitch of z-axis CPML memory rows (>= 2*npml, multiple of 32)
    int pml, rigid, freesurf;// face masks
    float dzI, dyI, dxI;
    long long vol;           // floats per field volume = pz*ny1*nx1
};

// one CPML memory term (a derivative field of the reference): memory array + a,b,kI vectors
struct PmlTerm {
    float* mem;              // batch slot 0
    const float* a; const float* b; const float* kI;   // 2*npml each
    long long bstride;       // floats between batch slots
};

struct StepArgs {
    // wavefields (batch slot 0 of this pw); *_out == *_in in forward mode
    float* tau[6];
    float* v[3];
    const float* c[C_N];     // dmod in unified layout (shared by all shots)
    PmlTerm pv[9];           // terms of k_vel, order documented at the kernel
    PmlTerm ps[9];           // terms of k_stress
    long long wstride;       // floats between batch slots of the wavefield set
    int nbatch;
    // FD-Born (2-D acoustic): when set, the scalar kernels also store the CPML-corrected derivatives of this
    // wavefield (dpdx, dpdz in the velocity kernel; dvxdx, dvzdz in the stress kernel), the scattering sources
    float* dout[2];
    long long dstride;       // floats between batch slots of dout
    // out-of-place stepping (the OOP instantiations only; adjoint runs with GPI_PINGPONG=1): the velocity kernel reads v and
    // writes v_o, the stress kernel reads tau and writes tau_o -- the set just read stays behind as the previous time level,
    // which is what save_tp! (save_tp.jl:5-12) otherwise copies
    float* tau_o[6];
    float* v_o[3];
};

__device__ __forceinline__ long long uidx(const Geom& g, int k, int j, int i) {
    return (long long)k + (long long)g.pz * ((long long)j + (long long)g.ny1 * (long long)i);
}

// CPML memory-variable update on one derivative value (cpml.jl:175-183):
//     m = b*m + a*d ;  d = d*kI + m
// AXIS: 0=z 1=y 2=x.  The derivative field covers unified coordinates [s0, s0+len) along AXIS;
// its first/last npml entries are the min/max slabs (cpml.jl:190-211).
// Storage (engine's own): x terms [k, j, s], y terms [k, s, i] with s the reference's slab index
// 0..2*npml-1.  z terms [zi, j, i] with a row of pzm = 2*zhalf floats: the min slab keeps the unified
// coordinate (zi = k), the max slab sits at zi = zhalf + (k - kb), kb = the 4-aligned start of the slab,
// so that the vector kernels can move z memory with aligned float4 accesses.  For z terms the
// coefficient vectors are pre-expanded to k-indexed tables (identity a=b=0, kI=1 outside the slabs).
__device__ __forceinline__ int slab_index(int u, int s0, int len, int npml, bool hmin, bool hmax) {
    const int r = u - s0;
    if (hmin && r >= 0 && r < npml) return r;
    const int rm = r - (len - npml);
    if (hmax && rm >= 0 && rm < npml) return npml + rm;
    return -1;
}
__device__ __forceinline__ int zslab_base(int s0, int len, int npml) { return ((s0 + len - npml) >> 2) << 2; }

template <int AXIS>    // k is the GLOBAL z index
__device__ __forceinline__ float cpml(const Geom& g, const PmlTerm& t, float d, int k, int j, int i,
                                      int s0, int len, int b) {
    const int u = AXIS == 0 ? k : (AXIS == 1 ? j : i);
    const int minbit = AXIS == 0 ? ZMIN : (AXIS == 1 ? YMIN : XMIN);
    const int maxbit = AXIS == 0 ? ZMAX : (AXIS == 1 ? YMAX : XMAX);
    const int s = slab_index(u, s0, len, g.npml, (g.pml & minbit) != 0, (g.pml & maxbit) != 0);
    if (s >= 0) {
        long long mi;
        int ci = s;
        if (AXIS == 2)      mi = (long long)(k - g.koff) + (long long)g.pz * ((long long)j + (long long)g.ny1 * s);
        else if (AXIS == 1) mi = (long long)(k - g.koff) + (long long)g.pz * ((long long)s + 2LL * g.npml * i);
        else {
            const int zi = s < g.npml ? k : (g.pzm >> 1) + (k - zslab_base(s0, len, g.npml));
            mi = (long long)zi + (long long)g.pzm * ((long long)j + (long long)g.ny1 * i);
            ci = k;
        }
        float* mp = t.mem + (long long)b * t.bstride + mi;
        float m = *mp;
        m = __fadd_rn(__fmul_rn(__ldg(t.b + ci), m), __fmul_rn(__ldg(t.a + ci), d));
        *mp = m;
        d = __fadd_rn(__fmul_rn(d, __ldg(t.kI + ci)), m);
    }
    return d;
}

// thread -> unified cell.  3-D: grid.z folds (x tile, batch); 2-D: grid.y = x tiles, grid.z = batch.
template <int ND>
__device__ __forceinline__ bool cell(const Geom& g, int nbatch, int& k, int& j, int& i, int& b) {
    k = blockIdx.x * blockDim.x + threadIdx.x;
    if (ND == 3) {
        j = blockIdx.y * blockDim.y + threadIdx.y;
        const int ntx = (g.nx1 + blockDim.z - 1) / blockDim.z;
        b = blockIdx.z / ntx;
        i = (blockIdx.z - b * ntx) * blockDim.z + threadIdx.z;
    } else {
        j = 0;
        i = blockIdx.y * blockDim.y + threadIdx.y;
        b = blockIdx.z;
    }
    (void)nbatch;
    return k >= g.klo && k <= g.khi && j < g.ny1 && i <= g.nx;      // k: local z index of an owned node
}

// ------------------------------------------------------------------------------------------------
// k_vel: update_dstress! + update_v! (+ dirichlet) fused.
//   acoustic (advance_acou.jl:258-283):  v_inn = v_inn + b * dpd?          (PLUS)
//   elastic  (advance_elastic.jl:69-110): v_inn = v_inn - b * (sum of dtau) (MINUS, reference order)
// CPML term order in a.pv:
//   acoustic : 0 dpdx, 1 dpdy, 2 dpdz
//   elastic  : 0 dtauxxdx 1 dtauxydy 2 dtauxzdz | 3 dtauxydx 4 dtauyydy 5 dtauyzdz | 6 dtauxzdx 7 dtauyzdy 8 dtauzzdz
// ------------------------------------------------------------------------------------------------
template <int ND, int EL, int OOP = 0>
__device__ __forceinline__ void vel_cell(const Geom& g, const StepArgs& a, int kl, int j, int i, int b) {
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, kl, j, i);
    const int k = kl + g.koff;                  // global z index: every range predicate below is global
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;

    // interior ranges of the three velocity components (@inn of compute_v!)
    const bool iny  = (ND == 2) || (j >= 1 && j <= ny - 2);
    const bool inyh = (ND == 2) || (j >= 1 && j <= ny - 1);
    const bool vxin = k >= 1 && k <= nz - 2 && iny && i >= 1 && i <= nx - 1;
    const bool vzin = k >= 1 && k <= nz - 1 && iny && i >= 1 && i <= nx - 2;
    const bool vyin = (ND == 3) && k >= 1 && k <= nz - 2 && inyh && i >= 1 && i <= nx - 2;

    // array extents (who owns an entry to write)
    const bool vxown = k <= nz - 1 && (ND == 2 || j <= ny - 1);            // i <= nx always
    const bool vzown = i <= nx - 1 && (ND == 2 || j <= ny - 1);            // k <= nz always
    const bool vyown = (ND == 3) && k <= nz - 1 && i <= nx - 1;            // j <= ny always

    float* vx = (OOP ? a.v_o[V_X] : a.v[V_X]) + w; float* vz = (OOP ? a.v_o[V_Z] : a.v[V_Z]) + w;        // written
    float* vy = (ND == 3) ? (OOP ? a.v_o[V_Y] : a.v[V_Y]) + w : nullptr;
    float nvx = 0.f, nvy = 0.f, nvz = 0.f;
    const float* vxi = OOP ? a.v[V_X] + w : vx; const float* vzi = OOP ? a.v[V_Z] + w : vz;                  // read
    const float* vyi = (ND == 3) ? (OOP ? a.v[V_Y] + w : vy) : nullptr;
    if (vxown) nvx = vxi[c];
    if (vzown) nvz = vzi[c];
    if (ND == 3 && vyown) nvy = vyi[c];

    if (!EL) {
        const float* p = a.tau[T_XX] + w;
        const float pc = (vxin || vzin || vyin) ? p[c] : 0.f;
        if (vxin) {
            float d = __fmul_rn(__fsub_rn(pc, p[c - sx]), g.dxI);
            d = cpml<2>(g, a.pv[0], d, k, j, i, 1, nx - 1, b);
            if (ND == 2 && a.dout[0]) a.dout[0][c + (long long)b * a.dstride] = d;
            nvx = __fadd_rn(nvx, __fmul_rn(__ldg(a.c[C_BX] + c), d));
        }
        if (ND == 3 && vyin) {
            float d = __fmul_rn(__fsub_rn(pc, p[c - sy]), g.dyI);
            d = cpml<1>(g, a.pv[1], d, k, j, i, 1, ny - 1, b);
            nvy = __fadd_rn(nvy, __fmul_rn(__ldg(a.c[C_BY] + c), d));
        }
        if (vzin) {
            float d = __fmul_rn(__fsub_rn(pc, p[c - 1]), g.dzI);
            d = cpml<0>(g, a.pv[2], d, k, j, i, 1, nz - 1, b);
            if (ND == 2 && a.dout[1]) a.dout[1][c + (long long)b * a.dstride] = d;
            nvz = __fadd_rn(nvz, __fmul_rn(__ldg(a.c[C_BZ] + c), d));
        }
    } else if (ND == 2) {
        const float* txx = a.tau[T_XX] + w; const float* tzz = a.tau[T_ZZ] + w; const float* txz = a.tau[T_XZ] + w;
        if (vxin) {
            float dxx = __fmul_rn(__fsub_rn(txx[c], txx[c - sx]), g.dxI);      // @d_xi(tauxx)
            dxx = cpml<2>(g, a.pv[0], dxx, k, j, i, 1, nx - 1, b);
            float dxz = __fmul_rn(__fsub_rn(txz[c + 1], txz[c]), g.dzI);       // @d_za(tauxz)
            dxz = cpml<0>(g, a.pv[2], dxz, k, j, i, 1, nz - 2, b);
            nvx = __fsub_rn(nvx, __fmul_rn(__ldg(a.c[C_BX] + c), __fadd_rn(dxx, dxz)));
        }
        if (vzin) {
            float dzx = __fmul_rn(__fsub_rn(txz[c + sx], txz[c]), g.dxI);      // @d_xa(tauxz)
            dzx = cpml<2>(g, a.pv[6], dzx, k, j, i, 1, nx - 2, b);
            float dzz = __fmul_rn(__fsub_rn(tzz[c], tzz[c - 1]), g.dzI);       // @d_zi(tauzz)
            dzz = cpml<0>(g, a.pv[8], dzz, k, j, i, 1, nz - 1, b);
            nvz = __fsub_rn(nvz, __fmul_rn(__ldg(a.c[C_BZ] + c), __fadd_rn(dzx, dzz)));
        }
    } else {
        const float* txx = a.tau[T_XX] + w; const float* tyy = a.tau[T_YY] + w; const float* tzz = a.tau[T_ZZ] + w;
        const float* txy = a.tau[T_XY] + w; const float* txz = a.tau[T_XZ] + w; const float* tyz = a.tau[T_YZ] + w;
        // shared centre values (each read once per thread)
        const bool any = vxin || vyin || vzin;
        const float cxy = (vxin || vyin) ? txy[c] : 0.f;
        const float cxz = (vxin || vzin) ? txz[c] : 0.f;
        const float cyz = (vyin || vzin) ? tyz[c] : 0.f;
        (void)any;
        if (vxin) {
            float dxx = __fmul_rn(__fsub_rn(txx[c], txx[c - sx]), g.dxI);      // @d_xi(tauxx)
            dxx = cpml<2>(g, a.pv[0], dxx, k, j, i, 1, nx - 1, b);
            float dxy = __fmul_rn(__fsub_rn(txy[c + sy], cxy), g.dyI);         // @d_ya(tauxy)
            dxy = cpml<1>(g, a.pv[1], dxy, k, j, i, 1, ny - 2, b);
            float dxz = __fmul_rn(__fsub_rn(txz[c + 1], cxz), g.dzI);          // @d_za(tauxz)
            dxz = cpml<0>(g, a.pv[2], dxz, k, j, i, 1, nz - 2, b);
            nvx = __fsub_rn(nvx, __fmul_rn(__ldg(a.c[C_BX] + c), __fadd_rn(__fadd_rn(dxx, dxy), dxz)));
        }
        if (vyin) {
            float dyx = __fmul_rn(__fsub_rn(txy[c + sx], cxy), g.dxI);         // @d_xa(tauxy)
            dyx = cpml<2>(g, a.pv[3], dyx, k, j, i, 1, nx - 2, b);
            float dyy = __fmul_rn(__fsub_rn(tyy[c], tyy[c - sy]), g.dyI);      // @d_yi(tauyy)
            dyy = cpml<1>(g, a.pv[4], dyy, k, j, i, 1, ny - 1, b);
            float dyz = __fmul_rn(__fsub_rn(tyz[c + 1], cyz), g.dzI);          // @d_za(tauyz)
            dyz = cpml<0>(g, a.pv[5], dyz, k, j, i, 1, nz - 2, b);
            nvy = __fsub_rn(nvy, __fmul_rn(__ldg(a.c[C_BY] + c), __fadd_rn(__fadd_rn(dyx, dyy), dyz)));
        }
        if (vzin) {
            float dzx = __fmul_rn(__fsub_rn(txz[c + sx], cxz), g.dxI);         // @d_xa(tauxz)
            dzx = cpml<2>(g, a.pv[6], dzx, k, j, i, 1, nx - 2, b);
            float dzy = __fmul_rn(__fsub_rn(tyz[c + sy], cyz), g.dyI);         // @d_ya(tauyz)
            dzy = cpml<1>(g, a.pv[7], dzy, k, j, i, 1, ny - 2, b);
            float dzz = __fmul_rn(__fsub_rn(tzz[c], tzz[c - 1]), g.dzI);       // @d_zi(tauzz)
            dzz = cpml<0>(g, a.pv[8], dzz, k, j, i, 1, nz - 1, b);
            nvz = __fsub_rn(nvz, __fmul_rn(__ldg(a.c[C_BZ] + c), __fadd_rn(__fadd_rn(dzx, dzy), dzz)));
        }
    }

    // ---- rigid faces (dirichlet.jl:35-74), reference call order x, (y,) z; later faces override.
    // Ghost entries (v?[1] / v?[n+1] along their own axis) are written by the thread that owns the
    // mirrored inner node, so no thread reads a neighbour's new value.
    const int R = g.rigid;
    const bool tz0 = (R & ZMIN) && k == 0, tz1 = (R & ZMAX) && k == nz - 1;
    const bool ty0 = ND == 3 && (R & YMIN) && j == 0, ty1 = ND == 3 && (R & YMAX) && j == ny - 1;
    const bool tx0 = (R & XMIN) && i == 0, tx1 = (R & XMAX) && i == nx - 1;
    // the tangential zeroing of face q covers the tauii index range of the other two axes
    const bool injj = (ND == 2) || j <= ny - 1;
    if (vxown) {
        float val = nvx;
        // zeroed by y faces (i <= nx-1) and z faces (i <= nx-1)
        const bool zero = i <= nx - 1 && (ty0 || ty1 || tz0 || tz1);
        const bool ghost_min = (R & XMIN) && i == 0, ghost_max = (R & XMAX) && i == nx;
        if (zero) val = 0.f;
        if (!(ghost_min || ghost_max)) vx[c] = val;
        if ((R & XMIN) && i == 1) vx[c - sx] = (ty0 || ty1 || tz0 || tz1) ? 0.f : -nvx;     // vx[1] = -vx[2]
        if ((R & XMAX) && i == nx - 1) vx[c + sx] = -nvx;                                   // vx[n+1] = -vx[n]
    }
    if (vzown) {
        float val = nvz;
        const bool zero = k <= nz - 1 && (tx0 || tx1 || ty0 || ty1);
        const bool ghost_min = (R & ZMIN) && k == 0, ghost_max = (R & ZMAX) && k == nz;
        if (zero) val = 0.f;
        if (!(ghost_min || ghost_max)) vz[c] = val;
        // the z faces are applied last: their ghosts mirror the value left by the x/y faces
        if ((R & ZMIN) && k == 1) vz[c - 1] = -val;
        if ((R & ZMAX) && k == nz - 1) vz[c + 1] = -val;
    }
    if (ND == 3 && vyown) {
        float val = nvy;
        const bool zero_before = j <= ny - 1 && (tx0 || tx1);      // x faces run before the y faces
        const bool zero_after  = j <= ny - 1 && (tz0 || tz1);      // z faces run after
        const bool ghost_min = (R & YMIN) && j == 0, ghost_max = (R & YMAX) && j == ny;
        if (zero_before) val = 0.f;
        const float mirrored = val;
        if (zero_after) val = 0.f;
        if (!(ghost_min || ghost_max)) vy[c] = val;
        if ((R & YMIN) && j == 1) vy[c - sy] = zero_after ? 0.f : -mirrored;
        if ((R & YMAX) && j == ny - 1) vy[c + sy] = -mirrored;
    }
    (void)injj;
}

template <int ND, int EL>
__global__ void __launch_bounds__(256) k_vel(const Geom g, const StepArgs a) {
    int k, j, i, b;
    if (!cell<ND>(g, a.nbatch, k, j, i, b)) return;
    vel_cell<ND, EL>(g, a, k, j, i, b);
}

// ------------------------------------------------------------------------------------------------
// k_stress: update_dv! + update_stress! (+ free surface) fused.
//   acoustic (advance_acou.jl:288-313): p = p + (dvxdx + dvzdz [+ dvydy]) * dtK
//   elastic  (advance_elastic.jl:155-230)
// CPML term order in a.ps:
//   0 dvxdx 1 dvydy 2 dvzdz | 3 dvxdy 4 dvydx (tauxy) | 5 dvxdz 6 dvzdx (tauxz) | 7 dvydz 8 dvzdy (tauyz)
// ------------------------------------------------------------------------------------------------
template <int ND, int EL, int OOP = 0>
__device__ __forceinline__ void stress_cell(const Geom& g, const StepArgs& a, int kl, int j, int i, int b) {
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, kl, j, i);
    const int k = kl + g.koff;                  // global z index
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;

    const float* vx = a.v[V_X] + w; const float* vz = a.v[V_Z] + w; const float* vy = (ND == 3) ? a.v[V_Y] + w : nullptr;
    const bool nin = k <= nz - 1 && (ND == 2 || j <= ny - 1) && i <= nx - 1;       // tauii / p nodes

    float cvx = 0.f, cvy = 0.f, cvz = 0.f;
    const bool need_c = k <= nz - 1 || i <= nx - 1;    // cheap superset
    if (need_c) {
        if (k <= nz - 1 && (ND == 2 || j <= ny - 1)) cvx = vx[c];
        if (i <= nx - 1 && (ND == 2 || j <= ny - 1)) cvz = vz[c];
        if (ND == 3 && k <= nz - 1 && i <= nx - 1) cvy = vy[c];
    }

    float dxx = 0.f, dyy = 0.f, dzz = 0.f;
    if (nin) {
        dxx = __fmul_rn(__fsub_rn(vx[c + sx], cvx), g.dxI);                    // @d_xa(vx)
        dxx = cpml<2>(g, a.ps[0], dxx, k, j, i, 0, nx, b);
        if (ND == 3) {
            dyy = __fmul_rn(__fsub_rn(vy[c + sy], cvy), g.dyI);                // @d_ya(vy)
            dyy = cpml<1>(g, a.ps[1], dyy, k, j, i, 0, ny, b);
        }
        dzz = __fmul_rn(__fsub_rn(vz[c + 1], cvz), g.dzI);                     // @d_za(vz)
        dzz = cpml<0>(g, a.ps[2], dzz, k, j, i, 0, nz, b);
    }

    if (!EL) {
        if (ND == 2 && nin && a.dout[0]) { a.dout[0][c + (long long)b * a.dstride] = dxx; a.dout[1][c + (long long)b * a.dstride] = dzz; }
        if (nin) {
            const float* p = a.tau[T_XX] + w;
            float* po = (OOP ? a.tau_o[T_XX] : a.tau[T_XX]) + w;
            const float s = (ND == 3) ? __fadd_rn(__fadd_rn(dxx, dzz), dyy) : __fadd_rn(dxx, dzz);
            po[c] = __fadd_rn(p[c], __fmul_rn(s, __ldg(a.c[C_K] + c)));
        }
        return;
    }

    const bool fs = (g.freesurf & ZMIN) != 0;
    if (nin) {
        const float M = __ldg(a.c[C_K] + c), L = __ldg(a.c[C_L] + c);
        const float* txx = a.tau[T_XX] + w; const float* tzz = a.tau[T_ZZ] + w;
        float* oxx = (OOP ? a.tau_o[T_XX] : a.tau[T_XX]) + w; float* ozz = (OOP ? a.tau_o[T_ZZ] : a.tau[T_ZZ]) + w;
        float nzz;
        if (ND == 3) {
            const float* tyy = a.tau[T_YY] + w;
            float* oyy = (OOP ? a.tau_o[T_YY] : a.tau[T_YY]) + w;
            oxx[c] = __fsub_rn(__fsub_rn(txx[c], __fmul_rn(M, dxx)), __fmul_rn(L, __fadd_rn(dyy, dzz)));
            oyy[c] = __fsub_rn(__fsub_rn(tyy[c], __fmul_rn(M, dyy)), __fmul_rn(L, __fadd_rn(dxx, dzz)));
            nzz    = __fsub_rn(__fsub_rn(tzz[c], __fmul_rn(M, dzz)), __fmul_rn(L, __fadd_rn(dyy, dxx)));
        } else {
            oxx[c] = __fsub_rn(__fsub_rn(txx[c], __fmul_rn(M, dxx)), __fmul_rn(L, dzz));
            nzz    = __fsub_rn(__fsub_rn(tzz[c], __fmul_rn(M, dzz)), __fmul_rn(L, dxx));
        }
        // free surface (advance_elastic.jl:215-230): tauzz[1] = -tauzz[2], written by the owner of node 2
        if (!(fs && k == 0)) ozz[c] = nzz;
        if (fs && k == 1) ozz[c - 1] = -nzz;
    }
    // shear stresses on their own (half) grids
    const bool jh = (ND == 2) || (j >= 1 && j <= ny - 1);     // half nodes along y
    const bool jj = (ND == 2) || (j >= 1 && j <= ny - 2);     // inner nodes along y
    // tauxz: z half, y inner, x half
    if (k >= 1 && k <= nz - 1 && jj && i >= 1 && i <= nx - 1) {
        const float* txz = a.tau[T_XZ] + w;
        float* oxz = (OOP ? a.tau_o[T_XZ] : a.tau[T_XZ]) + w;
        float dxz = __fmul_rn(__fsub_rn(cvx, vx[c - 1]), g.dzI);               // @d_zi(vx)
        dxz = cpml<0>(g, a.ps[5], dxz, k, j, i, 1, nz - 1, b);
        float dzx = __fmul_rn(__fsub_rn(cvz, vz[c - sx]), g.dxI);              // @d_xi(vz)
        dzx = cpml<2>(g, a.ps[6], dzx, k, j, i, 1, nx - 1, b);
        float n = __fsub_rn(txz[c], __fmul_rn(__ldg(a.c[C_MUXZ] + c), __fadd_rn(dxz, dzx)));
        if (fs && k == 1) n = 0.f;                                             // free_surface!(tauxz)
        oxz[c] = n;
    }
    if (ND == 3) {
        // tauxy: z inner, y half, x half
        if (k >= 1 && k <= nz - 2 && jh && i >= 1 && i <= nx - 1) {
            const float* txy = a.tau[T_XY] + w;
            float* oxy = (OOP ? a.tau_o[T_XY] : a.tau[T_XY]) + w;
            float dxy = __fmul_rn(__fsub_rn(cvx, vx[c - sy]), g.dyI);          // @d_yi(vx)
            dxy = cpml<1>(g, a.ps[3], dxy, k, j, i, 1, ny - 1, b);
            float dyx = __fmul_rn(__fsub_rn(cvy, vy[c - sx]), g.dxI);          // @d_xi(vy)
            dyx = cpml<2>(g, a.ps[4], dyx, k, j, i, 1, nx - 1, b);
            oxy[c] = __fsub_rn(txy[c], __fmul_rn(__ldg(a.c[C_MUXY] + c), __fadd_rn(dxy, dyx)));
        }
        // tauyz: z half, y half, x inner
        if (k >= 1 && k <= nz - 1 && jh && i >= 1 && i <= nx - 2) {
            const float* tyz = a.tau[T_YZ] + w;
            float* oyz = (OOP ? a.tau_o[T_YZ] : a.tau[T_YZ]) + w;
            float dyz = __fmul_rn(__fsub_rn(cvy, vy[c - 1]), g.dzI);           // @d_zi(vy)
            dyz = cpml<0>(g, a.ps[7], dyz, k, j, i, 1, nz - 1, b);
            float dzy = __fmul_rn(__fsub_rn(cvz, vz[c - sy]), g.dyI);          // @d_yi(vz)
            dzy = cpml<1>(g, a.ps[8], dzy, k, j, i, 1, ny - 1, b);
            float n = __fsub_rn(tyz[c], __fmul_rn(__ldg(a.c[C_MUYZ] + c), __fadd_rn(dyz, dzy)));
            if (fs && k == 1) n = 0.f;                                         // free_surface!(tauyz)
            oyz[c] = n;
        }
    }
}

template <int ND, int EL>
__global__ void __launch_bounds__(256) k_stress(const Geom g, const StepArgs a) {
    int k, j, i, b;
    if (!cell<ND>(g, a.nbatch, k, j, i, b)) return;
    stress_cell<ND, EL>(g, a, k, j, i, b);
}

#include "kernels3d.cuh"
#include "kernels2v.cuh"
#include "kernels2a.cuh"
#if !defined(GPI_HOST_EMU) || defined(GPI_EMU_T3)   // tests/emu: the kernel-level harnesses leave the TMA kernels out; the emulated engine (cuda_rt_shim.h) takes their host forms
// Tile geometry and warp layout of the TMA tile kernels (kernels3t.cuh): ZC z cells x R rows per tile, MINB resident CTAs per SM;
// layouts as four decimal digits NCOMP consumer warps per row, NPROD producer warps, NSHELL shell warps, SHELLC.  Measured on B200
// (profiles/r02/tuning.md): 4 + 2 + 2 warps at 128 registers hide the shell; more consumer or producer warps, 8-row tiles and
// 96-cell tiles for narrow z-slab windows all run at the same ~0.95 us per tile and SM -- the rows the TMA unit moves bound the tile rate.
#ifndef T3_V_LAYOUT
#define T3_V_LAYOUT 1220
#endif
#ifndef T3_S_LAYOUT
#define T3_S_LAYOUT 1220
#endif
#ifndef T3_R
#define T3_R 4
#endif
#ifndef T3_MINB
#define T3_MINB 2
#endif
#ifndef T3_ZC
#define T3_ZC 128
#endif
#define T3_NS t3
#include "kernels3t.cuh"
#endif
#include "kernels4.cuh"
#include "kernels4v.cuh"

// ------------------------------------------------------------------------------------------------
// k_dmod: update_dmod! + store_invav*! (medium.jl:143-221).  The reference's `dt / @av_*(b)` and
// `inv(a) + 2.0*inv(b)` carry Float64 literals, so those are evaluated in double and rounded once.
// mod slots: 0 invK|invlambda, 1 rho, 2 invmu
// ------------------------------------------------------------------------------------------------

template <int ND, int EL>
__global__ void k_dmod(const Geom g, const float* __restrict__ m0, const float* __restrict__ rho,
                       const float* __restrict__ imu, float* const* __restrict__ out, float dt) {
    int k, j, i, b;
    if (!cell<ND>(g, 1, k, j, i, b)) return;
    const long long c = uidx(g, k, j, i);
    k += g.koff;                                // global z index for the range predicates
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const bool iny = (ND == 2) || (j >= 1 && j <= ny - 2);
    const bool inyh = (ND == 2) || (j >= 1 && j <= ny - 1);
    if (k >= 1 && k <= nz - 2 && iny && i >= 1 && i <= nx - 1)         // @av_xi(rho) on the vx interior
        out[C_BX][c] = inv_av(dt, __fadd_rn(rho[c - sx], rho[c]), 0.5f);
    if (k >= 1 && k <= nz - 1 && iny && i >= 1 && i <= nx - 2)         // @av_zi(rho)
        out[C_BZ][c] = inv_av(dt, __fadd_rn(rho[c - 1], rho[c]), 0.5f);
    if (ND == 3 && k >= 1 && k <= nz - 2 && inyh && i >= 1 && i <= nx - 2)   // @av_yi(rho)
        out[C_BY][c] = inv_av(dt, __fadd_rn(rho[c - sy], rho[c]), 0.5f);
    const bool nin = k <= nz - 1 && (ND == 2 || j <= ny - 1) && i <= nx - 1;
    if (nin) {
        if (!EL) out[C_K][c] = __fmul_rn(__fdiv_rn(1.0f, m0[c]), dt);                                  // dtK
        else {
            const float il = __fdiv_rn(1.0f, m0[c]), im = __fdiv_rn(1.0f, imu[c]);
            out[C_L][c] = __fmul_rn(il, dt);                                                           // dtlambda
            out[C_K][c] = __fmul_rn((float)((double)il + 2.0 * (double)im), dt);                       // dtM
        }
    }
    if (EL) {
        if (ND == 2) {
            if (k >= 1 && k <= nz - 1 && i >= 1 && i <= nx - 1) {      // @av(invmu) (diff2D.jl:222-225)
                const float s = __fadd_rn(__fadd_rn(__fadd_rn(imu[c - 1 - sx], imu[c - sx]), imu[c - 1]), imu[c]);
                out[C_MUXZ][c] = inv_av(dt, s, 0.25f);
            }
        } else {
            if (k >= 1 && k <= nz - 1 && j >= 1 && j <= ny - 2 && i >= 1 && i <= nx - 1) {   // @av_xzi
                const float s = __fadd_rn(__fadd_rn(__fadd_rn(imu[c - 1 - sx], imu[c - sx]), imu[c - 1]), imu[c]);
                out[C_MUXZ][c] = inv_av(dt, s, 0.25f);
            }
            if (k >= 1 && k <= nz - 2 && j >= 1 && j <= ny - 1 && i >= 1 && i <= nx - 1) {   // @av_xyi
                const float s = __fadd_rn(__fadd_rn(__fadd_rn(imu[c - sy - sx], imu[c - sx]), imu[c - sy]), imu[c]);
                out[C_MUXY][c] = inv_av(dt, s, 0.25f);
            }
            if (k >= 1 && k <= nz - 1 && j >= 1 && j <= ny - 1 && i >= 1 && i <= nx - 2) {   // @av_yzi
                const float s = __fadd_rn(__fadd_rn(__fadd_rn(imu[c - 1 - sy], imu[c - sy]), imu[c - 1]), imu[c]);
                out[C_MUYZ][c] = inv_av(dt, s, 0.25f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sources and receivers: one block per shot of the batch; inject, barrier, record.
// Replaces the reference's dense SpMV into a full-grid buffer plus a full-grid muladd sweep
// (source.jl:61-177) by O(ns * 2^N) scattered updates with the same per-cell summation order,
// and the per-step SpMV' (receiver.jl:3-14) by direct gathers into the [nt, nr] record block.
// ------------------------------------------------------------------------------------------------
struct InjOp {
    int kind;                 // 0: velocity (muladd_with_density_v?!), 1: stress (muladd_tauii!)
    int axis;                 // velocity: axis of the 2-point rho average (0 z, 1 y, 2 x)
    int ntarget;              // 1 (velocity, p) or 2/3 (elastic: every normal stress)
    float* target[3];
    const float* coef;        // stress: dtK / dtM ; velocity: rho
    int nrows;                // distinct cells hit
    const int* row_cell;      // unified linear index of each cell (fits int for <= 2^31 cells; checked on host)
    const int* row_ptr;       // nrows+1
    const int* ent_col;       // source index (column) of each entry, column order preserved
    const float* ent_val;
    const float* wav;         // [nt, ns]
};
struct RecOp {
    const float* field;
    int nr;
    const int* colptr;        // nr+1 (0-based)
    const int* tap_cell;      // unified linear index per stored entry, CSC order
    const float* tap_val;
    float* rec;               // [nt, nr]
};
enum { MAX_OPS = 8 };
struct PostDesc { int ninj, nrec; InjOp inj[MAX_OPS]; RecOp rec[MAX_OPS]; };

// woff: floats added to every wavefield pointer of the descriptors (0, or the distance to the other time-level set in ping-pong runs)
//
// The kernel is a chain of dependent loads (descriptor -> row / column pointers -> entries -> wavelet / field -> store), 10 - 13 us per
// launch when written naively and 17 % of a 16-shot 2-D time step.  Everything but the wavefields and the wavelet sample is the same at
// every step, so each thread first requests the static tables of BOTH phases (its injection row: pointers, cell, averaging operands; its
// receiver: column pointers, tap cells and weights, up to PRE_TAPS of them for the first PRE_OPS record operators) and only then walks
// the injection chain; after the barrier a receiver costs one round trip (the field values).  Summation orders are unchanged.
enum { PRE_OPS = 3, PRE_TAPS = 8 };
__global__ void k_post(const Geom g, const PostDesc* __restrict__ descs, int it /*1-based*/, int rec_it /*1-based row to write*/,
                       int nt, float dt, int flags /* bit0 inject, bit1 record */, long long woff, int i_lo = 0, int i_hi = 0x7fffffff /* inject into x planes [i_lo, i_hi) only */) {
    const PostDesc& d = descs[blockIdx.x];
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int ninj = (flags & 1) ? d.ninj : 0;
    const int nrec = ((flags & 2) && rec_it >= 1 && rec_it <= nt) ? d.nrec : 0;
    // ---- static tables of this thread's first receiver of the first record operators
    int pn[PRE_OPS], pcell[PRE_OPS][PRE_TAPS];
    float pval[PRE_OPS][PRE_TAPS];
#pragma unroll
    for (int o = 0; o < PRE_OPS; o++) {
        pn[o] = -1;
        if (o < nrec && (int)threadIdx.x < d.rec[o].nr) {
            const RecOp& op = d.rec[o];
            const int e0 = op.colptr[threadIdx.x], n = op.colptr[threadIdx.x + 1] - e0;
            if (n <= PRE_TAPS) {
                pn[o] = n;
#pragma unroll
                for (int q = 0; q < PRE_TAPS; q++) if (q < n) { pcell[o][q] = op.tap_cell[e0 + q]; pval[o][q] = op.tap_val[e0 + q]; }
            }
        }
    }
    // the tables above are written before the run only; the wavefields and records below by the kernels ahead in the stream
    pdl_wait(); pdl_release();
    // ---- sources
    for (int o = 0; o < ninj; o++) {
        const InjOp& op = d.inj[o];
        for (int r = threadIdx.x; r < op.nrows; r += blockDim.x) {
            const int e0 = op.row_ptr[r], e1 = op.row_ptr[r + 1];
            const long long c = op.row_cell[r];
            { const int ip = (int)(c / sx); if (ip < i_lo || ip >= i_hi) continue; }
            const long long off = op.axis == 0 ? 1 : (op.axis == 1 ? sy : sx);
            // operands that do not depend on the wavelet travel while the row is summed
            const float c0 = op.kind == 1 ? op.coef[c] : op.coef[c - (g.h + 1) * off];
            const float c1 = op.kind == 1 ? 0.f : op.coef[c - g.h * off];
            float old[3];
            for (int t = 0; t < op.ntarget; t++) old[t] = op.target[t][c + woff];
            float buf = 0.f;                                         // mul!(buf, S, w): buf[row] += nzval * w[col]
            for (int e = e0; e < e1; e++)
                buf = __fadd_rn(buf, __fmul_rn(op.ent_val[e], op.wav[(size_t)(it - 1) + (size_t)nt * op.ent_col[e]]));
            if (op.kind == 1) {
                const float add = __fmul_rn(buf, c0);                // pw = pw + (pv * dtK)
                for (int t = 0; t < op.ntarget; t++) op.target[t][c + woff] = __fadd_rn(old[t], add);
            } else {                                                 // pw = pw + (pv / av(rho) * dt), literal typing as in the reference (wide_t)
                const float s = __fadd_rn(c0, c1);                   // @av_?i: integer nodes u-1-h, u-h
                // every operation rounded on its own (nvcc would otherwise be free to contract the product and the sum into one fma,
                // which the CPU path does not do)
                op.target[0][c + woff] = (float)wadd((wide_t)old[0], wmul(wdiv((wide_t)buf, wmul((wide_t)s, (wide_t)0.5f)), (wide_t)dt));
            }
        }
    }
    __syncthreads();
    // ---- receivers
    for (int o = 0; o < nrec; o++) {
        const RecOp& op = d.rec[o];
        for (int ir = threadIdx.x; ir < op.nr; ir += blockDim.x) {
            float tmp = 0.f;                                         // mul!(rec, transpose(R), field)
            bool done = false;
#pragma unroll
            for (int po = 0; po < PRE_OPS; po++) if (po == o && ir == (int)threadIdx.x && pn[po] >= 0) {
                float fv[PRE_TAPS];
#pragma unroll
                for (int q = 0; q < PRE_TAPS; q++) if (q < pn[po]) fv[q] = op.field[pcell[po][q] + woff];
#pragma unroll
                for (int q = 0; q < PRE_TAPS; q++) if (q < pn[po]) tmp = __fadd_rn(tmp, __fmul_rn(pval[po][q], fv[q]));
                done = true;
            }
            if (!done)
                for (int e = op.colptr[ir]; e < op.colptr[ir + 1]; e++)
                    tmp = __fadd_rn(tmp, __fmul_rn(op.tap_val[e], op.field[op.tap_cell[e] + woff]));
            op.rec[(size_t)(rec_it - 1) + (size_t)nt * ir] = tmp;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// boundary store for time reversal (boundary.jl:17-52, 217-264): 3+3 planes per axis just inside the
// PML, saved negated in :forward_save and forced back in :adjoint.  Store layout (engine's own):
// axis x: [6][ny1][pz] ; axis y: [nx1][6][pz] ; axis z: [6][ny1][nx1].
// ------------------------------------------------------------------------------------------------
// One launch covers every (shot of the batch, stored field, axis, plane): blockIdx.z folds them, so a time step
// of :forward_save / :adjoint costs one boundary launch instead of (shots x fields x axes).  The planes of
// different axes overlap at the corners of the box; a save reads the same field there and a force writes the same
// saved value, so the reference's x, (y,) z call order does not matter.
struct BndField {
    float* f0;                // batch slot 0 of the field
    int lo[3], hi[3];         // first unified index of the min / max triple along each axis
    int k0, j0, i0, nk, nj, ni;   // field extents in unified coordinates [lower, upper)
};
struct BndArgs {
    int nf, naxes, nbound;
    int axes[3];              // axes that exist (2-D: z, x)
    BndField f[6];            // 3-D elastic stores all six stresses
    float* const* stores;     // [b][field][axis 0..2]: the shot's store of nt slots
    long long slot_off[3];    // (slot index) x (floats per slot) per axis for this time step
    long long wstride;        // floats between batch slots of the wavefield set
};
template <int SAVE>
__global__ void k_boundary(const Geom g, const BndArgs a) {
    pdl_wait(); pdl_release();
    int z = blockIdx.z;
    const int p = z % (2 * a.nbound); z /= 2 * a.nbound;      // plane 0..2*nbound-1
    const int ia = z % a.naxes; z /= a.naxes;
    const int fi = z % a.nf; const int b = z / a.nf;
    const int axis = a.axes[ia];
    const BndField& F = a.f[fi];
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y;
    const int q = p < a.nbound ? F.lo[axis] + p : F.hi[axis] + p - a.nbound;
    int k, j, i;
    long long si;
    if (axis == 2)      { k = t0; j = t1; i = q; if (k >= g.pz || j >= g.ny1) return; si = (long long)k + (long long)g.pz * (j + (long long)g.ny1 * p); }
    else if (axis == 1) { k = t0; i = t1; j = q; if (k >= g.pz || i >= g.nx1) return; si = (long long)k + (long long)g.pz * (p + 2LL * a.nbound * i); }
    else                { i = t0; j = t1; k = q; if (i >= g.nx1 || j >= g.ny1) return; si = (long long)i + (long long)g.nx1 * (j + (long long)g.ny1 * p); }
    if (k < F.k0 || k >= F.nk || j < F.j0 || j >= F.nj || i < F.i0 || i >= F.ni) return;
    float* field = F.f0 + (long long)b * a.wstride + uidx(g, k, j, i);
    float* store = a.stores[(b * a.nf + fi) * 3 + axis] + a.slot_off[axis] + si;
    if (SAVE == 2)  *store = *field;                       // plain copy: the pre-force values kept aside in ping-pong adjoint runs
    else if (SAVE)  *store = __fmul_rn(*field, -1.0f);     // rmul!(b, -1)
    else            *field = *store;
}

// ------------------------------------------------------------------------------------------------
// gradient imaging, 2-D acoustic (gradient.jl:17-56), one pass over the grid:
//   g_invK += p2_tp * (p1_tp - p1) * dtI
//   g_rho_inn -= av_xi(bx) + av_zi(bz),  b? = v?2_tp * (v?1 - v?1_tp) * dtI   (buffers recomputed on the fly)
// The reference's one-cell shift in combine_gmodrho! (gradient.jl:53-56) is kept.
// ------------------------------------------------------------------------------------------------
__global__ void k_grad2d(const Geom g, const float* __restrict__ p1, const float* __restrict__ p1tp, const float* __restrict__ p2tp,
                         const float* __restrict__ vx1, const float* __restrict__ vx1tp, const float* __restrict__ vx2tp,
                         const float* __restrict__ vz1, const float* __restrict__ vz1tp, const float* __restrict__ vz2tp,
                         float* __restrict__ gK, float* __restrict__ gR, float dtI, long long wstride, long long gstride, int unshifted) {
    int k, j, i, b;
    if (!cell<2>(g, 1, k, j, i, b)) return;
    if (k > g.nz - 1 || i > g.nx - 1) return;            // (k, i): tauii node; its storage index is shifted by g.h at order 4
    const long long w = (long long)b * wstride, gw = (long long)b * gstride;
    const long long c = uidx(g, k + g.h, 0, i + g.h), sx = g.pz;
    gK[gw + c] = __fadd_rn(gK[gw + c], __fmul_rn(__fmul_rn(p2tp[w + c], __fsub_rn(p1tp[w + c], p1[w + c])), dtI));
    if (unshifted) {
        // GPI_RUN_UNSHIFTED_RHO: cell (k, i) takes the velocity nodes that bound it -- vx nodes i and i+1, vz nodes k and
        // k+1, interior nodes only -- which makes the imaging the exact transpose of the FD-Born map
        auto bufx = [&](int kk, int ii) { const long long q = w + uidx(g, kk, 0, ii);
            return (kk >= 1 && kk <= g.nz - 2 && ii >= 1 && ii <= g.nx - 1) ? __fmul_rn(__fmul_rn(vx2tp[q], __fsub_rn(vx1[q], vx1tp[q])), dtI) : 0.f; };
        auto bufz = [&](int kk, int ii) { const long long q = w + uidx(g, kk, 0, ii);
            return (kk >= 1 && kk <= g.nz - 1 && ii >= 1 && ii <= g.nx - 2) ? __fmul_rn(__fmul_rn(vz2tp[q], __fsub_rn(vz1[q], vz1tp[q])), dtI) : 0.f; };
        const float ax = __fadd_rn(bufx(k, i), bufx(k, i + 1)), az = __fadd_rn(bufz(k, i), bufz(k + 1, i));
        gR[gw + c] = (float)wsub(wsub((wide_t)gR[gw + c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
        return;
    }
    const int o = 1 + 2 * g.h;                           // @inn(g): tauii nodes [O, n-1-O]
    if (k >= o && k <= g.nz - 1 - o && i >= o && i <= g.nx - 1 - o) {
        auto bufx = [&](long long q) { return __fmul_rn(__fmul_rn(vx2tp[w + q], __fsub_rn(vx1[w + q], vx1tp[w + q])), dtI); };
        auto bufz = [&](long long q) { return __fmul_rn(__fmul_rn(vz2tp[w + q], __fsub_rn(vz1[w + q], vz1tp[w + q])), dtI); };
        // @av_xi(vxbuffer) = vxbuffer[izi, ix] + vxbuffer[izi, ix+1]: velocity nodes 3h+1 and 3h below the cell's own index
        const long long q1 = 3 * g.h + 1, q0 = 3 * g.h;
        const float ax = __fadd_rn(bufx(c - q1 * sx), bufx(c - q0 * sx));
        const float az = __fadd_rn(bufz(c - q1), bufz(c - q0));      // @av_zi(vzbuffer)
        gR[gw + c] = (float)wsub(wsub((wide_t)gR[gw + c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
    }
}

// ------------------------------------------------------------------------------------------------
// gradient imaging, 2-D elastic (nothing upstream -- gradient.jl has acoustic methods only).  The
// adjoint-state construction of gradlame!/gradrho! written for the compliance form of the stress update:
//   e = (txx2 + tzz2)_tp * ((txx1 + tzz1)_tp - (txx1 + tzz1)) * dtI     isotropic part,  dS = dc/4, c = 1/(lambda + mu)
//   d = (txx2 - tzz2)_tp * ((txx1 - tzz1)_tp - (txx1 - tzz1)) * dtI     deviatoric part, dS = d(invmu)/4
//   s = txz2_tp * (txz1_tp - txz1) * dtI  on the shear nodes, a quarter of it to each cell of the node's @av(invmu)
//   g_invlambda += e/4 * dc/d(invlambda);  g_invmu += e/4 * dc/d(invmu) + d/4 + sum s/4;  g_rho as in k_grad2d
// with c = invlambda * invmu / (invlambda + invmu).  Operation order = oracle/fdtd_oracle.c::compute_gradient_el2d.
// ------------------------------------------------------------------------------------------------
struct GradE2Args {
    const float *xx1, *zz1, *xz1, *xx1tp, *zz1tp, *xz1tp, *xx2tp, *zz2tp, *xz2tp;
    const float *vx1, *vx1tp, *vx2tp, *vz1, *vz1tp, *vz2tp;
    const float *il, *im;                 // invlambda, invmu (shared by all shots)
    float *gL, *gM, *gR;                  // batch slot 0
    long long wstride, gstride;
};
__global__ void k_grad2d_el(const Geom g, const GradE2Args a, float dtI) {
    int k, j, i, b;
    if (!cell<2>(g, 1, k, j, i, b)) return;
    if (k > g.nz - 1 || i > g.nx - 1) return;            // (k, i): tauii node
    const long long w = (long long)b * a.wstride, gw = (long long)b * a.gstride;
    const long long c = uidx(g, k + g.h, 0, i + g.h), sx = g.pz;
    const long long q = c + w;
    const float e = __fmul_rn(__fmul_rn(__fadd_rn(a.xx2tp[q], a.zz2tp[q]), __fsub_rn(__fadd_rn(a.xx1tp[q], a.zz1tp[q]), __fadd_rn(a.xx1[q], a.zz1[q]))), dtI);
    const float d = __fmul_rn(__fmul_rn(__fsub_rn(a.xx2tp[q], a.zz2tp[q]), __fsub_rn(__fsub_rn(a.xx1tp[q], a.zz1tp[q]), __fsub_rn(a.xx1[q], a.zz1[q]))), dtI);
    const float la = a.il[c], mb = a.im[c], ab = __fadd_rn(la, mb);
    const float ab2 = __fmul_rn(ab, ab);
    const float dca = __fdiv_rn(__fmul_rn(mb, mb), ab2), dcb = __fdiv_rn(__fmul_rn(la, la), ab2);
    const float qe = __fmul_rn(0.25f, e);
    a.gL[gw + c] = __fadd_rn(a.gL[gw + c], __fmul_rn(qe, dca));
    float gm = __fadd_rn(__fadd_rn(a.gM[gw + c], __fmul_rn(qe, dcb)), __fmul_rn(0.25f, d));
    // shear nodes whose @av(invmu) contains this cell: array indices (iz - dz, ix - dx), dz, dx in {0, 1}, i.e. unified
    // (k + 1 + h - dz, i + 1 + h - dx); the tauxz array covers unified [1 + h, n - 1 - h]
    float acc = 0.f;
#pragma unroll
    for (int dx = 0; dx < 2; dx++)
#pragma unroll
        for (int dz = 0; dz < 2; dz++) {
            const int ku = k + 1 + g.h - dz, iu = i + 1 + g.h - dx;
            if (ku < 1 + g.h || ku > g.nz - 1 - g.h || iu < 1 + g.h || iu > g.nx - 1 - g.h) continue;
            const long long s = uidx(g, ku + g.h, 0, iu + g.h) + w;
            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(a.xz2tp[s], __fsub_rn(a.xz1tp[s], a.xz1[s])), dtI));
        }
    a.gM[gw + c] = __fadd_rn(gm, __fmul_rn(0.25f, acc));
    const int o = 1 + 2 * g.h;
    if (k >= o && k <= g.nz - 1 - o && i >= o && i <= g.nx - 1 - o) {
        auto bufx = [&](long long x) { return __fmul_rn(__fmul_rn(a.vx2tp[w + x], __fsub_rn(a.vx1[w + x], a.vx1tp[w + x])), dtI); };
        auto bufz = [&](long long x) { return __fmul_rn(__fmul_rn(a.vz2tp[w + x], __fsub_rn(a.vz1[w + x], a.vz1tp[w + x])), dtI); };
        const long long q1 = 3 * g.h + 1, q0 = 3 * g.h;
        const float ax = __fadd_rn(bufx(c - q1 * sx), bufx(c - q0 * sx));
        const float az = __fadd_rn(bufz(c - q1), bufz(c - q0));
        a.gR[gw + c] = (float)wsub(wsub((wide_t)a.gR[gw + c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
    }
}

// ------------------------------------------------------------------------------------------------
// gradient imaging, 3-D elastic: k_grad2d_el with the 3-D isotropic compliance
//   eps = dev(tau) / (2 mu) + tr(tau) / (3 (3 lambda + 2 mu)) I,   c3 = invlambda invmu / (3 invmu + 2 invlambda)
//   eT = T2_tp (T1_tp - T1) dtI, T = txx + tyy + tzz;   eD = sum_i dev2_ii_tp (dev1_ii_tp - dev1_ii) dtI, dev_ii = t_ii - T/3
//   g_invlambda += eT/3 dc3/dinvlambda;   g_invmu += eT/3 dc3/dinvmu + eD/2 + (shear nodes of tauxz, tauxy, tauyz)/4;  g_rho as k_grad3d
// Operation order = oracle/fdtd_oracle.c::compute_gradient_el3d.
// ------------------------------------------------------------------------------------------------
struct GradE3Args {
    const float *t1[6], *t1tp[6], *t2tp[6];      // T_XX .. T_YZ slots
    const float *v1[3], *v1tp[3], *v2tp[3];
    const float *il, *im;
    float *gL, *gM, *gR;
};
__global__ void k_grad3d_el(const Geom g, const GradE3Args a, float dtI) {
    int k, j, i, b;
    if (!cell<3>(g, 1, k, j, i, b)) return;
    if (k > g.nz - 1 || j > g.ny - 1 || i > g.nx - 1) return;
    const int h = g.h, O = 1 + 2 * g.h;
    const long long c = uidx(g, k + h, j + h, i + h), sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nrm[3] = {T_XX, T_YY, T_ZZ};
    float t1 = 0.f, t1p = 0.f, t2p = 0.f;
#pragma unroll
    for (int q = 0; q < 3; q++) { t1 = __fadd_rn(t1, a.t1[nrm[q]][c]); t1p = __fadd_rn(t1p, a.t1tp[nrm[q]][c]); t2p = __fadd_rn(t2p, a.t2tp[nrm[q]][c]); }
    const float third = __fdiv_rn(1.0f, 3.0f);
    const float eT = __fmul_rn(__fmul_rn(t2p, __fsub_rn(t1p, t1)), dtI);
    float eD = 0.f;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float d1 = __fsub_rn(a.t1[nrm[q]][c], __fmul_rn(t1, third)), d1p = __fsub_rn(a.t1tp[nrm[q]][c], __fmul_rn(t1p, third));
        const float d2p = __fsub_rn(a.t2tp[nrm[q]][c], __fmul_rn(t2p, third));
        eD = __fadd_rn(eD, __fmul_rn(__fmul_rn(d2p, __fsub_rn(d1p, d1)), dtI));
    }
    const float la = a.il[c], mb = a.im[c];
    const float den = __fadd_rn(__fmul_rn(3.0f, mb), __fmul_rn(2.0f, la)), den2 = __fmul_rn(den, den);
    const float dca = __fdiv_rn(__fmul_rn(3.0f, __fmul_rn(mb, mb)), den2), dcb = __fdiv_rn(__fmul_rn(2.0f, __fmul_rn(la, la)), den2);
    const float te = __fmul_rn(third, eT);
    a.gL[c] = __fadd_rn(a.gL[c], __fmul_rn(te, dca));
    float gm = __fadd_rn(__fadd_rn(a.gM[c], __fmul_rn(te, dcb)), __fmul_rn(0.5f, eD));
    auto inH = [&](int u, int n) { return u >= 1 + h && u <= n - 1 - h; };
    auto inJ = [&](int u, int n) { return u >= O && u <= n - 1 - O; };
    auto term = [&](int slot, int ku, int ju, int iu) {
        const long long s = uidx(g, ku + h, ju + h, iu + h);
        return __fmul_rn(__fmul_rn(a.t2tp[slot][s], __fsub_rn(a.t1tp[slot][s], a.t1[slot][s])), dtI);
    };
    float acc = 0.f;
    for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {          // tauxz: (z, x) averaged, y inner
        const int ku = k + 1 - d1 + h, iu = i + 1 - d2 + h;
        if (inH(ku, g.nz) && inJ(j, g.ny) && inH(iu, g.nx)) acc = __fadd_rn(acc, term(T_XZ, ku, j, iu));
    }
    for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {          // tauxy: (y, x) averaged, z inner
        const int ju = j + 1 - d1 + h, iu = i + 1 - d2 + h;
        if (inJ(k, g.nz) && inH(ju, g.ny) && inH(iu, g.nx)) acc = __fadd_rn(acc, term(T_XY, k, ju, iu));
    }
    for (int d2 = 0; d2 < 2; d2++) for (int d1 = 0; d1 < 2; d1++) {          // tauyz: (z, y) averaged, x inner
        const int ku = k + 1 - d1 + h, ju = j + 1 - d2 + h;
        if (inH(ku, g.nz) && inH(ju, g.ny) && inJ(i, g.nx)) acc = __fadd_rn(acc, term(T_YZ, ku, ju, i));
    }
    a.gM[c] = __fadd_rn(gm, __fmul_rn(0.25f, acc));
    if (k >= O && k <= g.nz - 1 - O && j >= O && j <= g.ny - 1 - O && i >= O && i <= g.nx - 1 - O) {
        auto buf = [&](int q, long long x) { return __fmul_rn(__fmul_rn(a.v2tp[q][x], __fsub_rn(a.v1[q][x], a.v1tp[q][x])), dtI); };
        const long long q1 = 3 * h + 1, q0 = 3 * h;
        const float ax = __fadd_rn(buf(V_X, c - q1 * sx), buf(V_X, c - q0 * sx));
        const float ay = __fadd_rn(buf(V_Y, c - q1 * sy), buf(V_Y, c - q0 * sy));
        const float az = __fadd_rn(buf(V_Z, c - q1), buf(V_Z, c - q0));
        a.gR[c] = (float)wsub(wsub(wsub((wide_t)a.gR[c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)ay, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
    }
}

// ------------------------------------------------------------------------------------------------
// gradient imaging, 3-D acoustic.  gradlame! (gradient.jl:17-29) is dimension-free; gradrho! exists
// upstream for 2-D only (gradient.jl:31,58-61).  Same construction with the y term between x and z:
//   g_rho_inn -= av_xi(bx) + av_yi(by) + av_zi(bz),   b? = v?2_tp * (v?1 - v?1_tp) * dtI
// `unshifted` (GPI_RUN_UNSHIFTED_RHO, order 2): the cell takes the interior velocity nodes that bound it instead of
// upstream's one-cell-shifted pairs, which is what finite differences of the loss confirm (tests/test_adjoint3d.py).
// ------------------------------------------------------------------------------------------------
struct Grad3Args {
    const float *p1, *p1tp, *p2tp;
    const float *v1[3], *v1tp[3], *v2tp[3];      // V_X, V_Y, V_Z
    float *gK, *gR;
};
__global__ void k_grad3d(const Geom g, const Grad3Args a, float dtI, int unshifted) {
    int k, j, i, b;
    if (!cell<3>(g, 1, k, j, i, b)) return;
    if (k > g.nz - 1 || j > g.ny - 1 || i > g.nx - 1) return;        // (k, j, i): tauii node
    const long long c = uidx(g, k + g.h, j + g.h, i + g.h), sy = g.pz, sx = (long long)g.pz * g.ny1;
    a.gK[c] = __fadd_rn(a.gK[c], __fmul_rn(__fmul_rn(a.p2tp[c], __fsub_rn(a.p1tp[c], a.p1[c])), dtI));
    auto buf = [&](int q, long long x) { return __fmul_rn(__fmul_rn(a.v2tp[q][x], __fsub_rn(a.v1[q][x], a.v1tp[q][x])), dtI); };
    if (unshifted) {
        // interior nodes of each component (@inn of compute_v!): vx (J, J, H), vy (J, H, J), vz (H, J, J)
        auto bx = [&](int ii) { return (k >= 1 && k <= g.nz - 2 && j >= 1 && j <= g.ny - 2 && ii >= 1 && ii <= g.nx - 1) ? buf(V_X, uidx(g, k, j, ii)) : 0.f; };
        auto by = [&](int jj) { return (k >= 1 && k <= g.nz - 2 && jj >= 1 && jj <= g.ny - 1 && i >= 1 && i <= g.nx - 2) ? buf(V_Y, uidx(g, k, jj, i)) : 0.f; };
        auto bz = [&](int kk) { return (kk >= 1 && kk <= g.nz - 1 && j >= 1 && j <= g.ny - 2 && i >= 1 && i <= g.nx - 2) ? buf(V_Z, uidx(g, kk, j, i)) : 0.f; };
        const float ax = __fadd_rn(bx(i), bx(i + 1)), ay = __fadd_rn(by(j), by(j + 1)), az = __fadd_rn(bz(k), bz(k + 1));
        a.gR[c] = (float)wsub(wsub(wsub((wide_t)a.gR[c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)ay, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
        return;
    }
    const int o = 1 + 2 * g.h;
    if (k >= o && k <= g.nz - 1 - o && j >= o && j <= g.ny - 1 - o && i >= o && i <= g.nx - 1 - o) {
        const long long q1 = 3 * g.h + 1, q0 = 3 * g.h;             // see k_grad2d
        const float ax = __fadd_rn(buf(V_X, c - q1 * sx), buf(V_X, c - q0 * sx));
        const float ay = __fadd_rn(buf(V_Y, c - q1 * sy), buf(V_Y, c - q0 * sy));
        const float az = __fadd_rn(buf(V_Z, c - q1), buf(V_Z, c - q0));
        a.gR[c] = (float)wsub(wsub(wsub((wide_t)a.gR[c], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)ay, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
    }
}

// ------------------------------------------------------------------------------------------------
// FD-Born scattering sources, 2-D acoustic (born.jl:1-12; upstream passes argument lists that match no kernel,
// the intent is the commented legacy code born.jl:32-99): the second wavefield is driven by the first one's
// derivatives times the perturbation of the update coefficients,
//     v?2_inn = v?2_inn + d(dt / av(rho)) * dpd?1          (after update_v!, propagate.jl:205)
//     p2      = p2 + (dvxdx1 + dvzdz1) * d(dt K)           (after update_stress!, propagate.jl:226)
// k_born_coef linearises the coefficient perturbations in (d invK, d rho), so that the map from the medium
// perturbation to the scattered data is exactly linear:
//     d(dt K) = -((K K) d invK) dt,  K = 1 / invK;     d(dt / av rho) = -(dt av(d rho)) / (av rho)^2   (Float64, rounded once)
// ------------------------------------------------------------------------------------------------
__global__ void k_born_coef(const Geom g, const float* __restrict__ invK, const float* __restrict__ rho,
                            const float* __restrict__ dinvK, const float* __restrict__ drho,
                            float* __restrict__ ddtK, float* __restrict__ dbx, float* __restrict__ dbz, float dt) {
    int k, j, i, b;
    if (!cell<2>(g, 1, k, j, i, b)) return;
    const long long c = uidx(g, k, 0, i), sx = g.pz;
    const int nz = g.nz, nx = g.nx;
    const double ddt = (double)dt;
    if (k <= nz - 1 && i <= nx - 1) {
        const float K = __fdiv_rn(1.0f, invK[c]);
        ddtK[c] = -__fmul_rn(__fmul_rn(__fmul_rn(K, K), dinvK[c]), dt);
    }
    if (k >= 1 && k <= nz - 2 && i >= 1 && i <= nx - 1) {
        const double av = (double)__fadd_rn(rho[c - sx], rho[c]) * 0.5, dav = (double)__fadd_rn(drho[c - sx], drho[c]) * 0.5;
        dbx[c] = (float)(-(ddt * dav) / (av * av));
    }
    if (k >= 1 && k <= nz - 1 && i >= 1 && i <= nx - 2) {
        const double av = (double)__fadd_rn(rho[c - 1], rho[c]) * 0.5, dav = (double)__fadd_rn(drho[c - 1], drho[c]) * 0.5;
        dbz[c] = (float)(-(ddt * dav) / (av * av));
    }
}
// KIND 0: velocities of pw 2 (d0 = dpdx1, d1 = dpdz1, c0 = dbx, c1 = dbz); KIND 1: pressure of pw 2 (d0 = dvxdx1, d1 = dvzdz1, c0 = d(dtK))
template <int KIND>
__global__ void k_born_add(const Geom g, float* __restrict__ f0, float* __restrict__ f1, const float* __restrict__ d0,
                           const float* __restrict__ d1, const float* __restrict__ c0, const float* __restrict__ c1,
                           long long wstride, long long dstride) {
    int k, j, i, b;
    if (!cell<2>(g, 1, k, j, i, b)) return;
    const long long c = uidx(g, k, 0, i), w = (long long)b * wstride, dw = (long long)b * dstride;
    const int nz = g.nz, nx = g.nx;
    if (KIND == 0) {
        if (k >= 1 && k <= nz - 2 && i >= 1 && i <= nx - 1) f0[c + w] = __fadd_rn(f0[c + w], __fmul_rn(c0[c], d0[c + dw]));
        if (k >= 1 && k <= nz - 1 && i >= 1 && i <= nx - 2) f1[c + w] = __fadd_rn(f1[c + w], __fmul_rn(c1[c], d1[c + dw]));
    } else {
        if (k <= nz - 1 && i <= nx - 1) f0[c + w] = __fadd_rn(f0[c + w], __fmul_rn(__fadd_rn(d0[c + dw], d1[c + dw]), c0[c]));
    }
}

// ------------------------------------------------------------------------------------------------
// z-slab halo planes (multi-GPU domain decomposition, no counterpart in the reference).  z is the
// fast axis, so a z = const plane is strided in memory: pack gathers up to three planes into one
// contiguous buffer [field][i][j] that travels over NVLink; unpack scatters it into the halo plane.
// ------------------------------------------------------------------------------------------------
struct HaloArgs { int n; float* field[3]; int k[3]; int i0, ni; };      // x planes [i0, i0 + ni) of the z plane; buffer [field][i - i0][j]
template <int PACK>
__global__ void k_halo(const Geom g, const HaloArgs a, float* __restrict__ buf) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y + a.i0;
    if (j >= g.ny1) return;
    const long long plane = (long long)g.ny1 * a.ni;
    const long long q = (long long)j + (long long)g.ny1 * blockIdx.y;
    for (int f = 0; f < a.n; f++) {
        float* p = a.field[f] + uidx(g, a.k[f], j, i);
        if (PACK) buf[f * plane + q] = *p;
        else      *p = buf[f * plane + q];
    }
}

// g_total += g_shot, in shot order (sum_grads!, gradient.jl:2-11)
__global__ void k_axpy1(float* __restrict__ y, const float* __restrict__ x, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = __fadd_rn(y[t], x[t]);
}

// ------------------------------------------------------------------------------------------------
// source illumination (compute_illum!, fdtd.jl:570-581: illum[i] += abs2(p[i]) at every time step, Float32 square added
// into a Float64 array; stack_illums!, fdtd.jl:556-565: summed over the supersources in shot order).  Upstream has the
// two calls commented out (propagate.jl:114,236) and allocates a dummy; `illum_flag` (fdtd.jl:59,73) documents the intent:
// the wavefield energy of pw 1 as a preconditioner.  One accumulator per resident shot (blockIdx.y).
// ------------------------------------------------------------------------------------------------
__global__ void k_illum(double* __restrict__ acc, const float* __restrict__ p, long long n, long long wstride) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float v = p[(long long)blockIdx.y * wstride + t];
    double* a = acc + (long long)blockIdx.y * n + t;
    *a = __dadd_rn(*a, (double)__fmul_rn(v, v));
}
__global__ void k_axpy1d(double* __restrict__ y, const double* __restrict__ x, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = __dadd_rn(y[t], x[t]);
}

}  // namespace gpi
