// kernels3d.cuh -- the 3-D stencil kernels of the roofline case (included inside namespace gpi).
//
// Same arithmetic as vel_cell / stress_cell (kernels.cuh), organised for HBM throughput on sm_100a:
//
//   * one thread owns FOUR consecutive z cells (z is the fast axis, pitch a multiple of 128 B), so every
//     field, coefficient and CPML-memory access is an aligned 128-bit load/store; the z-1 / z+1
//     neighbours come from the thread's own vector plus one scalar load per differentiated field;
//   * all loads of a thread are issued up front, unconditionally (18-20 independent 16-byte requests in
//     flight per thread), which is what a latency x bandwidth product of ~35 KB per SM needs -- the scalar
//     kernels issue 4-byte loads behind range predicates and reach only 20-35 % of HBM;
//   * the (z, y) plane is linearised into float4 groups so no lanes are wasted on a ragged z extent;
//     x comes from blockIdx.y (consecutive planes are scheduled back to back, so the +-1-plane
//     neighbours are L2 hits: measured DRAM traffic equals the algorithmic bytes);
//   * range predicates in z are per element and applied only at the store; rows on the outer shell
//     of the box (rigid faces, ghost cells, ragged x / y ranges; 2.4 % of the rows) run the scalar
//     reference-order code of kernels.cuh;
//   * CPML: x / y slab membership is uniform per thread (coefficients are broadcast loads, memory is
//     moved as float4); z slabs use k-indexed coefficient tables (identity outside the slab) and the
//     float4-aligned z-memory layout described at cpml<> in kernels.cuh.
//
// Arithmetic stays __fmul_rn / __fadd_rn / __fsub_rn in the reference's association order: results are
// bit-identical to the scalar kernels and to the CPU restatement.

// Launch shape, tuned on B200 (profiles/r01/tuning.md): one warp per block, 12 blocks per SM (168
// registers per thread).  Small blocks retire and refill independently, which keeps more loads in flight
// than 3 x 128 threads at the same register budget (27.6 vs 26.3 Gcell-updates/s on 338^3).
#ifndef GPI_VEC_THREADS
#define GPI_VEC_THREADS 32
#endif
#ifndef GPI_VEC_MINBLOCKS
#define GPI_VEC_MINBLOCKS 12
#endif

#ifndef GPI_VEC_W
#define GPI_VEC_W 4          // z cells per thread: 4 (128-bit accesses) or 2 (64-bit accesses, half the registers)
#endif
constexpr int VW = GPI_VEC_W;
struct F4 { float v[VW]; };
// Loads and stores of the fast path are volatile asm: the compiler keeps volatile asm statements in
// program order, so every load of a thread is issued before its first store and -- the point -- before
// the arithmetic that consumes the first loaded value.  With plain C++ loads nvcc sinks the loads whose
// values are needed last (the velocities and buoyancies of k_vel3v) below the CPML arithmetic to save
// registers, which costs a second, serialised DRAM round trip per thread.
#ifdef GPI_HOST_EMU
// tests/emu: the kernels compiled as host C++ (threads run one after the other); plain accessors, no prefetch
inline F4 ld4(const float* p) { F4 r; for (int e = 0; e < VW; e++) r.v[e] = p[e]; return r; }
inline F4 ldg4(const float* p) { return ld4(p); }
inline void st4(float* p, const F4& r) { for (int e = 0; e < VW; e++) p[e] = r.v[e]; }
inline void pf(const float*) {}
inline void pf2(const float*) {}
inline float ld1(const float* p) { return *p; }
#ifndef GPI_PF_AHEAD
#define GPI_PF_AHEAD 0
#endif
#else
#if GPI_VEC_W == 4
__device__ __forceinline__ F4 ld4(const float* p) {
    F4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ F4 ldg4(const float* p) {
    F4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(float* p, const F4& r) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
}
#else
__device__ __forceinline__ F4 ld4(const float* p) {
    F4 r;
    asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
    return r;
}
__device__ __forceinline__ F4 ldg4(const float* p) {
    F4 r;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(float* p, const F4& r) {
    asm volatile("st.global.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(r.v[0]), "f"(r.v[1]) : "memory");
}
#endif
// ptxas still sinks the loads whose values are consumed last below the CPML arithmetic; a prefetch of
// those lines, issued with the first loads, turns that second round trip into a cache hit.
#if defined(GPI_PF_NONE)
#define GPI_PF "// no prefetch %0"
#elif defined(GPI_PF_L2)
#define GPI_PF "prefetch.global.L2 [%0];"
#else
#define GPI_PF "prefetch.global.L1 [%0];"
#endif
__device__ __forceinline__ void pf(const float* p) { asm volatile(GPI_PF :: "l"(p)); }
// Plane-ahead L2 prefetch: blocks are scheduled plane by plane (blockIdx.y = x plane), so the lines this
// thread's (k, j) group will need GPI_PF_AHEAD planes later can be requested from DRAM now, with no register
// and no scoreboard cost; when the later block loads them they are L2 hits (~250 cycles instead of ~700).
#ifndef GPI_PF_AHEAD
#define GPI_PF_AHEAD 0
#endif
__device__ __forceinline__ void pf2(const float* p) {
#ifdef GPI_PF_SPARSE
    if ((threadIdx.x & 7) == 0)              // one request per 128-byte line (8 lanes x 16 B)
#endif
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
__device__ __forceinline__ float ld1(const float* p) {
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
#endif   // GPI_HOST_EMU
// value at z-1 of element 0 / at z+1 of element 3: one scalar load each, issued with all the other loads
// of the thread.  They hit the 128-byte lines the neighbouring lanes fetch anyway (no extra DRAM or L2
// traffic); warp shuffles would save the L1 request but make the first use of a loaded register precede
// the issue of the remaining loads (ptxas schedules SHFL early), which serialises two DRAM round trips.
__device__ __forceinline__ float z_prev(const float* p, int k0) { return k0 > 0 ? ld1(p - 1) : 0.f; }
__device__ __forceinline__ float z_next(const float* p, bool more) { return more ? ld1(p + VW) : 0.f; }

// d[e] * scale after a forward difference
__device__ __forceinline__ F4 diff4(const F4& hi, const F4& lo, float sI) {
    F4 r;
#pragma unroll
    for (int e = 0; e < VW; e++) r.v[e] = __fmul_rn(__fsub_rn(hi.v[e], lo.v[e]), sI);
    return r;
}
// forward difference along z onto half nodes: f(k) - f(k-1)
__device__ __forceinline__ F4 diff4_zm(const F4& c, float prev, float sI) {
    F4 r;
    r.v[0] = __fmul_rn(__fsub_rn(c.v[0], prev), sI);
#pragma unroll
    for (int e = 1; e < VW; e++) r.v[e] = __fmul_rn(__fsub_rn(c.v[e], c.v[e - 1]), sI);
    return r;
}
// forward difference along z from half nodes onto integer nodes: f(k+1) - f(k)
__device__ __forceinline__ F4 diff4_zp(const F4& c, float next, float sI) {
    F4 r;
#pragma unroll
    for (int e = 0; e < VW - 1; e++) r.v[e] = __fmul_rn(__fsub_rn(c.v[e + 1], c.v[e]), sI);
    r.v[VW - 1] = __fmul_rn(__fsub_rn(next, c.v[VW - 1]), sI);
    return r;
}

// CPML on four z-consecutive values, split in three phases so that every DRAM-latency load of a thread
// (fields, coefficients AND memory variables) is in flight before the first dependent instruction and
// before any store (stores would otherwise fence later loads: the compiler must assume aliasing):
//   pml_open_*  : slab test, address, issue the 16-byte load of the memory variables
//   pml_apply_* : m = b*m + a*d ; d = d*kI + m   (coefficients are L1-resident broadcast / table loads)
//   pml_close   : store the memory variables
struct Pml4 { float* mp; int s; F4 m; };

// Diagnostic builds (tuning only, results are WRONG on purpose): GPI_EXP_NOPML skips every CPML term,
// GPI_EXP_NONBR replaces the x / y neighbour loads by the centre line (L1 hits) -- they bound what the CPML
// traffic and the neighbour requests cost.  build.sh never defines them.
#ifdef GPI_EXP_NONBR
#define GPI_NBR(x) 0
#else
#define GPI_NBR(x) (x)
#endif
#ifndef GPI_EXP_NOPML_AXES
#define GPI_EXP_NOPML_AXES 0          // diagnostic: bit q set = skip the CPML terms of axis q (0 z, 1 y, 2 x)
#endif
template <int AXIS>   // 1 = y, 2 = x: slab index s uniform over the four cells (-1: not in a slab)
__device__ __forceinline__ void pml_open(Pml4& q, const Geom& g, const PmlTerm& t, int s, int k0, int j, int i, int b) {
    q.s = s; q.mp = nullptr;
#ifdef GPI_EXP_NOPML
    return;
#endif
    if ((GPI_EXP_NOPML_AXES >> AXIS) & 1) return;
    if (s < 0) return;
    long long mi;
    if (AXIS == 2) mi = (long long)k0 + (long long)g.pz * ((long long)j + (long long)g.ny1 * s);
    else           mi = (long long)k0 + (long long)g.pz * ((long long)s + 2LL * g.npml * i);
    q.mp = t.mem + (long long)b * t.bstride + mi;
#ifdef GPI_EXP_PMLNOMEM
    q.m = F4{};
#else
    q.m = ld4(q.mp);
#endif
}
// z terms: float4-aligned memory rows (cpml<> in kernels.cuh); s holds k0 for the table look-up
__device__ __forceinline__ void pml_open_z(Pml4& q, const Geom& g, const PmlTerm& t, int s0, int len, int k0l, int j, int i, int b) {
    const int npml = g.npml;
    const int k0 = k0l + g.koff;                 // global index of the first cell (koff % 4 == 0 keeps the alignment)
    int zi = -1;
#ifdef GPI_EXP_NOPML
    q.s = k0; q.mp = nullptr; return;
#endif
    if (GPI_EXP_NOPML_AXES & 1) { q.s = k0; q.mp = nullptr; return; }
    if ((g.pml & ZMIN) && k0 < s0 + npml) zi = k0;
    else if (g.pml & ZMAX) {
        const int kb = zslab_base(s0, len, npml);
        if (k0 >= kb && k0 < s0 + len) zi = (g.pzm >> 1) + (k0 - kb);
    }
    q.s = k0; q.mp = nullptr;
    if (zi < 0) return;
    q.mp = t.mem + (long long)b * t.bstride + (long long)zi + (long long)g.pzm * ((long long)j + (long long)g.ny1 * i);
#ifdef GPI_EXP_PMLNOMEM
    q.m = F4{};
#else
    q.m = ld4(q.mp);
#endif
}
__device__ __forceinline__ void pml_apply(Pml4& q, const PmlTerm& t, F4& d) {
#ifdef GPI_EXP_PMLNOARITH
    return;
#endif
    if (!q.mp) return;
    const float a = __ldg(t.a + q.s), bb = __ldg(t.b + q.s), kI = __ldg(t.kI + q.s);
#pragma unroll
    for (int e = 0; e < VW; e++) {
        q.m.v[e] = __fadd_rn(__fmul_rn(bb, q.m.v[e]), __fmul_rn(a, d.v[e]));
        d.v[e] = __fadd_rn(__fmul_rn(d.v[e], kI), q.m.v[e]);
    }
}
__device__ __forceinline__ void pml_apply_z(Pml4& q, const PmlTerm& t, F4& d) {
#ifdef GPI_EXP_PMLNOARITH
    return;
#endif
    if (!q.mp) return;
    const F4 a = ldg4(t.a + q.s), bb = ldg4(t.b + q.s), kI = ldg4(t.kI + q.s);
#pragma unroll
    for (int e = 0; e < VW; e++) {
        q.m.v[e] = __fadd_rn(__fmul_rn(bb.v[e], q.m.v[e]), __fmul_rn(a.v[e], d.v[e]));
        d.v[e] = __fadd_rn(__fmul_rn(d.v[e], kI.v[e]), q.m.v[e]);
    }
}
__device__ __forceinline__ void pml_close(const Pml4& q) {
#ifndef GPI_EXP_PMLNOMEM
    if (q.mp) st4(q.mp, q.m);
#endif
}

// thread -> (group of four z cells, row j); plane i = blockIdx.y, batch slot = blockIdx.z.
// GPI_VEC_ROWS = 1: the (z, y) plane is linearised group by group (a warp = 128 consecutive z cells of a row,
// straddling rows).  GPI_VEC_ROWS = R (4 or 8): a warp = R rows x (32/R) groups, i.e. one whole 128-byte (R = 4)
// or 64-byte (R = 8) segment per row.  Every access is still a full-line request, but a warp now sits in ONE z
// chunk: only the warps of the first / last chunks of a row contain z-slab cells, so the divergent z-CPML code is
// executed by ~4/11 of the warps instead of ~3/4 (and without divergence in the chunks that lie inside the slab).
#ifndef GPI_VEC_ROWS
#define GPI_VEC_ROWS 1
#endif
struct Vec3Idx { int k0, j, i, b; bool valid; };
__host__ __device__ inline int vec3_threads(int pz, int ny1) {
    const int nq = pz / VW;
    if (GPI_VEC_ROWS == 1) return nq * ny1;
    const int lpr = 32 / GPI_VEC_ROWS;                       // lanes (groups) per row inside a warp
    return (nq / lpr) * ((ny1 + GPI_VEC_ROWS - 1) / GPI_VEC_ROWS) * 32;
}
__device__ __forceinline__ Vec3Idx vec3_index(const Geom& g) {
    Vec3Idx q;
    const int nq = g.pz / VW;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (GPI_VEC_ROWS == 1) {
        q.valid = gid < nq * g.ny1;
        q.j = gid / nq;
        q.k0 = (gid - q.j * nq) * VW;
    } else {
        constexpr int R = GPI_VEC_ROWS, LPR = 32 / R;
        const int nzc = nq / LPR;                            // z chunks per row (pz is a multiple of 32 floats)
        const int w = gid >> 5, l = gid & 31;
        const int jb = w / nzc, zc = w - jb * nzc;
        q.j = jb * R + l / LPR;
        q.k0 = (zc * LPR + (l % LPR)) * VW;
        q.valid = q.j < g.ny1;
    }
    q.i = blockIdx.y + g.ioff; q.b = blockIdx.z;
    return q;
}

// ------------------------------------------------------------------------------------------------
// velocity kernel, 3-D (see vel_cell for the reference citations and the CPML term order)
// ------------------------------------------------------------------------------------------------
template <int EL, int OOP = 0>      // OOP: out of place, reads a.v and writes a.v_o (ping-pong adjoint runs)
__global__ void __launch_bounds__(GPI_VEC_THREADS, GPI_VEC_MINBLOCKS) k_vel3v(const Geom g, const StepArgs a) {
    const Vec3Idx q = vec3_index(g);
    const int k0 = q.k0, j = q.j, i = q.i, b = q.b;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const bool fast = q.valid && i >= 2 && i <= nx - 2 && j >= 2 && j <= ny - 2;
    if (!q.valid) return;
    if (!fast) {
#pragma unroll 1
        for (int e = 0; e < VW; e++) if (k0 + e >= g.klo && k0 + e <= g.khi) vel_cell<3, EL, OOP>(g, a, k0 + e, j, i, b);
        return;
    }
    const int kg0 = k0 + g.koff;                 // global z index of element 0
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, k0, j, i) + w;
    const long long sy = GPI_NBR(g.pz), sx = GPI_NBR((long long)g.pz * g.ny1);
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;

    float* vx = a.v[V_X] + c; float* vy = a.v[V_Y] + c; float* vz = a.v[V_Z] + c;
    float* ovx = OOP ? a.v_o[V_X] + c : vx; float* ovy = OOP ? a.v_o[V_Y] + c : vy; float* ovz = OOP ? a.v_o[V_Z] + c : vz;     // written
    pf(vx); pf(vy); pf(vz); pf(a.c[C_BX] + c - w); pf(a.c[C_BY] + c - w); pf(a.c[C_BZ] + c - w);
    if (GPI_PF_AHEAD > 0 && i + GPI_PF_AHEAD <= nx - 2) {
        const long long o = GPI_PF_AHEAD * sx;
        pf2(vx + o); pf2(vy + o); pf2(vz + o);
        pf2(a.c[C_BX] + c - w + o); pf2(a.c[C_BY] + c - w + o); pf2(a.c[C_BZ] + c - w + o);
        pf2(a.tau[T_XX] + c + o);
        if (EL) { pf2(a.tau[T_YY] + c + o); pf2(a.tau[T_ZZ] + c + o); pf2(a.tau[T_XY] + c + o + sx); pf2(a.tau[T_XZ] + c + o + sx); pf2(a.tau[T_YZ] + c + o); }
    }
    F4 nvx = ld4(vx), nvy = ld4(vy), nvz = ld4(vz);

    if (!EL) {
        const float* p = a.tau[T_XX] + c;
        const F4 pc = ld4(p), pmx = ld4(p - sx), pmy = ld4(p - sy);
        const F4 bx = ldg4(a.c[C_BX] + c - w), by = ldg4(a.c[C_BY] + c - w), bz = ldg4(a.c[C_BZ] + c - w);
        Pml4 m0, m1, m2;
        pml_open<2>(m0, g, a.pv[0], slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax), k0, j, i, b);
        pml_open<1>(m1, g, a.pv[1], slab_index(j, 1, ny - 1, g.npml, hymin, hymax), k0, j, i, b);
        pml_open_z(m2, g, a.pv[2], 1, nz - 1, k0, j, i, b);
        const float pprev = z_prev(p, k0);
        F4 dx = diff4(pc, pmx, g.dxI);        pml_apply(m0, a.pv[0], dx);
        F4 dy = diff4(pc, pmy, g.dyI);        pml_apply(m1, a.pv[1], dy);
        F4 dz = diff4_zm(pc, pprev, g.dzI);   pml_apply_z(m2, a.pv[2], dz);
        pml_close(m0); pml_close(m1); pml_close(m2);
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k >= 1 && k <= nz - 2) {
                nvx.v[e] = __fadd_rn(nvx.v[e], __fmul_rn(bx.v[e], dx.v[e]));
                nvy.v[e] = __fadd_rn(nvy.v[e], __fmul_rn(by.v[e], dy.v[e]));
            }
            if (own && k >= 1 && k <= nz - 1) nvz.v[e] = __fadd_rn(nvz.v[e], __fmul_rn(bz.v[e], dz.v[e]));
        }
    } else {
        const float* txx = a.tau[T_XX] + c; const float* tyy = a.tau[T_YY] + c; const float* tzz = a.tau[T_ZZ] + c;
        const float* txy = a.tau[T_XY] + c; const float* txz = a.tau[T_XZ] + c; const float* tyz = a.tau[T_YZ] + c;
        const F4 xx = ld4(txx), xxm = ld4(txx - sx);
        const F4 yy = ld4(tyy), yym = ld4(tyy - sy);
        const F4 zz = ld4(tzz);
        const F4 xy = ld4(txy), xypy = ld4(txy + sy), xypx = ld4(txy + sx);
        const F4 xz = ld4(txz), xzpx = ld4(txz + sx);
        const F4 yz = ld4(tyz), yzpy = ld4(tyz + sy);
        const F4 bx = ldg4(a.c[C_BX] + c - w), by = ldg4(a.c[C_BY] + c - w), bz = ldg4(a.c[C_BZ] + c - w);
        const int sx0 = slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax);      // dtauxxdx
        const int sx1 = slab_index(i, 1, nx - 2, g.npml, hxmin, hxmax);      // dtauxydx, dtauxzdx
        const int sy0 = slab_index(j, 1, ny - 2, g.npml, hymin, hymax);      // dtauxydy, dtauyzdy
        const int sy1 = slab_index(j, 1, ny - 1, g.npml, hymin, hymax);      // dtauyydy
        Pml4 m0, m1, m2, m3, m4, m5, m6, m7, m8;
        pml_open<2>(m0, g, a.pv[0], sx0, k0, j, i, b);   pml_open<1>(m1, g, a.pv[1], sy0, k0, j, i, b);   pml_open_z(m2, g, a.pv[2], 1, nz - 2, k0, j, i, b);
        pml_open<2>(m3, g, a.pv[3], sx1, k0, j, i, b);   pml_open<1>(m4, g, a.pv[4], sy1, k0, j, i, b);   pml_open_z(m5, g, a.pv[5], 1, nz - 2, k0, j, i, b);
        pml_open<2>(m6, g, a.pv[6], sx1, k0, j, i, b);   pml_open<1>(m7, g, a.pv[7], sy0, k0, j, i, b);   pml_open_z(m8, g, a.pv[8], 1, nz - 1, k0, j, i, b);
        const float zzprev = z_prev(tzz, k0);
        const float xznext = z_next(txz, more);
        const float yznext = z_next(tyz, more);

        // vx: dtauxxdx + dtauxydy + dtauxzdz
        F4 dxx = diff4(xx, xxm, g.dxI);            pml_apply(m0, a.pv[0], dxx);
        F4 dxy = diff4(xypy, xy, g.dyI);           pml_apply(m1, a.pv[1], dxy);
        F4 dxz = diff4_zp(xz, xznext, g.dzI);      pml_apply_z(m2, a.pv[2], dxz);
        // vy: dtauxydx + dtauyydy + dtauyzdz
        F4 dyx = diff4(xypx, xy, g.dxI);           pml_apply(m3, a.pv[3], dyx);
        F4 dyy = diff4(yy, yym, g.dyI);            pml_apply(m4, a.pv[4], dyy);
        F4 dyz = diff4_zp(yz, yznext, g.dzI);      pml_apply_z(m5, a.pv[5], dyz);
        // vz: dtauxzdx + dtauyzdy + dtauzzdz
        F4 dzx = diff4(xzpx, xz, g.dxI);           pml_apply(m6, a.pv[6], dzx);
        F4 dzy = diff4(yzpy, yz, g.dyI);           pml_apply(m7, a.pv[7], dzy);
        F4 dzz = diff4_zm(zz, zzprev, g.dzI);      pml_apply_z(m8, a.pv[8], dzz);
        pml_close(m0); pml_close(m1); pml_close(m2); pml_close(m3); pml_close(m4); pml_close(m5); pml_close(m6); pml_close(m7); pml_close(m8);
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k >= 1 && k <= nz - 2) {
                nvx.v[e] = __fsub_rn(nvx.v[e], __fmul_rn(bx.v[e], __fadd_rn(__fadd_rn(dxx.v[e], dxy.v[e]), dxz.v[e])));
                nvy.v[e] = __fsub_rn(nvy.v[e], __fmul_rn(by.v[e], __fadd_rn(__fadd_rn(dyx.v[e], dyy.v[e]), dyz.v[e])));
            }
            if (own && k >= 1 && k <= nz - 1)
                nvz.v[e] = __fsub_rn(nvz.v[e], __fmul_rn(bz.v[e], __fadd_rn(__fadd_rn(dzx.v[e], dzy.v[e]), dzz.v[e])));
        }
    }

    // rigid z faces (dirichlet.jl:35-74); the x / y faces only touch shell rows (scalar path)
    const int R = g.rigid;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    if (head && (R & ZMIN)) { nvx.v[0] = 0.f; nvy.v[0] = 0.f; nvz.v[0] = -nvz.v[1]; }
    if (!tail) {
        st4(ovx, nvx); st4(ovy, nvy); st4(ovz, nvz);
    } else {
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k <= nz - 1) {
                const bool zero = (R & ZMAX) && k == nz - 1;
                ovx[e] = zero ? 0.f : nvx.v[e];
                ovy[e] = zero ? 0.f : nvy.v[e];
                ovz[e] = nvz.v[e];
                if ((R & ZMAX) && k == nz - 1) ovz[e + 1] = -nvz.v[e];      // vz[nz+1] = -vz[nz]
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stress kernel, 3-D (see stress_cell for the reference citations and the CPML term order)
// ------------------------------------------------------------------------------------------------
template <int EL, int OOP = 0>      // OOP: out of place, reads a.tau and writes a.tau_o
__global__ void __launch_bounds__(GPI_VEC_THREADS, GPI_VEC_MINBLOCKS) k_stress3v(const Geom g, const StepArgs a) {
    const Vec3Idx q = vec3_index(g);
    const int k0 = q.k0, j = q.j, i = q.i, b = q.b;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const bool fast = q.valid && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
    if (!q.valid) return;
    if (!fast) {
#pragma unroll 1
        for (int e = 0; e < VW; e++) if (k0 + e >= g.klo && k0 + e <= g.khi) stress_cell<3, EL, OOP>(g, a, k0 + e, j, i, b);
        return;
    }
    const int kg0 = k0 + g.koff;                 // global z index of element 0
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, k0, j, i) + w;
    const long long sy = GPI_NBR(g.pz), sx = GPI_NBR((long long)g.pz * g.ny1);
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;

    // ---- phase 1: every load of the thread ---------------------------------------------------------
    const float* vx = a.v[V_X] + c; const float* vy = a.v[V_Y] + c; const float* vz = a.v[V_Z] + c;
    if (GPI_PF_AHEAD > 0 && i + GPI_PF_AHEAD <= nx - 2) {
        const long long o = GPI_PF_AHEAD * sx;
        pf2(vx + o + sx); pf2(vy + o); pf2(vz + o);
        pf2(a.tau[T_XX] + c + o); pf2(a.c[C_K] + c - w + o);
        if (EL) {
            pf2(a.tau[T_YY] + c + o); pf2(a.tau[T_ZZ] + c + o); pf2(a.tau[T_XY] + c + o); pf2(a.tau[T_XZ] + c + o); pf2(a.tau[T_YZ] + c + o);
            pf2(a.c[C_L] + c - w + o); pf2(a.c[C_MUXZ] + c - w + o); pf2(a.c[C_MUXY] + c - w + o); pf2(a.c[C_MUYZ] + c - w + o);
        }
    }
    const F4 cvx = ld4(vx), cvy = ld4(vy), cvz = ld4(vz);
    const F4 vxpx = ld4(vx + sx), vypy = ld4(vy + sy);
    Pml4 m0, m1, m2;
    pml_open<2>(m0, g, a.ps[0], slab_index(i, 0, nx, g.npml, hxmin, hxmax), k0, j, i, b);
    pml_open<1>(m1, g, a.ps[1], slab_index(j, 0, ny, g.npml, hymin, hymax), k0, j, i, b);
    pml_open_z(m2, g, a.ps[2], 0, nz, k0, j, i, b);

    if (!EL) {
        float* p = a.tau[T_XX] + c;
        F4 pc = ld4(p);
        const F4 K = ldg4(a.c[C_K] + c - w);
        const float vznext = z_next(vz, more);
        F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_apply(m0, a.ps[0], dxx);       // @d_xa(vx)
        F4 dyy = diff4(vypy, cvy, g.dyI);           pml_apply(m1, a.ps[1], dyy);       // @d_ya(vy)
        F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_apply_z(m2, a.ps[2], dzz);     // @d_za(vz)
#pragma unroll
        for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo))
            pc.v[e] = __fadd_rn(pc.v[e], __fmul_rn(__fadd_rn(__fadd_rn(dxx.v[e], dzz.v[e]), dyy.v[e]), K.v[e]));
        st4(OOP ? a.tau_o[T_XX] + c : p, pc);
        pml_close(m0); pml_close(m1); pml_close(m2);
        return;
    }

    const bool fs = (g.freesurf & ZMIN) != 0;
    float* txx = a.tau[T_XX] + c; float* tyy = a.tau[T_YY] + c; float* tzz = a.tau[T_ZZ] + c;
    float* txy = a.tau[T_XY] + c; float* txz = a.tau[T_XZ] + c; float* tyz = a.tau[T_YZ] + c;
    pf(txx); pf(tyy); pf(tzz); pf(txy); pf(txz); pf(tyz);
    pf(a.c[C_K] + c - w); pf(a.c[C_L] + c - w); pf(a.c[C_MUXZ] + c - w); pf(a.c[C_MUXY] + c - w); pf(a.c[C_MUYZ] + c - w);
    F4 xx = ld4(txx), yy = ld4(tyy), zz = ld4(tzz), xy = ld4(txy), xz = ld4(txz), yz = ld4(tyz);
    const F4 M = ldg4(a.c[C_K] + c - w), L = ldg4(a.c[C_L] + c - w);
    const F4 muxz = ldg4(a.c[C_MUXZ] + c - w), muxy = ldg4(a.c[C_MUXY] + c - w), muyz = ldg4(a.c[C_MUYZ] + c - w);
    const F4 vxmy = ld4(vx - sy), vymx = ld4(vy - sx), vzmx = ld4(vz - sx), vzmy = ld4(vz - sy);
    const int sxh = slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax);      // dvzdx, dvydx
    const int syh = slab_index(j, 1, ny - 1, g.npml, hymin, hymax);      // dvxdy, dvzdy
    Pml4 m3, m4, m5, m6, m7, m8;
    pml_open<1>(m3, g, a.ps[3], syh, k0, j, i, b);   pml_open<2>(m4, g, a.ps[4], sxh, k0, j, i, b);
    pml_open_z(m5, g, a.ps[5], 1, nz - 1, k0, j, i, b);   pml_open<2>(m6, g, a.ps[6], sxh, k0, j, i, b);
    pml_open_z(m7, g, a.ps[7], 1, nz - 1, k0, j, i, b);   pml_open<1>(m8, g, a.ps[8], syh, k0, j, i, b);
    const float vznext = z_next(vz, more);
    const float vxprev = z_prev(vx, k0);
    const float vyprev = z_prev(vy, k0);

    // ---- phase 2: arithmetic in the reference's order ------------------------------------------------
    F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_apply(m0, a.ps[0], dxx);       // @d_xa(vx)
    F4 dyy = diff4(vypy, cvy, g.dyI);           pml_apply(m1, a.ps[1], dyy);       // @d_ya(vy)
    F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_apply_z(m2, a.ps[2], dzz);     // @d_za(vz)
#pragma unroll
    for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo)) {
        xx.v[e] = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dzz.v[e])));
        yy.v[e] = __fsub_rn(__fsub_rn(yy.v[e], __fmul_rn(M.v[e], dyy.v[e])), __fmul_rn(L.v[e], __fadd_rn(dxx.v[e], dzz.v[e])));
        zz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dxx.v[e])));
    }
    if (fs && kg0 == 0) zz.v[0] = -zz.v[1];                     // free_surface_mirror!: tauzz[1] = -tauzz[2]

    // tauxz: z half, y inner, x half
    F4 dxz = diff4_zm(cvx, vxprev, g.dzI);      pml_apply_z(m5, a.ps[5], dxz);     // @d_zi(vx)
    F4 dzx = diff4(cvz, vzmx, g.dxI);           pml_apply(m6, a.ps[6], dzx);       // @d_xi(vz)
    // tauxy: z inner, y half, x half
    F4 dxy = diff4(cvx, vxmy, g.dyI);           pml_apply(m3, a.ps[3], dxy);       // @d_yi(vx)
    F4 dyx = diff4(cvy, vymx, g.dxI);           pml_apply(m4, a.ps[4], dyx);       // @d_xi(vy)
    // tauyz: z half, y half, x inner
    F4 dyz = diff4_zm(cvy, vyprev, g.dzI);      pml_apply_z(m7, a.ps[7], dyz);     // @d_zi(vy)
    F4 dzy = diff4(cvz, vzmy, g.dyI);           pml_apply(m8, a.ps[8], dzy);       // @d_yi(vz)
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
        if (own && k >= 1 && k <= nz - 1) {
            float n = __fsub_rn(xz.v[e], __fmul_rn(muxz.v[e], __fadd_rn(dxz.v[e], dzx.v[e])));
            if (fs && k == 1) n = 0.f;                             // free_surface!(tauxz)
            xz.v[e] = n;
            n = __fsub_rn(yz.v[e], __fmul_rn(muyz.v[e], __fadd_rn(dyz.v[e], dzy.v[e])));
            if (fs && k == 1) n = 0.f;                             // free_surface!(tauyz)
            yz.v[e] = n;
        }
        if (own && k >= 1 && k <= nz - 2) xy.v[e] = __fsub_rn(xy.v[e], __fmul_rn(muxy.v[e], __fadd_rn(dxy.v[e], dyx.v[e])));
    }

    // ---- phase 3: stores -----------------------------------------------------------------------------
    if (OOP) {
        st4(a.tau_o[T_XX] + c, xx); st4(a.tau_o[T_YY] + c, yy); st4(a.tau_o[T_ZZ] + c, zz);
        st4(a.tau_o[T_XZ] + c, xz); st4(a.tau_o[T_XY] + c, xy); st4(a.tau_o[T_YZ] + c, yz);
    } else {
        st4(txx, xx); st4(tyy, yy); st4(tzz, zz); st4(txz, xz); st4(txy, xy); st4(tyz, yz);
    }
    pml_close(m0); pml_close(m1); pml_close(m2); pml_close(m3); pml_close(m4); pml_close(m5); pml_close(m6); pml_close(m7); pml_close(m8);
}
