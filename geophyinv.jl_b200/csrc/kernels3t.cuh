// kernels3t.cuh -- TMA-pipelined 3-D elastic stencil kernels (included inside namespace gpi, after kernels3d.cuh).
//
// Same arithmetic, same order of operations and the same results (bit for bit) as k_vel3v / k_stress3v and the
// scalar vel_cell / stress_cell; what changes is how the operands reach the SM.
//
// k_vel3v / k_stress3v are bound by the time a warp spends NOT waiting for memory: ncu (profiles/r01) shows ~830
// warp instructions per 128 cells, two thirds of them integer address / predicate work, at 10 resident warps per
// SM (168 registers), so only 4-5 warps per SM have loads in flight at any time.  Here the loads are decoupled
// from the warps that compute:
//
//   * persistent CTAs (2 per SM), each = 4 consumer warps + 1 producer warp, walk a static round-robin list of
//     tiles; a tile = R = 4 rows (y) x ZC = 128 cells (z) of one x plane, ordered z-chunk, row block, plane so
//     that the tiles in flight at any moment cover ~1.2 consecutive planes (+-1-plane operands are L2 hits);
//   * one elected producer thread moves every operand of a tile -- 15 (velocity) or 17 (stress) boxes of 4 or 5
//     rows x 544 bytes: the 15 / 20 arrays at the tile's own rows with their y +-1 halo row, the x +-1 plane
//     rows, each with a 16-byte halo on both z sides -- with TMA tensor copies (cp.async.bulk.tensor.3d global
//     -> shared through one CUtensorMap per box, completion on an mbarrier) into a 2-stage shared-memory ring
//     (per-row cp.async.bulk copies were tried first: ptxas serialises them lane by lane, ~85 cycles per row,
//     which made the producer the bottleneck); the loads of tile n+1 are in flight while tile n is computed, with no
//     registers and no scoreboard entries tied to them (2 CTAs x 34-38 KB per SM continuously in flight);
//   * consumers read operands from shared memory (conflict-free 16-byte reads; the z +-1 neighbours are plain
//     4-byte shared reads), do the reference-order arithmetic, and write results with coalesced 16-byte global
//     stores; a full / empty mbarrier pair per stage is the only synchronisation (no __syncthreads in the loop);
//   * CPML memory variables are read-modify-write streams private to a cell: consumers load them straight from
//     global memory BEFORE waiting for the tile (the producer has requested those lines into L2 one to two
//     tiles earlier with cp.async.bulk.prefetch.L2), and store them back after the update;
//   * the outer shell of the box (rigid faces, ghost cells, ragged x / y ranges; 2.4 % of the rows) is done by the
//     consumer warps with the scalar reference-order code after their last tile, which fills the tail of the
//     persistent schedule.
//
// The z-slab window (Geom.koff / klo / khi) is honoured exactly as in kernels3d.cuh, so the slab decomposition
// uses the same kernels.

namespace t3 {

#ifndef T3_STAGES
#define T3_STAGES 2
#endif
#ifndef T3_MINB
#define T3_MINB 2                    // resident CTAs per SM the register budget is sized for
#endif
constexpr int R = 4;                 // rows per tile
constexpr int ZC = 128;              // z cells per tile
constexpr int PITCH = ZC + 8;        // floats per staged row (16-byte halo on both sides)
constexpr int ROWBYTES = PITCH * 4;
constexpr int STAGES = T3_STAGES;
constexpr int NCW = R;               // consumer warps (one per row)
constexpr int NTHREADS = 32 * (NCW + 1);

// operand boxes of a stage.  Every box is rows x PITCH floats, dense, and starts on a 128-byte boundary (a 4-row box
// is 2176 B = 17 x 128; a 5-row box 2720 B is padded to 2816 B).  Offsets below are in FLOATS from the stage base.
constexpr int B4 = 4 * PITCH, B5 = 5 * PITCH + 24;
// velocity kernel: box id -> offset
enum { V_XX0 = 0, V_XXM = V_XX0 + B4, V_YY = V_XXM + B4 /*5 rows j-1..j+3*/, V_ZZ = V_YY + B5, V_XY0 = V_ZZ + B4 /*5 rows j..j+4*/,
       V_XYP = V_XY0 + B5, V_XZ0 = V_XYP + B4, V_XZP = V_XZ0 + B4, V_YZ = V_XZP + B4 /*5 rows*/, V_VX = V_YZ + B5, V_VY = V_VX + B4,
       V_VZ = V_VY + B4, V_BX = V_VZ + B4, V_BY = V_BX + B4, V_BZ = V_BY + B4, V_FLOATS = V_BZ + B4, V_NBOX = 15 };
// stress kernel
enum { S_VX0 = 0 /*5 rows j-1..j+3*/, S_VXP = S_VX0 + B5, S_VY0 = S_VXP + B4 /*5 rows j..j+4*/, S_VYM = S_VY0 + B5,
       S_VZ0 = S_VYM + B4 /*5 rows j-1..j+3*/, S_VZM = S_VZ0 + B5, S_XX = S_VZM + B4, S_YY = S_XX + B4, S_ZZ = S_YY + B4,
       S_XY = S_ZZ + B4, S_XZ = S_XY + B4, S_YZ = S_XZ + B4, S_K = S_YZ + B4, S_L = S_K + B4, S_MUXZ = S_L + B4,
       S_MUXY = S_MUXZ + B4, S_MUYZ = S_MUXY + B4, S_FLOATS = S_MUYZ + B4, S_NBOX = 17 };
constexpr int MAXBOX = 17;

// one operand box: which array, plane / first-row offset relative to (i, j0), rows, offset inside the stage
struct BoxSpec { int arr; int di, dj, rows, off; };      // arr: 0-5 tau, 6-8 v, 9.. coefficient slot + 9
__host__ __device__ inline BoxSpec box_spec(int kind, int b) {
    if (kind == 0) {
        const BoxSpec t[V_NBOX] = {
            {T_XX, 0, 0, 4, V_XX0}, {T_XX, -1, 0, 4, V_XXM}, {T_YY, 0, -1, 5, V_YY}, {T_ZZ, 0, 0, 4, V_ZZ},
            {T_XY, 0, 0, 5, V_XY0}, {T_XY, 1, 0, 4, V_XYP}, {T_XZ, 0, 0, 4, V_XZ0}, {T_XZ, 1, 0, 4, V_XZP},
            {T_YZ, 0, 0, 5, V_YZ}, {6 + V_X, 0, 0, 4, V_VX}, {6 + V_Y, 0, 0, 4, V_VY}, {6 + V_Z, 0, 0, 4, V_VZ},
            {9 + C_BX, 0, 0, 4, V_BX}, {9 + C_BY, 0, 0, 4, V_BY}, {9 + C_BZ, 0, 0, 4, V_BZ}};
        return t[b];
    }
    const BoxSpec t[S_NBOX] = {
        {6 + V_X, 0, -1, 5, S_VX0}, {6 + V_X, 1, 0, 4, S_VXP}, {6 + V_Y, 0, 0, 5, S_VY0}, {6 + V_Y, -1, 0, 4, S_VYM},
        {6 + V_Z, 0, -1, 5, S_VZ0}, {6 + V_Z, -1, 0, 4, S_VZM}, {T_XX, 0, 0, 4, S_XX}, {T_YY, 0, 0, 4, S_YY},
        {T_ZZ, 0, 0, 4, S_ZZ}, {T_XY, 0, 0, 4, S_XY}, {T_XZ, 0, 0, 4, S_XZ}, {T_YZ, 0, 0, 4, S_YZ},
        {9 + C_K, 0, 0, 4, S_K}, {9 + C_L, 0, 0, 4, S_L}, {9 + C_MUXZ, 0, 0, 4, S_MUXZ}, {9 + C_MUXY, 0, 0, 4, S_MUXY},
        {9 + C_MUYZ, 0, 0, 4, S_MUYZ}};
    return t[b];
}
// the TMA descriptors of one kernel (one per box), passed as a __grid_constant__ kernel parameter
struct alignas(64) Maps { unsigned char m[MAXBOX][128]; };

struct Sched {
    int ilo, ihi, jlo, jhi;          // fast region (inclusive)
    int njb, nzc;                    // row blocks per plane, z chunks per row
    int ntiles;
    int sp[4], nsp;                  // shell planes (all rows)
    int sr[4], nsr;                  // shell rows of the fast planes
};

__host__ __device__ inline size_t smem_bytes(int stage_floats) { return (size_t)STAGES * stage_floats * 4 + 2 * STAGES * 8 + 128; }

// ---- PTX: mbarrier + bulk copies ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const float* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ F4 lds4(const float* p) { return *reinterpret_cast<const F4*>(p); }

// tile index -> plane, first row, first z cell
__device__ __forceinline__ void tile_coords(const Sched& sc, int t, int& i, int& j0, int& kc0) {
    const int per_plane = sc.njb * sc.nzc;
    const int ip = t / per_plane, rem = t - ip * per_plane;
    const int jb = rem / sc.nzc, zc = rem - jb * sc.nzc;
    i = sc.ilo + ip; j0 = sc.jlo + jb * R; kc0 = zc * ZC;
}

// CPML terms of a kernel: axis (0 z, 1 y, 2 x) and the extent [s0, s0+len) of the derivative field along it
struct TermSpec { int axis, s0, len; };
template <int KIND>
__device__ __forceinline__ TermSpec term_spec(const Geom& g, int t) {
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    if (KIND == 0) {
        switch (t) {
            case 0: return {2, 1, nx - 1}; case 1: return {1, 1, ny - 2}; case 2: return {0, 1, nz - 2};
            case 3: return {2, 1, nx - 2}; case 4: return {1, 1, ny - 1}; case 5: return {0, 1, nz - 2};
            case 6: return {2, 1, nx - 2}; case 7: return {1, 1, ny - 2}; default: return {0, 1, nz - 1};
        }
    } else {
        switch (t) {
            case 0: return {2, 0, nx}; case 1: return {1, 0, ny}; case 2: return {0, 0, nz};
            case 3: return {1, 1, ny - 1}; case 4: return {2, 1, nx - 1}; case 5: return {0, 1, nz - 1};
            case 6: return {2, 1, nx - 1}; case 7: return {0, 1, nz - 1}; default: return {1, 1, ny - 1};
        }
    }
}

// producer lanes 0..8, one CPML term each: request the memory-variable lines of a tile into L2 (the consumers
// read them from global memory one to two tiles later)
template <int KIND>
__device__ __forceinline__ void prefetch_pml(const Geom& g, const StepArgs& a, int lane, int i, int j0, int kc0) {
    if (lane >= 9) return;
    const PmlTerm& t = (KIND == 0 ? a.pv : a.ps)[lane];
    const TermSpec ts = term_spec<KIND>(g, lane);
    const int npml = g.npml;
    const int zlen = min(ZC, g.pz - kc0);
    if (ts.axis == 2) {
        const int s = slab_index(i, ts.s0, ts.len, npml, g.pml & XMIN, g.pml & XMAX);
        if (s < 0) return;
#pragma unroll
        for (int r = 0; r < R; r++)
            l2_prefetch(t.mem + (long long)kc0 + (long long)g.pz * ((long long)(j0 + r) + (long long)g.ny1 * s), zlen * 4);
    } else if (ts.axis == 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int s = slab_index(j0 + r, ts.s0, ts.len, npml, g.pml & YMIN, g.pml & YMAX);
            if (s >= 0) l2_prefetch(t.mem + (long long)kc0 + (long long)g.pz * ((long long)s + 2LL * npml * i), zlen * 4);
        }
    } else {
        // z rows of R consecutive j are contiguous: [zi (pzm), j, i]
        const int kg0 = kc0 + g.koff, kg1 = kg0 + zlen;              // global z range of the chunk
        const bool lo = (g.pml & ZMIN) && kg0 < ts.s0 + npml;
        const bool hi = (g.pml & ZMAX) && kg1 > zslab_base(ts.s0, ts.len, npml);
        if (lo || hi) l2_prefetch(t.mem + (long long)g.pzm * ((long long)j0 + (long long)g.ny1 * i), R * g.pzm * 4);
    }
}

// shell lines after the last tile: scalar reference-order code (vel_cell / stress_cell)
template <int KIND>
__device__ __forceinline__ void shell_work(const Geom& g, const StepArgs& a, const Sched& sc, int cw, int lane) {
    const int nfast = sc.ihi - sc.ilo + 1;
    const int nlines = sc.nsp * g.ny1 + nfast * sc.nsr;
    const int nwork = nlines * sc.nzc;
    for (int wi = blockIdx.x * NCW + cw; wi < nwork; wi += gridDim.x * NCW) {
        const int L = wi / sc.nzc, zc = wi - L * sc.nzc;
        int i, j;
        if (L < sc.nsp * g.ny1) { const int q = L / g.ny1; i = sc.sp[q]; j = L - q * g.ny1; }
        else { const int L2 = L - sc.nsp * g.ny1; const int q = L2 / sc.nsr; i = sc.ilo + q; j = sc.sr[L2 - q * sc.nsr]; }
        const int k0 = zc * ZC + 4 * lane;
        if (k0 >= g.pz) continue;
#pragma unroll 1
        for (int e = 0; e < 4; e++) if (k0 + e >= g.klo && k0 + e <= g.khi) {
            if (KIND == 0) vel_cell<3, 1>(g, a, k0 + e, j, i, 0);
            else           stress_cell<3, 1>(g, a, k0 + e, j, i, 0);
        }
    }
}

template <int KIND>
__device__ __forceinline__ void producer(const Geom& g, const StepArgs& a, const Sched& sc, const Maps& tm, float* stage0,
                                         uint32_t full0, uint32_t empty0, int lane) {
    constexpr int NBOX = KIND == 0 ? V_NBOX : S_NBOX;
    constexpr int SFLOATS = KIND == 0 ? V_FLOATS : S_FLOATS;
    int txbytes = 0;
#pragma unroll
    for (int b = 0; b < NBOX; b++) txbytes += box_spec(KIND, b).rows * ROWBYTES;
    int n = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++) {
        const int s = n % STAGES;
        const uint32_t use = n / STAGES;
        mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
        int i, j0, kc0;
        tile_coords(sc, t, i, j0, kc0);
        if (lane == 0) {
            const uint32_t bar = full0 + 8 * s;
            mbar_expect_tx(bar, txbytes);
            const uint32_t dst0 = s32(stage0 + (size_t)s * SFLOATS);
#pragma unroll
            for (int b = 0; b < NBOX; b++) {
                const BoxSpec bs = box_spec(KIND, b);
                tma_box(dst0 + bs.off * 4, tm.m[b], kc0 - 4, j0 + bs.dj, i + bs.di, bar);
            }
        }
        __syncwarp();
        prefetch_pml<KIND>(g, a, lane, i, j0, kc0);
    }
}

// ------------------------------------------------------------------------------------------------
// velocity kernel (term order and citations: vel_cell in kernels.cuh)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vel_tile(const Geom& g, const StepArgs& a, const float* S, int lane, int r,
                                         int i, int j, int k0, uint32_t fullbar, uint32_t parity, bool active) {
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int kg0 = k0 + g.koff;
    const long long c = uidx(g, k0, j, i);
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;
    Pml4 m0, m1, m2, m3, m4, m5, m6, m7, m8;
    if (active) {
        const int sx0 = slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax);      // dtauxxdx
        const int sx1 = slab_index(i, 1, nx - 2, g.npml, hxmin, hxmax);      // dtauxydx, dtauxzdx
        const int sy0 = slab_index(j, 1, ny - 2, g.npml, hymin, hymax);      // dtauxydy, dtauyzdy
        const int sy1 = slab_index(j, 1, ny - 1, g.npml, hymin, hymax);      // dtauyydy
        pml_open<2>(m0, g, a.pv[0], sx0, k0, j, i, 0);   pml_open<1>(m1, g, a.pv[1], sy0, k0, j, i, 0);   pml_open_z(m2, g, a.pv[2], 1, nz - 2, k0, j, i, 0);
        pml_open<2>(m3, g, a.pv[3], sx1, k0, j, i, 0);   pml_open<1>(m4, g, a.pv[4], sy1, k0, j, i, 0);   pml_open_z(m5, g, a.pv[5], 1, nz - 2, k0, j, i, 0);
        pml_open<2>(m6, g, a.pv[6], sx1, k0, j, i, 0);   pml_open<1>(m7, g, a.pv[7], sy0, k0, j, i, 0);   pml_open_z(m8, g, a.pv[8], 1, nz - 1, k0, j, i, 0);
    }
    mbar_wait(fullbar, parity);
    if (!active) return;

    const float* P = S + 4 * lane + 4;                    // this thread's four cells inside a staged row
    auto row = [&](int box, int rr) { return P + box + rr * PITCH; };
    const F4 xx = lds4(row(V_XX0, r)), xxm = lds4(row(V_XXM, r));
    const F4 yy = lds4(row(V_YY, r + 1)), yym = lds4(row(V_YY, r));
    const F4 zz = lds4(row(V_ZZ, r));
    const F4 xy = lds4(row(V_XY0, r)), xypy = lds4(row(V_XY0, r + 1)), xypx = lds4(row(V_XYP, r));
    const F4 xz = lds4(row(V_XZ0, r)), xzpx = lds4(row(V_XZP, r));
    const F4 yz = lds4(row(V_YZ, r)), yzpy = lds4(row(V_YZ, r + 1));
    const float zzprev = k0 > 0 ? row(V_ZZ, r)[-1] : 0.f;
    const float xznext = more ? row(V_XZ0, r)[VW] : 0.f;
    const float yznext = more ? row(V_YZ, r)[VW] : 0.f;

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

    F4 nvx = lds4(row(V_VX, r)), nvy = lds4(row(V_VY, r)), nvz = lds4(row(V_VZ, r));
    const F4 bx = lds4(row(V_BX, r)), by = lds4(row(V_BY, r)), bz = lds4(row(V_BZ, r));
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
    float* vx = a.v[V_X] + c; float* vy = a.v[V_Y] + c; float* vz = a.v[V_Z] + c;
    // rigid z faces (dirichlet.jl:35-74); the x / y faces only touch shell rows (scalar path)
    const int Rg = g.rigid;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    if (head && (Rg & ZMIN)) { nvx.v[0] = 0.f; nvy.v[0] = 0.f; nvz.v[0] = -nvz.v[1]; }
    if (!tail) {
        st4(vx, nvx); st4(vy, nvy); st4(vz, nvz);
    } else {
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k <= nz - 1) {
                const bool zero = (Rg & ZMAX) && k == nz - 1;
                vx[e] = zero ? 0.f : nvx.v[e];
                vy[e] = zero ? 0.f : nvy.v[e];
                vz[e] = nvz.v[e];
                if ((Rg & ZMAX) && k == nz - 1) vz[e + 1] = -nvz.v[e];       // vz[nz+1] = -vz[nz]
            }
        }
    }
    pml_close(m0); pml_close(m1); pml_close(m2); pml_close(m3); pml_close(m4); pml_close(m5); pml_close(m6); pml_close(m7); pml_close(m8);
}

// ------------------------------------------------------------------------------------------------
// stress kernel (term order and citations: stress_cell in kernels.cuh)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stress_tile(const Geom& g, const StepArgs& a, const float* S, int lane, int r,
                                            int i, int j, int k0, uint32_t fullbar, uint32_t parity, bool active) {
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int kg0 = k0 + g.koff;
    const long long c = uidx(g, k0, j, i);
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;
    Pml4 m0, m1, m2, m3, m4, m5, m6, m7, m8;
    if (active) {
        const int sxh = slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax);      // dvzdx, dvydx
        const int syh = slab_index(j, 1, ny - 1, g.npml, hymin, hymax);      // dvxdy, dvzdy
        pml_open<2>(m0, g, a.ps[0], slab_index(i, 0, nx, g.npml, hxmin, hxmax), k0, j, i, 0);
        pml_open<1>(m1, g, a.ps[1], slab_index(j, 0, ny, g.npml, hymin, hymax), k0, j, i, 0);
        pml_open_z(m2, g, a.ps[2], 0, nz, k0, j, i, 0);
        pml_open<1>(m3, g, a.ps[3], syh, k0, j, i, 0);   pml_open<2>(m4, g, a.ps[4], sxh, k0, j, i, 0);
        pml_open_z(m5, g, a.ps[5], 1, nz - 1, k0, j, i, 0);   pml_open<2>(m6, g, a.ps[6], sxh, k0, j, i, 0);
        pml_open_z(m7, g, a.ps[7], 1, nz - 1, k0, j, i, 0);   pml_open<1>(m8, g, a.ps[8], syh, k0, j, i, 0);
    }
    mbar_wait(fullbar, parity);
    if (!active) return;

    const float* P = S + 4 * lane + 4;
    auto row = [&](int box, int rr) { return P + box + rr * PITCH; };
    const F4 cvx = lds4(row(S_VX0, r + 1)), cvy = lds4(row(S_VY0, r)), cvz = lds4(row(S_VZ0, r + 1));
    const F4 vxpx = lds4(row(S_VXP, r)), vypy = lds4(row(S_VY0, r + 1));
    const F4 vxmy = lds4(row(S_VX0, r)), vymx = lds4(row(S_VYM, r)), vzmx = lds4(row(S_VZM, r)), vzmy = lds4(row(S_VZ0, r));
    const float vznext = more ? row(S_VZ0, r + 1)[VW] : 0.f;
    const float vxprev = k0 > 0 ? row(S_VX0, r + 1)[-1] : 0.f;
    const float vyprev = k0 > 0 ? row(S_VY0, r)[-1] : 0.f;
    const bool fs = (g.freesurf & ZMIN) != 0;

    F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_apply(m0, a.ps[0], dxx);       // @d_xa(vx)
    F4 dyy = diff4(vypy, cvy, g.dyI);           pml_apply(m1, a.ps[1], dyy);       // @d_ya(vy)
    F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_apply_z(m2, a.ps[2], dzz);     // @d_za(vz)
    F4 xx = lds4(row(S_XX, r)), yy = lds4(row(S_YY, r)), zz = lds4(row(S_ZZ, r));
    const F4 M = lds4(row(S_K, r)), L = lds4(row(S_L, r));
#pragma unroll
    for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo)) {
        xx.v[e] = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dzz.v[e])));
        yy.v[e] = __fsub_rn(__fsub_rn(yy.v[e], __fmul_rn(M.v[e], dyy.v[e])), __fmul_rn(L.v[e], __fadd_rn(dxx.v[e], dzz.v[e])));
        zz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dxx.v[e])));
    }
    if (fs && kg0 == 0) zz.v[0] = -zz.v[1];                     // free_surface_mirror!: tauzz[1] = -tauzz[2]
    float* txx = a.tau[T_XX] + c; float* tyy = a.tau[T_YY] + c; float* tzz = a.tau[T_ZZ] + c;
    st4(txx, xx); st4(tyy, yy); st4(tzz, zz);
    pml_close(m0); pml_close(m1); pml_close(m2);

    // tauxz: z half, y inner, x half
    F4 dxz = diff4_zm(cvx, vxprev, g.dzI);      pml_apply_z(m5, a.ps[5], dxz);     // @d_zi(vx)
    F4 dzx = diff4(cvz, vzmx, g.dxI);           pml_apply(m6, a.ps[6], dzx);       // @d_xi(vz)
    // tauxy: z inner, y half, x half
    F4 dxy = diff4(cvx, vxmy, g.dyI);           pml_apply(m3, a.ps[3], dxy);       // @d_yi(vx)
    F4 dyx = diff4(cvy, vymx, g.dxI);           pml_apply(m4, a.ps[4], dyx);       // @d_xi(vy)
    // tauyz: z half, y half, x inner
    F4 dyz = diff4_zm(cvy, vyprev, g.dzI);      pml_apply_z(m7, a.ps[7], dyz);     // @d_zi(vy)
    F4 dzy = diff4(cvz, vzmy, g.dyI);           pml_apply(m8, a.ps[8], dzy);       // @d_yi(vz)
    F4 xy = lds4(row(S_XY, r)), xz = lds4(row(S_XZ, r)), yz = lds4(row(S_YZ, r));
    const F4 muxz = lds4(row(S_MUXZ, r)), muxy = lds4(row(S_MUXY, r)), muyz = lds4(row(S_MUYZ, r));
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
    float* txy = a.tau[T_XY] + c; float* txz = a.tau[T_XZ] + c; float* tyz = a.tau[T_YZ] + c;
    st4(txz, xz); st4(txy, xy); st4(tyz, yz);
    pml_close(m3); pml_close(m4); pml_close(m5); pml_close(m6); pml_close(m7); pml_close(m8);
}

template <int KIND>
__global__ void __launch_bounds__(NTHREADS, T3_MINB) k_step3t(const Geom g, const StepArgs a, const Sched sc, const __grid_constant__ Maps tm) {
    constexpr int SFLOATS = KIND == 0 ? V_FLOATS : S_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((128 - (s32(smem_dyn) & 127)) & 127);      // boxes need 128-byte alignment
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SFLOATS * 4);
    const uint32_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) { producer<KIND>(g, a, sc, tm, stage0, full0, empty0, lane); return; }

    const int r = warp;
    int n = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++) {
        const int s = n % STAGES;
        const uint32_t use = n / STAGES;
        int i, j0, kc0;
        tile_coords(sc, t, i, j0, kc0);
        const int j = j0 + r, k0 = kc0 + 4 * lane;
        const bool active = j <= sc.jhi && k0 < g.pz;
        const float* S = stage0 + (size_t)s * SFLOATS;
        if (KIND == 0) vel_tile(g, a, S, lane, r, i, j, k0, full0 + 8 * s, use & 1, active);
        else           stress_tile(g, a, S, lane, r, i, j, k0, full0 + 8 * s, use & 1, active);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
#ifndef T3_NOSHELL
    shell_work<KIND>(g, a, sc, r, lane);
#endif
}

}  // namespace t3
