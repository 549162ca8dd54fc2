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
//   * one elected producer thread moves every operand of a tile with TMA tensor copies (cp.async.bulk.tensor.3d
//     global -> shared through one CUtensorMap per box, completion on an mbarrier) into a 2-stage shared-memory
//     ring: 15 (velocity) or 17 (stress) boxes of 4 or 5 rows -- the 15 / 20 arrays at the tile's own rows with
//     their y +-1 halo row, the x +-1 plane rows, a 16-byte z halo where a z neighbour is needed -- plus, only in
//     the CPML slabs, up to 9 boxes of memory variables.  Out-of-range coordinates read zeros.  (Per-row
//     cp.async.bulk copies were tried first: ptxas serialises them lane by lane, ~85 cycles per row.)
//   * the producer also writes a small tile header (plane, first row, slab indices, row masks) into the stage, so
//     the consumers do no tile decoding, no slab tests and no operand address arithmetic at all: they read
//     operands from shared memory (conflict-free 16-byte reads; z +-1 neighbours by warp shuffle), do
//     the reference-order arithmetic and write results (fields and CPML memory) with coalesced 16-byte global
//     stores; a full / empty mbarrier pair per stage is the only synchronisation;
//   * the outer shell of the box (rigid faces, ghost cells, ragged x / y ranges; 2.4 % of the rows) is a separate
//     small launch of the scalar reference-order code (k_shell3) on a side stream, beside the tile kernel: it
//     touches cells no tile touches and needs no shared memory.
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
constexpr int PH = ZC + 8;           // floats per staged row of a box with a 16-byte z halo on both sides
constexpr int STAGES = T3_STAGES;
constexpr int NCW = R;               // consumer warps (one per row)
constexpr int NTHREADS = 32 * (NCW + 1);
constexpr int PZM = 96;              // floats per z-CPML memory row (Geom.pzm must equal this)

// Operand boxes of a stage.  Every box is rows x pitch floats, dense, and starts on a 128-byte boundary.
// Sizes in floats: 4 x 128 = 512, 5 x 128 = 640, 4 x 136 = 544, 5 x 136 = 680 (padded to 704).
constexpr int B4 = 512, B5 = 640, H4 = 544, H5 = 704;
// velocity kernel: box -> offset
enum { V_XX0 = 0, V_XXM = V_XX0 + B4, V_YY = V_XXM + B4 /*5 rows j-1..j+3*/, V_ZZ = V_YY + B5 /*halo*/, V_XY0 = V_ZZ + H4 /*5 rows j..j+4*/,
       V_XYP = V_XY0 + B5, V_XZ0 = V_XYP + B4 /*halo*/, V_XZP = V_XZ0 + H4, V_YZ = V_XZP + B4 /*5 rows, halo*/, V_VX = V_YZ + H5,
       V_VY = V_VX + B4, V_VZ = V_VY + B4, V_BX = V_VZ + B4, V_BY = V_BX + B4, V_BZ = V_BY + B4, V_MAIN = V_BZ + B4, V_NBOX = 15 };
// stress kernel
enum { S_VX0 = 0 /*5 rows j-1..j+3, halo*/, S_VXP = S_VX0 + H5, S_VY0 = S_VXP + B4 /*5 rows j..j+4, halo*/, S_VYM = S_VY0 + H5,
       S_VZ0 = S_VYM + B4 /*5 rows j-1..j+3, halo*/, S_VZM = S_VZ0 + H5, S_XX = S_VZM + B4, S_YY = S_XX + B4, S_ZZ = S_YY + B4,
       S_XY = S_ZZ + B4, S_XZ = S_XY + B4, S_YZ = S_XZ + B4, S_K = S_YZ + B4, S_L = S_K + B4, S_MUXZ = S_L + B4,
       S_MUXY = S_MUXZ + B4, S_MUYZ = S_MUXY + B4, S_MAIN = S_MUYZ + B4, S_NBOX = 17 };
// after the main boxes: CPML memory boxes (x terms, y terms: 4 x 128; z terms: 4 x 96), then the tile header
constexpr int P_X = 0, P_Y = 3 * B4, P_Z = 6 * B4, P_HDR = 6 * B4 + 3 * 4 * PZM, P_FLOATS = P_HDR + 32;
constexpr int MAXBOX = S_NBOX + 9;
// tile header (ints)
enum { H_I = 0, H_J0, H_KC0, H_SX /*3*/, H_SY0 = H_SX + 3 /*3*/, H_YMASK = H_SY0 + 3 /*3*/, H_ZLOAD = H_YMASK + 3, H_N };

// one operand box: which array, plane / first-row offset relative to (i, j0), rows, z halo, offset inside the stage
struct BoxSpec { int arr; int di, dj, rows, halo, off; };      // arr: 0-5 tau, 6-8 v, 9.. coefficient slot + 9
__host__ __device__ inline BoxSpec box_spec(int kind, int b) {
    if (kind == 0) {
        const BoxSpec t[V_NBOX] = {
            {T_XX, 0, 0, 4, 0, V_XX0}, {T_XX, -1, 0, 4, 0, V_XXM}, {T_YY, 0, -1, 5, 0, V_YY}, {T_ZZ, 0, 0, 4, 1, V_ZZ},
            {T_XY, 0, 0, 5, 0, V_XY0}, {T_XY, 1, 0, 4, 0, V_XYP}, {T_XZ, 0, 0, 4, 1, V_XZ0}, {T_XZ, 1, 0, 4, 0, V_XZP},
            {T_YZ, 0, 0, 5, 1, V_YZ}, {6 + V_X, 0, 0, 4, 0, V_VX}, {6 + V_Y, 0, 0, 4, 0, V_VY}, {6 + V_Z, 0, 0, 4, 0, V_VZ},
            {9 + C_BX, 0, 0, 4, 0, V_BX}, {9 + C_BY, 0, 0, 4, 0, V_BY}, {9 + C_BZ, 0, 0, 4, 0, V_BZ}};
        return t[b];
    }
    const BoxSpec t[S_NBOX] = {
        {6 + V_X, 0, -1, 5, 1, S_VX0}, {6 + V_X, 1, 0, 4, 0, S_VXP}, {6 + V_Y, 0, 0, 5, 1, S_VY0}, {6 + V_Y, -1, 0, 4, 0, S_VYM},
        {6 + V_Z, 0, -1, 5, 1, S_VZ0}, {6 + V_Z, -1, 0, 4, 0, S_VZM}, {T_XX, 0, 0, 4, 0, S_XX}, {T_YY, 0, 0, 4, 0, S_YY},
        {T_ZZ, 0, 0, 4, 0, S_ZZ}, {T_XY, 0, 0, 4, 0, S_XY}, {T_XZ, 0, 0, 4, 0, S_XZ}, {T_YZ, 0, 0, 4, 0, S_YZ},
        {9 + C_K, 0, 0, 4, 0, S_K}, {9 + C_L, 0, 0, 4, 0, S_L}, {9 + C_MUXZ, 0, 0, 4, 0, S_MUXZ}, {9 + C_MUXY, 0, 0, 4, 0, S_MUXY},
        {9 + C_MUYZ, 0, 0, 4, 0, S_MUYZ}};
    return t[b];
}
// CPML terms of a kernel grouped by axis: index into StepArgs.pv / .ps, and (s0, len) of the derivative field
// (same numbers as the cpml<> calls of vel_cell / stress_cell)
__host__ __device__ inline int term_index(int kind, int axis /*0 z 1 y 2 x*/, int q) {
    const int v[3][3] = {{2, 5, 8}, {1, 4, 7}, {0, 3, 6}};
    const int s[3][3] = {{2, 5, 7}, {1, 3, 8}, {0, 4, 6}};
    return kind == 0 ? v[axis][q] : s[axis][q];
}
__host__ __device__ inline void term_extent(const Geom& g, int kind, int axis, int q, int& s0, int& len) {
    const int n = axis == 0 ? g.nz : axis == 1 ? g.ny : g.nx;
    if (kind == 0) {
        // x: dtauxxdx (1, n-1) dtauxydx (1, n-2) dtauxzdx (1, n-2); y: dtauxydy (1, n-2) dtauyydy (1, n-1) dtauyzdy (1, n-2)
        // z: dtauxzdz (1, n-2) dtauyzdz (1, n-2) dtauzzdz (1, n-1)
        s0 = 1;
        if (axis == 2) len = q == 0 ? n - 1 : n - 2;
        else if (axis == 1) len = q == 1 ? n - 1 : n - 2;
        else len = q == 2 ? n - 1 : n - 2;
    } else {
        // first term of every axis is the normal derivative on the tauii grid (0, n); the others live on half grids (1, n-1)
        s0 = q == 0 ? 0 : 1; len = q == 0 ? n : n - 1;
    }
}
// the TMA descriptors of one kernel (main boxes, then x, y, z CPML boxes); lives in global memory
struct alignas(64) Maps { unsigned char m[MAXBOX][128]; };

struct Sched {
    int ilo, ihi, jlo, jhi;          // fast region (inclusive)
    int njb, nzc;                    // row blocks per plane, z chunks per row
    int ntiles;
    int sp[4], nsp;                  // shell planes (all rows)
    int sr[4], nsr;                  // shell rows of the fast planes
};

template <int KIND> struct K {
    static constexpr int MAIN = KIND == 0 ? (int)V_MAIN : (int)S_MAIN;
    static constexpr int NBOX = KIND == 0 ? (int)V_NBOX : (int)S_NBOX;
    static constexpr int SFLOATS = MAIN + P_FLOATS;
};
__host__ __device__ inline size_t smem_bytes(int kind) {
    return (size_t)STAGES * (kind == 0 ? K<0>::SFLOATS : K<1>::SFLOATS) * 4 + 2 * STAGES * 8 + 128 + 9 * PZM * 4;
}

// ---- PTX: mbarrier + TMA ---------------------------------------------------------------------------------------------
#ifdef GPI_HOST_EMU
// tests/emu (cuda_rt_shim.h): host forms of the primitives, so that the tile decomposition, the producer's box list and byte
// accounting, the tile header and the consumers' shared-memory indexing run on the CPU.  A CTA is emulated serially (tile by
// tile: producer, then the 128 consumer lanes), so the barriers have nothing to order; the 64-bit barrier word counts the
// bytes still expected instead, and a consumer that finds it non-zero has caught a wrong expect_tx (a hang on the GPU).
typedef uintptr_t sptr_t;
struct EmuTensorMap { const float* base; unsigned long long dim[3]; unsigned box[3]; };       // what the emulated encoder stores in a CUtensorMap
inline sptr_t s32(const void* p) { return (sptr_t)p; }
inline void mbar_init(sptr_t bar, uint32_t) { *(long long*)bar = 0; }
inline void mbar_expect_tx(sptr_t bar, uint32_t bytes) { *(long long*)bar += bytes; }
inline void mbar_arrive(sptr_t) {}
inline void mbar_wait(sptr_t, uint32_t) {}
inline void mbar_check_complete(sptr_t bar) { if (*(long long*)bar != 0) { fprintf(stderr, "t3 emulation: %lld bytes of a stage never arrived / were never expected\n", *(long long*)bar); abort(); } }
inline void tma_box(sptr_t dst, const void* tmap, int c0, int c1, int c2, sptr_t bar) {
    const EmuTensorMap& m = *reinterpret_cast<const EmuTensorMap*>(tmap);
    float* d = reinterpret_cast<float*>(dst);
    for (unsigned z = 0; z < m.box[2]; z++) for (unsigned y = 0; y < m.box[1]; y++) for (unsigned x = 0; x < m.box[0]; x++) {
        const long long k = (long long)c0 + x, j = (long long)c1 + y, i = (long long)c2 + z;
        const bool in = k >= 0 && j >= 0 && i >= 0 && k < (long long)m.dim[0] && j < (long long)m.dim[1] && i < (long long)m.dim[2];
        *d++ = in ? m.base[k + (long long)m.dim[0] * (j + (long long)m.dim[1] * i)] : 0.f;      // out-of-range coordinates read zeros
    }
    *(long long*)bar -= (long long)m.box[0] * m.box[1] * m.box[2] * 4;                            // complete_tx counts the whole box
}
inline F4 lds4(const float* p) { return *reinterpret_cast<const F4*>(p); }
// z neighbours: the value the shuffle fetches from the adjacent lane is the staged float next to this lane's group
inline float z_prev_s(unsigned, const F4&, const float* p, int, bool has) { return has ? p[-1] : 0.f; }
inline float z_next_s(unsigned, const F4&, const float* p, int, bool has) { return has ? p[VW] : 0.f; }
#else
typedef uint32_t sptr_t;             // shared-window address
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
__device__ __forceinline__ F4 lds4(const float* p) { return *reinterpret_cast<const F4*>(p); }
// z neighbours of a group of four: the adjacent lane's edge value by shuffle, the warp's own edge from the staged halo
// (4-byte shared reads by 32 lanes 16 bytes apart would be 4-way bank conflicts)
__device__ __forceinline__ float z_prev_s(unsigned mask, const F4& c, const float* p, int lane, bool has) {
    float v = __shfl_up_sync(mask, c.v[VW - 1], 1);
    if (lane == 0) v = p[-1];
    return has ? v : 0.f;
}
__device__ __forceinline__ float z_next_s(unsigned mask, const F4& c, const float* p, int lane, bool has) {
    float v = __shfl_down_sync(mask, c.v[0], 1);       // the lane above may be outside the mask (k0 >= pz): then has == false
    if (lane == 31) v = p[VW];
    return has ? v : 0.f;
}
#endif   // GPI_HOST_EMU

// ------------------------------------------------------------------------------------------------
// producer: one thread per CTA
// ------------------------------------------------------------------------------------------------
// The tile sequence of a CTA: t = blockIdx.x, blockIdx.x + gridDim.x, ... < ntiles, n = how many tiles came before (stage and
// phase bookkeeping).  The CPU emulation calls producer / consumer once per tile and passes the range (t_first, t_step, t_end, n0).
#ifdef GPI_HOST_EMU
#define T3_TILE_LOOP(t, n) int n = n0; for (int t = t_first; t < t_end; t += t_step, n++)
#else
#define T3_TILE_LOOP(t, n) int n = 0; for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++)
#endif
template <int KIND>
__device__ __forceinline__ void producer(const Geom& g, const Sched& sc, const Maps* tm, float* stage0,
                                         sptr_t full0, sptr_t empty0, int t_first, int t_step, int t_end, int n0) {
    constexpr int NBOX = K<KIND>::NBOX, MAIN = K<KIND>::MAIN, SFLOATS = K<KIND>::SFLOATS;
    int main_bytes = 0;
#pragma unroll
    for (int b = 0; b < NBOX; b++) { const BoxSpec bs = box_spec(KIND, b); main_bytes += bs.rows * (bs.halo ? PH : ZC) * 4; }
    const int npml = g.npml;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;
    // term extents (loop invariant)
    int xs0[3], xlen[3], ys0[3], ylen[3], zs0[3], zlen[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { term_extent(g, KIND, 2, q, xs0[q], xlen[q]); term_extent(g, KIND, 1, q, ys0[q], ylen[q]); term_extent(g, KIND, 0, q, zs0[q], zlen[q]); }
    const int per_plane = sc.njb * sc.nzc;
    T3_TILE_LOOP(t, n) {
        const int s = n % STAGES;
        const uint32_t use = n / STAGES;
        const int ip = t / per_plane, rem = t - ip * per_plane;
        const int jb = rem / sc.nzc, zc = rem - jb * sc.nzc;
        const int i = sc.ilo + ip, j0 = sc.jlo + jb * R, kc0 = zc * ZC;
        // CPML boxes of this tile
        int sx[3], sy0[3], ymask[3], bytes = main_bytes;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            sx[q] = slab_index(i, xs0[q], xlen[q], npml, hxmin, hxmax);
            if (sx[q] >= 0) bytes += R * ZC * 4;
            ymask[q] = 0; sy0[q] = 0;
#pragma unroll
            for (int r = R - 1; r >= 0; r--) {
                const int sr = slab_index(j0 + r, ys0[q], ylen[q], npml, hymin, hymax);
                if (sr >= 0) { ymask[q] |= 1 << r; sy0[q] = sr - r; }
            }
            if (ymask[q]) bytes += R * ZC * 4;
        }
        bool zload = false;
        {
            const int kg0 = kc0 + g.koff, kg1 = kg0 + min(ZC, g.pz - kc0);       // global z range of the chunk
#pragma unroll
            for (int q = 0; q < 3; q++) {
                zload |= (g.pml & ZMIN) && kg0 < zs0[q] + npml;
                zload |= (g.pml & ZMAX) && kg1 > zslab_base(zs0[q], zlen[q], npml);
            }
        }
        if (zload) bytes += 3 * R * PZM * 4;

        mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
        float* S = stage0 + (size_t)s * SFLOATS;
        int* hdr = reinterpret_cast<int*>(S + MAIN + P_HDR);
        hdr[H_I] = i; hdr[H_J0] = j0; hdr[H_KC0] = kc0; hdr[H_ZLOAD] = zload ? 1 : 0;
#pragma unroll
        for (int q = 0; q < 3; q++) { hdr[H_SX + q] = sx[q]; hdr[H_SY0 + q] = sy0[q]; hdr[H_YMASK + q] = ymask[q]; }
        const sptr_t bar = full0 + 8 * s;
        mbar_expect_tx(bar, bytes);                 // release: the header is visible to whoever observes the phase
        const sptr_t dst0 = s32(S);
#pragma unroll
        for (int b = 0; b < NBOX; b++) {
            const BoxSpec bs = box_spec(KIND, b);
            tma_box(dst0 + bs.off * 4, tm->m[b], kc0 - (bs.halo ? 4 : 0), j0 + bs.dj, i + bs.di, bar);
        }
        const sptr_t dstp = dst0 + MAIN * 4;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (sx[q] >= 0) tma_box(dstp + (P_X + q * B4) * 4, tm->m[NBOX + q], kc0, j0, sx[q], bar);          // [k, j, s]
            if (ymask[q])   tma_box(dstp + (P_Y + q * B4) * 4, tm->m[NBOX + 3 + q], kc0, sy0[q], i, bar);      // [k, s, i]
            if (zload)      tma_box(dstp + (P_Z + q * 4 * PZM) * 4, tm->m[NBOX + 6 + q], 0, j0, i, bar);       // [zi, j, i]
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CPML on four z-consecutive values with the memory variables staged in shared memory
// ------------------------------------------------------------------------------------------------
// x / y terms: slab index s uniform over the four cells; gm = global address of the four memory variables
__device__ __forceinline__ void pml_s(const float* sm, float* gm, const float (&cf)[3], F4& d) {
    F4 m = lds4(sm);
    const float a = cf[0], bb = cf[1], kI = cf[2];
#pragma unroll
    for (int e = 0; e < VW; e++) {
        m.v[e] = __fadd_rn(__fmul_rn(bb, m.v[e]), __fmul_rn(a, d.v[e]));
        d.v[e] = __fadd_rn(__fmul_rn(d.v[e], kI), m.v[e]);
    }
    st4(gm, m);
}
// z terms: zi = position inside the 96-float memory row (-1: this group of four is outside the slabs);
// coefficient tables are indexed by the global z index (identity outside the slabs)
__device__ __forceinline__ int z_slot(const Geom& g, int s0, int len, int kg0) {
    const int npml = g.npml;
    if ((g.pml & ZMIN) && kg0 < s0 + npml) return kg0;
    if (g.pml & ZMAX) {
        const int kb = zslab_base(s0, len, npml);
        if (kg0 >= kb && kg0 < s0 + len) return (PZM >> 1) + (kg0 - kb);
    }
    return -1;
}
// zt: this term's coefficients re-indexed like the memory row ([a | b | kI][PZM], filled once per CTA)
__device__ __forceinline__ void pml_z(const float* smrow, float* gmrow, const float* zt, int zi, F4& d) {
    if (zi < 0) return;
    F4 m = lds4(smrow + zi);
    const F4 a = lds4(zt + zi), bb = lds4(zt + PZM + zi), kI = lds4(zt + 2 * PZM + zi);
#pragma unroll
    for (int e = 0; e < VW; e++) {
        m.v[e] = __fadd_rn(__fmul_rn(bb.v[e], m.v[e]), __fmul_rn(a.v[e], d.v[e]));
        d.v[e] = __fadd_rn(__fmul_rn(d.v[e], kI.v[e]), m.v[e]);
    }
    st4(gmrow + zi, m);
}

// per-thread view of a tile
struct TileCtx {
    const float* S;        // stage base
    const float* P;        // CPML boxes base
    const float* ZT;       // z coefficient tables of the three z terms (shared memory)
    int i, j, k0, kg0, r, lane;
    unsigned mask;         // lanes of this warp that own cells of the tile (shuffle mask)
    int sx[3], sy[3];      // slab index of the x terms (plane) and of the y terms (this row), -1 = none
    float cx[3][3], cy[3][3];   // their a, b, kI, requested as soon as the tile header is read (L1 / L2 hits, but ~200+ cycles)
    bool zload;
    long long c;           // unified index of the first cell
};

template <int KIND>
__device__ __forceinline__ void open_tile(TileCtx& q, const Geom& g, const StepArgs& a, const float* S, int r, int lane) {
    const int* hdr = reinterpret_cast<const int*>(S + K<KIND>::MAIN + P_HDR);
    q.S = S; q.P = S + K<KIND>::MAIN; q.r = r; q.lane = lane;
    q.i = hdr[H_I]; q.j = hdr[H_J0] + r; q.k0 = hdr[H_KC0] + 4 * lane; q.kg0 = q.k0 + g.koff;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        q.sx[t] = hdr[H_SX + t];
        q.sy[t] = ((hdr[H_YMASK + t] >> r) & 1) ? hdr[H_SY0 + t] + r : -1;
    }
    q.zload = hdr[H_ZLOAD] != 0;
    q.c = uidx(g, q.k0, q.j, q.i);
#pragma unroll
    for (int t = 0; t < 3; t++) {
        const PmlTerm& tx = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, 2, t)];
        const PmlTerm& ty = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, 1, t)];
        const int s = q.sx[t], u = q.sy[t];
        if (s >= 0) { q.cx[t][0] = ld1(tx.a + s); q.cx[t][1] = ld1(tx.b + s); q.cx[t][2] = ld1(tx.kI + s); }
        if (u >= 0) { q.cy[t][0] = ld1(ty.a + u); q.cy[t][1] = ld1(ty.b + u); q.cy[t][2] = ld1(ty.kI + u); }
    }
}
// apply the CPML term (axis, q) of kernel KIND to d
template <int KIND, int AXIS, int Q>
__device__ __forceinline__ void pml_t(const TileCtx& q, const Geom& g, const StepArgs& a, F4& d) {
#ifdef GPI_EXP_NOPML
    return;
#endif
    const PmlTerm& t = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, AXIS, Q)];
    if (AXIS == 2) {
        const int s = q.sx[Q];
        if (s < 0) return;
        pml_s(q.P + P_X + Q * B4 + q.r * ZC + 4 * q.lane,
              t.mem + (long long)q.k0 + (long long)g.pz * ((long long)q.j + (long long)g.ny1 * s), q.cx[Q], d);
    } else if (AXIS == 1) {
        const int s = q.sy[Q];
        if (s < 0) return;
        pml_s(q.P + P_Y + Q * B4 + q.r * ZC + 4 * q.lane,
              t.mem + (long long)q.k0 + (long long)g.pz * ((long long)s + 2LL * g.npml * q.i), q.cy[Q], d);
    } else {
        if (!q.zload) return;
        int s0, len;
        term_extent(g, KIND, 0, Q, s0, len);
        pml_z(q.P + P_Z + (Q * 4 + q.r) * PZM, t.mem + (long long)g.pzm * ((long long)q.j + (long long)g.ny1 * q.i),
              q.ZT + Q * 3 * PZM, z_slot(g, s0, len, q.kg0), d);
    }
}

// ------------------------------------------------------------------------------------------------
// velocity kernel (term order and citations: vel_cell in kernels.cuh)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vel_tile(const Geom& g, const StepArgs& a, const TileCtx& q) {
    const int nz = g.nz, k0 = q.k0, kg0 = q.kg0, r = q.r;
    const bool more = k0 + VW < g.pz;
    const float* B = q.S + 4 * q.lane;                    // plain boxes
    const float* H = q.S + 4 * q.lane + 4;                // boxes with a z halo
    const F4 xx = lds4(B + V_XX0 + r * ZC), xxm = lds4(B + V_XXM + r * ZC);
    const F4 yy = lds4(B + V_YY + (r + 1) * ZC), yym = lds4(B + V_YY + r * ZC);
    const F4 zz = lds4(H + V_ZZ + r * PH);
    const F4 xy = lds4(B + V_XY0 + r * ZC), xypy = lds4(B + V_XY0 + (r + 1) * ZC), xypx = lds4(B + V_XYP + r * ZC);
    const F4 xz = lds4(H + V_XZ0 + r * PH), xzpx = lds4(B + V_XZP + r * ZC);
    const F4 yz = lds4(H + V_YZ + r * PH), yzpy = lds4(H + V_YZ + (r + 1) * PH);
    const float zzprev = z_prev_s(q.mask, zz, H + V_ZZ + r * PH, q.lane, k0 > 0);
    const float xznext = z_next_s(q.mask, xz, H + V_XZ0 + r * PH, q.lane, more);
    const float yznext = z_next_s(q.mask, yz, H + V_YZ + r * PH, q.lane, more);

    // vx: dtauxxdx + dtauxydy + dtauxzdz
    F4 dxx = diff4(xx, xxm, g.dxI);            pml_t<0, 2, 0>(q, g, a, dxx);
    F4 dxy = diff4(xypy, xy, g.dyI);           pml_t<0, 1, 0>(q, g, a, dxy);
    F4 dxz = diff4_zp(xz, xznext, g.dzI);      pml_t<0, 0, 0>(q, g, a, dxz);
    // vy: dtauxydx + dtauyydy + dtauyzdz
    F4 dyx = diff4(xypx, xy, g.dxI);           pml_t<0, 2, 1>(q, g, a, dyx);
    F4 dyy = diff4(yy, yym, g.dyI);            pml_t<0, 1, 1>(q, g, a, dyy);
    F4 dyz = diff4_zp(yz, yznext, g.dzI);      pml_t<0, 0, 1>(q, g, a, dyz);
    // vz: dtauxzdx + dtauyzdy + dtauzzdz
    F4 dzx = diff4(xzpx, xz, g.dxI);           pml_t<0, 2, 2>(q, g, a, dzx);
    F4 dzy = diff4(yzpy, yz, g.dyI);           pml_t<0, 1, 2>(q, g, a, dzy);
    F4 dzz = diff4_zm(zz, zzprev, g.dzI);      pml_t<0, 0, 2>(q, g, a, dzz);

    F4 nvx = lds4(B + V_VX + r * ZC), nvy = lds4(B + V_VY + r * ZC), nvz = lds4(B + V_VZ + r * ZC);
    const F4 bx = lds4(B + V_BX + r * ZC), by = lds4(B + V_BY + r * ZC), bz = lds4(B + V_BZ + r * ZC);
    float* vx = a.v[V_X] + q.c; float* vy = a.v[V_Y] + q.c; float* vz = a.v[V_Z] + q.c;
    const int Rg = g.rigid;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    const bool allown = k0 >= g.klo && k0 + VW - 1 <= g.khi;
    if (!head && !tail && allown) {
        // every cell of the group is an interior node of vx, vy and vz: no predicates
#pragma unroll
        for (int e = 0; e < VW; e++) {
            nvx.v[e] = __fsub_rn(nvx.v[e], __fmul_rn(bx.v[e], __fadd_rn(__fadd_rn(dxx.v[e], dxy.v[e]), dxz.v[e])));
            nvy.v[e] = __fsub_rn(nvy.v[e], __fmul_rn(by.v[e], __fadd_rn(__fadd_rn(dyx.v[e], dyy.v[e]), dyz.v[e])));
            nvz.v[e] = __fsub_rn(nvz.v[e], __fmul_rn(bz.v[e], __fadd_rn(__fadd_rn(dzx.v[e], dzy.v[e]), dzz.v[e])));
        }
        st4(vx, nvx); st4(vy, nvy); st4(vz, nvz);
        return;
    }
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
    // rigid z faces (dirichlet.jl:35-74); the x / y faces only touch shell rows (scalar path)
    if (head && (Rg & ZMIN)) { nvx.v[0] = 0.f; nvy.v[0] = 0.f; nvz.v[0] = -nvz.v[1]; }
    if (!tail && allown) {
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
}

// ------------------------------------------------------------------------------------------------
// stress kernel (term order and citations: stress_cell in kernels.cuh)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stress_tile(const Geom& g, const StepArgs& a, const TileCtx& q) {
    const int nz = g.nz, k0 = q.k0, kg0 = q.kg0, r = q.r;
    const bool more = k0 + VW < g.pz;
    const float* B = q.S + 4 * q.lane;
    const float* H = q.S + 4 * q.lane + 4;
    const F4 cvx = lds4(H + S_VX0 + (r + 1) * PH), cvy = lds4(H + S_VY0 + r * PH), cvz = lds4(H + S_VZ0 + (r + 1) * PH);
    const F4 vxpx = lds4(B + S_VXP + r * ZC), vypy = lds4(H + S_VY0 + (r + 1) * PH);
    const F4 vxmy = lds4(H + S_VX0 + r * PH), vymx = lds4(B + S_VYM + r * ZC), vzmx = lds4(B + S_VZM + r * ZC), vzmy = lds4(H + S_VZ0 + r * PH);
    const float vznext = z_next_s(q.mask, cvz, H + S_VZ0 + (r + 1) * PH, q.lane, more);
    const float vxprev = z_prev_s(q.mask, cvx, H + S_VX0 + (r + 1) * PH, q.lane, k0 > 0);
    const float vyprev = z_prev_s(q.mask, cvy, H + S_VY0 + r * PH, q.lane, k0 > 0);
    const bool fs = (g.freesurf & ZMIN) != 0;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    const bool allown = k0 >= g.klo && k0 + VW - 1 <= g.khi;
    const bool plain = !head && !tail && allown;        // every cell is an interior node of all six stresses (k >= 4: no free-surface row)

    F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_t<1, 2, 0>(q, g, a, dxx);      // @d_xa(vx)
    F4 dyy = diff4(vypy, cvy, g.dyI);           pml_t<1, 1, 0>(q, g, a, dyy);      // @d_ya(vy)
    F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_t<1, 0, 0>(q, g, a, dzz);      // @d_za(vz)
    F4 xx = lds4(B + S_XX + r * ZC), yy = lds4(B + S_YY + r * ZC), zz = lds4(B + S_ZZ + r * ZC);
    const F4 M = lds4(B + S_K + r * ZC), L = lds4(B + S_L + r * ZC);
#pragma unroll
    for (int e = 0; e < VW; e++) if (plain || (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo))) {
        xx.v[e] = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dzz.v[e])));
        yy.v[e] = __fsub_rn(__fsub_rn(yy.v[e], __fmul_rn(M.v[e], dyy.v[e])), __fmul_rn(L.v[e], __fadd_rn(dxx.v[e], dzz.v[e])));
        zz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dxx.v[e])));
    }
    if (fs && kg0 == 0) zz.v[0] = -zz.v[1];                     // free_surface_mirror!: tauzz[1] = -tauzz[2]
    float* txx = a.tau[T_XX] + q.c; float* tyy = a.tau[T_YY] + q.c; float* tzz = a.tau[T_ZZ] + q.c;
    st4(txx, xx); st4(tyy, yy); st4(tzz, zz);

    // tauxz: z half, y inner, x half
    F4 dxz = diff4_zm(cvx, vxprev, g.dzI);      pml_t<1, 0, 1>(q, g, a, dxz);      // @d_zi(vx)
    F4 dzx = diff4(cvz, vzmx, g.dxI);           pml_t<1, 2, 2>(q, g, a, dzx);      // @d_xi(vz)
    // tauxy: z inner, y half, x half
    F4 dxy = diff4(cvx, vxmy, g.dyI);           pml_t<1, 1, 1>(q, g, a, dxy);      // @d_yi(vx)
    F4 dyx = diff4(cvy, vymx, g.dxI);           pml_t<1, 2, 1>(q, g, a, dyx);      // @d_xi(vy)
    // tauyz: z half, y half, x inner
    F4 dyz = diff4_zm(cvy, vyprev, g.dzI);      pml_t<1, 0, 2>(q, g, a, dyz);      // @d_zi(vy)
    F4 dzy = diff4(cvz, vzmy, g.dyI);           pml_t<1, 1, 2>(q, g, a, dzy);      // @d_yi(vz)
    F4 xy = lds4(B + S_XY + r * ZC), xz = lds4(B + S_XZ + r * ZC), yz = lds4(B + S_YZ + r * ZC);
    const F4 muxz = lds4(B + S_MUXZ + r * ZC), muxy = lds4(B + S_MUXY + r * ZC), muyz = lds4(B + S_MUYZ + r * ZC);
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = kg0 + e; const bool own = plain || (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
        if (plain || (own && k >= 1 && k <= nz - 1)) {
            float n = __fsub_rn(xz.v[e], __fmul_rn(muxz.v[e], __fadd_rn(dxz.v[e], dzx.v[e])));
            if (!plain && fs && k == 1) n = 0.f;                   // free_surface!(tauxz)
            xz.v[e] = n;
            n = __fsub_rn(yz.v[e], __fmul_rn(muyz.v[e], __fadd_rn(dyz.v[e], dzy.v[e])));
            if (!plain && fs && k == 1) n = 0.f;                   // free_surface!(tauyz)
            yz.v[e] = n;
        }
        if (plain || (own && k >= 1 && k <= nz - 2)) xy.v[e] = __fsub_rn(xy.v[e], __fmul_rn(muxy.v[e], __fadd_rn(dxy.v[e], dyx.v[e])));
    }
    float* txy = a.tau[T_XY] + q.c; float* txz = a.tau[T_XZ] + q.c; float* tyz = a.tau[T_YZ] + q.c;
    st4(txz, xz); st4(txy, xy); st4(tyz, yz);
}

// z-CPML coefficient tables, re-indexed like the memory rows: zi < PZM/2 is the min slab (zi = k), the upper half
// is the max slab from its 4-aligned start; identity (a = b = 0, kI = 1) where the k-indexed table has no entry
template <int KIND>
__device__ __forceinline__ void fill_zt(const Geom& g, const StepArgs& a, float* ZT, int first, int step) {
    for (int idx = first; idx < 9 * PZM; idx += step) {
        const int Q = idx / (3 * PZM), cc = (idx / PZM) % 3, zi = idx % PZM;
        int s0, len;
        term_extent(g, KIND, 0, Q, s0, len);
        const int k = zi < (PZM >> 1) ? zi : zslab_base(s0, len, g.npml) + zi - (PZM >> 1);
        const PmlTerm& t = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, 0, Q)];
        const float* tab = cc == 0 ? t.a : cc == 1 ? t.b : t.kI;
        ZT[idx] = (k >= 0 && k <= g.nz && tab) ? __ldg(tab + k) : (cc == 2 ? 1.f : 0.f);
    }
}
// one consumer lane: the same tile sequence as the producer (see there for the range arguments)
template <int KIND>
__device__ __forceinline__ void consumer(const Geom& g, const StepArgs& a, const Sched& sc, const float* stage0, const float* ZT,
                                         sptr_t full0, sptr_t empty0, int warp, int lane, int t_first, int t_step, int t_end, int n0) {
    constexpr int SFLOATS = K<KIND>::SFLOATS;
    T3_TILE_LOOP(t, n) {
        const int s = n % STAGES;
        const uint32_t use = n / STAGES;
        const float* S = stage0 + (size_t)s * SFLOATS;
        mbar_wait(full0 + 8 * s, use & 1);
#ifdef GPI_HOST_EMU
        mbar_check_complete(full0 + 8 * s);
#endif
        TileCtx q;
        q.ZT = ZT;
        open_tile<KIND>(q, g, a, S, warp, lane);
        const bool active = q.j <= sc.jhi && q.k0 < g.pz;
        q.mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            if (KIND == 0) vel_tile(g, a, q); else stress_tile(g, a, q);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
}

#ifdef GPI_HOST_EMU
// CPU emulation of one CTA (run by the launch's "thread 0" of each block; the other 159 return): shared memory is a heap block,
// the tiles of the CTA are taken one after the other -- producer, then the 4 x 32 consumer lanes.
template <int KIND>
void k_step3t(const Geom g, const StepArgs a, const Sched sc, const Maps* tm) {
    if (threadIdx.x != 0) return;
    constexpr int SFLOATS = K<KIND>::SFLOATS;
    const size_t nbytes = smem_bytes(KIND);
    unsigned char* smem_raw = static_cast<unsigned char*>(aligned_alloc(128, (nbytes + 127) / 128 * 128));
    memset(smem_raw, 0xff, nbytes);                       // shared memory starts as garbage (NaN patterns), not zeros
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SFLOATS * 4);
    const sptr_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    float* ZT = reinterpret_cast<float*>(bars + 2 * STAGES);
    fill_zt<KIND>(g, a, ZT, 0, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NCW); }
    int n = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++) {
        producer<KIND>(g, sc, tm, stage0, full0, empty0, t, 1, t + 1, n);
        for (int warp = 0; warp < NCW; warp++) for (int lane = 0; lane < 32; lane++)
            consumer<KIND>(g, a, sc, stage0, ZT, full0, empty0, warp, lane, t, 1, t + 1, n);
    }
    free(smem_raw);
}
#else
template <int KIND>
__global__ void __launch_bounds__(NTHREADS, T3_MINB) k_step3t(const Geom g, const StepArgs a, const Sched sc, const Maps* __restrict__ tm) {
    constexpr int SFLOATS = K<KIND>::SFLOATS;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((128 - (s32(smem_dyn) & 127)) & 127);      // boxes need 128-byte alignment
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SFLOATS * 4);
    const sptr_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    float* ZT = reinterpret_cast<float*>(bars + 2 * STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fill_zt<KIND>(g, a, ZT, threadIdx.x, NTHREADS);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) producer<KIND>(g, sc, tm, stage0, full0, empty0, 0, 0, 0, 0);
        return;
    }
    consumer<KIND>(g, a, sc, stage0, ZT, full0, empty0, warp, lane, 0, 0, 0, 0);
}
#endif

// ------------------------------------------------------------------------------------------------
// shell: the rows / planes outside the fast region, scalar reference-order code.  grid.x = line (a row of one
// plane), threads = z cells
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128, 8) k_shell3(const Geom g, const StepArgs a, const Sched sc) {
    const int L = blockIdx.x;
    int i, j;
    if (L < sc.nsp * g.ny1) { const int q = L / g.ny1; i = sc.sp[q]; j = L - q * g.ny1; }
    else { const int L2 = L - sc.nsp * g.ny1; const int q = L2 / sc.nsr; i = sc.ilo + q; j = sc.sr[L2 - q * sc.nsr]; }
    for (int k = threadIdx.x; k < g.pz; k += blockDim.x) if (k >= g.klo && k <= g.khi) {
        if (KIND == 0) vel_cell<3, 1>(g, a, k, j, i, 0);
        else           stress_cell<3, 1>(g, a, k, j, i, 0);
    }
}

}  // namespace t3
