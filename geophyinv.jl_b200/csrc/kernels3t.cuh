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
//   * persistent CTAs (2 per SM), each = consumer warps + producer warps + shell warps (layout per kernel: struct L below), walk a
//     static round-robin list of tiles; a tile = R = 4 rows (y) x ZC = 128 cells (z) of one x plane, ordered z-chunk, row block, plane
//     so that the tiles in flight at any moment cover ~1.2 consecutive planes (+-1-plane operands are L2 hits);
//   * one elected thread per producer warp moves its share of the operands of a tile with TMA tensor copies
//     (cp.async.bulk.tensor.3d global -> shared through one CUtensorMap per box, completion on an mbarrier that counts one arrival
//     per producer plus the announced bytes) into a 2-stage shared-memory ring: 15 (velocity) or 17 (stress) boxes of 4 or 5 rows --
//     the 15 / 20 arrays at the tile's own rows with their y +-1 halo row, the x +-1 plane rows, a 16-byte z halo where a z neighbour
//     is needed -- plus, only in the CPML slabs, up to 9 boxes of memory variables.  Out-of-range coordinates read zeros.  (Per-row
//     cp.async.bulk copies were tried first: ptxas serialises them lane by lane, ~85 cycles per row.)
//   * everything about a tile that does not depend on the fields (plane, first row, slab indices, row masks) comes from a tile table
//     built once per handle on the host (fill_tile_table); producer 0 copies the record into the stage as the tile header, so neither
//     the producers nor the consumers decode tiles, test slabs or compute operand addresses: the consumers read operands from shared
//     memory (conflict-free 16-byte reads; z +-1 neighbours by warp shuffle), do the reference-order arithmetic and write results
//     (fields and CPML memory) with coalesced 16-byte global stores; a full / empty mbarrier pair per stage is the only synchronisation.
//     (r01 decoded tiles in the single producer thread: ~700 dependent instructions per tile, as long as the tile period itself.)
//   * the outer shell of the box (rigid faces, ghost cells, ragged x / y ranges; 2.4 % of the rows) is the scalar reference-order
//     code, walked by the shell warps of every CTA beside the tile pipeline (shell_lines): it touches cells no tile touches and needs
//     no shared memory.  A dependent load costs ~3 us while the tiles saturate HBM, so the shell must never sit on the pipeline's
//     critical path (a separate launch ran as a 40 - 80 us tail; shell cells between two tiles of a consumer warp cost 60 %).
//
// The z-slab window (Geom.koff / klo / khi) is honoured exactly as in kernels3d.cuh, so the slab decomposition
// uses the same kernels.

// Tile geometry (T3_ZC, T3_R), resident CTAs per SM (T3_MINB) and warp layouts (T3_V_LAYOUT, T3_S_LAYOUT) come from the includer.
namespace T3_NS {

#ifndef T3_STAGES
#define T3_STAGES 2
#endif
constexpr int R = T3_R;               // rows per tile
constexpr int MINB = T3_MINB;        // resident CTAs per SM the register and shared-memory budgets are sized for
constexpr int ZC = T3_ZC;            // z cells per tile (a multiple of 32)
constexpr int PH = ZC + 8;           // floats per staged row of a box with a 16-byte z halo on both sides
constexpr int STAGES = T3_STAGES;
// Warp layout of a CTA, per kernel (measured on B200, profiles/r02/tuning.md):
//   NCOMP  consumer warps per row: 3 = one per output group (velocity: vx | vy | vz; stress: normal | xz, yz | xy), 1 = one warp takes
//          the three groups in turn;
//   NPROD  producer warps (one elected thread each); operand box b belongs to producer b % NPROD;
//   NSHELL warps that walk the shell lines beside the tile pipeline; SHELLC = 1: the consumer warps take one 32-cell shell unit every
//          16th tile instead (their slack between two tiles is longer than a shell cell's dependent-load chain).
// The macros T3_V_* / T3_S_* come from the includer (kernels.cuh: one set per tile width).
constexpr int LAYOUT_V[4] = {(T3_V_LAYOUT) / 1000, (T3_V_LAYOUT) / 100 % 10, (T3_V_LAYOUT) / 10 % 10, (T3_V_LAYOUT) % 10};      // NCOMP, NPROD, NSHELL, SHELLC
constexpr int LAYOUT_S[4] = {(T3_S_LAYOUT) / 1000, (T3_S_LAYOUT) / 100 % 10, (T3_S_LAYOUT) / 10 % 10, (T3_S_LAYOUT) % 10};
template <int KIND> struct L {
    static constexpr int NCOMP = KIND == 0 ? LAYOUT_V[0] : LAYOUT_S[0];
    static constexpr int NPROD = KIND == 0 ? LAYOUT_V[1] : LAYOUT_S[1];
    static constexpr int NSHELL = KIND == 0 ? LAYOUT_V[2] : LAYOUT_S[2];
    static constexpr int SHELLC = KIND == 0 ? LAYOUT_V[3] : LAYOUT_S[3];
    static constexpr int NCW = R * NCOMP;                 // consumer warps
    static constexpr int NTHREADS = 32 * (NCW + NPROD + NSHELL);
    static constexpr bool INKERNEL_SHELL = NSHELL > 0 || SHELLC;       // false: the shell is the separate launch k_shell3
    static_assert(NCOMP == 1 || NCOMP == 3, "NCOMP: 1 or 3");
    static_assert(NPROD >= 1 && NPROD <= 4, "NPROD: 1..4");
    static_assert(!SHELLC || NCW <= 16, "SHELLC: the consumer warps take turns modulo 16 tiles");
};
constexpr int SHELL_EVERY = 16;
constexpr int PZM = 96;              // floats per z-CPML memory row (Geom.pzm must equal this)

// Operand boxes of a stage.  Every box is rows x pitch floats, dense, and starts on a 128-byte boundary.
// Sizes in floats at ZC = 128: 4 x 128 = 512, 5 x 128 = 640, 4 x 136 = 544, 5 x 136 = 680 (padded to 704); every size is a multiple of 32.
constexpr int B4 = R * ZC, B5 = (R + 1) * ZC, H4 = R * PH, H5 = ((R + 1) * PH + 31) / 32 * 32;       // B4 / H4: R rows, B5 / H5: R + 1 rows (y halo)
static_assert(ZC % 32 == 0 && B4 % 32 == 0 && B5 % 32 == 0 && H4 % 32 == 0 && H5 % 32 == 0, "operand boxes start on 128-byte boundaries");
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
constexpr int P_X = 0, P_Y = 3 * B4, P_Z = 6 * B4, P_HDR = 6 * B4 + 3 * R * PZM, P_FLOATS = P_HDR + 32;
constexpr int MAXBOX = S_NBOX + 9;
// tile header (ints) = one record of the tile table
enum { H_I = 0, H_J0, H_KC0, H_SX /*3*/, H_SY0 = H_SX + 3 /*3*/, H_YMASK = H_SY0 + 3 /*3*/, H_ZLOAD = H_YMASK + 3, H_N, H_REC = 16 };
// The tile table: everything about tile t that does not depend on the fields -- plane, first row, z chunk, the slab index of the plane
// for the three x terms, first slab index and row mask for the three y terms, whether the z-memory rows are staged.  Computed once per
// (handle, kernel) on the host (fill_tile_table) so that the producers do no tile decoding: a single thread running ~350 dependent integer
// instructions per tile was what bounded the r01 kernels (profiles/r02/tuning.md).
struct alignas(16) TileRec { int v[H_REC]; };

// one operand box: which array, plane / first-row offset relative to (i, j0), rows, z halo, offset inside the stage
struct BoxSpec { int arr; int di, dj, rows, halo, off; };      // arr: 0-5 tau, 6-8 v, 9.. coefficient slot + 9
__host__ __device__ inline BoxSpec box_spec(int kind, int b) {
    if (kind == 0) {
        const BoxSpec t[V_NBOX] = {
            {T_XX, 0, 0, R, 0, V_XX0}, {T_XX, -1, 0, R, 0, V_XXM}, {T_YY, 0, -1, R + 1, 0, V_YY}, {T_ZZ, 0, 0, R, 1, V_ZZ},
            {T_XY, 0, 0, R + 1, 0, V_XY0}, {T_XY, 1, 0, R, 0, V_XYP}, {T_XZ, 0, 0, R, 1, V_XZ0}, {T_XZ, 1, 0, R, 0, V_XZP},
            {T_YZ, 0, 0, R + 1, 1, V_YZ}, {6 + V_X, 0, 0, R, 0, V_VX}, {6 + V_Y, 0, 0, R, 0, V_VY}, {6 + V_Z, 0, 0, R, 0, V_VZ},
            {9 + C_BX, 0, 0, R, 0, V_BX}, {9 + C_BY, 0, 0, R, 0, V_BY}, {9 + C_BZ, 0, 0, R, 0, V_BZ}};
        return t[b];
    }
    const BoxSpec t[S_NBOX] = {
        {6 + V_X, 0, -1, R + 1, 1, S_VX0}, {6 + V_X, 1, 0, R, 0, S_VXP}, {6 + V_Y, 0, 0, R + 1, 1, S_VY0}, {6 + V_Y, -1, 0, R, 0, S_VYM},
        {6 + V_Z, 0, -1, R + 1, 1, S_VZ0}, {6 + V_Z, -1, 0, R, 0, S_VZM}, {T_XX, 0, 0, R, 0, S_XX}, {T_YY, 0, 0, R, 0, S_YY},
        {T_ZZ, 0, 0, R, 0, S_ZZ}, {T_XY, 0, 0, R, 0, S_XY}, {T_XZ, 0, 0, R, 0, S_XZ}, {T_YZ, 0, 0, R, 0, S_YZ},
        {9 + C_K, 0, 0, R, 0, S_K}, {9 + C_L, 0, 0, R, 0, S_L}, {9 + C_MUXZ, 0, 0, R, 0, S_MUXZ}, {9 + C_MUXY, 0, 0, R, 0, S_MUXY},
        {9 + C_MUYZ, 0, 0, R, 0, S_MUYZ}};
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
    const TileRec* tiles;            // [ntiles] (device)
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
// tile: producer, then the NCW x 32 consumer lanes), so the barriers have nothing to order; the 64-bit barrier word counts the
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
    if (lane == ZC / 4 - 1) v = p[VW];
    return has ? v : 0.f;
}
#endif   // GPI_HOST_EMU

// ------------------------------------------------------------------------------------------------
// tile table (host) and producers: one elected thread per producer warp
// ------------------------------------------------------------------------------------------------
inline int slab_index_h(int u, int s0, int len, int npml, bool hmin, bool hmax) {
    const int r = u - s0;
    if (hmin && r >= 0 && r < npml) return r;
    const int rm = r - (len - npml);
    if (hmax && rm >= 0 && rm < npml) return npml + rm;
    return -1;
}
inline void fill_tile_table(const Geom& g, const Sched& sc, int kind, TileRec* out) {
    const int npml = g.npml;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX, hymin = g.pml & YMIN, hymax = g.pml & YMAX;
    int xs0[3], xlen[3], ys0[3], ylen[3], zs0[3], zlen[3];
    for (int q = 0; q < 3; q++) { term_extent(g, kind, 2, q, xs0[q], xlen[q]); term_extent(g, kind, 1, q, ys0[q], ylen[q]); term_extent(g, kind, 0, q, zs0[q], zlen[q]); }
    const int per_plane = sc.njb * sc.nzc;
    for (int t = 0; t < sc.ntiles; t++) {
        int* h = out[t].v;
        for (int q = 0; q < H_REC; q++) h[q] = 0;
        const int ip = t / per_plane, rem = t - ip * per_plane;
        const int jb = rem / sc.nzc, zc = rem - jb * sc.nzc;
        const int i = sc.ilo + ip, j0 = sc.jlo + jb * R, kc0 = zc * ZC;
        h[H_I] = i; h[H_J0] = j0; h[H_KC0] = kc0;
        for (int q = 0; q < 3; q++) {
            h[H_SX + q] = slab_index_h(i, xs0[q], xlen[q], npml, hxmin, hxmax);
            for (int r = R - 1; r >= 0; r--) {
                const int sr = slab_index_h(j0 + r, ys0[q], ylen[q], npml, hymin, hymax);
                if (sr >= 0) { h[H_YMASK + q] |= 1 << r; h[H_SY0 + q] = sr - r; }
            }
        }
        const int kg0 = kc0 + g.koff, kg1 = kg0 + (ZC < g.pz - kc0 ? ZC : g.pz - kc0);       // global z range of the chunk
        bool zload = false;
        for (int q = 0; q < 3; q++) {
            zload |= (g.pml & ZMIN) && kg0 < zs0[q] + npml;
            zload |= (g.pml & ZMAX) && kg1 > ((zs0[q] + zlen[q] - npml) >> 2 << 2);
        }
        h[H_ZLOAD] = zload ? 1 : 0;
    }
}

// The tile sequence of a CTA: t = blockIdx.x, blockIdx.x + gridDim.x, ... < ntiles, n = how many tiles came before (stage and
// phase bookkeeping).  The CPU emulation calls producer / consumer once per tile and passes the range (t_first, t_step, t_end, n0).
#ifdef GPI_HOST_EMU
#define T3_TILE_LOOP(t, n) int n = n0; for (int t = t_first; t < t_end; t += t_step, n++)
#define T3_NEXT_TILE(t) ((t) + t_step)
struct I4 { int x, y, z, w; };
inline I4 ldrec(const TileRec* r, int q) { return I4{r->v[4 * q], r->v[4 * q + 1], r->v[4 * q + 2], r->v[4 * q + 3]}; }
inline void sts_i4(int* p, const I4& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
#else
#define T3_TILE_LOOP(t, n) int n = 0; for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++)
#define T3_NEXT_TILE(t) ((t) + (int)gridDim.x)
typedef int4 I4;
__device__ __forceinline__ I4 ldrec(const TileRec* r, int q) { return __ldg(reinterpret_cast<const int4*>(r->v) + q); }
__device__ __forceinline__ void sts_i4(int* p, const I4& v) { *reinterpret_cast<int4*>(p) = v; }
#endif
// producer P of NPROD: waits for the stage, (P == 0: writes the tile header,) announces the bytes of ITS boxes and issues them.
// The full barrier of a stage counts NPROD arrivals plus every announced byte.
template <int KIND, int P>
__device__ __forceinline__ void producer(const Geom& g, const Sched& sc, const Maps* tm, float* stage0,
                                         sptr_t full0, sptr_t empty0, int t_first, int t_step, int t_end, int n0) {
    constexpr int NBOX = K<KIND>::NBOX, MAIN = K<KIND>::MAIN, SFLOATS = K<KIND>::SFLOATS, NPROD = L<KIND>::NPROD;
    int main_bytes = 0;
#pragma unroll
    for (int b = 0; b < NBOX; b++) if (b % NPROD == P) { const BoxSpec bs = box_spec(KIND, b); main_bytes += bs.rows * (bs.halo ? PH : ZC) * 4; }
    I4 r0, r1, r2, r3;             // record of the tile about to be issued (loaded one tile ahead)
    {
#ifdef GPI_HOST_EMU
        const int tf = t_first;
#else
        const int tf = blockIdx.x;
#endif
        if (tf < sc.ntiles) { r0 = ldrec(sc.tiles + tf, 0); r1 = ldrec(sc.tiles + tf, 1); r2 = ldrec(sc.tiles + tf, 2); r3 = ldrec(sc.tiles + tf, 3); }
    }
    T3_TILE_LOOP(t, n) {
        const int s = n % STAGES;
        const uint32_t use = n / STAGES;
        const I4 c0 = r0, c1 = r1, c2 = r2, c3 = r3;
        {
            const int tn = T3_NEXT_TILE(t);
            if (tn < sc.ntiles) { r0 = ldrec(sc.tiles + tn, 0); r1 = ldrec(sc.tiles + tn, 1); r2 = ldrec(sc.tiles + tn, 2); r3 = ldrec(sc.tiles + tn, 3); }
        }
        // record layout: H_I H_J0 H_KC0 sx0 | sx1 sx2 sy0 sy1 | sy2 ym0 ym1 ym2 | zload
        const int i = c0.x, j0 = c0.y, kc0 = c0.z;
        const int sx[3] = {c0.w, c1.x, c1.y}, sy0[3] = {c1.z, c1.w, c2.x}, ymask[3] = {c2.y, c2.z, c2.w};
        const bool zload = c3.x != 0;
        int bytes = main_bytes;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if ((NBOX + q) % NPROD == P && sx[q] >= 0) bytes += R * ZC * 4;
            if ((NBOX + 3 + q) % NPROD == P && ymask[q]) bytes += R * ZC * 4;
            if ((NBOX + 6 + q) % NPROD == P && zload) bytes += R * PZM * 4;
        }
        mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
        float* S = stage0 + (size_t)s * SFLOATS;
        if (P == 0) {
            int* hdr = reinterpret_cast<int*>(S + MAIN + P_HDR);
            sts_i4(hdr, c0); sts_i4(hdr + 4, c1); sts_i4(hdr + 8, c2); sts_i4(hdr + 12, c3);
        }
        const sptr_t bar = full0 + 8 * s;
        mbar_expect_tx(bar, bytes);                 // release: the header is visible to whoever observes the phase
        const sptr_t dst0 = s32(S);
#pragma unroll
        for (int b = 0; b < NBOX; b++) if (b % NPROD == P) {
            const BoxSpec bs = box_spec(KIND, b);
            tma_box(dst0 + bs.off * 4, tm->m[b], kc0 - (bs.halo ? 4 : 0), j0 + bs.dj, i + bs.di, bar);
        }
        const sptr_t dstp = dst0 + MAIN * 4;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if ((NBOX + q) % NPROD == P && sx[q] >= 0)  tma_box(dstp + (P_X + q * B4) * 4, tm->m[NBOX + q], kc0, j0, sx[q], bar);          // [k, j, s]
            if ((NBOX + 3 + q) % NPROD == P && ymask[q]) tma_box(dstp + (P_Y + q * B4) * 4, tm->m[NBOX + 3 + q], kc0, sy0[q], i, bar);      // [k, s, i]
            if ((NBOX + 6 + q) % NPROD == P && zload)    tma_box(dstp + (P_Z + q * R * PZM) * 4, tm->m[NBOX + 6 + q], 0, j0, i, bar);       // [zi, j, i]
        }
    }
}
template <int KIND>
__device__ __forceinline__ void producer_any(int p, const Geom& g, const Sched& sc, const Maps* tm, float* stage0,
                                             sptr_t full0, sptr_t empty0, int t_first, int t_step, int t_end, int n0) {
    constexpr int NPROD = L<KIND>::NPROD;
    if (p == 0) producer<KIND, 0>(g, sc, tm, stage0, full0, empty0, t_first, t_step, t_end, n0);
    else if (NPROD > 1 && p == 1) producer<KIND, 1 % NPROD>(g, sc, tm, stage0, full0, empty0, t_first, t_step, t_end, n0);
    else if (NPROD > 2 && p == 2) producer<KIND, 2 % NPROD>(g, sc, tm, stage0, full0, empty0, t_first, t_step, t_end, n0);
    else if (NPROD > 3 && p == 3) producer<KIND, 3 % NPROD>(g, sc, tm, stage0, full0, empty0, t_first, t_step, t_end, n0);
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

// which CPML terms an output group touches (AXIS: 0 z, 1 y, 2 x; Q: position in the axis' term list, see term_index)
//   velocity: group C = component C, whose x, y and z derivative are term C of every axis
//   stress:   group 0 = tauxx, tauyy, tauzz (the three normal derivatives, Q = 0 on every axis);
//             group 1 = tauxz, tauyz (d_zi vx, d_xi vz | d_zi vy, d_yi vz);  group 2 = tauxy (d_yi vx, d_xi vy)
__host__ __device__ constexpr bool group_uses(int kind, int comp, int axis, int q) {
    return comp == 3 ? true
         : kind == 0 ? q == comp
         : comp == 0 ? q == 0
         : comp == 1 ? (axis == 2 ? q == 2 : axis == 1 ? q == 2 : q >= 1)
         :             (axis != 0 && q == 1);
}
template <int KIND, int COMP>
__device__ __forceinline__ void open_tile(TileCtx& q, const Geom& g, const StepArgs& a, const float* S, int r, int lane) {
    const int* hdr = reinterpret_cast<const int*>(S + K<KIND>::MAIN + P_HDR);
    q.S = S; q.P = S + K<KIND>::MAIN; q.r = r; q.lane = lane;
    q.i = hdr[H_I]; q.j = hdr[H_J0] + r; q.k0 = hdr[H_KC0] + 4 * lane; q.kg0 = q.k0 + g.koff;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        if (group_uses(KIND, COMP, 2, t)) q.sx[t] = hdr[H_SX + t];
        if (group_uses(KIND, COMP, 1, t)) q.sy[t] = ((hdr[H_YMASK + t] >> r) & 1) ? hdr[H_SY0 + t] + r : -1;
    }
    q.zload = hdr[H_ZLOAD] != 0;
    q.c = uidx(g, q.k0, q.j, q.i);
#pragma unroll
    for (int t = 0; t < 3; t++) {
        if (group_uses(KIND, COMP, 2, t)) {
            const PmlTerm& tx = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, 2, t)];
            const int s = q.sx[t];
            if (s >= 0) { q.cx[t][0] = ld1(tx.a + s); q.cx[t][1] = ld1(tx.b + s); q.cx[t][2] = ld1(tx.kI + s); }
        }
        if (group_uses(KIND, COMP, 1, t)) {
            const PmlTerm& ty = (KIND == 0 ? a.pv : a.ps)[term_index(KIND, 1, t)];
            const int u = q.sy[t];
            if (u >= 0) { q.cy[t][0] = ld1(ty.a + u); q.cy[t][1] = ld1(ty.b + u); q.cy[t][2] = ld1(ty.kI + u); }
        }
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
        pml_z(q.P + P_Z + (Q * R + q.r) * PZM, t.mem + (long long)g.pzm * ((long long)q.j + (long long)g.ny1 * q.i),
              q.ZT + Q * 3 * PZM, z_slot(g, s0, len, q.kg0), d);
    }
}

// ------------------------------------------------------------------------------------------------
// velocity kernel (term order and citations: vel_cell in kernels.cuh); one warp = one component of one row
// ------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void vel_tile(const Geom& g, const StepArgs& a, const TileCtx& q) {
    const int nz = g.nz, k0 = q.k0, kg0 = q.kg0, r = q.r;
    const bool more = k0 + VW < g.pz;
    const float* B = q.S + 4 * q.lane;                    // plain boxes
    const float* H = q.S + 4 * q.lane + 4;                // boxes with a z halo
    F4 dx, dy, dz;                                        // the component's x, y, z derivative; v -= b * ((dx + dy) + dz)
    if (C == 0) {          // vx: dtauxxdx + dtauxydy + dtauxzdz
        const F4 xx = lds4(B + V_XX0 + r * ZC), xxm = lds4(B + V_XXM + r * ZC);
        const F4 xy = lds4(B + V_XY0 + r * ZC), xypy = lds4(B + V_XY0 + (r + 1) * ZC);
        const F4 xz = lds4(H + V_XZ0 + r * PH);
        const float xznext = z_next_s(q.mask, xz, H + V_XZ0 + r * PH, q.lane, more);
        dx = diff4(xx, xxm, g.dxI);            pml_t<0, 2, 0>(q, g, a, dx);
        dy = diff4(xypy, xy, g.dyI);           pml_t<0, 1, 0>(q, g, a, dy);
        dz = diff4_zp(xz, xznext, g.dzI);      pml_t<0, 0, 0>(q, g, a, dz);
    } else if (C == 1) {   // vy: dtauxydx + dtauyydy + dtauyzdz
        const F4 xy = lds4(B + V_XY0 + r * ZC), xypx = lds4(B + V_XYP + r * ZC);
        const F4 yy = lds4(B + V_YY + (r + 1) * ZC), yym = lds4(B + V_YY + r * ZC);
        const F4 yz = lds4(H + V_YZ + r * PH);
        const float yznext = z_next_s(q.mask, yz, H + V_YZ + r * PH, q.lane, more);
        dx = diff4(xypx, xy, g.dxI);           pml_t<0, 2, 1>(q, g, a, dx);
        dy = diff4(yy, yym, g.dyI);            pml_t<0, 1, 1>(q, g, a, dy);
        dz = diff4_zp(yz, yznext, g.dzI);      pml_t<0, 0, 1>(q, g, a, dz);
    } else {               // vz: dtauxzdx + dtauyzdy + dtauzzdz
        const F4 xz = lds4(H + V_XZ0 + r * PH), xzpx = lds4(B + V_XZP + r * ZC);
        const F4 yz = lds4(H + V_YZ + r * PH), yzpy = lds4(H + V_YZ + (r + 1) * PH);
        const F4 zz = lds4(H + V_ZZ + r * PH);
        const float zzprev = z_prev_s(q.mask, zz, H + V_ZZ + r * PH, q.lane, k0 > 0);
        dx = diff4(xzpx, xz, g.dxI);           pml_t<0, 2, 2>(q, g, a, dx);
        dy = diff4(yzpy, yz, g.dyI);           pml_t<0, 1, 2>(q, g, a, dy);
        dz = diff4_zm(zz, zzprev, g.dzI);      pml_t<0, 0, 2>(q, g, a, dz);
    }
    F4 nv = lds4(B + V_VX + C * B4 + r * ZC);
    const F4 b = lds4(B + V_BX + C * B4 + r * ZC);
    float* v = a.v[C == 0 ? V_X : C == 1 ? V_Y : V_Z] + q.c;
    const int Rg = g.rigid;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    const bool allown = k0 >= g.klo && k0 + VW - 1 <= g.khi;
    if (!head && !tail && allown) {
        // every cell of the group is an interior node of vx, vy and vz: no predicates
#pragma unroll
        for (int e = 0; e < VW; e++) nv.v[e] = __fsub_rn(nv.v[e], __fmul_rn(b.v[e], __fadd_rn(__fadd_rn(dx.v[e], dy.v[e]), dz.v[e])));
        st4(v, nv);
        return;
    }
    const int klast = C == 2 ? nz - 1 : nz - 2;           // last updated node: vx, vy inner in z; vz half
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
        if (own && k >= 1 && k <= klast) nv.v[e] = __fsub_rn(nv.v[e], __fmul_rn(b.v[e], __fadd_rn(__fadd_rn(dx.v[e], dy.v[e]), dz.v[e])));
    }
    // rigid z faces (dirichlet.jl:35-74); the x / y faces only touch shell rows (scalar path)
    if (head && (Rg & ZMIN)) nv.v[0] = C == 2 ? -nv.v[1] : 0.f;
    if (!tail && allown) { st4(v, nv); return; }
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
        if (own && k <= nz - 1) {
            const bool face = (Rg & ZMAX) && k == nz - 1;
            if (C == 2) { v[e] = nv.v[e]; if (face) v[e + 1] = -nv.v[e]; }       // vz[nz+1] = -vz[nz]
            else v[e] = face ? 0.f : nv.v[e];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stress kernel (term order and citations: stress_cell in kernels.cuh); one warp = one output group of one row
// ------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void stress_tile(const Geom& g, const StepArgs& a, const TileCtx& q) {
    const int nz = g.nz, k0 = q.k0, kg0 = q.kg0, r = q.r;
    const bool more = k0 + VW < g.pz;
    const float* B = q.S + 4 * q.lane;
    const float* H = q.S + 4 * q.lane + 4;
    const bool fs = (g.freesurf & ZMIN) != 0;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    const bool allown = k0 >= g.klo && k0 + VW - 1 <= g.khi;
    const bool plain = !head && !tail && allown;        // every cell is an interior node of all six stresses (k >= 4: no free-surface row)
    if (C == 0) {          // tauxx, tauyy, tauzz
        const F4 cvx = lds4(H + S_VX0 + (r + 1) * PH), vxpx = lds4(B + S_VXP + r * ZC);
        const F4 cvy = lds4(H + S_VY0 + r * PH), vypy = lds4(H + S_VY0 + (r + 1) * PH);
        const F4 cvz = lds4(H + S_VZ0 + (r + 1) * PH);
        const float vznext = z_next_s(q.mask, cvz, H + S_VZ0 + (r + 1) * PH, q.lane, more);
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
        st4(a.tau[T_XX] + q.c, xx); st4(a.tau[T_YY] + q.c, yy); st4(a.tau[T_ZZ] + q.c, zz);
    } else if (C == 1) {   // tauxz: z half, y inner, x half;  tauyz: z half, y half, x inner
        const F4 cvx = lds4(H + S_VX0 + (r + 1) * PH), cvy = lds4(H + S_VY0 + r * PH), cvz = lds4(H + S_VZ0 + (r + 1) * PH);
        const F4 vzmx = lds4(B + S_VZM + r * ZC), vzmy = lds4(H + S_VZ0 + r * PH);
        const float vxprev = z_prev_s(q.mask, cvx, H + S_VX0 + (r + 1) * PH, q.lane, k0 > 0);
        const float vyprev = z_prev_s(q.mask, cvy, H + S_VY0 + r * PH, q.lane, k0 > 0);
        F4 dxz = diff4_zm(cvx, vxprev, g.dzI);      pml_t<1, 0, 1>(q, g, a, dxz);      // @d_zi(vx)
        F4 dzx = diff4(cvz, vzmx, g.dxI);           pml_t<1, 2, 2>(q, g, a, dzx);      // @d_xi(vz)
        F4 dyz = diff4_zm(cvy, vyprev, g.dzI);      pml_t<1, 0, 2>(q, g, a, dyz);      // @d_zi(vy)
        F4 dzy = diff4(cvz, vzmy, g.dyI);           pml_t<1, 1, 2>(q, g, a, dzy);      // @d_yi(vz)
        F4 xz = lds4(B + S_XZ + r * ZC), yz = lds4(B + S_YZ + r * ZC);
        const F4 muxz = lds4(B + S_MUXZ + r * ZC), muyz = lds4(B + S_MUYZ + r * ZC);
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
        }
        st4(a.tau[T_XZ] + q.c, xz); st4(a.tau[T_YZ] + q.c, yz);
    } else {               // tauxy: z inner, y half, x half
        const F4 cvx = lds4(H + S_VX0 + (r + 1) * PH), vxmy = lds4(H + S_VX0 + r * PH);
        const F4 cvy = lds4(H + S_VY0 + r * PH), vymx = lds4(B + S_VYM + r * ZC);
        F4 dxy = diff4(cvx, vxmy, g.dyI);           pml_t<1, 1, 1>(q, g, a, dxy);      // @d_yi(vx)
        F4 dyx = diff4(cvy, vymx, g.dxI);           pml_t<1, 2, 1>(q, g, a, dyx);      // @d_xi(vy)
        F4 xy = lds4(B + S_XY + r * ZC);
        const F4 muxy = lds4(B + S_MUXY + r * ZC);
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = plain || (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (plain || (own && k >= 1 && k <= nz - 2)) xy.v[e] = __fsub_rn(xy.v[e], __fmul_rn(muxy.v[e], __fadd_rn(dxy.v[e], dyx.v[e])));
        }
        st4(a.tau[T_XY] + q.c, xy);
    }
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
// The outer shell of the box (rigid faces, ghost cells, ragged x / y ranges: every row of the planes sc.sp[], the rows sc.sr[] of the
// fast planes; 2.4 % of the cells at C3) is the scalar reference-order code.  It touches cells no tile touches and reads only fields
// this half step does not write, so it needs no ordering against the tiles: NSHELL warps of every CTA walk the shell lines beside the
// tile pipeline, in issue slots the consumers leave idle (as a separate launch it could not become resident next to two tile CTAs that
// fill the register file, and ran as a 40 - 80 us tail).  line -> (plane, row): as k_shell3.
__device__ __forceinline__ void shell_line_ji(const Geom& g, const Sched& sc, int L, int& i, int& j) {
    if (L < sc.nsp * g.ny1) { const int q = L / g.ny1; i = sc.sp[q]; j = L - q * g.ny1; }
    else { const int L2 = L - sc.nsp * g.ny1; const int q = L2 / sc.nsr; i = sc.ilo + q; j = sc.sr[L2 - q * sc.nsr]; }
}
template <int KIND>
__device__ __forceinline__ void shell_lines(const Geom& g, const StepArgs& a, const Sched& sc, int first, int step, int lane) {
    const int nlines = sc.nsp * g.ny1 + (sc.ihi - sc.ilo + 1) * sc.nsr;
    for (int L = first; L < nlines; L += step) {
        int i, j;
        shell_line_ji(g, sc, L, i, j);
        for (int k = lane; k < g.pz; k += 32) if (k >= g.klo && k <= g.khi) {
            if (KIND == 0) vel_cell<3, 1>(g, a, k, j, i, 0);
            else           stress_cell<3, 1>(g, a, k, j, i, 0);
        }
    }
}
// shell unit u = 32 consecutive z cells of a shell line (SHELLC: one unit per consumer warp every SHELL_EVERY tiles, the rest after the last tile)
__device__ __forceinline__ int shell_units(const Geom& g, const Sched& sc) { return (sc.nsp * g.ny1 + (sc.ihi - sc.ilo + 1) * sc.nsr) * ((g.pz + 31) >> 5); }
template <int KIND>
__device__ __forceinline__ void shell_unit(const Geom& g, const StepArgs& a, const Sched& sc, int u, int lane) {
    const int nzq = (g.pz + 31) >> 5;
    const int L = u / nzq, k = (u - L * nzq) * 32 + lane;
    int i, j;
    shell_line_ji(g, sc, L, i, j);
    if (k < g.pz && k >= g.klo && k <= g.khi) {
        if (KIND == 0) vel_cell<3, 1>(g, a, k, j, i, 0);
        else           stress_cell<3, 1>(g, a, k, j, i, 0);
    }
}
// consumer warp `warp` of CTA `cta` (of nctas): its q-th unit
template <int KIND>
__device__ __forceinline__ int shell_unit_of(int cta, int nctas, int warp, int q) { return cta * L<KIND>::NCW + warp + q * nctas * L<KIND>::NCW; }
// units left after a CTA's ntl tiles (ntl tiles gave warp w a turn at n % SHELL_EVERY == w)
template <int KIND>
__device__ __forceinline__ void shell_tail(const Geom& g, const StepArgs& a, const Sched& sc, int cta, int nctas, int ntl, int warp, int lane) {
    const int nu = shell_units(g, sc);
    int q = ntl > warp ? (ntl - warp + SHELL_EVERY - 1) / SHELL_EVERY : 0;
    for (int u = shell_unit_of<KIND>(cta, nctas, warp, q); u < nu; u = shell_unit_of<KIND>(cta, nctas, warp, ++q)) shell_unit<KIND>(g, a, sc, u, lane);
}

// one consumer lane of output group COMP, row r: the same tile sequence as the producer (see there for the range arguments)
template <int KIND, int COMP>
__device__ __forceinline__ void consumer_group(const Geom& g, const StepArgs& a, const Sched& sc, const float* stage0, const float* ZT,
                                               sptr_t full0, sptr_t empty0, int warp, int r, int lane, int t_first, int t_step, int t_end, int n0) {
    constexpr int SFLOATS = K<KIND>::SFLOATS;
    const int cta = blockIdx.x, nctas = gridDim.x;
    const int nu = L<KIND>::SHELLC ? shell_units(g, sc) : 0;
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
        open_tile<KIND, COMP>(q, g, a, S, r, lane);
        const bool active = q.j <= sc.jhi && q.k0 < g.pz && 4 * lane < ZC;
        q.mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            if (COMP == 3) {         // one warp, the three output groups in turn
                if (KIND == 0) { vel_tile<0>(g, a, q); vel_tile<1>(g, a, q); vel_tile<2>(g, a, q); }
                else { stress_tile<0>(g, a, q); stress_tile<1>(g, a, q); stress_tile<2>(g, a, q); }
            } else if (KIND == 0) vel_tile<COMP>(g, a, q); else stress_tile<COMP>(g, a, q);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (L<KIND>::SHELLC && n % SHELL_EVERY == warp) {          // this warp's turn: one shell unit in the slack before the next tile lands
            const int u = shell_unit_of<KIND>(cta, nctas, warp, n / SHELL_EVERY);
            if (u < nu) shell_unit<KIND>(g, a, sc, u, lane);
        }
    }
#ifndef GPI_HOST_EMU
    if (L<KIND>::SHELLC) shell_tail<KIND>(g, a, sc, cta, nctas, n, warp, lane);     // n = tiles this CTA had
#endif
}
// consumer warp w: row w / NCOMP, output group w % NCOMP (the four warps an SM sub-partition gets hold all three groups)
template <int KIND>
__device__ __forceinline__ void consumer(const Geom& g, const StepArgs& a, const Sched& sc, const float* stage0, const float* ZT,
                                         sptr_t full0, sptr_t empty0, int warp, int lane, int t_first, int t_step, int t_end, int n0) {
    constexpr int NCOMP = L<KIND>::NCOMP;
    const int r = warp / NCOMP, comp = warp - r * NCOMP;
    if (NCOMP == 1)     consumer_group<KIND, 3>(g, a, sc, stage0, ZT, full0, empty0, warp, r, lane, t_first, t_step, t_end, n0);
    else if (comp == 0) consumer_group<KIND, 0>(g, a, sc, stage0, ZT, full0, empty0, warp, r, lane, t_first, t_step, t_end, n0);
    else if (comp == 1) consumer_group<KIND, 1>(g, a, sc, stage0, ZT, full0, empty0, warp, r, lane, t_first, t_step, t_end, n0);
    else                consumer_group<KIND, 2>(g, a, sc, stage0, ZT, full0, empty0, warp, r, lane, t_first, t_step, t_end, n0);
}

#ifdef GPI_HOST_EMU
// CPU emulation of one CTA (run by the launch's "thread 0" of each block; the other 159 return): shared memory is a heap block,
// the tiles of the CTA are taken one after the other -- producer, then the NCW x 32 consumer lanes.
template <int KIND>
void k_step3t(const Geom g, const StepArgs a, const Sched sc, const Maps* tm) {
    if (threadIdx.x != 0) return;
    constexpr int SFLOATS = K<KIND>::SFLOATS, NPROD = L<KIND>::NPROD, NCW = L<KIND>::NCW, NSHELL = L<KIND>::NSHELL;
    const size_t nbytes = smem_bytes(KIND);
    unsigned char* smem_raw = static_cast<unsigned char*>(aligned_alloc(128, (nbytes + 127) / 128 * 128));
    memset(smem_raw, 0xff, nbytes);                       // shared memory starts as garbage (NaN patterns), not zeros
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SFLOATS * 4);
    const sptr_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    float* ZT = reinterpret_cast<float*>(bars + 2 * STAGES);
    fill_zt<KIND>(g, a, ZT, 0, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, NPROD); mbar_init(empty0 + 8 * s, NCW); }
    int n = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, n++) {
        for (int p = 0; p < NPROD; p++) producer_any<KIND>(p, g, sc, tm, stage0, full0, empty0, t, 1, t + 1, n);
        for (int warp = 0; warp < NCW; warp++) for (int lane = 0; lane < 32; lane++)
            consumer<KIND>(g, a, sc, stage0, ZT, full0, empty0, warp, lane, t, 1, t + 1, n);
    }
    for (int w = 0; w < NSHELL; w++) for (int lane = 0; lane < 32; lane++)
        shell_lines<KIND>(g, a, sc, blockIdx.x * NSHELL + w, gridDim.x * NSHELL, lane);
    if (L<KIND>::SHELLC) for (int warp = 0; warp < NCW; warp++) for (int lane = 0; lane < 32; lane++)
        shell_tail<KIND>(g, a, sc, blockIdx.x, gridDim.x, n, warp, lane);
    free(smem_raw);
}
#else
template <int KIND>
__global__ void __launch_bounds__(L<KIND>::NTHREADS, MINB) k_step3t(const Geom g, const StepArgs a, const Sched sc, const Maps* __restrict__ tm) {
    constexpr int SFLOATS = K<KIND>::SFLOATS, NPROD = L<KIND>::NPROD, NCW = L<KIND>::NCW, NSHELL = L<KIND>::NSHELL, NTHREADS = L<KIND>::NTHREADS;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((128 - (s32(smem_dyn) & 127)) & 127);      // boxes need 128-byte alignment
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SFLOATS * 4);
    const sptr_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    float* ZT = reinterpret_cast<float*>(bars + 2 * STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fill_zt<KIND>(g, a, ZT, threadIdx.x, NTHREADS);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, NPROD); mbar_init(empty0 + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCW + NPROD) { shell_lines<KIND>(g, a, sc, blockIdx.x * NSHELL + (warp - NCW - NPROD), gridDim.x * NSHELL, lane); return; }
    if (warp >= NCW) {
        if (lane == 0) producer_any<KIND>(warp - NCW, g, sc, tm, stage0, full0, empty0, 0, 0, 0, 0);
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

// what engine.cu needs of one variant
struct Tr {
    typedef Maps MapsT; typedef Sched SchedT; typedef TileRec TileRecT;
    static constexpr int ZC_ = ZC, PH_ = PH, R_ = R, MINB_ = MINB, PZM_ = PZM, V_NBOX_ = V_NBOX, S_NBOX_ = S_NBOX;
    static BoxSpec box(int kind, int b) { return box_spec(kind, b); }
    static int term(int kind, int axis, int q) { return term_index(kind, axis, q); }
    static size_t smem(int kind) { return smem_bytes(kind); }
    static void tiles(const Geom& g, const Sched& sc, int kind, TileRec* out) { fill_tile_table(g, sc, kind, out); }
    template <int KIND> struct Lay : L<KIND> {};
    typedef void (*TileFn)(const Geom, const StepArgs, const Sched, const Maps*);
    typedef void (*ShellFn)(const Geom, const StepArgs, const Sched);
    template <int KIND> static TileFn tile_kernel() { return k_step3t<KIND>; }
    template <int KIND> static ShellFn shell_kernel() { return k_shell3<KIND>; }
};

}  // namespace T3_NS
