// engine.cu -- C-ABI implementation (include/gpifdtd.h) of the B200 FDTD engine.
//
// Host side of the drop-in boundary for GeoPhyInv.jl's src/fdtd (citations relative to
// /root/reference): owns all device memory, converts the reference's per-field arrays to the
// unified-box layout described in kernels.cuh, and drives the time loop of mod_x_proc!
// (propagate.jl:138-261) as two fused stencil launches plus one small source/receiver launch per
// half step.  There is no CPU fallback: every entry point needs a CUDA device.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gpifdtd.h"
#include "kernels.cuh"

using namespace gpi;

namespace {

// ---- field metadata (fields.jl:92-671 reduced to node types, see kernels.cuh) -------------------
// per axis (z,y,x), O = order - 1, h = (order - 2) / 2; (length, storage offset of the array's first entry):
// 'I' tauii nodes (n, h), 'V' velocity nodes (n+O, 0), 'H' half (n-O, 1+2h), 'J' inner (n-2O, O+h)
const char* field_types(int f) {
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
    return nullptr;
}
int type_len(char t, int n, int order) { const int O = order - 1; return t == 'I' ? n : t == 'V' ? n + O : t == 'H' ? n - O : n - 2 * O; }
int type_off(char t, int order) { const int h = (order - 2) / 2; return t == 'I' ? h : t == 'V' ? 0 : t == 'H' ? 1 + 2 * h : 1 + 3 * h; }
bool has_y(int f) {
    switch (f) {
    case GPI_VY: case GPI_TAUYY: case GPI_TAUXY: case GPI_TAUYZ: case GPI_DPDY: case GPI_DVYDY:
    case GPI_DVXDY: case GPI_DVYDX: case GPI_DVYDZ: case GPI_DVZDY: case GPI_DTAUYYDY:
    case GPI_DTAUXYDX: case GPI_DTAUXYDY: case GPI_DTAUYZDY: case GPI_DTAUYZDZ: return true;
    }
    return false;
}
bool field_exists(int nd, int phys, int f) {
    if (f < 0 || f >= GPI_NFIELD || !field_types(f)) return false;
    if (nd == 2 && has_y(f)) return false;
    if (phys == GPI_ACOUSTIC) {
        if (f >= GPI_TAUXX && f <= GPI_TAUYZ) return false;
        if (f >= GPI_DVXDY && f <= GPI_DVZDY) return false;
        if (f >= GPI_DTAUXXDX) return false;
        return true;
    }
    return !(f == GPI_P || f == GPI_DPDX || f == GPI_DPDY || f == GPI_DPDZ);
}
int field_shape(int nd, int f, const int n[3], int out[3], int off[3], int order = 2) {
    const char* t = field_types(f);
    if (!t || (nd == 2 && has_y(f))) return 1;
    for (int q = 0; q < 3; q++) {
        if (q == 1 && nd == 2) { out[q] = 1; off[q] = 0; continue; }
        out[q] = type_len(t[q], n[q], order); off[q] = type_off(t[q], order);
    }
    return 0;
}
int dfield_axis(int f) {
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
// wavefield -> (is velocity?, slot)
bool wf_slot(int f, bool& isv, int& slot) {
    switch (f) {
    case GPI_P: case GPI_TAUXX: isv = false; slot = T_XX; return true;
    case GPI_TAUYY: isv = false; slot = T_YY; return true;
    case GPI_TAUZZ: isv = false; slot = T_ZZ; return true;
    case GPI_TAUXY: isv = false; slot = T_XY; return true;
    case GPI_TAUXZ: isv = false; slot = T_XZ; return true;
    case GPI_TAUYZ: isv = false; slot = T_YZ; return true;
    case GPI_VX: isv = true; slot = V_X; return true;
    case GPI_VY: isv = true; slot = V_Y; return true;
    case GPI_VZ: isv = true; slot = V_Z; return true;
    }
    return false;
}

// ---- per-(pw, shot, field) acquisition data ------------------------------------------------------
struct SparseDev {
    bool set = false;
    int ncol = 0, nnz = 0, nrows = 0;
    int *colptr = nullptr, *tap_cell = nullptr;   float* tap_val = nullptr;                 // receiver view (CSC order)
    int *row_cell = nullptr, *row_ptr = nullptr, *ent_col = nullptr;  float* ent_val = nullptr;   // injection view (row lists)
};
struct ShotData {
    SparseDev spray[GPI_NWAVEFIELD], interp[GPI_NWAVEFIELD];
    float* wav[GPI_NWAVEFIELD] = {};  int ns[GPI_NWAVEFIELD] = {};
    float* rec[GPI_NWAVEFIELD] = {};  int nr[GPI_NWAVEFIELD] = {};
    float* bnd[GPI_NWAVEFIELD][3] = {};          // [nt][slot] boundary planes (pw 1 only)
    float* snap[GPI_NWAVEFIELD] = {};            // final-state snapshots (pw 1 only), unified volumes
    std::vector<float*> usnaps;                  // user snapshots, unified volumes
};

struct Id128 { char b[128]; };     // ncclUniqueId is 128 opaque bytes, passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

}  // namespace

struct gpi_handle {
    gpi_config c;
    Geom g;
    int nd, el, npw, B;                 // B = shot batch
    int device;
    cudaStream_t stream = nullptr;  bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t side = nullptr;  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // the shell kernels of kernels3t.cuh run beside the tile kernels
    std::string err;

    // wavefields: W[b][pw][slot][vol]; TP same shape (adjoint only)
    int slot_tau[6], slot_v[3], nslots;
    float *W = nullptr, *TP = nullptr;
    long long pwstride, bstride;
    // CPML memory: one block per (b, pw) holding all terms
    struct Term { int dfield; int axis; long long off; long long size; bool vel; int idx; };
    std::vector<Term> terms;  long long mem_per_pw = 0;  float* MEM = nullptr;
    float* pmlcoef = nullptr;           // [GPI_NFIELD][3][2*npml]  (x / y terms: indexed by slab position)
    float* pmlztab = nullptr;           // [GPI_NFIELD][3][pz]      (z terms: indexed by the unified z coordinate)
    // medium
    float* mod[GPI_NPARAM] = {};        // unified volumes
    float* dmod[C_N] = {};  float** dmod_table = nullptr;
    // FD-Born (2-D acoustic): medium perturbation, scattering coefficients (d dtK, d bx, d bz), derivative scratch [B][2][vol]
    float* modp[GPI_NPARAM] = {};  float* born_c[3] = {};  float* born_d = nullptr;  bool born_ready = false;
    // gradients (every physics / dimensionality with npw = 2): total + per batch slot
    float* gtot[GPI_NPARAM] = {};  float* gshot = nullptr;  int ngrad = 0;  int gparam[3] = {};   // gshot[b][ngrad][vol]; gparam: parameter of each slot
    std::vector<ShotData> shots[2];
    std::vector<int32_t> itsnaps;
    PostDesc *post_v = nullptr, *post_s = nullptr, *h_post_v = nullptr, *h_post_s = nullptr;
    float** bnd_table = nullptr;        // boundary stores of the resident batch, [b][field][axis]
    // adjoint runs (order 2; float4 kernel families; default, GPI_PINGPONG=0 restores the copy): W and TP alternate as the time levels instead of save_tp!'s copy; the forced
    // boundary planes of the level that stays behind get their pre-force values back from `stash` ([b][field][axis], one slot)
    bool pingpong = true;  float* stash = nullptr;  float** stash_table = nullptr;
    float* stage = nullptr;  size_t stage_floats = 0;       // pinned host staging
    float* dscratch = nullptr;  size_t dscratch_floats = 0; // device scratch (raw interior medium before padding)
    gpi_timers timers{};
    std::vector<cudaEvent_t> evpool;  size_t evused = 0;   // sampled per-kernel timing (pairs)
    std::vector<int> evkind;                                // 0 = velocity kernel, 1 = stress kernel, 2 = z-slab halo exchange
    int sample_every = 16;
    // tuning
    dim3 blk3{64, 2, 2}, blk2{128, 2, 1};
    bool tma3 = true;  int num_sms = 148;  int tma3_ctas = 0;  bool tma3_force = false;   // GPI_TMA3=2: TMA kernels whatever the tile utilisation
    int shell_mode = 1;                                 // GPI_SHELL=0: shell kernel serialised behind the tile kernel (diagnostic)
    bool o4vec = true;                                  // order-4 kernels with four z cells per thread (kernels4v.cuh); GPI_O4VEC=0 selects the scalar ones
    // CUDA graphs for 2-D runs (forward, forward_save, adjoint; no slabs): the launches of a batch's time loop are captured once per run configuration and
    // replayed by later runs -- one graph launch instead of 3 - 4 host launches per time step, which is what bounds small 2-D grids
    // (C2, one resident shot: 30.5 -> 43.3 Gcell-updates/s; 8 shots: 91.6 -> 99.8).  GPI_GRAPH = 1 (default): the first run of a
    // configuration goes launch by launch (nothing to amortise yet; its kernel samples stay the timers' kernel figures), the second is
    // captured, later ones replay; 2: capture at the first run; 0: off.
    struct GraphEntry { unsigned long long key; void* exec; double launches; bool swapped; };      // swapped: the run ends on the other time-level set (ping-pong adjoint, odd nt)
    std::vector<GraphEntry> graphs;  int graph_mode = 1;  bool capturing = false;
    double kept_samples[4] = {0, 0, 0, 0};      // vel_ms, vel_n, stress_ms, stress_n of the last launch-by-launch run
    int o4by = 4;                                       // GPI_O4_BY: rows per block of the order-4 3-D kernels (1, 2, 4; planes = 4 / rows)
    int pzalign = 8;                                    // GPI_PZ_ALIGN (4, 8, 16, 32 floats)
    bool fuse2a = true;      // fused 2-D acoustic adjoint (kernels2a.cuh); GPI_FUSE2A=0 opts out
    // Programmatic dependent launch for the 2-D chain (k_vel2v, k_stress2v, k_stress2a, k_post, k_boundary; kernels.cuh: pdl_wait / pdl_release):
    // the launches carry the programmatic-stream-serialization attribute, so the next grid's CTAs are scheduled while the last wave of the
    // current one drains; captured into the time-loop graphs as programmatic edges.  C2 with 1 / 8 resident shots 42.9 -> 45.2 / 100.8 -> 105.1,
    // C4 87.7 -> 89.2 Gcell-updates/s, bit-identical (profiles/r02/ab_pdl.txt, tests/test_pdl_gpu.py).  GPI_PDL=0 opts out.
    bool pdl = true;
    bool illum_on = false;  double* illum_acc = nullptr;  double* illum_stack = nullptr;   // gpi_set_illum: [B][vol] per resident shot, [vol] stacked over the shots
    bool nvtx = false;       // GPI_NVTX=1: NVTX ranges around the phases of the entry points (a run, each resident batch, medium update, all-reduce) for nsys / ncu --nvtx
    void* t3_tiles[2] = {nullptr, nullptr};      // tile tables of the two tile kernels (geometry only: built once per handle)
    struct TmaSet { const float* key = nullptr; void* d[2] = {nullptr, nullptr}; } tmaps[2];  int tmap_victim = 0;   // TMA descriptors (device copies) per pw: [0] velocity, [1] stress kernel
    void* encode_tiled = nullptr;                                            // cuTensorMapEncodeTiled (driver entry point)   // 3-D elastic: TMA-pipelined persistent kernels (kernels3t.cuh); GPI_TMA3=0 selects k_*3v
    bool vec2 = true;                                   // 2-D: float4-per-thread kernels (kernels2v.cuh); GPI_SCALAR2D=1 selects the scalar ones
    int blkv = GPI_VEC_THREADS;  bool vec3 = true;      // 3-D: float4-per-thread kernels (kernels3d.cuh); GPI_SCALAR3D=1 selects the scalar ones
    // nccl
    NcclApi nccl;  void* comm = nullptr;  int rank = 0, nranks = 1;
    // z-slab domain decomposition
    bool slab = false;  int srank = 0, snranks = 1;  int ka = 0, kb = 0;      // owned global unified z range [ka, kb)
    int pzt = 0;                                                                // length of the k-indexed z coefficient tables
    float* halo_send[2] = {nullptr, nullptr};  float* halo_recv[2] = {nullptr, nullptr};   // [0] towards rank-1, [1] towards rank+1
    // pipelined exchange (GPI_SLAB_PIPE, default on): every stencil launch of a slab handle is split into two x halves; the halo planes of a
    // half travel on the side stream while the other half is computed (gpi_run).  xr_*: x range of the launch being issued (0 planes = all).
    bool slab_pipe = true;  int xr_lo = 0, xr_n = 0;
    cudaEvent_t ev_half[2] = {nullptr, nullptr}, ev_xv[2] = {nullptr, nullptr}, ev_xt[2] = {nullptr, nullptr};
};

static thread_local std::string g_create_err;      // per host thread: handles may be created from different threads

#define FAIL(h, ...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); (h)->err = _b; return 1; } while (0)
#define CU(h, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { char _b[512]; \
    snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); (h)->err = _b; return 1; } } while (0)
#define GUARD(h) do { if (!(h)) return 1; if (cudaSetDevice((h)->device) != cudaSuccess) { (h)->err = "cudaSetDevice failed"; return 1; } } while (0)

namespace {

float* wf_ptr(gpi_handle* h, float* base, int b, int ipw, int f) {
    bool isv; int slot;
    if (!wf_slot(f, isv, slot)) return nullptr;
    int s = isv ? h->slot_v[slot] : h->slot_tau[slot];
    if (s < 0) return nullptr;
    return base + (long long)b * h->bstride + (long long)ipw * h->pwstride + (long long)s * h->g.vol;
}

int ensure_stage(gpi_handle* h, size_t nfloats) {
    if (h->stage_floats >= nfloats) return 0;
    if (h->stage) cudaFreeHost(h->stage);
    h->stage = nullptr; h->stage_floats = 0;
    CU(h, cudaMallocHost((void**)&h->stage, nfloats * sizeof(float)));
    h->stage_floats = nfloats;
    return 0;
}

// host array in the field's own GLOBAL shape (column-major [z,(y),x]) <-> unified volume on the device.
// A slab handle keeps local z indices kl = k - koff for k in [koff, koff + pz): uploads copy every row it can
// hold (owned rows and the halo rows next to them), downloads return the owned rows and zero elsewhere.
// `k_first`/`nk` describe a z-window of the source: src holds global field rows k_first .. k_first+nk-1.
int upload_field(gpi_handle* h, int f, const float* src, float* dvol, int k_first = 0, int nk = -1) {
    const Geom& g = h->g;
    int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
    if (field_shape(h->nd, f, n, sh, off, h->c.order)) FAIL(h, "field %d has no shape in %d-D", f, h->nd);
    if (nk < 0) nk = sh[0];
    if (ensure_stage(h, (size_t)g.vol)) return 1;
    // the staging buffer starts from the device copy so that rows outside the window keep their values
    if (k_first != 0 || nk != sh[0]) {
        CU(h, cudaMemcpyAsync(h->stage, dvol, (size_t)g.vol * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    } else memset(h->stage, 0, (size_t)g.vol * sizeof(float));
    // field row iz (0-based in the global field array) sits at global unified k = iz + off[0]
    const int iz_lo = std::max(std::max(0, k_first), g.koff - off[0]);
    const int iz_hi = std::min(std::min(sh[0], k_first + nk), g.koff + g.pz - off[0]);      // exclusive
    if (iz_hi > iz_lo) for (int ix = 0; ix < sh[2]; ix++) for (int iy = 0; iy < sh[1]; iy++) {
        const float* sp = src + (size_t)nk * ((size_t)iy + (size_t)sh[1] * ix) + (iz_lo - k_first);
        float* d = h->stage + (iz_lo + off[0] - g.koff) + (size_t)g.pz * ((size_t)(iy + off[1]) + (size_t)g.ny1 * (ix + off[2]));
        memcpy(d, sp, (size_t)(iz_hi - iz_lo) * sizeof(float));
    }
    CU(h, cudaMemcpyAsync(dvol, h->stage, (size_t)g.vol * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}
template <typename T>
int download_field(gpi_handle* h, int f, const T* dvol, T* dst) {
    const Geom& g = h->g;
    int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
    if (field_shape(h->nd, f, n, sh, off, h->c.order)) FAIL(h, "field %d has no shape in %d-D", f, h->nd);
    if (ensure_stage(h, (size_t)g.vol * (sizeof(T) / sizeof(float)))) return 1;
    CU(h, cudaMemcpyAsync(h->stage, dvol, (size_t)g.vol * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    const int iz_lo = std::max(0, g.koff + g.klo - off[0]);
    const int iz_hi = std::min(sh[0], g.koff + g.khi + 1 - off[0]);                          // exclusive
    if (h->slab) memset(dst, 0, (size_t)sh[0] * sh[1] * sh[2] * sizeof(T));
    if (iz_hi > iz_lo) for (int ix = 0; ix < sh[2]; ix++) for (int iy = 0; iy < sh[1]; iy++) {
        T* d = dst + (size_t)sh[0] * ((size_t)iy + (size_t)sh[1] * ix) + iz_lo;
        const T* sp = (const T*)h->stage + (iz_lo + off[0] - g.koff) + (size_t)g.pz * ((size_t)(iy + off[1]) + (size_t)g.ny1 * (ix + off[2]));
        memcpy(d, sp, (size_t)(iz_hi - iz_lo) * sizeof(T));
    }
    return 0;
}

void free_sparse(SparseDev& s) {
    cudaFree(s.colptr); cudaFree(s.tap_cell); cudaFree(s.tap_val);
    cudaFree(s.row_cell); cudaFree(s.row_ptr); cudaFree(s.ent_col); cudaFree(s.ent_val);
    s = SparseDev();
}

template <typename T>
int to_device(gpi_handle* h, T** d, const std::vector<T>& v) {
    *d = nullptr;
    CU(h, cudaMalloc((void**)d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CU(h, cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

dim3 grid_for(const gpi_handle* h, dim3 blk, int nbatch) {
    const Geom& g = h->g;
    if (h->nd == 3) {
        int ntx = (g.nx1 + blk.z - 1) / blk.z;
        return dim3((g.khi + 1 + blk.x - 1) / blk.x, (g.ny1 + blk.y - 1) / blk.y, ntx * nbatch);
    }
    return dim3((g.khi + 1 + blk.x - 1) / blk.x, (g.nx1 + blk.y - 1) / blk.y, nbatch);
}

// `merged`: both wavefields of every resident shot in ONE launch.  The wavefield sets are laid out [b][pw][slot] and the
// CPML memory [b][pw][term], so slot b' = b * npw + ipw of a launch with the per-pw strides is wavefield ipw of shot b.
void fill_args(gpi_handle* h, StepArgs& a, int ipw, int nbatch, bool merged = false, float* base = nullptr) {
    memset(&a, 0, sizeof a);
    if (!base) base = h->W;
    for (int s = 0; s < 6; s++) a.tau[s] = h->slot_tau[s] >= 0 ? base + (long long)ipw * h->pwstride + (long long)h->slot_tau[s] * h->g.vol : nullptr;
    for (int s = 0; s < 3; s++) a.v[s] = h->slot_v[s] >= 0 ? base + (long long)ipw * h->pwstride + (long long)h->slot_v[s] * h->g.vol : nullptr;
    for (int s = 0; s < C_N; s++) a.c[s] = h->dmod[s];
    const int np2 = 2 * h->c.npml;
    for (const auto& t : h->terms) {
        PmlTerm p;
        p.mem = h->MEM + (long long)ipw * h->mem_per_pw + t.off;
        if (t.axis == 0 && h->c.order == 2) {
            p.a = h->pmlztab + ((size_t)t.dfield * 3 + 0) * h->pzt;
            p.b = h->pmlztab + ((size_t)t.dfield * 3 + 1) * h->pzt;
            p.kI = h->pmlztab + ((size_t)t.dfield * 3 + 2) * h->pzt;
        } else {
            p.a = h->pmlcoef + ((size_t)t.dfield * 3 + 0) * np2;
            p.b = h->pmlcoef + ((size_t)t.dfield * 3 + 1) * np2;
            p.kI = h->pmlcoef + ((size_t)t.dfield * 3 + 2) * np2;
        }
        p.bstride = merged ? h->mem_per_pw : (long long)h->npw * h->mem_per_pw;
        (t.vel ? a.pv : a.ps)[t.idx] = p;
    }
    a.wstride = merged ? h->pwstride : h->bstride;
    a.nbatch = merged ? nbatch * h->npw : nbatch;
}

// TMA descriptors of the operand boxes of kernels3t.cuh.  Every field / coefficient array is a rank-3 tensor
// (z, y, x) = (pz, ny1, nx1) with the unified-box strides, box = (128 or 136) x rows x 1; CPML memory arrays are
// (pz, ny1, 2 npml) for x terms, (pz, 2 npml, nx1) for y terms, (pzm, ny1, nx1) for z terms.  Out-of-range
// coordinates read zeros.
int encode_map(gpi_handle* h, void* out, const float* base, const cuuint64_t dims[3], cuuint32_t b0, cuuint32_t b1) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    cuuint64_t strides[2] = {dims[0] * 4, dims[0] * dims[1] * 4};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = ((EncodeFn)h->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base,
                                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) FAIL(h, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return 0;
}
template <class T>
int build_tmaps(gpi_handle* h, const StepArgs& a, int kind, void** dout) {
    const Geom& g = h->g;
    const int nbox = kind == 0 ? (int)T::V_NBOX_ : (int)T::S_NBOX_;
    typename T::MapsT hm;
    memset(&hm, 0, sizeof hm);
    const cuuint64_t fdims[3] = {(cuuint64_t)g.pz, (cuuint64_t)g.ny1, (cuuint64_t)g.nx1};
    for (int b = 0; b < nbox; b++) {
        const auto bs = T::box(kind, b);
        const float* base = bs.arr < 6 ? a.tau[bs.arr] : bs.arr < 9 ? a.v[bs.arr - 6] : a.c[bs.arr - 9];
        if (!base) FAIL(h, "TMA descriptor: operand %d of kernel %d is not allocated", bs.arr, kind);
        if (encode_map(h, hm.m[b], base, fdims, bs.halo ? T::PH_ : T::ZC_, bs.rows)) return 1;
    }
    const cuuint64_t np2 = 2 * (cuuint64_t)g.npml;
    const cuuint64_t xdims[3] = {(cuuint64_t)g.pz, (cuuint64_t)g.ny1, np2};
    const cuuint64_t ydims[3] = {(cuuint64_t)g.pz, np2, (cuuint64_t)g.nx1};
    const cuuint64_t zdims[3] = {(cuuint64_t)g.pzm, (cuuint64_t)g.ny1, (cuuint64_t)g.nx1};
    for (int q = 0; q < 3; q++) {
        const PmlTerm* terms = kind == 0 ? a.pv : a.ps;
        const float* mx = terms[T::term(kind, 2, q)].mem;
        const float* my = terms[T::term(kind, 1, q)].mem;
        const float* mz = terms[T::term(kind, 0, q)].mem;
        if (mx && (g.pml & (XMIN | XMAX)) && encode_map(h, hm.m[nbox + q], mx, xdims, T::ZC_, T::R_)) return 1;
        if (my && (g.pml & (YMIN | YMAX)) && encode_map(h, hm.m[nbox + 3 + q], my, ydims, T::ZC_, T::R_)) return 1;
        if (mz && (g.pml & (ZMIN | ZMAX)) && encode_map(h, hm.m[nbox + 6 + q], mz, zdims, T::PZM_, T::R_)) return 1;
    }
    if (!*dout) CU(h, cudaMalloc(dout, sizeof hm));
    CU(h, cudaMemcpyAsync(*dout, &hm, sizeof hm, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}
template <class T, int KIND>
int launch_step3t(gpi_handle* h, const StepArgs& a) {
    const Geom& g = h->g;
    // descriptors are cached per wavefield set, keyed by the set's first pointer (ping-pong runs swap W and TP: a new key rebuilds them)
    gpi_handle::TmaSet* set = nullptr;
    for (auto& ts : h->tmaps) if (ts.key == a.v[0]) set = &ts;
    if (!set) {
        set = &h->tmaps[h->tmap_victim]; h->tmap_victim ^= 1;      // two sets (pw 1, pw 2), replaced in turn
        if (build_tmaps<T>(h, a, 0, &set->d[0]) || build_tmaps<T>(h, a, 1, &set->d[1])) return 1;
        set->key = a.v[0];
    }
    typename T::SchedT sc{};
    if (KIND == 0) {
        sc.ilo = 2; sc.ihi = g.nx - 2; sc.jlo = 2; sc.jhi = g.ny - 2;
        sc.nsp = 4; sc.sp[0] = 0; sc.sp[1] = 1; sc.sp[2] = g.nx - 1; sc.sp[3] = g.nx;
        sc.nsr = 4; sc.sr[0] = 0; sc.sr[1] = 1; sc.sr[2] = g.ny - 1; sc.sr[3] = g.ny;
    } else {
        sc.ilo = 1; sc.ihi = g.nx - 2; sc.jlo = 1; sc.jhi = g.ny - 2;
        sc.nsp = 3; sc.sp[0] = 0; sc.sp[1] = g.nx - 1; sc.sp[2] = g.nx;
        sc.nsr = 3; sc.sr[0] = 0; sc.sr[1] = g.ny - 1; sc.sr[2] = g.ny;
    }
    sc.njb = (sc.jhi - sc.jlo + 1 + T::R_ - 1) / T::R_;
    sc.nzc = (g.pz + T::ZC_ - 1) / T::ZC_;
    sc.ntiles = (sc.ihi - sc.ilo + 1) * sc.njb * sc.nzc;
    if (!h->t3_tiles[KIND]) {
        std::vector<typename T::TileRecT> tab((size_t)sc.ntiles);
        T::tiles(g, sc, KIND, tab.data());
        CU(h, cudaMalloc(&h->t3_tiles[KIND], tab.size() * sizeof(typename T::TileRecT)));
        CU(h, cudaMemcpy(h->t3_tiles[KIND], tab.data(), tab.size() * sizeof(typename T::TileRecT), cudaMemcpyHostToDevice));
    }
    sc.tiles = static_cast<const typename T::TileRecT*>(h->t3_tiles[KIND]);
    const int nctas = std::max(1, std::min(sc.ntiles, h->tma3_ctas > 0 ? h->tma3_ctas : T::MINB_ * h->num_sms));
    typedef typename T::template Lay<KIND> Lay;
    const auto k_tile = T::template tile_kernel<KIND>();
    const auto k_shell = T::template shell_kernel<KIND>();
    const typename T::MapsT* maps = static_cast<const typename T::MapsT*>(set->d[KIND]);
    const int nlines = sc.nsp * g.ny1 + (sc.ihi - sc.ilo + 1) * sc.nsr;
    if (Lay::INKERNEL_SHELL || h->shell_mode == 2) {      // the shell lines are walked by warps of the tile kernel itself (GPI_SHELL=2 with a NSHELL = 0 build: no shell at all, timing only)
        k_tile<<<nctas, Lay::NTHREADS, T::smem(KIND), h->stream>>>(g, a, sc, maps);
        return 0;
    }
    // NSHELL = 0 builds: the shell is a launch of its own.  It touches cells no tile touches and reads only fields this half step
    // does not write, so it runs on a side stream; the tile kernel is launched FIRST: its tiles are assigned statically, so every
    // CTA must become resident at once.
    if (h->shell_mode == 0) {          // GPI_SHELL=0 (diagnostic): shell after the tiles on the same stream
        k_tile<<<nctas, Lay::NTHREADS, T::smem(KIND), h->stream>>>(g, a, sc, maps);
        k_shell<<<nlines, 128, 0, h->stream>>>(g, a, sc);
        h->timers.launches += 1;
        return 0;
    }
    CU(h, cudaEventRecord(h->ev_fork, h->stream));
    k_tile<<<nctas, Lay::NTHREADS, T::smem(KIND), h->stream>>>(g, a, sc, maps);
    CU(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    k_shell<<<nlines, 128, 0, h->side>>>(g, a, sc);
    CU(h, cudaEventRecord(h->ev_join, h->side));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->timers.launches += 1;
    return 0;
}
// Pipelined exchange (x-split launches of k_*3v, halos on the side stream) or serial exchange with whatever kernel family fits: every
// rank must decide alike (the NCCL calls have to match), so the rule uses global quantities only -- the AVERAGE slab height against the
// tile-utilisation threshold of tma3_eligible.  Slabs tall enough for the TMA tiles (C5 on 2 GPUs) keep them and the serial exchange
// (4 % of the step there); narrower ones (4, 8 GPUs) run the register-staged kernels anyway and hide the exchange behind them.
bool slab_pipelined(const gpi_handle* h) {
    if (!(h->slab && h->slab_pipe && h->nd == 3 && h->c.order == 2 && h->vec3 && h->snranks > 1)) return false;
    if (!(h->el && h->tma3) || h->tma3_force) return !h->tma3_force;
    const int avg = (h->g.nz + h->snranks) / h->snranks, chunks = (avg + 2 + t3::ZC - 1) / t3::ZC;
    return 4 * avg < 3 * chunks * t3::ZC;
}
bool tma3_eligible(const gpi_handle* h) {
    const Geom& g = h->g;
    if (slab_pipelined(h)) return false;        // every rank must take the same path (the NCCL calls have to match): the x-split launches are k_*3v
    const int zc = t3::ZC;
    const int zext = g.khi - g.klo + 1, zchunks = (g.pz + zc - 1) / zc;
    const bool tiles_fill = h->tma3_force || 4 * zext >= 3 * zchunks * zc;
    return h->nd == 3 && h->el && h->c.order == 2 && h->vec3 && h->tma3 && tiles_fill && g.pzm == t3::PZM && g.nx >= 2 * g.npml + 8 && g.ny >= 2 * g.npml + 8;
}
template <int EL>
void launch_step_kernels3v(gpi_handle* h, const StepArgs& a, bool vel, int nbatch) {
    const Geom& g = h->g;
    // The TMA tiles are 128 z cells wide: a narrow z-slab window (4 or 8 slabs of C5) would leave most lanes of its
    // last chunk idle, while the register-staged kernels linearise (z, y) and waste nothing -- they take over below
    // 75 % tile utilisation (C3: 339 of 384 = 0.88 -> TMA; C5 on 4 GPUs: 150 of 256 = 0.59 -> k_*3v).
    const bool oop = vel ? a.v_o[V_X] != nullptr : a.tau_o[T_XX] != nullptr;       // ping-pong adjoint runs: the register-staged kernels
    if (EL && nbatch == 1 && !oop && tma3_eligible(h)) {
        if ((vel ? launch_step3t<t3::Tr, 0>(h, a) : launch_step3t<t3::Tr, 1>(h, a)) == 0) return;
        h->tma3 = false;                        // descriptor creation failed: fall back to the register-staged kernels
    }
    const int ngroups = vec3_threads(g.pz, g.ny1);
    dim3 blk(h->blkv), grd((ngroups + h->blkv - 1) / h->blkv, g.nx1, nbatch);
    if (oop) {
        if (vel) k_vel3v<EL, 1><<<grd, blk, 0, h->stream>>>(g, a);
        else     k_stress3v<EL, 1><<<grd, blk, 0, h->stream>>>(g, a);
        return;
    }
    if (h->xr_n > 0) {          // one x half of a pipelined slab step
        Geom gx = g; gx.ioff = h->xr_lo; grd.y = h->xr_n;
        if (vel) k_vel3v<EL><<<grd, blk, 0, h->stream>>>(gx, a);
        else     k_stress3v<EL><<<grd, blk, 0, h->stream>>>(gx, a);
        return;
    }
    if (vel) k_vel3v<EL><<<grd, blk, 0, h->stream>>>(g, a);
    else     k_stress3v<EL><<<grd, blk, 0, h->stream>>>(g, a);
}
// NVTX range for the lifetime of the object (gpi_handle::nvtx); the profiler sees the phases the per-phase timers of gpi_get_timers count
struct Range {
    bool on;
    Range(const gpi_handle* h, const char* fmt, ...) : on(h && h->nvtx) {
        if (!on) return;
        char buf[160]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        nvtxRangePushA(buf);
    }
    ~Range() { if (on) nvtxRangePop(); }
    Range(const Range&) = delete;
};
// launch with the programmatic-stream-serialization attribute (gpi_handle::pdl; always off in the emulated engine)
#ifndef GPI_HOST_EMU
template <typename... P, typename... A>
static inline void launch_pdl(gpi_handle* h, void (*kern)(P...), dim3 grd, dim3 blk, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grd; cfg.blockDim = blk; cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);      // errors surface at the cudaGetLastError() that closes the batch
}
#define GPI_PDL(h, kern, grd, blk, ...) launch_pdl(h, kern, grd, blk, __VA_ARGS__)
#else
#define GPI_PDL(h, kern, grd, blk, ...) ((void)0)
#endif
template <int ND, int EL>
void launch_step_kernels(gpi_handle* h, const StepArgs& a, bool vel, int nbatch) {
    if (ND == 3 && h->vec3) { launch_step_kernels3v<EL>(h, a, vel, nbatch); return; }
    if (ND == 2 && h->vec2 && !a.dout[0]) {        // the derivative write-out of FD-Born lives in the scalar kernels
        const int nthreads = (h->g.pz / VW) * h->g.nx1;
        dim3 blk(128), grd((nthreads + 127) / 128, nbatch);
        if (vel ? a.v_o[V_X] != nullptr : a.tau_o[T_XX] != nullptr) {      // out of place (ping-pong adjoint runs)
            if (h->pdl) { if (vel) GPI_PDL(h, (k_vel2v<EL, 1>), grd, blk, h->g, a); else GPI_PDL(h, (k_stress2v<EL, 1>), grd, blk, h->g, a); return; }
            if (vel) k_vel2v<EL, 1><<<grd, blk, 0, h->stream>>>(h->g, a);
            else     k_stress2v<EL, 1><<<grd, blk, 0, h->stream>>>(h->g, a);
            return;
        }
        if (h->pdl) { if (vel) GPI_PDL(h, (k_vel2v<EL, 0>), grd, blk, h->g, a); else GPI_PDL(h, (k_stress2v<EL, 0>), grd, blk, h->g, a); return; }
        if (vel) k_vel2v<EL><<<grd, blk, 0, h->stream>>>(h->g, a);
        else     k_stress2v<EL><<<grd, blk, 0, h->stream>>>(h->g, a);
        return;
    }
    dim3 blk = ND == 3 ? h->blk3 : h->blk2;
    dim3 grd = grid_for(h, blk, nbatch);
    if (vel) k_vel<ND, EL><<<grd, blk, 0, h->stream>>>(h->g, a);
    else     k_stress<ND, EL><<<grd, blk, 0, h->stream>>>(h->g, a);
}
// one CUDA-event pair around a sampled launch; the pair's elapsed time is the kernel's duration because
// the stream is in-order (the first event completes when the preceding work has drained)
cudaEvent_t sample_event(gpi_handle* h) {
    if (h->evused == h->evpool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        h->evpool.push_back(e);
    }
    return h->evpool[h->evused++];
}
// order 4 (kernels4.cuh): the fused velocity kernel, then one face-sized launch per axis for the rigid faces in the
// reference's x, (y,) z order (dirichlet.jl:3-78; update_v!, advance_acou.jl:36-58); the fused stress kernel
template <int ND, int EL>
void launch_step_kernels4(gpi_handle* h, const StepArgs& a, bool vel, int nbatch) {
    const Geom& g = h->g;
    dim3 blk = ND == 3 ? h->blk3 : h->blk2;
    dim3 grd = grid_for(h, blk, nbatch);
    if (h->o4vec) {        // four z cells per thread (kernels4v.cuh)
        const int ngx = ((g.pz / 4) + 31) / 32;
        if (ND == 3) { const int by = h->o4by, bx = 4 / by; blk = dim3(32, by, bx); grd = dim3(ngx, (g.ny1 + by - 1) / by, ((g.nx1 + bx - 1) / bx) * nbatch); }
        else         { blk = dim3(32, 4, 1); grd = dim3(ngx, (g.nx1 + 3) / 4, nbatch); }
        if (!vel) { k_stress4v<ND, EL><<<grd, blk, 0, h->stream>>>(g, a); return; }
        k_vel4v<ND, EL><<<grd, blk, 0, h->stream>>>(g, a);
        blk = ND == 3 ? h->blk3 : h->blk2;
    } else {
        if (!vel) { k_stress4<ND, EL><<<grd, blk, 0, h->stream>>>(g, a); return; }
        k_vel4<ND, EL><<<grd, blk, 0, h->stream>>>(g, a);
    }
    const int nn[3] = {g.nz, g.ny, g.nx};
    for (int axis = 2; axis >= 0; axis--) {
        if (axis == 1 && ND == 2) continue;
        const int minbit = axis == 0 ? ZMIN : axis == 1 ? YMIN : XMIN;
        if (!(g.rigid & (minbit | (minbit << 1)))) continue;
        const int o1 = axis == 0 ? (ND == 3 ? 1 : 2) : 0;
        const int o2 = axis == 0 ? (ND == 3 ? 2 : -1) : (axis == 1 ? 2 : (ND == 3 ? 1 : -1));
        dim3 fb(128), fg((nn[o1] + 127) / 128, o2 >= 0 ? nn[o2] : 1, nbatch);
        k_dirichlet4<ND><<<fg, fb, 0, h->stream>>>(g, a, axis);
        h->timers.launches += 1;
    }
}
void launch_step(gpi_handle* h, const StepArgs& a, bool vel, int nbatch, bool sample = false, bool half = false) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (sample) { e0 = sample_event(h); e1 = sample_event(h); }
    if (e0 && e1) { cudaEventRecord(e0, h->stream); h->evkind.push_back((vel ? 0 : 1) + (half ? 3 : 0)); }      // 3 / 4: one x half of a pipelined slab step
    struct Closer { gpi_handle* h; cudaEvent_t e; ~Closer() { if (e) cudaEventRecord(e, h->stream); } } closer{h, (e0 && e1) ? e1 : nullptr};
    if (h->c.order == 4) {
        if (h->nd == 2 && !h->el) launch_step_kernels4<2, 0>(h, a, vel, nbatch);
        else if (h->nd == 2)      launch_step_kernels4<2, 1>(h, a, vel, nbatch);
        else if (!h->el)          launch_step_kernels4<3, 0>(h, a, vel, nbatch);
        else                      launch_step_kernels4<3, 1>(h, a, vel, nbatch);
        h->timers.launches += 1;
        return;
    }
    if (h->nd == 2 && !h->el) launch_step_kernels<2, 0>(h, a, vel, nbatch);
    else if (h->nd == 2)      launch_step_kernels<2, 1>(h, a, vel, nbatch);
    else if (!h->el)          launch_step_kernels<3, 0>(h, a, vel, nbatch);
    else                      launch_step_kernels<3, 1>(h, a, vel, nbatch);
    h->timers.launches += 1;
}

// boundary-stored fields (boundary.jl:113-264)
int boundary_fields(const gpi_handle* h, int out[6]) {
    if (!h->el) { out[0] = GPI_P; return 1; }
    if (h->nd == 3) {      // no upstream method (boundary.jl:215-264): the 2-D construction applied to all six stresses
        const int f[6] = {GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ, GPI_TAUXY, GPI_TAUXZ, GPI_TAUYZ};
        for (int q = 0; q < 6; q++) out[q] = f[q];
        return 6;
    }
    out[0] = GPI_TAUXX; out[1] = GPI_TAUXZ; out[2] = GPI_TAUZZ; return 3;
}
long long bnd_slot_floats(const gpi_handle* h, int axis) {
    const Geom& g = h->g; const int nb2 = 2 * h->c.nbound;
    if (axis == 2) return (long long)g.pz * g.ny1 * nb2;
    if (axis == 1) return (long long)g.pz * nb2 * g.nx1;
    return (long long)g.nx1 * g.ny1 * nb2;
}
// one launch for the whole batch: every stored field, axis and plane (k_boundary)
int launch_boundary(gpi_handle* h, int save /* 0 force, 1 save (negated), 2 copy out */, int nb, int slot /* 0-based time slot */,
                    float* base = nullptr, float* const* table = nullptr) {
    const Geom& g = h->g;
    BndArgs a; memset(&a, 0, sizeof a);
    int bf[6]; a.nf = boundary_fields(h, bf);
    a.nbound = h->c.nbound;
    a.naxes = 0;
    for (int q = 2; q >= 0; q--) if (!(q == 1 && h->nd == 2)) a.axes[a.naxes++] = q;
    int umax = 0, vmax = 0;
    for (int i = 0; i < a.nf; i++) {
        int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
        field_shape(h->nd, bf[i], n, sh, off, h->c.order);
        BndField& F = a.f[i];
        F.f0 = wf_ptr(h, base ? base : h->W, 0, 0, bf[i]);
        for (int ia = 0; ia < a.naxes; ia++) {
            const int axis = a.axes[ia];
            const int minbit = axis == 0 ? ZMIN : axis == 1 ? YMIN : XMIN;
            const int np = (h->c.pml_faces & minbit) ? h->c.npml : 0;   // min-face flag also used for the max side (boundary.jl:24,33)
            F.lo[axis] = np + off[axis]; F.hi[axis] = sh[axis] - np - a.nbound + off[axis];
            if (F.hi[axis] < F.lo[axis]) FAIL(h, "grid too small for the boundary store along axis %d", axis);
        }
        F.k0 = off[0]; F.j0 = off[1]; F.i0 = off[2];
        F.nk = off[0] + sh[0]; F.nj = off[1] + sh[1]; F.ni = off[2] + sh[2];
    }
    for (int ia = 0; ia < a.naxes; ia++) {
        const int axis = a.axes[ia];
        umax = std::max(umax, axis == 0 ? g.nx1 : g.pz);
        vmax = std::max(vmax, axis == 1 ? g.nx1 : g.ny1);
        a.slot_off[axis] = (long long)slot * bnd_slot_floats(h, axis);
    }
    a.stores = table ? table : h->bnd_table;
    a.wstride = h->bstride;
    dim3 blk(128), grd((umax + 127) / 128, vmax, 2 * a.nbound * a.naxes * a.nf * nb);
    if (h->pdl && h->nd == 2) {
        if (save == 2)  GPI_PDL(h, k_boundary<2>, grd, blk, g, a);
        else if (save)  GPI_PDL(h, k_boundary<1>, grd, blk, g, a);
        else            GPI_PDL(h, k_boundary<0>, grd, blk, g, a);
    }
    else if (save == 2)  k_boundary<2><<<grd, blk, 0, h->stream>>>(g, a);
    else if (save)  k_boundary<1><<<grd, blk, 0, h->stream>>>(g, a);
    else            k_boundary<0><<<grd, blk, 0, h->stream>>>(g, a);
    h->timers.launches += 1;
    return 0;
}
// device table of the boundary stores of the batch's shots: [b][field][axis]
int build_bnd_table(gpi_handle* h, int shot0, int nb) {
    int bf[6]; const int nbf = boundary_fields(h, bf);
    std::vector<float*> t((size_t)nb * nbf * 3, nullptr);
    for (int b = 0; b < nb; b++) for (int i = 0; i < nbf; i++) for (int q = 0; q < 3; q++)
        t[((size_t)b * nbf + i) * 3 + q] = h->shots[0][shot0 + b].bnd[bf[i]][q];
    if (!h->bnd_table) CU(h, cudaMalloc((void**)&h->bnd_table, (size_t)h->B * 18 * sizeof(float*)));
    CU(h, cudaMemcpyAsync(h->bnd_table, t.data(), t.size() * sizeof(float*), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));      // `t` is a pageable temporary
    return 0;
}

// ping-pong adjoint runs: one store slot per (batch slot, stored field, axis) for the pre-force values of the forced planes
int ensure_stash(gpi_handle* h) {
    if (h->stash_table) return 0;
    int bf[6]; const int nbf = boundary_fields(h, bf);
    long long per_field = 0, off[3] = {0, 0, 0};
    for (int q = 0; q < 3; q++) { if (q == 1 && h->nd == 2) continue; off[q] = per_field; per_field += (bnd_slot_floats(h, q) + 31) / 32 * 32; }
    CU(h, cudaMalloc((void**)&h->stash, (size_t)h->B * nbf * per_field * sizeof(float)));
    CU(h, cudaMemset(h->stash, 0, (size_t)h->B * nbf * per_field * sizeof(float)));
    std::vector<float*> t((size_t)h->B * nbf * 3, nullptr);
    for (int b = 0; b < h->B; b++) for (int i = 0; i < nbf; i++) for (int q = 0; q < 3; q++) {
        if (q == 1 && h->nd == 2) continue;
        t[((size_t)b * nbf + i) * 3 + q] = h->stash + ((size_t)b * nbf + i) * per_field + off[q];
    }
    CU(h, cudaMalloc((void**)&h->stash_table, t.size() * sizeof(float*)));
    CU(h, cudaMemcpy(h->stash_table, t.data(), t.size() * sizeof(float*), cudaMemcpyHostToDevice));
    return 0;
}

// Replicate-pad an un-extended medium array into the unified volume (media.jl:260-275: Pad(:replicate) on
// the PML faces): node (k, j, i) of the extended grid reads the interior node clamped to the array.
__global__ void k_pad_replicate(const Geom g, const float* __restrict__ src, float* __restrict__ dst,
                                int mz, int my, int mx, int lz, int ly, int lx) {
    const int kl = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, i = blockIdx.z;
    const int k = kl + g.koff;
    if (kl >= g.pz || k >= g.nz) return;
    const int sz = min(max(k - lz, 0), mz - 1), sy = min(max(j - ly, 0), my - 1), sx = min(max(i - lx, 0), mx - 1);
    dst[uidx(g, kl + g.h, g.ny1 > 1 ? j + g.h : j, i + g.h)] = src[(long long)sz + (long long)mz * ((long long)sy + (long long)my * sx)];
}

// The same padding fused with the derived-parameter broadcasts of media.jl:103-130 (invK!, invlambda!, invmu!, rho!: one scalar
// function per cell, Float32): the host hands over vp, (vs,) rho as they are and the independent parameters of the physics
// (medium.jl:81-95) come out on the extended grid.  inv(x) is one(x) / x; 2 * abs2(vs) is a Float32 product.
template <int EL>
__global__ void k_pad_derive(const Geom g, const float* __restrict__ vp, const float* __restrict__ vs, const float* __restrict__ rho,
                             float* __restrict__ m0 /* invK | invlambda */, float* __restrict__ m1 /* invmu */, float* __restrict__ mrho,
                             int mz, int my, int mx, int lz, int ly, int lx) {
    const int kl = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, i = blockIdx.z;
    const int k = kl + g.koff;
    if (kl >= g.pz || k >= g.nz) return;
    const int sz = min(max(k - lz, 0), mz - 1), sy = min(max(j - ly, 0), my - 1), sx = min(max(i - lx, 0), mx - 1);
    const long long s = (long long)sz + (long long)mz * ((long long)sy + (long long)my * sx);
    const long long d = uidx(g, kl + g.h, g.ny1 > 1 ? j + g.h : j, i + g.h);
    const float a = vp[s], r = rho[s];
    if (!EL) m0[d] = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(a, a), r));                                  // inv(abs2(vp) * rho)
    else {
        const float b2 = __fmul_rn(vs[s], vs[s]);
        m0[d] = __fdiv_rn(1.0f, __fmul_rn(__fsub_rn(__fmul_rn(a, a), __fmul_rn(b2, 2.0f)), r));       // inv((abs2(vp) - 2 * abs2(vs)) * rho)
        m1[d] = __fdiv_rn(1.0f, __fmul_rn(b2, r));                                                    // inv(abs2(vs) * rho)
    }
    mrho[d] = r;
}

__global__ void k_negate_copy(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = __fmul_rn(src[t], -1.0f);
}

}  // namespace

// =================================================================================================
// lifecycle
// =================================================================================================
extern "C" int gpi_abi_version(void) { return GPI_ABI_VERSION; }

extern "C" const char* gpi_last_error(const gpi_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int gpi_field_shape_order(int ndims, int physics, int order, int field_id, const int32_t n[3], int32_t out[3]) {
    if ((order != 2 && order != 4) || !field_exists(ndims, physics, field_id)) return 1;
    int nn[3] = {n[0], ndims == 3 ? n[1] : 1, n[2]}, o[3], off[3];
    if (field_shape(ndims, field_id, nn, o, off, order)) return 1;
    out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
    return 0;
}
extern "C" int gpi_field_shape(int ndims, int physics, int field_id, const int32_t n[3], int32_t out[3]) {
    return gpi_field_shape_order(ndims, physics, 2, field_id, n, out);
}

static int create_impl(gpi_handle* h) {
    const gpi_config& c = h->c;
    Geom& g = h->g;
    g.nz = c.n[0]; g.ny = h->nd == 3 ? c.n[1] : 1; g.nx = c.n[2];
    // z-slab window: global unified nodes k in [0, nz] split evenly; koff is a multiple of four so that
    // the vector kernels' global table / z-memory indices keep their 16-byte alignment
    g.h = (c.order - 2) / 2;                       // order 4: one extra node on the min side of every axis
    g.ioff = 0;
    g.koff = 0; g.klo = 0; g.khi = g.nz + 2 * g.h;
    h->ka = 0; h->kb = g.nz + 1;
    if (h->slab) {
        // Planes inside the z-CPML slabs move three extra memory variables per kernel (48 B on top of 140 B per cell and
        // step in 3-D elastic, 16 on top of 64 in 3-D acoustic), so the end ranks get proportionally fewer planes:
        // boundaries at equal cumulated weight, weight = 1 + alpha inside a z slab.
        const int nodes = g.nz + 1;
        const double alpha = h->el ? 48.0 / 140.0 : 16.0 / 64.0;
        std::vector<double> cum(nodes + 1, 0.0);
        for (int k = 0; k < nodes; k++) {
            const bool inslab = ((c.pml_faces & ZMIN) && k < c.npml) || ((c.pml_faces & ZMAX) && k >= nodes - 1 - c.npml);
            cum[k + 1] = cum[k] + 1.0 + (inslab ? alpha : 0.0);
        }
        auto cut = [&](int r) {
            if (r <= 0) return 0;
            if (r >= h->snranks) return nodes;
            const double target = cum[nodes] * r / h->snranks;
            return (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        };
        h->ka = cut(h->srank);
        h->kb = cut(h->srank + 1);
        if (h->kb - h->ka < 4) FAIL(h, "z-slab of rank %d has only %d planes", h->srank, h->kb - h->ka);
        g.koff = h->srank == 0 ? 0 : ((h->ka - 1) / 4) * 4;
        g.klo = h->ka - g.koff;
        g.khi = h->kb - 1 - g.koff;
    }
    {
        // z pitch: nodes 0..khi+1 rounded up to `pzalign` floats.  Every vector access needs 16 bytes (4 floats); the default keeps
        // whole 32-byte sectors per row (8 floats) -- rounding rows up to 128-byte lines costs 28 % of the traffic of a 75-plane
        // z-slab window (77 -> 96 floats) and 2.3 % at C3 (339 -> 352), and buys nothing: rows are contiguous in memory anyway
        int pzalign = h->pzalign;
        if (pzalign != 4 && pzalign != 8 && pzalign != 16 && pzalign != 32) pzalign = 8;
        g.pz = ((g.khi + 2 + pzalign - 1) / pzalign) * pzalign;
    }
    h->pzt = ((g.nz + 64 + 31) / 32) * 32;
    g.ny1 = h->nd == 3 ? g.ny + 1 + 2 * g.h : 1;
    g.nx1 = g.nx + 1 + 2 * g.h;
    g.npml = c.npml;
    g.pzm = ((2 * ((c.npml + 3 + 3) / 4 * 4) + 31) / 32) * 32;     // two float4-aligned halves (kernels.cuh, cpml<>)
    g.pml = c.pml_faces; g.rigid = c.rigid_faces; g.freesurf = c.stressfree_faces;
    g.dzI = (float)c.dI[0]; g.dyI = (float)c.dI[1]; g.dxI = (float)c.dI[2];
    g.vol = (long long)g.pz * g.ny1 * g.nx1;
    if (const char* e = getenv("GPI_VOL_PAD")) g.vol += (atoll(e) + 31) / 32 * 32;      // tuning: floats of padding between consecutive field volumes
    if (g.vol >= (1LL << 31)) FAIL(h, "grid of %lld unified cells exceeds the 32-bit cell index of the source/receiver tables", g.vol);
    // the min and max CPML slabs of a derivative field must not overlap (the reference would apply both)
    for (int q = 0; q < 3; q++) {
        if (q == 1 && h->nd == 2) continue;
        const int lo = q == 0 ? ZMIN : q == 1 ? YMIN : XMIN, hi = q == 0 ? ZMAX : q == 1 ? YMAX : XMAX;
        const int inner = c.n[q] - 2 * (c.order - 1);      // shortest derivative field along the axis
        if ((c.pml_faces & lo) && (c.pml_faces & hi) && inner < 2 * c.npml)
            FAIL(h, "axis %d: %d nodes cannot hold two %d-cell CPML slabs", q, c.n[q], c.npml);
        if ((c.pml_faces & (lo | hi)) && inner < c.npml) FAIL(h, "axis %d shorter than the CPML slab", q);
    }

    // wavefield slots
    for (int s = 0; s < 6; s++) h->slot_tau[s] = -1;
    for (int s = 0; s < 3; s++) h->slot_v[s] = -1;
    int ns = 0;
    if (!h->el) h->slot_tau[T_XX] = ns++;
    else {
        h->slot_tau[T_XX] = ns++; if (h->nd == 3) h->slot_tau[T_YY] = ns++; h->slot_tau[T_ZZ] = ns++;
        if (h->nd == 3) h->slot_tau[T_XY] = ns++;
        h->slot_tau[T_XZ] = ns++;
        if (h->nd == 3) h->slot_tau[T_YZ] = ns++;
    }
    h->slot_v[V_X] = ns++; if (h->nd == 3) h->slot_v[V_Y] = ns++; h->slot_v[V_Z] = ns++;
    h->nslots = ns;
    h->pwstride = (long long)ns * g.vol;
    h->bstride = (long long)h->npw * h->pwstride;

    // shot batch: 2-D problems are tiny and launch-bound, so several shots share every launch
    int B = c.shot_batch > 0 ? c.shot_batch : (h->nd == 2 ? 16 : 1);
    B = std::max(1, std::min(B, c.nshots));
    h->B = B;

    // CPML terms (cpml.jl:109-120: one memory array per derivative field)
    struct TD { int f; bool vel; int idx; };
    std::vector<TD> td;
    if (!h->el) {
        td = {{GPI_DPDX, true, 0}, {GPI_DPDZ, true, 2}, {GPI_DVXDX, false, 0}, {GPI_DVZDZ, false, 2}};
        if (h->nd == 3) { td.push_back({GPI_DPDY, true, 1}); td.push_back({GPI_DVYDY, false, 1}); }
    } else if (h->nd == 2) {
        td = {{GPI_DTAUXXDX, true, 0}, {GPI_DTAUXZDZ, true, 2}, {GPI_DTAUXZDX, true, 6}, {GPI_DTAUZZDZ, true, 8},
              {GPI_DVXDX, false, 0}, {GPI_DVZDZ, false, 2}, {GPI_DVXDZ, false, 5}, {GPI_DVZDX, false, 6}};
    } else {
        td = {{GPI_DTAUXXDX, true, 0}, {GPI_DTAUXYDY, true, 1}, {GPI_DTAUXZDZ, true, 2},
              {GPI_DTAUXYDX, true, 3}, {GPI_DTAUYYDY, true, 4}, {GPI_DTAUYZDZ, true, 5},
              {GPI_DTAUXZDX, true, 6}, {GPI_DTAUYZDY, true, 7}, {GPI_DTAUZZDZ, true, 8},
              {GPI_DVXDX, false, 0}, {GPI_DVYDY, false, 1}, {GPI_DVZDZ, false, 2},
              {GPI_DVXDY, false, 3}, {GPI_DVYDX, false, 4}, {GPI_DVXDZ, false, 5}, {GPI_DVZDX, false, 6},
              {GPI_DVYDZ, false, 7}, {GPI_DVZDY, false, 8}};
    }
    long long off = 0;
    for (auto& t : td) {
        gpi_handle::Term T;
        T.dfield = t.f; T.axis = dfield_axis(t.f); T.vel = t.vel; T.idx = t.idx; T.off = off;
        const int minbit = T.axis == 0 ? ZMIN : T.axis == 1 ? YMIN : XMIN, maxbit = minbit << 1;
        if (!(c.pml_faces & (minbit | maxbit))) T.size = 32;    // never touched
        else if (T.axis == 2) T.size = (long long)g.pz * g.ny1 * 2 * c.npml;
        else if (T.axis == 1) T.size = (long long)g.pz * 2 * c.npml * g.nx1;
        else T.size = (long long)g.pzm * g.ny1 * g.nx1;
        off += (T.size + 31) / 32 * 32;
        h->terms.push_back(T);
    }
    h->mem_per_pw = off;

    CU(h, cudaMalloc((void**)&h->W, (size_t)B * h->bstride * sizeof(float)));
    CU(h, cudaMemset(h->W, 0, (size_t)B * h->bstride * sizeof(float)));
    if (h->npw == 2) {
        CU(h, cudaMalloc((void**)&h->TP, (size_t)B * h->bstride * sizeof(float)));
        CU(h, cudaMemset(h->TP, 0, (size_t)B * h->bstride * sizeof(float)));
    }
    CU(h, cudaMalloc((void**)&h->MEM, (size_t)B * h->npw * h->mem_per_pw * sizeof(float)));
    CU(h, cudaMemset(h->MEM, 0, (size_t)B * h->npw * h->mem_per_pw * sizeof(float)));

    // CPML coefficients: a = b = 0, kI = 1 until update_pml! (cpml.jl:114)
    const int np2 = 2 * c.npml;
    std::vector<float> coef((size_t)GPI_NFIELD * 3 * np2, 0.f);
    for (int f = 0; f < GPI_NFIELD; f++) for (int i = 0; i < np2; i++) coef[((size_t)f * 3 + 2) * np2 + i] = 1.f;
    if (to_device(h, &h->pmlcoef, coef)) return 1;
    std::vector<float> ztab((size_t)GPI_NFIELD * 3 * h->pzt, 0.f);
    for (int f = 0; f < GPI_NFIELD; f++) for (int i = 0; i < h->pzt; i++) ztab[((size_t)f * 3 + 2) * h->pzt + i] = 1.f;
    if (to_device(h, &h->pmlztab, ztab)) return 1;

    // medium
    const size_t vb = (size_t)g.vol * sizeof(float);
    auto alloc_vol = [&](float** p) -> int { CU(h, cudaMalloc((void**)p, vb)); CU(h, cudaMemset(*p, 0, vb)); return 0; };
    if (alloc_vol(&h->mod[GPI_RHO])) return 1;
    if (!h->el) { if (alloc_vol(&h->mod[GPI_INVK])) return 1; }
    else { if (alloc_vol(&h->mod[GPI_INVLAMBDA]) || alloc_vol(&h->mod[GPI_INVMU])) return 1; }
    int need[C_N] = {1, h->nd == 3, 1, 1, h->el, h->el, h->el && h->nd == 3, h->el && h->nd == 3};
    for (int s = 0; s < C_N; s++) if (need[s]) { if (alloc_vol(&h->dmod[s])) return 1; }
    CU(h, cudaMalloc((void**)&h->dmod_table, C_N * sizeof(float*)));
    CU(h, cudaMemcpy(h->dmod_table, h->dmod, C_N * sizeof(float*), cudaMemcpyHostToDevice));

    // gradients exist upstream for acoustic media (fdtd.jl:164-170; imaging 2-D only, gradient.jl:31); here also 3-D acoustic
    // and 2-D elastic (kernels.cuh k_grad3d, k_grad2d_el)
    if (h->npw == 2) {
        if (!h->el) { h->ngrad = 2; h->gparam[0] = GPI_INVK; h->gparam[1] = GPI_RHO; }
        else        { h->ngrad = 3; h->gparam[0] = GPI_INVLAMBDA; h->gparam[1] = GPI_INVMU; h->gparam[2] = GPI_RHO; }
        for (int q = 0; q < h->ngrad; q++) if (alloc_vol(&h->gtot[h->gparam[q]])) return 1;
        CU(h, cudaMalloc((void**)&h->gshot, (size_t)B * h->ngrad * vb));
        CU(h, cudaMemset(h->gshot, 0, (size_t)B * h->ngrad * vb));
    }

    // per-shot state
    for (int ipw = 0; ipw < h->npw; ipw++) h->shots[ipw].resize(c.nshots);
    int bf[6]; const int nbf = boundary_fields(h, bf);
    for (int is = 0; is < c.nshots; is++) {
        ShotData& s = h->shots[0][is];
        if (c.store_boundary) {
            for (int i = 0; i < nbf; i++) {
                for (int q = 0; q < 3; q++) {
                    if (q == 1 && h->nd == 2) continue;
                    size_t nb = (size_t)bnd_slot_floats(h, q) * c.nt * sizeof(float);
                    CU(h, cudaMalloc((void**)&s.bnd[bf[i]][q], nb));
                    CU(h, cudaMemset(s.bnd[bf[i]][q], 0, nb));
                }
                if (alloc_vol(&s.snap[bf[i]])) return 1;
            }
            const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ};
            for (int i = 0; i < 3; i++) if (field_exists(h->nd, c.physics, vf[i])) { if (alloc_vol(&s.snap[vf[i]])) return 1; }
        }
        if (c.nsnaps > 0) for (int ipw = 0; ipw < h->npw; ipw++) {
            h->shots[ipw][is].usnaps.resize(c.nsnaps, nullptr);
            for (int k = 0; k < c.nsnaps; k++) if (alloc_vol(&h->shots[ipw][is].usnaps[k])) return 1;
        }
    }
    CU(h, cudaMalloc((void**)&h->post_v, (size_t)B * sizeof(PostDesc)));
    CU(h, cudaMalloc((void**)&h->post_s, (size_t)B * sizeof(PostDesc)));
    CU(h, cudaMallocHost((void**)&h->h_post_v, (size_t)B * sizeof(PostDesc)));
    CU(h, cudaMallocHost((void**)&h->h_post_s, (size_t)B * sizeof(PostDesc)));
    CU(h, cudaEventCreate(&h->ev0));
    CU(h, cudaEventCreate(&h->ev1));
    {
        // the side stream carries the halo exchange of pipelined z-slab steps: highest priority, so that its small pack / NCCL / unpack
        // kernels get their blocks placed as soon as SM resources free up instead of queueing behind the stencil kernel's whole grid
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CU(h, cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
    }
    CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    if (h->slab) for (int q = 0; q < 2; q++) {
        CU(h, cudaEventCreateWithFlags(&h->ev_half[q], cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_xv[q], cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_xt[q], cudaEventDisableTiming));
    }
    if (h->slab) for (int d = 0; d < 2; d++) {
        const size_t nb = (size_t)3 * g.ny1 * g.nx1 * sizeof(float);
        CU(h, cudaMalloc((void**)&h->halo_send[d], nb));
        CU(h, cudaMalloc((void**)&h->halo_recv[d], nb));
    }
    return 0;
}

extern "C" int gpi_create(const gpi_config* cfg, gpi_handle** out) {
    if (!cfg || !out) { g_create_err = "gpi_create: null argument"; return 1; }
    *out = nullptr;
    if (cfg->abi_version != GPI_ABI_VERSION) { g_create_err = "gpi_create: ABI version mismatch"; return 1; }
    if (cfg->order != 2 && cfg->order != 4) { g_create_err = "gpi_create: orders 2 and 4 are implemented (orders 6/8 are broken upstream)"; return 1; }
    if (cfg->npml != 40 + (cfg->order - 1)) { g_create_err = "gpi_create: npml must be 40 + (order - 1) (GeoPhyInv.jl:90)"; return 1; }
    if (cfg->order == 4 && cfg->slab_nranks > 1) { g_create_err = "gpi_create: z-slabs are implemented for order 2 (an order-4 cut needs two halo planes)"; return 1; }
    if (cfg->ndims != 2 && cfg->ndims != 3) { g_create_err = "gpi_create: ndims must be 2 or 3"; return 1; }
    if (cfg->physics != GPI_ACOUSTIC && cfg->physics != GPI_ELASTIC) { g_create_err = "gpi_create: unknown physics"; return 1; }
    if (cfg->npw < 1 || cfg->npw > 2 || cfg->nshots < 1 || cfg->nt < 1 || cfg->npml < 1 || cfg->nbound < 1 || cfg->nbound > 4) { g_create_err = "gpi_create: bad npw/nshots/nt/npml/nbound"; return 1; }
    if (cfg->n[0] < 4 || cfg->n[2] < 4 || (cfg->ndims == 3 && cfg->n[1] < 4)) { g_create_err = "gpi_create: grid too small"; return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_err = "gpi_create: no CUDA device (this engine has no CPU fallback)"; return 1; }
    gpi_handle* h = new gpi_handle();
    h->c = *cfg; h->nd = cfg->ndims; h->el = cfg->physics == GPI_ELASTIC; h->npw = cfg->npw;
    if (cfg->slab_nranks > 1) {
        if (cfg->ndims != 3 || cfg->npw != 1 || cfg->store_boundary || cfg->slab_rank < 0 || cfg->slab_rank >= cfg->slab_nranks) {
            g_create_err = "gpi_create: z-slabs need ndims = 3, npw = 1, no boundary store and 0 <= slab_rank < slab_nranks"; delete h; return 1;
        }
        h->slab = true; h->srank = cfg->slab_rank; h->snranks = cfg->slab_nranks;
    }
    if (cfg->device >= 0) h->device = cfg->device; else cudaGetDevice(&h->device);
    if (cudaSetDevice(h->device) != cudaSuccess) { g_create_err = "gpi_create: cudaSetDevice failed"; delete h; return 1; }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { g_create_err = "gpi_create: stream creation failed"; delete h; return 1; }
    h->own_stream = true;
    if (const char* e = getenv("GPI_SAMPLE_EVERY")) h->sample_every = atoi(e);
    if (const char* e = getenv("GPI_BLOCK3")) { int a, b, c3; if (sscanf(e, "%d,%d,%d", &a, &b, &c3) == 3 && a * b * c3 <= 256) h->blk3 = dim3(a, b, c3); }
    if (const char* e = getenv("GPI_SCALAR3D")) h->vec3 = atoi(e) == 0;
    if (const char* e = getenv("GPI_SCALAR2D")) h->vec2 = atoi(e) == 0;
    if (const char* e = getenv("GPI_TMA3")) { h->tma3 = atoi(e) != 0; h->tma3_force = atoi(e) == 2; }
    if (const char* e = getenv("GPI_TMA3_CTAS")) h->tma3_ctas = atoi(e);
    if (const char* e = getenv("GPI_SHELL")) h->shell_mode = atoi(e);
    if (const char* e = getenv("GPI_FUSE2A")) h->fuse2a = atoi(e) != 0;
    if (const char* e = getenv("GPI_NVTX")) h->nvtx = atoi(e) != 0;
#ifdef GPI_HOST_EMU
    h->pdl = false;          // the emulation runs one grid after the other through plain launches
#else
    if (const char* e = getenv("GPI_PDL")) h->pdl = atoi(e) != 0;
#endif
    if (const char* e = getenv("GPI_PZ_ALIGN")) h->pzalign = atoi(e);
    if (const char* e = getenv("GPI_SLAB_PIPE")) h->slab_pipe = atoi(e) != 0;
    if (const char* e = getenv("GPI_GRAPH")) h->graph_mode = atoi(e);
    if (const char* e = getenv("GPI_O4_BY")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) h->o4by = v; }
    if (const char* e = getenv("GPI_O4VEC")) h->o4vec = atoi(e) != 0;
    if (const char* e = getenv("GPI_PINGPONG")) h->pingpong = atoi(e) != 0;
    if (h->nd == 3 && h->el) {
        cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
        cudaError_t e0 = cudaFuncSetAttribute(t3::k_step3t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t3::smem_bytes(0));
        cudaError_t e1 = cudaFuncSetAttribute(t3::k_step3t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t3::smem_bytes(1));
        cudaFuncSetAttribute(t3::k_step3t<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(t3::k_step3t<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e2 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres);
        if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess || !h->encode_tiled) { h->tma3 = false; cudaGetLastError(); }
    }
    if (const char* e = getenv("GPI_BLOCKV")) { int a = atoi(e); if (a >= 32 && a <= GPI_VEC_THREADS && a % 32 == 0) h->blkv = a; }
    if (const char* e = getenv("GPI_BLOCK2")) { int a, b; if (sscanf(e, "%d,%d", &a, &b) == 2 && a * b <= 256) h->blk2 = dim3(a, b, 1); }
    if (create_impl(h)) { g_create_err = "gpi_create: " + h->err; gpi_destroy(h); return 1; }
    *out = h;
    return 0;
}

extern "C" int gpi_destroy(gpi_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->comm && h->nccl.CommDestroy) h->nccl.CommDestroy(h->comm);
    cudaFree(h->W); cudaFree(h->TP); cudaFree(h->MEM); cudaFree(h->pmlcoef); cudaFree(h->pmlztab);
    for (auto& p : h->mod) cudaFree(p);
    for (auto& p : h->dmod) cudaFree(p);
    cudaFree(h->dmod_table);
    for (auto& p : h->gtot) cudaFree(p);
    cudaFree(h->gshot);
    cudaFree(h->illum_acc); cudaFree(h->illum_stack);
    for (int ipw = 0; ipw < 2; ipw++) for (auto& s : h->shots[ipw]) {
        for (int f = 0; f < GPI_NWAVEFIELD; f++) {
            free_sparse(s.spray[f]); free_sparse(s.interp[f]);
            cudaFree(s.wav[f]); cudaFree(s.rec[f]); cudaFree(s.snap[f]);
            for (int q = 0; q < 3; q++) cudaFree(s.bnd[f][q]);
        }
        for (auto p : s.usnaps) cudaFree(p);
    }
    cudaFree(h->post_v); cudaFree(h->post_s); cudaFree(h->bnd_table); cudaFree(h->stash); cudaFree(h->stash_table);
    if (h->h_post_v) cudaFreeHost(h->h_post_v);
    if (h->h_post_s) cudaFreeHost(h->h_post_s);
    if (h->stage) cudaFreeHost(h->stage);
    cudaFree(h->dscratch);
    for (auto& p : h->modp) cudaFree(p);
    for (auto& p : h->born_c) cudaFree(p);
    cudaFree(h->born_d);
    for (auto& ts : h->tmaps) { cudaFree(ts.d[0]); cudaFree(ts.d[1]); }
    cudaFree(h->t3_tiles[0]); cudaFree(h->t3_tiles[1]);
#ifndef GPI_HOST_EMU
    for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy((cudaGraphExec_t)e.exec);
#endif
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side) cudaStreamDestroy(h->side);
    for (auto e : h->evpool) cudaEventDestroy(e);
    for (int d = 0; d < 2; d++) { cudaFree(h->halo_send[d]); cudaFree(h->halo_recv[d]); }
    for (int q = 0; q < 2; q++) { if (h->ev_half[q]) cudaEventDestroy(h->ev_half[q]); if (h->ev_xv[q]) cudaEventDestroy(h->ev_xv[q]); if (h->ev_xt[q]) cudaEventDestroy(h->ev_xt[q]); }
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int gpi_set_stream(gpi_handle* h, void* s) {
    GUARD(h);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)s; h->own_stream = false;
    return 0;
}
extern "C" int gpi_synchronize(gpi_handle* h) { GUARD(h); CU(h, cudaStreamSynchronize(h->stream)); return 0; }

// =================================================================================================
// medium, CPML, acquisition, wavelets
// =================================================================================================
extern "C" int gpi_set_medium(gpi_handle* h, int p, const float* a) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->mod[p]) FAIL(h, "medium parameter %d is not part of this physics", p);
    if (!a) FAIL(h, "null medium array");
    return upload_field(h, h->el ? GPI_TAUXX : GPI_P, a, h->mod[p]);
}
extern "C" int gpi_set_medium_rows(gpi_handle* h, int p, const float* rows, int k_first, int nk) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->mod[p]) FAIL(h, "medium parameter %d is not part of this physics", p);
    if (!rows || k_first < 0 || nk < 1 || k_first + nk > h->g.nz) FAIL(h, "bad medium row window [%d, %d)", k_first, k_first + nk);
    return upload_field(h, h->el ? GPI_TAUXX : GPI_P, rows, h->mod[p], k_first, nk);
}
// update!(pa, medium) without the host-side padarray: the interior array goes to the device as it is (one
// contiguous H2D copy, 1/2.3 of the extended bytes at C3) and the replicate padding happens there.
extern "C" int gpi_set_medium_interior(gpi_handle* h, int p, const float* a, const int32_t n_in[3], const int32_t lo[3]) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->mod[p]) FAIL(h, "medium parameter %d is not part of this physics", p);
    if (!a || !n_in || !lo) FAIL(h, "null medium array");
    const Geom& g = h->g;
    const int mz = n_in[0], my = h->nd == 3 ? n_in[1] : 1, mx = n_in[2];
    const int lz = lo[0], ly = h->nd == 3 ? lo[1] : 0, lx = lo[2];
    if (mz < 1 || my < 1 || mx < 1 || lz < 0 || ly < 0 || lx < 0 || mz + lz > g.nz || my + ly > g.ny || mx + lx > g.nx)
        FAIL(h, "interior medium [%d,%d,%d] + padding [%d,%d,%d] does not fit the extended grid [%d,%d,%d]", mz, my, mx, lz, ly, lx, g.nz, g.ny, g.nx);
    const size_t nf = (size_t)mz * my * mx;
    if (h->dscratch_floats < nf) {
        cudaFree(h->dscratch);
        h->dscratch = nullptr; h->dscratch_floats = 0;
        CU(h, cudaMalloc((void**)&h->dscratch, nf * sizeof(float)));
        h->dscratch_floats = nf;
    }
    CU(h, cudaMemcpyAsync(h->dscratch, a, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    dim3 blk(128), grd((g.pz + 127) / 128, g.ny, g.nx);
    k_pad_replicate<<<grd, blk, 0, h->stream>>>(g, h->dscratch, h->mod[p], mz, my, mx, lz, ly, lx);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));     // `a` is borrowed for the duration of the call only
    return 0;
}
// update!(pa, medium) in one call: vp, (vs,) rho of the UN-extended medium -> every independent parameter on the extended grid
// (k_pad_derive).  Replaces one Medium getter + gpi_set_medium_interior per parameter; same H2D bytes, no host-side arithmetic.
extern "C" int gpi_set_medium_fields(gpi_handle* h, const float* vp, const float* vs, const float* rho, const int32_t n_in[3], const int32_t lo[3]) {
    GUARD(h);
    Range range(h, "%s", "gpi_set_medium_fields (update!(pa, medium))");
    if (!vp || !rho || !n_in || !lo || (h->el && !vs)) FAIL(h, "null medium array (an elastic medium needs vp, vs and rho)");
    const Geom& g = h->g;
    const int mz = n_in[0], my = h->nd == 3 ? n_in[1] : 1, mx = n_in[2];
    const int lz = lo[0], ly = h->nd == 3 ? lo[1] : 0, lx = lo[2];
    if (mz < 1 || my < 1 || mx < 1 || lz < 0 || ly < 0 || lx < 0 || mz + lz > g.nz || my + ly > g.ny || mx + lx > g.nx)
        FAIL(h, "interior medium [%d,%d,%d] + padding [%d,%d,%d] does not fit the extended grid [%d,%d,%d]", mz, my, mx, lz, ly, lx, g.nz, g.ny, g.nx);
    const size_t nf = (size_t)mz * my * mx, need = 3 * nf;
    if (h->dscratch_floats < need) {
        cudaFree(h->dscratch);
        h->dscratch = nullptr; h->dscratch_floats = 0;
        CU(h, cudaMalloc((void**)&h->dscratch, need * sizeof(float)));
        h->dscratch_floats = need;
    }
    float* dvp = h->dscratch; float* dvs = h->dscratch + nf; float* drho = h->dscratch + 2 * nf;
    CU(h, cudaMemcpyAsync(dvp, vp, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (h->el) CU(h, cudaMemcpyAsync(dvs, vs, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(drho, rho, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    dim3 blk(128), grd((g.pz + 127) / 128, g.ny, g.nx);
    if (h->el) k_pad_derive<1><<<grd, blk, 0, h->stream>>>(g, dvp, dvs, drho, h->mod[GPI_INVLAMBDA], h->mod[GPI_INVMU], h->mod[GPI_RHO], mz, my, mx, lz, ly, lx);
    else       k_pad_derive<0><<<grd, blk, 0, h->stream>>>(g, dvp, nullptr, drho, h->mod[GPI_INVK], nullptr, h->mod[GPI_RHO], mz, my, mx, lz, ly, lx);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));     // the host arrays are borrowed for the duration of the call only
    return 0;
}
extern "C" int gpi_slab_range(gpi_handle* h, int32_t* k_begin, int32_t* k_end) {
    if (!h || !k_begin || !k_end) return 1;
    *k_begin = h->ka; *k_end = h->kb;
    return 0;
}
extern "C" int gpi_get_medium(gpi_handle* h, int p, float* out) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->mod[p]) FAIL(h, "medium parameter %d is not part of this physics", p);
    return download_field(h, h->el ? GPI_TAUXX : GPI_P, h->mod[p], out);
}
extern "C" int gpi_update_dmod(gpi_handle* h) {
    GUARD(h);
    Range range(h, "%s", "gpi_update_dmod");
    dim3 blk = h->nd == 3 ? h->blk3 : h->blk2, grd = grid_for(h, blk, 1);
    const float dt = (float)h->c.dt;
    const float* m0 = h->el ? h->mod[GPI_INVLAMBDA] : h->mod[GPI_INVK];
    if (h->c.order == 4) {
        if (h->nd == 2 && !h->el) k_dmod4<2, 0><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], nullptr, h->dmod_table, dt);
        else if (h->nd == 2)      k_dmod4<2, 1><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], h->mod[GPI_INVMU], h->dmod_table, dt);
        else if (!h->el)          k_dmod4<3, 0><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], nullptr, h->dmod_table, dt);
        else                      k_dmod4<3, 1><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], h->mod[GPI_INVMU], h->dmod_table, dt);
        CU(h, cudaGetLastError());
        CU(h, cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (h->nd == 2 && !h->el) k_dmod<2, 0><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], nullptr, h->dmod_table, dt);
    else if (h->nd == 2)      k_dmod<2, 1><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], h->mod[GPI_INVMU], h->dmod_table, dt);
    else if (!h->el)          k_dmod<3, 0><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], nullptr, h->dmod_table, dt);
    else                      k_dmod<3, 1><<<grd, blk, 0, h->stream>>>(h->g, m0, h->mod[GPI_RHO], h->mod[GPI_INVMU], h->dmod_table, dt);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int gpi_set_medium_pert(gpi_handle* h, int p, const float* a) {
    GUARD(h);
    if (h->nd != 2 || h->el) FAIL(h, "FD-Born exists for 2-D acoustic media only (born.jl:1-12)");
    if (h->c.order != 2) FAIL(h, "FD-Born is defined for order 2 only");
    if (p != GPI_INVK && p != GPI_RHO) FAIL(h, "medium perturbation %d is not a parameter of this physics", p);
    if (!a) FAIL(h, "null medium perturbation");
    if (!h->modp[p]) { CU(h, cudaMalloc((void**)&h->modp[p], (size_t)h->g.vol * sizeof(float))); CU(h, cudaMemset(h->modp[p], 0, (size_t)h->g.vol * sizeof(float))); }
    h->born_ready = false;
    return upload_field(h, GPI_P, a, h->modp[p]);
}
extern "C" int gpi_update_born(gpi_handle* h) {
    GUARD(h);
    if (h->nd != 2 || h->el || h->npw != 2) FAIL(h, "FD-Born needs a 2-D acoustic experiment with npw = 2");
    if (!h->modp[GPI_INVK] || !h->modp[GPI_RHO]) FAIL(h, "gpi_update_born: set the perturbations of invK and rho first");
    const size_t vb = (size_t)h->g.vol * sizeof(float);
    for (auto& p : h->born_c) if (!p) { CU(h, cudaMalloc((void**)&p, vb)); CU(h, cudaMemset(p, 0, vb)); }
    if (!h->born_d) CU(h, cudaMalloc((void**)&h->born_d, (size_t)h->B * 2 * vb));
    dim3 blk = h->blk2, grd = grid_for(h, blk, 1);
    k_born_coef<<<grd, blk, 0, h->stream>>>(h->g, h->mod[GPI_INVK], h->mod[GPI_RHO], h->modp[GPI_INVK], h->modp[GPI_RHO],
                                            h->born_c[0], h->born_c[1], h->born_c[2], (float)h->c.dt);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    h->born_ready = true;
    return 0;
}
extern "C" int gpi_set_pml(gpi_handle* h, int f, const float* a, const float* b, const float* kI) {
    GUARD(h);
    if (f < GPI_NWAVEFIELD || !field_exists(h->nd, h->c.physics, f)) FAIL(h, "field %d is not a derivative field of this physics", f);
    const size_t np2 = 2 * h->c.npml;
    CU(h, cudaMemcpy(h->pmlcoef + ((size_t)f * 3 + 0) * np2, a, np2 * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->pmlcoef + ((size_t)f * 3 + 1) * np2, b, np2 * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->pmlcoef + ((size_t)f * 3 + 2) * np2, kI, np2 * sizeof(float), cudaMemcpyHostToDevice));
    if (dfield_axis(f) == 0) {
        // z terms: expand onto the unified z coordinate; identity (a = b = 0, kI = 1) outside the slabs
        const Geom& g = h->g;
        int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
        field_shape(h->nd, f, n, sh, off, h->c.order);
        const int s0 = off[0], len = sh[0], npml = h->c.npml;
        const int pzt = h->pzt;
        std::vector<float> ta(pzt, 0.f), tb(pzt, 0.f), tk(pzt, 1.f);
        for (int k = 0; k < pzt; k++) {
            const int r = k - s0;
            int s = -1;
            if ((h->c.pml_faces & ZMIN) && r >= 0 && r < npml) s = r;
            else if ((h->c.pml_faces & ZMAX) && r - (len - npml) >= 0 && r < len) s = npml + r - (len - npml);
            if (s >= 0) { ta[k] = a[s]; tb[k] = b[s]; tk[k] = kI[s]; }
        }
        CU(h, cudaMemcpy(h->pmlztab + ((size_t)f * 3 + 0) * pzt, ta.data(), pzt * sizeof(float), cudaMemcpyHostToDevice));
        CU(h, cudaMemcpy(h->pmlztab + ((size_t)f * 3 + 1) * pzt, tb.data(), pzt * sizeof(float), cudaMemcpyHostToDevice));
        CU(h, cudaMemcpy(h->pmlztab + ((size_t)f * 3 + 2) * pzt, tk.data(), pzt * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" int gpi_set_sparse(gpi_handle* h, int kind, int ipw, int issp, int f, int ncol,
                              const int64_t* colptr, const int64_t* rowval, const float* nzval) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || issp < 0 || issp >= h->c.nshots) FAIL(h, "bad pw/shot index (%d, %d)", ipw, issp);
    if (f < 0 || f >= GPI_NWAVEFIELD || !field_exists(h->nd, h->c.physics, f)) FAIL(h, "field %d is not a wavefield of this physics", f);
    if (kind != GPI_SPRAY && kind != GPI_INTERP) FAIL(h, "kind must be GPI_SPRAY or GPI_INTERP");
    if (ncol < 0 || !colptr || (ncol > 0 && (!rowval || !nzval))) FAIL(h, "null sparse arrays");
    ShotData& s = h->shots[ipw][issp];
    SparseDev& m = kind == GPI_SPRAY ? s.spray[f] : s.interp[f];
    free_sparse(m);
    const Geom& g = h->g;
    int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
    field_shape(h->nd, f, n, sh, off, h->c.order);
    const long long flen = (long long)sh[0] * sh[1] * sh[2];
    const int64_t nnz = colptr[ncol] - 1;
    std::vector<int> cp(ncol + 1), cell(nnz);
    std::vector<float> val(nnz);
    bool isv; int slot; wf_slot(f, isv, slot);
    std::map<int, std::vector<std::pair<int, float>>> rows;
    for (int jcol = 0; jcol < ncol; jcol++) {
        cp[jcol] = (int)(colptr[jcol] - 1);
        if (colptr[jcol + 1] < colptr[jcol]) FAIL(h, "colptr not monotone");
        for (int64_t e = colptr[jcol] - 1; e < colptr[jcol + 1] - 1; e++) {
            const int64_t r = rowval[e] - 1;
            if (r < 0 || r >= flen) FAIL(h, "row index %lld outside the field array of %lld entries", (long long)rowval[e], flen);
            const int iz = (int)(r % sh[0]), iy = (int)((r / sh[0]) % sh[1]), ix = (int)(r / ((long long)sh[0] * sh[1]));
            const int k = iz + off[0], j = iy + off[1], i = ix + off[2];       // k: global unified z index
            const int kl = k - g.koff;
            const bool owned = kl >= g.klo && kl <= g.khi;
            // a tap outside this handle's slab contributes nothing here (its owner adds it; records are summed over ranks)
            cell[e] = owned ? (int)((long long)kl + (long long)g.pz * (j + (long long)g.ny1 * i)) : (int)g.klo;
            val[e] = owned ? nzval[e] : 0.f;
            if (!owned) continue;
            // injection view: velocity sources only act on the @inn range (source.jl:166-177)
            bool ok = true;
            if (isv) {
                // unified coordinates (storage - h); J = inner nodes [O, n-1-O], H = half nodes [1+h, n-1-h]
                const int hh = g.h, O = h->c.order - 1;
                const int ku = k - hh, ju = h->nd == 3 ? j - hh : 0, iu = i - hh;
                auto inJ = [&](int u, int n) { return u >= O && u <= n - 1 - O; };
                auto inH = [&](int u, int n) { return u >= 1 + hh && u <= n - 1 - hh; };
                const bool jJ = h->nd == 2 || inJ(ju, g.ny), jH = h->nd == 2 || inH(ju, g.ny);
                if (slot == V_X) ok = inJ(ku, g.nz) && jJ && inH(iu, g.nx);
                if (slot == V_Z) ok = inH(ku, g.nz) && jJ && inJ(iu, g.nx);
                if (slot == V_Y) ok = inJ(ku, g.nz) && jH && inJ(iu, g.nx);
            }
            if (ok) rows[cell[e]].push_back({jcol, nzval[e]});
        }
    }
    cp[ncol] = (int)nnz;
    std::vector<int> rc, rp{0}, ec; std::vector<float> ev;
    for (auto& kv : rows) {
        rc.push_back(kv.first);
        for (auto& e : kv.second) { ec.push_back(e.first); ev.push_back(e.second); }
        rp.push_back((int)ec.size());
    }
    m.ncol = ncol; m.nnz = (int)nnz; m.nrows = (int)rc.size();
    if (to_device(h, &m.colptr, cp) || to_device(h, &m.tap_cell, cell) || to_device(h, &m.tap_val, val) ||
        to_device(h, &m.row_cell, rc) || to_device(h, &m.row_ptr, rp) || to_device(h, &m.ent_col, ec) || to_device(h, &m.ent_val, ev)) return 1;
    m.set = true;
    if (kind == GPI_INTERP) {
        cudaFree(s.rec[f]); s.rec[f] = nullptr; s.nr[f] = ncol;
        const size_t nb = (size_t)h->c.nt * std::max(ncol, 1) * sizeof(float);
        CU(h, cudaMalloc((void**)&s.rec[f], nb));
        CU(h, cudaMemset(s.rec[f], 0, nb));
    }
    return 0;
}

extern "C" int gpi_set_wavelets(gpi_handle* h, int ipw, int issp, int f, int ns, const float* w) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || issp < 0 || issp >= h->c.nshots) FAIL(h, "bad pw/shot index (%d, %d)", ipw, issp);
    if (f < 0 || f >= GPI_NWAVEFIELD || !field_exists(h->nd, h->c.physics, f)) FAIL(h, "field %d is not a wavefield of this physics", f);
    ShotData& s = h->shots[ipw][issp];
    const size_t nb = (size_t)h->c.nt * (ns > 0 ? ns : 0) * sizeof(float);
    if (w && ns > 0 && s.wav[f] && s.ns[f] == ns) {      // same shape as before (every FWI iteration): no free / malloc
        CU(h, cudaMemcpy(s.wav[f], w, nb, cudaMemcpyHostToDevice));
        return 0;
    }
    cudaFree(s.wav[f]); s.wav[f] = nullptr; s.ns[f] = 0;
    if (!w || ns <= 0) return 0;           // removes the source field
    CU(h, cudaMalloc((void**)&s.wav[f], nb));
    CU(h, cudaMemcpy(s.wav[f], w, nb, cudaMemcpyHostToDevice));
    s.ns[f] = ns;
    return 0;
}

extern "C" int gpi_set_snap_steps(gpi_handle* h, int nsnaps, const int32_t* its) {
    GUARD(h);
    if (nsnaps != h->c.nsnaps) FAIL(h, "nsnaps differs from the configuration");
    h->itsnaps.assign(its, its + nsnaps);
    return 0;
}

extern "C" int gpi_reset(gpi_handle* h, int what) {
    GUARD(h);
    const size_t vb = (size_t)h->g.vol * sizeof(float);
    if (what & GPI_RESET_WAVEFIELDS) {
        CU(h, cudaMemsetAsync(h->W, 0, (size_t)h->B * h->bstride * sizeof(float), h->stream));
        if (h->TP) CU(h, cudaMemsetAsync(h->TP, 0, (size_t)h->B * h->bstride * sizeof(float), h->stream));
        CU(h, cudaMemsetAsync(h->MEM, 0, (size_t)h->B * h->npw * h->mem_per_pw * sizeof(float), h->stream));
    }
    for (int ipw = 0; ipw < h->npw; ipw++) for (auto& s : h->shots[ipw]) {
        if (what & GPI_RESET_RECORDS) for (int f = 0; f < GPI_NWAVEFIELD; f++) if (s.rec[f])
            CU(h, cudaMemsetAsync(s.rec[f], 0, (size_t)h->c.nt * std::max(s.nr[f], 1) * sizeof(float), h->stream));
        if (what & GPI_RESET_BOUNDARY) for (int f = 0; f < GPI_NWAVEFIELD; f++) {
            if (s.snap[f]) CU(h, cudaMemsetAsync(s.snap[f], 0, vb, h->stream));
            for (int q = 0; q < 3; q++) if (s.bnd[f][q]) CU(h, cudaMemsetAsync(s.bnd[f][q], 0, (size_t)bnd_slot_floats(h, q) * h->c.nt * sizeof(float), h->stream));
        }
        if (what & GPI_RESET_SNAPS) for (auto p : s.usnaps) CU(h, cudaMemsetAsync(p, 0, vb, h->stream));
    }
    if (what & GPI_RESET_GRADIENTS) {
        for (auto& p : h->gtot) if (p) CU(h, cudaMemsetAsync(p, 0, vb, h->stream));
        if (h->gshot) CU(h, cudaMemsetAsync(h->gshot, 0, (size_t)h->B * h->ngrad * vb, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// =================================================================================================
// the hot loop
// =================================================================================================
namespace {

InjOp make_inj(gpi_handle* h, const SparseDev& m, const float* wav, int kind, int axis) {
    InjOp op; memset(&op, 0, sizeof op);
    op.kind = kind; op.axis = axis;
    op.nrows = m.nrows; op.row_cell = m.row_cell; op.row_ptr = m.row_ptr; op.ent_col = m.ent_col; op.ent_val = m.ent_val;
    op.wav = wav;
    (void)h;
    return op;
}

// build the per-slot source/receiver descriptors of one batch (add_*_source!, record!)
int build_post(gpi_handle* h, int shot0, int nb, int activepw, int src_flags) {
    const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ}, vaxis[3] = {2, 1, 0};
    const int sf[4] = {GPI_P, GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ};
    for (int b = 0; b < nb; b++) {
        const int is = shot0 + b;
        PostDesc& pv = h->h_post_v[b]; PostDesc& ps = h->h_post_s[b];
        memset(&pv, 0, sizeof pv); memset(&ps, 0, sizeof ps);
        // ---- velocity sources (source.jl:124-157)
        if ((activepw & 1) && (src_flags & 1)) for (int i = 0; i < 3; i++) {
            const int f = vf[i]; ShotData& s = h->shots[0][is];
            if (!s.wav[f]) continue;
            if (!s.spray[f].set) FAIL(h, "shot %d: wavelets set for field %d but no spray matrix", is, f);
            if (s.spray[f].ncol != s.ns[f]) FAIL(h, "shot %d field %d: %d wavelets but %d spray columns", is, f, s.ns[f], s.spray[f].ncol);
            InjOp op = make_inj(h, s.spray[f], s.wav[f], 0, vaxis[i]);
            op.ntarget = 1; op.target[0] = wf_ptr(h, h->W, b, 0, f); op.coef = h->mod[GPI_RHO];
            if (pv.ninj >= MAX_OPS) FAIL(h, "too many source fields");
            pv.inj[pv.ninj++] = op;
        }
        if (h->npw == 2 && (activepw & 2) && (src_flags & 2)) for (int i = 0; i < 3; i++) {
            const int f = vf[i]; ShotData& s2 = h->shots[1][is]; ShotData& s1 = h->shots[0][is];
            if (!s2.wav[f]) continue;
            // adjoint sources are sprayed through pw 1's receiver matrix (source.jl:142-156)
            if (!s1.interp[f].set) FAIL(h, "shot %d: adjoint wavelets for field %d but pw 1 has no receiver matrix for it", is, f);
            if (s1.interp[f].ncol != s2.ns[f]) FAIL(h, "shot %d field %d: %d adjoint wavelets but %d receivers", is, f, s2.ns[f], s1.interp[f].ncol);
            InjOp op = make_inj(h, s1.interp[f], s2.wav[f], 0, vaxis[i]);
            op.ntarget = 1; op.target[0] = wf_ptr(h, h->W, b, 1, f); op.coef = h->mod[GPI_RHO];
            if (pv.ninj >= MAX_OPS) FAIL(h, "too many source fields");
            pv.inj[pv.ninj++] = op;
        }
        // ---- stress sources, pw 1 only (source.jl:61-119)
        if ((activepw & 1) && (src_flags & 1)) for (int i = 0; i < 4; i++) {
            const int f = sf[i]; ShotData& s = h->shots[0][is];
            if (!field_exists(h->nd, h->c.physics, f) || !s.wav[f]) continue;
            if (!s.spray[f].set) FAIL(h, "shot %d: wavelets set for field %d but no spray matrix", is, f);
            if (s.spray[f].ncol != s.ns[f]) FAIL(h, "shot %d field %d: %d wavelets but %d spray columns", is, f, s.ns[f], s.spray[f].ncol);
            InjOp op = make_inj(h, s.spray[f], s.wav[f], 1, 0);
            op.coef = h->dmod[C_K];                    // dtK (acoustic) / dtM (elastic)
            if (!h->el) { op.ntarget = 1; op.target[0] = wf_ptr(h, h->W, b, 0, GPI_P); }
            else {                                     // every normal stress gets the same term (source.jl:99-118)
                op.ntarget = 0;
                op.target[op.ntarget++] = wf_ptr(h, h->W, b, 0, GPI_TAUXX);
                if (h->nd == 3) op.target[op.ntarget++] = wf_ptr(h, h->W, b, 0, GPI_TAUYY);
                op.target[op.ntarget++] = wf_ptr(h, h->W, b, 0, GPI_TAUZZ);
            }
            if (ps.ninj >= MAX_OPS) FAIL(h, "too many source fields");
            ps.inj[ps.ninj++] = op;
        }
        // ---- receivers (receiver.jl:3-14; only p and velocities are ever recorded, propagate.jl:177,208)
        for (int ipw = 0; ipw < h->npw; ipw++) {
            if (!(activepw & (1 << ipw))) continue;
            ShotData& s = h->shots[ipw][is];
            for (int i = 0; i < 3; i++) {
                const int f = vf[i];
                if (!s.interp[f].set || !s.rec[f]) continue;
                RecOp op; op.field = wf_ptr(h, h->W, b, ipw, f); op.nr = s.interp[f].ncol; op.colptr = s.interp[f].colptr;
                op.tap_cell = s.interp[f].tap_cell; op.tap_val = s.interp[f].tap_val; op.rec = s.rec[f];
                if (pv.nrec >= MAX_OPS) FAIL(h, "too many receiver fields");
                pv.rec[pv.nrec++] = op;
            }
            if (!h->el && s.interp[GPI_P].set && s.rec[GPI_P]) {
                RecOp op; op.field = wf_ptr(h, h->W, b, ipw, GPI_P); op.nr = s.interp[GPI_P].ncol; op.colptr = s.interp[GPI_P].colptr;
                op.tap_cell = s.interp[GPI_P].tap_cell; op.tap_val = s.interp[GPI_P].tap_val; op.rec = s.rec[GPI_P];
                ps.rec[ps.nrec++] = op;
            }
        }
    }
    CU(h, cudaMemcpyAsync(h->post_v, h->h_post_v, (size_t)nb * sizeof(PostDesc), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->post_s, h->h_post_s, (size_t)nb * sizeof(PostDesc), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// z-slab halo exchange with the two z neighbours.  phase 0 (before the velocity kernel):
// tauzz | p travels up (plane khi -> neighbour's klo-1), tauxz and tauyz travel down (plane klo ->
// neighbour's khi+1).  phase 1 (before the stress kernel): vx, vy travel up, vz travels down.
// One pack kernel per direction, one grouped NCCL send/recv over NVLink, one unpack kernel per direction,
// all on the engine's stream.
int exchange_halos(gpi_handle* h, int phase, bool sample = false, int i0 = 0, int ni = -1, int half = 0, cudaStream_t st = nullptr) {
    if (!h->slab) return 0;
    if (!h->comm) FAIL(h, "z-slab handles need gpi_nccl_init before gpi_run");
    if (!st) st = h->stream;
    const Geom& g = h->g;
    if (ni < 0) ni = g.nx1;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (sample) { e0 = sample_event(h); e1 = sample_event(h); }
    if (e0 && e1) { cudaEventRecord(e0, st); h->evkind.push_back(ni == g.nx1 ? 2 : 5); }      // 5: one x half
    struct Closer { cudaStream_t s; cudaEvent_t e; ~Closer() { if (e) cudaEventRecord(e, s); } } closer{st, (e0 && e1) ? e1 : nullptr};
    const size_t plane = (size_t)g.ny1 * ni;
    // the second half's messages live behind the first half's in the buffers (3 fields x ny1 x nx1 floats in all)
    const size_t boff = half ? (size_t)3 * g.ny1 * i0 : 0;
    int up[3], dn[3], nup = 0, ndn = 0;
    if (phase == 0) {
        up[nup++] = h->el ? GPI_TAUZZ : GPI_P;
        if (h->el) { dn[ndn++] = GPI_TAUXZ; dn[ndn++] = GPI_TAUYZ; }
    } else {
        if (h->el) { up[nup++] = GPI_VX; up[nup++] = GPI_VY; }
            dn[ndn++] = GPI_VZ;
    }
    const bool has_up = h->srank < h->snranks - 1, has_dn = h->srank > 0;
    dim3 blk(128), grd((g.ny1 + 127) / 128, ni);
    auto args = [&](const int* f, int n, int k) { HaloArgs a; a.n = n; a.i0 = i0; a.ni = ni; for (int q = 0; q < 3; q++) { a.field[q] = q < n ? wf_ptr(h, h->W, 0, 0, f[q]) : nullptr; a.k[q] = k; } return a; };
    if (has_up && nup) { k_halo<1><<<grd, blk, 0, st>>>(g, args(up, nup, g.khi), h->halo_send[1] + boff); h->timers.launches += 1; }
    if (has_dn && ndn) { k_halo<1><<<grd, blk, 0, st>>>(g, args(dn, ndn, g.klo), h->halo_send[0] + boff); h->timers.launches += 1; }
    const int F32 = 7;   // ncclFloat32
    int rc = h->nccl.GroupStart();
    if (has_up) {
        if (nup && !rc) rc = h->nccl.Send(h->halo_send[1] + boff, nup * plane, F32, h->srank + 1, h->comm, st);
        if (ndn && !rc) rc = h->nccl.Recv(h->halo_recv[1] + boff, ndn * plane, F32, h->srank + 1, h->comm, st);
    }
    if (has_dn) {
        if (ndn && !rc) rc = h->nccl.Send(h->halo_send[0] + boff, ndn * plane, F32, h->srank - 1, h->comm, st);
        if (nup && !rc) rc = h->nccl.Recv(h->halo_recv[0] + boff, nup * plane, F32, h->srank - 1, h->comm, st);
    }
    const int rc2 = h->nccl.GroupEnd();
    if (rc || rc2) FAIL(h, "NCCL halo exchange failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(rc ? rc : rc2) : "?");
    if (has_up && ndn) { k_halo<0><<<grd, blk, 0, st>>>(g, args(dn, ndn, g.khi + 1), h->halo_recv[1] + boff); h->timers.launches += 1; }
    if (has_dn && nup) { k_halo<0><<<grd, blk, 0, st>>>(g, args(up, nup, g.klo - 1), h->halo_recv[0] + boff); h->timers.launches += 1; }
    return 0;
}

bool any_post(const PostDesc* d, int nb, bool with_rec) {
    for (int b = 0; b < nb; b++) if (d[b].ninj > 0 || (with_rec && d[b].nrec > 0)) return true;
    return false;
}

}  // namespace

extern "C" int gpi_run(gpi_handle* h, int mode, int activepw, int src_flags) {
    GUARD(h);
    const bool born = (mode & GPI_RUN_BORN) != 0;
    const int unshifted = (mode & GPI_RUN_UNSHIFTED_RHO) ? 1 : 0;
    mode &= ~(GPI_RUN_BORN | GPI_RUN_UNSHIFTED_RHO);
    if (born && (h->nd != 2 || h->el || h->npw != 2 || (activepw & 3) != 3)) FAIL(h, "FD-Born needs a 2-D acoustic experiment with both wavefields active");
    if ((born || unshifted) && h->c.order != 2) FAIL(h, "FD-Born and its exact-transpose imaging are defined for order 2 only");
    if (born && mode == GPI_MODE_ADJOINT) FAIL(h, "FD-Born scattering sources exist in the forward modes only (born.jl:27-30)");
    if (born && !h->born_ready) FAIL(h, "FD-Born: call gpi_set_medium_pert and gpi_update_born first");
    if (mode != GPI_MODE_FORWARD && mode != GPI_MODE_FORWARD_SAVE && mode != GPI_MODE_ADJOINT) FAIL(h, "unknown mode %d", mode);
    if (!(activepw & 1)) FAIL(h, "pw 1 must be active");
    if ((activepw & 2) && h->npw < 2) FAIL(h, "pw 2 requested but the experiment was built with npw = 1");
    if (h->slab && mode != GPI_MODE_FORWARD) FAIL(h, "z-slab handles run forward modelling only");
    if (h->slab && (!h->comm || h->nranks != h->snranks || h->rank != h->srank)) FAIL(h, "z-slab handles need gpi_nccl_init(rank = slab_rank, nranks = slab_nranks) before gpi_run");
    if (mode == GPI_MODE_FORWARD_SAVE && !h->c.store_boundary) FAIL(h, "forward_save needs store_boundary=1 at construction (fdtd.jl:445-455)");
    if (mode == GPI_MODE_ADJOINT && !h->c.store_boundary) FAIL(h, "adjoint needs the boundary store of a forward_save run");
    const bool grad = mode == GPI_MODE_ADJOINT && (activepw & 2) && h->npw == 2;
    // default (GPI_PINGPONG=0 opts out; measured r02: C4 65.0 -> 76.4 Gcell-updates/s): the two wavefield sets alternate as "this step" / "previous step" (what save_tp! copies, save_tp.jl:5-12)
    const bool pp = h->pingpong && mode == GPI_MODE_ADJOINT && h->TP && h->c.order == 2 && (h->nd == 2 ? h->vec2 : h->vec3) && !born && !h->slab;
    if (pp && ensure_stash(h)) return 1;
    if (grad && !h->gshot) FAIL(h, "gradient imaging needs an experiment built with npw = 2");
    if (unshifted && h->el) FAIL(h, "the exact-transpose rho imaging is defined for acoustic media");
    const Geom& g = h->g;
    const int nt = h->c.nt;
    const size_t vb = (size_t)g.vol * sizeof(float);
    int bf[6]; const int nbf = boundary_fields(h, bf);
    const int vf[3] = {GPI_VX, GPI_VY, GPI_VZ};
    Range run_range(h, "gpi_run %s%s pw %d", mode == GPI_MODE_FORWARD ? "forward" : mode == GPI_MODE_FORWARD_SAVE ? "forward_save" : "adjoint", born ? " born" : "", activepw);
    h->timers = gpi_timers{};
    h->evused = 0; h->evkind.clear();
    CU(h, cudaEventRecord(h->ev0, h->stream));

    if (h->illum_on) CU(h, cudaMemsetAsync(h->illum_stack, 0, (size_t)g.vol * sizeof(double), h->stream));      // initialize!(pac): fill!(illum_stack, 0.0) (types.jl:171)
    int graph_batches = 0;
    for (int shot0 = 0; shot0 < h->c.nshots; shot0 += h->B) {
        const int nb = std::min(h->B, h->c.nshots - shot0);
        Range batch_range(h, "shots %d..%d (%d time steps)", shot0, shot0 + nb - 1, nt);
        // reset_w2! (types.jl:100-113)
        CU(h, cudaMemsetAsync(h->W, 0, (size_t)nb * h->bstride * sizeof(float), h->stream));
        if (h->TP) CU(h, cudaMemsetAsync(h->TP, 0, (size_t)nb * h->bstride * sizeof(float), h->stream));
        CU(h, cudaMemsetAsync(h->MEM, 0, (size_t)nb * h->npw * h->mem_per_pw * sizeof(float), h->stream));
        if (grad) CU(h, cudaMemsetAsync(h->gshot, 0, (size_t)nb * h->ngrad * vb, h->stream));
        if (h->illum_on) CU(h, cudaMemsetAsync(h->illum_acc, 0, (size_t)nb * g.vol * sizeof(double), h->stream));
        if (mode == GPI_MODE_ADJOINT) {       // boundary_force_snap_tau!/v! (boundary.jl:173-212)
            for (int b = 0; b < nb; b++) {
                ShotData& s = h->shots[0][shot0 + b];
                for (int i = 0; i < nbf; i++) CU(h, cudaMemcpyAsync(wf_ptr(h, h->W, b, 0, bf[i]), s.snap[bf[i]], vb, cudaMemcpyDeviceToDevice, h->stream));
                for (int i = 0; i < 3; i++) if (s.snap[vf[i]]) CU(h, cudaMemcpyAsync(wf_ptr(h, h->W, b, 0, vf[i]), s.snap[vf[i]], vb, cudaMemcpyDeviceToDevice, h->stream));
            }
        }
        if (build_post(h, shot0, nb, activepw, src_flags)) return 1;
        if (mode != GPI_MODE_FORWARD && build_bnd_table(h, shot0, nb)) return 1;
        const bool do_post_v = any_post(h->h_post_v, nb, true);
        const bool inj_s = any_post(h->h_post_s, nb, false), rec_s = any_post(h->h_post_s, nb, true);
        StepArgs args[2];
        for (int ipw = 0; ipw < h->npw; ipw++) fill_args(h, args[ipw], ipw, nb);
        // adjoint / two-wavefield runs without the FD-Born write-out: pw 1 and pw 2 of all resident shots share each launch
        const bool merge_pw = h->npw == 2 && (activepw & 3) == 3 && !born && h->c.order == 2 && h->nd == 2;
        StepArgs margs;
        if (merge_pw) fill_args(h, margs, 0, nb, true);
        // 2-D acoustic adjoint with ping-pong levels: stress update of both wavefields and the imaging in one pass (kernels2a.cuh)
        int n_stress_ops = 0;
        for (int b = 0; b < nb; b++) n_stress_ops = std::max(n_stress_ops, h->h_post_s[b].ninj);
        const bool fuse2a = h->fuse2a && pp && grad && merge_pw && !h->el && !unshifted && n_stress_ops == 0;      // stress sources act between update and imaging
        Grad2aArgs ga2{};
        if (fuse2a) {
            int n[3] = {g.nz, g.ny, g.nx}, sh[3], off[3];
            field_shape(h->nd, GPI_P, n, sh, off, h->c.order);
            ga2.gK = h->gshot; ga2.gR = h->gshot + g.vol; ga2.gstride = 2 * g.vol;
            ga2.dtI = (float)h->c.dtI; ga2.stash = h->stash_table; ga2.nb = h->c.nbound;
            const int npx = (h->c.pml_faces & XMIN) ? h->c.npml : 0, npz = (h->c.pml_faces & ZMIN) ? h->c.npml : 0;     // as launch_boundary
            ga2.xlo = npx + off[2]; ga2.xhi = sh[2] - npx - h->c.nbound + off[2];
            ga2.zlo = npz + off[0]; ga2.zhi = sh[0] - npz - h->c.nbound + off[0];
            ga2.k0 = off[0]; ga2.nk = off[0] + sh[0]; ga2.i0 = off[2]; ga2.ni = off[2] + sh[2];
        }
        if (born) { args[0].dout[0] = h->born_d; args[0].dout[1] = h->born_d + g.vol; args[0].dstride = 2 * g.vol; }
        // everything from here to the end of the batch is a fixed sequence of launches on h->stream: the body runs directly, or once
        // under stream capture and from then on as a graph launch (gpi_handle::graph_mode)
        bool graphing = false;
        gpi_handle::GraphEntry* ge = nullptr;
#ifndef GPI_HOST_EMU
        if (h->graph_mode > 0 && !h->slab && h->nd == 2) {
            // what decides the sequence besides the handle's own fixed state
            unsigned long long key = 1469598103934665603ULL;
            auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ULL; };
            mix(mode); mix(activepw); mix(src_flags); mix(shot0); mix(nb); mix(do_post_v); mix(inj_s); mix(rec_s); mix(born); mix(nt);
            mix((unsigned long long)(uintptr_t)h->stream); mix((unsigned long long)(uintptr_t)h->W); mix(pp); mix(unshifted); mix(h->illum_on); mix(h->itsnaps.size());
            for (int v : h->itsnaps) mix(v);
            for (int b = 0; b < nb; b++) for (int ipw = 0; ipw < h->npw; ipw++) for (auto p : h->shots[ipw][shot0 + b].usnaps) mix((unsigned long long)(uintptr_t)p);
            for (auto& e : h->graphs) if (e.key == key) ge = &e;
            if (ge) graphing = true;                       // seen before: capture now (if not done yet) and replay
            else {
                h->graphs.push_back({key, nullptr, 0.0, false});
                if (h->graph_mode >= 2) { ge = &h->graphs.back(); graphing = true; }
            }
        }
#endif
        auto steps = [&]() -> int {
            // record!(1, ..., [:p]) at the start of step 1 (zero unless the fields were loaded from snapshots)
            const bool pdl2 = h->pdl && h->nd == 2;
            if (rec_s) {
                if (pdl2) GPI_PDL(h, k_post, dim3(nb), dim3(128), g, h->post_s, 1, 1, nt, (float)h->c.dt, 2, 0LL, 0, 0x7fffffff);
                else k_post<<<nb, 128, 0, h->stream>>>(g, h->post_s, 1, 1, nt, (float)h->c.dt, 2, 0LL);
                h->timers.launches += 1;
            }
            // time levels: `cur` holds the fields of this step, `prev` those of the step before (adjoint runs; always W / TP without ping-pong)
            float* cur = h->W; float* prev = h->TP;
            float* const W0 = h->W;            // the descriptors and StepArgs above were built on this base
            long long woff = 0;                // cur - W0
            auto rebase = [&](const StepArgs& src, float* A, float* B, bool vel) {     // out-of-place step A -> B
                StepArgs a = src;
                for (int q = 0; q < 6; q++) if (src.tau[q]) { a.tau[q] = A + (src.tau[q] - W0); if (!vel) a.tau_o[q] = B + (src.tau[q] - W0); }
                for (int q = 0; q < 3; q++) if (src.v[q]) { a.v[q] = (vel ? A : B) + (src.v[q] - W0); if (vel) a.v_o[q] = B + (src.v[q] - W0); }
                return a;
            };

            const bool pipe = slab_pipelined(h) && mode == GPI_MODE_FORWARD && !born && h->npw == 1;
            const int xlo[2] = {0, (g.nx1 + 1) / 2}, xn[2] = {(g.nx1 + 1) / 2, g.nx1 - (g.nx1 + 1) / 2};
            for (int it = 1; it <= nt; it++) {
                float* A = cur; float* Bn = prev;      // ping-pong: this step reads level A and writes level Bn
                if (pp) {
                    // the planes boundary_force! overwrites keep their values aside: A stays behind as the previous level, which
                    // save_tp! copies BEFORE the force (propagate.jl:186-188)
                    if (launch_boundary(h, 2, nb, 0, A, h->stash_table)) return 1;
                    if (launch_boundary(h, 0, nb, nt - it, A)) return 1;
                    woff = Bn - W0;
                } else if (mode == GPI_MODE_ADJOINT) {
                    // save_tp! (save_tp.jl:5-12): one device copy of every wavefield of the batch
                    CU(h, cudaMemcpyAsync(h->TP, h->W, (size_t)nb * h->bstride * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
                    // boundary_force!(nt - it + 1) on pw 1 (propagate.jl:188), x then (y) then z
                    if (launch_boundary(h, false, nb, nt - it)) return 1;
                }
                const bool sample = h->sample_every > 0 && (it % h->sample_every) == 0 && !graphing;
                if (pipe) {
                    // pipelined z-slab step: each half step is launched as two x halves; the halo planes of a half travel on the side stream
                    // (pack -> NCCL send / recv -> unpack) while the other half -- or the first half of the next kernel -- is computed.  A
                    // kernel of half q waits for the halos the previous kernel's half q sent; sources are injected right behind the half
                    // they fall into (k_post's x-range filter) so that the edge planes leave with them.
                    for (int ph = 0; ph < 2; ph++) {
                        const bool vel = ph == 0;
                        for (int q = 0; q < 2; q++) {
                            CU(h, cudaStreamWaitEvent(h->stream, vel ? h->ev_xt[q] : h->ev_xv[q], 0));
                            h->xr_lo = xlo[q]; h->xr_n = xn[q];
                            launch_step(h, args[0], vel, nb, sample, /*half=*/true);
                            h->xr_n = 0;
                            const bool last = q == 1;
                            if (vel) {
                                if (do_post_v && (last || any_post(h->h_post_v, nb, false))) {
                                    k_post<<<nb, 128, 0, h->stream>>>(g, h->post_v, it, it, nt, (float)h->c.dt, last ? 3 : 1, woff, xlo[q], xlo[q] + xn[q]);
                                    h->timers.launches += 1;
                                }
                            } else if (inj_s || (last && rec_s && it < nt)) {
                                k_post<<<nb, 128, 0, h->stream>>>(g, h->post_s, it, it + 1, nt, (float)h->c.dt, last ? 3 : 1, woff, xlo[q], xlo[q] + xn[q]);
                                h->timers.launches += 1;
                            }
                            CU(h, cudaEventRecord(h->ev_half[q], h->stream));
                            CU(h, cudaStreamWaitEvent(h->side, h->ev_half[q], 0));
                            if (exchange_halos(h, vel ? 1 : 0, sample, xlo[q], xn[q], q, h->side)) return 1;
                            CU(h, cudaEventRecord(vel ? h->ev_xv[q] : h->ev_xt[q], h->side));
                        }
                    }
                } else {
                    if (merge_pw) launch_step(h, pp ? rebase(margs, A, Bn, true) : margs, true, margs.nbatch, sample);
                    else for (int ipw = 0; ipw < h->npw; ipw++) if (activepw & (1 << ipw)) launch_step(h, pp ? rebase(args[ipw], A, Bn, true) : args[ipw], true, nb, sample && ipw == 0);
                    if (born) {        // add_born_sources_velocity! (propagate.jl:205)
                        dim3 blk = h->blk2, grd = grid_for(h, blk, nb);
                        k_born_add<0><<<grd, blk, 0, h->stream>>>(g, args[1].v[V_X], args[1].v[V_Z], h->born_d, h->born_d + g.vol, h->born_c[1], h->born_c[2], h->bstride, 2 * g.vol);
                        h->timers.launches += 1;
                    }
                    if (do_post_v) {
                        if (pdl2) GPI_PDL(h, k_post, dim3(nb), dim3(128), g, h->post_v, it, it, nt, (float)h->c.dt, 3, woff, 0, 0x7fffffff);
                        else k_post<<<nb, 128, 0, h->stream>>>(g, h->post_v, it, it, nt, (float)h->c.dt, 3, woff);
                        h->timers.launches += 1;
                    }
                    if (exchange_halos(h, 1, sample)) return 1;
                    if (fuse2a) {
                        cudaEvent_t e0 = nullptr, e1 = nullptr;
                        if (sample) { e0 = sample_event(h); e1 = sample_event(h); }
                        if (e0 && e1) { cudaEventRecord(e0, h->stream); h->evkind.push_back(1); }
                        ga2.vxA = wf_ptr(h, A, 0, 0, GPI_VX); ga2.vzA = wf_ptr(h, A, 0, 0, GPI_VZ);
                        const int nthreads = (g.pz / VW) * g.nx1;
                        dim3 blk(128), grd((nthreads + 127) / 128, nb);
                        if (pdl2) GPI_PDL(h, k_stress2a, grd, blk, g, rebase(margs, A, Bn, false), ga2);
                        else k_stress2a<<<grd, blk, 0, h->stream>>>(g, rebase(margs, A, Bn, false), ga2);
                        if (e0 && e1) cudaEventRecord(e1, h->stream);
                        h->timers.launches += 1;
                    }
                    else if (merge_pw) launch_step(h, pp ? rebase(margs, A, Bn, false) : margs, false, margs.nbatch, sample);
                    else for (int ipw = 0; ipw < h->npw; ipw++) if (activepw & (1 << ipw)) launch_step(h, pp ? rebase(args[ipw], A, Bn, false) : args[ipw], false, nb, sample && ipw == 0);
                    if (born) {        // add_born_sources_stress! (propagate.jl:226)
                        dim3 blk = h->blk2, grd = grid_for(h, blk, nb);
                        k_born_add<1><<<grd, blk, 0, h->stream>>>(g, args[1].tau[T_XX], nullptr, h->born_d, h->born_d + g.vol, h->born_c[0], nullptr, h->bstride, 2 * g.vol);
                        h->timers.launches += 1;
                    }
                    // stress sources at step it, then the pressure record of step it+1 (record! runs at the start of a step)
                    if (inj_s || (rec_s && it < nt)) {
                        if (pdl2) GPI_PDL(h, k_post, dim3(nb), dim3(128), g, h->post_s, it, it + 1, nt, (float)h->c.dt, 3, woff, 0, 0x7fffffff);
                        else k_post<<<nb, 128, 0, h->stream>>>(g, h->post_s, it, it + 1, nt, (float)h->c.dt, 3, woff);
                        h->timers.launches += 1;
                    }
                    if (pp) {
                        // the previous level as save_tp! would have left it (the fused pass has read the pre-force values from the stash already,
                        // and level A is overwritten by the next step: no restore)
                        if (!fuse2a && launch_boundary(h, 0, nb, 0, A, h->stash_table)) return 1;
                        cur = Bn; prev = A;
                    }
                    if (exchange_halos(h, 0, sample)) return 1;
                }
                if (mode == GPI_MODE_FORWARD_SAVE && launch_boundary(h, true, nb, it - 1)) return 1;
                if (grad && h->el && h->nd == 3) {
                    const int tf[6] = {GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ, GPI_TAUXY, GPI_TAUXZ, GPI_TAUYZ};     // T_XX .. T_YZ
                    for (int b = 0; b < nb; b++) {
                        GradE3Args ga;
                        for (int q = 0; q < 6; q++) { ga.t1[q] = wf_ptr(h, cur, b, 0, tf[q]); ga.t1tp[q] = wf_ptr(h, prev, b, 0, tf[q]); ga.t2tp[q] = wf_ptr(h, prev, b, 1, tf[q]); }
                        for (int q = 0; q < 3; q++) { ga.v1[q] = wf_ptr(h, cur, b, 0, vf[q]); ga.v1tp[q] = wf_ptr(h, prev, b, 0, vf[q]); ga.v2tp[q] = wf_ptr(h, prev, b, 1, vf[q]); }
                        ga.il = h->mod[GPI_INVLAMBDA]; ga.im = h->mod[GPI_INVMU];
                        ga.gL = h->gshot + (size_t)b * 3 * g.vol; ga.gM = ga.gL + g.vol; ga.gR = ga.gL + 2 * g.vol;
                        dim3 blk = h->blk3, grd = grid_for(h, blk, 1);
                        k_grad3d_el<<<grd, blk, 0, h->stream>>>(g, ga, (float)h->c.dtI);
                        h->timers.launches += 1;
                    }
                } else if (grad && h->el) {
                    GradE2Args ga;
                    ga.xx1 = wf_ptr(h, cur, 0, 0, GPI_TAUXX); ga.zz1 = wf_ptr(h, cur, 0, 0, GPI_TAUZZ); ga.xz1 = wf_ptr(h, cur, 0, 0, GPI_TAUXZ);
                    ga.xx1tp = wf_ptr(h, prev, 0, 0, GPI_TAUXX); ga.zz1tp = wf_ptr(h, prev, 0, 0, GPI_TAUZZ); ga.xz1tp = wf_ptr(h, prev, 0, 0, GPI_TAUXZ);
                    ga.xx2tp = wf_ptr(h, prev, 0, 1, GPI_TAUXX); ga.zz2tp = wf_ptr(h, prev, 0, 1, GPI_TAUZZ); ga.xz2tp = wf_ptr(h, prev, 0, 1, GPI_TAUXZ);
                    ga.vx1 = wf_ptr(h, cur, 0, 0, GPI_VX); ga.vx1tp = wf_ptr(h, prev, 0, 0, GPI_VX); ga.vx2tp = wf_ptr(h, prev, 0, 1, GPI_VX);
                    ga.vz1 = wf_ptr(h, cur, 0, 0, GPI_VZ); ga.vz1tp = wf_ptr(h, prev, 0, 0, GPI_VZ); ga.vz2tp = wf_ptr(h, prev, 0, 1, GPI_VZ);
                    ga.il = h->mod[GPI_INVLAMBDA]; ga.im = h->mod[GPI_INVMU];
                    ga.gL = h->gshot; ga.gM = h->gshot + g.vol; ga.gR = h->gshot + 2 * g.vol;
                    ga.wstride = h->bstride; ga.gstride = 3 * g.vol;
                    dim3 blk = h->blk2, grd = grid_for(h, blk, nb);
                    k_grad2d_el<<<grd, blk, 0, h->stream>>>(g, ga, (float)h->c.dtI);
                    h->timers.launches += 1;
                } else if (grad && h->nd == 3) {
                    for (int b = 0; b < nb; b++) {
                        Grad3Args ga;
                        ga.p1 = wf_ptr(h, cur, b, 0, GPI_P); ga.p1tp = wf_ptr(h, prev, b, 0, GPI_P); ga.p2tp = wf_ptr(h, prev, b, 1, GPI_P);
                        for (int q = 0; q < 3; q++) {
                            ga.v1[q] = wf_ptr(h, cur, b, 0, vf[q]); ga.v1tp[q] = wf_ptr(h, prev, b, 0, vf[q]); ga.v2tp[q] = wf_ptr(h, prev, b, 1, vf[q]);
                        }
                        ga.gK = h->gshot + (size_t)b * 2 * g.vol; ga.gR = ga.gK + g.vol;
                        dim3 blk = h->blk3, grd = grid_for(h, blk, 1);
                        k_grad3d<<<grd, blk, 0, h->stream>>>(g, ga, (float)h->c.dtI, unshifted);
                        h->timers.launches += 1;
                    }
                } else if (grad && !fuse2a) {
                    dim3 blk = h->blk2, grd = grid_for(h, blk, nb);
                    k_grad2d<<<grd, blk, 0, h->stream>>>(g,
                        wf_ptr(h, cur, 0, 0, GPI_P), wf_ptr(h, prev, 0, 0, GPI_P), wf_ptr(h, prev, 0, 1, GPI_P),
                        wf_ptr(h, cur, 0, 0, GPI_VX), wf_ptr(h, prev, 0, 0, GPI_VX), wf_ptr(h, prev, 0, 1, GPI_VX),
                        wf_ptr(h, cur, 0, 0, GPI_VZ), wf_ptr(h, prev, 0, 0, GPI_VZ), wf_ptr(h, prev, 0, 1, GPI_VZ),
                        h->gshot, h->gshot + g.vol, (float)h->c.dtI, h->bstride, 2 * g.vol, unshifted);
                    h->timers.launches += 1;
                }
                if (h->illum_on) {      // compute_illum! (fdtd.jl:570-581; its place in the loop: propagate.jl:236): pressure of pw 1 after the stress update and sources of this step
                    k_illum<<<dim3((unsigned)((g.vol + 255) / 256), nb), 256, 0, h->stream>>>(h->illum_acc, wf_ptr(h, cur, 0, 0, GPI_P), g.vol, h->bstride);
                    h->timers.launches += 1;
                }
                if (!h->itsnaps.empty()) for (int ks = 0; ks < (int)h->itsnaps.size(); ks++) if (h->itsnaps[ks] == it)
                    for (int ipw = 0; ipw < h->npw; ipw++) if (activepw & (1 << ipw)) for (int b = 0; b < nb; b++)
                        CU(h, cudaMemcpyAsync(h->shots[ipw][shot0 + b].usnaps[ks], wf_ptr(h, cur, b, ipw, h->c.snaps_field), vb, cudaMemcpyDeviceToDevice, h->stream));
            }
            if (pipe) for (int q = 0; q < 2; q++) CU(h, cudaStreamWaitEvent(h->stream, h->ev_xt[q], 0));      // the last stress halos
            // final state for the initial-value problem of the time reversal (propagate.jl:251-258)
            if (mode == GPI_MODE_FORWARD_SAVE)
                for (int b = 0; b < nb; b++) for (int i = 0; i < nbf; i++) {
                    k_negate_copy<<<(unsigned)((g.vol + 255) / 256), 256, 0, h->stream>>>(h->shots[0][shot0 + b].snap[bf[i]], wf_ptr(h, h->W, b, 0, bf[i]), g.vol);
                    h->timers.launches += 1;
                }
            if (cur != W0) {       // ping-pong run that ended on the other set: from here on it is the wavefield set
                StepArgs fin; fill_args(h, fin, 0, nb, false, cur);
                launch_step(h, fin, true, nb);
                std::swap(h->W, h->TP);
            } else launch_step(h, args[0], true, nb);
            if (mode == GPI_MODE_FORWARD_SAVE)
                for (int b = 0; b < nb; b++) for (int i = 0; i < 3; i++) if (h->shots[0][shot0 + b].snap[vf[i]])
                    CU(h, cudaMemcpyAsync(h->shots[0][shot0 + b].snap[vf[i]], wf_ptr(h, h->W, b, 0, vf[i]), vb, cudaMemcpyDeviceToDevice, h->stream));
            // stack_illums! (fdtd.jl:556-565): shot order
            if (h->illum_on) for (int b = 0; b < nb; b++) {
                k_axpy1d<<<(unsigned)((g.vol + 255) / 256), 256, 0, h->stream>>>(h->illum_stack, h->illum_acc + (size_t)b * g.vol, g.vol);
                h->timers.launches += 1;
            }
            // sum_grads! (gradient.jl:2-11): stack in shot order
            if (grad) for (int b = 0; b < nb; b++) {
                for (int q = 0; q < h->ngrad; q++)
                    k_axpy1<<<(unsigned)((g.vol + 255) / 256), 256, 0, h->stream>>>(h->gtot[h->gparam[q]], h->gshot + ((size_t)b * h->ngrad + q) * g.vol, g.vol);
                h->timers.launches += h->ngrad;
            }
            return 0;
        };
        if (!graphing) { if (steps()) return 1; }
#ifndef GPI_HOST_EMU
        else {
            if (!ge->exec) {
                const double l0 = h->timers.launches;
                cudaGraph_t graph = nullptr;
                CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                h->capturing = true;
                float* const w_before = h->W;
                const int rc = steps();
                h->capturing = false;
                ge->swapped = h->W != w_before;
                if (ge->swapped) std::swap(h->W, h->TP);          // undone here, applied after the launch below like in every replay
                const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
                if (ce != cudaSuccess || !graph) FAIL(h, "stream capture of the time loop failed: %s", cudaGetErrorString(ce));
                cudaGraphExec_t exec = nullptr;
                const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) FAIL(h, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
                ge->exec = (void*)exec; ge->launches = h->timers.launches - l0;
                h->timers.launches = l0;
            }
            CU(h, cudaGraphLaunch((cudaGraphExec_t)ge->exec, h->stream));
            if (ge->swapped) std::swap(h->W, h->TP);
            h->timers.launches += ge->launches;
            graph_batches++;
        }
#endif
        // z-slabs: every rank holds the partial sums of the taps it owns; one sum all-reduce per record block
        if (h->slab) for (int b = 0; b < nb; b++) for (int f = 0; f < GPI_NWAVEFIELD; f++) {
            ShotData& sd = h->shots[0][shot0 + b];
            if (!sd.rec[f] || sd.nr[f] == 0) continue;
            int r = h->nccl.AllReduce(sd.rec[f], sd.rec[f], (size_t)nt * sd.nr[f], /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->comm, h->stream);
            if (r != 0) FAIL(h, "ncclAllReduce of the records failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(r) : "?");
        }
        CU(h, cudaGetLastError());
        // the pinned descriptor buffers are reused by the next batch
        CU(h, cudaStreamSynchronize(h->stream));
    }
    CU(h, cudaEventRecord(h->ev1, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->timers.run_ms = ms;
    for (size_t q = 0; q < h->evkind.size(); q++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->evpool[2 * q], h->evpool[2 * q + 1]) != cudaSuccess) continue;
        if (h->evkind[q] == 0)      { h->timers.vel_ms += t; h->timers.vel_n += 1; }
        else if (h->evkind[q] == 1) { h->timers.stress_ms += t; h->timers.stress_n += 1; }
        else if (h->evkind[q] == 3) { h->timers.vel_ms += t; h->timers.vel_n += 0.5; }          // two half launches make one kernel pass
        else if (h->evkind[q] == 4) { h->timers.stress_ms += t; h->timers.stress_n += 0.5; }
        else if (h->evkind[q] == 5) { h->timers.exch_ms += t; h->timers.exch_n += 0.5; }
        else                        { h->timers.exch_ms += t; h->timers.exch_n += 1; }
    }
    if (graph_batches > 0 && h->timers.vel_n + h->timers.stress_n == 0) {      // replayed run: the kernel figures of the last launch-by-launch run
        h->timers.vel_ms = h->kept_samples[0]; h->timers.vel_n = h->kept_samples[1]; h->timers.stress_ms = h->kept_samples[2]; h->timers.stress_n = h->kept_samples[3];
    } else if (graph_batches == 0) {
        h->kept_samples[0] = h->timers.vel_ms; h->kept_samples[1] = h->timers.vel_n; h->kept_samples[2] = h->timers.stress_ms; h->kept_samples[3] = h->timers.stress_n;
    }
    h->timers.stencil_ms = h->timers.vel_ms + h->timers.stress_ms;
    const int npw_active = ((activepw & 1) ? 1 : 0) + ((activepw & 2) ? 1 : 0);
    h->timers.steps = (double)nt * h->c.nshots;
    h->timers.cell_updates = (double)nt * h->c.nshots * npw_active * (double)(h->slab ? std::min(h->kb, g.nz) - h->ka : g.nz) * g.ny * g.nx;
    return 0;
}

// =================================================================================================
// results
// =================================================================================================
extern "C" int gpi_get_records(gpi_handle* h, int ipw, int issp, int f, float* out) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || issp < 0 || issp >= h->c.nshots) FAIL(h, "bad pw/shot index (%d, %d)", ipw, issp);
    if (f < 0 || f >= GPI_NWAVEFIELD) FAIL(h, "bad field id %d", f);
    ShotData& s = h->shots[ipw][issp];
    if (!s.rec[f]) FAIL(h, "no receivers set for field %d of shot %d", f, issp);
    CU(h, cudaMemcpyAsync(out, s.rec[f], (size_t)h->c.nt * s.nr[f] * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int gpi_records_device_ptr(gpi_handle* h, int ipw, int issp, int f, void** dptr, int64_t* nbytes) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || issp < 0 || issp >= h->c.nshots || f < 0 || f >= GPI_NWAVEFIELD) FAIL(h, "bad index");
    ShotData& s = h->shots[ipw][issp];
    if (!s.rec[f]) FAIL(h, "no receivers set for field %d of shot %d", f, issp);
    *dptr = s.rec[f]; *nbytes = (int64_t)h->c.nt * s.nr[f] * sizeof(float);
    return 0;
}
extern "C" int gpi_get_gradient(gpi_handle* h, int p, float* out) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->gtot[p]) FAIL(h, "no gradient for parameter %d (gradients exist for acoustic npw=2 experiments)", p);
    return download_field(h, h->el ? GPI_TAUXX : GPI_P, h->gtot[p], out);
}
extern "C" int gpi_gradient_device_ptr(gpi_handle* h, int p, void** dptr, int64_t* nfloats) {
    GUARD(h);
    if (p < 0 || p >= GPI_NPARAM || !h->gtot[p]) FAIL(h, "no gradient for parameter %d", p);
    *dptr = h->gtot[p]; *nfloats = h->g.vol;
    return 0;
}
extern "C" int gpi_get_snap(gpi_handle* h, int ipw, int issp, int isnap, float* out) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || issp < 0 || issp >= h->c.nshots || isnap < 0 || isnap >= h->c.nsnaps) FAIL(h, "bad snapshot index");
    return download_field(h, h->c.snaps_field, h->shots[ipw][issp].usnaps[isnap], out);
}
// source illumination (gpifdtd.h): accumulators are allocated at the first gpi_set_illum(h, 1)
extern "C" int gpi_set_illum(gpi_handle* h, int on) {
    GUARD(h);
    if (on && h->el) FAIL(h, "the illumination is defined on the pressure field: acoustic experiments only (fdtd.jl:570-581)");
    if (on && h->slab) FAIL(h, "z-slab handles do not accumulate the illumination");
    if (on && !h->illum_acc) {
        CU(h, cudaMalloc((void**)&h->illum_acc, (size_t)h->B * h->g.vol * sizeof(double)));
        CU(h, cudaMalloc((void**)&h->illum_stack, (size_t)h->g.vol * sizeof(double)));
        CU(h, cudaMemset(h->illum_stack, 0, (size_t)h->g.vol * sizeof(double)));
    }
    h->illum_on = on != 0;
    return 0;
}
extern "C" int gpi_get_illum(gpi_handle* h, double* out) {
    GUARD(h);
    if (!h->illum_stack) FAIL(h, "no illumination: call gpi_set_illum(h, 1) before gpi_run");
    return download_field(h, GPI_P, h->illum_stack, out);
}
extern "C" int gpi_get_field(gpi_handle* h, int ipw, int ib, int f, float* out) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || ib < 0 || ib >= h->B) FAIL(h, "bad pw/batch index");
    float* p = wf_ptr(h, h->W, ib, ipw, f);
    if (!p || !field_exists(h->nd, h->c.physics, f)) FAIL(h, "field %d is not a wavefield of this physics", f);
    return download_field(h, f, p, out);
}
extern "C" int gpi_set_field(gpi_handle* h, int ipw, int ib, int f, const float* in) {
    GUARD(h);
    if (ipw < 0 || ipw >= h->npw || ib < 0 || ib >= h->B) FAIL(h, "bad pw/batch index");
    float* p = wf_ptr(h, h->W, ib, ipw, f);
    if (!p || !field_exists(h->nd, h->c.physics, f)) FAIL(h, "field %d is not a wavefield of this physics", f);
    return upload_field(h, f, in, p);
}
extern "C" int gpi_get_timers(gpi_handle* h, gpi_timers* out) { if (!h || !out) return 1; *out = h->timers; return 0; }
extern "C" int gpi_kernel_family(gpi_handle* h) {
    if (!h) return -1;
    if (h->c.order == 4) return GPI_KERNELS_ORDER4;
    if (h->nd == 3) return !h->vec3 ? GPI_KERNELS_SCALAR : (h->B == 1 && tma3_eligible(h)) ? GPI_KERNELS_TMA : slab_pipelined(h) ? GPI_KERNELS_VEC4_PIPELINED : GPI_KERNELS_VEC4;
    return h->vec2 ? GPI_KERNELS_VEC4 : GPI_KERNELS_SCALAR;
}

// =================================================================================================
// NCCL (loaded lazily so the library also loads where libnccl is absent)
// =================================================================================================
namespace {
int load_nccl(gpi_handle* h, NcclApi& n) {
    if (n.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto nm : names) { n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.lib) break; }
    if (!n.lib) { if (h) h->err = "libnccl.so.2 not found"; return 1; }
    n.GetUniqueId = (int (*)(void*))dlsym(n.lib, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(n.lib, "ncclCommInitRank");
    n.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclAllReduce");
    n.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclSend");
    n.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclRecv");
    n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
    n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
    n.GetErrorString = (const char* (*)(int))dlsym(n.lib, "ncclGetErrorString");
    if (!n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.CommDestroy || !n.Send || !n.Recv || !n.GroupStart || !n.GroupEnd) {
        if (h) h->err = "libnccl is missing symbols (need ncclSend / ncclRecv / ncclGroupStart / ncclGroupEnd: NCCL >= 2.7)"; return 1;
    }
    return 0;
}
NcclApi g_nccl;
}  // namespace

extern "C" int gpi_nccl_unique_id(void* id128) {
    if (!id128 || load_nccl(nullptr, g_nccl)) return 1;
    return g_nccl.GetUniqueId(id128);
}
extern "C" int gpi_nccl_init(gpi_handle* h, const void* id128, int rank, int nranks) {
    GUARD(h);
    if (load_nccl(h, h->nccl)) return 1;
    Id128 id; memcpy(&id, id128, sizeof id);
    int r = h->nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (r != 0) FAIL(h, "ncclCommInitRank failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(r) : "?");
    h->rank = rank; h->nranks = nranks;
    return 0;
}
// the cross-worker half of sum_grads! (gradient.jl:2-11): one sum all-reduce per parameter over NVLink
extern "C" int gpi_allreduce_gradients(gpi_handle* h) {
    GUARD(h);
    Range range(h, "%s", "gpi_allreduce_gradients (sum_grads! across ranks)");
    if (!h->comm) { if (h->nranks == 1) return 0; FAIL(h, "gpi_nccl_init has not been called"); }
    CU(h, cudaEventRecord(h->ev0, h->stream));
    for (int p = 0; p < GPI_NPARAM; p++) if (h->gtot[p]) {
        int r = h->nccl.AllReduce(h->gtot[p], h->gtot[p], (size_t)h->g.vol, /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->comm, h->stream);
        if (r != 0) FAIL(h, "ncclAllReduce failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(r) : "?");
    }
    CU(h, cudaEventRecord(h->ev1, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->timers.allreduce_ms = ms;
    return 0;
}
