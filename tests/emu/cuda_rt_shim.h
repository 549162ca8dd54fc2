// Host stand-in for the CUDA runtime, so that geophyinv.jl_b200/csrc/engine.cu -- the WHOLE C-ABI library: handle life cycle, uploads,
// the time loop of gpi_run, every kernel launch -- compiles as plain C++ and runs on the CPU.  TEST INFRASTRUCTURE ONLY
// (tests/test_emu_engine.py builds it into a temporary directory as libgpifdtd_emu.so): it lets the no-GPU suite drive the very
// host code and kernels the B200 runs, through the same ctypes binding, and compare them with the oracle bit for bit.  It is not a
// CPU fallback: nothing in the package can load it (engine.py only knows libgpifdtd.so), and it exports gpi_emu_marker so that a
// loaded copy is recognisable.
//
//   * "device" memory is host memory; copies are memcpy; streams are in-order by construction (every call is synchronous); events are
//     dummies (elapsed time 0);
//   * a launch  k<<<grid, block, smem, stream>>>(args)  is rewritten by tests/emu/make_emu_engine.py into
//     emu::launch(grid, block, [&] { k(args); })  -- blocks in parallel over OpenMP threads (blocks of a launch are independent on the
//     GPU as well), the threads of a block one after the other; kernels that use __syncthreads (k_post) run the threads of a block as
//     real host threads around a barrier (emu::launch_mt);
//   * the TMA-pipelined kernels (kernels3t.cuh) run through the host forms of their PTX primitives defined there under GPI_HOST_EMU:
//     cudaGetDriverEntryPoint hands the engine an emulated cuTensorMapEncodeTiled (same argument checks as the driver's), a tensor
//     copy is a box copy with zero fill, a CTA is a serial loop over its tiles (producer, then the 128 consumer lanes), and the
//     barrier words count expected bytes so that a wrong expect_tx is caught.  What this leaves out is exactly what only the
//     hardware can show: the asynchrony of the pipeline and the warp shuffles.
#pragma once
#define EMU_TLS thread_local
#define EMU_HAVE_SYNCTHREADS 1
#define GPI_EMU_T3 1
#include <pthread.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>
namespace emu { extern thread_local pthread_barrier_t* block_barrier; }
static inline void __syncthreads() { if (emu::block_barrier) pthread_barrier_wait(emu::block_barrier); }
#include "cuda_shim.h"

// ---- types ----------------------------------------------------------------------------------------------------------------------
struct dim3 : emu_dim3 {
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) { x = a; y = b; z = c; }
};
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEnableDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };

// ---- memory ---------------------------------------------------------------------------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }

// ---- device, streams, events ----------------------------------------------------------------------------------------------------
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline unsigned __ballot_sync(unsigned, bool) { return 0xffffffffu; }      // only feeds shuffle masks, which the host forms ignore
static inline void __syncwarp() {}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline int nvtxRangePushA(const char*) { return 0; }      // NVTX (engine.cu: struct Range)
static inline int nvtxRangePop() { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- driver API: the tensor-map encoder ------------------------------------------------------------------------------------------
typedef unsigned long long cuuint64_t;
typedef unsigned int cuuint32_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
struct CUtensorMap { unsigned char opaque[128]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
// what kernels3t.cuh's host tma_box reads back (gpi::t3::EmuTensorMap has the same layout)
struct EmuTensorMapRaw { const float* base; unsigned long long dim[3]; unsigned box[3]; };
static CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                           const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    // the driver's own requirements (CUDA driver API, cuTensorMapEncodeTiled): 16-byte aligned base, strides multiples of 16 bytes,
    // box extents 1..256, inner box extent a multiple of 16 bytes, unit element strides here
    if (!out || dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT32 || rank != 3 || ((uintptr_t)base & 15)) return 1;
    for (int q = 0; q < 3; q++) if (dims[q] == 0 || box[q] == 0 || box[q] > 256 || estr[q] != 1) return 1;
    if ((box[0] * 4) % 16) return 1;
    if (strides[0] % 16 || strides[1] % 16 || strides[0] != dims[0] * 4 || strides[1] != dims[0] * dims[1] * 4) return 1;
    EmuTensorMapRaw m; m.base = (const float*)base;
    for (int q = 0; q < 3; q++) { m.dim[q] = dims[q]; m.box[q] = box[q]; }
    memset(out, 0, sizeof *out); memcpy(out, &m, sizeof m);
    return CUDA_SUCCESS;
}
static inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, int, cudaDriverEntryPointQueryResult* q) {
    const bool ok = strcmp(name, "cuTensorMapEncodeTiled") == 0;
    *fn = ok ? (void*)&emu_cuTensorMapEncodeTiled : nullptr;
    if (q) *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return ok ? cudaSuccess : cudaErrorEmu;
}

// ---- launches -------------------------------------------------------------------------------------------------------------------
namespace emu {
thread_local pthread_barrier_t* block_barrier = nullptr;
static long long launches = 0;

template <typename F> static void launch(dim3 grid, dim3 block, F&& body) {
    launches++;
    const long long nblocks = (long long)grid.x * grid.y * grid.z;
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < nblocks; q++) {
        gridDim = grid; blockDim = block;
        blockIdx.x = (unsigned)(q % grid.x); blockIdx.y = (unsigned)((q / grid.x) % grid.y); blockIdx.z = (unsigned)(q / ((long long)grid.x * grid.y));
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
            body();
        }
    }
}
// kernels with __syncthreads: the threads of a block are host threads around a barrier, blocks one after the other.  The host
// threads live in a pool that is created once (a launch per time step would otherwise spawn 128 threads every time).
struct Pool {
    std::vector<std::thread> th;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_cond_t go = PTHREAD_COND_INITIALIZER, done = PTHREAD_COND_INITIALIZER;
    unsigned long long gen = 0;  unsigned active = 0, pending = 0;
    std::function<void(unsigned)> job;
    void worker(unsigned id) {
        unsigned long long seen = 0;
        for (;;) {
            pthread_mutex_lock(&mu);
            while (gen == seen) pthread_cond_wait(&go, &mu);
            seen = gen;
            const bool mine = id < active;
            pthread_mutex_unlock(&mu);
            if (!mine) continue;
            job(id);
            pthread_mutex_lock(&mu);
            if (--pending == 0) pthread_cond_signal(&done);
            pthread_mutex_unlock(&mu);
        }
    }
    void run(unsigned n, std::function<void(unsigned)> f) {
        while (th.size() < n) { const unsigned id = (unsigned)th.size(); th.emplace_back([this, id] { worker(id); }); th.back().detach(); }
        pthread_mutex_lock(&mu);
        job = std::move(f); active = n; pending = n; gen++;
        pthread_cond_broadcast(&go);
        while (pending) pthread_cond_wait(&done, &mu);
        pthread_mutex_unlock(&mu);
    }
};
static Pool pool;
template <typename F> static void launch_mt(dim3 grid, dim3 block, F&& body) {
    launches++;
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, nthreads);
        pool.run(nthreads, [&](unsigned t) {
            gridDim = grid; blockDim = block;
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
            threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
            block_barrier = &bar;
            body();
            block_barrier = nullptr;
        });
        pthread_barrier_destroy(&bar);
    }
}
}  // namespace emu

extern "C" int gpi_emu_marker(void) { return 1; }
