// Minimal host shim so that the templated CUDA kernels of geophyinv.jl_b200/csrc (the ones without shared memory, barriers or PTX)
// compile as plain C++ and can be run one thread at a time on the CPU.  Test infrastructure only (tests/test_emu_kernels4.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define GPI_HOST_EMU 1
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
#ifndef EMU_TLS
#define EMU_TLS                 // emu_engine (cuda_rt_shim.h) runs blocks on several host threads: thread_local there
#endif
static EMU_TLS emu_dim3 blockIdx, blockDim, threadIdx, gridDim;
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
// round-to-nearest single operations without contraction (compile with -ffp-contract=off)
#ifdef EMU_PLAIN_ARITH      // SSE2 arithmetic is exactly rounded and -ffp-contract=off -fno-fast-math forbids fusing / reassociation
#define EMU_R(T, e) return (e)
#else
#define EMU_R(T, e) volatile T r = (e); return r
#endif
static inline float __fadd_rn(float a, float b) { EMU_R(float, a + b); }
static inline float __fsub_rn(float a, float b) { EMU_R(float, a - b); }
static inline float __fmul_rn(float a, float b) { EMU_R(float, a * b); }
static inline float __fdiv_rn(float a, float b) { EMU_R(float, a / b); }
static inline double __dadd_rn(double a, double b) { EMU_R(double, a + b); }
static inline double __dsub_rn(double a, double b) { EMU_R(double, a - b); }
static inline double __dmul_rn(double a, double b) { EMU_R(double, a * b); }
static inline double __ddiv_rn(double a, double b) { EMU_R(double, a / b); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
#ifndef EMU_HAVE_SYNCTHREADS
static inline void __syncthreads() {}      // never reached by the kernels the emulation runs (k_post is compiled, not run)
#endif
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
