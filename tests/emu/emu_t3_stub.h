// Inert stand-ins for the names of kernels3t.cuh that engine.cu mentions (test infrastructure, see cuda_rt_shim.h): the TMA-pipelined
// kernels are mbarrier / cp.async.bulk.tensor PTX and have no host form.  cudaGetDriverEntryPoint fails in the emulated runtime, so
// gpi_create clears handle->tma3 and none of this is ever reached; it only has to compile.
#pragma once
#define T3_MINB 2
namespace gpi { namespace t3 {
constexpr int R = 4, ZC = 128, PH = ZC + 8, PZM = 96, NTHREADS = 160, V_NBOX = 15, S_NBOX = 17, MAXBOX = S_NBOX + 9;
struct BoxSpec { int arr; int di, dj, rows, halo, off; };
inline BoxSpec box_spec(int, int) { return BoxSpec{0, 0, 0, 4, 0, 0}; }
inline int term_index(int, int, int) { return 0; }
struct Maps { unsigned char m[MAXBOX][128]; };
struct Sched { int ilo, ihi, jlo, jhi, njb, nzc, ntiles, sp[4], nsp, sr[4], nsr; };
inline size_t smem_bytes(int) { return 0; }
template <int KIND> void k_step3t(const Geom&, const StepArgs&, const Sched&, const Maps*) { abort(); }
template <int KIND> void k_shell3(const Geom&, const StepArgs&, const Sched&) { abort(); }
} }
