/*
 * gpifdtd.h -- C ABI of the B200-native staggered-grid velocity-stress FDTD engine.
 *
 * This is the drop-in boundary for GeoPhyInv.jl's `src/fdtd` hot path: one export per seam
 * where the reference's Julia code touches device arrays.  Julia drives it through `ccall`
 * (see INTEGRATION.md and julia/GPIFdtdB200.jl); the Python mirror in
 * geophyinv.jl_b200/host drives it through ctypes.  No torch / C++ types cross this boundary:
 * plain pointers, sizes and POD structs only.
 *
 * Conventions (all inherited from the reference):
 *   - host arrays are column-major with index order [z,(y),x], z fastest
 *     (reference src/GeoPhyInv.jl:108-119); 2-D problems use n[1] (ny) = 1;
 *   - `n` is the EXTENDED grid (medium + npml cells on every PML face, fdtd.jl:137);
 *   - sparse spray / interpolation matrices arrive as CSC with 1-based row indices into the
 *     FIELD's own staggered array (reference src/fdtd/fdtd.jl:469-493, src/fdtd/ageom.jl:33-58);
 *   - records leave as column-major (nt x nr) Float32 blocks (reference src/database/database.jl:19-34);
 *   - every export returns 0 on success, non-zero on error; gpi_last_error() gives the message.
 *     Nothing throws or aborts across the ABI.
 *
 * Threading: one handle per GPU per process; calls on one handle must be serialised by the
 * caller (the reference has one in-flight remotecall per worker, propagate.jl:100-106).
 */
#ifndef GPIFDTD_H
#define GPIFDTD_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPI_ABI_VERSION 2

/* ---- enumerations -------------------------------------------------------------------------- */

/* physics, reference src/physics_types.jl:17-49 */
enum { GPI_ACOUSTIC = 0, GPI_ELASTIC = 1 };

/* attrib_mod.mode, reference src/physics_types.jl:17-49, src/fdtd/propagate.jl:38-60 */
enum { GPI_MODE_FORWARD = 0, GPI_MODE_FORWARD_SAVE = 1, GPI_MODE_ADJOINT = 2 };
/* OR into the mode of gpi_run: FD-Born scattering sources from pw 1 into pw 2 (FdtdAcoustic{Born}, born.jl:1-12) */
#define GPI_RUN_BORN 0x100
/* OR into the mode of an adjoint run: g_rho takes the velocity nodes that bound a cell instead of upstream's one-cell-shifted
 * pair (combine_gmodrho!, gradient.jl:53-56), which makes the imaging the exact transpose of the FD-Born map (LinearMap) */
#define GPI_RUN_UNSHIFTED_RHO 0x200

/* face bit masks (pml_faces / rigid_faces / stressfree_faces), reference src/fdtd/fdtd.jl:65-68 */
enum {
    GPI_ZMIN = 1, GPI_ZMAX = 2, GPI_YMIN = 4, GPI_YMAX = 8, GPI_XMIN = 16, GPI_XMAX = 32
};

/* independent medium parameters `mod`, reference src/fdtd/medium.jl:81-95 */
enum { GPI_INVK = 0, GPI_RHO = 1, GPI_INVLAMBDA = 2, GPI_INVMU = 3, GPI_NPARAM = 4 };

/* field tags, reference src/fields.jl:28-42.  Wavefields first, then derivative fields
 * (which exist in the reference as arrays and here only as CPML memory + coefficient slots). */
enum {
    GPI_P = 0, GPI_VX, GPI_VY, GPI_VZ,
    GPI_TAUXX, GPI_TAUYY, GPI_TAUZZ, GPI_TAUXY, GPI_TAUXZ, GPI_TAUYZ,
    GPI_NWAVEFIELD,                       /* = 10 */
    GPI_DPDX = GPI_NWAVEFIELD, GPI_DPDY, GPI_DPDZ,
    GPI_DVXDX, GPI_DVYDY, GPI_DVZDZ,
    GPI_DVXDY, GPI_DVXDZ, GPI_DVYDX, GPI_DVYDZ, GPI_DVZDX, GPI_DVZDY,
    GPI_DTAUXXDX, GPI_DTAUYYDY, GPI_DTAUZZDZ,
    GPI_DTAUXYDX, GPI_DTAUXYDY, GPI_DTAUXZDX, GPI_DTAUXZDZ, GPI_DTAUYZDY, GPI_DTAUYZDZ,
    GPI_NFIELD                            /* = 31 */
};

/* kind argument of gpi_set_sparse */
enum { GPI_SPRAY = 0, GPI_INTERP = 1 };

/* what argument of gpi_reset (bit mask), reference src/fdtd/types.jl:41-113,169-176 */
enum {
    GPI_RESET_WAVEFIELDS = 1,  /* reset_w2!: fields, CPML memory, buffers        */
    GPI_RESET_RECORDS    = 2,  /* initialize!(pass): records                      */
    GPI_RESET_GRADIENTS  = 4,  /* per-shot and stacked gradients                  */
    GPI_RESET_BOUNDARY   = 8,  /* initialize_boundary!: boundary stores and snaps */
    GPI_RESET_SNAPS      = 16
};

/* ---- configuration ------------------------------------------------------------------------- */

/* Everything `P_common` / `P_x_worker_x_pw` need to allocate (reference src/fdtd/fdtd.jl:61-528):
 * sizes from `ic` (fdtd.jl:301-307), float constants from `fc` (fdtd.jl:313-332; passed as the
 * Float32 values the reference stores, widened to double here so the struct is precision-agnostic). */
typedef struct gpi_config {
    int32_t abi_version;      /* GPI_ABI_VERSION */
    int32_t ndims;            /* 2 or 3                                   (_fd_ndims)  */
    int32_t physics;          /* GPI_ACOUSTIC | GPI_ELASTIC                            */
    int32_t order;            /* 2 or 4 (6/8 are broken upstream)         (_fd_order)  */
    int32_t n[3];             /* extended nz, ny, nx; ny = 1 when ndims == 2           */
    int32_t nt;               /* time steps                                            */
    int32_t npml;             /* 40 + (order-1) = 41 | 43                 (_fd_npml)   */
    int32_t nbound;           /* 3                                        (_fd_nbound) */
    int32_t pml_faces;        /* bit mask of GPI_ZMIN..GPI_XMAX                        */
    int32_t rigid_faces;      /* bit mask; the host passes unique(rigid U pml), fdtd.jl:215 */
    int32_t stressfree_faces; /* bit mask; only GPI_ZMIN acts, elastic only            */
    int32_t npw;              /* 1 or 2 propagating wavefields                         */
    int32_t nshots;           /* supersources owned by this handle (this worker's chunk) */
    int32_t store_boundary;   /* 1: allocate nt boundary slots (built in :forward_save, fdtd.jl:445-455) */
    int32_t nsnaps;           /* number of snapshot times (0 = none)                   */
    int32_t snaps_field;      /* field id to snapshot                                  */
    int32_t device;           /* CUDA device ordinal; -1 = current device              */
    int32_t shot_batch;       /* shots propagated concurrently (0 = engine decides)    */
    int32_t slab_rank;        /* z-slab domain decomposition (3-D, forward mode): this handle owns slab   */
    int32_t slab_nranks;      /* slab_rank of slab_nranks (0 or 1 = whole grid on this GPU); n[] stays GLOBAL */
    double  dt, dtI;          /* fc[:dt], fc[:dtI]                                      */
    double  d[3];             /* fc[:dz], fc[:dy], fc[:dx]                              */
    double  dI[3];            /* fc[:dzI], fc[:dyI], fc[:dxI]                           */
} gpi_config;

/* phase timers (ms), the engine-side analogue of the reference's TimerOutputs sections
 * (reference src/fdtd/propagate.jl:176-233) */
typedef struct gpi_timers {
    double run_ms;            /* whole gpi_run, device time                 */
    double steps;             /* time steps executed (all shots)            */
    double cell_updates;      /* extended-grid cell updates executed        */
    double stencil_ms;        /* device time inside the two stencil kernels */
    double launches;          /* kernels launched by the last gpi_run       */
    /* per-kernel CUDA-event samples taken inside the last gpi_run (every GPI_SAMPLE_EVERY-th step,
     * default 16; pw 1 launches only): summed duration and number of sampled launches.  Runs replayed from a CUDA graph
     * (2-D forward, INTEGRATION.md: GPI_GRAPH) take no samples and report those of the last launch-by-launch run */
    double vel_ms, vel_n;     /* fused velocity kernel  (update_dstress! + update_v!)  */
    double stress_ms, stress_n; /* fused stress kernel  (update_dv! + update_stress!)  */
    /* ABI 2: multi-GPU phases.  exch_*: z-slab halo exchanges sampled like the kernels (time the compute stream waits
     * for / spends in the exchange of one half step); allreduce_ms: device time of the last gpi_allreduce_gradients */
    double exch_ms, exch_n;
    double allreduce_ms;
} gpi_timers;

typedef struct gpi_handle gpi_handle;

/* ---- lifecycle: replaces P_x_worker_x_pw / P_x_worker_x_pw_x_ss (fdtd.jl:340-528) ----------- */
int  gpi_create(const gpi_config* cfg, gpi_handle** out);
int  gpi_destroy(gpi_handle* h);
const char* gpi_last_error(const gpi_handle* h);   /* h may be NULL: last create error */
int  gpi_abi_version(void);

/* ---- medium: replaces copyto!(mod[name], exmedium, name) + update_dmod! (medium.jl:131-221) -- */
int  gpi_set_medium(gpi_handle* h, int param_id, const float* ex_array /* [nz,(ny),nx] */);
/* same, from a z-window of the global array: rows[nk, (ny), nx] holds global rows k_first .. k_first+nk-1
 * (a slab handle only needs its own rows plus one halo row on each side; see gpi_slab_range) */
int  gpi_set_medium_rows(gpi_handle* h, int param_id, const float* rows, int k_first, int nk);
/* same, from the UN-extended array a[mz,(my),mx] (column-major): one contiguous H2D copy, then the
 * replicate padding of padarray! (media.jl:260-275) runs on the device; lo[q] = cells of padding on the
 * min face of axis q (npml on PML faces, else 0), n_in = (mz, my, mx) (my ignored in 2-D) */
int  gpi_set_medium_interior(gpi_handle* h, int param_id, const float* a, const int32_t n_in[3], const int32_t lo[3]);
/* update!(pa, medium) in one call: vp, vs (NULL for acoustic media), rho of the UN-extended medium, each [mz,(my),mx]; the derived-
 * parameter broadcasts of media.jl:103-130 (invK | invlambda, invmu, rho: Float32, inv(x) = 1/x) and the replicate padding run on the
 * device and fill every independent parameter of the physics (medium.jl:81-95).  Same n_in / lo as gpi_set_medium_interior. */
int  gpi_set_medium_fields(gpi_handle* h, const float* vp, const float* vs, const float* rho, const int32_t n_in[3], const int32_t lo[3]);
int  gpi_get_medium(gpi_handle* h, int param_id, float* out);
int  gpi_update_dmod(gpi_handle* h);
/* FD-Born (2-D acoustic): replaces copyto!(pac.δmod[name], exmedium_pert) followed by δmod .-= mod of update!(pac, medium, medium_pert)
 * (medium.jl:103-127): the perturbation of invK | rho on the extended grid, then the scattering coefficients */
int  gpi_set_medium_pert(gpi_handle* h, int param_id, const float* ex_array /* [nz, nx] */);
int  gpi_update_born(gpi_handle* h);

/* ---- CPML: replaces copyto!(pml[df][:a|:b|:kI], ...) in update_pml! (cpml.jl:100-102) -------- */
int  gpi_set_pml(gpi_handle* h, int dfield_id, const float* a, const float* b, const float* kI /* 2*npml each */);

/* ---- acquisition: replaces update!(pass, ipw, iss, ageomss, pac, Srcs|Recs) (ageom.jl:33-58) -- */
int  gpi_set_sparse(gpi_handle* h, int kind, int ipw, int issp, int field_id, int ncol,
                    const int64_t* colptr /* ncol+1, 1-based */, const int64_t* rowval /* 1-based */,
                    const float* nzval);

/* ---- wavelets: replaces fill_wavelets! (source.jl:24-58); w is [nt, ns] column-major, already
 *      transformed by get_source (source.jl:3-19) ---------------------------------------------- */
int  gpi_set_wavelets(gpi_handle* h, int ipw, int issp, int field_id, int ns, const float* w);

/* ---- the hot loop: replaces mod_x_proc!(pac, pap, activepw, src_flags) (propagate.jl:138-261)
 *      for every shot owned by the handle.  Blocking. ------------------------------------------ */
int  gpi_run(gpi_handle* h, int mode, int activepw_mask /* bit0 = pw1, bit1 = pw2 */,
             int src_flag_mask /* same bits */);

/* ---- results: replaces update_datamat! (receiver.jl:17-34), sum_grads! (gradient.jl:2-11),
 *      snaps (getprop.jl:10-25) ---------------------------------------------------------------- */
int  gpi_get_records(gpi_handle* h, int ipw, int issp, int field_id, float* out /* [nt,nr] */);
/* gradients exist for every experiment built with npw = 2: invK, rho (acoustic) | invlambda, invmu, rho (elastic); upstream images
 * 2-D acoustic media only (gradient.jl:17-61), the other methods follow the same construction */
int  gpi_get_gradient(gpi_handle* h, int param_id, float* out /* [nz,(ny),nx], summed over local shots */);
int  gpi_get_snap(gpi_handle* h, int ipw, int issp, int isnap, float* out /* snaps_field shape */);
int  gpi_set_snap_steps(gpi_handle* h, int nsnaps, const int32_t* itsnaps /* 1-based steps */);
/* Source illumination, `illum_flag` of SeisForwExpt (fdtd.jl:59,73): the energy of pw 1's pressure field, illum += abs2(p) at every time
 * step (compute_illum!, fdtd.jl:570-581; Float64 accumulation of the Float32 square) stacked over the supersources in shot order
 * (stack_illums!, fdtd.jl:556-565).  Upstream has both calls commented out (propagate.jl:114,236) and allocates a dummy (fdtd.jl:504-506);
 * this is the documented intent.  Acoustic experiments; zeroed at the start of every gpi_run like initialize!(pac) does (types.jl:171).
 * gpi_get_illum returns the extended grid ([nz,(ny),nx] of :p); the host takes the interior view as stack_illums! does. */
int  gpi_set_illum(gpi_handle* h, int on);
int  gpi_get_illum(gpi_handle* h, double* out /* [nz,(ny),nx] */);
int  gpi_get_field(gpi_handle* h, int ipw, int ibatch, int field_id, float* out /* field's own shape */);
int  gpi_set_field(gpi_handle* h, int ipw, int ibatch, int field_id, const float* in);
int  gpi_reset(gpi_handle* h, int what);

/* ---- multi-GPU: one process per GPU; the FWI gradient all-reduce that replaces the host
 *      SharedArray accumulation of sum_grads! (gradient.jl:2-11, propagate.jl:110-117) ---------- */
int  gpi_nccl_unique_id(void* id128 /* 128 bytes out */);
int  gpi_nccl_init(gpi_handle* h, const void* id128, int rank, int nranks);
int  gpi_allreduce_gradients(gpi_handle* h);
/* z-slab domain decomposition (new capability, no reference counterpart: one 3-D shot spread over the GPUs of a
 * box).  Create every handle with the same GLOBAL n[] and slab_rank / slab_nranks, call gpi_nccl_init with the
 * slab rank, then use the ordinary calls: host arrays stay global-shaped (uploads take the rows the slab needs,
 * downloads fill the rows it owns and zero the rest), gpi_run exchanges one halo plane of tauzz|p, tauxz, tauyz /
 * vx, vy, vz per half step with the z neighbours and all-reduces the records at the end.
 * gpi_slab_range: global unified z range [k_begin, k_end) owned by the handle (whole grid: 0 .. n[0]+1). */
int  gpi_slab_range(gpi_handle* h, int32_t* k_begin, int32_t* k_end);

/* ---- raw device access for zero-copy callers (CUDA.jl CuArray / torch tensors) ---------------- */
int  gpi_records_device_ptr(gpi_handle* h, int ipw, int issp, int field_id, void** dptr, int64_t* nbytes);
int  gpi_gradient_device_ptr(gpi_handle* h, int param_id, void** dptr, int64_t* nfloats);
int  gpi_set_stream(gpi_handle* h, void* cuda_stream);
int  gpi_synchronize(gpi_handle* h);

/* ---- instrumentation ------------------------------------------------------------------------- */
int  gpi_get_timers(gpi_handle* h, gpi_timers* out);
/* which stencil kernels gpi_run launches for this handle (for the bench's labels and the tests' coverage checks):
 * scalar one-thread-per-cell, float4-per-thread (k_*2v / k_*3v), TMA-pipelined 3-D elastic tiles (t3::k_step3t), order 4;
 * VEC4_PIPELINED: a z-slab handle whose launches are split into two x halves with the halo exchange on a side stream */
enum { GPI_KERNELS_SCALAR = 0, GPI_KERNELS_VEC4 = 1, GPI_KERNELS_TMA = 2, GPI_KERNELS_VEC4_PIPELINED = 3, GPI_KERNELS_ORDER4 = 4 };
int  gpi_kernel_family(gpi_handle* h);
int  gpi_field_shape(int ndims, int physics, int field_id, const int32_t n[3], int32_t out[3]);      /* order 2 */
/* shape of a field's own staggered array (fields.jl:92-671) for _fd_order = 2 | 4: velocity axes n + (order-1),
 * half-node axes n - (order-1), inner axes n - 2 (order-1) */
int  gpi_field_shape_order(int ndims, int physics, int order, int field_id, const int32_t n[3], int32_t out[3]);

#ifdef __cplusplus
}
#endif
#endif /* GPIFDTD_H */
