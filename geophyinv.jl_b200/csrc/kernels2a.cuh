// kernels2a.cuh -- 2-D acoustic adjoint step with the gradient imaging fused into the stress kernel (included inside
// namespace gpi, after kernels2v.cuh).
//
// The reference's adjoint step is save_tp! (save_tp.jl:5-12) -> boundary_force! -> update_v! -> update_stress! -> sources ->
// compute_gradient! (gradient.jl:17-56): the imaging pass re-reads p, vx, vz of this and the previous time level of both
// wavefields (36 B per cell) although the stress kernel has just had most of them in registers.  With ping-pong time levels
// (StepArgs.tau_o: level A is read, level B written) one thread of k_stress2a updates BOTH wavefields of a shot at its four
// cells and images in the same pass:
//     g_invK += p2(A) * (p1(A) - p1(B)) * dtI                                  p1(A), p2(A): the update's own operands, p1(B): its result
//     g_rho  -= 0.5 * av_x(vx2(A) * (vx1(B) - vx1(A)) * dtI) + 0.5 * av_z(...)  vx1(B), vz1(B): the update's own operands
// Per cell and adjoint step the pass moves 68 B (stress update of two wavefields 36, A-level velocities 16, gradient RMW 16)
// where k_stress2v x 2 + k_grad2d moved 92 B.  Same operations on the same values in the same order as k_grad2d -- the gradient
// stays bit-identical -- with two places where the values differ from what the stress kernel holds:
//   * boundary_force! has overwritten p1(A) on the forced planes, but save_tp! copies BEFORE the force: the imaging takes the
//     pre-force value from the stash k_boundary<2> filled (unforced());
//   * stress sources of pw 1 are added AFTER the stress update and compute_gradient! sees them: a batch that has any runs the
//     unfused kernels (upstream's adjoint runs cannot have one: get_source(w, field, Val{-1}) exists for velocity fields only,
//     source.jl:8, and velocity sources are injected before the stress pass).
struct Grad2aArgs {
    float* gK; float* gR;        // batch slot 0
    long long gstride;           // floats between the shots' gradient blocks
    const float* vxA; const float* vzA;     // level A velocities, wavefield slot 0 (pw 1 of shot 0); slots are StepArgs.wstride apart
    float dtI;
    float* const* stash;         // [shot][axis 0..2]: pre-force values of pw 1's p on the forced planes (one stored field in acoustic media)
    int nb;                      // planes per side
    int xlo, xhi, zlo, zhi;      // first unified index of the min / max triple along x and z
    int k0, nk, i0, ni;          // extents of p in unified coordinates [lower, upper)
};
// value of pw 1's p at (k, i) of shot b as save_tp! saw it: the stash where boundary_force! wrote, else the field value `raw`
__device__ __forceinline__ float unforced(const Geom& g, const Grad2aArgs& ga, int b, int k, int i, float raw) {
    int p = -1;
    if (i >= ga.xlo && i < ga.xlo + ga.nb) p = i - ga.xlo; else if (i >= ga.xhi && i < ga.xhi + ga.nb) p = ga.nb + i - ga.xhi;
    if (p >= 0 && k >= ga.k0 && k < ga.nk && k < g.pz && i >= ga.i0 && i < ga.ni) return ga.stash[b * 3 + 2][(long long)k + (long long)g.pz * p];
    p = -1;
    if (k >= ga.zlo && k < ga.zlo + ga.nb) p = k - ga.zlo; else if (k >= ga.zhi && k < ga.zhi + ga.nb) p = ga.nb + k - ga.zhi;
    if (p >= 0 && i >= ga.i0 && i < ga.ni && k >= ga.k0 && k < ga.nk) return ga.stash[b * 3 + 0][(long long)i + (long long)g.nx1 * p];
    return raw;
}
__device__ __forceinline__ bool near_forced(const Grad2aArgs& ga, int k0, int i) {
    const bool x = (i >= ga.xlo && i < ga.xlo + ga.nb) || (i >= ga.xhi && i < ga.xhi + ga.nb);
    const bool z = (k0 + VW > ga.zlo && k0 < ga.zlo + ga.nb) || (k0 + VW > ga.zhi && k0 < ga.zhi + ga.nb);
    return x || z;
}

// one launch = every resident shot (blockIdx.y), both wavefields: StepArgs `a` is the merged form (slot 2b = pw 1, 2b + 1 = pw 2)
#ifndef GPI_S2A_MINB
#define GPI_S2A_MINB 5
#endif
__global__ void __launch_bounds__(128, GPI_S2A_MINB) k_stress2a(const Geom g, const StepArgs a, const Grad2aArgs ga) {
    pdl_wait(); pdl_release();
    const Vec2Idx q = vec2_index(g);
    if (!q.valid) return;
    const int k0 = q.k0, i = q.i, b = q.b;
    const int nz = g.nz, nx = g.nx;
    const long long w1 = (long long)(2 * b) * a.wstride, w2 = w1 + a.wstride;
    const long long gw = (long long)b * ga.gstride;
    if (!(i >= 1 && i <= nx - 2)) {
        // shell columns: the scalar reference-order update, then compute_gmodKI! on the tauii nodes of the column (g_rho is @inn: not here)
#pragma unroll 1
        for (int e = 0; e < VW; e++) {
            const int k = k0 + e;
            if (k < g.klo || k > g.khi) continue;
            stress_cell<2, 0, 1>(g, a, k, 0, i, 2 * b);
            stress_cell<2, 0, 1>(g, a, k, 0, i, 2 * b + 1);
            if (k <= nz - 1 && i <= nx - 1) {
                const long long c = uidx(g, k, 0, i);
                const float p1tp = unforced(g, ga, b, k, i, a.tau[T_XX][c + w1]);
                ga.gK[gw + c] = __fadd_rn(ga.gK[gw + c], __fmul_rn(__fmul_rn(a.tau[T_XX][c + w2], __fsub_rn(p1tp, a.tau_o[T_XX][c + w1])), ga.dtI));
            }
        }
        return;
    }
    const int kg0 = k0 + g.koff;
    const long long c0 = uidx(g, k0, 0, i);
    const long long sx = g.pz;
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX;
    const int sxi = slab_index(i, 0, nx, g.npml, hxmin, hxmax);

#ifndef GPI_S2A_PHASED
#define GPI_S2A_PHASED 1
#endif
    // ---- phase 1 loads: the stress update of both wavefields and g_invK (stores would fence later loads, so they come first)
    const float* vx1 = a.v[V_X] + c0 + w1; const float* vz1 = a.v[V_Z] + c0 + w1;       // level B (this step's velocities)
    const float* vx2 = a.v[V_X] + c0 + w2; const float* vz2 = a.v[V_Z] + c0 + w2;
    const float* vxA1 = ga.vxA + c0 + w1; const float* vzA1 = ga.vzA + c0 + w1;        // level A (previous step)
    const float* vxA2 = ga.vxA + c0 + w2; const float* vzA2 = ga.vzA + c0 + w2;
    const F4 cvx1 = ld4(vx1), cvz1 = ld4(vz1), vxpx1 = ld4(vx1 + sx);
    const F4 cvx2 = ld4(vx2), cvz2 = ld4(vz2), vxpx2 = ld4(vx2 + sx);
    const float vznext1 = z_next(vz1, more), vznext2 = z_next(vz2, more);
    Pml4 m01, m21, m02, m22;
    pml_open<2>(m01, g, a.ps[0], sxi, k0, 0, i, 2 * b);      pml_open_z(m21, g, a.ps[2], 0, nz, k0, 0, i, 2 * b);
    pml_open<2>(m02, g, a.ps[0], sxi, k0, 0, i, 2 * b + 1);  pml_open_z(m22, g, a.ps[2], 0, nz, k0, 0, i, 2 * b + 1);
    F4 p1 = ld4(a.tau[T_XX] + c0 + w1), p2 = ld4(a.tau[T_XX] + c0 + w2);               // level A
    const F4 K = ldg4(a.c[C_K] + c0);
    F4 gK = ld4(ga.gK + gw + c0);
#if GPI_S2A_PHASED
    // phase 2 operands (g_rho: level A velocities of both wavefields at this column and the one before, the gradient itself) are
    // requested into L2 now and loaded after the phase 1 stores: holding all 26 vectors at once costs 128 registers (4 blocks per SM)
    pf2(vxA1); pf2(vxA2); pf2(vzA1); pf2(vzA2); pf2(ga.gR + gw + c0);
#else
    const F4 vx1m = ld4(vx1 - sx);
    const F4 vx1tp = ld4(vxA1), vx1tpm = ld4(vxA1 - sx), vx2tp = ld4(vxA2), vx2tpm = ld4(vxA2 - sx);
    const F4 vz1tp = ld4(vzA1), vz2tp = ld4(vzA2);
    const float vz1prev = z_prev(vz1, k0), vz1tpprev = z_prev(vzA1, k0), vz2tpprev = z_prev(vzA2, k0);
    F4 gR = ld4(ga.gR + gw + c0);
#endif

    // ---- update_stress! of both wavefields (k_stress2v<0, 1>)
    const F4 p1tp_forced = p1, p2tp = p2;
    {
        F4 dxx = diff4(vxpx1, cvx1, g.dxI);          pml_apply(m01, a.ps[0], dxx);       // @d_xa(vx)
        F4 dzz = diff4_zp(cvz1, vznext1, g.dzI);     pml_apply_z(m21, a.ps[2], dzz);     // @d_za(vz)
#pragma unroll
        for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo))
            p1.v[e] = __fadd_rn(p1.v[e], __fmul_rn(__fadd_rn(dxx.v[e], dzz.v[e]), K.v[e]));
    }
    {
        F4 dxx = diff4(vxpx2, cvx2, g.dxI);          pml_apply(m02, a.ps[0], dxx);
        F4 dzz = diff4_zp(cvz2, vznext2, g.dzI);     pml_apply_z(m22, a.ps[2], dzz);
#pragma unroll
        for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo))
            p2.v[e] = __fadd_rn(p2.v[e], __fmul_rn(__fadd_rn(dxx.v[e], dzz.v[e]), K.v[e]));
    }
    // ---- compute_gradient! (k_grad2d, shifted form): this = level B, previous = level A as save_tp! saw it
    F4 p1tp = p1tp_forced;
    if (near_forced(ga, k0, i)) {
#pragma unroll
        for (int e = 0; e < VW; e++) p1tp.v[e] = unforced(g, ga, b, k0 + e, i, p1tp_forced.v[e]);
    }
#pragma unroll
    for (int e = 0; e < VW; e++)
        if (k0 + e <= nz - 1) gK.v[e] = __fadd_rn(gK.v[e], __fmul_rn(__fmul_rn(p2tp.v[e], __fsub_rn(p1tp.v[e], p1.v[e])), ga.dtI));
#if GPI_S2A_PHASED
    st4(a.tau_o[T_XX] + c0 + w1, p1); st4(a.tau_o[T_XX] + c0 + w2, p2);
    pml_close(m01); pml_close(m21); pml_close(m02); pml_close(m22);
    st4(ga.gK + gw + c0, gK);
    // ---- phase 2 loads (L2 hits: requested above; the column before is the neighbouring thread's own line)
    const F4 vx1m = ld4(vx1 - sx);
    const F4 vx1tp = ld4(vxA1), vx1tpm = ld4(vxA1 - sx), vx2tp = ld4(vxA2), vx2tpm = ld4(vxA2 - sx);
    const F4 vz1tp = ld4(vzA1), vz2tp = ld4(vzA2);
    const float vz1prev = z_prev(vz1, k0), vz1tpprev = z_prev(vzA1, k0), vz2tpprev = z_prev(vzA2, k0);
    F4 gR = ld4(ga.gR + gw + c0);
#endif
    float bz[VW + 1];       // vzbuffer at k0 - 1 .. k0 + 3
    bz[0] = __fmul_rn(__fmul_rn(vz2tpprev, __fsub_rn(vz1prev, vz1tpprev)), ga.dtI);
#pragma unroll
    for (int e = 0; e < VW; e++) bz[e + 1] = __fmul_rn(__fmul_rn(vz2tp.v[e], __fsub_rn(cvz1.v[e], vz1tp.v[e])), ga.dtI);
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = k0 + e;
        if (k >= 1 && k <= nz - 2) {       // @inn(g) along z; i is in [1, nx-2] here
            const float bxm = __fmul_rn(__fmul_rn(vx2tpm.v[e], __fsub_rn(vx1m.v[e], vx1tpm.v[e])), ga.dtI);
            const float bxc = __fmul_rn(__fmul_rn(vx2tp.v[e], __fsub_rn(cvx1.v[e], vx1tp.v[e])), ga.dtI);
            const float ax = __fadd_rn(bxm, bxc);                 // @av_xi(vxbuffer): nodes i-1, i
            const float az = __fadd_rn(bz[e], bz[e + 1]);         // @av_zi(vzbuffer): nodes k-1, k
            gR.v[e] = (float)wsub(wsub((wide_t)gR.v[e], wmul((wide_t)ax, (wide_t)0.5f)), wmul((wide_t)az, (wide_t)0.5f));
        }
    }
#if !GPI_S2A_PHASED
    st4(a.tau_o[T_XX] + c0 + w1, p1); st4(a.tau_o[T_XX] + c0 + w2, p2);
    pml_close(m01); pml_close(m21); pml_close(m02); pml_close(m22);
    st4(ga.gK + gw + c0, gK);
#endif
    st4(ga.gR + gw + c0, gR);
}
