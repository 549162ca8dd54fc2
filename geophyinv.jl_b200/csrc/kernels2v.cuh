// kernels2v.cuh -- vectorised 2-D stencil kernels (included inside namespace gpi, after kernels3d.cuh).
//
// Same arithmetic and order of operations as vel_cell<2,*> / stress_cell<2,*> (kernels.cuh), organised like the
// 3-D k_*3v kernels: one thread owns FOUR consecutive z cells of one column, every field / coefficient / CPML
// access is an aligned 128-bit request, all loads of a thread are issued before the first dependent instruction,
// the z +-1 neighbours are one scalar load per differentiated field.  The (z, x) plane is linearised into groups
// of four so no lanes are wasted on a ragged z extent; the batch slot (resident supersource) is blockIdx.y.
// Columns on the outer shell (i < 2, i > nx-2: rigid faces, ghost cells, ragged x ranges) run the scalar code.

struct Vec2Idx { int k0, i, b; bool valid; };
__device__ __forceinline__ Vec2Idx vec2_index(const Geom& g) {
    Vec2Idx q;
    const int nq = g.pz / VW;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    q.i = gid / nq;
    q.k0 = (gid - q.i * nq) * VW;
    q.valid = q.i < g.nx1;
    q.b = blockIdx.y;
    return q;
}

// ------------------------------------------------------------------------------------------------
// velocity kernel, 2-D
// ------------------------------------------------------------------------------------------------
template <int EL, int OOP = 0>      // OOP: out of place, reads a.v and writes a.v_o (StepArgs)
__global__ void __launch_bounds__(128) k_vel2v(const Geom g, const StepArgs a) {
    pdl_wait(); pdl_release();
    const Vec2Idx q = vec2_index(g);
    if (!q.valid) return;
    const int k0 = q.k0, i = q.i, b = q.b;
    const int nz = g.nz, nx = g.nx;
    if (!(i >= 2 && i <= nx - 2)) {
#pragma unroll 1
        for (int e = 0; e < VW; e++) if (k0 + e >= g.klo && k0 + e <= g.khi) vel_cell<2, EL, OOP>(g, a, k0 + e, 0, i, b);
        return;
    }
    const int kg0 = k0 + g.koff;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, k0, 0, i) + w;
    const long long sx = g.pz;
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX;

    float* vx = (OOP ? a.v_o[V_X] : a.v[V_X]) + c; float* vz = (OOP ? a.v_o[V_Z] : a.v[V_Z]) + c;       // written
    F4 nvx = ld4(a.v[V_X] + c), nvz = ld4(a.v[V_Z] + c);
    const F4 bx = ldg4(a.c[C_BX] + c - w), bz = ldg4(a.c[C_BZ] + c - w);
    if (!EL) {
        const float* p = a.tau[T_XX] + c;
        const F4 pc = ld4(p), pmx = ld4(p - sx);
        Pml4 m0, m2;
        pml_open<2>(m0, g, a.pv[0], slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax), k0, 0, i, b);
        pml_open_z(m2, g, a.pv[2], 1, nz - 1, k0, 0, i, b);
        const float pprev = z_prev(p, k0);
        F4 dx = diff4(pc, pmx, g.dxI);        pml_apply(m0, a.pv[0], dx);
        F4 dz = diff4_zm(pc, pprev, g.dzI);   pml_apply_z(m2, a.pv[2], dz);
        pml_close(m0); pml_close(m2);
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k >= 1 && k <= nz - 2) nvx.v[e] = __fadd_rn(nvx.v[e], __fmul_rn(bx.v[e], dx.v[e]));
            if (own && k >= 1 && k <= nz - 1) nvz.v[e] = __fadd_rn(nvz.v[e], __fmul_rn(bz.v[e], dz.v[e]));
        }
    } else {
        const float* txx = a.tau[T_XX] + c; const float* tzz = a.tau[T_ZZ] + c; const float* txz = a.tau[T_XZ] + c;
        const F4 xx = ld4(txx), xxm = ld4(txx - sx), zz = ld4(tzz), xz = ld4(txz), xzpx = ld4(txz + sx);
        Pml4 m0, m2, m6, m8;
        pml_open<2>(m0, g, a.pv[0], slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax), k0, 0, i, b);     // dtauxxdx
        pml_open_z(m2, g, a.pv[2], 1, nz - 2, k0, 0, i, b);                                            // dtauxzdz
        pml_open<2>(m6, g, a.pv[6], slab_index(i, 1, nx - 2, g.npml, hxmin, hxmax), k0, 0, i, b);     // dtauxzdx
        pml_open_z(m8, g, a.pv[8], 1, nz - 1, k0, 0, i, b);                                            // dtauzzdz
        const float zzprev = z_prev(tzz, k0);
        const float xznext = z_next(txz, more);
        F4 dxx = diff4(xx, xxm, g.dxI);            pml_apply(m0, a.pv[0], dxx);
        F4 dxz = diff4_zp(xz, xznext, g.dzI);      pml_apply_z(m2, a.pv[2], dxz);
        F4 dzx = diff4(xzpx, xz, g.dxI);           pml_apply(m6, a.pv[6], dzx);
        F4 dzz = diff4_zm(zz, zzprev, g.dzI);      pml_apply_z(m8, a.pv[8], dzz);
        pml_close(m0); pml_close(m2); pml_close(m6); pml_close(m8);
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k >= 1 && k <= nz - 2) nvx.v[e] = __fsub_rn(nvx.v[e], __fmul_rn(bx.v[e], __fadd_rn(dxx.v[e], dxz.v[e])));
            if (own && k >= 1 && k <= nz - 1) nvz.v[e] = __fsub_rn(nvz.v[e], __fmul_rn(bz.v[e], __fadd_rn(dzx.v[e], dzz.v[e])));
        }
    }
    // rigid z faces (dirichlet.jl:35-74); the x faces only touch shell columns (scalar path)
    const int R = g.rigid;
    const bool head = kg0 == 0, tail = kg0 + VW - 1 >= nz - 1;
    if (head && (R & ZMIN)) { nvx.v[0] = 0.f; nvz.v[0] = -nvz.v[1]; }
    if (!tail) {
        st4(vx, nvx); st4(vz, nvz);
    } else {
#pragma unroll
        for (int e = 0; e < VW; e++) {
            const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
            if (own && k <= nz - 1) {
                const bool zero = (R & ZMAX) && k == nz - 1;
                vx[e] = zero ? 0.f : nvx.v[e];
                vz[e] = nvz.v[e];
                if ((R & ZMAX) && k == nz - 1) vz[e + 1] = -nvz.v[e];       // vz[nz+1] = -vz[nz]
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stress kernel, 2-D
// ------------------------------------------------------------------------------------------------
template <int EL, int OOP = 0>      // OOP: out of place, reads a.tau and writes a.tau_o
__global__ void __launch_bounds__(128) k_stress2v(const Geom g, const StepArgs a) {
    pdl_wait(); pdl_release();
    const Vec2Idx q = vec2_index(g);
    if (!q.valid) return;
    const int k0 = q.k0, i = q.i, b = q.b;
    const int nz = g.nz, nx = g.nx;
    if (!(i >= 1 && i <= nx - 2)) {
#pragma unroll 1
        for (int e = 0; e < VW; e++) if (k0 + e >= g.klo && k0 + e <= g.khi) stress_cell<2, EL, OOP>(g, a, k0 + e, 0, i, b);
        return;
    }
    const int kg0 = k0 + g.koff;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, k0, 0, i) + w;
    const long long sx = g.pz;
    const bool more = k0 + VW < g.pz;
    const bool hxmin = g.pml & XMIN, hxmax = g.pml & XMAX;

    const float* vx = a.v[V_X] + c; const float* vz = a.v[V_Z] + c;
    const F4 cvx = ld4(vx), cvz = ld4(vz), vxpx = ld4(vx + sx);
    Pml4 m0, m2;
    pml_open<2>(m0, g, a.ps[0], slab_index(i, 0, nx, g.npml, hxmin, hxmax), k0, 0, i, b);
    pml_open_z(m2, g, a.ps[2], 0, nz, k0, 0, i, b);
    const float vznext = z_next(vz, more);
    if (!EL) {
        float* p = (OOP ? a.tau_o[T_XX] : a.tau[T_XX]) + c;          // written
        F4 pc = ld4(a.tau[T_XX] + c);
        const F4 K = ldg4(a.c[C_K] + c - w);
        F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_apply(m0, a.ps[0], dxx);       // @d_xa(vx)
        F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_apply_z(m2, a.ps[2], dzz);     // @d_za(vz)
#pragma unroll
        for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo))
            pc.v[e] = __fadd_rn(pc.v[e], __fmul_rn(__fadd_rn(dxx.v[e], dzz.v[e]), K.v[e]));
        st4(p, pc);
        pml_close(m0); pml_close(m2);
        return;
    }
    const bool fs = (g.freesurf & ZMIN) != 0;
    float* txx = (OOP ? a.tau_o[T_XX] : a.tau[T_XX]) + c; float* tzz = (OOP ? a.tau_o[T_ZZ] : a.tau[T_ZZ]) + c;      // written
    float* txz = (OOP ? a.tau_o[T_XZ] : a.tau[T_XZ]) + c;
    F4 xx = ld4(a.tau[T_XX] + c), zz = ld4(a.tau[T_ZZ] + c), xz = ld4(a.tau[T_XZ] + c);
    const F4 M = ldg4(a.c[C_K] + c - w), L = ldg4(a.c[C_L] + c - w), mu = ldg4(a.c[C_MUXZ] + c - w);
    const F4 vzmx = ld4(vz - sx);
    Pml4 m5, m6;
    pml_open_z(m5, g, a.ps[5], 1, nz - 1, k0, 0, i, b);                                               // dvxdz
    pml_open<2>(m6, g, a.ps[6], slab_index(i, 1, nx - 1, g.npml, hxmin, hxmax), k0, 0, i, b);        // dvzdx
    const float vxprev = z_prev(vx, k0);

    F4 dxx = diff4(vxpx, cvx, g.dxI);           pml_apply(m0, a.ps[0], dxx);       // @d_xa(vx)
    F4 dzz = diff4_zp(cvz, vznext, g.dzI);      pml_apply_z(m2, a.ps[2], dzz);     // @d_za(vz)
#pragma unroll
    for (int e = 0; e < VW; e++) if (kg0 + e <= nz - 1 && (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo)) {
        xx.v[e] = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], dzz.v[e]));
        zz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], dxx.v[e]));
    }
    if (fs && kg0 == 0) zz.v[0] = -zz.v[1];                     // free_surface_mirror!: tauzz[1] = -tauzz[2]
    F4 dxz = diff4_zm(cvx, vxprev, g.dzI);      pml_apply_z(m5, a.ps[5], dxz);     // @d_zi(vx)
    F4 dzx = diff4(cvz, vzmx, g.dxI);           pml_apply(m6, a.ps[6], dzx);       // @d_xi(vz)
#pragma unroll
    for (int e = 0; e < VW; e++) {
        const int k = kg0 + e; const bool own = (unsigned)(k0 + e - g.klo) <= (unsigned)(g.khi - g.klo);
        if (own && k >= 1 && k <= nz - 1) {
            float n = __fsub_rn(xz.v[e], __fmul_rn(mu.v[e], __fadd_rn(dxz.v[e], dzx.v[e])));
            if (fs && k == 1) n = 0.f;                             // free_surface!(tauxz)
            xz.v[e] = n;
        }
    }
    st4(txx, xx); st4(tzz, zz); st4(txz, xz);
    pml_close(m0); pml_close(m2); pml_close(m5); pml_close(m6);
}
