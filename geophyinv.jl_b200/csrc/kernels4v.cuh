// kernels4v.cuh -- fourth-order kernels, four z cells per thread (included inside namespace gpi, after kernels4.cuh).
//
// Same arithmetic per cell as k_vel4 / k_stress4 (kernels4.cuh) -- d4() in Float64, CPML in Float32, the reference's
// association order -- so the results are bit-identical; what changes is the work per thread.  ncu on the scalar kernels
// (profiles/r01/ncu_all_kernels.md) shows DRAM traffic equal to the algorithmic bytes but 3.5x the instructions of the
// order-2 TMA kernel: index arithmetic, range predicates, slab tests and coefficient loads repeated for every cell.  Here
// one thread owns four consecutive storage cells of a z row (ks0 a multiple of four, so every x / y neighbour access is
// an aligned 128-bit load), evaluates the (j, i) predicates, the x / y slab indices and their coefficients once, moves
// x / y CPML memory as float4, and takes its z neighbours from a 12-float window (two 8-byte halo loads).
// Rigid faces stay in k_dirichlet4.

struct W12 { float w[12]; };       // storage ks0-4 .. ks0+7; only 2..9 are ever read (two halo cells on each side)
__device__ __forceinline__ W12 ldz4(const float* __restrict__ f /* at ks0 */, bool hasL, bool hasR) {
    W12 r;
    const F4 c = ld4(f);
#pragma unroll
    for (int e = 0; e < 4; e++) r.w[4 + e] = c.v[e];
    float2 l = make_float2(0.f, 0.f), h = make_float2(0.f, 0.f);
    if (hasL) l = *reinterpret_cast<const float2*>(f - 2);
    if (hasR) h = *reinterpret_cast<const float2*>(f + 4);
    r.w[0] = 0.f; r.w[1] = 0.f; r.w[2] = l.x; r.w[3] = l.y; r.w[8] = h.x; r.w[9] = h.y; r.w[10] = 0.f; r.w[11] = 0.f;
    return r;
}
// differences along z from the window: onto a half-type node (dH) / onto an integer-type node (dI), element e
__device__ __forceinline__ float dHz(const W12& w, int e, float sI) { return d4(w.w[4 + e], w.w[3 + e], w.w[2 + e], w.w[5 + e], sI); }
__device__ __forceinline__ float dIz(const W12& w, int e, float sI) { return d4(w.w[5 + e], w.w[4 + e], w.w[3 + e], w.w[6 + e], sI); }
// differences along x / y: four aligned vectors at c + m*s
__device__ __forceinline__ F4 dH4(const float* __restrict__ f, long long c, long long s, float sI) {
    const F4 a0 = ld4(f + c), am1 = ld4(f + c - s), am2 = ld4(f + c - 2 * s), ap1 = ld4(f + c + s);
    F4 r;
#pragma unroll
    for (int e = 0; e < 4; e++) r.v[e] = d4(a0.v[e], am1.v[e], am2.v[e], ap1.v[e], sI);
    return r;
}
__device__ __forceinline__ F4 dI4(const float* __restrict__ f, long long c, long long s, float sI) {
    const F4 ap1 = ld4(f + c + s), a0 = ld4(f + c), am1 = ld4(f + c - s), ap2 = ld4(f + c + 2 * s);
    F4 r;
#pragma unroll
    for (int e = 0; e < 4; e++) r.v[e] = d4(ap1.v[e], a0.v[e], am1.v[e], ap2.v[e], sI);
    return r;
}
// CPML of an x / y term on four cells: one slab index, one coefficient triple, memory as float4
template <int AXIS, int TYPE>
__device__ __forceinline__ void cpml4v(const Geom& g, const PmlTerm& t, F4& d, const bool (&m)[4], int ks0, int js, int is, int u, int n, int b) {
    const int s0 = TYPE == 0 ? 0 : (TYPE == 1 ? 1 + H4 : O4);
    const int len = TYPE == 0 ? n : (TYPE == 1 ? n - O4 : n - 2 * O4);
    const int minbit = AXIS == 1 ? YMIN : XMIN, maxbit = AXIS == 1 ? YMAX : XMAX;
    const int s = slab_index(u, s0, len, g.npml, (g.pml & minbit) != 0, (g.pml & maxbit) != 0);
    if (s < 0) return;
    const long long mi = AXIS == 2 ? (long long)ks0 + (long long)g.pz * ((long long)js + (long long)g.ny1 * s)
                                   : (long long)ks0 + (long long)g.pz * ((long long)s + 2LL * g.npml * is);
    float* mp = t.mem + (long long)b * t.bstride + mi;
    F4 mem = ld4(mp);
    const float ca = __ldg(t.a + s), cb = __ldg(t.b + s), ck = __ldg(t.kI + s);
#pragma unroll
    for (int e = 0; e < 4; e++) if (m[e]) {
        mem.v[e] = __fadd_rn(__fmul_rn(cb, mem.v[e]), __fmul_rn(ca, d.v[e]));
        d.v[e] = __fadd_rn(__fmul_rn(d.v[e], ck), mem.v[e]);
    }
    st4(mp, mem);
}
// CPML of a z term on the masked cells (slab index per cell; cpml4<0, TYPE> of kernels4.cuh)
template <int TYPE>
__device__ __forceinline__ void cpml4vz(const Geom& g, const PmlTerm& t, F4& d, const bool (&m)[4], int ks0, int js, int is, int n, int b) {
#pragma unroll
    for (int e = 0; e < 4; e++) if (m[e]) d.v[e] = cpml4<0, TYPE>(g, t, d.v[e], ks0 + e, js, is, ks0 + e - H4, n, b);
}

template <int ND>
__device__ __forceinline__ bool cell4v(const Geom& g, int& ks0, int& js, int& is, int& b) {
    ks0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (ND == 3) {
        js = blockIdx.y * blockDim.y + threadIdx.y;
        const int ntx = (g.nx1 + blockDim.z - 1) / blockDim.z;
        b = blockIdx.z / ntx;
        is = (blockIdx.z - b * ntx) * blockDim.z + threadIdx.z;
    } else {
        js = 0;
        is = blockIdx.y * blockDim.y + threadIdx.y;
        b = blockIdx.z;
    }
    return ks0 < g.pz && js < g.ny1 && is < g.nx1;
}

// The arithmetic below takes its operands derivative by derivative, and every CPML term stores its memory variables before the next
// derivative's loads can issue (possible aliasing): a thread pays one DRAM round trip (2 - 3 us under load) per derivative, nine times in
// the 3-D elastic kernels.  Requesting the thread's own lines of every array it will touch into L2 first -- and, for the differentiated
// fields, the lines one and two planes ahead, which no earlier block has touched yet -- turns those round trips into L2 hits.
// (Requesting the CPML memory variables as well made it slower again: 22.9 -> 21.6 Gcell-updates/s on the C3 grid.)
// GPI_O4_PREFETCH=0 (compile time) leaves it out.
#ifndef GPI_O4_PREFETCH
#define GPI_O4_PREFETCH 1
#endif
template <int ND, int EL>
__device__ __forceinline__ void prefetch4v(const Geom& g, const StepArgs& a, long long w, long long c, bool vel) {
#if GPI_O4_PREFETCH
    const long long sx = (long long)g.pz * g.ny1;
    float* const* diff = vel ? a.tau : a.v;        // differentiated fields: 6 stresses | 3 velocities
    float* const* upd = vel ? a.v : a.tau;         // updated fields
    const int nd = vel ? 6 : 3, nu = vel ? 3 : 6;
#pragma unroll
    for (int q = 0; q < nd; q++) if (diff[q]) { pf2(diff[q] + w + c); pf2(diff[q] + w + c + sx); pf2(diff[q] + w + c + 2 * sx); }
#pragma unroll
    for (int q = 0; q < nu; q++) if (upd[q]) pf2(upd[q] + w + c);
    if (vel) { pf2(a.c[C_BX] + c); pf2(a.c[C_BZ] + c); if (ND == 3) pf2(a.c[C_BY] + c); }
    else {
        pf2(a.c[C_K] + c);
        if (EL) { pf2(a.c[C_L] + c); pf2(a.c[C_MUXZ] + c); if (ND == 3) { pf2(a.c[C_MUXY] + c); pf2(a.c[C_MUYZ] + c); } }
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// k_vel4v
// ------------------------------------------------------------------------------------------------
template <int ND, int EL>
__global__ void __launch_bounds__(128) k_vel4v(const Geom g, const StepArgs a) {
    int ks0, js, is, b;
    if (!cell4v<ND>(g, ks0, js, is, b)) return;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, ks0, js, is);
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int j = ND == 3 ? js - H4 : 0, i = is - H4;
    const bool jJ = ND == 2 || inJ(j, ny), jH = ND == 2 || inH(j, ny);
    const bool hasL = ks0 >= 4, hasR = ks0 + 8 <= g.pz;
    bool kJ[4], kH[4], anyJ = false, anyH = false;
#pragma unroll
    for (int e = 0; e < 4; e++) { const int k = ks0 + e - H4; kJ[e] = inJ(k, nz) && ks0 + e <= g.khi; kH[e] = inH(k, nz) && ks0 + e <= g.khi; anyJ |= kJ[e]; anyH |= kH[e]; }
    const bool dox = anyJ && jJ && inH(i, nx);             // vx: (J, J, H)
    const bool doy = ND == 3 && anyJ && jH && inJ(i, nx);  // vy: (J, H, J)
    const bool doz = anyH && jJ && inJ(i, nx);             // vz: (H, J, J)
    if (!(dox || doy || doz)) return;
    prefetch4v<ND, EL>(g, a, w, c, true);

    if (!EL) {
        const float* p = a.tau[T_XX] + w;
        if (dox) {
            float* vx = a.v[V_X] + w + c;
            F4 d = dH4(p, c, sx, g.dxI);
            cpml4v<2, 1>(g, a.pv[0], d, kJ, ks0, js, is, i, nx, b);
            F4 v = ld4(vx); const F4 bb = ldg4(a.c[C_BX] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kJ[e]) v.v[e] = __fadd_rn(v.v[e], __fmul_rn(bb.v[e], d.v[e]));
            st4(vx, v);
        }
        if (doy) {
            float* vy = a.v[V_Y] + w + c;
            F4 d = dH4(p, c, sy, g.dyI);
            cpml4v<1, 1>(g, a.pv[1], d, kJ, ks0, js, is, j, ny, b);
            F4 v = ld4(vy); const F4 bb = ldg4(a.c[C_BY] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kJ[e]) v.v[e] = __fadd_rn(v.v[e], __fmul_rn(bb.v[e], d.v[e]));
            st4(vy, v);
        }
        if (doz) {
            float* vz = a.v[V_Z] + w + c;
            const W12 pw = ldz4(p + c, hasL, hasR);
            F4 d;
#pragma unroll
            for (int e = 0; e < 4; e++) d.v[e] = dHz(pw, e, g.dzI);
            cpml4vz<1>(g, a.pv[2], d, kH, ks0, js, is, nz, b);
            F4 v = ld4(vz); const F4 bb = ldg4(a.c[C_BZ] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kH[e]) v.v[e] = __fadd_rn(v.v[e], __fmul_rn(bb.v[e], d.v[e]));
            st4(vz, v);
        }
        return;
    }
    const float* txx = a.tau[T_XX] + w; const float* tzz = a.tau[T_ZZ] + w; const float* txz = a.tau[T_XZ] + w;
    if (ND == 2) {
        if (dox) {
            float* vx = a.v[V_X] + w + c;
            F4 dxx = dH4(txx, c, sx, g.dxI);                                     // @d_xi(tauxx)
            cpml4v<2, 1>(g, a.pv[0], dxx, kJ, ks0, js, is, i, nx, b);
            const W12 zw = ldz4(txz + c, hasL, hasR);
            F4 dxz;
#pragma unroll
            for (int e = 0; e < 4; e++) dxz.v[e] = dIz(zw, e, g.dzI);            // @d_za(tauxz)
            cpml4vz<2>(g, a.pv[2], dxz, kJ, ks0, js, is, nz, b);
            F4 v = ld4(vx); const F4 bb = ldg4(a.c[C_BX] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kJ[e]) v.v[e] = __fsub_rn(v.v[e], __fmul_rn(bb.v[e], __fadd_rn(dxx.v[e], dxz.v[e])));
            st4(vx, v);
        }
        if (doz) {
            float* vz = a.v[V_Z] + w + c;
            F4 dzx = dI4(txz, c, sx, g.dxI);                                     // @d_xa(tauxz)
            cpml4v<2, 2>(g, a.pv[6], dzx, kH, ks0, js, is, i, nx, b);
            const W12 zw = ldz4(tzz + c, hasL, hasR);
            F4 dzz;
#pragma unroll
            for (int e = 0; e < 4; e++) dzz.v[e] = dHz(zw, e, g.dzI);            // @d_zi(tauzz)
            cpml4vz<1>(g, a.pv[8], dzz, kH, ks0, js, is, nz, b);
            F4 v = ld4(vz); const F4 bb = ldg4(a.c[C_BZ] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kH[e]) v.v[e] = __fsub_rn(v.v[e], __fmul_rn(bb.v[e], __fadd_rn(dzx.v[e], dzz.v[e])));
            st4(vz, v);
        }
        return;
    }
    const float* tyy = a.tau[T_YY] + w; const float* txy = a.tau[T_XY] + w; const float* tyz = a.tau[T_YZ] + w;
    if (dox) {
        float* vx = a.v[V_X] + w + c;
        F4 dxx = dH4(txx, c, sx, g.dxI);                                         // @d_xi(tauxx)
        cpml4v<2, 1>(g, a.pv[0], dxx, kJ, ks0, js, is, i, nx, b);
        F4 dxy = dI4(txy, c, sy, g.dyI);                                         // @d_ya(tauxy)
        cpml4v<1, 2>(g, a.pv[1], dxy, kJ, ks0, js, is, j, ny, b);
        const W12 zw = ldz4(txz + c, hasL, hasR);
        F4 dxz;
#pragma unroll
        for (int e = 0; e < 4; e++) dxz.v[e] = dIz(zw, e, g.dzI);                // @d_za(tauxz)
        cpml4vz<2>(g, a.pv[2], dxz, kJ, ks0, js, is, nz, b);
        F4 v = ld4(vx); const F4 bb = ldg4(a.c[C_BX] + c);
#pragma unroll
        for (int e = 0; e < 4; e++) if (kJ[e]) v.v[e] = __fsub_rn(v.v[e], __fmul_rn(bb.v[e], __fadd_rn(__fadd_rn(dxx.v[e], dxy.v[e]), dxz.v[e])));
        st4(vx, v);
    }
    if (doy) {
        float* vy = a.v[V_Y] + w + c;
        F4 dyx = dI4(txy, c, sx, g.dxI);                                         // @d_xa(tauxy)
        cpml4v<2, 2>(g, a.pv[3], dyx, kJ, ks0, js, is, i, nx, b);
        F4 dyy = dH4(tyy, c, sy, g.dyI);                                         // @d_yi(tauyy)
        cpml4v<1, 1>(g, a.pv[4], dyy, kJ, ks0, js, is, j, ny, b);
        const W12 zw = ldz4(tyz + c, hasL, hasR);
        F4 dyz;
#pragma unroll
        for (int e = 0; e < 4; e++) dyz.v[e] = dIz(zw, e, g.dzI);                // @d_za(tauyz)
        cpml4vz<2>(g, a.pv[5], dyz, kJ, ks0, js, is, nz, b);
        F4 v = ld4(vy); const F4 bb = ldg4(a.c[C_BY] + c);
#pragma unroll
        for (int e = 0; e < 4; e++) if (kJ[e]) v.v[e] = __fsub_rn(v.v[e], __fmul_rn(bb.v[e], __fadd_rn(__fadd_rn(dyx.v[e], dyy.v[e]), dyz.v[e])));
        st4(vy, v);
    }
    if (doz) {
        float* vz = a.v[V_Z] + w + c;
        F4 dzx = dI4(txz, c, sx, g.dxI);                                         // @d_xa(tauxz)
        cpml4v<2, 2>(g, a.pv[6], dzx, kH, ks0, js, is, i, nx, b);
        F4 dzy = dI4(tyz, c, sy, g.dyI);                                         // @d_ya(tauyz)
        cpml4v<1, 2>(g, a.pv[7], dzy, kH, ks0, js, is, j, ny, b);
        const W12 zw = ldz4(tzz + c, hasL, hasR);
        F4 dzz;
#pragma unroll
        for (int e = 0; e < 4; e++) dzz.v[e] = dHz(zw, e, g.dzI);                // @d_zi(tauzz)
        cpml4vz<1>(g, a.pv[8], dzz, kH, ks0, js, is, nz, b);
        F4 v = ld4(vz); const F4 bb = ldg4(a.c[C_BZ] + c);
#pragma unroll
        for (int e = 0; e < 4; e++) if (kH[e]) v.v[e] = __fsub_rn(v.v[e], __fmul_rn(bb.v[e], __fadd_rn(__fadd_rn(dzx.v[e], dzy.v[e]), dzz.v[e])));
        st4(vz, v);
    }
}

// ------------------------------------------------------------------------------------------------
// k_stress4v
// ------------------------------------------------------------------------------------------------
template <int ND, int EL>
__global__ void __launch_bounds__(128) k_stress4v(const Geom g, const StepArgs a) {
    int ks0, js, is, b;
    if (!cell4v<ND>(g, ks0, js, is, b)) return;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, ks0, js, is);
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int j = ND == 3 ? js - H4 : 0, i = is - H4;
    const bool jI = ND == 2 || inI(j, ny), jJ = ND == 2 || inJ(j, ny), jH = ND == 2 || inH(j, ny);
    const bool hasL = ks0 >= 4, hasR = ks0 + 8 <= g.pz;
    bool kI[4], kJ[4], kH[4], anyI = false, anyJ = false, anyH = false;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int k = ks0 + e - H4; const bool own = ks0 + e <= g.khi;
        kI[e] = own && inI(k, nz); kJ[e] = own && inJ(k, nz); kH[e] = own && inH(k, nz);
        anyI |= kI[e]; anyJ |= kJ[e]; anyH |= kH[e];
    }
    const float* vx = a.v[V_X] + w; const float* vz = a.v[V_Z] + w; const float* vy = (ND == 3) ? a.v[V_Y] + w : nullptr;
    const bool don = anyI && jI && inI(i, nx);
    const bool fs = EL && (g.freesurf & ZMIN) != 0;
    prefetch4v<ND, EL>(g, a, w, c, false);

    if (don) {
        F4 dxx = dI4(vx, c, sx, g.dxI);                                          // @d_xa(vx)
        cpml4v<2, 0>(g, a.ps[0], dxx, kI, ks0, js, is, i, nx, b);
        F4 dyy;
#pragma unroll
        for (int e = 0; e < 4; e++) dyy.v[e] = 0.f;
        if (ND == 3) {
            dyy = dI4(vy, c, sy, g.dyI);                                         // @d_ya(vy)
            cpml4v<1, 0>(g, a.ps[1], dyy, kI, ks0, js, is, j, ny, b);
        }
        const W12 zw = ldz4(vz + c, hasL, hasR);
        F4 dzz;
#pragma unroll
        for (int e = 0; e < 4; e++) dzz.v[e] = dIz(zw, e, g.dzI);                // @d_za(vz)
        cpml4vz<0>(g, a.ps[2], dzz, kI, ks0, js, is, nz, b);
        if (!EL) {
            float* p = a.tau[T_XX] + w + c;
            F4 pc = ld4(p); const F4 K = ldg4(a.c[C_K] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kI[e]) {
                const float s = (ND == 3) ? __fadd_rn(__fadd_rn(dxx.v[e], dzz.v[e]), dyy.v[e]) : __fadd_rn(dxx.v[e], dzz.v[e]);
                pc.v[e] = __fadd_rn(pc.v[e], __fmul_rn(s, K.v[e]));
            }
            st4(p, pc);
        } else {
            float* txx = a.tau[T_XX] + w + c; float* tzz = a.tau[T_ZZ] + w + c;
            const F4 M = ldg4(a.c[C_K] + c), L = ldg4(a.c[C_L] + c);
            F4 xx = ld4(txx), zz = ld4(tzz), nzz = zz;
            if (ND == 3) {
                float* tyy = a.tau[T_YY] + w + c;
                F4 yy = ld4(tyy);
#pragma unroll
                for (int e = 0; e < 4; e++) if (kI[e]) {
                    xx.v[e]  = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dzz.v[e])));
                    yy.v[e]  = __fsub_rn(__fsub_rn(yy.v[e], __fmul_rn(M.v[e], dyy.v[e])), __fmul_rn(L.v[e], __fadd_rn(dxx.v[e], dzz.v[e])));
                    nzz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], __fadd_rn(dyy.v[e], dxx.v[e])));
                }
                st4(tyy, yy);
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) if (kI[e]) {
                    xx.v[e]  = __fsub_rn(__fsub_rn(xx.v[e], __fmul_rn(M.v[e], dxx.v[e])), __fmul_rn(L.v[e], dzz.v[e]));
                    nzz.v[e] = __fsub_rn(__fsub_rn(zz.v[e], __fmul_rn(M.v[e], dzz.v[e])), __fmul_rn(L.v[e], dxx.v[e]));
                }
            }
            // free surface (advance_elastic.jl:215-230): tauzz[1] = -tauzz[2]: unified 0 <- -(unified 1); both live in the group ks0 = 0
            // (storage 1 and 2).  Unified node 0 keeps its value when the surface is free, then takes the mirror of node 1.
            if (fs && ks0 == 0) { nzz.v[1] = zz.v[1]; if (kI[2]) nzz.v[1] = -nzz.v[2]; }
            st4(txx, xx); st4(tzz, nzz);
        }
    }
    if (!EL) return;
    // tauxz: z half, y inner, x half
    if (anyH && jJ && inH(i, nx)) {
        float* txz = a.tau[T_XZ] + w + c;
        const W12 zw = ldz4(vx + c, hasL, hasR);
        F4 dxz;
#pragma unroll
        for (int e = 0; e < 4; e++) dxz.v[e] = dHz(zw, e, g.dzI);                // @d_zi(vx)
        cpml4vz<1>(g, a.ps[5], dxz, kH, ks0, js, is, nz, b);
        F4 dzx = dH4(vz, c, sx, g.dxI);                                          // @d_xi(vz)
        cpml4v<2, 1>(g, a.ps[6], dzx, kH, ks0, js, is, i, nx, b);
        F4 t = ld4(txz); const F4 mu = ldg4(a.c[C_MUXZ] + c);
#pragma unroll
        for (int e = 0; e < 4; e++) if (kH[e]) {
            float n = __fsub_rn(t.v[e], __fmul_rn(mu.v[e], __fadd_rn(dxz.v[e], dzx.v[e])));
            if (fs && ks0 + e - H4 == 1 + H4) n = 0.f;                           // free_surface!(tauxz): index 1 of the array
            t.v[e] = n;
        }
        st4(txz, t);
    }
    if (ND == 3) {
        // tauxy: z inner, y half, x half
        if (anyJ && jH && inH(i, nx)) {
            float* txy = a.tau[T_XY] + w + c;
            F4 dxy = dH4(vx, c, sy, g.dyI);                                      // @d_yi(vx)
            cpml4v<1, 1>(g, a.ps[3], dxy, kJ, ks0, js, is, j, ny, b);
            F4 dyx = dH4(vy, c, sx, g.dxI);                                      // @d_xi(vy)
            cpml4v<2, 1>(g, a.ps[4], dyx, kJ, ks0, js, is, i, nx, b);
            F4 t = ld4(txy); const F4 mu = ldg4(a.c[C_MUXY] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kJ[e]) t.v[e] = __fsub_rn(t.v[e], __fmul_rn(mu.v[e], __fadd_rn(dxy.v[e], dyx.v[e])));
            st4(txy, t);
        }
        // tauyz: z half, y half, x inner
        if (anyH && jH && inJ(i, nx)) {
            float* tyz = a.tau[T_YZ] + w + c;
            const W12 zw = ldz4(vy + c, hasL, hasR);
            F4 dyz;
#pragma unroll
            for (int e = 0; e < 4; e++) dyz.v[e] = dHz(zw, e, g.dzI);            // @d_zi(vy)
            cpml4vz<1>(g, a.ps[7], dyz, kH, ks0, js, is, nz, b);
            F4 dzy = dH4(vz, c, sy, g.dyI);                                      // @d_yi(vz)
            cpml4v<1, 1>(g, a.ps[8], dzy, kH, ks0, js, is, j, ny, b);
            F4 t = ld4(tyz); const F4 mu = ldg4(a.c[C_MUYZ] + c);
#pragma unroll
            for (int e = 0; e < 4; e++) if (kH[e]) {
                float n = __fsub_rn(t.v[e], __fmul_rn(mu.v[e], __fadd_rn(dyz.v[e], dzy.v[e])));
                if (fs && ks0 + e - H4 == 1 + H4) n = 0.f;                       // free_surface!(tauyz)
                t.v[e] = n;
            }
            st4(tyz, t);
        }
    }
}
