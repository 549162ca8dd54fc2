// kernels4.cuh -- fourth-order staggered-grid kernels (_fd_order = 4; included inside namespace gpi).
//
// The reference selects the order at compile time (src/GeoPhyInv.jl:85-92) and gets the order-4 scheme from the
// same kernels through its macros (src/fdtd/diff2D.jl:47-97, diff3D.jl:60-134): O = order - 1 = 3 is the shift of
// `@inn` and of the `_i` macros, velocity arrays grow to n+3 nodes, half-node arrays shrink to n-3 and inner arrays
// to n-6 (src/fields.jl:92-671), rigid faces mirror two ghost nodes (dirichlet.jl:12-24), npml = 43.  Here:
//
//   * the unified box gains one node on the min side of every axis (Geom.h = 1): storage index = unified
//     coordinate + h, box = (n + 1 + 2h) per axis.  Unified coordinates keep the meaning of kernels.cuh:
//     integer-type nodes u sit at position u, half-type nodes u at u - 1/2:
//         I  tauii/p      reference index ix -> u = ix-1        range [0,   n-1]
//         V  velocities                       -> u = ix-1-h      range [-h,  n+h]
//         H  half nodes                       -> u = ix+h        range [1+h, n-1-h]
//         J  inner nodes                      -> u = ix-1+O      range [O,   n-1-O]
//   * a difference onto a half node u from integer-type neighbours is  27 f(u) - 27 f(u-1) + f(u-2) - f(u+1),
//     onto an integer node u from half-type neighbours  27 f(u+1) - 27 f(u) + f(u-1) - f(u+2)  (the 1/24 is in
//     d?I, fdtd.jl:318-319).  The `27.0` literals take the kernel number type under ParallelStencil's @parallel (wide_t in
//     kernels.cuh: Float32 by default, Float64 with -DGPI_LITERALS_F64); d4() evaluates the macro's operations one by one in that
//     type (no contraction), so the results match the reference's kernel text bit for bit (tests/test_reference_pinned.py);
//   * `@av_?i` averages are NOT centred at order 4 (their `+ 1` neighbour does not scale with the order):
//     node u of a half-type axis averages the integer nodes u-1-h and u-h.  Reproduced as upstream has it;
//   * one thread per cell, derivatives in registers, CPML in the same pass (slab-indexed coefficients for all three
//     axes, z memory rows indexed by the slab position).  Rigid faces are a separate face-sized launch per axis
//     in the reference's x, (y,) z order: with two ghost nodes per face the in-kernel mirroring of kernels.cuh
//     would need neighbours' new values.
//
// This is the correctness path for order 4: coalesced along z, every operand read once per thread from L1/L2,
// but no vectorisation or TMA staging yet (the order-2 kernels carry the roofline work).

constexpr int O4 = 3, H4 = 1;

__device__ __forceinline__ float d4(float hi1, float lo1, float lo2, float hi2, float sI) {
    const wide_t a = wsub(wmul((wide_t)hi1, (wide_t)27.0f), wmul((wide_t)lo1, (wide_t)27.0f));
    const wide_t b = wsub(wadd(a, (wide_t)lo2), (wide_t)hi2);
    return (float)wmul(b, (wide_t)sI);
}
// onto a half-type node from integer-type neighbours (`@d_?i` of tauii/p/v along a non-staggered axis)
__device__ __forceinline__ float dH(const float* __restrict__ f, long long c, long long s, float sI) {
    return d4(f[c], f[c - s], f[c - 2 * s], f[c + s], sI);
}
// onto an integer-type node from half-type neighbours (`@d_?a`)
__device__ __forceinline__ float dI(const float* __restrict__ f, long long c, long long s, float sI) {
    return d4(f[c + s], f[c], f[c - s], f[c + 2 * s], sI);
}
__device__ __forceinline__ bool inI(int u, int n) { return u >= 0 && u <= n - 1; }
__device__ __forceinline__ bool inH(int u, int n) { return u >= 1 + H4 && u <= n - 1 - H4; }
__device__ __forceinline__ bool inJ(int u, int n) { return u >= O4 && u <= n - 1 - O4; }

// CPML on one derivative value (cpml.jl:175-183).  TYPE of the derivative field along its axis: 0 I, 1 H, 2 J.
// ks, js, is: storage indices of the cell; u: unified coordinate along AXIS.
template <int AXIS, int TYPE>
__device__ __forceinline__ float cpml4(const Geom& g, const PmlTerm& t, float d, int ks, int js, int is, int u, int n, int b) {
    const int s0 = TYPE == 0 ? 0 : (TYPE == 1 ? 1 + H4 : O4);
    const int len = TYPE == 0 ? n : (TYPE == 1 ? n - O4 : n - 2 * O4);
    const int minbit = AXIS == 0 ? ZMIN : (AXIS == 1 ? YMIN : XMIN);
    const int maxbit = AXIS == 0 ? ZMAX : (AXIS == 1 ? YMAX : XMAX);
    const int s = slab_index(u, s0, len, g.npml, (g.pml & minbit) != 0, (g.pml & maxbit) != 0);
    if (s >= 0) {
        long long mi;
        if (AXIS == 2)      mi = (long long)ks + (long long)g.pz * ((long long)js + (long long)g.ny1 * s);
        else if (AXIS == 1) mi = (long long)ks + (long long)g.pz * ((long long)s + 2LL * g.npml * is);
        else                mi = (long long)s + (long long)g.pzm * ((long long)js + (long long)g.ny1 * is);
        float* mp = t.mem + (long long)b * t.bstride + mi;
        float m = *mp;
        m = __fadd_rn(__fmul_rn(__ldg(t.b + s), m), __fmul_rn(__ldg(t.a + s), d));
        *mp = m;
        d = __fadd_rn(__fmul_rn(d, __ldg(t.kI + s)), m);
    }
    return d;
}

// thread -> storage cell of the order-4 box
template <int ND>
__device__ __forceinline__ bool cell4(const Geom& g, int& k, int& j, int& i, int& b) {
    k = blockIdx.x * blockDim.x + threadIdx.x;
    if (ND == 3) {
        j = blockIdx.y * blockDim.y + threadIdx.y;
        const int ntx = (g.nx1 + blockDim.z - 1) / blockDim.z;
        b = blockIdx.z / ntx;
        i = (blockIdx.z - b * ntx) * blockDim.z + threadIdx.z;
    } else {
        j = 0;
        i = blockIdx.y * blockDim.y + threadIdx.y;
        b = blockIdx.z;
    }
    return k <= g.khi && j < g.ny1 && i < g.nx1;
}

// ------------------------------------------------------------------------------------------------
// k_vel4: update_dstress! + update_v! without the rigid faces (same term order in a.pv as k_vel)
// ------------------------------------------------------------------------------------------------
template <int ND, int EL>
__global__ void __launch_bounds__(256) k_vel4(const Geom g, const StepArgs a) {
    int ks, js, is, b;
    if (!cell4<ND>(g, ks, js, is, b)) return;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, ks, js, is);
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int k = ks - H4, j = ND == 3 ? js - H4 : 0, i = is - H4;          // unified coordinates
    const bool jJ = ND == 2 || inJ(j, ny), jH = ND == 2 || inH(j, ny);
    const bool vxin = inJ(k, nz) && jJ && inH(i, nx);
    const bool vzin = inH(k, nz) && jJ && inJ(i, nx);
    const bool vyin = ND == 3 && inJ(k, nz) && jH && inJ(i, nx);
    if (!(vxin || vyin || vzin)) return;
    float* vx = a.v[V_X] + w; float* vz = a.v[V_Z] + w; float* vy = (ND == 3) ? a.v[V_Y] + w : nullptr;

    if (!EL) {
        const float* p = a.tau[T_XX] + w;
        if (vxin) {
            float d = dH(p, c, sx, g.dxI);
            d = cpml4<2, 1>(g, a.pv[0], d, ks, js, is, i, nx, b);
            vx[c] = __fadd_rn(vx[c], __fmul_rn(__ldg(a.c[C_BX] + c), d));
        }
        if (ND == 3 && vyin) {
            float d = dH(p, c, sy, g.dyI);
            d = cpml4<1, 1>(g, a.pv[1], d, ks, js, is, j, ny, b);
            vy[c] = __fadd_rn(vy[c], __fmul_rn(__ldg(a.c[C_BY] + c), d));
        }
        if (vzin) {
            float d = dH(p, c, 1, g.dzI);
            d = cpml4<0, 1>(g, a.pv[2], d, ks, js, is, k, nz, b);
            vz[c] = __fadd_rn(vz[c], __fmul_rn(__ldg(a.c[C_BZ] + c), d));
        }
    } else if (ND == 2) {
        const float* txx = a.tau[T_XX] + w; const float* tzz = a.tau[T_ZZ] + w; const float* txz = a.tau[T_XZ] + w;
        if (vxin) {
            float dxx = dH(txx, c, sx, g.dxI);                                   // @d_xi(tauxx)
            dxx = cpml4<2, 1>(g, a.pv[0], dxx, ks, js, is, i, nx, b);
            float dxz = dI(txz, c, 1, g.dzI);                                    // @d_za(tauxz)
            dxz = cpml4<0, 2>(g, a.pv[2], dxz, ks, js, is, k, nz, b);
            vx[c] = __fsub_rn(vx[c], __fmul_rn(__ldg(a.c[C_BX] + c), __fadd_rn(dxx, dxz)));
        }
        if (vzin) {
            float dzx = dI(txz, c, sx, g.dxI);                                   // @d_xa(tauxz)
            dzx = cpml4<2, 2>(g, a.pv[6], dzx, ks, js, is, i, nx, b);
            float dzz = dH(tzz, c, 1, g.dzI);                                    // @d_zi(tauzz)
            dzz = cpml4<0, 1>(g, a.pv[8], dzz, ks, js, is, k, nz, b);
            vz[c] = __fsub_rn(vz[c], __fmul_rn(__ldg(a.c[C_BZ] + c), __fadd_rn(dzx, dzz)));
        }
    } else {
        const float* txx = a.tau[T_XX] + w; const float* tyy = a.tau[T_YY] + w; const float* tzz = a.tau[T_ZZ] + w;
        const float* txy = a.tau[T_XY] + w; const float* txz = a.tau[T_XZ] + w; const float* tyz = a.tau[T_YZ] + w;
        if (vxin) {
            float dxx = dH(txx, c, sx, g.dxI);                                   // @d_xi(tauxx)
            dxx = cpml4<2, 1>(g, a.pv[0], dxx, ks, js, is, i, nx, b);
            float dxy = dI(txy, c, sy, g.dyI);                                   // @d_ya(tauxy)
            dxy = cpml4<1, 2>(g, a.pv[1], dxy, ks, js, is, j, ny, b);
            float dxz = dI(txz, c, 1, g.dzI);                                    // @d_za(tauxz)
            dxz = cpml4<0, 2>(g, a.pv[2], dxz, ks, js, is, k, nz, b);
            vx[c] = __fsub_rn(vx[c], __fmul_rn(__ldg(a.c[C_BX] + c), __fadd_rn(__fadd_rn(dxx, dxy), dxz)));
        }
        if (vyin) {
            float dyx = dI(txy, c, sx, g.dxI);                                   // @d_xa(tauxy)
            dyx = cpml4<2, 2>(g, a.pv[3], dyx, ks, js, is, i, nx, b);
            float dyy = dH(tyy, c, sy, g.dyI);                                   // @d_yi(tauyy)
            dyy = cpml4<1, 1>(g, a.pv[4], dyy, ks, js, is, j, ny, b);
            float dyz = dI(tyz, c, 1, g.dzI);                                    // @d_za(tauyz)
            dyz = cpml4<0, 2>(g, a.pv[5], dyz, ks, js, is, k, nz, b);
            vy[c] = __fsub_rn(vy[c], __fmul_rn(__ldg(a.c[C_BY] + c), __fadd_rn(__fadd_rn(dyx, dyy), dyz)));
        }
        if (vzin) {
            float dzx = dI(txz, c, sx, g.dxI);                                   // @d_xa(tauxz)
            dzx = cpml4<2, 2>(g, a.pv[6], dzx, ks, js, is, i, nx, b);
            float dzy = dI(tyz, c, sy, g.dyI);                                   // @d_ya(tauyz)
            dzy = cpml4<1, 2>(g, a.pv[7], dzy, ks, js, is, j, ny, b);
            float dzz = dH(tzz, c, 1, g.dzI);                                    // @d_zi(tauzz)
            dzz = cpml4<0, 1>(g, a.pv[8], dzz, ks, js, is, k, nz, b);
            vz[c] = __fsub_rn(vz[c], __fmul_rn(__ldg(a.c[C_BZ] + c), __fadd_rn(__fadd_rn(dzx, dzy), dzz)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// rigid faces of one axis (dirichlet.jl:3-78): tangential components zero on the face's tauii line,
// two ghost nodes of the normal component mirrored.  One thread per (other-axes) tauii node.
// ------------------------------------------------------------------------------------------------
template <int ND>
__global__ void k_dirichlet4(const Geom g, const StepArgs a, int axis) {
    const int nn[3] = {g.nz, g.ny, g.nx};
    const int o1 = axis == 0 ? (ND == 3 ? 1 : 2) : 0;                                        // fastest other axis
    const int o2 = axis == 0 ? (ND == 3 ? 2 : -1) : (axis == 1 ? 2 : (ND == 3 ? 1 : -1));    // second other axis (or none)
    const int t1 = blockIdx.x * blockDim.x + threadIdx.x, t2 = blockIdx.y, b = blockIdx.z;
    if (t1 >= nn[o1] || (o2 >= 0 && t2 >= nn[o2])) return;
    const long long w = (long long)b * a.wstride;
    const long long str[3] = {1, g.pz, (long long)g.pz * g.ny1};
    const int hs[3] = {H4, ND == 3 ? H4 : 0, H4};                                // storage shift of integer-type nodes per axis
    // The launch indices (t1, t2) are 0-based ARRAY indices of every array the face touches (the reference runs
    // 1:n of the tauii grid and indexes vx, vy, vz alike, advance_acou.jl:50-57): along an axis where a component
    // is staggered its entry t is the velocity node t - h (storage t), elsewhere the tauii node t (storage t + h).
    auto cell_of = [&](int comp /* axis the component is staggered along */, int along_axis_storage) {
        long long q = (long long)along_axis_storage * str[axis];
        q += (long long)(t1 + (comp == o1 ? 0 : hs[o1])) * str[o1];
        if (o2 >= 0) q += (long long)(t2 + (comp == o2 ? 0 : hs[o2])) * str[o2];
        return q;
    };
    const int n = nn[axis];
    const int minbit = axis == 0 ? ZMIN : (axis == 1 ? YMIN : XMIN), maxbit = minbit << 1;
    float* vq = a.v[axis == 0 ? V_Z : (axis == 1 ? V_Y : V_X)] + w;
    for (int side = 0; side < 2; side++) {
        if (!(g.rigid & (side ? maxbit : minbit))) continue;
        // tangential components: index 1 / n along the face axis (integer-type for them) -> storage h / n-1+h
        for (int r = 0; r < 3; r++) {
            if (r == axis || (r == 1 && ND == 2)) continue;
            (a.v[r == 0 ? V_Z : (r == 1 ? V_Y : V_X)] + w)[cell_of(r, (side ? n - 1 : 0) + hs[axis])] = 0.f;
        }
        // normal component: min  v[1] = -v[4], v[2] = -v[3]  (storage 0 <- 3, 1 <- 2);
        //                   max  v[n+3] = -v[n], v[n+2] = -v[n+1]  (storage n+2 <- n-1, n+1 <- n)
        for (int ifd = 1; ifd <= 2; ifd++) {
            const int sg = side ? n + 3 - ifd : ifd - 1, ss = side ? n + ifd - 2 : 4 - ifd;
            vq[cell_of(axis, sg)] = -vq[cell_of(axis, ss)];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_stress4: update_dv! + update_stress! (+ free surface), term order in a.ps as k_stress
// ------------------------------------------------------------------------------------------------
template <int ND, int EL>
__global__ void __launch_bounds__(256) k_stress4(const Geom g, const StepArgs a) {
    int ks, js, is, b;
    if (!cell4<ND>(g, ks, js, is, b)) return;
    const long long w = (long long)b * a.wstride;
    const long long c = uidx(g, ks, js, is);
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int k = ks - H4, j = ND == 3 ? js - H4 : 0, i = is - H4;
    const float* vx = a.v[V_X] + w; const float* vz = a.v[V_Z] + w; const float* vy = (ND == 3) ? a.v[V_Y] + w : nullptr;
    const bool nin = inI(k, nz) && (ND == 2 || inI(j, ny)) && inI(i, nx);
    const bool jJ = ND == 2 || inJ(j, ny), jH = ND == 2 || inH(j, ny);

    float dxx = 0.f, dyy = 0.f, dzz = 0.f;
    if (nin) {
        dxx = dI(vx, c, sx, g.dxI);                                              // @d_xa(vx)
        dxx = cpml4<2, 0>(g, a.ps[0], dxx, ks, js, is, i, nx, b);
        if (ND == 3) {
            dyy = dI(vy, c, sy, g.dyI);                                          // @d_ya(vy)
            dyy = cpml4<1, 0>(g, a.ps[1], dyy, ks, js, is, j, ny, b);
        }
        dzz = dI(vz, c, 1, g.dzI);                                               // @d_za(vz)
        dzz = cpml4<0, 0>(g, a.ps[2], dzz, ks, js, is, k, nz, b);
    }
    if (!EL) {
        if (nin) {
            float* p = a.tau[T_XX] + w;
            const float s = (ND == 3) ? __fadd_rn(__fadd_rn(dxx, dzz), dyy) : __fadd_rn(dxx, dzz);
            p[c] = __fadd_rn(p[c], __fmul_rn(s, __ldg(a.c[C_K] + c)));
        }
        return;
    }
    const bool fs = (g.freesurf & ZMIN) != 0;
    if (nin) {
        const float M = __ldg(a.c[C_K] + c), L = __ldg(a.c[C_L] + c);
        float* txx = a.tau[T_XX] + w; float* tzz = a.tau[T_ZZ] + w;
        float nzz;
        if (ND == 3) {
            float* tyy = a.tau[T_YY] + w;
            txx[c] = __fsub_rn(__fsub_rn(txx[c], __fmul_rn(M, dxx)), __fmul_rn(L, __fadd_rn(dyy, dzz)));
            tyy[c] = __fsub_rn(__fsub_rn(tyy[c], __fmul_rn(M, dyy)), __fmul_rn(L, __fadd_rn(dxx, dzz)));
            nzz    = __fsub_rn(__fsub_rn(tzz[c], __fmul_rn(M, dzz)), __fmul_rn(L, __fadd_rn(dyy, dxx)));
        } else {
            txx[c] = __fsub_rn(__fsub_rn(txx[c], __fmul_rn(M, dxx)), __fmul_rn(L, dzz));
            nzz    = __fsub_rn(__fsub_rn(tzz[c], __fmul_rn(M, dzz)), __fmul_rn(L, dxx));
        }
        // free surface (advance_elastic.jl:215-230): tauzz[1] = -tauzz[2], written by the owner of index 2
        if (!(fs && k == 0)) tzz[c] = nzz;
        if (fs && k == 1) tzz[c - 1] = -nzz;
    }
    // tauxz: z half, y inner, x half
    if (inH(k, nz) && jJ && inH(i, nx)) {
        float* txz = a.tau[T_XZ] + w;
        float dxz = dH(vx, c, 1, g.dzI);                                         // @d_zi(vx)
        dxz = cpml4<0, 1>(g, a.ps[5], dxz, ks, js, is, k, nz, b);
        float dzx = dH(vz, c, sx, g.dxI);                                        // @d_xi(vz)
        dzx = cpml4<2, 1>(g, a.ps[6], dzx, ks, js, is, i, nx, b);
        float n = __fsub_rn(txz[c], __fmul_rn(__ldg(a.c[C_MUXZ] + c), __fadd_rn(dxz, dzx)));
        if (fs && k == 1 + H4) n = 0.f;                                          // free_surface!(tauxz): index 1 of the array
        txz[c] = n;
    }
    if (ND == 3) {
        // tauxy: z inner, y half, x half
        if (inJ(k, nz) && jH && inH(i, nx)) {
            float* txy = a.tau[T_XY] + w;
            float dxy = dH(vx, c, sy, g.dyI);                                    // @d_yi(vx)
            dxy = cpml4<1, 1>(g, a.ps[3], dxy, ks, js, is, j, ny, b);
            float dyx = dH(vy, c, sx, g.dxI);                                    // @d_xi(vy)
            dyx = cpml4<2, 1>(g, a.ps[4], dyx, ks, js, is, i, nx, b);
            txy[c] = __fsub_rn(txy[c], __fmul_rn(__ldg(a.c[C_MUXY] + c), __fadd_rn(dxy, dyx)));
        }
        // tauyz: z half, y half, x inner
        if (inH(k, nz) && jH && inJ(i, nx)) {
            float* tyz = a.tau[T_YZ] + w;
            float dyz = dH(vy, c, 1, g.dzI);                                     // @d_zi(vy)
            dyz = cpml4<0, 1>(g, a.ps[7], dyz, ks, js, is, k, nz, b);
            float dzy = dH(vz, c, sy, g.dyI);                                    // @d_yi(vz)
            dzy = cpml4<1, 1>(g, a.ps[8], dzy, ks, js, is, j, ny, b);
            float n = __fsub_rn(tyz[c], __fmul_rn(__ldg(a.c[C_MUYZ] + c), __fadd_rn(dyz, dzy)));
            if (fs && k == 1 + H4) n = 0.f;                                      // free_surface!(tauyz)
            tyz[c] = n;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_dmod4: update_dmod! + store_invav*! at order 4 (medium.jl:143-221 through the order-4 macros).
// Averages read the integer nodes u-1-h and u-h along a half-type axis (see the header).
// ------------------------------------------------------------------------------------------------
template <int ND, int EL>
__global__ void k_dmod4(const Geom g, const float* __restrict__ m0, const float* __restrict__ rho,
                        const float* __restrict__ imu, float* const* __restrict__ out, float dt) {
    int ks, js, is, b;
    if (!cell4<ND>(g, ks, js, is, b)) return;
    const long long c = uidx(g, ks, js, is);
    const long long sy = g.pz, sx = (long long)g.pz * g.ny1;
    const int nz = g.nz, ny = g.ny, nx = g.nx;
    const int k = ks - H4, j = ND == 3 ? js - H4 : 0, i = is - H4;
    const bool jJ = ND == 2 || inJ(j, ny), jH = ND == 2 || inH(j, ny);
    const long long lo = -(H4 + 1), hi = -H4;            // the two averaged integer nodes, in node steps
    if (inJ(k, nz) && jJ && inH(i, nx))                  // @av_xi(rho) on the vx interior
        out[C_BX][c] = inv_av(dt, __fadd_rn(rho[c + lo * sx], rho[c + hi * sx]), 0.5f);
    if (inH(k, nz) && jJ && inJ(i, nx))                  // @av_zi(rho)
        out[C_BZ][c] = inv_av(dt, __fadd_rn(rho[c + lo], rho[c + hi]), 0.5f);
    if (ND == 3 && inJ(k, nz) && jH && inJ(i, nx))       // @av_yi(rho)
        out[C_BY][c] = inv_av(dt, __fadd_rn(rho[c + lo * sy], rho[c + hi * sy]), 0.5f);
    const bool nin = inI(k, nz) && (ND == 2 || inI(j, ny)) && inI(i, nx);
    if (nin) {
        if (!EL) out[C_K][c] = __fmul_rn(__fdiv_rn(1.0f, m0[c]), dt);
        else {
            const float il = __fdiv_rn(1.0f, m0[c]), im = __fdiv_rn(1.0f, imu[c]);
            out[C_L][c] = __fmul_rn(il, dt);
            out[C_K][c] = __fmul_rn((float)((double)il + 2.0 * (double)im), dt);
        }
    }
    if (EL) {
        auto av4 = [&](long long sa, long long sb) {     // A[a,b] + A[a+1,b] + A[a,b+1] + A[a+1,b+1], a = first listed axis
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(imu[c + lo * sa + lo * sb], imu[c + hi * sa + lo * sb]), imu[c + lo * sa + hi * sb]), imu[c + hi * sa + hi * sb]);
            return inv_av(dt, s, 0.25f);
        };
        if (inH(k, nz) && jJ && inH(i, nx)) out[C_MUXZ][c] = av4(1, sx);                       // @av / @av_xzi: (z, x)
        if (ND == 3) {
            if (inJ(k, nz) && jH && inH(i, nx)) out[C_MUXY][c] = av4(sy, sx);                  // @av_xyi: (y, x)
            if (inH(k, nz) && jH && inJ(i, nx)) out[C_MUYZ][c] = av4(1, sy);                   // @av_yzi: (z, y)
        }
    }
}
