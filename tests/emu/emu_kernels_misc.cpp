// CPU emulation, bounds and round-trip checks of the kernels that have no second implementation to compare with:
// the elastic / 3-D gradient imaging kernels (orders 2 and 4: every neighbour offset must stay inside the operand arrays --
// the harness runs under AddressSanitizer with exactly-sized vectors) and the batched boundary store (save, wipe, force:
// the stored planes come back negated, everything else stays wiped), 3-D elastic with six fields and 2-D with a shot batch.
#include "cuda_shim.h"
#include "../../geophyinv.jl_b200/csrc/kernels.cuh"
#include <cstdio>
#include <random>
#include <vector>
using namespace gpi;

template <typename K, typename A, typename... R> static void launch(K kernel, emu_dim3 grid, emu_dim3 block, const Geom& g, const A& a, R... rest) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz; threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
            kernel(g, a, rest...);
        }
}
static Geom geom(int nd, int nz, int ny, int nx, int h, int npml, int faces) {
    Geom g; memset(&g, 0, sizeof g);
    g.nz = nz; g.ny = nd == 3 ? ny : 1; g.nx = nx; g.h = h; g.klo = 0; g.khi = nz + 2 * h;
    g.pz = ((g.khi + 2 + 31) / 32) * 32; g.ny1 = nd == 3 ? ny + 1 + 2 * h : 1; g.nx1 = nx + 1 + 2 * h; g.npml = npml;
    g.pzm = 32; g.pml = faces; g.rigid = faces; g.vol = (long long)g.pz * g.ny1 * g.nx1;
    return g;
}
static emu_dim3 d3(unsigned x, unsigned y, unsigned z) { emu_dim3 d; d.x = x; d.y = y; d.z = z; return d; }

static int grad_case(int nd, int h, unsigned seed) {
    const Geom g = geom(nd, 19, 14, 17, h, 3, 0);
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(-1.f, 1.f);
    auto rnd = [&](size_t n) { std::vector<float> v(n); for (auto& x : v) x = U(rng); return v; };
    const size_t vol = (size_t)g.vol;
    const int nb = nd == 2 ? 2 : 1;
    std::vector<float> W = rnd(nb * 18 * vol), TP = rnd(nb * 18 * vol), il = rnd(vol), im = rnd(vol), G3(nb * 3 * vol, 0.f);
    for (auto& x : il) x = 1.5f + 0.4f * x; for (auto& x : im) x = 2.0f + 0.5f * x;
    size_t touched = 0;
    if (nd == 2) {
        GradE2Args a;
        const float* w1 = W.data(); const float* tp1 = TP.data(); const float* tp2 = TP.data() + 9 * vol;      // [b][pw][slot]
        a.xx1 = w1; a.zz1 = w1 + vol; a.xz1 = w1 + 2 * vol; a.xx1tp = tp1; a.zz1tp = tp1 + vol; a.xz1tp = tp1 + 2 * vol;
        a.xx2tp = tp2; a.zz2tp = tp2 + vol; a.xz2tp = tp2 + 2 * vol;
        a.vx1 = w1 + 3 * vol; a.vx1tp = tp1 + 3 * vol; a.vx2tp = tp2 + 3 * vol; a.vz1 = w1 + 4 * vol; a.vz1tp = tp1 + 4 * vol; a.vz2tp = tp2 + 4 * vol;
        a.il = il.data(); a.im = im.data(); a.gL = G3.data(); a.gM = G3.data() + vol; a.gR = G3.data() + 2 * vol;
        a.wstride = 18 * vol; a.gstride = 3 * vol;
        launch(k_grad2d_el, d3((g.khi + 16) / 16, (g.nx1 + 1) / 2, nb), d3(16, 2, 1), g, a, 0.5f);
    } else {
        GradE3Args a;
        for (int q = 0; q < 6; q++) { a.t1[q] = W.data() + q * vol; a.t1tp[q] = TP.data() + q * vol; a.t2tp[q] = TP.data() + (9 + q) * vol; }
        for (int q = 0; q < 3; q++) { a.v1[q] = W.data() + (6 + q) * vol; a.v1tp[q] = TP.data() + (6 + q) * vol; a.v2tp[q] = TP.data() + (15 + q) * vol; }
        a.il = il.data(); a.im = im.data(); a.gL = G3.data(); a.gM = G3.data() + vol; a.gR = G3.data() + 2 * vol;
        launch(k_grad3d_el, d3((g.khi + 16) / 16, (g.ny1 + 1) / 2, (g.nx1 + 1) / 2), d3(16, 2, 2), g, a, 0.5f);
        Grad3Args b;                                                                                            // acoustic 3-D, both variants
        b.p1 = W.data(); b.p1tp = TP.data(); b.p2tp = TP.data() + 9 * vol;
        for (int q = 0; q < 3; q++) { b.v1[q] = W.data() + (6 + q) * vol; b.v1tp[q] = TP.data() + (6 + q) * vol; b.v2tp[q] = TP.data() + (15 + q) * vol; }
        std::vector<float> G2(2 * vol, 0.f); b.gK = G2.data(); b.gR = G2.data() + vol;
        launch(k_grad3d, d3((g.khi + 16) / 16, (g.ny1 + 1) / 2, (g.nx1 + 1) / 2), d3(16, 2, 2), g, b, 0.5f, 0);
        if (h == 0) launch(k_grad3d, d3((g.khi + 16) / 16, (g.ny1 + 1) / 2, (g.nx1 + 1) / 2), d3(16, 2, 2), g, b, 0.5f, 1);
        for (float x : G2) touched += x != 0.f;
    }
    for (float x : G3) touched += x != 0.f;
    printf("  gradient imaging %d-D, order %d: %zu gradient entries written\n", nd, 2 + 2 * h, touched);
    return touched == 0;
}

static int boundary_case(int nd, int h, unsigned seed) {
    const int npml = 3, nbound = 3, nb = nd == 2 ? 3 : 1, nf = nd == 3 ? 6 : 3;
    const int faces = ZMIN | ZMAX | XMIN | XMAX | (nd == 3 ? YMIN | YMAX : 0);
    const Geom g = geom(nd, 26, 21, 24, h, npml, faces);
    const size_t vol = (size_t)g.vol;
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(0.5f, 1.5f);
    std::vector<float> W((size_t)nb * nf * vol); for (auto& x : W) x = U(rng);
    const std::vector<float> W0 = W;
    // node types of the stored stresses along (z, y, x): tauxx/yy/zz integer; tauxy (J,H,H); tauxz (H,J,H); tauyz (H,H,J)  (2-D: xx, xz (H,H), zz)
    const char* types3[6] = {"III", "III", "III", "JHH", "HJH", "HHJ"}; const char* types2[3] = {"II", "HH", "II"};
    const int O = 1 + 2 * h;
    auto len = [&](char t, int n) { return t == 'I' ? n : t == 'H' ? n - O : n - 2 * O; };
    auto off = [&](char t) { return t == 'I' ? h : t == 'H' ? 1 + 2 * h : 1 + 3 * h; };
    BndArgs a; memset(&a, 0, sizeof a);
    a.nf = nf; a.nbound = nbound; a.naxes = nd;
    if (nd == 3) { a.axes[0] = 2; a.axes[1] = 1; a.axes[2] = 0; } else { a.axes[0] = 2; a.axes[1] = 0; }
    const int nn[3] = {g.nz, g.ny, g.nx};
    long long slot[3] = {(long long)g.nx1 * g.ny1 * 2 * nbound, (long long)g.pz * 2 * nbound * g.nx1, (long long)g.pz * g.ny1 * 2 * nbound};
    std::vector<std::vector<float>> stores;
    std::vector<float*> table((size_t)nb * nf * 3, nullptr);
    for (int b = 0; b < nb; b++) for (int f = 0; f < nf; f++) for (int q = 0; q < 3; q++) {
        if (nd == 2 && q == 1) continue;
        stores.emplace_back((size_t)slot[q] * 2, 0.f);                       // two time slots, the second is used
        table[((size_t)b * nf + f) * 3 + q] = stores.back().data();
    }
    for (int f = 0; f < nf; f++) {
        BndField& F = a.f[f]; F.f0 = W.data() + (size_t)f * vol;
        const char* t = nd == 3 ? types3[f] : types2[f];
        for (int q = 0; q < 3; q++) {
            if (nd == 2 && q == 1) { continue; }
            const char ty = nd == 3 ? t[q] : t[q == 0 ? 0 : 1];
            const int sh = len(ty, nn[q]), of = off(ty);
            F.lo[q] = npml + of; F.hi[q] = sh - npml - nbound + of;
            (q == 0 ? F.k0 : q == 1 ? F.j0 : F.i0) = of; (q == 0 ? F.nk : q == 1 ? F.nj : F.ni) = of + sh;
        }
        if (nd == 2) { F.j0 = 0; F.nj = 1; }
    }
    for (int q = 0; q < 3; q++) a.slot_off[q] = slot[q];                     // time slot 1
    a.stores = table.data(); a.wstride = (long long)nf * vol;
    const int umax = g.pz > g.nx1 ? g.pz : g.nx1, vmax = g.nx1 > g.ny1 ? g.nx1 : g.ny1;
    const emu_dim3 grd = d3((umax + 127) / 128, vmax, 2 * nbound * a.naxes * nf * nb), blk = d3(128, 1, 1);
    launch(k_boundary<1>, grd, blk, g, a);
    for (auto& x : W) x = 0.f;                                                // wipe, then force the stored planes back
    launch(k_boundary<0>, grd, blk, g, a);
    size_t restored = 0, wrong = 0;
    for (size_t q = 0; q < W.size(); q++) { if (W[q] != 0.f) { restored++; if (W[q] != -W0[q]) wrong++; } }
    printf("  boundary store %d-D, order %d, %d fields x %d shots: %zu plane values restored, %zu wrong\n", nd, 2 + 2 * h, nf, nb, restored, wrong);
    return restored == 0 || wrong != 0;
}

// medium coefficients, rigid faces of the order-4 path, FD-Born coefficient / add kernels, z-plane halo pack / unpack: bounds only
static int small_kernels_case(int nd, int h, unsigned seed) {
    const int faces = ZMIN | ZMAX | XMIN | XMAX | (nd == 3 ? YMIN | YMAX : 0);
    const Geom g = geom(nd, 22, 15, 18, h, 3, faces);
    const size_t vol = (size_t)g.vol;
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(0.5f, 1.5f);
    auto rnd = [&](size_t n) { std::vector<float> v(n); for (auto& x : v) x = U(rng); return v; };
    std::vector<float> m0 = rnd(vol), rho = rnd(vol), imu = rnd(vol);
    std::vector<std::vector<float>> dm(C_N, std::vector<float>(vol, 0.f));
    float* table[C_N]; for (int q = 0; q < C_N; q++) table[q] = dm[q].data();
    const emu_dim3 blk = nd == 3 ? d3(16, 2, 2) : d3(16, 2, 1);
    const emu_dim3 grd = nd == 3 ? d3((g.khi + 16) / 16, (g.ny1 + 1) / 2, (g.nx1 + 1) / 2) : d3((g.khi + 16) / 16, (g.nx1 + 1) / 2, 1);
    size_t written = 0;
    for (int el = 0; el < 2; el++) {
        for (auto& v : dm) std::fill(v.begin(), v.end(), 0.f);
        float* const* tp = table;
        if (h == 0) { if (nd == 2) { if (el) launch(k_dmod<2, 1>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)imu.data(), tp, 1e-3f); else launch(k_dmod<2, 0>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)nullptr, tp, 1e-3f); }
                      else         { if (el) launch(k_dmod<3, 1>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)imu.data(), tp, 1e-3f); else launch(k_dmod<3, 0>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)nullptr, tp, 1e-3f); } }
        else        { if (nd == 2) { if (el) launch(k_dmod4<2, 1>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)imu.data(), tp, 1e-3f); else launch(k_dmod4<2, 0>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)nullptr, tp, 1e-3f); }
                      else         { if (el) launch(k_dmod4<3, 1>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)imu.data(), tp, 1e-3f); else launch(k_dmod4<3, 0>, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)nullptr, tp, 1e-3f); } }
        for (auto& v : dm) for (float x : v) written += x != 0.f;
    }
    if (h == 1) {                                    // rigid faces, two ghost nodes per face
        std::vector<float> W = rnd(9 * vol);
        StepArgs a; memset(&a, 0, sizeof a);
        for (int q = 0; q < 3; q++) a.v[q] = W.data() + (size_t)(6 + q) * vol;
        a.wstride = 9 * vol; a.nbatch = 1;
        const int nn[3] = {g.nz, g.ny, g.nx};
        for (int axis = 2; axis >= 0; axis--) {
            if (axis == 1 && nd == 2) continue;
            const int o1 = axis == 0 ? (nd == 3 ? 1 : 2) : 0, o2 = axis == 0 ? (nd == 3 ? 2 : -1) : (axis == 1 ? 2 : (nd == 3 ? 1 : -1));
            const emu_dim3 fg = d3((nn[o1] + 127) / 128, o2 >= 0 ? nn[o2] : 1, 1);
            if (nd == 3) launch(k_dirichlet4<3>, fg, d3(128, 1, 1), g, a, axis); else launch(k_dirichlet4<2>, fg, d3(128, 1, 1), g, a, axis);
        }
    }
    if (nd == 2 && h == 0) {                         // FD-Born kernels
        std::vector<float> dK = rnd(vol), dr = rnd(vol), c0(vol), c1(vol), c2(vol), f0 = rnd(2 * vol), f1 = rnd(2 * vol), dd = rnd(4 * vol);
        launch(k_born_coef, grd, blk, g, (const float*)m0.data(), (const float*)rho.data(), (const float*)dK.data(), (const float*)dr.data(), c0.data(), c1.data(), c2.data(), 1e-3f);
        launch(k_born_add<0>, d3(grd.x, grd.y, 2), blk, g, f0.data(), f1.data(), (const float*)dd.data(), (const float*)(dd.data() + vol), (const float*)c1.data(), (const float*)c2.data(), (long long)vol, (long long)(2 * vol));
        launch(k_born_add<1>, d3(grd.x, grd.y, 2), blk, g, f0.data(), (float*)nullptr, (const float*)dd.data(), (const float*)(dd.data() + vol), (const float*)c0.data(), (const float*)nullptr, (long long)vol, (long long)(2 * vol));
    }
    if (nd == 3 && h == 0) {                         // z-plane halo pack / unpack of three fields
        std::vector<float> W = rnd(3 * vol), buf(3 * (size_t)g.ny1 * g.nx1, 0.f);
        HaloArgs ha; ha.n = 3; ha.i0 = 0; ha.ni = g.nx1; for (int q = 0; q < 3; q++) { ha.field[q] = W.data() + (size_t)q * vol; ha.k[q] = g.khi; }
        launch(k_halo<1>, d3((g.ny1 + 127) / 128, g.nx1, 1), d3(128, 1, 1), g, ha, buf.data());
        for (int q = 0; q < 3; q++) ha.k[q] = 0;
        launch(k_halo<0>, d3((g.ny1 + 127) / 128, g.nx1, 1), d3(128, 1, 1), g, ha, buf.data());
        for (int q = 0; q < 3; q++) for (int i = 0; i < g.nx1; i++) for (int j = 0; j < g.ny1; j++)
            if (W[q * vol + 0 + (size_t)g.pz * (j + (size_t)g.ny1 * i)] != W[q * vol + g.khi + (size_t)g.pz * (j + (size_t)g.ny1 * i)]) return 1;
    }
    printf("  medium coefficients / rigid faces / Born / halo kernels %d-D, order %d: %zu coefficients written\n", nd, 2 + 2 * h, written);
    return written == 0;
}

int main() {
    int bad = 0;
    for (int h = 0; h < 2; h++) { bad += small_kernels_case(2, h, 11 + h); bad += small_kernels_case(3, h, 13 + h); }
    for (int h = 0; h < 2; h++) { bad += grad_case(2, h, 1 + h); bad += grad_case(3, h, 3 + h); }
    for (int h = 0; h < 2; h++) { bad += boundary_case(2, h, 5 + h); bad += boundary_case(3, h, 7 + h); }
    printf(bad ? "EMU_MISMATCH\n" : "EMU_OK\n");
    return bad ? 1 : 0;
}
