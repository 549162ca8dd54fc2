// CPU emulation check: the four-cells-per-thread order-4 kernels (kernels4v.cuh) against the one-cell-per-thread kernels
// (kernels4.cuh, validated bit for bit against the oracle on the GPU) on random fields, with CPML slabs on every face.
// Threads run sequentially; neither kernel family has inter-thread communication.  Exit code 0 = bit-identical.
#include "cuda_shim.h"
#include "../../geophyinv.jl_b200/csrc/kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace gpi;

template <typename K> static void launch(K kernel, emu_dim3 grid, emu_dim3 block, const Geom& g, const StepArgs& a) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz; threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
            kernel(g, a);
        }
}

struct State {
    std::vector<float> W, MEM, C;          // wavefields [nbatch][9][vol], CPML memory, coefficients [C_N][vol]
};

template <int ND, int EL> static int run_case(int nz, int ny, int nx, int npml, int faces, int freesurf, int nbatch, unsigned seed) {
    Geom g; memset(&g, 0, sizeof g);
    g.nz = nz; g.ny = ND == 3 ? ny : 1; g.nx = nx; g.h = 1; g.koff = 0; g.klo = 0; g.khi = nz + 2;
    g.pz = ((g.khi + 2 + 31) / 32) * 32; g.ny1 = ND == 3 ? ny + 3 : 1; g.nx1 = nx + 3; g.npml = npml;
    g.pzm = ((2 * npml + 31) / 32) * 32; g.pml = faces; g.rigid = faces; g.freesurf = freesurf;
    g.dzI = 1.0f / 240.0f; g.dyI = 1.0f / 264.0f; g.dxI = 1.0f / 216.0f; g.vol = (long long)g.pz * g.ny1 * g.nx1;
    const long long vol = g.vol;
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(-1.f, 1.f);
    State s0;
    s0.W.resize((size_t)nbatch * 9 * vol); for (auto& x : s0.W) x = U(rng);
    s0.C.resize((size_t)C_N * vol); for (auto& x : s0.C) x = 0.5f + 0.25f * U(rng);
    // CPML terms: 9 velocity-kernel + 9 stress-kernel slots, each a generous block
    const long long msz = std::max<long long>((long long)g.pz * g.ny1 * 2 * npml, std::max<long long>((long long)g.pz * 2 * npml * g.nx1, (long long)g.pzm * g.ny1 * g.nx1));
    s0.MEM.resize((size_t)nbatch * 18 * msz); for (auto& x : s0.MEM) x = 0.1f * U(rng);
    std::vector<float> coef(18 * 3 * 2 * npml); for (auto& x : coef) x = 0.5f + 0.4f * U(rng);
    int bad = 0;
    for (int vel = 1; vel >= 0; vel--) {
        State sa = s0, sb = s0;
        auto args = [&](State& s) {
            StepArgs a; memset(&a, 0, sizeof a);
            for (int q = 0; q < 6; q++) a.tau[q] = s.W.data() + (size_t)q * vol;
            for (int q = 0; q < 3; q++) a.v[q] = s.W.data() + (size_t)(6 + q) * vol;
            for (int q = 0; q < C_N; q++) a.c[q] = s.C.data() + (size_t)q * vol;
            for (int q = 0; q < 18; q++) {
                PmlTerm t; t.mem = s.MEM.data() + (size_t)q * msz; t.a = coef.data() + (size_t)(q * 3 + 0) * 2 * npml;
                t.b = coef.data() + (size_t)(q * 3 + 1) * 2 * npml; t.kI = coef.data() + (size_t)(q * 3 + 2) * 2 * npml; t.bstride = 18 * msz;
                if (q < 9) a.pv[q] = t; else a.ps[q - 9] = t;
            }
            a.wstride = 9 * vol; a.nbatch = nbatch;
            return a;
        };
        StepArgs aa = args(sa), ab = args(sb);
        emu_dim3 blk, grd, blkv, grdv;
        const int ngx = ((g.pz / 4) + 31) / 32;
        if (ND == 3) {
            blk.x = 16; blk.y = 2; blk.z = 2; grd.x = (g.khi + 1 + 15) / 16; grd.y = (g.ny1 + 1) / 2; grd.z = ((g.nx1 + 1) / 2) * nbatch;
            blkv.x = 32; blkv.y = 2; blkv.z = 2; grdv.x = ngx; grdv.y = (g.ny1 + 1) / 2; grdv.z = ((g.nx1 + 1) / 2) * nbatch;
        } else {
            blk.x = 16; blk.y = 2; blk.z = 1; grd.x = (g.khi + 1 + 15) / 16; grd.y = (g.nx1 + 1) / 2; grd.z = nbatch;
            blkv.x = 32; blkv.y = 4; blkv.z = 1; grdv.x = ngx; grdv.y = (g.nx1 + 3) / 4; grdv.z = nbatch;
        }
        if (vel) { launch(k_vel4<ND, EL>, grd, blk, g, aa); launch(k_vel4v<ND, EL>, grdv, blkv, g, ab); }
        else     { launch(k_stress4<ND, EL>, grd, blk, g, aa); launch(k_stress4v<ND, EL>, grdv, blkv, g, ab); }
        size_t dw = 0, dm = 0, changed = 0;
        for (size_t q = 0; q < sa.W.size(); q++) { if (memcmp(&sa.W[q], &sb.W[q], 4)) dw++; if (memcmp(&sa.W[q], &s0.W[q], 4)) changed++; }
        for (size_t q = 0; q < sa.MEM.size(); q++) if (memcmp(&sa.MEM[q], &sb.MEM[q], 4)) dm++;
        printf("  ND=%d EL=%d %s: %zu cells updated, wavefield mismatches %zu, CPML-memory mismatches %zu\n", ND, EL, vel ? "velocity" : "stress", changed, dw, dm);
        if (dw || dm || !changed) bad++;
    }
    return bad;
}

int main() {
    const int all = ZMIN | ZMAX | YMIN | YMAX | XMIN | XMAX;
    int bad = 0;
    printf("3-D elastic, CPML on all faces, free surface flag\n");   bad += run_case<3, 1>(21, 18, 19, 3, all, ZMIN, 1, 1);
    printf("3-D elastic, partial faces, z extent 29\n");             bad += run_case<3, 1>(29, 17, 20, 4, ZMAX | XMIN | YMAX, 0, 1, 2);
    printf("3-D acoustic\n");                                        bad += run_case<3, 0>(26, 18, 17, 3, all, 0, 1, 3);
    printf("2-D elastic, 3 resident shots, free surface\n");         bad += run_case<2, 1>(37, 1, 23, 3, ZMAX | XMIN | XMAX, ZMIN, 3, 4);
    printf("2-D acoustic, 2 resident shots\n");                      bad += run_case<2, 0>(30, 1, 41, 5, ZMIN | ZMAX | XMIN | XMAX, 0, 2, 5);
    printf("2-D acoustic, no CPML, z extent 124 (pz = nz + 4)\n");   bad += run_case<2, 0>(124, 1, 20, 3, 0, 0, 1, 6);
    printf(bad ? "EMU_MISMATCH\n" : "EMU_OK\n");
    return bad ? 1 : 0;
}
