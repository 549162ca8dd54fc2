// CPU emulation check, order 2: the float4-per-thread kernels (k_vel2v / k_stress2v, k_vel3v / k_stress3v) against the
// one-thread-per-cell reference-order kernels (k_vel / k_stress) on random fields -- CPML slabs with the engine's z-memory
// layout and k-indexed z coefficient tables, rigid faces, free surface, several resident shots, ragged z extents.
// Threads run sequentially; the kernels have no inter-thread communication.  Exit code 0 = bit-identical.
#include "cuda_shim.h"
#include "../../geophyinv.jl_b200/csrc/kernels.cuh"
#include <algorithm>
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
struct State { std::vector<float> W, MEM; };

// (s0, len) along z of the derivative field behind each CPML slot (kernels.cuh: vel_cell / stress_cell)
static void zextent(int nd, int el, bool vel, int idx, int nz, int& s0, int& len) {
    s0 = 0; len = 0;
    if (vel) {
        if (!el) { if (idx == 2) { s0 = 1; len = nz - 1; } }                        // dpdz: half nodes
        else if (idx == 2 || idx == 5) { s0 = 1; len = nz - 2; }                   // dtauxzdz, dtauyzdz: inner nodes
        else if (idx == 8) { s0 = 1; len = nz - 1; }                               // dtauzzdz: half nodes
    } else {
        if (idx == 2) { s0 = 0; len = nz; }                                        // dvzdz: integer nodes
        else if (idx == 5 || idx == 7) { s0 = 1; len = nz - 1; }                   // dvxdz, dvydz: half nodes
    }
    (void)nd;
}

template <int ND, int EL> static int run_case(int nz, int ny, int nx, int npml, int faces, int rigid, int freesurf, int nbatch, unsigned seed) {
    Geom g; memset(&g, 0, sizeof g);
    g.nz = nz; g.ny = ND == 3 ? ny : 1; g.nx = nx; g.h = 0; g.koff = 0; g.klo = 0; g.khi = nz;
    g.pz = ((g.khi + 2 + 31) / 32) * 32; g.ny1 = ND == 3 ? ny + 1 : 1; g.nx1 = nx + 1; g.npml = npml;
    g.pzm = ((2 * ((npml + 3 + 3) / 4 * 4) + 31) / 32) * 32; g.pml = faces; g.rigid = rigid; g.freesurf = freesurf;
    g.dzI = 1.0f / 10.0f; g.dyI = 1.0f / 11.0f; g.dxI = 1.0f / 9.0f; g.vol = (long long)g.pz * g.ny1 * g.nx1;
    const long long vol = g.vol;
    const int pzt = ((nz + 64 + 31) / 32) * 32;
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(-1.f, 1.f);
    State s0;
    s0.W.resize((size_t)nbatch * 9 * vol); for (auto& x : s0.W) x = U(rng);
    std::vector<float> C((size_t)C_N * vol); for (auto& x : C) x = 0.5f + 0.25f * U(rng);
    const long long msz = std::max<long long>((long long)g.pz * g.ny1 * 2 * npml, std::max<long long>((long long)g.pz * 2 * npml * g.nx1, (long long)g.pzm * g.ny1 * g.nx1));
    s0.MEM.resize((size_t)nbatch * 18 * msz); for (auto& x : s0.MEM) x = 0.1f * U(rng);
    // x / y terms: slab-indexed coefficient vectors; z terms: k-indexed tables, identity outside the slabs (gpi_set_pml)
    std::vector<float> coef((size_t)18 * 3 * 2 * npml); for (auto& x : coef) x = 0.5f + 0.4f * U(rng);
    std::vector<float> ztab((size_t)18 * 3 * pzt);
    for (int q = 0; q < 18; q++) {
        int z0, zl; zextent(ND, EL, q < 9, q % 9, nz, z0, zl);
        for (int k = 0; k < pzt; k++) {
            const int r = k - z0; int s = -1;
            if ((faces & ZMIN) && r >= 0 && r < npml) s = r;
            else if ((faces & ZMAX) && zl > 0 && r - (zl - npml) >= 0 && r < zl) s = npml + r - (zl - npml);
            for (int cc = 0; cc < 3; cc++) ztab[((size_t)q * 3 + cc) * pzt + k] = s >= 0 ? coef[((size_t)q * 3 + cc) * 2 * npml + s] : (cc == 2 ? 1.f : 0.f);
        }
    }
    auto axis_of = [&](bool vel, int idx) {      // axis of each CPML slot (vel_cell / stress_cell term order)
        if (vel) { if (!EL) return idx == 0 ? 2 : idx == 1 ? 1 : 0; const int ax[9] = {2, 1, 0, 2, 1, 0, 2, 1, 0}; return ax[idx]; }
        const int ax[9] = {2, 1, 0, 1, 2, 0, 2, 0, 1}; return ax[idx];
    };
    int bad = 0;
    {
        // three full time steps (velocity kernel, stress kernel) with each family from the same state.  The float4 kernels also
        // update the memory variables of the cells of a group that lie outside the derivative field (their range predicates act at
        // the store of the fields); those entries belong to no cell of the field, so they are compared only through their effect
        // on the wavefields, the entries the reference-order kernels use are compared directly.
        State sa = s0, sb = s0;
        auto args = [&](State& s) {
            StepArgs a; memset(&a, 0, sizeof a);
            for (int q = 0; q < 6; q++) a.tau[q] = s.W.data() + (size_t)q * vol;
            for (int q = 0; q < 3; q++) a.v[q] = s.W.data() + (size_t)(6 + q) * vol;
            for (int q = 0; q < C_N; q++) a.c[q] = C.data() + (size_t)q * vol;
            for (int q = 0; q < 18; q++) {
                PmlTerm t; t.mem = s.MEM.data() + (size_t)q * msz; t.bstride = 18 * msz;
                if (axis_of(q < 9, q % 9) == 0) { t.a = ztab.data() + ((size_t)q * 3 + 0) * pzt; t.b = ztab.data() + ((size_t)q * 3 + 1) * pzt; t.kI = ztab.data() + ((size_t)q * 3 + 2) * pzt; }
                else { t.a = coef.data() + ((size_t)q * 3 + 0) * 2 * npml; t.b = coef.data() + ((size_t)q * 3 + 1) * 2 * npml; t.kI = coef.data() + ((size_t)q * 3 + 2) * 2 * npml; }
                if (q < 9) a.pv[q] = t; else a.ps[q - 9] = t;
            }
            a.wstride = 9 * vol; a.nbatch = nbatch;
            return a;
        };
        StepArgs aa = args(sa), ab = args(sb);
        emu_dim3 blk, grd, blkv, grdv;
        for (int step = 0; step < 3; step++) for (int vel = 1; vel >= 0; vel--) {
            if (ND == 3) {
                blk.x = 16; blk.y = 2; blk.z = 2; grd.x = (g.khi + 1 + 15) / 16; grd.y = (g.ny1 + 1) / 2; grd.z = ((g.nx1 + 1) / 2) * nbatch;
                const int ng = vec3_threads(g.pz, g.ny1);
                blkv.x = 32; blkv.y = 1; blkv.z = 1; grdv.x = (ng + 31) / 32; grdv.y = g.nx1; grdv.z = nbatch;
                if (vel) { launch(k_vel<ND, EL>, grd, blk, g, aa); launch(k_vel3v<EL>, grdv, blkv, g, ab); }
                else     { launch(k_stress<ND, EL>, grd, blk, g, aa); launch(k_stress3v<EL>, grdv, blkv, g, ab); }
            } else {
                blk.x = 16; blk.y = 2; blk.z = 1; grd.x = (g.khi + 1 + 15) / 16; grd.y = (g.nx1 + 1) / 2; grd.z = nbatch;
                const int nth = (g.pz / VW) * g.nx1;
                blkv.x = 128; blkv.y = 1; blkv.z = 1; grdv.x = (nth + 127) / 128; grdv.y = nbatch; grdv.z = 1;
                if (vel) { launch(k_vel<ND, EL>, grd, blk, g, aa); launch(k_vel2v<EL>, grdv, blkv, g, ab); }
                else     { launch(k_stress<ND, EL>, grd, blk, g, aa); launch(k_stress2v<EL>, grdv, blkv, g, ab); }
            }
        }
        size_t dw = 0, dm = 0, changed = 0, used = 0;
        for (size_t q = 0; q < sa.W.size(); q++) { if (memcmp(&sa.W[q], &sb.W[q], 4)) dw++; if (memcmp(&sa.W[q], &s0.W[q], 4)) changed++; }
        for (size_t q = 0; q < sa.MEM.size(); q++) if (memcmp(&sa.MEM[q], &s0.MEM[q], 4)) { used++; if (memcmp(&sa.MEM[q], &sb.MEM[q], 4)) dm++; }
        printf("  ND=%d EL=%d, 3 steps: %zu field values updated, mismatches %zu; %zu CPML memory variables in use, mismatches %zu\n", ND, EL, changed, dw, used, dm);
        if (dw || dm || !changed) bad++;
    }
    return bad;
}

int main() {
    const int all = ZMIN | ZMAX | YMIN | YMAX | XMIN | XMAX, all2 = ZMIN | ZMAX | XMIN | XMAX;
    int bad = 0;
    printf("3-D elastic, CPML + rigid on all faces\n");                 bad += run_case<3, 1>(21, 18, 19, 3, all, all, 0, 1, 1);
    printf("3-D elastic, free surface, partial faces, z extent 30\n");  bad += run_case<3, 1>(30, 17, 20, 4, ZMAX | XMIN | XMAX | YMAX, ZMAX | XMIN | XMAX | YMAX, ZMIN, 1, 2);
    printf("3-D acoustic\n");                                           bad += run_case<3, 0>(26, 18, 17, 3, all, all, 0, 1, 3);
    printf("2-D elastic, 3 resident shots, free surface\n");            bad += run_case<2, 1>(37, 1, 23, 3, ZMAX | XMIN | XMAX, ZMAX | XMIN | XMAX, ZMIN, 3, 4);
    printf("2-D acoustic, 2 resident shots\n");                         bad += run_case<2, 0>(30, 1, 41, 5, all2, all2, 0, 2, 5);
    printf("2-D acoustic, no CPML, rigid x faces, z extent 126\n");     bad += run_case<2, 0>(126, 1, 20, 3, 0, XMIN | XMAX, 0, 1, 6);
    printf(bad ? "EMU_MISMATCH\n" : "EMU_OK\n");
    return bad ? 1 : 0;
}
