// "Simple" (SB04) cloud microphysics on sm_100a.
//
// Reference: src/physics/mp_simple.f90 -- mp_simple_driver :595-646, mp_simple :481-566,
// mp_conversions :381-420, cloud_conversion :198-280, cloud2hydrometeor :295-315,
// phase_change :333-362, sediment :437-459, sat_mr :146-182.
//
// Parallelisation.  The reference walks one column at a time; but inside a column every
// level is independent in mp_conversions and in the re-evaporation loops, and one
// sedimentation sub-step is level-parallel too once written per level:
//     q_new(k) = (q_old(k) - out(k)) + in(k),   in(k) = fall*q_old(k+1)*rho(k+1)/(rho(k)*dz(k))
// (same operation order as the Fortran's two sequential loops, :449-457).  So a CTA owns a
// tile of 32 columns x all levels with ONE THREAD PER CELL (CPT cells per thread when
// nz > 32): a warp is 32 consecutive i at one level, so every global access is one
// coalesced 128-byte line, the whole column state lives in registers from the first load
// to the last store (each field is read once and written once: the 14 field-passes of
// SURVEY.md section 8d), and only q_old(k+1) and three column reductions go through shared
// memory.
//
// Numerics: `exp` of the Fortran is glibc's expf on the reference platform; exp_glibc()
// below evaluates the same algorithm (same table, same double-precision polynomial), so
// together with IEEE division and -fmad=false the kernel reproduces the CPU oracle bit for
// bit rather than to a tolerance.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "mp_simple.cuh"

namespace icar {

namespace {

#define ICAR_DYNAMIC_SMEM_F(name) extern __shared__ float name[]
#include "mp_simple_kernel.inc"

}  // namespace

int mp_simple_launch(const Geom& g, const float* p, float* th, const float* pii, const float* rho,
                     float* qv, float* qc, float* qr, float* qs, float* rain, float* snow, int ld2,
                     float dt, const float* dz, const MpTiles& tiles, int kts, int kte, float c2r,
                     float c2s, unsigned long long* diag, cudaStream_t st, double* acc_rain, double* acc_snow) {
    int gx = 0, gy = 0;
    for (int t = 0; t < tiles.n; ++t) {
        if (tiles.ite[t] < tiles.its[t] || tiles.jte[t] < tiles.jts[t]) continue;
        const int w = tiles.ite[t] - tiles.its[t] + 1, h = tiles.jte[t] - tiles.jts[t] + 1;
        gx = std::max(gx, ((tiles.along_y[t] ? h : w) + TX - 1) / TX);
        gy = std::max(gy, tiles.along_y[t] ? w : h);
    }
    if (gx == 0 || gy == 0) return 0;
    int cpt = (g.nz + 9) / 10;                              // cells per thread: ~10 z-threads per CTA
    if (cpt > 8) return -1;                                 // nz > 80 is not built
    const int nzt = (g.nz + cpt - 1) / cpt;
    dim3 block(TX, nzt);
    dim3 grid(gx, gy, tiles.n);
    size_t smem = (size_t)(2 * (g.nz + 1) + 3) * TX * sizeof(float);
#define LAUNCH(C)                                                                                 \
    k_mp_simple<C><<<grid, block, smem, st>>>(g, p, th, pii, rho, qv, qc, qr, qs, rain, snow, ld2, \
                                              dt, dz, tiles, kts, kte, c2r, c2s, diag, acc_rain, acc_snow)
    switch (cpt) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break;
        case 5: LAUNCH(5); break;
        case 6: LAUNCH(6); break;
        case 7: LAUNCH(7); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    return 1;
}

bool mp_simple_preload() {
    return preload_kernel(k_mp_simple<1>) && preload_kernel(k_mp_simple<2>) && preload_kernel(k_mp_simple<3>) &&
           preload_kernel(k_mp_simple<4>) && preload_kernel(k_mp_simple<5>) && preload_kernel(k_mp_simple<6>) &&
           preload_kernel(k_mp_simple<7>) && preload_kernel(k_mp_simple<8>);
}

}  // namespace icar
