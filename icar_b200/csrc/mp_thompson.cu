// Thompson et al. (2008) bulk microphysics column update on sm_100a.
//
// Reference: src/physics/mp_thompson.f90 -- mp_gt_driver :772-1044 (3-D <-> column wrapper, precipitation
// sums, water-vapour floor) and mp_thompson :1057-2844 (the column physics).
//
// Parallelisation.  The reference walks one column at a time, but almost every statement of the column
// routine is independent per level: the load/clamp block, thermodynamics, snow moments, all process
// rates, the conservation rescaling, the tendencies and the state update never look at another level.
// What couples levels is small:
//   * no_micro            -- an OR over the column                              (:1245-1345, exit :1363)
//   * graupel intercept   -- N0_min is a running minimum from the model top     (:1456-1468, :2379-2391)
//   * fall speeds         -- levels without a species inherit the speed above   (:2544-2545 ...)
//   * nstep / ksed1       -- column maxima per species                          (:2548-2555 ...)
//   * sedimentation       -- flux divergence needs sed(k+1), per sub-step       (:2660-2770)
// So, as for mp_simple, a CTA owns TX (16 or 8) columns x all levels with ONE THREAD PER CELL (threadIdx.x = column
// within the tile, threadIdx.y = level): the cell's state and its ~55 REAL*8 process rates live in
// registers, the couplings above go through a few shared-memory rows, and each field is read once and
// written once.  Levels are processed in the Fortran's statement order within a cell, so the result
// differs from the CPU restatement only through the device's powf/expf/log10f/pow/exp (and the three cheaper forms of
// x**y described in the fragment).
//
// Typing follows the Fortran: REAL -> float, DOUBLE PRECISION -> double, mixed expressions promote
// where the Fortran's do.  -fmad=false (see Makefile): no contraction.
#include <algorithm>
#include <cstdlib>

#include <cstring>

#include "mp_thompson.cuh"

namespace icar {

namespace {

#define ICAR_DYNAMIC_SMEM_D(name) extern __shared__ double name[]
#include "mp_thompson_kernel.inc"

}  // namespace

int th_driver_launch(const Geom& g, const ThConst& c, const ThTables& tb, float* qv, float* qc, float* qr, float* qi,
                     float* qs, float* qg, float* ni, float* nr, float* th, const float* pii, const float* p,
                     const float* dz, float dt, float* rainnc, const float* rainncv, float* snownc, float* snowncv,
                     float* graupelnc, float* graupelncv, float* sr, int ld2, const MpTiles& tiles, int kts, int kte,
                     unsigned long long* diag, cudaStream_t st, double* acc_precip, double* acc_snow,
                     double* acc_graupel) {
    const int nk = kte - kts + 1;
    const int TX = th_columns_per_cta(nk);
    int gx = 0, gy = 0;
    for (int t = 0; t < tiles.n; ++t) {
        if (tiles.ite[t] < tiles.its[t] || tiles.jte[t] < tiles.jts[t]) continue;
        const int w = tiles.ite[t] - tiles.its[t] + 1, h = tiles.jte[t] - tiles.jts[t] + 1;
        gx = std::max(gx, ((tiles.along_y[t] ? h : w) + TX - 1) / TX);
        gy = std::max(gy, tiles.along_y[t] ? w : h);
    }
    if (gx == 0 || gy == 0) return 0;
    if (nk < 1 || nk * TX > 1024) return -1;                // one thread per level
    const int nkt = (nk * TX + 31) / 32 * 32 / TX;          // whole warps
    Args A{qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz, rainnc, snownc, snowncv, graupelnc, graupelncv, sr, rainncv,
           acc_precip, acc_snow, acc_graupel};
    dim3 block(TX, nkt), grid(gx, gy, tiles.n);
    const size_t smem = (size_t)nk * TX * sizeof(double) + (size_t)7 * (nk + 1) * TX * sizeof(float) + (size_t)10 * TX * sizeof(int);
    const int nthreads = TX * nkt;
#define TH_LAUNCH(TXC, MAXT, MINB) k_mp_thompson<TXC, MAXT, MINB><<<grid, block, smem, st>>>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag)
    // registers: what the register file gives the resident CTAs -- 96 for two CTAs of <= 320 threads or one of <= 640,
    // 80 up to 800 threads, 64 above.  Occupancy matters more than the spills: at 40 levels one CTA per SM with 168
    // registers and no spills ran at 10.7 ms per step against 7.8; at 50 levels 25 warps with 64 registers beat 13 with 96.
    if (TX == 16) {
        if (nthreads <= 320) { TH_LAUNCH(16, 320, 2); }
        else if (nthreads <= 640) { TH_LAUNCH(16, 640, 1); }
        else if (nthreads <= 800) { TH_LAUNCH(16, 800, 1); }
        else { TH_LAUNCH(16, 1024, 1); }
    }
    else if (nthreads <= 640) { TH_LAUNCH(8, 640, 1); }
    else { TH_LAUNCH(8, 1024, 1); }
#undef TH_LAUNCH
    return 1;
}

bool th_driver_preload() {
    return preload_kernel(k_mp_thompson<16, 320, 2>) && preload_kernel(k_mp_thompson<16, 640, 1>) &&
           preload_kernel(k_mp_thompson<16, 800, 1>) && preload_kernel(k_mp_thompson<16, 1024, 1>) &&
           preload_kernel(k_mp_thompson<8, 640, 1>) && preload_kernel(k_mp_thompson<8, 1024, 1>);
}

}  // namespace icar
