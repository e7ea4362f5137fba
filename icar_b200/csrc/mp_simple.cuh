// Launchers of the mp_simple kernels (mp_simple.cu).  Return = kernels queued on `st`.
#pragma once
#include "common.cuh"

namespace icar {

// up to 4 disjoint 0-based inclusive tiles processed by one launch
struct MpTiles {
    int n;
    int its[4], ite[4], jts[4], jte[4];
    int along_y[4];     // 1: the 32 lanes of a warp run along j (for strips only a few columns wide)
};

// mp_simple_driver (mp_simple.f90:595-646) on the 0-based inclusive tile its..ite, jts..jte,
// kts..kte.  rain/snow are (nx,ny) REAL*4 with row pitch ld2.  c2r/c2s = cloud2rain/cloud2snow
// (:619-620), diag = 3 device counters (see icar_mp_diagnostics).
int mp_simple_launch(const Geom& g, const float* p, float* th, const float* pii, const float* rho,
                     float* qv, float* qc, float* qr, float* qs, float* rain, float* snow, int ld2,
                     float dt, const float* dz, const MpTiles& tiles, int kts, int kte, float c2r,
                     float c2s, unsigned long long* diag, cudaStream_t st, double* acc_rain = nullptr,
                     double* acc_snow = nullptr);
// (acc_rain / acc_snow non-null: resident mode -- rain / snow are ignored and the call's precipitation is
//  added straight to domain%accumulated_precipitation / accumulated_snowfall, REAL*8, mp_driver.f90:587-595)

bool mp_simple_preload();     // see preload_kernel (common.cuh)

}  // namespace icar
