// Launchers of the MPDATA kernels (definitions in mpdata_staged.cu / mpdata_fused.cu).
// Every launcher returns the number of kernels it queued on `st`.
#pragma once
#include "common.cuh"

#define ICAR_MAX_SCALARS 12

namespace icar {

// the advected scalars of one advect() call: q = field before pass 1, q2 = scratch for the field
// after pass 1, out = buffer that receives the advected field (all distinct)
struct ScalarSet {
    int n;
    const float* q[ICAR_MAX_SCALARS];
    float* q2[ICAR_MAX_SCALARS];
    float* out[ICAR_MAX_SCALARS];
};

// per-face scalar-independent coefficients of the fused corrector (corrector_fused.inc: k_face_coefficients)
struct FusedCoef {
    float *t0x, *cbx, *ccx, *t0y, *cby, *ccy, *ggy, *t0z, *cbz, *ccz;
};

// launch plan of the fused corrector for a tile; ok == false when it cannot serve this geometry
// (fewer than 3 or more than 64 levels, or not enough opt-in shared memory): the caller then uses the staged path
struct FusedPlan {
    bool ok;
    int cpt, nzp, stages;
    size_t smem;
};
FusedPlan fused_plan(const Geom& g);

// order 2 + FCT, rho == 1, fp32 (mpdata_fused.cu): face coefficients, pass 1 for all scalars, fused corrector -- the
// reference's arithmetic operation for operation.  U, V, W, Wn from staged_wind_normalise.  Returns < 0 when the TMA
// tensor maps cannot be built (driver without cuTensorMapEncodeTiled).
// marks (optional): 4 events recorded on `st` before the coefficient kernel and after each of the three
// kernels, for the per-kernel profile (icar_profile_read).
// part: 0 = everything; 1 = k_face_coefficients only (depends on the winds, not on the scalars); 2 = pass 1 +
// corrector only.  A multi-image step queues part 1 before it waits for the halo exchange.
int fused_advect_all(const Geom& g, const ScalarSet& s, const float* U, const float* V, const float* W,
                     const float* Wn, const float* jaco, const float* dz, const FusedCoef& co,
                     cudaStream_t st, cudaEvent_t* marks = nullptr, int part = 0, const int* own = nullptr);
// own (optional): {i0, i1, j0, j1}, the half-open tile-local ranges of the cells whose result is stored (a slab stores
// only the cells it owns); default: the tile without its boundary ring.

// div2x of corrector_prims.inc (what the corrector divides with) and the compiler's a / b on n device floats (n even)
int selftest_division(const float* a, const float* b, float* packed, float* ieee, long long n, cudaStream_t st);

// pass 1 (adv_mpdata.f90:44-105) of every scalar in one launch; also copies the boundary ring to out
int staged_upwind_all(const Geom& g, const ScalarSet& s, const float* U, const float* V, const float* W,
                      const float* dz, const float* jaco, cudaStream_t st);

template <class R>
int staged_wind_normalise(const Geom& g, const R* u, const R* v, const R* w, const R* ju,
                          const R* jv, const R* jw, const R* rho, R dt, R dx, int scheme, R* U,
                          R* V, R* W, cudaStream_t st, const R* dz = nullptr, R* Wn = nullptr);

template <class R>
int staged_wn_G(const Geom& g, const R* W, const R* dz, const R* jaco, const R* rho, R* Wn, R* G,
                cudaStream_t st);

template <class R>
int staged_advect3d(const Geom& g, R* q, const R* rho, const R* jaco, const R* dz, const R* U,
                    const R* V, const R* W, const R* Wn, const R* G, int order, bool fct, R* qa,
                    R* qb, R* u2, R* v2, R* w2, R** result, cudaStream_t st);

bool fused_preload();         // see preload_kernel (common.cuh)
bool staged_preload();

}  // namespace icar
