// Launchers of the MPDATA kernels (definitions in mpdata_staged.cu / mpdata_fused.cu).
// Every launcher returns the number of kernels it queued on `st`.
#pragma once
#include "common.cuh"

namespace icar {

template <class R>
int staged_wind_normalise(const Geom& g, const R* u, const R* v, const R* w, const R* ju,
                          const R* jv, const R* jw, const R* rho, R dt, R dx, int scheme, R* U,
                          R* V, R* W, cudaStream_t st);

template <class R>
int staged_wn_G(const Geom& g, const R* W, const R* dz, const R* jaco, const R* rho, R* Wn, R* G,
                cudaStream_t st);

template <class R>
int staged_advect3d(const Geom& g, R* q, const R* rho, const R* jaco, const R* dz, const R* U,
                    const R* V, const R* W, const R* Wn, const R* G, int order, bool fct, R* qa,
                    R* qb, R* u2, R* v2, R* w2, R** result, cudaStream_t st);

}  // namespace icar
