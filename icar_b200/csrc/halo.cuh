// Launchers of the halo pack/unpack kernels (halo.cu).  Return = kernels queued on `st`.
#pragma once
#include "common.cuh"

#define ICAR_MAX_EXCHANGED 12   /* domain%halo_send lists 12 exchangeable scalars (domain_obj.f90:104-122) */

namespace icar {

// Gathers columns [x0[i], x0[i]+h) of `nfields` cell fields into buf[i][s][j][k][h] / scatters them
// back, for one or two strips (west and east) in a single launch.
int halo_pack_launch(const Geom& g, float* const* fields, int nfields, int h, int nsides, const int* x0,
                     float* const* buf, cudaStream_t st);
int halo_unpack_launch(const Geom& g, float* const* fields, int nfields, int h, int nsides, const int* x0,
                       float* const* buf, cudaStream_t st);

}  // namespace icar
