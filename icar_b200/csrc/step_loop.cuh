// The cheap parts of step()'s loop body around mp / advect (src/main/time_step.f90:464-547), so that a
// device-resident step loop needs no host round trip: diagnostic_update's exner / density (:85-101),
// compute_dt's CFL maxima (:217-330) and domain%enforce_limits (domain_obj.f90:2228-2243).
#pragma once
#include "common.cuh"

namespace icar {

// exner = (pressure / 100000) ** (Rd/cp); density = pressure / (Rd * potential_temperature * exner)
// (time_step.f90:85,96,101; atm_utilities.f90:682-691; Rd = 287.058, cp = 1012, icar_constants.f90:391-393)
int diagnostic_update_launch(const Geom& g, const float* pressure, const float* th, float* exner, float* density,
                             cudaStream_t st);

// The rest of diagnostic_update, the fields only output / other physics read (time_step.f90:88-108, 165-194):
//   temperature = th * exner; pressure_i / temperature_i(k > kms) = (x(k-1) + x(k)) / 2, (kms) = x(kms) + (x(kms) - x(kms+1)) / 2
//   u_mass = (u(i+1) + u(i)) / 2; v_mass = (v(j+1) + v(j)) / 2
//   w_real(ims+1:ime-1, k, jms+1:jme-1) = (uw(i) + uw(i+1)) * 0.5 + (vw(j) + vw(j+1)) * 0.5 + jacobian * (w(k-1) + w(k)) * 0.5
//       with uw = u * dzdx, vw = v * dzdy, w(kms-1) = 0; the boundary ring of w_real is never written (stays 0)
// Any output pointer may be null.  dzdx is u-shaped, dzdy v-shaped (domain_obj.f90:893-900).
int diagnostic_thermo_launch(const Geom& g, const float* pressure, const float* th, const float* exner,
                             float* temperature, float* p_i, float* t_i, cudaStream_t st);
int diagnostic_winds_launch(const Geom& g, const float* u, const float* v, const float* w, const float* dzdx,
                            const float* dzdy, const float* jaco, float* u_mass, float* v_mass, float* w_real,
                            cudaStream_t st);

// compute_dt's four maxima over this image's tile (time_step.f90:234-289, use_density = .false.):
//   out[0] = max over cells of  max|u|/dx + max|v|/dx + max|w|/dz(k)   (the strictness 2/3/4 loop)
//   out[1..3] = maxval(abs(u)), maxval(abs(v)), maxval(abs(w))
// stored as the bit patterns of non-negative floats (integer order == float order); out must be zeroed.
int cfl_maxima_launch(const Geom& g, const float* u, const float* v, const float* w, const float* dz_levels,
                      float dx, unsigned* out4, cudaStream_t st);

// domain%apply_forcing (domain_obj.f90:2383-2448) for one 3-D REAL variable: q = q + dqdt * dt%seconds()
// (the product and the sum are REAL*8, rounded once on assignment).  boundaries_only: force_boundaries =
// .True. variables (advected scalars) are forced on the domain edges this image owns -- west/east columns
// for rows jms+1..jme-1 (:2409-2414), whole south/north rows (:2415-2420); edges[4] = {west, east, south, north}.
// width = nx (+1 for u), rows_y = ny (+1 for v), pitch = row pitch of the variable.
int apply_forcing_launch(float* q, const float* dqdt, int width, int nz, int rows_y, int pitch, double seconds,
                         bool boundaries_only, const bool edges[4], cudaStream_t st);

// balance_uvw (src/physics/wind.f90:81-172) with calc_divergence(horz_only) (:175-228): the w that balances
// the horizontal divergence of u*jacobian_u, v*jacobian_v, integrated upward from w = 0 at the ground.
// One thread per column (the level loop carries w(k-1)); lanes run along x.
int balance_uvw_launch(const Geom& g, const float* u, const float* v, float* w, const float* ju, const float* jv,
                       const float* jw, const float* dz, float dx, cudaStream_t st);

// output_t%save_data writes a 3-D variable as reshape(data_3d, order=[1,3,2]) (src/io/output_obj.f90:423-426): the file
// holds (x, y, z), memory holds (x, z, y).  to_file: dense[(k*rows_y + j)*width + i] = field(i, k, j); otherwise the
// reverse (read_restart_data, src/io/restart.f90:57-63).  x is the fastest index on both sides.
int output_reorder_launch(float* field, float* dense, int width, int nz, int rows_y, int pitch, bool to_file,
                          cudaStream_t st);

// where (q < 0) q = 0 for every listed field (domain_obj.f90:2228-2243)
int enforce_limits_launch(const Geom& g, float* const* fields, int nfields, cudaStream_t st);

}  // namespace icar
