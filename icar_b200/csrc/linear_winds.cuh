// Linear-theory wind perturbation on the device: the direct evaluation of
// src/physics/linear_winds.f90 for the background wind at hand instead of the reference's look-up table
// over speed / direction / stability (7 TB at 2048 x 2048 x 60).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "common.cuh"

namespace icar {

struct LinWinds;   // terrain spectrum, wavenumber grids, cuFFT plan, work arrays of linear_theory_type

// add_buffer_topo (linear_winds.f90:351-418) on host arrays: terrain (nx, ny) REAL -> out (nx + 2 buffer,
// ny + 2 buffer) complex double, interleaved (re, im).  Host only: it runs once, at set-up, and its smoothing
// pass is sequential by construction (it reads cells it has just replaced).
void lw_add_buffer_topo(const float* terrain, int nx, int ny, int smooth_window, int buffer, double* out);

// setup_linwinds (:1180-1222) + initialize_linear_theory_data (:426-470): buffers the terrain twice (window 5,
// then a flat rim of 2), forward FFT / N, fftshift; wavenumber grids k, l.  Returns 0 or an error text.
int lw_create(LinWinds** out, const float* terrain, int nx, int ny, int buffer, float dx, cudaStream_t st,
              std::string* err);
void lw_destroy(LinWinds* lw);
// size of the buffered (FFT) grid and the total buffer (original + 2, :1207)
void lw_dims(const LinWinds* lw, int* fnx, int* fny, int* buffer);
// the terrain lw_create was given (domain%global_terrain) and its size
const float* lw_terrain(const LinWinds* lw, int* nx, int* ny);
// domain%terrain_frequency, (fnx, fny) complex double, to host
int lw_terrain_frequency(LinWinds* lw, double* host, std::string* err);

// linear_perturbation_constz (:239-277): leaves lt_data%u_perturb / v_perturb on the device.
// The mean over the sub-layer heights is taken in wavenumber space (the inverse transform is linear), so a
// layer costs two inverse FFTs whatever its number of sub-layers.
int lw_perturbation_constz(LinWinds* lw, float U, float V, float Nsq, float z_bottom, float z_top,
                           float minimum_step, long long* launches, std::string* err);
// linear_perturbation_varyingz (:280-344): z_bottom, z_top are HOST (nx, ny) fields of the un-buffered domain.
int lw_perturbation_varyingz(LinWinds* lw, float U, float V, float Nsq, const float* z_bottom, const float* z_top,
                             float minimum_step, long long* launches, std::string* err);
// u_perturb / v_perturb of the last call, (fnx, fny) complex double each, to host
int lw_download_perturbation(LinWinds* lw, double* u, double* v, std::string* err);

// The stagger of initialize_spatial_winds (:756-762) and the update of spatial_winds (:1094-1111) for level k of
// a tile whose first cell is global cell (i0, j0):
//   new_u(i, j) = real(u_perturb(buffer + i - 1, buffer + j) + u_perturb(buffer + i, buffer + j)) / 2
//   u_perturbation = u_perturbation * (1 - update_fraction) + update_fraction * new_u
//   u = u + u_perturbation * linear_contribution                         (v likewise, staggered in y)
int lw_apply_level(LinWinds* lw, const Geom& g, int i0, int j0, int k, float* u3d, float* v3d,
                   float linear_contribution, float update_fraction, long long* launches, std::string* err);


// lt_parameters as spatial_winds reads them (include/icar_b200.h: icar_lt_options has the same members)
struct LtOptions {
    int variable_N, smooth_nsq, vert_smooth, stability_window_size;
    float max_stability, min_stability, N_squared, linear_contribution, linear_update_fraction;
    float dirmax, dirmin, spdmax, spdmin;
    int n_dir_values, n_nsq_values, n_spd_values;
    float minimum_layer_size;
};
struct LtFields {          // resident fields of the tile (device, cell layout); hydrometeors may be null
    const float *th, *pii, *qv, *qc, *qi, *qr, *qs;
};

// domain%z%data_3d of the tile (host, (nx, nz, ny) dense): read by calc_stability; kept on the device
int lw_set_z(LinWinds* lw, const Geom& g, const float* z_host, std::string* err);

// spatial_winds (:840-1127), reverse = .False.: stability field and its smoothing, column-mean background wind, table
// positions, the table entries the tile needs (each computed when first needed, level by level, with the routines
// the reference's table builder calls, :700-762), trilinear interpolation, relaxation, update of u3d / v3d.
// Layer bounds: global_z_interface + global_dz_interface (space_varying_dz, host, (gnx, nz, gny)) or layer_height +
// dz_levels (host, nz).  entries_computed (optional) receives the number of (entry, level) perturbations evaluated.
int lw_spatial_winds(LinWinds* lw, const Geom& g, int i0, int j0, const LtFields& f, const LtOptions& o, float* u3d,
                     float* v3d, const float* global_z_interface, const float* global_dz_interface,
                     const float* layer_height, const float* dz_levels, long long* entries_computed,
                     long long* launches, std::string* err);

}  // namespace icar
