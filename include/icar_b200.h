/* icar_b200.h -- C ABI of the B200-native replacement for ICAR's advection +
 * microphysics hot path (src/physics/advection*, src/physics/mp_*).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and returns an
 * int status (0 = ICAR_OK); a message for the last failure is available from
 * icar_last_error().  The reference reports errors with `stop` / `error stop`
 * (time_step.f90:322-328, mp_simple.f90:354-356, exchangeable_obj.f90:38); the
 * Fortran shim (icar_b200/fortran/icar_b200_mod.f90) turns a non-zero status
 * into `error stop`.
 *
 * Arrays are the reference's own: Fortran column-major, 32-bit real,
 *   3-D  (ims:ime, kms:kme, jms:jme)       x fastest, z middle, y slowest
 *   u    (ims:ime+1, kms:kme, jms:jme)     staggered in x  (domain_obj.f90, u_grid)
 *   v    (ims:ime, kms:kme, jms:jme+1)     staggered in y
 *   2-D  (ims:ime, jms:jme)
 * Index arguments are the Fortran 1-based inclusive bounds (grid_t, grid_obj.f90:144-255).
 *
 * Two ways in:
 *   resident  -- a context owns device copies of the domain fields; the plugin
 *                calls icar_mp / icar_halo_* / icar_advect operate on them.  This
 *                is the production path (fields stay in HBM across steps).
 *   host      -- leaf calls with host pointers that mirror the reference's leaf
 *                subroutines one-to-one (advect3d, mp_simple_driver); each call
 *                uploads, runs the same kernels and downloads.
 *
 * Threading: like the reference (one coarray image enters mp()/advect() with one
 * thread), a context is entered by one host thread at a time.  One context = one GPU.
 */
#ifndef ICAR_B200_H
#define ICAR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ICAR_OK 0
#define ICAR_ERR_INVALID 1   /* bad argument / call order                      */
#define ICAR_ERR_CUDA 2      /* CUDA runtime failure (no device, OOM, launch)  */
#define ICAR_ERR_PHYSICS 3   /* condition on which the reference would `stop`  */
#define ICAR_ERR_COMM 4      /* NCCL / neighbour exchange failure              */

typedef struct icar_ctx icar_ctx;

/* grid_t (src/objects/grid_h.f90; bounds computed in grid_obj.f90:144-255). */
typedef struct icar_grid {
    int ims, ime, kms, kme, jms, jme; /* memory bounds of this image (incl. halo)  */
    int its, ite, kts, kte, jts, jte; /* physics tile (excludes BC cell / halo)    */
    int ids, ide, kds, kde, jds, jde; /* whole model domain                        */
    int halo_size;                    /* kDEFAULT_HALO_SIZE = 1 in the reference   */
} icar_grid;

/* The members of options_t this path reads. Defaults: icar_default_options(). */
typedef struct icar_options {
    int advection;                /* options%physics%advection: 1 kADV_UPWIND, 2 kADV_MPDATA
                                     (icar_constants.f90:371-372)                            */
    int microphysics;             /* options%physics%microphysics: 0 none, 1 kMP_THOMPSON,
                                     2 kMP_SB04 (icar_constants.f90:347-348)                 */
    int mpdata_order;             /* adv_options%mpdata_order, default 2 (options_obj.f90:1580) */
    int flux_corrected_transport; /* adv_options%flux_corrected_transport, default 1 (:1579)   */
    int advect_density;           /* parameters%advect_density, default 0 (options_obj.f90:1023) */
    int update_interval;          /* mp_options%update_interval [s], default 0 (:1282)       */
    int top_mp_level;             /* mp_options%top_mp_level, default 0 = all (:1283)         */
    float dx;                     /* domain%dx [m]                                           */
    int exact_arithmetic;         /* no reference counterpart.  0 (default): MPDATA order 2 + FCT runs the
                                     fused kernel, whose anti-diffusive correction is evaluated with
                                     pre-multiplied coefficients, shared pair sums and MUFU-based division
                                     (differences ~1e-8 relative, tolerance 1e-6 in the tests);
                                     1: always run the staged kernels, which keep the Fortran's operand
                                     order and are bit-identical to the CPU oracle.                        */
} icar_options;

/* kVARS-style identifiers of the domain_t members the path touches
 * (src/objects/domain_h.f90:18-363).  3-D real unless noted. */
enum icar_var {
    ICAR_VAR_U = 0,                 /* domain%u%data_3d            (nx+1,nz,ny)            */
    ICAR_VAR_V = 1,                 /* domain%v%data_3d            (nx,nz,ny+1)            */
    ICAR_VAR_W = 2,                 /* domain%w%data_3d                                    */
    ICAR_VAR_PRESSURE = 3,          /* domain%pressure%data_3d                             */
    ICAR_VAR_EXNER = 4,             /* domain%exner%data_3d                                */
    ICAR_VAR_DENSITY = 5,           /* domain%density%data_3d                              */
    ICAR_VAR_POTENTIAL_TEMPERATURE = 6,
    ICAR_VAR_WATER_VAPOR = 7,
    ICAR_VAR_CLOUD_WATER = 8,
    ICAR_VAR_RAIN_IN_AIR = 9,
    ICAR_VAR_SNOW_IN_AIR = 10,
    ICAR_VAR_JACOBIAN = 11,         /* domain%jacobian                                     */
    ICAR_VAR_JACOBIAN_U = 12,       /* domain%jacobian_u           (nx+1,nz,ny)            */
    ICAR_VAR_JACOBIAN_V = 13,       /* domain%jacobian_v           (nx,nz,ny+1)            */
    ICAR_VAR_JACOBIAN_W = 14,       /* domain%jacobian_w                                   */
    ICAR_VAR_ADVECTION_DZ = 15,     /* domain%advection_dz                                 */
    ICAR_VAR_DZ_MASS = 16,          /* domain%dz_mass%data_3d                              */
    ICAR_VAR_PRECIPITATION = 17,    /* domain%accumulated_precipitation%data_2dd (nx,ny) REAL*8 */
    ICAR_VAR_SNOWFALL = 18,         /* domain%accumulated_snowfall%data_2dd     (nx,ny) REAL*8 */
    /* Thompson (kMP_THOMPSON) only -- mp_driver.f90:128-131, 415-424 */
    ICAR_VAR_CLOUD_ICE = 19,        /* domain%cloud_ice_mass%data_3d                       */
    ICAR_VAR_ICE_NUMBER = 20,       /* domain%cloud_ice_number%data_3d                     */
    ICAR_VAR_RAIN_NUMBER = 21,      /* domain%rain_number%data_3d                          */
    ICAR_VAR_GRAUPEL_IN_AIR = 22,   /* domain%graupel_mass%data_3d                         */
    ICAR_VAR_GRAUPEL = 23,          /* domain%graupel%data_2dd                 (nx,ny) REAL*8 */
    ICAR_VAR_COUNT = 24
};

/* mp_options_type (src/objects/opt_types.f90:31-45): the Thompson sensitivity parameters that
 * thompson_init copies into its module variables (mp_thompson.f90:397-417).  Defaults:
 * icar_default_mp_options() = options_obj.f90:1259-1280. */
typedef struct icar_mp_options {
    float Nt_c, TNO, am_s, rho_g;
    float av_s, bv_s, fv_s, av_i;
    float av_g, bv_g;
    float Ef_si, Ef_rs, Ef_rg, Ef_ri;
    float C_cubes, C_sqrd;
    float mu_r;
    float t_adjust;
    int Ef_rw_l, Ef_sw_l;
} icar_mp_options;

/* ------------------------------------------------------------------ lifecycle */

/* Fills *opt with the reference's namelist defaults (options_obj.f90:1023,1259-1284,1568-1593):
 * advection=2, microphysics=2, mpdata_order=2, FCT on, advect_density off. */
void icar_default_options(icar_options* opt);

/* grid_t%set_grid_dimensions restricted to an x-slab decomposition (ximages = nimages,
 * yimages = 1): grid_obj.f90:116-138 (my_n / my_start) and :209-255 (update_with_halos).
 * image is 1-based like this_image(). */
int icar_grid_xslab(int nx_global, int nz, int ny_global, int nimages, int image, int halo_size,
                    icar_grid* out);

/* Creates a context on CUDA device `device` for one image's tile.  Replaces the allocation half
 * of domain%init (domain_obj.f90:160-330) plus adv_init / mp_init (advection_driver.f90:24-37,
 * mp_driver.f90:74-198) for the schemes named in `opt`.  Fails with ICAR_ERR_CUDA when no
 * device / the CUDA build is unusable: there is no CPU fallback. */
int icar_create(int device, const icar_grid* grid, const icar_options* opt, icar_ctx** ctx);
int icar_destroy(icar_ctx* ctx);

/* Number of CUDA devices this process sees (0 without a GPU): the Fortran shim places image i on device
 * mod(i - 1, count), one image per GPU of the node. */
int icar_device_count(int* count);
const char* icar_last_error(const icar_ctx* ctx); /* ctx may be NULL: last creation error */

/* ------------------------------------------------------- resident domain state */

/* Copies one domain_t member between a host array (reference layout) and its device copy.
 * Replaces nothing in the reference (fields live in host memory there); these are the
 * hooks output / restart use (src/io/output_obj.f90 save_file, src/io/restart.f90:11-90). */
int icar_upload(icar_ctx* ctx, int var, const void* host);
int icar_download(icar_ctx* ctx, int var, void* host);
/* Device pointer and row pitch (in elements; rows are x-lines) of the current copy of `var`.
 * Advected scalars are double-buffered: the pointer changes after every icar_advect. */
int icar_device_ptr(icar_ctx* ctx, int var, void** dptr, long long* pitch_elems);

/* domain%model_time in seconds (used by mp()'s update-interval logic, mp_driver.f90:698-713). */
int icar_set_model_time(icar_ctx* ctx, double seconds);
int icar_get_model_time(icar_ctx* ctx, double* seconds);

/* ------------------------------------------------------------ plugin call surface */

/* advect(domain, options, dt)  -- src/physics/advection_driver.f90:51-77 -> mpdata()
 * (adv_mpdata.f90:463-524) or upwind() (advect.f90:380-418): wind normalisation, then every
 * advected scalar (mp_simple: qv, qc, qr, qs, th) through advect3d. */
int icar_advect(icar_ctx* ctx, float dt);

/* mp(domain, options, dt_in, halo, subset) -- src/physics/mp_driver.f90:673-772.
 * halo / subset < 0 mean "argument not present".  Implements the update_interval / mp_dt
 * bookkeeping (:698-713), top_mp_level (:723-725), process_halo's four strips (:609-658)
 * and the REAL*8 precipitation accumulation of process_subdomain (:587-595). */
int icar_mp(icar_ctx* ctx, float dt_in, int halo, int subset);

/* domain%halo_send / halo_retrieve / halo_exchange -- src/objects/domain_obj.f90:104-155,
 * exchangeable_obj.f90:138-356, for an x-slab decomposition (east/west neighbours only).
 * No-ops on a single image. */
int icar_halo_send(icar_ctx* ctx);
int icar_halo_retrieve(icar_ctx* ctx);
int icar_halo_exchange(icar_ctx* ctx);

/* The physics part of one pass of step()'s loop, src/main/time_step.f90:512-529 + :547:
 * mp(halo) -> halo_send -> mp(subset) -> halo_retrieve -> advect ; model_time += dt. */
int icar_physics_step(icar_ctx* ctx, float dt);

/* `nsteps` passes of that loop body at a fixed dt -- the `do while (domain%model_time < end_time)` loop of
 * step(), src/main/time_step.f90:464-549, between two update_dt changes (the caller shortens the last
 * step itself, :469-471, with one icar_physics_step).  use_graph != 0: pairs of steps are replayed from a
 * CUDA graph captured on first use (a pair, because advect() swaps every field with its output buffer),
 * which removes the per-launch host cost; the results are those of nsteps icar_physics_step calls, bit for
 * bit.  The first two steps of a context, odd remainders, and steps taken while the per-kernel profile is
 * on always run eagerly. */
int icar_physics_steps(icar_ctx* ctx, float dt, int nsteps, int use_graph);

/* ---- the rest of step()'s loop body, so that the loop can run without host round trips ("next" rows) */

/* The part of diagnostic_update (src/main/time_step.f90:49-198) that mp / advect read:
 * exner = exner_function(pressure) (:85, atm_utilities.f90:682-691) and
 * density = pressure / (Rd * potential_temperature * exner) (:96,:101), on the device copies. */
int icar_diagnostic_update(icar_ctx* ctx);

/* The rest of diagnostic_update (src/main/time_step.f90:88-108, 165-194): the fields that output and the other
 * physics read, computed on the device from the resident state and written to HOST arrays in the reference's
 * layout (any output may be NULL):
 *   temperature (nx,nz,ny) = potential_temperature * exner                                            (:96)
 *   pressure_interface, temperature_interface (nx,nz,ny): level means above kms, half-difference at kms (:88-90, :97-98)
 *   surface_pressure (nx,ny) = pressure_interface(:, kms, :)                                          (:93)
 *   u_mass, v_mass (nx,nz,ny) = means of the two faces of a cell                                      (:103-108)
 *   w_real (nx,nz,ny): the real vertical motion, u * dzdx and v * dzdy averaged to the mass grid plus
 *       jacobian * (w(k-1) + w(k)) / 2, on ims+1:ime-1, jms+1:jme-1; the boundary ring, which the reference never
 *       writes, comes back 0                                                                          (:165-194)
 * dzdx (nx+1,nz,ny) and dzdy (nx,nz,ny+1) are domain%dzdx / domain%dzdy (domain_obj.f90:893-900, 1417-1451), host
 * arrays, needed for w_real only.  exner must be current (icar_diagnostic_update).  Synchronous. */
int icar_diagnostic_output(icar_ctx* ctx, const float* dzdx, const float* dzdy, float* temperature,
                           float* pressure_interface, float* temperature_interface, float* surface_pressure,
                           float* u_mass, float* v_mass, float* w_real);

/* update_dt / compute_dt (src/main/time_step.f90:217-330, 375-423) on the device-resident winds:
 * the CFL maxima for cfl_strictness 1..5 (use_density = .false.), dt = cfl_reduction_factor / maxwind,
 * `co_min` over the images of the communicator (:413; ncclAllReduce(min) of one double) and the
 * 120 s cap (:421).  dz_levels = options%parameters%dz_levels(kms:kme).  Returns ICAR_ERR_PHYSICS where
 * the reference stops with "ERROR time step too small" (dt < 0.1 s, :322-328). */
int icar_compute_dt(icar_ctx* ctx, const float* dz_levels, float cfl_reduction_factor, int cfl_strictness,
                    double* dt_seconds);

/* domain%apply_forcing(dt) (src/objects/domain_obj.f90:2383-2448) for the 3-D variables of this path.
 * icar_upload_tendency stores var%dqdt_3d (the per-second change update_delta_fields derives from two
 * forcing time steps; same shape as the variable) on the device; icar_apply_forcing then does, for every
 * variable with a tendency,  q = q + dqdt * dt_seconds  (REAL*8 product and sum, rounded once):
 * throughout the tile for the diagnosed fields (u, v, w, pressure: force_boundaries = .False.,
 * domain_obj.f90:178-184, and :2431-2434 for w), on the domain edges this image owns for the advected
 * scalars (force_boundaries = .True., :2404-2421). */
int icar_upload_tendency(icar_ctx* ctx, int var, const float* host);
int icar_apply_forcing(icar_ctx* ctx, double dt_seconds);

/* balance_uvw(u, v, w, jaco_u, jaco_v, jaco_w, dz, dx, jaco, options) (src/physics/wind.f90:81-172, with
 * calc_divergence(horz_only=.True.), :175-228) on the device-resident winds: w is replaced by the vertical
 * wind that balances du/dx + dv/dy, integrated upward from the ground; dz = domain%advection_dz as at the
 * call sites wind.f90:330-332.  update_winds runs this once per forcing interval ("next" rank 3). */
int icar_balance_uvw(icar_ctx* ctx);

/* ---- linear-theory winds (src/physics/linear_winds.f90; "next" rank 3, BASELINE configs[4]).  update_winds calls
 * linear_perturb once per forcing interval (wind.f90:310-354), then balance_uvw.  The reference goes through a
 * look-up table over 10 speeds x 36 directions x 10 stabilities (initialize_spatial_winds :596-830, interpolated in
 * spatial_winds :840-1127) -- 3600 x nx x nz x ny x 2 REALs, 7 TB at 2048 x 2048 x 60.  This library evaluates the
 * perturbation directly for the background state (U, V, N^2) the caller passes: one LUT entry, computed when it
 * is needed, with the same routines the LUT builder calls.  2-D arrays are Fortran (nx, ny), first index fastest. */

/* add_buffer_topo(terrain, buffer_topo, smooth_window, buffer) (:351-418).  out: (nx + 2 buffer, ny + 2 buffer)
 * complex double, interleaved.  Pure host function (no GPU needed). */
void icar_add_buffer_topo(const float* terrain, int nx, int ny, int smooth_window, int buffer, double* out);

/* setup_linwinds (:1180-1222) + initialize_linear_theory_data (:426-472): the terrain buffered twice (window 5
 * over `buffer` cells, then a rim of 2), forward FFT / N, fftshift -> domain%terrain_frequency on the device;
 * wavenumber grids from opt.dx.  global_terrain = domain%global_terrain (ids:ide, jds:jde); every image passes the
 * same array.  buffer = lt_options%buffer (default 50, options_obj.f90:1452). */
int icar_linear_winds_init(icar_ctx* ctx, const float* global_terrain, int nx_global, int ny_global, int buffer);
/* size of the buffered (FFT) grid and the total buffer (lt buffer + 2, :1207) */
int icar_linear_winds_dims(icar_ctx* ctx, int* fft_nx, int* fft_ny, int* buffer);
/* domain%terrain_frequency, (fft_nx, fft_ny) complex double, interleaved */
int icar_linear_winds_terrain_frequency(icar_ctx* ctx, double* host);

/* linear_perturbation(U, V, Nsq, z_bottom, z_top, minimum_step, fourier_terrain, lt_data) (:239-344), both
 * specifics of the generic: z_bottom_2d / z_top_2d NULL -> linear_perturbation_constz with the scalar bounds;
 * otherwise linear_perturbation_varyingz with (nx_global, ny_global) host fields.  u_perturb / v_perturb (may be
 * NULL) receive lt_data%u_perturb / v_perturb, (fft_nx, fft_ny) complex double; they also stay on the device. */
int icar_linear_perturbation(icar_ctx* ctx, float U, float V, float Nsq, float z_bottom, float z_top,
                             const float* z_bottom_2d, const float* z_top_2d, float minimum_step,
                             double* u_perturb, double* v_perturb);

/* linear_perturb(domain, options, vsmooth, reverse, useDensity, update) (:1317-1366) on the resident u, v of this
 * image's tile, for a uniform background (U, V, Nsq): for every level the layer's perturbation as the LUT builder
 * computes it (:716-737: space_varying_dz -> bounds global_z_interface(:,z,:) - terrain and + global_dz_interface(:,z,:),
 * both (nx_global, nz, ny_global); else layer_height(z) -+ dz_levels(z)/2 with layer_height = z(ims,:,jms) - terrain(ims,jms)),
 * staggered to the u / v faces (:756-762), then, as spatial_winds does (:1094-1111),
 *   u_perturbation = u_perturbation * (1 - linear_update_fraction) + linear_update_fraction * new
 *   u = u + u_perturbation * linear_contribution                                     (v likewise).
 * Defaults of lt_parameters: minimum_layer_size 100, linear_contribution 1, linear_update_fraction 0.2
 * (options_obj.f90:1459-1479).  Follow with icar_balance_uvw, as update_winds does. */
int icar_linear_perturb(icar_ctx* ctx, float U, float V, float Nsq, const float* global_z_interface,
                        const float* global_dz_interface, const float* layer_height, const float* dz_levels,
                        float minimum_layer_size, float linear_contribution, float linear_update_fraction);

/* lt_parameters as spatial_winds reads them (options_obj.f90:1401-1530); defaults: icar_default_lt_options(). */
typedef struct icar_lt_options {
    int variable_N;               /* N^2 from the state (calc_stability) instead of N_squared, default 1          */
    int smooth_nsq;               /* vertical running mean + smooth_array of log N^2, default 1                     */
    int vert_smooth;              /* vsmooth, default 10                                                            */
    int stability_window_size;    /* winsz, default 10                                                              */
    float max_stability, min_stability;   /* 6e-4, 1e-7                                                             */
    float N_squared;              /* 3e-5 (used when variable_N = 0)                                                */
    float linear_contribution, linear_update_fraction;   /* 1, 0.2                                                  */
    float dirmax, dirmin, spdmax, spdmin; /* table axes: 2 pi, 0, 30, 0 (the N^2 axis spans log(min..max_stability)) */
    int n_dir_values, n_nsq_values, n_spd_values;        /* 24, 5, 6                                                */
    float minimum_layer_size;     /* 100                                                                            */
} icar_lt_options;
void icar_default_lt_options(icar_lt_options* opt);

/* domain%z%data_3d of this image's tile (host, (ims:ime, kms:kme, jms:jme)): calc_stability reads it when variable_N. */
int icar_linear_winds_set_z(icar_ctx* ctx, const float* z);

/* linear_perturb -> spatial_winds(domain, reverse = .False., vsmooth, winsz, update) (linear_winds.f90:840-1127) on
 * the resident state of the tile, with the reference's own semantics: per-cell log N^2 (calc_stability, or
 * N_squared), its vertical and horizontal smoothing, the column-mean background wind of every (i, j) column, each
 * cell's position in the (speed, direction, log N^2) table, trilinear interpolation between the eight entries
 * around it, relaxation of the stored perturbation and the update of u, v.  What differs is WHEN a table entry is
 * computed: the reference fills the whole table at start-up (initialize_spatial_winds :596-830; n_spd x n_dir x n_nsq
 * x nx x nz x ny x 2 REALs, 720 entries by default -- 1.4 TB at 2048 x 2048 x 60); here the entries a level needs
 * are evaluated when that level is processed (same routines, same values) and dropped afterwards.  A uniform flow
 * needs 8 of the 720.  Layer bounds as in icar_linear_perturb.  *entries_computed (may be NULL) receives the number
 * of (entry, level) evaluations.  Reads potential_temperature, exner, water_vapor and whichever of cloud_water,
 * cloud_ice, rain_in_air, snow_in_air the scheme allocates; follow with icar_balance_uvw. */
int icar_linear_perturb_spatial(icar_ctx* ctx, const icar_lt_options* opt, const float* global_z_interface,
                                const float* global_dz_interface, const float* layer_height, const float* dz_levels,
                                long long* entries_computed);

/* ---- restart / output files ("next" rank 4).  The reference writes one NetCDF file per image (output_t%save_file,
 * src/io/output_obj.f90:41-80, 404-458) and reads it back in restart_model (src/io/restart.f90:11-90); there is no
 * NetCDF library here, so the file is a self-describing little-endian binary (layout: icar_b200/restart.py, which
 * also converts it to NumPy arrays).  3-D variables are stored in the reference's FILE order (x, y, z) =
 * reshape(data_3d, order=[1,3,2]) (output_obj.f90:423-426); the reorder runs on the device. */

/* Writes every resident variable that has been uploaded or computed, the model time and mp()'s update-interval
 * memory (mp_driver.f90:698-713) for this image's tile. */
int icar_write_restart(icar_ctx* ctx, const char* path);
/* restart_model / read_restart_data: the file must have been written for the same tile (grid_t bounds) and the
 * same microphysics scheme -- otherwise ICAR_ERR_INVALID, where the reference stops in restart_domain_error.
 * Continuing from a file is bit-identical to not having stopped. */
int icar_read_restart(icar_ctx* ctx, const char* path);
/* Header of a file without a GPU: grid, scheme, model time and the variable ids it holds (vars: room for
 * ICAR_VAR_COUNT ints).  Any out pointer may be NULL.  Pure host function. */
int icar_restart_info(const char* path, icar_grid* grid, int* microphysics, double* model_time, int* nvars,
                      int* vars);

/* domain%enforce_limits (src/objects/domain_obj.f90:2228-2243): where (q < 0) q = 0 on every scalar the
 * microphysics scheme allocates. */
int icar_enforce_limits(icar_ctx* ctx);

/* step(domain, end_time, options) (src/main/time_step.f90:438-549): internal time steps until the model time reaches
 * end_time (seconds, same clock as icar_set_model_time), each one
 *     update_dt (:466, icar_compute_dt) -> clip to end_time (:469-471) -> diagnostic_update (:474) ->
 *     if dt > 1e-3: mp(halo) / halo_send / mp(subset) / halo_retrieve / advect (:512-529, icar_physics_step with
 *     real(dt)), apply_forcing(dt) (:537), enforce_limits in the last two updates (:540-542) -> model_time += dt (:547)
 * with no host round trip other than the one double of update_dt.  rad / lsm / pbl / convect (:489-509) are not part
 * of this path (0 in every BASELINE configuration).  steps_taken (may be NULL) receives the number of passes.
 * Errors are those of the calls it is made of ("time step too small" included, on every image together).
 * The clock is REAL*8 seconds; the reference's Time_type counts days in REAL*8, so (i) it may take one more pass whose dt
 * is a rounding leftover below 1e-3 s, in which nothing runs, and (ii) where `end_time - model_time` EQUALS 2 dt in
 * seconds the rounding of its day count decides the `<` of :540 (enforce_limits one update earlier); here equal is not
 * less.  oracle/f90run executes step() itself against this control flow (tests/test_reference_source.py::test_live_step_loop). */
int icar_step(icar_ctx* ctx, double end_time, const float* dz_levels, float cfl_reduction_factor, int cfl_strictness,
              int* steps_taken);

/* Blocks until all work queued by the calls above has finished on the device. */
int icar_synchronize(icar_ctx* ctx);

/* Device-side diagnostics accumulated since creation; replaces mp_simple's print blocks
 * (mp_simple.f90:248-278, :354-356).  out[0] = saturation iterations that hit 15,
 * out[1] = "mp_simple ERROR" out-of-bounds events, out[2] = phase_change `stop` events. */
int icar_mp_diagnostics(icar_ctx* ctx, long long out[3]);

/* Number of CUDA kernels this context has launched since creation. */
long long icar_kernel_launches(const icar_ctx* ctx);

/* The instantiation of the fused MPDATA corrector that serves this context's tile, as ncu prints it
 * ("k_corrector_fused<1, 40, 2>": pairs of levels per thread, level slots, TMA stages), or "" when the staged
 * kernels do (mpdata_order != 2, FCT off, advect_density, exact_arithmetic, more than 64 levels).  bench.py uses it to
 * match its roofline line with the committed ncu capture of the same symbol. */
int icar_fused_kernel_name(icar_ctx* ctx, char name[64]);

/* Self-test of the division the fused MPDATA corrector uses (csrc/corrector_prims.inc: div2x, the fast path of the
 * compiler's own div.rn.f32 on f32x2 pairs): packed[i] = div2x(a[i], b[i]) and ieee[i] = a[i] / b[i] for n (even) host
 * floats.  The reference divides with the Fortran `/` (adv_mpdata.f90:129-251, adv_mpdata_FCT_core.f90:103-114); the
 * two outputs must be equal bit for bit for normal operands (tests/test_gpu_parity.py). */
int icar_selftest_division(icar_ctx* ctx, const float* a, const float* b, float* packed, float* ieee, long long n);

/* Device-side stopwatch on the context's compute stream (CUDA events): icar_timer_start queues
 * the start mark, icar_timer_stop queues the stop mark, waits for it and returns the elapsed
 * device time in milliseconds between the two marks.  Replaces timer_t (timer_obj.f90:13-60),
 * which brackets physics with wall-clock timers in driver.f90:155-157. */
int icar_timer_start(icar_ctx* ctx);
int icar_timer_stop(icar_ctx* ctx, float* elapsed_ms);

/* Per-kernel device timing, for the roofline of the dominant kernel.  While enabled, every kernel family
 * the context launches is bracketed by CUDA events on the stream it is launched on (timer_t brackets
 * only whole phases, timer_obj.f90:13-60; the reference has no per-kernel timers).  icar_profile_read
 * waits for the queued work, returns per family the summed milliseconds and the number of launches
 * since the last read, and resets the sums.  names = max_entries x char[32]. */
int icar_profile_enable(icar_ctx* ctx, int on);
int icar_profile_read(icar_ctx* ctx, int max_entries, char (*names)[32], float* ms, long long* launches,
                      int* n_out);

/* ------------------------------------------------- Thompson microphysics (microphysics = 1) */

void icar_default_mp_options(icar_mp_options* opt);

/* thompson_init(mp_options) -- src/physics/mp_thompson.f90:342-766: derives the gamma-function
 * constants and builds the 29 REAL*8 lookup tables (qr_acr_qg :2853, qr_acr_qs :3014, freezeH2O :3273,
 * qi_aut_qs :3413, table_Efrw :3464, table_Efsw :3533) ON THE DEVICE (about 0.2 s; the reference
 * spends minutes and therefore caches them in *_mpt.dat files).  opt may be NULL (defaults).
 * icar_mp with microphysics = 1 calls this with the defaults if the caller has not. */
int icar_thompson_init(icar_ctx* ctx, const icar_mp_options* opt);

/* Copies lookup table `table_id` (0..28, in the allocation order of mp_thompson.f90:353-390:
 * tcg_racg, tmr_racg, tcr_gacr, tmg_gacr, tnr_racg, tnr_gacr, tcs_racs1, tmr_racs1, tcs_racs2,
 * tmr_racs2, tcr_sacr1, tms_sacr1, tcr_sacr2, tms_sacr2, tnr_racs1, tnr_racs2, tnr_sacr1, tnr_sacr2,
 * tpi_qcfz, tni_qcfz, tpi_qrfz, tpg_qrfz, tni_qrfz, tnr_qrfz, tps_iaus, tni_iaus, tpi_ide, t_Efrw,
 * t_Efsw) to `host` in the Fortran's column-major order -- the record layout of the reference's
 * qr_acr_qg_mpt.dat / qr_acr_qs_mpt.dat / freezeH2O_mpt.dat.  *count receives the element count;
 * host may be NULL to query it. */
int icar_thompson_table(icar_ctx* ctx, int table_id, double* host, long long* count);

/* mp_gt_driver(qv,qc,qr,qi,qs,qg,ni,nr,th,pii,p,dz,dt_in,itimestep, RAINNC,RAINNCV, SNOWNC,SNOWNCV,
 * GRAUPELNC,GRAUPELNCV,SR, ids..kde, ims..kme, its..kte) -- src/physics/mp_thompson.f90:772-1044, on host
 * arrays; same argument order.  SNOWNC, SNOWNCV, GRAUPELNC, GRAUPELNCV are OPTIONAL in the Fortran and
 * may be NULL; RAINNCV is only read (for SR, :915) and may be NULL (read as 0). */
int icar_mp_gt_driver(icar_ctx* ctx, float* qv, float* qc, float* qr, float* qi, float* qs, float* qg,
                      float* ni, float* nr, float* th, const float* pii, const float* p, const float* dz,
                      float dt_in, int itimestep, float* RAINNC, const float* RAINNCV, float* SNOWNC,
                      float* SNOWNCV, float* GRAUPELNC, float* GRAUPELNCV, float* SR, int ids, int ide,
                      int jds, int jde, int kds, int kde, int ims, int ime, int jms, int jme, int kms,
                      int kme, int its, int ite, int jts, int jte, int kts, int kte);

/* ------------------------------------------------- multi-image (one process per GPU) */

/* Same for a y-slab decomposition (ximages = 1, yimages = nimages): every image keeps full x-rows,
 * which keeps all kernels at full SIMD width and makes the exchanged faces contiguous planes.  This is what
 * bench.py uses for N > 1 by default; x-slabs (the reference's first axis) work the same way. */
int icar_grid_yslab(int nx_global, int nz, int ny_global, int nimages, int image, int halo_size,
                    icar_grid* out);

/* Which local columns (x-slabs) or planes (y-slabs), as 0-based offsets from ims / jms, one image
 * exchanges with its two neighbours for `halo_size` = h -- exchangeable_obj.f90:249-356 restated
 * as index ranges (put_west/put_south send data(start+h : start+2h-1), put_east/put_north send
 * data(n-2h+1 : n-h); retrieve_* fill data(start : start+h-1) and data(n-h+1 : n)):
 *   plan[0] first index SENT to the lower neighbour (west / south)
 *   plan[1] first index SENT to the upper neighbour (east / north)
 *   plan[2] first index RECEIVED from the lower neighbour
 *   plan[3] first index RECEIVED from the upper neighbour
 *   plan[4] 1 if a lower neighbour exists, plan[5] 1 if an upper neighbour exists
 *   plan[6] exchange axis: 0 = x (columns), 1 = y (planes);  plan[7] = h
 * The axis is read off the grid (which of ims/ime or jms/jme stops short of the domain bounds).
 * Pure host function (no GPU needed); icar_halo_send / icar_halo_retrieve use the same plan. */
int icar_halo_plan(const icar_grid* grid, int image, int nimages, int plan[8]);

/* Connecting the images.  Replaces exchangeable_t%set_neighbors (exchangeable_obj.f90:68-118): neighbours are
 * image-1 (west / south) and image+1 (east / north).
 *
 * The exchange is the reference's own pattern (exchangeable_obj.f90:249-356: coarray PUTs into the neighbour's
 * halo buffers + `sync images`): icar_halo_send stores this image's boundary cells straight into the neighbours'
 * ghost cells (peer-mapped device memory over NVLink, one kernel for all exchanged scalars and both sides) and raises
 * a flag there; icar_halo_retrieve waits for the neighbours' flags.  No message layer is on that path.
 *
 * One process per image (one GPU each, all on one box): image 1 calls icar_comm_unique_id and broadcasts the 128
 * bytes with whatever the host program already has (coarray co_broadcast, MPI, torch.distributed); then every image
 * calls icar_comm_init, which builds the NCCL communicator (used for the co_min of update_dt, icar_compute_dt) and
 * maps the neighbours' memory with CUDA IPC.
 *
 * All images in one process (one host thread driving several contexts, on one or several GPUs): icar_comm_init_local
 * with the contexts in image order.  There is no communicator then: icar_compute_dt returns the image's own time step
 * and the caller takes the minimum.  The caller must queue the step of every image before it waits for any of them;
 * with several images on ONE device, CUDA_DEVICE_MAX_CONNECTIONS must be at least 3 x the number of contexts on it
 * (an image's kernels wait on the device for its neighbours', so their streams must not share a hardware queue). */
int icar_comm_unique_id(char id[128]);
int icar_comm_init(icar_ctx* ctx, int image, int nimages, const char id[128]);
int icar_comm_init_local(icar_ctx** ctxs, int nimages);

/* ------------------------------------------------- host-pointer leaf calls (stateless) */

/* mpdata()'s wind normalisation, adv_mpdata.f90:500-506 (scheme 2) / setup_module_winds,
 * advect.f90:344-347 (scheme 1).  rho may be NULL (advect_density = .false. => rho = 1).
 * The normalised U_m, V_m, W_m stay on the device inside ctx, like the module arrays. */
int icar_mpdata_set_winds(icar_ctx* ctx, const float* u, const float* v, const float* w,
                          const float* jacobian_u, const float* jacobian_v,
                          const float* jacobian_w, const float* rho, float dt, float dx,
                          int scheme);

/* advect3d(q, rho, ims,ime,kms,kme,jms,jme, jaco, dz, options) -- adv_mpdata.f90:356-418.
 * q is updated in place.  Uses the winds of the last icar_mpdata_set_winds. */
int icar_mpdata_advect3d(icar_ctx* ctx, float* q, const float* rho, int ims, int ime, int kms,
                         int kme, int jms, int jme, const float* jaco, const float* dz,
                         int mpdata_order, int flux_corrected_transport);
/* Same in REAL*8, used for the 1e-12 mass-conservation check (no reference counterpart: the
 * reference is built with 32-bit reals). */
int icar_mpdata_set_winds_f64(icar_ctx* ctx, const double* u, const double* v, const double* w,
                              const double* jacobian_u, const double* jacobian_v,
                              const double* jacobian_w, const double* rho, double dt, double dx,
                              int scheme);
int icar_mpdata_advect3d_f64(icar_ctx* ctx, double* q, const double* rho, int ims, int ime,
                             int kms, int kme, int jms, int jme, const double* jaco,
                             const double* dz, int mpdata_order, int flux_corrected_transport);

/* mp_simple_driver(pressure,th,pii,rho,qv,qc,qr,qs,rain,snow,dt,dz, ims..kme, its..kte)
 * -- src/physics/mp_simple.f90:595-646.  th,qv,qc,qr,qs,rain,snow are updated in place. */
int icar_mp_simple_driver(icar_ctx* ctx, const float* pressure, float* th, const float* pii,
                          const float* rho, float* qv, float* qc, float* qr, float* qs,
                          float* rain, float* snow, float dt, const float* dz, int ims, int ime,
                          int jms, int jme, int kms, int kme, int its, int ite, int jts, int jte,
                          int kts, int kte);

/* Host mirror of the members of domain_t one step reads and writes.  All pointers are host
 * arrays in the reference layout; NULL members are left untouched. */
typedef struct icar_host_domain {
    /* read every step */
    const float *u, *v, *w, *pressure, *exner, *density;
    /* read and written every step */
    float *potential_temperature, *water_vapor, *cloud_water, *rain_in_air, *snow_in_air;
    double *accumulated_precipitation, *accumulated_snowfall;
    /* Thompson only (NULL otherwise): read and written every step */
    float *cloud_ice, *ice_number, *rain_number, *graupel_in_air;
    double *accumulated_graupel;
} icar_host_domain;

/* One physics step on HOST arrays: uploads the per-step inputs, runs icar_physics_step, and
 * downloads the per-step outputs.  The static metric fields (jacobians, dz) must have been
 * uploaded once with icar_upload (the reference passes `domain` to adv_init, so a shim can do
 * that there).  icar_host_register pins the host arrays so the copies run at PCIe speed. */
int icar_physics_step_host(icar_ctx* ctx, float dt, const icar_host_domain* dom);
/* The same with diagnostic_update's exner / density (src/main/time_step.f90:85-101) taken on the device from the
 * pressure and potential temperature of this call: dom->exner and dom->density must be NULL.  Two of the
 * thirteen per-step input fields do not cross PCIe. */
int icar_step_host(icar_ctx* ctx, float dt, const icar_host_domain* dom);
/* The host step with separate input and output arrays and a row range for the outputs, so that one model
 * step can be pipelined over several contexts that each hold a band of rows (upload of the next band,
 * kernels of this one and download of the previous one overlap; icar_b200/pipeline.py).  Every non-NULL
 * member of `in` is uploaded (whole tile); every non-NULL output member of `out` receives the tile's rows
 * row0..row1 (0-based y index, inclusive) at its place in an array whose row 0 is the tile's row 0.
 * derive_diagnostics != 0: exner / density as in icar_step_host. */
int icar_physics_step_host_io_async(icar_ctx* ctx, float dt, const icar_host_domain* in, const icar_host_domain* out,
                                    int row0, int row1, int derive_diagnostics);
/* (the _async form above only QUEUES the uploads, the kernels and the downloads on the context's streams and returns:
 * the caller queues one band after the other, so that the copy engines serve the bands in order, and waits with
 * icar_synchronize; host arrays must be pinned, icar_host_register, and must not be touched before that.) */
int icar_physics_step_host_io(icar_ctx* ctx, float dt, const icar_host_domain* in, const icar_host_domain* out,
                              int row0, int row1, int derive_diagnostics);
int icar_host_register(icar_ctx* ctx, const icar_host_domain* dom);
int icar_host_unregister(icar_ctx* ctx, const icar_host_domain* dom);

#ifdef __cplusplus
}
#endif
#endif /* ICAR_B200_H */
