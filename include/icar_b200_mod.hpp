// C++ mirror of ICAR's advection / microphysics plugin surface over libicar_b200 (include/icar_b200.h).
//
// The reference is Fortran and its binding is icar_b200/fortran/icar_b200_mod.f90; no Fortran compiler exists in
// this repository's build image, so this header is the same module in a host language that does compile here:
// the three procedures per plugin family the reference uses, with the reference's names, argument meaning and
// error behaviour,
//     adv_var_request / adv_init / advect          src/physics/advection_driver.f90:24-77
//     mp_var_request  / mp_init  / mp              src/physics/mp_driver.f90:74-237, :673-772
// plus what replaces domain%halo_send / halo_retrieve (domain_obj.f90:104-144) for device-resident scalars and
// the download before output (driver.f90:169-195).  A failing call throws (the reference prints and `stop`s).
//
// domain_t / options_t carry only the members this path reads, under the reference's member names
// (src/objects/domain_h.f90, options_h.f90, opt_types.f90); arrays are dense, in the reference's (i,k,j) order.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "icar_b200.h"

namespace icar_b200_mod {

constexpr int kADV_UPWIND = 1, kADV_MPDATA = 2;          // icar_constants.f90
constexpr int kMP_THOMPSON = 1, kMP_SB04 = 2;

struct variable_t {                                      // variable_t: data_3d or data_2d
    std::vector<float> data_3d;
    std::vector<double> data_2dd;                        // REAL*8 accumulators (accumulated_precipitation ...)
};

struct physics_type { int advection = kADV_MPDATA, microphysics = kMP_SB04; };                     // opt_types.f90
struct adv_options_type { bool flux_corrected_transport = true; int mpdata_order = 2; };
struct mp_options_type : icar_mp_options {
    int update_interval = 0, top_mp_level = 0;
    mp_options_type() { icar_default_mp_options(this); }
};
struct parameter_options { bool advect_density = false; };

struct options_t {
    physics_type physics;
    adv_options_type adv_options;
    mp_options_type mp_options;
    parameter_options parameters;
    std::vector<int> vars_to_advect, vars_to_allocate;   // filled by the *_var_request procedures
};

struct domain_t {
    icar_grid grid{};                                    // grid_t
    float dx = 4000.f;
    double model_time = 0.0;                             // seconds
    variable_t u, v, w, pressure, exner, density, potential_temperature, water_vapor, cloud_water_mass, rain_mass,
        snow_mass, cloud_ice_mass, cloud_ice_number, rain_number, graupel_mass, dz_mass;
    variable_t accumulated_precipitation, accumulated_snowfall, graupel;
    std::vector<float> jacobian, jacobian_u, jacobian_v, jacobian_w, advection_dz;
};

namespace detail {
inline icar_ctx*& ctx() { static icar_ctx* c = nullptr; return c; }
inline bool& state_on_device() { static bool s = false; return s; }
inline bool& thompson() { static bool t = false; return t; }
inline void check(int rc, const char* what) {
    if (rc) throw std::runtime_error(std::string(what) + ": " + icar_last_error(ctx()) + " (status " + std::to_string(rc) + ")");
}
inline void up3(int var, const std::vector<float>& a) { if (!a.empty()) check(icar_upload(ctx(), var, a.data()), "icar_upload"); }
inline void down3(int var, std::vector<float>& a) { if (!a.empty()) check(icar_download(ctx(), var, a.data()), "icar_download"); }
}  // namespace detail

// advection_driver.f90:39-49 / mp_driver.f90:200-237: the scalars the schemes ask to have advected
inline void adv_var_request(options_t& options) {
    options.vars_to_allocate.insert(options.vars_to_allocate.end(), {ICAR_VAR_U, ICAR_VAR_V, ICAR_VAR_W});
}
inline void mp_var_request(options_t& options) {
    std::vector<int> s = {ICAR_VAR_POTENTIAL_TEMPERATURE, ICAR_VAR_WATER_VAPOR, ICAR_VAR_CLOUD_WATER, ICAR_VAR_RAIN_IN_AIR,
                          ICAR_VAR_SNOW_IN_AIR};
    if (options.physics.microphysics == kMP_THOMPSON)
        s.insert(s.end(), {ICAR_VAR_CLOUD_ICE, ICAR_VAR_ICE_NUMBER, ICAR_VAR_RAIN_NUMBER, ICAR_VAR_GRAUPEL_IN_AIR});
    options.vars_to_advect = s;
}

// mp_driver.f90:74-198: Thompson's thompson_init(options%mp_options); nothing for mp_simple.  init_physics calls it
// before adv_init (init.f90:136,146): without a context yet, adv_init applies the options.
inline void mp_init(const options_t& options) {
    if (options.physics.microphysics != kMP_THOMPSON || !detail::ctx()) return;
    detail::check(icar_thompson_init(detail::ctx(), &options.mp_options), "icar_thompson_init");
}

// advection_driver.f90:24-37: creates the device context and uploads the static metric terms
inline void adv_init(const domain_t& domain, const options_t& options, int device = 0) {
    icar_options o;
    icar_default_options(&o);
    o.advection = options.physics.advection;
    o.microphysics = options.physics.microphysics;
    o.mpdata_order = options.adv_options.mpdata_order;
    o.flux_corrected_transport = options.adv_options.flux_corrected_transport ? 1 : 0;
    o.advect_density = options.parameters.advect_density ? 1 : 0;
    o.update_interval = options.mp_options.update_interval;
    o.top_mp_level = options.mp_options.top_mp_level;
    o.dx = domain.dx;
    int rc = icar_create(device, &domain.grid, &o, &detail::ctx());
    if (rc) throw std::runtime_error(std::string("icar_create: ") + icar_last_error(nullptr) + " (status " + std::to_string(rc) + ")");
    detail::up3(ICAR_VAR_JACOBIAN, domain.jacobian);
    detail::up3(ICAR_VAR_JACOBIAN_U, domain.jacobian_u);
    detail::up3(ICAR_VAR_JACOBIAN_V, domain.jacobian_v);
    detail::up3(ICAR_VAR_JACOBIAN_W, domain.jacobian_w);
    detail::up3(ICAR_VAR_ADVECTION_DZ, domain.advection_dz);
    detail::up3(ICAR_VAR_DZ_MASS, domain.dz_mass.data_3d);
    detail::thompson() = options.physics.microphysics == kMP_THOMPSON;
    detail::state_on_device() = false;
    mp_init(options);
}

// Resident mode: the state goes up once, at the first mp() call; after that the loop body of step() that touches it
// runs on the device, and the reference call sites that would change the HOST copies are rerouted:
// diagnostic_update -> b200_diagnostic_update, domain%apply_forcing -> b200_apply_forcing, domain%enforce_limits ->
// b200_enforce_limits, new forcing tendencies -> b200_update_forcing, update_winds on the host -> b200_upload_winds.
inline void b200_upload_winds(const domain_t& domain) {
    using namespace detail;
    up3(ICAR_VAR_U, domain.u.data_3d); up3(ICAR_VAR_V, domain.v.data_3d); up3(ICAR_VAR_W, domain.w.data_3d);
}
inline void sync_to_device(const domain_t& domain) {
    using namespace detail;
    if (!state_on_device()) {
        b200_upload_winds(domain);
        up3(ICAR_VAR_PRESSURE, domain.pressure.data_3d);
        up3(ICAR_VAR_POTENTIAL_TEMPERATURE, domain.potential_temperature.data_3d);
        up3(ICAR_VAR_WATER_VAPOR, domain.water_vapor.data_3d);
        up3(ICAR_VAR_CLOUD_WATER, domain.cloud_water_mass.data_3d);
        up3(ICAR_VAR_RAIN_IN_AIR, domain.rain_mass.data_3d);
        up3(ICAR_VAR_SNOW_IN_AIR, domain.snow_mass.data_3d);
        if (thompson()) {
            up3(ICAR_VAR_CLOUD_ICE, domain.cloud_ice_mass.data_3d); up3(ICAR_VAR_ICE_NUMBER, domain.cloud_ice_number.data_3d);
            up3(ICAR_VAR_RAIN_NUMBER, domain.rain_number.data_3d); up3(ICAR_VAR_GRAUPEL_IN_AIR, domain.graupel_mass.data_3d);
        }
        if (domain.exner.data_3d.empty() || domain.density.data_3d.empty()) {
            check(icar_diagnostic_update(ctx()), "icar_diagnostic_update");       // exner, density from p and th
        } else {            // a caller that brings its own diagnosed fields (the reference scenarios do)
            up3(ICAR_VAR_EXNER, domain.exner.data_3d);
            up3(ICAR_VAR_DENSITY, domain.density.data_3d);
        }
        state_on_device() = true;
    }
    check(icar_set_model_time(ctx(), domain.model_time), "icar_set_model_time");
}
// diagnostic_update(domain, options), time_step.f90:49-198, for what mp / advect read
inline void b200_diagnostic_update() {
    if (detail::state_on_device()) detail::check(icar_diagnostic_update(detail::ctx()), "icar_diagnostic_update");
}
// the output-only fields of diagnostic_update (time_step.f90:88-108, 165-194) from the device state; any pointer may be null
inline void b200_diagnostic_output(const float* dzdx, const float* dzdy, float* temperature, float* pressure_interface,
                                   float* temperature_interface, float* surface_pressure, float* u_mass, float* v_mass,
                                   float* w_real) {
    detail::check(icar_diagnostic_output(detail::ctx(), dzdx, dzdy, temperature, pressure_interface, temperature_interface,
                                         surface_pressure, u_mass, v_mass, w_real), "icar_diagnostic_output");
}
// var%dqdt_3d of a forced variable (update_delta_fields, domain_obj.f90:2318-2381) goes to the device
inline void b200_update_forcing(int var, const std::vector<float>& dqdt_3d) {
    if (!dqdt_3d.empty()) detail::check(icar_upload_tendency(detail::ctx(), var, dqdt_3d.data()), "icar_upload_tendency");
}
// domain%apply_forcing(dt), domain_obj.f90:2383-2448; domain%enforce_limits(), :2228-2243
inline void b200_apply_forcing(double dt_seconds) { detail::check(icar_apply_forcing(detail::ctx(), dt_seconds), "icar_apply_forcing"); }
inline void b200_enforce_limits() { detail::check(icar_enforce_limits(detail::ctx()), "icar_enforce_limits"); }

// step(domain, end_time, options), time_step.f90:438-549, with the whole loop on the device; returns the passes taken
inline int b200_step(double end_time_seconds, const float* dz_levels, float cfl_reduction_factor, int cfl_strictness) {
    int n = 0;
    detail::check(icar_step(detail::ctx(), end_time_seconds, dz_levels, cfl_reduction_factor, cfl_strictness, &n), "icar_step");
    return n;
}

// mp(domain, options, dt_in, halo, subset) -- mp_driver.f90:673; halo / subset < 0 == "not present"
inline void mp(domain_t& domain, const options_t&, float dt_in, int halo = -1, int subset = -1) {
    if (halo >= 0 || subset < 0) sync_to_device(domain);          // first mp call of the step
    detail::check(icar_mp(detail::ctx(), dt_in, halo, subset), "icar_mp");
}

// domain%halo_send / halo_retrieve (domain_obj.f90:104-144) for the device-resident scalars
inline void b200_halo_send() { detail::check(icar_halo_send(detail::ctx()), "icar_halo_send"); }
inline void b200_halo_retrieve() { detail::check(icar_halo_retrieve(detail::ctx()), "icar_halo_retrieve"); }

// advect(domain, options, dt) -- advection_driver.f90:51
inline void advect(domain_t&, const options_t&, float dt) { detail::check(icar_advect(detail::ctx(), dt), "icar_advect"); }

// before output / restart / any host physics that reads the scalars (driver.f90:169-195)
inline void b200_sync_to_host(domain_t& domain) {
    using namespace detail;
    down3(ICAR_VAR_POTENTIAL_TEMPERATURE, domain.potential_temperature.data_3d);
    down3(ICAR_VAR_WATER_VAPOR, domain.water_vapor.data_3d);
    down3(ICAR_VAR_CLOUD_WATER, domain.cloud_water_mass.data_3d);
    down3(ICAR_VAR_RAIN_IN_AIR, domain.rain_mass.data_3d);
    down3(ICAR_VAR_SNOW_IN_AIR, domain.snow_mass.data_3d);
    if (thompson()) {
        down3(ICAR_VAR_CLOUD_ICE, domain.cloud_ice_mass.data_3d); down3(ICAR_VAR_ICE_NUMBER, domain.cloud_ice_number.data_3d);
        down3(ICAR_VAR_RAIN_NUMBER, domain.rain_number.data_3d); down3(ICAR_VAR_GRAUPEL_IN_AIR, domain.graupel_mass.data_3d);
    }
    auto down2 = [&](int var, std::vector<double>& a) { if (!a.empty()) check(icar_download(ctx(), var, a.data()), "icar_download"); };
    down2(ICAR_VAR_PRECIPITATION, domain.accumulated_precipitation.data_2dd);
    down2(ICAR_VAR_SNOWFALL, domain.accumulated_snowfall.data_2dd);
    if (thompson()) down2(ICAR_VAR_GRAUPEL, domain.graupel.data_2dd);
}

inline void b200_finalize() {
    if (detail::ctx()) icar_destroy(detail::ctx());
    detail::ctx() = nullptr;
}

}  // namespace icar_b200_mod
