// ORACLE -- TEST INFRASTRUCTURE ONLY (see mpdata_oracle.hpp).  C entry points
// over the templated restatement so tests/ and bench.py's cpu_baseline leg can
// call it through ctypes.  PARITY UNPINNED by the reference.
#include "mpdata_oracle.hpp"
#include "mp_simple_oracle.hpp"

#include <omp.h>

using namespace icar_oracle;

#define ORACLE_INSTANTIATE(SUF, R)                                                                  \
    extern "C" void oracle_wind_normalise_##SUF(const R* u, const R* v, const R* w, const R* ju,    \
                                                const R* jv, const R* jw, const R* rho, R dt, R dx, \
                                                int nx, int nz, int ny, int scheme, R* U_m,         \
                                                R* V_m, R* W_m) {                                   \
        wind_normalise<R>(u, v, w, ju, jv, jw, rho, dt, dx, nx, nz, ny, scheme, U_m, V_m, W_m);     \
    }                                                                                               \
    extern "C" void oracle_upwind_advection_##SUF(const R* qin, const R* u, const R* v, const R* w, \
                                                  const R* rho, R* q, const R* dz, int nx, int nz,  \
                                                  int ny, const R* jaco) {                          \
        upwind_advection<R>(qin, u, v, w, rho, q, dz, nx, nz, ny, jaco);                            \
    }                                                                                               \
    extern "C" void oracle_mpdata_fluxes_##SUF(const R* q, const R* u, const R* v, const R* w,      \
                                               R* u2, R* v2, R* w2, int nx, int nz, int ny,         \
                                               const R* G) {                                        \
        mpdata_fluxes<R>(q, u, v, w, u2, v2, w2, nx, nz, ny, G);                                    \
    }                                                                                               \
    extern "C" void oracle_flux_limiter_##SUF(const R* q, const R* q2, R* u, R* v, R* w, int nx,    \
                                              int nz, int ny) {                                     \
        flux_limiter<R>(q, q2, u, v, w, nx, nz, ny);                                                \
    }                                                                                               \
    extern "C" void oracle_advect3d_##SUF(R* q, const R* rho, int nx, int nz, int ny,               \
                                          const R* jaco, const R* dz, const R* U_m, const R* V_m,   \
                                          const R* W_m, int order, int fct) {                       \
        advect3d<R>(q, rho, nx, nz, ny, jaco, dz, U_m, V_m, W_m, order, fct != 0);                  \
    }                                                                                               \
    extern "C" void oracle_mp_simple_driver_##SUF(const R* p, R* th, const R* pii, const R* rho,    \
                                                  R* qv, R* qc, R* qr, R* qs, R* rain, R* snow,     \
                                                  R dt, const R* dz, int nx, int nz, int ny,        \
                                                  int its, int ite, int jts, int jte, int kts,      \
                                                  int kte, long long* diag3) {                      \
        MpSimpleDiag dg;                                                                            \
        mp_simple_driver<R>(p, th, pii, rho, qv, qc, qr, qs, rain, snow, dt, dz, nx, nz, ny, its,   \
                            ite, jts, jte, kts, kte, &dg);                                          \
        if (diag3) {                                                                                \
            diag3[0] = dg.not_converged;                                                            \
            diag3[1] = dg.out_of_bounds;                                                            \
            diag3[2] = dg.negative_stop;                                                            \
        }                                                                                           \
    }                                                                                               \
    /* one column, the scenario of src/tests/test_mp_simple.f90 (T is temperature, not theta) */    \
    extern "C" void oracle_mp_simple_column_##SUF(const R* p, R* T, const R* rho, R* qv, R* qc,     \
                                                  R* qr, R* qs, R* rain, R* snow, R dt,             \
                                                  const R* dz, int nz, int kts, int kte) {          \
        MpSimple<R> mp;                                                                             \
        mp.cloud2snow = std::exp(R(-1.0f) * (R(1) / R(2000.0f)) * dt);                              \
        mp.cloud2rain = std::exp(R(-1.0f) * (R(1) / R(500.0f)) * dt);                               \
        mp.diag = nullptr;                                                                          \
        std::vector<R> fr(nz), fl(nz);                                                              \
        mp.column(p, T, rho, qv, qc, qr, qs, *rain, *snow, dt, dz, nz, kts, kte, fr.data(),         \
                  fl.data());                                                                       \
    }

ORACLE_INSTANTIATE(f32, float)
ORACLE_INSTANTIATE(f64, double)

// One model step of the hot path as time_step.f90:512-529 runs it on a single
// image with mp_simple + MPDATA: mp over the physics tile (its..ite, jts..jte =
// everything but the outermost boundary-condition cell, grid_obj.f90:241-250),
// then wind normalisation and advect3d of the five advected scalars in the
// order of adv_mpdata.f90:512-522 (qv, qc, qr, qs, th).  This is the CPU
// baseline timed by bench.py; scratch = 3 fields (U_m,V_m,W_m).
extern "C" void oracle_step_simple_f32(const float* p, float* th, const float* pii,
                                       const float* rho, float* qv, float* qc, float* qr,
                                       float* qs, float* rain, float* snow, const float* dz_mass,
                                       const float* u, const float* v, const float* w,
                                       const float* ju, const float* jv, const float* jw,
                                       const float* jaco, const float* adv_dz, float mp_dt,
                                       float dt, float dx, int nx, int nz, int ny, int order,
                                       int fct, float* U_m, float* V_m, float* W_m) {
    mp_simple_driver<float>(p, th, pii, rho, qv, qc, qr, qs, rain, snow, mp_dt, dz_mass, nx, nz,
                            ny, 1, nx - 2, 1, ny - 2, 0, nz - 1, nullptr);
    wind_normalise<float>(u, v, w, ju, jv, jw, nullptr, dt, dx, nx, nz, ny, 2, U_m, V_m, W_m);
    float* scalars[5] = {qv, qc, qr, qs, th};
    for (int s = 0; s < 5; ++s)
        advect3d<float>(scalars[s], nullptr, nx, nz, ny, jaco, adv_dz, U_m, V_m, W_m, order,
                        fct != 0);
}

extern "C" int oracle_max_threads(void) { return omp_get_max_threads(); }
extern "C" void oracle_set_threads(int n) { omp_set_num_threads(n); }
