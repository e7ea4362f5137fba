// ORACLE -- TEST INFRASTRUCTURE ONLY (see mpdata_oracle.hpp).  C entry points
// over the templated restatement so tests/ and bench.py's cpu_baseline leg can
// call it through ctypes.  Parity: pinned to the reference source executed by oracle/f90run.
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

// ------------------------------------------------------------------ Thompson (mp = 1)
#include "mp_thompson_oracle.hpp"

extern "C" void* oracle_thompson_create(const ThompsonOptions* opt, int big_tables) {
    Thompson* t = new Thompson();
    t->init(opt ? *opt : ThompsonOptions(), big_tables != 0);
    return t;
}
extern "C" void oracle_thompson_destroy(void* h) { delete static_cast<Thompson*>(h); }

// table by id, in the order thompson_init allocates them (mp_thompson.f90:353-390); returns the
// element count and a pointer to the fp64 data (Fortran column-major)
extern "C" long long oracle_thompson_table(void* h, int id, const double** data) {
    Thompson* t = static_cast<Thompson*>(h);
    std::vector<double>* tabs[] = {
        &t->tcg_racg, &t->tmr_racg, &t->tcr_gacr, &t->tmg_gacr, &t->tnr_racg, &t->tnr_gacr,
        &t->tcs_racs1, &t->tmr_racs1, &t->tcs_racs2, &t->tmr_racs2, &t->tcr_sacr1, &t->tms_sacr1,
        &t->tcr_sacr2, &t->tms_sacr2, &t->tnr_racs1, &t->tnr_racs2, &t->tnr_sacr1, &t->tnr_sacr2,
        &t->tpi_qcfz, &t->tni_qcfz, &t->tpi_qrfz, &t->tpg_qrfz, &t->tni_qrfz, &t->tnr_qrfz,
        &t->tps_iaus, &t->tni_iaus, &t->tpi_ide, &t->t_Efrw, &t->t_Efsw};
    if (id < 0 || id >= (int)(sizeof(tabs) / sizeof(tabs[0]))) return -1;
    if (data) *data = tabs[id]->data();
    return (long long)tabs[id]->size();
}
// install externally supplied contents for table `id` (used to cache the two big tables on disk)
extern "C" int oracle_thompson_set_table(void* h, int id, const double* data, long long n) {
    const double* cur = nullptr;
    long long have = oracle_thompson_table(h, id, &cur);
    if (have != n) return 1;
    std::copy(data, data + n, const_cast<double*>(cur));
    return 0;
}
// a few (k,m) slabs of the two big collision tables, for CPU tests that cannot afford all 1369
extern "C" void oracle_thompson_big_slab(void* h, int which, int km, int ij0, int ij1) {
    Thompson* t = static_cast<Thompson*>(h);
    if (which == 0) t->qr_acr_qg_km(km, ij0, ij1);
    else t->qr_acr_qs_km(km, ij0, ij1);
}
// scalar constants thompson_init derives, for cross-checks: out[0..] = mu_c, Sc3, D0i, xm0s, xm0g,
// am_r, am_i, am_g, then cce[3] ccg[3] cie[7] cig[7] cre[13] crg[13] cse[18] csg[18] cge[12] cgg[12],
// then the 16 t?_q?_?? rate prefactors
extern "C" int oracle_thompson_constants(void* h, float* out) {
    Thompson* t = static_cast<Thompson*>(h);
    int n = 0;
    for (float v : {t->mu_c, t->Sc3, t->D0i, t->xm0s, t->xm0g, t->am_r, t->am_i, t->am_g}) out[n++] = v;
    for (float v : t->cce) out[n++] = v;
    for (float v : t->ccg) out[n++] = v;
    for (float v : t->cie) out[n++] = v;
    for (float v : t->cig) out[n++] = v;
    for (float v : t->cre) out[n++] = v;
    for (float v : t->crg) out[n++] = v;
    for (float v : t->cse) out[n++] = v;
    for (float v : t->csg) out[n++] = v;
    for (float v : t->cge) out[n++] = v;
    for (float v : t->cgg) out[n++] = v;
    for (float v : {t->t1_qr_qc, t->t1_qr_qi, t->t2_qr_qi, t->t1_qg_qc, t->t1_qs_qc, t->t1_qs_qi, t->t1_qr_ev,
                    t->t2_qr_ev, t->t1_qs_sd, t->t2_qs_sd, t->t1_qg_sd, t->t2_qg_sd, t->t1_qs_me, t->t2_qs_me,
                    t->t1_qg_me, t->t2_qg_me}) out[n++] = v;
    return n;
}
extern "C" void oracle_thompson_bins(void* h, double* out500) {   // Dc, Di, Dr, Ds, Dg (100 each)
    Thompson* t = static_cast<Thompson*>(h);
    const double* b[5] = {t->Dc, t->Di, t->Dr, t->Ds, t->Dg};
    for (int a = 0; a < 5; ++a) std::copy(b[a], b[a] + 100, out500 + 100 * a);
}
// mp_gt_driver (mp_thompson.f90:772-1044); bounds 0-based inclusive, ide0/jde0 see driver()
extern "C" long long oracle_thompson_driver(void* h, float* qv, float* qc, float* qr, float* qi, float* qs,
                                            float* qg, float* ni, float* nr, float* th, const float* pii,
                                            const float* p, const float* dz, float dt, float* rainnc,
                                            const float* rainncv, float* snownc, float* graupelnc, float* sr,
                                            int nx, int nz, int ny, int its, int ite, int jts, int jte, int kts,
                                            int kte, int ide0, int jde0) {
    long long nm = 0;
    static_cast<Thompson*>(h)->driver(qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz, dt, rainnc, rainncv, snownc,
                                      graupelnc, sr, nx, nz, ny, its, ite, jts, jte, kts, kte, ide0, jde0, &nm);
    return nm;
}
// one column (T is temperature); ppt[4] = rain, snow, graupel, ice, accumulated
extern "C" int oracle_thompson_column(void* h, float* qv, float* qc, float* qi, float* qr, float* qs, float* qg,
                                      float* ni, float* nr, float* T, const float* p, const float* dz, float* ppt,
                                      int nk, float dt) {
    return static_cast<Thompson*>(h)->column(qv, qc, qi, qr, qs, qg, ni, nr, T, p, dz, ppt[0], ppt[1], ppt[2],
                                             ppt[3], nk, dt) ? 1 : 0;
}

// One model step with Thompson + MPDATA on a single image (time_step.f90:512-529): mp_gt_driver over
// the physics tile, REAL*8 accumulation of the per-call precipitation (mp_driver.f90:587-595), then
// wind normalisation and advect3d of the nine advected scalars in the order of adv_mpdata.f90:512-522
// (qv, qc, qr, qs, th, qi, qg, ni, nr).  acc = accumulated precipitation / snowfall / graupel (REAL*8).
extern "C" void oracle_step_thompson_f32(void* h, const float* p, float* th, const float* pii, float* qv,
                                         float* qc, float* qr, float* qi, float* qs, float* qg, float* ni,
                                         float* nr, double* acc_precip, double* acc_snow, double* acc_graupel,
                                         float* sr, const float* dz_mass, const float* u, const float* v,
                                         const float* w, const float* ju, const float* jv, const float* jw,
                                         const float* jaco, const float* adv_dz, float mp_dt, float dt,
                                         float dx, int nx, int nz, int ny, int order, int fct, float* U_m,
                                         float* V_m, float* W_m) {
    std::vector<float> rain((size_t)nx * ny, 0.f), snow((size_t)nx * ny, 0.f), graupel((size_t)nx * ny, 0.f);
    static_cast<Thompson*>(h)->driver(qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz_mass, mp_dt, rain.data(),
                                      nullptr, snow.data(), graupel.data(), sr, nx, nz, ny, 1, nx - 2, 1, ny - 2,
                                      0, nz - 1, nx - 1, ny - 1, nullptr);
    for (size_t c = 0; c < (size_t)nx * ny; ++c) {
        acc_precip[c] = acc_precip[c] + (double)rain[c];
        acc_graupel[c] = acc_graupel[c] + (double)graupel[c];
        acc_snow[c] = acc_snow[c] + (double)snow[c];
    }
    wind_normalise<float>(u, v, w, ju, jv, jw, nullptr, dt, dx, nx, nz, ny, 2, U_m, V_m, W_m);
    float* scalars[9] = {qv, qc, qr, qs, th, qi, qg, ni, nr};
    for (int s = 0; s < 9; ++s)
        advect3d<float>(scalars[s], nullptr, nx, nz, ny, jaco, adv_dz, U_m, V_m, W_m, order, fct != 0);
}
