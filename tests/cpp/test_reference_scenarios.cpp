// The reference's own two test scenarios for this path, written against the C ABI / the C++ mirror of the plugin
// surface the way the Fortran programs are written against the modules:
//   src/tests/test_mpdata.f90:51-111      test_3d: step function in y, CFL 0.25, 3 x 3 x 100, wrap-around cells
//   src/tests/test_mp_simple.f90:4-33     a 5-level column cooled by 0.1 K per call, 100 calls
// and a third part that drives mp + advect through include/icar_b200_mod.hpp like step() does.
// Like the Fortran programs it prints; tests/test_cpp_host.py compares the printed rows with tests/golden/.
//
//   g++ -std=c++17 -Iinclude tests/cpp/test_reference_scenarios.cpp -Licar_b200 -licar_b200 -Wl,-rpath,$PWD/icar_b200
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "icar_b200_mod.hpp"

using namespace icar_b200_mod;

static icar_grid whole_domain(int nx, int nz, int ny) {      // grid_t of a single image (grid_obj.f90:116-255)
    icar_grid g{};
    g.ims = g.ids = 1; g.ime = g.ide = nx;
    g.kms = g.kds = g.kts = 1; g.kme = g.kde = g.kte = nz;
    g.jms = g.jds = 1; g.jme = g.jde = ny;
    g.its = 2; g.ite = nx - 1; g.jts = 2; g.jte = ny - 1;
    g.halo_size = 1;
    return g;
}

static void check(int rc, icar_ctx* ctx, const char* what) {
    if (rc) { std::fprintf(stderr, "%s failed (status %d): %s\n", what, rc, icar_last_error(ctx)); std::exit(3); }
}

// program test_mpdata / test_3d
static void test_3d(bool FCT) {
    const int nx = 3, ny = 100, nz = 3;
    const float cfl = 0.25f;
    const size_t n = (size_t)nx * nz * ny;
    auto at = [&](int i, int k, int j) { return (size_t)i + nx * ((size_t)k + nz * (size_t)j); };   // q(i,k,j), 0-based
    std::vector<float> q(n, 1.f), one(n, 1.f), w(n, 0.f);
    std::vector<float> u((size_t)(nx + 1) * nz * ny, 0.f), ju((size_t)(nx + 1) * nz * ny, 1.f);
    std::vector<float> v((size_t)nx * nz * (ny + 1), cfl), jv((size_t)nx * nz * (ny + 1), 1.f);
    for (int j = 0; j < ny; ++j)
        if (j + 1 > ny / 2) for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) q[at(i, k, j)] = 2.f;
    auto wrap = [&]() {                                   // q(:,:,1)=q(:,:,ny-1) ; q(:,:,ny)=q(:,:,2)
        std::memcpy(&q[at(0, 0, 0)], &q[at(0, 0, ny - 2)], sizeof(float) * nx * nz);
        std::memcpy(&q[at(0, 0, ny - 1)], &q[at(0, 0, 1)], sizeof(float) * nx * nz);
    };
    wrap();
    icar_grid g = whole_domain(nx, nz, ny);
    icar_options o;
    icar_default_options(&o);
    icar_ctx* ctx = nullptr;
    check(icar_create(0, &g, &o, &ctx), nullptr, "icar_create");
    // v = cfl with dt = dx = 1 and unit jacobians: V_m == 0.25 exactly (adv_mpdata.f90:500-506)
    check(icar_mpdata_set_winds(ctx, u.data(), v.data(), w.data(), ju.data(), jv.data(), one.data(), nullptr, 1.f, 1.f, 2),
          ctx, "icar_mpdata_set_winds");
    const int nloops = 1;                                  // the frozen vector holds the first revolution
    for (int loop = 1; loop <= nloops; ++loop) {
        for (int i = 1; i <= (int)((ny - 2) / cfl); ++i) {
            check(icar_mpdata_advect3d(ctx, q.data(), nullptr, 1, nx, 1, nz, 1, ny, one.data(), one.data(), 2, FCT ? 1 : 0),
                  ctx, "icar_mpdata_advect3d");
            wrap();
        }
        float mx = q[at(1, 0, 0)], mn = mx;
        std::printf("mpdata loop %d:", loop);
        for (int j = 0; j < ny; ++j) {
            const float x = q[at(1, 0, j)];
            mx = std::max(mx, x); mn = std::min(mn, x);
            if (j % (ny / 10) == 0) std::printf(" %.9g", x);
        }
        std::printf(" %.9g %.9g\n", mx, mn);
    }
    icar_destroy(ctx);
}

// program test_mp_simple
static void test_mp_simple() {
    const int nx = 3, nz = 5, ny = 3;
    const size_t n = (size_t)nx * nz * ny;
    auto at = [&](int i, int k, int j) { return (size_t)i + nx * ((size_t)k + nz * (size_t)j); };
    std::vector<float> p(n, 80000.f), th(n, 280.f), pii(n, 1.f), rho(n, 1.f), qv(n, 0.005f), qc(n, 0.f), qr(n, 0.f),
        qs(n, 0.f), dz(n, 200.f), rain((size_t)nx * ny, 0.f), snow((size_t)nx * ny, 0.f);
    icar_grid g = whole_domain(nx, nz, ny);
    icar_options o;
    icar_default_options(&o);
    icar_ctx* ctx = nullptr;
    check(icar_create(0, &g, &o, &ctx), nullptr, "icar_create");
    for (int call = 1; call <= 100; ++call) {
        check(icar_mp_simple_driver(ctx, p.data(), th.data(), pii.data(), rho.data(), qv.data(), qc.data(), qr.data(),
                                    qs.data(), rain.data(), snow.data(), 20.f, dz.data(), 1, nx, 1, ny, 1, nz, 2, 2, 2, 2, 1, nz),
              ctx, "icar_mp_simple_driver");
        std::printf("mp_simple call %d: %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", call, th[at(1, 0, 1)], qv[at(1, 0, 1)],
                    qc[at(1, 0, 1)], qr[at(1, 0, 1)], qs[at(1, 0, 1)], rain[1 + nx * 1], snow[1 + nx * 1]);
        for (int k = 0; k < nz; ++k) th[at(1, k, 1)] -= 0.1f;
    }
    icar_destroy(ctx);
}

// the plugin surface as step() uses it (time_step.f90:512-529) on a small flat-metric domain
static void test_plugin_surface() {
    const int nx = 24, nz = 10, ny = 20;
    const size_t n = (size_t)nx * nz * ny;
    domain_t domain;
    options_t options;
    domain.grid = whole_domain(nx, nz, ny);
    domain.dx = 4000.f;
    auto fill = [&](variable_t& var, float x) { var.data_3d.assign(n, x); };
    fill(domain.w, 0.f); fill(domain.pressure, 80000.f); fill(domain.exner, 0.938f); fill(domain.density, 1.f);
    fill(domain.potential_temperature, 285.f); fill(domain.water_vapor, 0.008f); fill(domain.cloud_water_mass, 0.f);
    fill(domain.rain_mass, 0.f); fill(domain.snow_mass, 0.f); fill(domain.dz_mass, 400.f);
    domain.u.data_3d.assign((size_t)(nx + 1) * nz * ny, 10.f);
    domain.v.data_3d.assign((size_t)nx * nz * (ny + 1), 5.f);
    domain.jacobian.assign(n, 1.f); domain.jacobian_w.assign(n, 1.f); domain.advection_dz.assign(n, 400.f);
    domain.jacobian_u.assign((size_t)(nx + 1) * nz * ny, 1.f);
    domain.jacobian_v.assign((size_t)nx * nz * (ny + 1), 1.f);
    domain.accumulated_precipitation.data_2dd.assign((size_t)nx * ny, 0.0);
    domain.accumulated_snowfall.data_2dd.assign((size_t)nx * ny, 0.0);
    for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) for (int i = 8; i < 14; ++i)      // a moist blob
        domain.water_vapor.data_3d[(size_t)i + nx * ((size_t)k + nz * (size_t)j)] = 0.012f;
    mp_var_request(options);
    adv_var_request(options);
    mp_init(options);                                        // init.f90:136
    adv_init(domain, options);                               // init.f90:146
    const float dt = 20.f;
    double mass0 = 0, mass1 = 0;
    for (float x : domain.water_vapor.data_3d) mass0 += x;
    for (int step = 0; step < 5; ++step) {
        mp(domain, options, dt, /*halo=*/1);                 // time_step.f90:512
        b200_halo_send();                                    // :515
        mp(domain, options, dt, -1, /*subset=*/1);           // :523
        b200_halo_retrieve();                                // :526
        advect(domain, options, dt);                         // :529
        domain.model_time += dt;                             // :547
    }
    b200_sync_to_host(domain);
    double cloud = 0, rain = 0;
    for (float x : domain.water_vapor.data_3d) mass1 += x;
    for (float x : domain.cloud_water_mass.data_3d) cloud += x;
    for (double x : domain.accumulated_precipitation.data_2dd) rain += x;
    std::printf("plugin surface: qv %.9g -> %.9g, cloud %.9g, precipitation %.9g, vars_to_advect %zu\n", mass0, mass1, cloud,
                rain, options.vars_to_advect.size());
    b200_finalize();
}

int main(int argc, char** argv) {
    try {
        const bool all = argc < 2;
        if (all || !std::strcmp(argv[1], "mpdata")) test_3d(true);
        if (all || !std::strcmp(argv[1], "mp_simple")) test_mp_simple();
        if (all || !std::strcmp(argv[1], "plugin")) test_plugin_surface();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error stop: %s\n", e.what());
        return 3;
    }
    return 0;
}
