// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of ICAR's MPDATA advection, following the reference Fortran
// statement by statement (same operand order, same loop coverage, same edge
// quirks).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, load or call this.
//
// PARITY PINNED TO THE REFERENCE'S OWN SOURCE.  The reference holds no golden
// vectors / known-answer tests for this path (src/tests/test_mpdata.f90 has no
// asserts and a stale signature) and no Fortran compiler
// exists in this image; so oracle/f90run EXECUTES src/physics/adv_mpdata.f90 and
// adv_mpdata_FCT_core.f90 themselves (translated statement by statement where
// they lie, IEEE binary32).  This restatement equals those outputs in every bit:
// advect3d (orders 1-3, FCT on / off, density), mpdata()'s wind normalisation and
// scalar list, the test_mpdata.f90 scenario, three coupled hill-case steps
// (tests/test_reference_source.py; tests/golden/reference_source_*.npz).
//
// Reference files followed (paths relative to /root/reference):
//   src/physics/adv_mpdata.f90:24-42    flux1
//   src/physics/adv_mpdata.f90:44-105   upwind_advection
//   src/physics/adv_mpdata.f90:107-255  mpdata_fluxes
//   src/physics/adv_mpdata.f90:257-354  flux_limiter
//   src/physics/adv_mpdata_FCT_core.f90:47-116  1-D FCT core
//   src/physics/adv_mpdata.f90:356-418  advect3d (order loop)
//   src/physics/adv_mpdata.f90:463-524  mpdata (wind normalisation)
//   src/physics/advect.f90:306-351      setup_module_winds (upwind variant)
//
// Array convention: Fortran column-major (i,k,j) with i fastest; all indices
// here are 0-based offsets from (ims,kms,jms).  Staggered module arrays keep
// the reference's compact shapes:
//   U_m (nx-1,nz,ny): element f is the x-face between cells f and f+1
//   V_m (nx,nz,ny-1): element g is the y-face between rows  g and g+1
//   W_m (nx,nz,ny)  : element k is the z-face above cell k
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include <algorithm>

namespace icar_oracle {

template <class R>
static inline R absr(R x) { return x < R(0) ? -x : x; }

// adv_mpdata.f90:24-42  f = ((U+ABS(U))*l + (U-ABS(U))*r)/2
template <class R>
static inline R flux1(R l, R r, R U) {
    return ((U + absr(U)) * l + (U - absr(U)) * r) / R(2);
}

template <class R>
struct Dims {
    int nx, nz, ny;
    inline size_t c(int i, int k, int j) const {              // cell / W face
        return (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)j);
    }
    inline size_t u(int f, int k, int j) const {              // U_m compact
        return (size_t)f + (size_t)(nx - 1) * ((size_t)k + (size_t)nz * (size_t)j);
    }
    inline size_t v(int i, int k, int g) const {              // V_m compact
        return (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)g);
    }
    inline size_t cells() const { return (size_t)nx * nz * ny; }
    inline size_t ufaces() const { return (size_t)(nx - 1) * nz * ny; }
    inline size_t vfaces() const { return (size_t)nx * nz * (ny - 1); }
};

// ---------------------------------------------------------------------------
// adv_mpdata.f90:500-506 (scheme==2) / advect.f90:344-347 (scheme==1)
// u is (nx+1,nz,ny), v is (nx,nz,ny+1), w, jac_w are (nx,nz,ny);
// jac_u (nx+1,nz,ny), jac_v (nx,nz,ny+1); rho may be null (== 1 everywhere,
// options%parameters%advect_density = .false.).
// ---------------------------------------------------------------------------
template <class R>
void wind_normalise(const R* u, const R* v, const R* w, const R* jac_u, const R* jac_v,
                    const R* jac_w, const R* rho, R dt, R dx, int nx, int nz, int ny,
                    int scheme, R* U_m, R* V_m, R* W_m) {
    Dims<R> d{nx, nz, ny};
    auto iu = [&](int i, int k, int j) {   // (nx+1,nz,ny)
        return (size_t)i + (size_t)(nx + 1) * ((size_t)k + (size_t)nz * (size_t)j);
    };
    auto rh = [&](int i, int k, int j) { return rho ? rho[d.c(i, k, j)] : R(1); };
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz; ++k) {
            for (int f = 0; f < nx - 1; ++f) {   // Fortran i = f+2 (1-based), face ims+1..ime
                R rs = rh(f + 1, k, j) + rh(f, k, j);
                R uu = u[iu(f + 1, k, j)], ju = jac_u[iu(f + 1, k, j)];
                U_m[d.u(f, k, j)] = (scheme == 2) ? (uu * dt * rs * R(0.5) * ju / dx)
                                                  : (uu * dt * ju * rs * R(0.5) / dx);
            }
            for (int i = 0; i < nx; ++i) {
                if (j >= 1) {                    // face between rows j-1 and j -> g = j-1
                    R rs = rh(i, k, j) + rh(i, k, j - 1);
                    R vv = v[d.c(i, k, j)], jv = jac_v[d.c(i, k, j)];   // (nx,nz,ny+1) shares strides
                    V_m[d.v(i, k, j - 1)] = (scheme == 2) ? (vv * dt * rs * R(0.5) * jv / dx)
                                                          : (vv * dt * jv * rs * R(0.5) / dx);
                }
                R ww = w[d.c(i, k, j)], jw = jac_w[d.c(i, k, j)];
                if (k < nz - 1)
                    W_m[d.c(i, k, j)] = ww * dt * jw * (rh(i, k + 1, j) + rh(i, k, j)) * R(0.5);
                else
                    W_m[d.c(i, k, j)] = ww * dt * jw * rh(i, k, j);
            }
        }
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:44-105  (identical stencil: advect.f90:106-178)
// ---------------------------------------------------------------------------
template <class R>
void upwind_advection(const R* qin, const R* u, const R* v, const R* w, const R* rho, R* q,
                      const R* dz, int nx, int nz, int ny, const R* jaco) {
    Dims<R> d{nx, nz, ny};
    std::memcpy(q, qin, d.cells() * sizeof(R));                       // :63-65
#pragma omp parallel
    {
    // the per-plane temporaries f1, f3, f4, f5 of the Fortran (:54-57), one set per thread
    std::vector<R> f1((size_t)(nx - 1) * nz), f3((size_t)(nx - 2) * nz),
        f4((size_t)(nx - 2) * nz), f5((size_t)(nx - 2) * (nz > 1 ? nz - 1 : 0));
#pragma omp for schedule(static)
    for (int j = 1; j < ny - 1; ++j) {                                  // :70
        for (int k = 0; k < nz; ++k) {
            for (int f = 0; f < nx - 1; ++f)                            // :74-75
                f1[f + (size_t)(nx - 1) * k] =
                    flux1(qin[d.c(f, k, j)], qin[d.c(f + 1, k, j)], u[d.u(f, k, j)]);
            for (int i = 1; i < nx - 1; ++i) {
                f3[(i - 1) + (size_t)(nx - 2) * k] =                    // :77-78  v(:,:,j+1) -> g=j
                    flux1(qin[d.c(i, k, j)], qin[d.c(i, k, j + 1)], v[d.v(i, k, j)]);
                f4[(i - 1) + (size_t)(nx - 2) * k] =                    // :80-81  v(:,:,j)   -> g=j-1
                    flux1(qin[d.c(i, k, j - 1)], qin[d.c(i, k, j)], v[d.v(i, k, j - 1)]);
                if (k < nz - 1)                                         // :83-84
                    f5[(i - 1) + (size_t)(nx - 2) * k] =
                        flux1(qin[d.c(i, k, j)], qin[d.c(i, k + 1, j)], w[d.c(i, k, j)]);
            }
        }
        for (int k = 0; k < nz; ++k)
            for (int i = 1; i < nx - 1; ++i) {
                size_t c = d.c(i, k, j);
                R rh = rho ? rho[c] : R(1);
                R f1p = f1[i + (size_t)(nx - 1) * k], f1m = f1[(i - 1) + (size_t)(nx - 1) * k];
                R f3v = f3[(i - 1) + (size_t)(nx - 2) * k], f4v = f4[(i - 1) + (size_t)(nx - 2) * k];
                R val = q[c] - ((f1p - f1m) + (f3v - f4v)) / (jaco[c] * rh);   // :88-89
                R den = dz[c] * jaco[c] * rh;
                if (nz == 1) { q[c] = val; continue; }
                if (k == 0) {                                                   // :95-96
                    val = val - f5[(i - 1)] / den;
                } else if (k == nz - 1) {                                       // :98-99
                    val = val - (qin[c] * w[c] - f5[(i - 1) + (size_t)(nx - 2) * (k - 1)]) / den;
                } else {                                                        // :92-93
                    val = val - (f5[(i - 1) + (size_t)(nx - 2) * k] -
                                 f5[(i - 1) + (size_t)(nx - 2) * (k - 1)]) / den;
                }
                q[c] = val;
            }
    }
    }
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:107-255.  `w` here is W_m/dz (advect3d passes (W_m/dz), :379),
// G = jaco*rho.  Outputs keep the compact face shapes.
// ---------------------------------------------------------------------------
template <class R>
void mpdata_fluxes(const R* q, const R* u, const R* v, const R* w, R* u2, R* v2, R* w2, int nx,
                   int nz, int ny, const R* G) {
    Dims<R> d{nx, nz, ny};
    const R eps = R(1e-10f);
    std::fill(u2, u2 + d.ufaces(), R(0));                               // :122-124
    std::fill(v2, v2 + d.vfaces(), R(0));
    std::fill(w2, w2 + d.cells(), R(0));
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {                                      // :129 (i in the Fortran)
        for (int k = 0; k < nz; ++k) {                                  // :130 (j in the Fortran)
            // ---- U component (:134-169); `if (i > 0)` is always true for jms >= 1
            for (int f = 0; f < nx - 1; ++f) {
                R rx = q[d.c(f + 1, k, j)], lx = q[d.c(f, k, j)];
                R denomx = rx + lx + eps;
                R uu = u[d.u(f, k, j)];
                R GG = G[d.c(f + 1, k, j)] + G[d.c(f, k, j)];
                R t = absr(uu) * (R(1) - absr(uu) / (R(0.5) * GG));
                t = t * (rx - lx) / denomx;
                if (j > 0 && j < ny - 1) {                              // UxV :148-155
                    R eq = (q[d.c(f + 1, k, j + 1)] - q[d.c(f + 1, k, j - 1)] +
                            q[d.c(f, k, j + 1)] - q[d.c(f, k, j - 1)]) /
                           (q[d.c(f + 1, k, j + 1)] + q[d.c(f + 1, k, j - 1)] +
                            q[d.c(f, k, j + 1)] + q[d.c(f, k, j - 1)] + eps);
                    // v(.,j,i)=face rows (j-1,j) -> g=j-1 ; v(.,j,i+1) -> g=j
                    R ev = R(0.25) * (v[d.v(f + 1, k, j - 1)] + v[d.v(f + 1, k, j)] +
                                      v[d.v(f, k, j - 1)] + v[d.v(f, k, j)]);
                    t = t - R(0.5) * uu * ev * eq / GG;
                }
                if (k > 0 && k < nz - 1) {                              // UxW :160-167
                    R eq = (q[d.c(f + 1, k + 1, j)] - q[d.c(f + 1, k - 1, j)] +
                            q[d.c(f, k + 1, j)] - q[d.c(f, k - 1, j)]) /
                           (q[d.c(f + 1, k + 1, j)] + q[d.c(f + 1, k - 1, j)] +
                            q[d.c(f, k + 1, j)] + q[d.c(f, k - 1, j)] + eps);
                    R ev = R(0.25) * (w[d.c(f + 1, k, j)] + w[d.c(f + 1, k - 1, j)] +
                                      w[d.c(f, k, j)] + w[d.c(f, k - 1, j)]);
                    t = t - R(0.5) * uu * ev * eq / GG;
                }
                u2[d.u(f, k, j)] = t;
            }
            // ---- V component (:172-211), face g=j-1 between rows j-1 and j
            if (j > 0) {
                for (int i = 0; i < nx; ++i) {
                    R r = q[d.c(i, k, j)], l = q[d.c(i, k, j - 1)];
                    R denom = r + l + eps;
                    R vv = v[d.v(i, k, j - 1)];
                    R GG = G[d.c(i, k, j)] + G[d.c(i, k, j - 1)];
                    R t = absr(vv) * (R(1) - absr(vv) / (R(0.5) * GG));
                    t = t * (r - l) / denom;
                    if (i > 0 && i < nx - 1) {                          // VxU :189-195 (edge_* = 0 at i=ims,ime)
                        R eq = (q[d.c(i + 1, k, j - 1)] - q[d.c(i - 1, k, j)] +
                                q[d.c(i + 1, k, j)] - q[d.c(i - 1, k, j - 1)]) /
                               (q[d.c(i + 1, k, j)] + q[d.c(i + 1, k, j - 1)] +
                                q[d.c(i - 1, k, j)] + q[d.c(i - 1, k, j - 1)] + eps);
                        // u(i+1)=face (i,i+1) -> f=i ; u(i)=face (i-1,i) -> f=i-1
                        R ev = R(0.25) * (u[d.u(i, k, j)] + u[d.u(i, k, j - 1)] +
                                          u[d.u(i - 1, k, j)] + u[d.u(i - 1, k, j - 1)]);
                        t = t - R(0.5) * vv * ev * eq / GG;
                    }
                    if (k > 0 && k < nz - 1) {                          // VxW :201-208
                        R eq = (q[d.c(i, k + 1, j - 1)] - q[d.c(i, k - 1, j)] +
                                q[d.c(i, k + 1, j)] - q[d.c(i, k - 1, j - 1)]) /
                               (q[d.c(i, k + 1, j - 1)] + q[d.c(i, k - 1, j)] +
                                q[d.c(i, k + 1, j)] + q[d.c(i, k - 1, j - 1)] + eps);
                        R ev = R(0.25) * (w[d.c(i, k, j)] + w[d.c(i, k - 1, j)] +
                                          w[d.c(i, k, j - 1)] + w[d.c(i, k - 1, j - 1)]);
                        t = t - R(0.5) * vv * ev * eq / GG;
                    }
                    v2[d.v(i, k, j - 1)] = t;
                }
            }
            // ---- W component (:217-250)
            if (k == nz - 1) {
                for (int i = 0; i < nx; ++i) w2[d.c(i, k, j)] = R(0);
            } else {
                for (int i = 0; i < nx; ++i) {
                    R r = q[d.c(i, k + 1, j)], l = q[d.c(i, k, j)];
                    R denom = r + l + eps;
                    R ww = w[d.c(i, k, j)];
                    R GG = G[d.c(i, k + 1, j)] + G[d.c(i, k, j)];
                    R t = absr(ww) * (R(1) - absr(ww) / (R(0.5) * GG));
                    t = t * (r - l) / denom;
                    if (i > 0 && i < nx - 1) {                          // WxU :230-236
                        R eq = (q[d.c(i + 1, k + 1, j)] - q[d.c(i - 1, k, j)] +
                                q[d.c(i + 1, k, j)] - q[d.c(i - 1, k + 1, j)]) /
                               (q[d.c(i + 1, k, j)] + q[d.c(i + 1, k + 1, j)] +
                                q[d.c(i - 1, k, j)] + q[d.c(i - 1, k + 1, j)] + eps);
                        R ev = R(0.25) * (u[d.u(i, k, j)] + u[d.u(i, k + 1, j)] +
                                          u[d.u(i - 1, k, j)] + u[d.u(i - 1, k + 1, j)]);
                        t = t - R(0.5) * ww * ev * eq / GG;
                    }
                    if (j > 0 && j < ny - 1) {                          // WxV :241-248
                        R eq = (q[d.c(i, k + 1, j + 1)] - q[d.c(i, k, j - 1)] +
                                q[d.c(i, k, j + 1)] - q[d.c(i, k + 1, j - 1)]) /
                               (q[d.c(i, k, j + 1)] + q[d.c(i, k + 1, j - 1)] +
                                q[d.c(i, k + 1, j + 1)] + q[d.c(i, k, j - 1)] + eps);
                        // v(.,k,j)->g=j-1, v(.,k+1,j)->g=j-1, v(.,k,j+1)->g=j, v(.,k+1,j+1)->g=j
                        R ev = R(0.25) * (v[d.v(i, k, j - 1)] + v[d.v(i, k + 1, j - 1)] +
                                          v[d.v(i, k, j)] + v[d.v(i, k + 1, j)]);
                        t = t - R(0.5) * ww * ev * eq / GG;
                    }
                    w2[d.c(i, k, j)] = t;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// adv_mpdata_FCT_core.f90:47-116 on one line of n cells (0..n-1), n-1 faces.
// q1 = field after pass 1, l = field before pass 1, U2 = pseudo-velocity
// (in/out), f = donor flux of the UNLIMITED U2.  Sequential, with the carried
// *_i2 -> *_i temporaries exactly as in the Fortran.
// ---------------------------------------------------------------------------
template <class R>
void fct_core(const R* q1, const R* l, R* U2, const R* f, int n, bool flux_is_w) {
    using std::max;
    using std::min;
    R qmax_i = 0, qmin_i = 0, qmax_i2 = 0, qmin_i2 = 0;
    R fin_i = 0, fout_i = 0, fin_i2 = 0, fout_i2 = 0;
    const R e15 = R(1e-15f);
    for (int i = 0; i < n - 1; ++i) {                                   // :47  do i=n_s,n-1
        if (i == 0) {                                                   // :50-55
            qmax_i = max(max(q1[i], q1[i + 1]), max(l[i], l[i + 1]));
            qmin_i = min(min(q1[i], q1[i + 1]), min(l[i], l[i + 1]));
            qmax_i2 = max(max(max(q1[i], q1[i + 1]), q1[i + 2]), max(max(l[i], l[i + 1]), l[i + 2]));
            qmin_i2 = min(min(min(q1[i], q1[i + 1]), q1[i + 2]), min(min(l[i], l[i + 1]), l[i + 2]));
        } else if (i != n - 2) {                                        // :56-61
            qmax_i = qmax_i2;
            qmin_i = qmin_i2;
            qmax_i2 = max(max(max(q1[i], q1[i + 1]), q1[i + 2]), max(max(l[i], l[i + 1]), l[i + 2]));
            qmin_i2 = min(min(min(q1[i], q1[i + 1]), q1[i + 2]), min(min(l[i], l[i + 1]), l[i + 2]));
        } else {                                                        // :62-68
            qmax_i = qmax_i2;
            qmin_i = qmin_i2;
            qmax_i2 = max(max(q1[i], q1[i + 1]), l[i]);
            qmin_i2 = min(min(q1[i], q1[i + 1]), l[i]);
        }
        if (i != 0) {                                                   // :72-85
            fin_i = fin_i2;
            fout_i = fout_i2;
        } else if (flux_is_w) {
            fin_i = R(0) - min(R(0), f[i]);
            fout_i = max(R(0), f[i]);
        } else {
            fin_i = 0;
            fout_i = 0;
        }
        if (i != n - 2) {                                               // :88-100
            fin_i2 = max(R(0), f[i]) - min(R(0), f[i + 1]);
            fout_i2 = max(R(0), f[i + 1]) - min(R(0), f[i]);
        } else if (flux_is_w) {
            fin_i2 = max(R(0), f[i]) - min(R(0), f[i]);
            fout_i2 = max(R(0), f[i]) - min(R(0), f[i]);
        } else {
            fin_i2 = 0;
            fout_i2 = 0;
        }
        if (U2[i] > R(0)) {                                             // :103-107
            R beta_out_i = (q1[i] - qmin_i) / (fout_i + e15);
            R beta_in_i2 = (qmax_i2 - q1[i + 1]) / (fin_i2 + e15);
            U2[i] = min(min(R(1), beta_in_i2), beta_out_i) * U2[i];
        } else if (U2[i] < R(0)) {                                      // :110-114
            R beta_in_i = (qmax_i - q1[i]) / (fin_i + e15);
            R beta_out_i2 = (q1[i + 1] - qmin_i2) / (fout_i2 + e15);
            U2[i] = min(min(R(1), beta_in_i), beta_out_i2) * U2[i];
        }
    }
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:257-354.  u (nx-1,nz,ny), v (nx,nz,ny-1), w (nx,nz,ny) in/out.
// ---------------------------------------------------------------------------
template <class R>
void flux_limiter(const R* q, const R* q2, R* u, R* v, R* w, int nx, int nz, int ny) {
    Dims<R> d{nx, nz, ny};
#pragma omp parallel
    {
        std::vector<R> q1x(nx), lx(nx), fx(nx - 1), U2x(nx - 1);
        std::vector<R> q1y(ny), ly(ny), fy(ny - 1), U2y(ny - 1);
        std::vector<R> q1z(nz), lz(nz), fz(nz > 1 ? nz - 1 : 0), U2z(nz > 1 ? nz - 1 : 0);
#pragma omp for schedule(static)
        for (int j = 1; j < ny - 1; ++j) {                              // :287
            for (int k = 0; k < nz; ++k) {                              // :294-304  x lines
                for (int i = 0; i < nx; ++i) { q1x[i] = q2[d.c(i, k, j)]; lx[i] = q[d.c(i, k, j)]; }
                for (int f = 0; f < nx - 1; ++f) U2x[f] = u[d.u(f, k, j)];
                for (int f = 0; f < nx - 1; ++f) fx[f] = flux1(q1x[f], q1x[f + 1], U2x[f]);
                fct_core(q1x.data(), lx.data(), U2x.data(), fx.data(), nx, false);
                for (int f = 0; f < nx - 1; ++f) u[d.u(f, k, j)] = U2x[f];
            }
            if (nz > 1)
                for (int i = 1; i < nx - 1; ++i) {                      // :313-323  z lines
                    for (int k = 0; k < nz; ++k) { q1z[k] = q2[d.c(i, k, j)]; lz[k] = q[d.c(i, k, j)]; }
                    for (int k = 0; k < nz - 1; ++k) U2z[k] = w[d.c(i, k, j)];
                    for (int k = 0; k < nz - 1; ++k) fz[k] = flux1(q1z[k], q1z[k + 1], U2z[k]);
                    fct_core(q1z.data(), lz.data(), U2z.data(), fz.data(), nz, true);
                    for (int k = 0; k < nz - 1; ++k) w[d.c(i, k, j)] = U2z[k];
                    w[d.c(i, nz - 1, j)] = R(0);
                }
        }
#pragma omp for schedule(static)
        for (int i = 0; i < nx; ++i) {                                  // :340-350  y lines, ALL i and k
            for (int k = 0; k < nz; ++k) {
                for (int j = 0; j < ny; ++j) { q1y[j] = q2[d.c(i, k, j)]; ly[j] = q[d.c(i, k, j)]; }
                for (int g = 0; g < ny - 1; ++g) U2y[g] = v[d.v(i, k, g)];
                for (int g = 0; g < ny - 1; ++g) fy[g] = flux1(q1y[g], q1y[g + 1], U2y[g]);
                fct_core(q1y.data(), ly.data(), U2y.data(), fy.data(), ny, false);
                for (int g = 0; g < ny - 1; ++g) v[d.v(i, k, g)] = U2y[g];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:356-418.  q in/out.  U_m,V_m,W_m are the module arrays.
// ---------------------------------------------------------------------------
template <class R>
void advect3d(R* q, const R* rho, int nx, int nz, int ny, const R* jaco, const R* dz,
              const R* U_m, const R* V_m, const R* W_m, int mpdata_order, bool fct) {
    Dims<R> d{nx, nz, ny};
    std::vector<R> q2(d.cells()), u2(d.ufaces()), v2(d.vfaces()), w2(d.cells());
    std::vector<R> wn, G;
    if (mpdata_order > 1) {
        wn.resize(d.cells());
        G.resize(d.cells());
#pragma omp parallel for schedule(static)
        for (long long c = 0; c < (long long)d.cells(); ++c) {
            wn[c] = W_m[c] / dz[c];                                     // (W_m/dz), :379
            G[c] = jaco[c] * (rho ? rho[c] : R(1));                     // (jaco*rho), :379
        }
    }
    for (int iord = 1; iord <= mpdata_order; ++iord) {
        if (iord == 1) {
            upwind_advection(q, U_m, V_m, W_m, rho, q2.data(), dz, nx, nz, ny, jaco);   // :374
        } else {
            mpdata_fluxes(q2.data(), U_m, V_m, wn.data(), u2.data(), v2.data(), w2.data(), nx, nz,
                          ny, G.data());                                                   // :379
#pragma omp parallel for schedule(static)
            for (long long c = 0; c < (long long)d.ufaces(); ++c) u2[c] = u2[c] * R(0.5);  // :383
#pragma omp parallel for schedule(static)
            for (long long c = 0; c < (long long)d.vfaces(); ++c) v2[c] = v2[c] * R(0.5);  // :384
#pragma omp parallel for schedule(static)
            for (long long c = 0; c < (long long)d.cells(); ++c) w2[c] = w2[c] * R(0.5) * dz[c];  // :385
            if (fct) flux_limiter(q, q2.data(), u2.data(), v2.data(), w2.data(), nx, nz, ny);     // :387
            upwind_advection(q2.data(), u2.data(), v2.data(), w2.data(), rho, q, dz, nx, nz, ny,
                             jaco);                                                        // :389
        }
        if (iord != mpdata_order) {
            if (iord > 1) std::memcpy(q2.data(), q, d.cells() * sizeof(R));                // :396-398
        } else if (iord == 1) {
            std::memcpy(q, q2.data(), d.cells() * sizeof(R));                              // :405-407
        }
    }
}

}  // namespace icar_oracle
