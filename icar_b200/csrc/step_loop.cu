// See step_loop.cuh.  All three kernels are plain HBM streams (one pass over the fields they name).
#include "step_loop.cuh"

namespace icar {

namespace {

constexpr float Rd = 287.058f, cp = 1012.0f;       // src/constants/icar_constants.f90:391-393

__global__ void __launch_bounds__(128)
k_diagnostic_update(Geom g, const float* __restrict__ p, const float* __restrict__ th, float* __restrict__ exner,
                    float* __restrict__ rho) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= g.nx) return;
    const size_t c = g.at(i, k, j);
    const float pr = p[c];
    const float ex = powf(pr / 100000.f, Rd / cp);            // exner_function, atm_utilities.f90:687-688
    const float t = th[c] * ex;                               // time_step.f90:96
    exner[c] = ex;
    rho[c] = pr / (Rd * t);                                   // time_step.f90:101
}

// diagnostic_update's interface fields (time_step.f90:88-99): temperature = th * exner; pressure_i / temperature_i are
// the level means ((k-1) + k) / 2 above the ground and a half-difference extrapolation at kms
__global__ void __launch_bounds__(128)
k_diagnostic_thermo(Geom g, const float* __restrict__ p, const float* __restrict__ th, const float* __restrict__ exner,
                    float* __restrict__ temperature, float* __restrict__ p_i, float* __restrict__ t_i) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= g.nx) return;
    const size_t c = g.at(i, k, j);
    // the neighbour level: k-1 above the ground, kms+1 at the ground (a one-level model has no interface rule: copy)
    const size_t o = (k > 0) ? c - g.ld : ((g.nz > 1) ? c + g.ld : c);
    const float pc = p[c], po = p[o];
    const float tc = th[c] * exner[c], to = th[o] * exner[o];
    if (temperature) temperature[c] = tc;
    if (p_i) p_i[c] = (k > 0) ? (po + pc) / 2.f : pc + (pc - po) / 2.f;
    if (t_i) t_i[c] = (k > 0) ? (to + tc) / 2.f : tc + (tc - to) / 2.f;
}

// u_mass, v_mass (time_step.f90:103-108) and the real vertical motion w_real (:165-194) on the mass grid
__global__ void __launch_bounds__(128)
k_diagnostic_winds(Geom g, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
                   const float* __restrict__ dzdx, const float* __restrict__ dzdy, const float* __restrict__ jaco,
                   float* __restrict__ u_mass, float* __restrict__ v_mass, float* __restrict__ w_real) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= g.nx) return;
    const size_t c = g.at(i, k, j), cu = g.atu(i, k, j), ps = g.ps();
    const float u0 = u[cu], u1 = u[cu + 1], v0 = v[c], v1 = v[c + ps];
    if (u_mass) u_mass[c] = (u1 + u0) / 2.f;
    if (v_mass) v_mass[c] = (v1 + v0) / 2.f;
    if (w_real) {
        float r = 0.f;                                    // the boundary ring is never written by the reference: stays 0
        if (i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1) {
            const float uw0 = u0 * dzdx[cu], uw1 = u1 * dzdx[cu + 1];
            const float vw0 = v0 * dzdy[c], vw1 = v1 * dzdy[c + ps];
            const float currw = w[c], lastw = (k > 0) ? w[c - g.ld] : 0.f;
            r = ((uw0 + uw1) * 0.5f + (vw0 + vw1) * 0.5f) + jaco[c] * (lastw + currw) * 0.5f;
        }
        w_real[c] = r;
    }
}

__global__ void __launch_bounds__(128)
k_cfl_maxima(Geom g, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
             const float* __restrict__ dzl, float dx, unsigned* __restrict__ out) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    float m3 = 0.f, mu = 0.f, mv = 0.f, mw = 0.f;
    if (i < g.nx) {
        const size_t c = g.at(i, k, j), cu = g.atu(i, k, j);
        const float u0 = fabsf(u[cu]), u1 = fabsf(u[cu + 1]);
        const float v0 = fabsf(v[c]), v1 = fabsf(v[c + g.ps()]);
        const float w0 = fabsf(w[c]), wl = (k == 0) ? w0 : fabsf(w[c - g.ld]);       // zoffset, :267-271
        m3 = fmaxf(u0, u1) / dx + fmaxf(v0, v1) / dx + fmaxf(w0, wl) / dzl[k];      // :282-284
        mu = fmaxf(u0, u1); mv = fmaxf(v0, v1); mw = w0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
        mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
        mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
        mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, __float_as_uint(m3));
        atomicMax(out + 1, __float_as_uint(mu));
        atomicMax(out + 2, __float_as_uint(mv));
        atomicMax(out + 3, __float_as_uint(mw));
    }
}

__global__ void __launch_bounds__(128)
k_apply_forcing(float* __restrict__ q, const float* __restrict__ dqdt, int width, int nz, int rows_y, int pitch,
                double seconds, int boundaries_only, int west, int east, int south, int north) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= width) return;
    if (boundaries_only) {
        const bool on_sn = (south && j == 0) || (north && j == rows_y - 1);
        const bool on_we = ((west && i == 0) || (east && i == width - 1)) && j > 0 && j < rows_y - 1;
        if (!on_sn && !on_we) return;
    }
    const size_t c = (size_t)i + (size_t)pitch * ((size_t)k + (size_t)nz * (size_t)j);
    q[c] = (float)((double)q[c] + ((double)dqdt[c] * seconds));
}

__global__ void __launch_bounds__(128)
k_balance_uvw(Geom g, const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ w,
              const float* __restrict__ ju, const float* __restrict__ jv, const float* __restrict__ jw,
              const float* __restrict__ dz, float dx) {
    const int i = blockIdx.x * 128 + threadIdx.x, j = blockIdx.y;
    if (i >= g.nx) return;
    float w_below = 0.f, jw_below = 1.f;
    for (int k = 0; k < g.nz; ++k) {
        const size_t c = g.at(i, k, j), cu = g.atu(i, k, j);
        // calc_divergence, horz_only: u_met = u*jaco_u, v_met = v*jaco_v; div = (diff_U + diff_V)/dx   (:209-216)
        const float diff_U = u[cu + 1] * ju[cu + 1] - u[cu] * ju[cu];
        const float diff_V = v[c + g.ps()] * jv[c + g.ps()] - v[c] * jv[c];
        const float div = (diff_U + diff_V) / dx;
        const float jwc = jw[c];
        float wk;
        if (k == 0) wk = 0.f - div * dz[c] / jwc;                              // :142
        else        wk = (w_below * jw_below - div * dz[c]) / jwc;            // :144
        w[c] = wk;
        w_below = wk; jw_below = jwc;
    }
}

struct FieldPtrs { float* p[12]; };

__global__ void __launch_bounds__(128)
k_enforce_limits(Geom g, FieldPtrs f, int nf) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= g.nx) return;
    const size_t c = g.at(i, k, j);
    for (int n = 0; n < nf; ++n) {
        const float v = f.p[n][c];
        if (v < 0.f) f.p[n][c] = 0.f;
    }
}

dim3 row_grid(const Geom& g) { return dim3((g.nx + 127) / 128, g.nz, g.ny); }

// (x, z, y) memory order <-> (x, y, z) file order of output_obj.f90:423-426 / restart.f90:57-63
__global__ void __launch_bounds__(128)
k_output_reorder(float* __restrict__ field, float* __restrict__ dense, int width, int nz, int rows_y, int pitch, int to_file) {
    const int i = blockIdx.x * 128 + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
    if (i >= width) return;
    const size_t m = (size_t)i + (size_t)pitch * ((size_t)k + (size_t)nz * (size_t)j);
    const size_t f = (size_t)i + (size_t)width * ((size_t)j + (size_t)rows_y * (size_t)k);
    if (to_file) dense[f] = field[m]; else field[m] = dense[f];
}

}  // namespace

int diagnostic_update_launch(const Geom& g, const float* pressure, const float* th, float* exner, float* density,
                             cudaStream_t st) {
    k_diagnostic_update<<<row_grid(g), 128, 0, st>>>(g, pressure, th, exner, density);
    return 1;
}

int diagnostic_thermo_launch(const Geom& g, const float* pressure, const float* th, const float* exner,
                             float* temperature, float* p_i, float* t_i, cudaStream_t st) {
    k_diagnostic_thermo<<<row_grid(g), 128, 0, st>>>(g, pressure, th, exner, temperature, p_i, t_i);
    return 1;
}

int diagnostic_winds_launch(const Geom& g, const float* u, const float* v, const float* w, const float* dzdx,
                            const float* dzdy, const float* jaco, float* u_mass, float* v_mass, float* w_real,
                            cudaStream_t st) {
    k_diagnostic_winds<<<row_grid(g), 128, 0, st>>>(g, u, v, w, dzdx, dzdy, jaco, u_mass, v_mass, w_real);
    return 1;
}

int cfl_maxima_launch(const Geom& g, const float* u, const float* v, const float* w, const float* dz_levels,
                      float dx, unsigned* out4, cudaStream_t st) {
    k_cfl_maxima<<<row_grid(g), 128, 0, st>>>(g, u, v, w, dz_levels, dx, out4);
    return 1;
}

int apply_forcing_launch(float* q, const float* dqdt, int width, int nz, int rows_y, int pitch, double seconds,
                         bool boundaries_only, const bool edges[4], cudaStream_t st) {
    k_apply_forcing<<<dim3((width + 127) / 128, nz, rows_y), 128, 0, st>>>(q, dqdt, width, nz, rows_y, pitch, seconds,
                                                                          boundaries_only ? 1 : 0, edges[0], edges[1],
                                                                          edges[2], edges[3]);
    return 1;
}

int balance_uvw_launch(const Geom& g, const float* u, const float* v, float* w, const float* ju, const float* jv,
                       const float* jw, const float* dz, float dx, cudaStream_t st) {
    k_balance_uvw<<<dim3((g.nx + 127) / 128, g.ny), 128, 0, st>>>(g, u, v, w, ju, jv, jw, dz, dx);
    return 1;
}

int output_reorder_launch(float* field, float* dense, int width, int nz, int rows_y, int pitch, bool to_file,
                          cudaStream_t st) {
    k_output_reorder<<<dim3((width + 127) / 128, nz, rows_y), 128, 0, st>>>(field, dense, width, nz, rows_y, pitch,
                                                                           to_file ? 1 : 0);
    return 1;
}

int enforce_limits_launch(const Geom& g, float* const* fields, int nfields, cudaStream_t st) {
    FieldPtrs f{};
    for (int n = 0; n < nfields && n < 12; ++n) f.p[n] = fields[n];
    k_enforce_limits<<<row_grid(g), 128, 0, st>>>(g, f, nfields);
    return 1;
}

}  // namespace icar
