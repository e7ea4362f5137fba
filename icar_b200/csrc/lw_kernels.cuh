// The kernels of the linear-theory wind perturbation (see linear_winds.cu for the routine-by-routine map to
// src/physics/linear_winds.f90).  They use nothing but lw_math.cuh, plain pointers and blockIdx / threadIdx, so that
// tests/cpp/lw_kernels_host.cpp can run these very bodies on the CPU (one loop iteration per thread, stubs for the
// built-ins) between NumPy transforms and compare a whole icar_linear_perturb with the restatement without a GPU.
// The product includes this file only from linear_winds.cu.
#pragma once
#include <cstddef>

#include "lw_math.cuh"

#if !defined(__CUDACC__)                         // host emulation (tests only)
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return {x, y}; }
struct lw_emu_idx { unsigned x, y, z; };
extern thread_local lw_emu_idx blockIdx, threadIdx;
#define __global__
#define __launch_bounds__(n)
#endif

namespace icar {
namespace lwk {

using lw::cplx;

constexpr int kThreads = 128;

// terrain_frequency = fftshift(FFT(terrain) / (nx ny))     (:1217-1222)
static __global__ void __launch_bounds__(kThreads)
k_lw_shift_norm(const double2* __restrict__ in, double2* __restrict__ out, int fnx, int fny, double cells) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i >= fnx) return;
    const double2 a = in[(size_t)j * fnx + i];
    const cplx v = lw::through_single({a.x / cells, a.y / cells});
    out[(size_t)lw::shifted(j, fny) * fnx + lw::shifted(i, fnx)] = make_double2(v.x, v.y);
}

// uhat, vhat at their FFT-order positions (ifftshift folded in), summed over `nsteps` heights that start at z0
// and advance by `step` in REAL arithmetic exactly as linear_perturbation_constz's loop does (:262-268)
static __global__ void __launch_bounds__(kThreads)
k_lw_spectrum(const double2* __restrict__ tf, const float* __restrict__ kx, const float* __restrict__ ly, int fnx,
              int fny, float U, float V, float Nsq, float z0, float step, int nsteps, double2* __restrict__ uhat,
              double2* __restrict__ vhat) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i >= fnx) return;
    const int si = lw::shifted(i, fnx), sj = lw::shifted(j, fny);      // ifftshift reads its (i, j) from (ii, jj)
    const double2 t = tf[(size_t)sj * fnx + si];
    const lw::Mode q = lw::make_mode(U, V, Nsq, kx[si], ly[sj], {t.x, t.y});
    cplx us = {0.0, 0.0}, vs = {0.0, 0.0};
    float z = z0;
    for (int s = 0; s < nsteps; ++s) {
        cplx uh, vh;
        lw::hats_at(q, z, &uh, &vh);
        us.x += uh.x; us.y += uh.y;
        vs.x += vh.x; vs.y += vh.y;
        z = z + step;
    }
    uhat[(size_t)j * fnx + i] = make_double2(us.x, us.y);
    vhat[(size_t)j * fnx + i] = make_double2(vs.x, vs.y);
}

// u_perturb = u_accumulator / n_steps   (:271-272)
static __global__ void __launch_bounds__(kThreads)
k_lw_divide(double2* __restrict__ u, double2* __restrict__ v, size_t n, double by) {
    const size_t c = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    u[c] = make_double2(u[c].x / by, u[c].y / by);
    v[c] = make_double2(v[c].x / by, v[c].y / by);
}

// internal_z_top / internal_z_bottom (:305-309): the domain's layer bounds placed at (buffer : buffer + n - 1)
// of the FFT grid -- as written, one cell west / south of where the terrain sits -- inside their extreme values
static __global__ void __launch_bounds__(kThreads)
k_lw_layer_bounds(const float* __restrict__ zb, const float* __restrict__ zt, int nx, int ny, int buffer, float zb_min,
                  float zt_max, float* __restrict__ ibot, float* __restrict__ itop, int fnx, int fny) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i >= fnx) return;
    const int di = i - (buffer - 1), dj = j - (buffer - 1);
    const bool inside = di >= 0 && di < nx && dj >= 0 && dj < ny;
    ibot[(size_t)j * fnx + i] = inside ? zb[(size_t)dj * nx + di] : zb_min;
    itop[(size_t)j * fnx + i] = inside ? zt[(size_t)dj * nx + di] : zt_max;
}

// one sampled height of linear_perturbation_varyingz (:326-333)
static __global__ void __launch_bounds__(kThreads)
k_lw_accumulate(const double2* __restrict__ up, const double2* __restrict__ vp, const float* __restrict__ ibot,
                const float* __restrict__ itop, float current_z, float step_size, double2* __restrict__ ua,
                double2* __restrict__ va, float* __restrict__ count, size_t n) {
    const size_t c = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const float f = lw::layer_fraction(current_z, ibot[c], itop[c], step_size);
    const double fd = (double)f;
    count[c] = count[c] + f;
    ua[c] = make_double2(ua[c].x + up[c].x * fd, ua[c].y + up[c].y * fd);
    va[c] = make_double2(va[c].x + vp[c].x * fd, va[c].y + vp[c].y * fd);
}

// u_perturb = u_accumulator / layer_count   (:338-339)
static __global__ void __launch_bounds__(kThreads)
k_lw_divide_by_count(const double2* __restrict__ ua, const double2* __restrict__ va, const float* __restrict__ count,
                     double2* __restrict__ up, double2* __restrict__ vp, size_t n) {
    const size_t c = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const double d = (double)count[c];
    up[c] = make_double2(ua[c].x / d, ua[c].y / d);
    vp[c] = make_double2(va[c].x / d, va[c].y / d);
}

// level k of u (nx + 1 faces per row): stagger (:756-758) and update (:1094-1098)
static __global__ void __launch_bounds__(kThreads)
k_lw_apply_u(const double2* __restrict__ up, int fnx, int buffer, int nx, int nz, int ldu, int i0, int j0, int k,
             float* __restrict__ u3d, float* __restrict__ upert, float contribution, float fraction) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i > nx) return;
    const size_t row = (size_t)(buffer + j0 + j) * fnx;
    const int col = buffer + i0 + i;
    const float fresh = (float)(up[row + col - 1].x + up[row + col].x) / 2.0f;
    const size_t c = (size_t)i + (size_t)ldu * ((size_t)k + (size_t)nz * (size_t)j);
    const float keep = upert[c] * (1.0f - fraction);
    const float p = keep + fraction * fresh;
    upert[c] = p;
    u3d[c] = u3d[c] + p * contribution;
}

// level k of v (ny + 1 rows): stagger (:760-762) and update (:1107-1111)
static __global__ void __launch_bounds__(kThreads)
k_lw_apply_v(const double2* __restrict__ vp, int fnx, int buffer, int nx, int nz, int ld, int i0, int j0, int k,
             float* __restrict__ v3d, float* __restrict__ vpert, float contribution, float fraction) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const int col = buffer + i0 + i;
    const size_t row = (size_t)(buffer + j0 + j) * fnx;
    const float fresh = (float)(vp[row - fnx + col].x + vp[row + col].x) / 2.0f;
    const size_t c = (size_t)i + (size_t)ld * ((size_t)k + (size_t)nz * (size_t)j);
    const float keep = vpert[c] * (1.0f - fraction);
    const float p = keep + fraction * fresh;
    vpert[c] = p;
    v3d[c] = v3d[c] + p * contribution;
}

}  // namespace lwk
}  // namespace icar
