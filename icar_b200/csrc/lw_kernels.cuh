// The kernels of the linear-theory wind perturbation (see linear_winds.cu for the routine-by-routine map to
// src/physics/linear_winds.f90).  They use nothing but lw_math.cuh, plain pointers and blockIdx / threadIdx, so that
// tests/cpp/lw_kernels_host.cpp can run these very bodies on the CPU (one loop iteration per thread; tests/cpp/host_emu.h
// supplies the built-ins) between NumPy transforms and compare a whole icar_linear_perturb with the restatement without a GPU.
// The product includes this file only from linear_winds.cu.
#pragma once
#include <cstddef>

#include "lw_math.cuh"

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


// ============================================================================================================
// spatial_winds (linear_winds.f90:840-1127): the per-cell look-up in the (speed, direction, log N^2) table

// The table axes (linear_space, array_utilities.f90:215-237; linear_winds.f90:657-661) and the lt_parameters
// spatial_winds reads.  Passed to the kernels by value.
struct LutAxes {
    int n_dir, n_nsq, n_spd;
    float dir[64], nsq[32], spd[32];
};
struct SpatialOpts {
    int variable_N, smooth_nsq, vsmooth, winsz;
    float min_stability, max_stability, N_squared, contribution, fraction;
};
struct CellFields {      // resident fields of the tile, cell layout (pitch ld); hydrometeors may be null
    const float *th, *pii, *z, *qv, *qc, *qi, *qr, *qs;
};

constexpr float kPi = 3.1415927f, kGravity = 9.81f, kRd = 287.058f, kRw = 461.5f, kCp = 1012.0f, kLv = 2260000.0f;

// calc_stability (atm_utilities.f90:394-467), variable_N branch; the argument names are the Fortran's
LW_HD float calc_stability(float th_top, float th_bot, float pii_top, float pii_bot, float z_top, float z_bot,
                           float qv_top, float qv_bot, float qc) {
    if (qc < 1e-7f) return kGravity * (logf(th_top) - logf(th_bot)) / (z_top - z_bot);
    const float t_top = th_top * pii_top, t_bot = th_bot * pii_bot;
    const float t = (t_top + t_bot) / 2.0f, qv = (qv_top + qv_bot) / 2.0f, dz = z_top - z_bot;
    const float L = kLv;
    const float sat_lapse = kGravity * ((1.0f + (L * qv) / (kRd * t)) / (kCp + (L * L * qv * (kRd / kRw)) / (kRd * t * t)));
    return (kGravity / t) * ((t_top - t_bot) / dz + sat_lapse) * (1.0f + (L * qv) / (kRd * t)) -
           (kGravity / (1.0f + qv + qc) * (qv_top - qv_bot) / dz);
}

// calc_direction (atm_utilities.f90:334-355)
LW_HD float calc_direction(float u, float v) {
    if (v < 0.0f) return atanf(u / v) + kPi;
    if (v == 0.0f) return (u > 0.0f) ? kPi / 2.0f : kPi * 1.5f;
    return (u >= 0.0f) ? atanf(u / v) : atanf(u / v) + (2.0f * kPi);
}

// `pos = 1; do step = 1, n; if (cur > values(step)) pos = step` (:1045-1071) + calc_weight (array_utilities.f90:263-288)
LW_HD void table_position(const float* values, int n, float cur, int* pos, int* next, float* weight) {
    int p = 0;
    for (int s = 0; s < n; ++s) if (cur > values[s]) p = s;
    if (cur < values[0]) { *pos = p; *next = 0; *weight = 1.0f; return; }
    if (p == n - 1) { *pos = p; *next = n - 1; *weight = 1.0f; return; }
    *pos = p; *next = p + 1;
    *weight = (values[p + 1] - cur) / (values[p + 1] - values[p]);
}

struct Lookup { int spos, nexts, dpos, nextd, npos, nextn; float sw, dw, nw; };

// mean of log N^2 over levels j-winsz..j+winsz of column (vi, uk) (:1073), then the three table positions
LW_HD Lookup cell_lookup(const LutAxes& ax, float u, float v, const float* nsq_col, int ld, int nz, int j, int winsz) {
    Lookup r;
    table_position(ax.dir, ax.n_dir, calc_direction(u, v), &r.dpos, &r.nextd, &r.dw);
    table_position(ax.spd, ax.n_spd, sqrtf(u * u + v * v), &r.spos, &r.nexts, &r.sw);
    const int bottom = (j - winsz > 0) ? j - winsz : 0, top = (j + winsz < nz - 1) ? j + winsz : nz - 1;
    float acc = 0.0f;
    for (int s = bottom; s <= top; ++s) acc = acc + nsq_col[(size_t)s * ld];
    table_position(ax.nsq, ax.n_nsq, acc / (float)(top - bottom + 1), &r.npos, &r.nextn, &r.nw);
    return r;
}
LW_HD int entry_id(const LutAxes& ax, int spd, int dir, int nsq) { return (dir * ax.n_spd + spd) * ax.n_nsq + nsq; }

// log N^2 of every level of one column, then the in-place vertical running mean (:898-944).  One thread per column.
static __global__ void __launch_bounds__(kThreads)
k_lw_nsq_column(int nx, int nz, int ld, CellFields f, SpatialOpts o, float* __restrict__ nsq) {
    const int i = blockIdx.x * kThreads + threadIdx.x, krow = blockIdx.y;
    if (i >= nx) return;
    const size_t base = (size_t)i + (size_t)ld * (size_t)nz * (size_t)krow;
    for (int j = 1; j <= nz; ++j) {
        float val = o.N_squared;
        if (o.variable_N) {
            const int top = (j + o.vsmooth < nz) ? j + o.vsmooth : nz;
            int bottom = j - (o.vsmooth - (top - j));
            if (bottom < 1) bottom = 1;
            const size_t b = base + (size_t)ld * (bottom - 1), t = base + (size_t)ld * (top - 1), c = base + (size_t)ld * (j - 1);
            float hyd = 0.0f;
            if (f.qc) hyd = hyd + f.qc[c];
            if (f.qi) hyd = hyd + f.qi[c];
            if (f.qr) hyd = hyd + f.qr[c];
            if (f.qs) hyd = hyd + f.qs[c];
            val = calc_stability(f.th[b], f.th[t], f.pii[b], f.pii[t], f.z[b], f.z[t], f.qv[b], f.qv[t], hyd);
            val = fmaxf(o.min_stability, fminf(o.max_stability, val));
        }
        nsq[base + (size_t)ld * (j - 1)] = logf(val);
    }
    if (!o.smooth_nsq) return;
    for (int j = 1; j <= nz; ++j) {                      // in place: the levels below j are already smoothed
        const int top = (j + o.vsmooth < nz) ? j + o.vsmooth : nz;
        int bottom = j - (o.vsmooth - (top - j));
        if (bottom < 1) bottom = 1;
        float acc = nsq[base + (size_t)ld * (j - 1)];
        for (int s = bottom; s <= j - 1; ++s) acc = acc + nsq[base + (size_t)ld * (s - 1)];
        for (int s = j + 1; s <= top; ++s) acc = acc + nsq[base + (size_t)ld * (s - 1)];
        nsq[base + (size_t)ld * (j - 1)] = acc / (float)(top - bottom + 1);
    }
}

// smooth_array_3d(nsquared, winsz, ydim = 3) (array_utilities.f90:290-362): a box mean of 2w+1 along y, then along x,
// with the edge value repeated outside the tile, REAL*8 sums (every window summed afresh, not as running sums)
static __global__ void __launch_bounds__(kThreads)
k_lw_smooth_y(const float* __restrict__ in, double* __restrict__ out, int nx, int nz, int ny, int ld, int w) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    double s = 0.0;
    for (int d = -w; d <= w; ++d) {
        int kk = k + d;
        kk = kk < 0 ? 0 : (kk > ny - 1 ? ny - 1 : kk);
        s += (double)in[(size_t)i + (size_t)ld * ((size_t)j + (size_t)nz * (size_t)kk)];
    }
    out[(size_t)i + (size_t)ld * ((size_t)j + (size_t)nz * (size_t)k)] = s / (double)(2 * w + 1);
}
static __global__ void __launch_bounds__(kThreads)
k_lw_smooth_x(const double* __restrict__ in, float* __restrict__ out, int nx, int nz, int ny, int ld, int w) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const size_t row = (size_t)ld * ((size_t)j + (size_t)nz * (size_t)k);
    double s = 0.0;
    for (int d = -w; d <= w; ++d) {
        int ii = i + d;
        ii = ii < 0 ? 0 : (ii > nx - 1 ? nx - 1 : ii);
        s += in[row + ii];
    }
    out[row + i] = (float)(s / (double)(2 * w + 1));
}

// u1d(i) = sum(u3d(i, :, uk)) / nz, v1d(i) = sum(v3d(vi, :, k)) / nz for rows k = row0 .. row1 (:965-973); i = 0 .. nx
static __global__ void __launch_bounds__(kThreads)
k_lw_colmean(const float* __restrict__ u3d, const float* __restrict__ v3d, int nx, int nz, int ny, int ld, int ldu,
             int row0, float* __restrict__ u1d, float* __restrict__ v1d) {
    const int i = blockIdx.x * kThreads + threadIdx.x, k = row0 + blockIdx.y;
    if (i > nx) return;
    const int uk = k < ny ? k : ny - 1, vi = i < nx ? i : nx - 1;
    float su = 0.0f, sv = 0.0f;
    for (int j = 0; j < nz; ++j) {
        su = su + u3d[(size_t)i + (size_t)ldu * ((size_t)j + (size_t)nz * (size_t)uk)];
        sv = sv + v3d[(size_t)vi + (size_t)ld * ((size_t)j + (size_t)nz * (size_t)k)];
    }
    u1d[(size_t)k * (nx + 1) + i] = su / (float)nz;
    v1d[(size_t)k * (nx + 1) + i] = sv / (float)nz;
}

// which table entries level j needs for rows row0 .. (one bit per entry)
static __global__ void __launch_bounds__(kThreads)
k_lw_mark_entries(LutAxes ax, const float* __restrict__ u1d, const float* __restrict__ v1d, const float* __restrict__ nsq,
                  int nx, int nz, int ny, int ld, int j, int winsz, int row0, unsigned* __restrict__ bitmap) {
    const int i = blockIdx.x * kThreads + threadIdx.x, k = row0 + blockIdx.y;
    if (i > nx) return;
    const int uk = k < ny ? k : ny - 1, vi = i < nx ? i : nx - 1;
    const float* col = nsq + (size_t)vi + (size_t)ld * (size_t)nz * (size_t)uk;
    const Lookup r = cell_lookup(ax, u1d[(size_t)k * (nx + 1) + i], v1d[(size_t)k * (nx + 1) + i], col, ld, nz, j, winsz);
    const int sp[2] = {r.spos, r.nexts}, dp[2] = {r.dpos, r.nextd}, np[2] = {r.npos, r.nextn};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                const int e = entry_id(ax, sp[a], dp[b], np[c]);
                atomicOr(bitmap + (e >> 5), 1u << (e & 31));
            }
}

// one table entry of one level, staggered (:756-762), on the GLOBAL domain: plane_u (gnx + 1, gny), plane_v (gnx, gny + 1)
static __global__ void __launch_bounds__(kThreads)
k_lw_stagger(const double2* __restrict__ up, const double2* __restrict__ vp, int fnx, int buffer, int gnx, int gny,
             float* __restrict__ plane_u, float* __restrict__ plane_v) {
    const int i = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
    if (i > gnx) return;
    const size_t row = (size_t)(buffer + j) * fnx;
    const int col = buffer + i;
    if (j < gny) plane_u[(size_t)j * (gnx + 1) + i] = (float)(up[row + col - 1].x + up[row + col].x) / 2.0f;
    if (i < gnx) plane_v[(size_t)j * gnx + i] = (float)(vp[row - fnx + col].x + vp[row + col].x) / 2.0f;
}

// the interpolation between the eight entries around a cell (:1084-1091, :1100-1107), relaxation and update of u or v
LW_HD float interpolate8(const Lookup& r, const LutAxes& ax, const int* slot_of, const float* planes, size_t stride, size_t at) {
#define LW_LUT(s, d, n) planes[(size_t)slot_of[entry_id(ax, s, d, n)] * stride + at]
    const float first = r.nw * (r.dw * LW_LUT(r.spos, r.dpos, r.npos) + (1.0f - r.dw) * LW_LUT(r.spos, r.nextd, r.npos)) +
                        (1.0f - r.nw) * (r.dw * LW_LUT(r.spos, r.dpos, r.nextn) + (1.0f - r.dw) * LW_LUT(r.spos, r.nextd, r.nextn));
    const float second = r.nw * (r.dw * LW_LUT(r.nexts, r.dpos, r.npos) + (1.0f - r.dw) * LW_LUT(r.nexts, r.nextd, r.npos)) +
                         (1.0f - r.nw) * (r.dw * LW_LUT(r.nexts, r.dpos, r.nextn) + (1.0f - r.dw) * LW_LUT(r.nexts, r.nextd, r.nextn));
#undef LW_LUT
    return r.sw * first + (1.0f - r.sw) * second;
}

// level j of u (rows 0 .. ny-1) or of v (rows row0 .. row0 + gridDim.y - 1); is_v selects
static __global__ void __launch_bounds__(kThreads)
k_lw_interp(LutAxes ax, SpatialOpts o, const float* __restrict__ u1d, const float* __restrict__ v1d,
            const float* __restrict__ nsq, int nx, int nz, int ny, int ld, int ldw, int j, int row0, int is_v,
            const int* __restrict__ slot_of, const float* __restrict__ planes, size_t stride, int gnx, int i0, int j0,
            float* __restrict__ wind, float* __restrict__ pert) {
    const int i = blockIdx.x * kThreads + threadIdx.x, k = row0 + blockIdx.y;
    if (is_v ? (i >= nx) : (i > nx)) return;
    const int uk = k < ny ? k : ny - 1, vi = i < nx ? i : nx - 1;
    const float* col = nsq + (size_t)vi + (size_t)ld * (size_t)nz * (size_t)uk;
    const Lookup r = cell_lookup(ax, u1d[(size_t)k * (nx + 1) + i], v1d[(size_t)k * (nx + 1) + i], col, ld, nz, j, o.winsz);
    const size_t at = is_v ? (size_t)(j0 + k) * gnx + (i0 + i) : (size_t)(j0 + k) * (gnx + 1) + (i0 + i);
    const float fresh = interpolate8(r, ax, slot_of, planes, stride, at);
    const size_t c = (size_t)i + (size_t)ldw * ((size_t)j + (size_t)nz * (size_t)k);
    const float keep = pert[c] * (1.0f - o.fraction);
    const float p = keep + o.fraction * fresh;
    pert[c] = p;
    wind[c] = wind[c] + p * o.contribution;
}

}  // namespace lwk
}  // namespace icar
