// TEST INFRASTRUCTURE.  Runs the kernel bodies of icar_b200/csrc/lw_kernels.cuh on the CPU: every "launch" is a loop
// over the grid the product uses in linear_winds.cu (128 threads per block, blockIdx.y = row), one iteration per
// thread.  tests/test_linear_winds_host.py puts NumPy transforms where the product calls cuFFT.
#include "host_emu.h"
#include "../../icar_b200/csrc/lw_kernels.cuh"

thread_local icar_emu_idx blockIdx, threadIdx, blockDim;

using namespace icar::lwk;

namespace {
template <class F>
void launch(unsigned gx, unsigned gy, F body) {
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx)
            for (unsigned t = 0; t < (unsigned)kThreads; ++t) {
                blockIdx = {bx, by, 0};
                threadIdx = {t, 0, 0};
                body();
            }
}
unsigned blocks(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }
}  // namespace

extern "C" {

void emu_shift_norm(const double2* in, double2* out, int fnx, int fny) {
    launch(blocks(fnx), fny, [&] { k_lw_shift_norm(in, out, fnx, fny, (double)(fnx * fny)); });
}
void emu_spectrum(const double2* tf, const float* kx, const float* ly, int fnx, int fny, float U, float V, float Nsq,
                  float z0, float step, int nsteps, double2* uhat, double2* vhat) {
    launch(blocks(fnx), fny, [&] { k_lw_spectrum(tf, kx, ly, fnx, fny, U, V, Nsq, z0, step, nsteps, uhat, vhat); });
}
void emu_divide(double2* u, double2* v, long long n, double by) {
    launch(blocks(n), 1, [&] { k_lw_divide(u, v, (size_t)n, by); });
}
void emu_layer_bounds(const float* zb, const float* zt, int nx, int ny, int buffer, float zb_min, float zt_max,
                      float* ibot, float* itop, int fnx, int fny) {
    launch(blocks(fnx), fny, [&] { k_lw_layer_bounds(zb, zt, nx, ny, buffer, zb_min, zt_max, ibot, itop, fnx, fny); });
}
void emu_accumulate(const double2* up, const double2* vp, const float* ibot, const float* itop, float current_z,
                    float step_size, double2* ua, double2* va, float* count, long long n) {
    launch(blocks(n), 1, [&] { k_lw_accumulate(up, vp, ibot, itop, current_z, step_size, ua, va, count, (size_t)n); });
}
void emu_divide_by_count(const double2* ua, const double2* va, const float* count, double2* up, double2* vp, long long n) {
    launch(blocks(n), 1, [&] { k_lw_divide_by_count(ua, va, count, up, vp, (size_t)n); });
}
void emu_apply_u(const double2* up, int fnx, int buffer, int nx, int ny, int nz, int ldu, int i0, int j0, int k,
                 float* u3d, float* upert, float contribution, float fraction) {
    launch(blocks(nx + 1), ny, [&] { k_lw_apply_u(up, fnx, buffer, nx, nz, ldu, i0, j0, k, u3d, upert, contribution, fraction); });
}
void emu_apply_v(const double2* vp, int fnx, int buffer, int nx, int ny, int nz, int ld, int i0, int j0, int k,
                 float* v3d, float* vpert, float contribution, float fraction) {
    launch(blocks(nx), ny + 1, [&] { k_lw_apply_v(vp, fnx, buffer, nx, nz, ld, i0, j0, k, v3d, vpert, contribution, fraction); });
}


// ---- spatial_winds kernels.  axes: n_dir, n_nsq, n_spd, then the three axes; opts: SpatialOpts as 4 ints + 5 floats
static LutAxes axes_of(const int* n, const float* dir, const float* nsq, const float* spd) {
    LutAxes ax{};
    ax.n_dir = n[0]; ax.n_nsq = n[1]; ax.n_spd = n[2];
    for (int i = 0; i < n[0]; ++i) ax.dir[i] = dir[i];
    for (int i = 0; i < n[1]; ++i) ax.nsq[i] = nsq[i];
    for (int i = 0; i < n[2]; ++i) ax.spd[i] = spd[i];
    return ax;
}
static SpatialOpts opts_of(const int* oi, const float* of) {
    return SpatialOpts{oi[0], oi[1], oi[2], oi[3], of[0], of[1], of[2], of[3], of[4]};
}

void emu_nsq_column(int nx, int nz, int ny, int ld, const float* th, const float* pii, const float* z, const float* qv,
                    const float* qc, const float* qi, const float* qr, const float* qs, const int* oi, const float* of,
                    float* nsq) {
    const CellFields f{th, pii, z, qv, qc, qi, qr, qs};
    const SpatialOpts o = opts_of(oi, of);
    launch(blocks(nx), ny, [&] { k_lw_nsq_column(nx, nz, ld, f, o, nsq); });
}
void emu_smooth(float* nsq, double* rowmeans, int nx, int nz, int ny, int ld, int w) {
    auto launch3 = [&](auto body) {
        for (unsigned k = 0; k < (unsigned)ny; ++k)
            for (unsigned j = 0; j < (unsigned)nz; ++j)
                for (unsigned bx = 0; bx < blocks(nx); ++bx)
                    for (unsigned t = 0; t < (unsigned)kThreads; ++t) { blockIdx = {bx, j, k}; threadIdx = {t, 0, 0}; body(); }
    };
    launch3([&] { k_lw_smooth_y(nsq, rowmeans, nx, nz, ny, ld, w); });
    launch3([&] { k_lw_smooth_x(rowmeans, nsq, nx, nz, ny, ld, w); });
}
void emu_colmean(const float* u3d, const float* v3d, int nx, int nz, int ny, int ld, int ldu, int row0, int nrows,
                 float* u1d, float* v1d) {
    launch(blocks(nx + 1), nrows, [&] { k_lw_colmean(u3d, v3d, nx, nz, ny, ld, ldu, row0, u1d, v1d); });
}
void emu_mark_entries(const int* n, const float* dir, const float* nsqv, const float* spd, const float* u1d, const float* v1d,
                      const float* nsq, int nx, int nz, int ny, int ld, int j, int winsz, int row0, int nrows, unsigned* bitmap) {
    const LutAxes ax = axes_of(n, dir, nsqv, spd);
    launch(blocks(nx + 1), nrows, [&] { k_lw_mark_entries(ax, u1d, v1d, nsq, nx, nz, ny, ld, j, winsz, row0, bitmap); });
}
void emu_stagger(const double2* up, const double2* vp, int fnx, int buffer, int gnx, int gny, float* plane_u, float* plane_v) {
    launch(blocks(gnx + 1), gny + 1, [&] { k_lw_stagger(up, vp, fnx, buffer, gnx, gny, plane_u, plane_v); });
}
void emu_interp(const int* n, const float* dir, const float* nsqv, const float* spd, const int* oi, const float* of,
                const float* u1d, const float* v1d, const float* nsq, int nx, int nz, int ny, int ld, int ldw, int j, int row0,
                int nrows, int is_v, const int* slot_of, const float* planes, long long stride, int gnx, int i0, int j0,
                float* wind, float* pert) {
    const LutAxes ax = axes_of(n, dir, nsqv, spd);
    const SpatialOpts o = opts_of(oi, of);
    launch(blocks(is_v ? nx : nx + 1), nrows, [&] {
        k_lw_interp(ax, o, u1d, v1d, nsq, nx, nz, ny, ld, ldw, j, row0, is_v, slot_of, planes, (size_t)stride, gnx, i0, j0, wind, pert);
    });
}

}
