// TEST INFRASTRUCTURE.  Runs the kernel bodies of icar_b200/csrc/lw_kernels.cuh on the CPU: every "launch" is a loop
// over the grid the product uses in linear_winds.cu (128 threads per block, blockIdx.y = row), one iteration per
// thread.  tests/test_linear_winds_host.py puts NumPy transforms where the product calls cuFFT.
#include "../../icar_b200/csrc/lw_kernels.cuh"

thread_local lw_emu_idx blockIdx, threadIdx;

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

}
