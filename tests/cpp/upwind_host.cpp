// TEST INFRASTRUCTURE.  Runs k_upwind_all_v4 (icar_b200/csrc/upwind_all.cuh: MPDATA pass 1 for all scalars, four cells
// per thread) on the CPU, one loop iteration per thread over the grid staged_upwind_all launches:
// dim3((ld / 4 + 127) / 128, nz, ny) blocks of 128 threads.  Built with -ffp-contract=off.
#include "host_emu.h"
#include "../../icar_b200/csrc/upwind_all.cuh"

thread_local icar_emu_idx blockIdx, threadIdx, blockDim;

using namespace icar;

template <int N>
static void run(const Geom& g, const ScalarSet& s, const float* U, const float* V, const float* W, const float* dz,
                const float* jaco) {
    const unsigned gx = (unsigned)((g.ld / 4 + 127) / 128);
    blockDim = {128, 1, 1};
    for (unsigned j = 0; j < (unsigned)g.ny; ++j)
        for (unsigned k = 0; k < (unsigned)g.nz; ++k)
            for (unsigned bx = 0; bx < gx; ++bx)
                for (unsigned t = 0; t < 128; ++t) {
                    blockIdx = {bx, k, j};
                    threadIdx = {t, 0, 0};
                    k_upwind_all_v4<N>(g, s, U, V, W, dz, jaco);
                }
}

extern "C" int emu_upwind_all(int nx, int nz, int ny, int ld, int n, const float* const* q, float* const* q2,
                              float* const* out, const float* U, const float* V, const float* W, const float* dz,
                              const float* jaco) {
    Geom g{nx, nz, ny, ld, ld};
    ScalarSet s{};
    s.n = n;
    for (int i = 0; i < n; ++i) { s.q[i] = q[i]; s.q2[i] = q2[i]; s.out[i] = out[i]; }
    switch (n) {
        case 5: run<5>(g, s, U, V, W, dz, jaco); break;
        case 4: run<4>(g, s, U, V, W, dz, jaco); break;
        case 3: run<3>(g, s, U, V, W, dz, jaco); break;
        case 2: run<2>(g, s, U, V, W, dz, jaco); break;
        case 1: run<1>(g, s, U, V, W, dz, jaco); break;
        default: return 1;
    }
    return 0;
}
