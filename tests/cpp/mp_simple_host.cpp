// TEST INFRASTRUCTURE.  Runs k_mp_simple (icar_b200/csrc/mp_simple_kernel.inc) on the CPU with ONE OS THREAD PER CUDA
// THREAD of a CTA: shared memory is a heap block, __syncthreads a pthread barrier, the shared-memory atomicMax and the
// diagnostics atomicAdd host atomics (host_emu.h).  The launch is mp_simple_launch's (mp_simple.cu): block (TX, nzt),
// grid (ceil(width / TX), rows, tiles), the same dynamic shared memory.
#include <pthread.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "host_emu.h"
#include "../../icar_b200/csrc/common.cuh"
#include "../../icar_b200/csrc/mp_simple.cuh"

thread_local icar_emu_idx blockIdx, threadIdx, blockDim;
static unsigned char* icar_emu_smem = nullptr;
static pthread_barrier_t cta_barrier;
static std::vector<pthread_barrier_t> warp_barrier;
static std::vector<int> warp_buffer;                                // [warp][32]
static int emu_tx = 32;
static inline void __syncthreads() { pthread_barrier_wait(&cta_barrier); }
static inline int __reduce_max_sync(unsigned, int v) {             // redux.sync.max over the 32 lanes of the caller's warp
    const int tid = threadIdx.y * emu_tx + threadIdx.x, w = tid / 32, lane = tid % 32;
    warp_buffer[w * 32 + lane] = v;
    pthread_barrier_wait(&warp_barrier[w]);
    int m = warp_buffer[w * 32];
    for (int l = 1; l < 32; ++l) m = std::max(m, warp_buffer[w * 32 + l]);
    pthread_barrier_wait(&warp_barrier[w]);
    return m;
}
using std::max;
using std::min;
#define ICAR_DYNAMIC_SMEM_F(name) float* name = reinterpret_cast<float*>(icar_emu_smem)

namespace icar {
namespace {
#include "../../icar_b200/csrc/mp_simple_kernel.inc"
}  // namespace
}  // namespace icar

using namespace icar;

extern "C" int emu_mp_simple(int nx, int nz, int ny, int ld, int ld2, const float* p, float* th, const float* pii,
                             const float* rho, float* qv, float* qc, float* qr, float* qs, float* rain, float* snow, float dt,
                             const float* dz, int its, int ite, int jts, int jte, int kts, int kte, int along_y, float c2r,
                             float c2s, int cpt, unsigned long long* diag) {
    Geom g{nx, nz, ny, ld, ld};
    MpTiles tiles{};
    tiles.n = 1;
    tiles.its[0] = its; tiles.ite[0] = ite; tiles.jts[0] = jts; tiles.jte[0] = jte; tiles.along_y[0] = along_y;
    const int w = ite - its + 1, h = jte - jts + 1;
    const int gx = ((along_y ? h : w) + TX - 1) / TX, gy = along_y ? w : h;
    if (cpt <= 0) cpt = (nz + 9) / 10;
    if (cpt > 8) return -1;
    const int nzt = (nz + cpt - 1) / cpt;
    const size_t smem = (size_t)(2 * (nz + 1) + 3) * TX * sizeof(float);
    const int nthreads = TX * nzt;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::vector<unsigned char> block(smem + 64, 0xA5);
            icar_emu_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(block.data()) + 63) & ~uintptr_t(63));
            pthread_barrier_init(&cta_barrier, nullptr, nthreads);
            emu_tx = TX;
            warp_barrier.assign((nthreads + 31) / 32, pthread_barrier_t{});
            for (auto& b : warp_barrier) pthread_barrier_init(&b, nullptr, 32);
            warp_buffer.assign((size_t)warp_barrier.size() * 32, 0);
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t)
                pool.emplace_back([=, &tiles, &g] {
                    blockIdx = {(unsigned)bx, (unsigned)by, 0};
                    threadIdx = {(unsigned)(t % TX), (unsigned)(t / TX), 0};
                    blockDim = {(unsigned)TX, (unsigned)nzt, 1};
#define RUN(C) k_mp_simple<C>(g, p, th, pii, rho, qv, qc, qr, qs, rain, snow, ld2, dt, dz, tiles, kts, kte, c2r, c2s, diag, nullptr, nullptr)
                    switch (cpt) {
                        case 1: RUN(1); break; case 2: RUN(2); break; case 3: RUN(3); break; case 4: RUN(4); break;
                        case 5: RUN(5); break; case 6: RUN(6); break; case 7: RUN(7); break; default: RUN(8); break;
                    }
#undef RUN
                });
            for (auto& th_ : pool) th_.join();
            pthread_barrier_destroy(&cta_barrier);
            for (auto& b : warp_barrier) pthread_barrier_destroy(&b);
        }
    return 1;
}
