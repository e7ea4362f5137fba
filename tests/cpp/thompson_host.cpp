// TEST INFRASTRUCTURE.  Runs k_mp_thompson (icar_b200/csrc/mp_thompson_kernel.inc) on the CPU with ONE OS THREAD PER
// CUDA THREAD of a CTA: shared memory is a heap block, __syncthreads a pthread barrier, the shared-memory atomicMax
// and the diagnostics atomicAdd host atomics (host_emu.h), powf / expf / log10f / pow ... the host's libm.  The launch
// is th_driver_launch's (mp_thompson.cu): block (TX, nkt), grid (ceil(width / TX), rows, tiles), the same dynamic
// shared memory.  The constants come from the product's own host routine icar::th_make_constants (linked from
// libicar_b200.so); the lookup tables are passed in by the test (the oracle's, or any others).
#include <pthread.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "host_emu.h"
#include "../../icar_b200/csrc/common.cuh"
#include "../../icar_b200/csrc/mp_thompson.cuh"

thread_local icar_emu_idx blockIdx, threadIdx, blockDim;
static unsigned char* icar_emu_smem = nullptr;
static pthread_barrier_t cta_barrier;
static std::vector<pthread_barrier_t> warp_barrier;
static std::vector<int> warp_buffer;                                // [warp][32]
static int emu_tx = 8;                                              // blockDim.x (TX of the kernel)
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
#define ICAR_DYNAMIC_SMEM_D(name) double* name = reinterpret_cast<double*>(icar_emu_smem)

namespace icar {
namespace {
#include "../../icar_b200/csrc/mp_thompson_kernel.inc"
}  // namespace
}  // namespace icar

using namespace icar;

extern "C" int emu_thompson_driver(int nx, int nz, int ny, int ld, int ld2, const ThOptions* opt, const double* const* tables,
                                   float* qv, float* qc, float* qr, float* qi, float* qs, float* qg, float* ni, float* nr,
                                   float* th, const float* pii, const float* p, const float* dz, float dt, float* rainnc,
                                   float* snownc, float* graupelnc, float* sr, int its, int ite, int jts, int jte, int kts,
                                   int kte, int along_y, unsigned long long* diag) {
    Geom g{nx, nz, ny, ld, ld};
    ThConst c;
    std::vector<double> bins(1000);
    th_make_constants(*opt, &c, bins.data());
    ThTables tb;
    for (int t = 0; t < TH_NTABLES; ++t) tb.t[t] = tables[t];
    MpTiles tiles{};
    tiles.n = 1;
    tiles.its[0] = its; tiles.ite[0] = ite; tiles.jts[0] = jts; tiles.jte[0] = jte; tiles.along_y[0] = along_y;
    const int w = ite - its + 1, h = jte - jts + 1;
    const int nk = kte - kts + 1;
    const int TX = th_columns_per_cta(nk);
    const int gx = ((along_y ? h : w) + TX - 1) / TX, gy = along_y ? w : h;
    if (nk < 1 || nk * TX > 1024) return -1;
    const int nkt = (nk * TX + 31) / 32 * 32 / TX;
    Args A{qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz, rainnc, snownc, nullptr, graupelnc, nullptr, sr, nullptr,
           nullptr, nullptr, nullptr};
    const size_t smem = (size_t)nk * TX * sizeof(double) + (size_t)7 * (nk + 1) * TX * sizeof(float) + (size_t)10 * TX * sizeof(int);
    const int nthreads = TX * nkt;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::vector<unsigned char> block(smem + 64, 0xA5);
            icar_emu_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(block.data()) + 63) & ~uintptr_t(63));
            pthread_barrier_init(&cta_barrier, nullptr, nthreads);
            emu_tx = TX;
            warp_barrier.assign(nthreads / 32, pthread_barrier_t{});
            for (auto& b : warp_barrier) pthread_barrier_init(&b, nullptr, 32);
            warp_buffer.assign(nthreads, 0);
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t)
                pool.emplace_back([=, &c, &tb, &A, &tiles, &g] {
                    blockIdx = {(unsigned)bx, (unsigned)by, 0};
                    threadIdx = {(unsigned)(t % TX), (unsigned)(t / TX), 0};
                    blockDim = {(unsigned)TX, (unsigned)nkt, 1};
                    if (TX == 16) {
                        if (nthreads <= 320) k_mp_thompson<16, 320, 2>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                        else if (nthreads <= 640) k_mp_thompson<16, 640, 1>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                        else if (nthreads <= 800) k_mp_thompson<16, 800, 1>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                        else k_mp_thompson<16, 1024, 1>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                    } else if (nthreads <= 640) k_mp_thompson<8, 640, 1>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                    else k_mp_thompson<8, 1024, 1>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag);
                });
            for (auto& th_ : pool) th_.join();
            pthread_barrier_destroy(&cta_barrier);
            for (auto& b : warp_barrier) pthread_barrier_destroy(&b);
        }
    return 1;
}
