// TEST INFRASTRUCTURE.  Runs the kernels of the fused MPDATA path (icar_b200/csrc/corrector_fused.inc:
// k_face_coefficients, k_corrector_fused) on the CPU.
//
// A CTA is emulated with ONE OS THREAD PER CUDA THREAD: __syncthreads is a pthread barrier over the CTA, a warp
// shuffle is an exchange through a per-warp buffer between two barriers over the warp, an mbarrier is a phase
// counter + outstanding byte count in the shared-memory word the kernel gave it, and a TMA box load is a copy with
// zero fill outside the field done by the issuing thread (the kernel's waiters spin on the phase like on hardware).
// Shared memory is one heap block per CTA; CTAs run one after the other.  The f32x2 operations are done one IEEE half
// at a time (f32x2.cuh) and the packed quotient is the IEEE `/` (corrector_prims.inc: the device sequence is the fast
// path of div.rn.f32, checked against `/` on the device itself), so the results are the device's bit for bit.
// Built with -ffp-contract=off -pthread.
#include <pthread.h>
#include <sched.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "host_emu.h"
#include "../../icar_b200/csrc/common.cuh"
#include "../../icar_b200/csrc/f32x2.cuh"
#include "../../icar_b200/csrc/mpdata.cuh"

thread_local icar_emu_idx blockIdx, threadIdx, blockDim;

// ---- what the kernels need from the "hardware"
struct CUtensorMap { const float* base; int nx, nz, ny, ld; };     // the kernels only pass its address on
static unsigned char* icar_emu_smem = nullptr;
static pthread_barrier_t cta_barrier;
static std::vector<pthread_barrier_t> warp_barrier;
static std::vector<float> warp_buffer;                              // [warp][32]

static inline void __syncthreads() { pthread_barrier_wait(&cta_barrier); }
static inline float icar_emu_shfl(float v, int delta) {            // delta = -1: value of lane - 1 (shfl.up), +1: lane + 1
    const int tid = threadIdx.y * 32 + threadIdx.x, w = tid / 32, lane = tid % 32;
    warp_buffer[w * 32 + lane] = v;
    pthread_barrier_wait(&warp_barrier[w]);
    const int src = lane + delta;
    const float r = (src < 0 || src > 31) ? v : warp_buffer[w * 32 + src];
    pthread_barrier_wait(&warp_barrier[w]);
    return r;
}
// an mbarrier: phase counter + what is still outstanding in the current phase (arrivals and TMA bytes); the arrival
// count it was initialised with lives beside it (the word itself has no room), written once before the CTA starts
// a waiter that finds nothing sleeps a little instead of spinning: hundreds of OS threads yielding in a loop spend
// their time in the scheduler
static inline void emu_pause() { timespec t{0, 20000}; nanosleep(&t, nullptr); }
struct EmuMbar { std::atomic<uint32_t> phase; std::atomic<int32_t> pending; };
static_assert(sizeof(EmuMbar) == sizeof(uint64_t), "an mbarrier is one 64-bit shared-memory word");
static int32_t emu_mbar_count[256];
static inline int32_t& emu_count_of(uint64_t* b) { return emu_mbar_count[(reinterpret_cast<uintptr_t>(b) >> 3) & 255]; }
static inline void emu_mbar_complete_one(uint64_t* b, int32_t units) {
    auto* m = reinterpret_cast<EmuMbar*>(b);
    if (m->pending.fetch_sub(units) - units == 0) {
        m->pending.store(emu_count_of(b));
        m->phase.fetch_add(1, std::memory_order_release);
    }
}
static inline void icar_emu_mbar_init(uint64_t* b, uint32_t count) {
    auto* m = reinterpret_cast<EmuMbar*>(b);
    emu_count_of(b) = (int32_t)count;
    m->phase.store(0);
    m->pending.store((int32_t)count);
}
static inline void icar_emu_mbar_arrive(uint64_t* b) { emu_mbar_complete_one(b, 1); }
static inline void icar_emu_mbar_expect(uint64_t* b, uint32_t bytes) {      // arrive.expect_tx: one arrival, `bytes` more to come
    reinterpret_cast<EmuMbar*>(b)->pending.fetch_add((int32_t)bytes - 1);
}
static inline bool icar_emu_mbar_try_wait(uint64_t* b, uint32_t parity) {
    const bool ok = (reinterpret_cast<EmuMbar*>(b)->phase.load(std::memory_order_acquire) & 1u) != parity;
    if (!ok) emu_pause();
    return ok;
}
static inline void icar_emu_yield() { emu_pause(); }
static inline void icar_emu_syncwarp() {
    const int w = (threadIdx.y * 32 + threadIdx.x) / 32;
    pthread_barrier_wait(&warp_barrier[w]);
}
static inline void icar_emu_tma_load_box(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int z, int y) {
    float* d = static_cast<float*>(dst);
    for (int zz = 0; zz < m->nz; ++zz)
        for (int xx = 0; xx < 32; ++xx) {
            const int gi = x + xx, gk = z + zz;
            const bool in = gi >= 0 && gi < m->nx && gk >= 0 && gk < m->nz && y >= 0 && y < m->ny;
            d[zz * 32 + xx] = in ? m->base[(size_t)gi + (size_t)m->ld * ((size_t)gk + (size_t)m->nz * (size_t)y)] : 0.f;
        }
    emu_mbar_complete_one(bar, m->nz * 32 * (int32_t)sizeof(float));
}
using std::max;
using std::min;
#define ICAR_DYNAMIC_SMEM(name) unsigned char* name = icar_emu_smem

namespace icar {
namespace {
// the hardware primitives of corrector_prims.inc, emulated (one OS thread per CUDA thread of a CTA)
inline float lane_lo(float v) { return icar_emu_shfl(v, -1); }
inline float lane_hi(float v) { return icar_emu_shfl(v, +1); }
inline void mbar_init(uint64_t* b, uint32_t count) { icar_emu_mbar_init(b, count); }
inline void fence_barrier_init() {}
inline void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) { icar_emu_mbar_expect(b, bytes); }
inline bool mbar_try_wait(uint64_t* b, uint32_t parity) { return icar_emu_mbar_try_wait(b, parity); }
inline void tma_load_box(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int z, int y) {
    icar_emu_tma_load_box(dst, map, bar, x, z, y);
}
#include "../../icar_b200/csrc/corrector_prims.inc"
#include "../../icar_b200/csrc/corrector_fused.inc"

// one CTA of `threads` = 32 x nzt CUDA threads
template <class Kernel>
void run_cta(unsigned bx, unsigned by, int nzt, size_t smem_bytes, Kernel kernel) {
    const int threads = 32 * nzt;
    std::vector<unsigned char> smem(smem_bytes + 128, 0xA5);          // shared memory starts uninitialised
    icar_emu_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    pthread_barrier_init(&cta_barrier, nullptr, threads);
    warp_barrier.assign(nzt, pthread_barrier_t{});
    for (auto& b : warp_barrier) pthread_barrier_init(&b, nullptr, 32);
    warp_buffer.assign((size_t)nzt * 32, 0.f);
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] {
            blockIdx = {bx, by, 0};
            threadIdx = {(unsigned)(t % 32), (unsigned)(t / 32), 0};
            blockDim = {32, (unsigned)nzt, 1};
            kernel();
        });
    for (auto& th : pool) th.join();
    for (auto& b : warp_barrier) pthread_barrier_destroy(&b);
    pthread_barrier_destroy(&cta_barrier);
}

template <class Kernel>
void run_grid(unsigned gx, unsigned gy, int nzt, size_t smem, Kernel kernel) {
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) run_cta(bx, by, nzt, smem, kernel);
}

}  // namespace
}  // namespace icar

using namespace icar;

extern "C" {

// k_face_coefficients (grid ((ld/4+127)/128, nz, ny) x 128); coef10 in FusedCoef's member order
void emu_face_coefficients(int nx, int nz, int ny, int ld, const float* U, const float* V, const float* Wn,
                           const float* G, float* const* coef10) {
    Geom g{nx, nz, ny, ld, ld};
    FusedCoef co{coef10[0], coef10[1], coef10[2], coef10[3], coef10[4], coef10[5], coef10[6], coef10[7], coef10[8], coef10[9]};
    blockDim = {128, 1, 1};
    for (unsigned j = 0; j < (unsigned)ny; ++j)
        for (unsigned k = 0; k < (unsigned)nz; ++k)
            for (unsigned bx = 0; bx < (unsigned)((ld / 4 + 127) / 128); ++bx)
                for (unsigned t = 0; t < 128; ++t) { blockIdx = {bx, k, j}; threadIdx = {t, 0, 0}; k_face_coefficients(g, U, V, Wn, G, co); }
}

// the corrector as fused_advect_all launches it: grid (xtiles * n, chunks), block (32, nzp / (2 np)), the dynamic shared
// memory of fused_plan.  Returns 2 when the instantiation needs more shared memory than a CTA has, 1 when there is no
// such instantiation.
int emu_corrector(int np, int nzp, int stages, int nx, int nz, int ny, int ld, int n, const float* const* q2,
                  const float* const* l, float* const* out, const float* const* coef10, const float* jaco, const float* dz,
                  int chunk) {
    Geom g{nx, nz, ny, ld, ld};
    FusedMaps maps{};
    OutSet outs{};
    outs.n = n;
    for (int s = 0; s < n; ++s) {
        outs.out[s] = out[s];
        maps.q2[s] = CUtensorMap{q2[s], nx, nz, ny, ld};
        maps.l[s] = CUtensorMap{l[s], nx, nz, ny, ld};
    }
    // coef10 is in FusedCoef's member order (t0x cbx ccx t0y cby ccy ggy t0z cbz ccz); the maps are in stage order
    // (t0y cby ccy ggy | t0x cbx ccx t0z cbz ccz jaco dz), as fused_advect_all builds them
    const int order[10] = {3, 4, 5, 6, 0, 1, 2, 7, 8, 9};
    for (int f = 0; f < 10; ++f) maps.coef[f] = CUtensorMap{coef10[order[f]], nx, nz, ny, ld};
    maps.coef[10] = CUtensorMap{jaco, nx, nz, ny, ld};
    maps.coef[11] = CUtensorMap{dz, nx, nz, ny, ld};
    const int nzt = nzp / (2 * np);
    const size_t smem = (size_t)(8 * (nzt + 1) * TW + stages * NFE * nzp * TW) * sizeof(float) + 16;
    if (smem > 232448) return 2;                                      // 227 KB of shared memory per CTA
    const int xtiles = nx / TOUT + 1, planes = ny - 2;
    const unsigned gx = (unsigned)(xtiles * n), gy = (unsigned)((planes + chunk - 1) / chunk);
#define CASE(N, Z, S) if (np == N && nzp == Z && stages == S) { run_grid(gx, gy, nzt, smem, [&] { k_corrector_fused<N, Z, S>(g, outs, maps, chunk, OutRange{1, nx - 1, 1, ny - 1}); }); return 0; }
    CASE(1, 20, 2)
    CASE(1, 40, 2)
    CASE(2, 48, 2)
    CASE(1, 52, 2)
    CASE(1, 56, 2)
    CASE(2, 60, 2)
    CASE(2, 64, 1)
    CASE(2, 64, 2)          // does not fit: exercises the refusal
#undef CASE
    return 1;
}

}
