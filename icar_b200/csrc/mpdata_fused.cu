// MPDATA order 2 + FCT, the production fp32 path, BIT-IDENTICAL to the reference arithmetic: three kernels per advect()
//
//   k_face_coefficients  once per call: the scalar-independent factors of the pseudo-velocity formulas
//                              (adv_mpdata.f90:129-251), rounded exactly as the Fortran rounds them
//   k_upwind_all               (mpdata_staged.cu) donor-cell pass 1 (adv_mpdata.f90:44-105), all scalars in one launch
//   k_corrector_fused          per scalar: pseudo-velocities + flux limiter (adv_mpdata.f90:257-354,
//                              adv_mpdata_FCT_core.f90) + second donor-cell pass (:389) in ONE kernel
//
// The kernel is a 2.5-D streaming stencil (corrector_fused.inc): a CTA owns a tile of 32 columns (28 outputs + 2 halo
// each side) x all levels and marches along y; blockIdx.x = scalar + nscalars * x-tile, so the CTAs of the scalars of a
// tile run side by side and the coefficient planes come from HBM once; the 14 boxes (32 x nz floats) a plane needs are
// fetched by TMA into a double-buffered shared-memory stage one plane ahead (mbarrier completion, zero fill outside the
// domain); y-coupling lives in register windows, x-coupling in warp shuffles, z-coupling between threads in a few
// shared-memory rows; two vertically adjacent cells are one f32x2 pair (FADD2 / FMUL2 / FFMA2).
// This TU is compiled with -fmad=false: the only fused multiply-adds are the explicit ones of the division sequence.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "f32x2.cuh"
#include "mpdata.cuh"
#include "tma_maps.cuh"

namespace icar {

namespace {

#define ICAR_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#include "corrector_prims.inc"
#include "corrector_fused.inc"

// the packed quotient next to the compiler's IEEE division, for icar_selftest_division
__global__ void k_selftest_division(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ packed,
                                    float* __restrict__ ieee, long long n) {
    const long long t = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (t + 1 >= n) return;
    const f2 q = div2x(f2{a[t], a[t + 1]}, f2{b[t], b[t + 1]});
    packed[t] = q.x; packed[t + 1] = q.y;
    ieee[t] = a[t] / b[t]; ieee[t + 1] = a[t + 1] / b[t + 1];
}

}  // namespace

int selftest_division(const float* a, const float* b, float* packed, float* ieee, long long n, cudaStream_t st) {
    k_selftest_division<<<(unsigned)((n / 2 + 255) / 256), 256, 0, st>>>(a, b, packed, ieee, n);
    return 1;
}

// Which instantiation serves nz levels, and its dynamic shared memory: 8 exchange rows of (nzt + 1) x 32 floats,
// 1 or 2 stages of 14 boxes of NZP x 32 floats, 2 mbarriers.
FusedPlan fused_plan(const Geom& g) {
    FusedPlan p{};
    if (g.nz < 3 || g.nz > 64 || g.nx < 3 || g.ny < 3) return p;
    p.stages = 2;
    if (g.nz <= 20) { p.cpt = 2; p.nzp = 20; }
    else if (g.nz <= 40) { p.cpt = 2; p.nzp = 40; }
    else if (g.nz <= 48) { p.cpt = 4; p.nzp = 48; }
    else if (g.nz <= 52) { p.cpt = 2; p.nzp = 52; }
    else if (g.nz <= 56) { p.cpt = 2; p.nzp = 56; }
    else if (g.nz <= 60) { p.cpt = 4; p.nzp = 60; }
    else { p.cpt = 4; p.nzp = 64; p.stages = 1; }
    const int nzt = p.nzp / p.cpt;
    p.smem = (size_t)(8 * (nzt + 1) * TW + p.stages * NFE * p.nzp * TW) * sizeof(float) + 16;
    int dev = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return p;
    }
    p.ok = p.smem <= (size_t)max_optin;
    return p;
}

int fused_advect_all(const Geom& g, const ScalarSet& s, const float* U, const float* V, const float* W,
                     const float* Wn, const float* jaco, const float* dz, const FusedCoef& co,
                     cudaStream_t st, cudaEvent_t* marks, int part, const int* own) {
    int launches = 0;
    if (s.n > FUSED_MAX_SCALARS) return -1;
    if (marks) cudaEventRecord(marks[0], st);
    if (part != 2) {
        k_face_coefficients<<<dim3((g.ld / 4 + 127) / 128, g.nz, g.ny), 128, 0, st>>>(g, U, V, Wn, jaco, co);
        launches += 1;
    }
    if (marks) cudaEventRecord(marks[1], st);
    if (part == 1) return launches;
    launches += staged_upwind_all(g, s, U, V, W, dz, jaco, st);   // also copies the boundary ring into `out`
    if (marks) cudaEventRecord(marks[2], st);

    FusedMaps maps;
    OutSet outs;
    outs.n = s.n;
    const float* cf[12] = {co.t0y, co.cby, co.ccy, co.ggy, co.t0x, co.cbx, co.ccx, co.t0z, co.cbz, co.ccz, jaco, dz};
    for (int n = 0; n < s.n; ++n) {
        outs.out[n] = s.out[n];
        if (!tma::field_map(g, s.q2[n], &maps.q2[n]) || !tma::field_map(g, s.q[n], &maps.l[n])) return -1;
    }
    for (int f = 0; f < 12; ++f)
        if (!tma::field_map(g, cf[f], &maps.coef[f])) return -1;

    const FusedPlan plan = fused_plan(g);
    if (!plan.ok) return -1;
    const int xtiles = g.nx / TOUT + 1;               // outputs of tile xt: columns 28*xt-2 .. 28*xt+25
    // y-chunks: pick the split that wastes least (partial last wave over the SMs vs. the 4 warm-up planes each chunk
    // re-reads); one CTA per SM is resident (shared-memory bound)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // cells whose result is stored: own = {i0, i1, j0, j1} (half-open, tile-local), default the tile without its ring
    OutRange range{1, g.nx - 1, 1, g.ny - 1};
    if (own) range = OutRange{std::max(own[0], 1), std::min(own[1], g.nx - 1), std::max(own[2], 1), std::min(own[3], g.ny - 1)};
    if (range.i1 <= range.i0 || range.j1 <= range.j0) return launches;
    const int per_row = xtiles * s.n, planes = range.j1 - range.j0;
    int best = 1;
    double best_cost = 1e30;
    for (int cand = 1; cand <= 64 && cand <= planes; ++cand) {
        const int len = (planes + cand - 1) / cand;
        const int ctas = per_row * ((planes + len - 1) / len);
        const int waves = (ctas + sms - 1) / sms;
        const double cost = (double)waves * (len + 4);
        if (cost < best_cost) { best_cost = cost; best = cand; }
    }
    const int chunk = (planes + best - 1) / best;
    const int chunks = (planes + chunk - 1) / chunk;
    dim3 grid(xtiles * s.n, chunks);
    dim3 block(TW, plan.nzp / plan.cpt);
#define LAUNCH_FUSED(NPAIRS, Z, S)                                                                            \
    {                                                                                                         \
        auto kern = k_corrector_fused<NPAIRS, Z, S>;                                                          \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) !=        \
            cudaSuccess) { (void)cudaGetLastError(); return -1; }                                             \
        kern<<<grid, block, plan.smem, st>>>(g, outs, maps, chunk, range);                                          \
    }
    if (plan.nzp == 20) LAUNCH_FUSED(1, 20, 2)
    else if (plan.nzp == 40) LAUNCH_FUSED(1, 40, 2)
    else if (plan.nzp == 48) LAUNCH_FUSED(2, 48, 2)
    else if (plan.nzp == 52) LAUNCH_FUSED(1, 52, 2)
    else if (plan.nzp == 56) LAUNCH_FUSED(1, 56, 2)
    else if (plan.nzp == 60) LAUNCH_FUSED(2, 60, 2)
    else LAUNCH_FUSED(2, 64, 1)
#undef LAUNCH_FUSED
    ++launches;
    if (marks) cudaEventRecord(marks[3], st);
    return launches;
}

bool fused_preload() {
    return preload_kernel(k_face_coefficients) && preload_kernel(k_corrector_fused<1, 20, 2>) &&
           preload_kernel(k_corrector_fused<1, 40, 2>) && preload_kernel(k_corrector_fused<2, 48, 2>) &&
           preload_kernel(k_corrector_fused<1, 52, 2>) && preload_kernel(k_corrector_fused<1, 56, 2>) &&
           preload_kernel(k_corrector_fused<2, 60, 2>) && preload_kernel(k_corrector_fused<2, 64, 1>);
}

}  // namespace icar
