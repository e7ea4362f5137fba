// MPDATA order 2 + FCT, the production fp32 path: three kernels per advect() call
//
//   k_face_coefficients  once per call: every scalar-independent factor of the pseudo-velocity
//                        formulas (adv_mpdata.f90:129-251, :383-385), one value per face
//   k_upwind_all         (mpdata_staged.cu) donor-cell pass 1 (adv_mpdata.f90:44-105) for ALL
//                        advected scalars in one launch, bit-exact
//   k_corrector_fused    per scalar: pseudo-velocities + flux limiter (adv_mpdata.f90:257-354,
//                        adv_mpdata_FCT_core.f90) + corrector donor-cell pass (:389) in ONE
//                        kernel, written as a 2.5-D streaming stencil:
//
//   * a CTA owns a tile of 32 columns (28 outputs + 2 halo each side) x all levels and marches
//     along y; blockIdx.x = scalar + nscalars * x-tile, so the CTAs of the 5 scalars of a tile run
//     side by side and the 11 coefficient planes come from HBM once and from L2 four times;
//   * every plane a CTA needs (q2, l and 11 coefficient fields: 13 boxes of 32 x nz floats) is
//     fetched by TMA (cp.async.bulk.tensor.3d, zero fill outside the domain) into a
//     double-buffered shared-memory stage one iteration ahead, completion on an mbarrier: no
//     thread ever waits on a global load, and the loads cost no issue slots and no registers;
//   * everything that couples cells along y lives in the owning thread's registers (sliding
//     windows), x-neighbours are warp shuffles (a warp is one 32-wide row of the tile),
//     z-neighbours between threads go through a few shared-memory rows;
//   * each face's pseudo-velocity, each cell's limiter bounds (beta_in / beta_out) and each limited
//     flux are computed exactly once (the staged kernels recompute them per consumer).
//
// Arithmetic of k_corrector_fused ("fast" mode): the same formulas as the Fortran, but
//   * the scalar-independent factors are pre-multiplied (k_face_coefficients),
//   * the transverse ratios use (a-b+c-d)/(a+b+c+d+eps) = (S+ - S-)/(S+ + S- + eps) with S the
//     face pair-sums, which are shared between terms,
//   * divisions are MUFU.RCP based (__fdividef, <= 2 ulp) and a*b+c contracts to FMA.
// These change roundings of the ANTI-DIFFUSIVE CORRECTION only (pass 1 stays bit-exact):
// differences of order 1e-7 of the field maximum per call; tests/test_gpu_parity.py states the
// tolerance and also runs the bit-exact staged path (icar_options.exact_arithmetic = 1).
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "common.cuh"
#include "f32x2.cuh"
#include "mpdata.cuh"
#include "tma_maps.cuh"

namespace icar {

namespace {

#define ICAR_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#include "corrector_prims.inc"
#include "corrector_fused.inc"

using tma::field_map;

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

}  // namespace

// Which instantiation serves nz levels, and its dynamic shared memory: 2 x 10 exchange rows of
// (nzt+2) x 32 floats, 1 or 2 stages of 13 boxes of NZP x 32 floats, 2 mbarriers.
FusedPlan fused_plan(const Geom& g) {
    FusedPlan p{};
    if (g.nz < 2 || g.nz > 64 || g.nx < 3 || g.ny < 3) return p;
    p.stages = 2;
    p.rowbufs = 2;
    p.nsync = 0;
    if (g.nz <= 20) { p.cpt = 2; p.nzp = 20; }
    else if (g.nz <= 40) { p.cpt = 2; p.nzp = 40; }
    else if (g.nz <= 48) { p.cpt = 4; p.nzp = 48; }
    else { p.cpt = 4; p.nzp = 64; p.stages = 1; }
    if (env_int("ICAR_FUSED_CPT", 0) == 4 && g.nz <= 48) { p.cpt = 4; p.nzp = 48; }
    // opt-in until it has run on hardware: 49..52 levels (BASELINE configs[3] has 50) still fit TWO stages
    // (2 x 13 x 52 x 128 B + rows = 211 KB), so the TMA loads stay one plane ahead instead of being exposed
    if (env_int("ICAR_FUSED_NZP52", 0) == 1 && g.nz > 48 && g.nz <= 52) { p.cpt = 4; p.nzp = 52; p.stages = 2; }
    // opt-in, not yet run on hardware (emulated on the CPU, tests/test_corrector_host.py): 53..60 levels (BASELINE
    // configs[4] has 60) with TWO stages and ONE set of exchange rows, packed kernel only (needs ICAR_CORR_V2=1):
    // 2 x 13 x 60 x 128 B + 10 x 17 x 128 B = 221 KB
    if (env_int("ICAR_FUSED_NZP60", 0) == 1 && env_int("ICAR_CORR_V2", 0) == 1 && g.nz > 52 && g.nz <= 60) {
        p.cpt = 4; p.nzp = 60; p.stages = 2; p.rowbufs = 1;
    }
    const int nzt = p.nzp / p.cpt;
    // opt-in, not yet run on hardware (emulated on the CPU): the packed kernel without CTA-wide barriers in its plane loop
    // (a warp synchronises with the warps tz - 1 and tz + 1 only), for the two-stage, two-row-set instantiations
    if (env_int("ICAR_CORR_NSYNC", 0) == 1 && env_int("ICAR_CORR_V2", 0) == 1 && p.stages == 2 && p.rowbufs == 2) p.nsync = 1;
    p.smem = (size_t)(p.rowbufs * 10 * (nzt + 2) * TW + p.stages * NF * p.nzp * TW) * sizeof(float) + 16 +
             (p.nsync ? (size_t)(2 + 3 * nzt) * 8 : 0);
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
                     cudaStream_t st, cudaEvent_t* marks, int part) {
    int launches = 0;
    if (s.n > FUSED_MAX_SCALARS) return -1;
    dim3 cgrid((g.nx + 31) / 32, (g.nz + 3) / 4, (g.ny + 3) / 4);
    if (marks) cudaEventRecord(marks[0], st);
    if (part != 2) {
        if (g.ld % 4 == 0 && !std::getenv("ICAR_COEF_SCALAR"))
            k_face_coefficients_v4<<<dim3((g.ld / 4 + 127) / 128, g.nz, g.ny), 128, 0, st>>>(g, U, V, Wn, jaco, dz, co);
        else
            k_face_coefficients<<<cgrid, dim3(32, 4, 4), 0, st>>>(g, U, V, Wn, jaco, dz, co);
        launches += 1;
    }
    if (marks) cudaEventRecord(marks[1], st);
    if (part == 1) return launches;
    launches += staged_upwind_all(g, s, U, V, W, dz, jaco, st);   // also copies the boundary ring into `out`
    if (marks) cudaEventRecord(marks[2], st);

    FusedMaps maps;
    OutSet outs;
    outs.n = s.n;
    const float* cf[11] = {co.ay, co.by, co.cy, co.ax, co.bx, co.cx, co.az, co.bz, co.cz, co.rj, co.rjz};
    for (int n = 0; n < s.n; ++n) {
        outs.out[n] = s.out[n];
        if (!field_map(g, s.q2[n], &maps.q2[n]) || !field_map(g, s.q[n], &maps.l[n])) return -1;
    }
    for (int f = 0; f < 11; ++f)
        if (!field_map(g, cf[f], &maps.coef[f])) return -1;

    FusedPlan plan = fused_plan(g);
    if (!plan.ok) return -1;
    const int xtiles = g.nx / TOUT + 1;               // outputs of tile xt: columns 28*xt-2 .. 28*xt+25
    // y-chunks: pick the split that wastes least (partial last wave of 148 SMs vs. the 4 warm-up
    // planes each chunk re-reads); one CTA per SM is resident (shared-memory bound)
    const int per_row = xtiles * s.n, planes = g.ny - 2;
    int best = 1;
    double best_cost = 1e30;
    for (int cand = 1; cand <= 64 && cand <= planes; ++cand) {
        const int len = (planes + cand - 1) / cand;
        const int ctas = per_row * ((planes + len - 1) / len);
        const int waves = (ctas + 147) / 148;
        const double cost = (double)waves * (len + 4);
        if (cost < best_cost) { best_cost = cost; best = cand; }
    }
    const int chunk = (planes + best - 1) / best;
    const int chunks = (planes + chunk - 1) / chunk;
    dim3 grid(xtiles * s.n, chunks);
    dim3 block(TW, plan.nzp / plan.cpt);
#define LAUNCH_FUSED(C, Z, S)                                                                                 \
    {                                                                                                         \
        auto kern = k_corrector_fused<C, Z, S>;                                                               \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) !=        \
            cudaSuccess) { (void)cudaGetLastError(); return -1; }                                             \
        kern<<<grid, block, plan.smem, st>>>(g, outs, maps, chunk);                                           \
    }
    // CPT = 2: the f32x2 form of the kernel (ICAR_CORR_SCALAR=1 selects the scalar one, the reference form)
    const bool packed = plan.cpt == 2 && !std::getenv("ICAR_CORR_SCALAR");
#define LAUNCH_PACKED(Z)                                                                                      \
    {                                                                                                         \
        auto kern = k_corrector_fused_p2<Z>;                                                                  \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) !=        \
            cudaSuccess) { (void)cudaGetLastError(); return -1; }                                             \
        kern<<<grid, block, plan.smem, st>>>(g, outs, maps, chunk);                                           \
    }
    // candidate form of the packed kernel, opt-in until it has been validated on hardware
    const bool v2 = env_int("ICAR_CORR_V2", 0) == 1;
#define LAUNCH_PK(NPAIRS, Z, S, ...)                                                                          \
    {                                                                                                         \
        auto kern = k_corrector_fused_pk<NPAIRS, Z, S, ##__VA_ARGS__>;                                                     \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) !=        \
            cudaSuccess) { (void)cudaGetLastError(); return -1; }                                             \
        kern<<<grid, block, plan.smem, st>>>(g, outs, maps, chunk);                                           \
    }
    if (v2 && plan.nsync && plan.cpt == 2 && plan.nzp == 20) LAUNCH_PK(1, 20, 2, 2, 1)
    else if (v2 && plan.nsync && plan.cpt == 2 && plan.nzp == 40) LAUNCH_PK(1, 40, 2, 2, 1)
    else if (v2 && plan.nsync && plan.nzp == 48) LAUNCH_PK(2, 48, 2, 2, 1)
    else if (v2 && plan.nsync && plan.nzp == 52) LAUNCH_PK(2, 52, 2, 2, 1)
    else if (v2 && plan.cpt == 2 && plan.nzp == 20) LAUNCH_PK(1, 20, 2)
    else if (v2 && plan.cpt == 2 && plan.nzp == 40) LAUNCH_PK(1, 40, 2)
    else if (v2 && plan.nzp == 48) LAUNCH_PK(2, 48, 2)
    else if (v2 && plan.nzp == 52) LAUNCH_PK(2, 52, 2)
    else if (v2 && plan.nzp == 60 && plan.rowbufs == 1) LAUNCH_PK(2, 60, 2, 1)
    else if (v2 && plan.nzp == 64) LAUNCH_PK(2, 64, 1)
    else if (packed && plan.nzp == 20) LAUNCH_PACKED(20)
    else if (packed && plan.nzp == 40) LAUNCH_PACKED(40)
    else if (plan.cpt == 2 && plan.nzp == 20) LAUNCH_FUSED(2, 20, 2)
    else if (plan.cpt == 2 && plan.nzp == 40) LAUNCH_FUSED(2, 40, 2)
    else if (plan.nzp == 48) LAUNCH_FUSED(4, 48, 2)
    else if (plan.nzp == 52) LAUNCH_FUSED(4, 52, 2)
    else LAUNCH_FUSED(4, 64, 1)
#undef LAUNCH_FUSED
#undef LAUNCH_PACKED
#undef LAUNCH_PK
    ++launches;
    if (marks) cudaEventRecord(marks[3], st);
    return launches;
}

}  // namespace icar
