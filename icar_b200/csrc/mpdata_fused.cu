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

namespace icar {

namespace {

constexpr int TW = 32;          // tile width in columns == warp size
constexpr int XH = 2;           // x halo columns on each side of the 28 output columns
constexpr int TOUT = TW - 2 * XH;
constexpr float EPS = 1e-10f;
constexpr float E15 = 1e-15f;
constexpr float BIG = 3.0e38f;
constexpr int NF = 13;          // boxes per stage
enum { F_Q2, F_L, F_AY, F_BY, F_CY, F_AX, F_BX, F_CX, F_AZ, F_BZ, F_CZ, F_RJ, F_RJZ };

constexpr int FUSED_MAX_SCALARS = 9;   // keeps the kernel parameters under 4 KB (29 maps x 128 B)

struct alignas(64) FusedMaps {
    CUtensorMap q2[FUSED_MAX_SCALARS];
    CUtensorMap l[FUSED_MAX_SCALARS];
    CUtensorMap coef[11];       // ay by cy ax bx cx az bz cz rj rjz  (stage order F_AY..F_RJZ)
};

struct OutSet {
    int n;
    float* out[FUSED_MAX_SCALARS];
};

// a / b as a * rcp(b): one MUFU.RCP + one FMUL.  (__fdividef wraps the same two instructions in a
// denormal-range rescue -- FSETP + two predicated FMULs -- that cost 21 % of this kernel's instructions;
// every divisor here is a sum + 1e-10 or + 1e-15 of non-negative terms, far inside the normal range.)
__device__ __forceinline__ float fdiv(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float lane_lo(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }    // column i-1
__device__ __forceinline__ float lane_hi(float v) { return __shfl_down_sync(0xffffffffu, v, 1); }  // column i+1

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
// one 32 x nz x 1 box of a (nx, nz, ny) field -> dense [nz][32] floats in shared memory
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int z, int y) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(z), "r"(y)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// Per-face scalar-independent coefficients.  With A = (r-l)/(r+l+eps) and B, C the two
// transverse ratios, the scaled pseudo-velocities of adv_mpdata.f90:129-251,383-385 are
//     u2 = ax*A - bx*B - cx*C        (x-face between i-1 and i, stored at i)
//     v2 = ay*A - by*B - cy*C        (y-face between j-1 and j, stored at j)
//     w2 = az*A - bz*B - cz*C        (z-face above k, stored at k; includes the *0.5*dz)
// Terms the Fortran skips at domain edges get a zero coefficient.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_face_coefficients(Geom g, const float* __restrict__ U, const float* __restrict__ V,
                    const float* __restrict__ Wn, const float* __restrict__ G,
                    const float* __restrict__ dz, FusedCoef co) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    const size_t ps = g.ps();
    const int ld = g.ld;
    const size_t c = g.at(i, k, j);
    const bool iin = (i > 0 && i < g.nx - 1), jin = (j > 0 && j < g.ny - 1), kin = (k > 0 && k < g.nz - 1);
    const float Gc = G[c];
    float ax = 0.f, bx = 0.f, cx = 0.f, ay = 0.f, by = 0.f, cy = 0.f, az = 0.f, bz = 0.f, cz = 0.f;
    if (i >= 1) {
        const float u = U[c], GG = Gc + G[c - 1], rGG = fdiv(1.f, GG);
        ax = 0.5f * (fabsf(u) * (1.f - fabsf(u) * (2.f * rGG)));
        if (jin) bx = 0.5f * (0.5f * u * (0.25f * (V[c] + V[c + ps] + V[c - 1] + V[c - 1 + ps]))) * rGG;
        if (kin) cx = 0.5f * (0.5f * u * (0.25f * (Wn[c] + Wn[c - ld] + Wn[c - 1] + Wn[c - 1 - ld]))) * rGG;
    }
    if (j >= 1) {
        const float v = V[c], GG = Gc + G[c - ps], rGG = fdiv(1.f, GG);
        ay = 0.5f * (fabsf(v) * (1.f - fabsf(v) * (2.f * rGG)));
        if (iin) by = 0.5f * (0.5f * v * (0.25f * (U[c + 1] + U[c + 1 - ps] + U[c] + U[c - ps]))) * rGG;
        if (kin) cy = 0.5f * (0.5f * v * (0.25f * (Wn[c] + Wn[c - ld] + Wn[c - ps] + Wn[c - ld - ps]))) * rGG;
    }
    if (k < g.nz - 1) {
        const float w = Wn[c], GG = G[c + ld] + Gc, hdz = 0.5f * dz[c], rGG = fdiv(1.f, GG);
        az = hdz * (fabsf(w) * (1.f - fabsf(w) * (2.f * rGG)));
        if (iin) bz = hdz * (0.5f * w * (0.25f * (U[c + 1] + U[c + 1 + ld] + U[c] + U[c + ld]))) * rGG;
        if (jin) cz = hdz * (0.5f * w * (0.25f * (V[c] + V[c + ld] + V[c + ps] + V[c + ld + ps]))) * rGG;
    }
    co.ax[c] = ax; co.bx[c] = bx; co.cx[c] = cx;
    co.ay[c] = ay; co.by[c] = by; co.cy[c] = cy;
    co.az[c] = az; co.bz[c] = bz; co.cz[c] = cz;
    co.rj[c] = fdiv(1.f, Gc);                 // rho == 1 on this path (advect_density = .false.): G == jaco
    co.rjz[c] = fdiv(1.f, dz[c] * Gc);
}

// The same coefficients, four consecutive cells per thread: every row the stencil touches is fetched as one
// 16-byte load (+ one scalar for the i-1 / i+4 neighbour), the eleven outputs leave as 16-byte stores.
// Same expressions as k_face_coefficients (which stays as the reference form, ICAR_COEF_SCALAR=1 selects it).
struct Row5 { float m, v[4], p; };       // values at i0-1, i0..i0+3, i0+4 of one x-row (0 outside the row)
__device__ __forceinline__ Row5 load_row(const float* __restrict__ a, size_t c, int i0, int nx, bool need_m, bool need_p) {
    Row5 r;
    const float4 q = *reinterpret_cast<const float4*>(a + c);
    r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    r.m = (need_m && i0 > 0) ? a[c - 1] : 0.f;
    r.p = (need_p && i0 + 4 < nx) ? a[c + 4] : 0.f;
    return r;
}
__device__ __forceinline__ float at5(const Row5& r, int e) { return e < 0 ? r.m : (e > 3 ? r.p : r.v[e]); }

__global__ void __launch_bounds__(128)
k_face_coefficients_v4(Geom g, const float* __restrict__ U, const float* __restrict__ V,
                       const float* __restrict__ Wn, const float* __restrict__ G,
                       const float* __restrict__ dz, FusedCoef co) {
    const int i0 = 4 * (blockIdx.x * 128 + threadIdx.x), k = blockIdx.y, j = blockIdx.z;
    if (i0 >= g.ld) return;
    const size_t ps = g.ps();
    const int ld = g.ld, nx = g.nx;
    const size_t c = g.at(i0, k, j);
    const bool jin = (j > 0 && j < g.ny - 1), kin = (k > 0 && k < g.nz - 1), jlo = j >= 1, kup = k < g.nz - 1;
    const Row5 zero{0.f, {0.f, 0.f, 0.f, 0.f}, 0.f};
    // rows: U needs i..i+1, V / Wn / G need i-1..i
    const Row5 u00 = load_row(U, c, i0, nx, false, true);
    const Row5 u0m = jlo ? load_row(U, c - ps, i0, nx, false, true) : zero;          // (k, j-1)
    const Row5 up0 = kup ? load_row(U, c + ld, i0, nx, false, true) : zero;          // (k+1, j)
    const Row5 v00 = load_row(V, c, i0, nx, true, false);
    const Row5 v0p = load_row(V, c + ps, i0, nx, true, false);                       // (k, j+1): V has ny+1 planes
    const Row5 vp0 = kup ? load_row(V, c + ld, i0, nx, false, false) : zero;         // (k+1, j)
    const Row5 vpp = kup ? load_row(V, c + ld + ps, i0, nx, false, false) : zero;    // (k+1, j+1)
    const Row5 w00 = load_row(Wn, c, i0, nx, true, false);
    const Row5 wm0 = (k > 0) ? load_row(Wn, c - ld, i0, nx, true, false) : zero;     // (k-1, j)
    const Row5 w0m = jlo ? load_row(Wn, c - ps, i0, nx, false, false) : zero;        // (k, j-1)
    const Row5 wmm = (jlo && k > 0) ? load_row(Wn, c - ld - ps, i0, nx, false, false) : zero;
    const Row5 g00 = load_row(G, c, i0, nx, true, false);
    const Row5 g0m = jlo ? load_row(G, c - ps, i0, nx, false, false) : zero;
    const Row5 gp0 = kup ? load_row(G, c + ld, i0, nx, false, false) : zero;
    const float4 dz4 = *reinterpret_cast<const float4*>(dz + c);
    const float dzz[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    float o[11][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + e;
        const bool in = i < nx, iin = (i > 0 && i < nx - 1);
        const float Gc = g00.v[e];
        float ax = 0.f, bx = 0.f, cx = 0.f, ay = 0.f, by = 0.f, cy = 0.f, az = 0.f, bz = 0.f, cz = 0.f;
        if (in && i >= 1) {
            const float u = u00.v[e], GG = Gc + at5(g00, e - 1), rGG = fdiv(1.f, GG);
            ax = 0.5f * (fabsf(u) * (1.f - fabsf(u) * (2.f * rGG)));
            if (jin) bx = 0.5f * (0.5f * u * (0.25f * (v00.v[e] + v0p.v[e] + at5(v00, e - 1) + at5(v0p, e - 1)))) * rGG;
            if (kin) cx = 0.5f * (0.5f * u * (0.25f * (w00.v[e] + wm0.v[e] + at5(w00, e - 1) + at5(wm0, e - 1)))) * rGG;
        }
        if (in && jlo) {
            const float v = v00.v[e], GG = Gc + g0m.v[e], rGG = fdiv(1.f, GG);
            ay = 0.5f * (fabsf(v) * (1.f - fabsf(v) * (2.f * rGG)));
            if (iin) by = 0.5f * (0.5f * v * (0.25f * (at5(u00, e + 1) + at5(u0m, e + 1) + u00.v[e] + u0m.v[e]))) * rGG;
            if (kin) cy = 0.5f * (0.5f * v * (0.25f * (w00.v[e] + wm0.v[e] + w0m.v[e] + wmm.v[e]))) * rGG;
        }
        if (in && kup) {
            const float w = w00.v[e], GG = gp0.v[e] + Gc, hdz = 0.5f * dzz[e], rGG = fdiv(1.f, GG);
            az = hdz * (fabsf(w) * (1.f - fabsf(w) * (2.f * rGG)));
            if (iin) bz = hdz * (0.5f * w * (0.25f * (at5(u00, e + 1) + at5(up0, e + 1) + u00.v[e] + up0.v[e]))) * rGG;
            if (jin) cz = hdz * (0.5f * w * (0.25f * (v00.v[e] + vp0.v[e] + v0p.v[e] + vpp.v[e]))) * rGG;
        }
        o[0][e] = ax; o[1][e] = bx; o[2][e] = cx; o[3][e] = ay; o[4][e] = by; o[5][e] = cy;
        o[6][e] = az; o[7][e] = bz; o[8][e] = cz;
        o[9][e] = in ? fdiv(1.f, Gc) : 0.f;
        o[10][e] = in ? fdiv(1.f, dzz[e] * Gc) : 0.f;
    }
    float* const outp[11] = {co.ax, co.bx, co.cx, co.ay, co.by, co.cy, co.az, co.bz, co.cz, co.rj, co.rjz};
#pragma unroll
    for (int f = 0; f < 11; ++f)
        *reinterpret_cast<float4*>(outp[f] + c) = make_float4(o[f][0], o[f][1], o[f][2], o[f][3]);
}

// ---------------------------------------------------------------------------------------------
// The fused corrector.  blockDim = (32, NZT); thread (tx, tz) owns cells (i, CPT*tz + c), c < CPT.
// ---------------------------------------------------------------------------------------------
// NZP = number of level slots per staged box (>= nz, multiple of CPT): compile-time, so that every
// shared-memory address below is base register + immediate.
// STAGES = 2: the TMA stage is double-buffered (the boxes of plane P+1 land while plane P is computed);
// STAGES = 1 (50 < nz <= 64 levels, where two stages of 13 boxes no longer fit 227 KB of shared memory):
// one stage, refilled at the top of every iteration -- the load latency is then exposed once per plane.
template <int CPT, int NZP, int STAGES>
__global__ void __launch_bounds__(TW * (NZP / CPT))
k_corrector_fused(const Geom g, const OutSet outs, const __grid_constant__ FusedMaps maps, const int chunk) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int nzt = NZP / CPT;
    const int tx = threadIdx.x, tz = threadIdx.y;
    const int scalar = blockIdx.x % outs.n, xt = blockIdx.x / outs.n;
    float* __restrict__ gout = outs.out[scalar];
    const int nx = g.nx, nz = g.nz, ny = g.ny;
    // TMA needs the box origin 16-byte aligned in the innermost dimension (scripts/probe/tma_probe.cu:
    // x0 = -1 traps, x0 = -4 / 40 work), so tiles start at a multiple of 4 columns: 28*xt - 4
    const int x0 = xt * TOUT - 4;                     // first column of the tile (may be negative)
    const int i = x0 + tx;                            // output columns are tx in [XH, TW-XH)
    const bool col = (i >= 0 && i < nx);
    const bool ilast = (i == nx - 1);
    const float xedge = (i == 0 || ilast) ? 0.f : 1.f;           // fin = fout = 0 at the x line ends
    const int ja = 1 + blockIdx.y * chunk;            // first output plane of this chunk
    const int jb = min(ja + chunk, ny - 1);           // one past the last output plane
    const int k0 = tz * CPT;
    constexpr int box = NZP * TW;                     // floats per staged box slot (TMA fills nz rows of it)

    // ---- shared memory carve-up
    enum { R_QLO, R_QHI, R_MHLO, R_MHHI, R_MLLO, R_MLHI, R_FZHI, R_BILO, R_BOLO, R_FLHI, R_COUNT };
    float* rows = reinterpret_cast<float*>(smraw);
    constexpr int rows_floats = 2 * R_COUNT * (nzt + 2) * TW;
    float* stage0 = rows + rows_floats;               // [STAGES][NF][nz][32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage0 + STAGES * NF * box);
    auto row = [&](int buf, int kind, int z) -> float& { return rows[((buf * R_COUNT + kind) * (nzt + 2) + (z + 1)) * TW + tx]; };
    const int tid = tz * TW + tx;
    const int niter = jb - ja + 4;                    // planes ja-2 .. jb+1
    const uint32_t stage_bytes = (uint32_t)(NF * nz * TW * sizeof(float));   // bytes TMA delivers per stage

    auto issue = [&](int it) {                        // thread 0: fetch the 13 boxes of iteration `it`
        const int P = ja - 2 + it, st = (STAGES == 2) ? (it & 1) : 0;
        float* dst = stage0 + st * NF * box;
        mbar_arrive_expect_tx(&mbar[st], stage_bytes);
        tma_load_box(dst + F_Q2 * box, &maps.q2[scalar], &mbar[st], x0, 0, P);
        tma_load_box(dst + F_L * box, &maps.l[scalar], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 0; f < 3; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 3; f < 11; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P - 2);
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    // level slots nz..NZP-1 of every box are never written by TMA: zero them once so that the
    // (non-existent) cells above the model top read as 0 without a predicate
    for (int t = tid; t < STAGES * NF * box; t += TW * nzt) stage0[t] = 0.f;
    // rows of the (non-existent) threads below tz=0 and above tz=nzt-1 read as "no neighbour"
    if (tz == 0) {
        for (int b = 0; b < 2; ++b) {
            row(b, R_QHI, -1) = 0.f; row(b, R_MHHI, -1) = -BIG; row(b, R_MLHI, -1) = BIG;
            row(b, R_FZHI, -1) = 0.f; row(b, R_FLHI, -1) = 0.f;
            row(b, R_QLO, nzt) = 0.f; row(b, R_MHLO, nzt) = -BIG; row(b, R_MLLO, nzt) = BIG;
            row(b, R_BILO, nzt) = 0.f; row(b, R_BOLO, nzt) = 0.f;
        }
    }
    __syncthreads();
    if (tid == 0) issue(0);

    // ---- per-cell register state (planes are relative to the newest loaded plane P)
    float a[CPT][4];          // q2 at planes P-3, P-2, P-1, P
    float sx[CPT][3];         // x-face pair sums  q2(i)+q2(i-1)   at planes P-3, P-2, P-1
    float sz[CPT][3];         // z-face pair sums  q2(k+1)+q2(k)   at planes P-3, P-2, P-1
    float mh[CPT][3], ml[CPT][3];   // max/min(q2, l) at planes P-2, P-1, P
    float v2a[CPT];           // unlimited v2 at face (P-2 | P-1)
    float fya[CPT];           // its donor flux
    float byi[CPT], byo[CPT]; // y-limiter beta_in / beta_out of cell plane P-2
    float fylp[CPT];          // limited y flux through face (P-3 | P-2)
    float qab[4], qbl[4];     // q2 of the row above (k0+CPT) / below (k0-1) at planes P-3..P
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
#pragma unroll
        for (int t = 0; t < 4; ++t) a[c][t] = 0.f;
#pragma unroll
        for (int t = 0; t < 3; ++t) { sx[c][t] = 0.f; sz[c][t] = 0.f; mh[c][t] = -BIG; ml[c][t] = BIG; }
        v2a[c] = 0.f; fya[c] = 0.f; byi[c] = 0.f; byo[c] = 0.f; fylp[c] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) { qab[t] = 0.f; qbl[t] = 0.f; }
    const int cell_off = k0 * TW + tx;                // this thread's first cell inside a box

    for (int it = 0; it < niter; ++it) {
        const int P = ja - 2 + it;
        const int buf = it & 1;
        const bool pin = (P >= 0 && P < ny);          // plane P exists
        const int Po = P - 2;                         // output plane of this iteration
        const bool out_iter = (Po >= ja);
        // every read of stage buf^1 (iteration it-1) precedes that iteration's first barrier,
        // which this thread has passed: the stage is free to be refilled
        if (STAGES == 2) { if (tid == 0 && it + 1 < niter) issue(it + 1); }
        else             { if (tid == 0 && it > 0) issue(it); }
        const int st = (STAGES == 2) ? buf : 0;
        while (!mbar_try_wait(&mbar[st], (uint32_t)((STAGES == 2) ? ((it >> 1) & 1) : (it & 1)))) {}
        const float* __restrict__ sg = stage0 + st * NF * box + cell_off;

        // ---------------- shift the windows, take plane P from the stage
        float cay[CPT], cby[CPT], ccy[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            a[c][0] = a[c][1]; a[c][1] = a[c][2]; a[c][2] = a[c][3];
            mh[c][0] = mh[c][1]; mh[c][1] = mh[c][2];
            ml[c][0] = ml[c][1]; ml[c][1] = ml[c][2];
            const bool cell = pin && col && (k0 + c < nz);
            const float q2v = sg[F_Q2 * box + c * TW];                   // TMA zero-fills outside the domain
            const float lv = sg[F_L * box + c * TW];
            cay[c] = sg[F_AY * box + c * TW];
            cby[c] = sg[F_BY * box + c * TW];
            ccy[c] = sg[F_CY * box + c * TW];
            a[c][3] = q2v;
            mh[c][2] = cell ? fmaxf(q2v, lv) : -BIG;
            ml[c][2] = cell ? fminf(q2v, lv) : BIG;
        }
        qab[0] = qab[1]; qab[1] = qab[2]; qab[2] = qab[3];
        qbl[0] = qbl[1]; qbl[1] = qbl[2]; qbl[2] = qbl[3];

        // ---------------- in-plane pseudo-velocities of plane Po = P-2 (data complete since P-1)
        // (all shuffles are executed by every lane of every warp)
        float u2[CPT], w2[CPT], fx[CPT], fz[CPT];
        float rj[CPT], rjz[CPT];
        if (out_iter) {
            const float sx_ab = qab[1] + lane_lo(qab[1]);      // x pair sum of the row above / below
            const float sx_bl = qbl[1] + lane_lo(qbl[1]);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const float cax = sg[F_AX * box + c * TW], cbx = sg[F_BX * box + c * TW],
                            ccx = sg[F_CX * box + c * TW], caz = sg[F_AZ * box + c * TW],
                            cbz = sg[F_BZ * box + c * TW], ccz = sg[F_CZ * box + c * TW];
                rj[c] = sg[F_RJ * box + c * TW];
                rjz[c] = sg[F_RJZ * box + c * TW];
                const float q = a[c][1];
                const float ql = lane_lo(q);
                const float qu = (c + 1 < CPT) ? a[(c + 1 < CPT) ? c + 1 : c][1] : qab[1];
                // x-face (i-1 | i)
                const float sxc = sx[c][1];                                  // q(i)+q(i-1)
                const float Ax = fdiv(q - ql, sxc + EPS);
                const float sp = sx[c][2], sn = sx[c][0];                    // planes Po+1, Po-1
                const float Bx = fdiv(sp - sn, sp + sn + EPS);
                const float sxu = (c + 1 < CPT) ? sx[(c + 1 < CPT) ? c + 1 : c][1] : sx_ab;
                const float sxd = (c > 0) ? sx[(c > 0) ? c - 1 : c][1] : sx_bl;
                const float Cx = fdiv(sxu - sxd, sxu + sxd + EPS);
                u2[c] = cax * Ax - cbx * Bx - ccx * Cx;
                // z-face above k
                const float szc = sz[c][1];                                  // q(k+1)+q(k)
                const float Az = fdiv(qu - q, szc + EPS);
                const float szr = lane_hi(szc), szl = lane_lo(szc);
                const float Bz = fdiv(szr - szl, szr + szl + EPS);
                const float zp = sz[c][2], zn = sz[c][0];
                const float Cz = fdiv(zp - zn, zp + zn + EPS);
                w2[c] = caz * Az - cbz * Bz - ccz * Cz;
                // donor fluxes of the UNLIMITED pseudo-velocities (FCT core: f = flux1(q1(c), q1(c+1), U2))
                fx[c] = donor(u2[c], ql, q);
                fz[c] = donor(w2[c], q, qu);
            }
        }
        // ---------------- publish rows for the z-neighbour threads
        row(buf, R_QLO, tz) = a[0][3];
        row(buf, R_QHI, tz) = a[CPT - 1][3];
        if (out_iter) {
            row(buf, R_MHLO, tz) = mh[0][0]; row(buf, R_MHHI, tz) = mh[CPT - 1][0];
            row(buf, R_MLLO, tz) = ml[0][0]; row(buf, R_MLHI, tz) = ml[CPT - 1][0];
            row(buf, R_FZHI, tz) = fz[CPT - 1];
        }
        __syncthreads();            // (no read of the TMA stage below this line)
        qab[3] = row(buf, R_QLO, tz + 1);
        qbl[3] = row(buf, R_QHI, tz - 1);

        // ---------------- y-face (P-1 | P): unlimited v2, then beta_y of plane P-1
        float v2b[CPT], fyb[CPT], syv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) syv[c] = a[c][3] + a[c][2];
        const float sy_ab = qab[3] + qab[2], sy_bl = qbl[3] + qbl[2];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float syr = lane_hi(syv[c]), syl = lane_lo(syv[c]);
            const float Ay = fdiv(a[c][3] - a[c][2], syv[c] + EPS);
            const float By = fdiv(syr - syl, syr + syl + EPS);
            const float syu = (c + 1 < CPT) ? syv[(c + 1 < CPT) ? c + 1 : c] : sy_ab;
            const float syd = (c > 0) ? syv[(c > 0) ? c - 1 : c] : sy_bl;
            const float Cy = fdiv(syu - syd, syu + syd + EPS);
            v2b[c] = cay[c] * Ay - cby[c] * By - ccy[c] * Cy;     // ay = by = cy = 0 for j = 0 (no face)
            fyb[c] = donor(v2b[c], a[c][2], a[c][3]);            // face (P-1 | P)
        }
        // beta_y of cell plane Pc = P-1 (needs fy of faces (P-2|P-1), (P-1|P) and mh/ml of P-2..P)
        float byi1[CPT], byo1[CPT];
        {
            const int Pc = P - 1;
            const bool ylast = (Pc == ny - 1);
            const float yedge = (Pc == 0 || ylast) ? 0.f : 1.f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const float q = a[c][2];
                const float hi = max3(mh[c][0], ylast ? q : mh[c][1], mh[c][2]);
                const float lo = min3(ml[c][0], ylast ? q : ml[c][1], ml[c][2]);
                const float fin = (fmaxf(0.f, fya[c]) - fminf(0.f, fyb[c])) * yedge;
                const float fout = (fmaxf(0.f, fyb[c]) - fminf(0.f, fya[c])) * yedge;
                byi1[c] = fdiv(hi - q, fin + E15);
                byo1[c] = fdiv(q - lo, fout + E15);
            }
        }
        // limited y flux through face (P-2 | P-1)
        float fyl[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float v = v2a[c];
            const float lim = (v > 0.f) ? min3(1.f, byi1[c], byo[c]) : min3(1.f, byi[c], byo1[c]);
            fyl[c] = lim * fya[c];                 // same donor cell as the unlimited flux: donor(lim*v) == lim*f
        }

        // ---------------- x / z limiter and corrector update of plane Po
        if (out_iter) {
            const float mhb = row(buf, R_MHHI, tz - 1), mlb = row(buf, R_MLHI, tz - 1);   // row below k0
            const float mha = row(buf, R_MHLO, tz + 1), mla = row(buf, R_MLLO, tz + 1);   // row above
            const float fzb = row(buf, R_FZHI, tz - 1);                                   // f_z of face above k0-1
            float bxi[CPT], bxo[CPT], bzi[CPT], bzo[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int k = k0 + c;
                const float q = a[c][1];
                const float mhc = mh[c][0], mlc = ml[c][0];
                // x direction
                {
                    const float mhl = lane_lo(mhc), mhr = lane_hi(mhc), mll = lane_lo(mlc), mlr = lane_hi(mlc);
                    const float fxp = lane_hi(fx[c]);
                    const float hi = max3(mhl, ilast ? q : mhc, mhr);       // lanes outside the domain hold -BIG
                    const float lo = min3(mll, ilast ? q : mlc, mlr);
                    const float fin = (fmaxf(0.f, fx[c]) - fminf(0.f, fxp)) * xedge;
                    const float fout = (fmaxf(0.f, fxp) - fminf(0.f, fx[c])) * xedge;
                    bxi[c] = fdiv(hi - q, fin + E15);
                    bxo[c] = fdiv(q - lo, fout + E15);
                }
                // z direction (flux_is_w rules at the line ends, FCT core :75-77, :92-94)
                {
                    const bool zlast = (k == nz - 1);
                    const float mhd = (c > 0) ? mh[(c > 0) ? c - 1 : c][0] : mhb;
                    const float mld = (c > 0) ? ml[(c > 0) ? c - 1 : c][0] : mlb;
                    const float mhu = (c + 1 < CPT) ? mh[(c + 1 < CPT) ? c + 1 : c][0] : mha;
                    const float mlu = (c + 1 < CPT) ? ml[(c + 1 < CPT) ? c + 1 : c][0] : mla;
                    const float fzm = (c > 0) ? fz[(c > 0) ? c - 1 : c] : fzb;
                    const float fzp = zlast ? fzm : fz[c];            // top: fin = fout = |f(nz-2)|
                    const float hi = max3(mhd, zlast ? q : mhc, mhu);     // the slot above the top level holds -BIG
                    const float lo = min3(mld, zlast ? q : mlc, mlu);
                    const float fin = fmaxf(0.f, fzm) - fminf(0.f, fzp);
                    const float fout = fmaxf(0.f, fzp) - fminf(0.f, fzm);
                    bzi[c] = fdiv(hi - q, fin + E15);
                    bzo[c] = fdiv(q - lo, fout + E15);
                }
            }
            row(buf, R_BILO, tz) = bzi[0];
            row(buf, R_BOLO, tz) = bzo[0];
            __syncthreads();
            const float bia = row(buf, R_BILO, tz + 1), boa = row(buf, R_BOLO, tz + 1);
            float fxl[CPT], fzl[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                // limited x flux through face (i-1 | i)
                const float bil = lane_lo(bxi[c]), bol = lane_lo(bxo[c]);
                const float u = u2[c];
                const float limx = (u > 0.f) ? min3(1.f, bxi[c], bol) : min3(1.f, bil, bxo[c]);
                fxl[c] = limx * fx[c];
                // limited z flux through the face above k
                const float biu = (c + 1 < CPT) ? bzi[(c + 1 < CPT) ? c + 1 : c] : bia;
                const float bou = (c + 1 < CPT) ? bzo[(c + 1 < CPT) ? c + 1 : c] : boa;
                const float w = w2[c];
                const float limz = (w > 0.f) ? min3(1.f, biu, bzo[c]) : min3(1.f, bzi[c], bou);
                fzl[c] = limz * fz[c];
            }
            row(buf, R_FLHI, tz) = fzl[CPT - 1];
            __syncthreads();
            const float fzlb = row(buf, R_FLHI, tz - 1);
            const bool store_col = col && (tx >= XH) && (tx < TW - XH) && (i >= 1) && (i <= nx - 2);
            float* __restrict__ op = gout + g.at(col ? i : 0, k0, Po);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const float fxr = lane_hi(fxl[c]);
                const float fzd = (c > 0) ? fzl[(c > 0) ? c - 1 : c] : fzlb;
                if (store_col && k0 + c < nz) {
                    const float q = a[c][1];
                    float val = q - ((fxr - fxl[c]) + (fyl[c] - fylp[c])) * rj[c];
                    val = val - (fzl[c] - fzd) * rjz[c];
                    op[c * g.ld] = val;
                }
            }
        }
        // ---------------- roll the y state forward
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            fylp[c] = fyl[c];
            v2a[c] = v2b[c];
            fya[c] = fyb[c];
            byi[c] = byi1[c]; byo[c] = byo1[c];
            sx[c][0] = sx[c][1]; sx[c][1] = sx[c][2];
            sz[c][0] = sz[c][1]; sz[c][1] = sz[c][2];
            sx[c][2] = a[c][3] + lane_lo(a[c][3]);
            sz[c][2] = a[c][3] + ((c + 1 < CPT) ? a[(c + 1 < CPT) ? c + 1 : c][3] : qab[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same kernel for CPT = 2 with the two cells of a thread carried as one f32x2 pair: sm_100 has
// packed FADD2 / FMUL2 / FFMA2 (add / sub / mul / fma .f32x2), one issue slot for both cells.  The kernel
// is issue-bound, and 166 of the 489 instructions of one plane iteration are
// FADD / FMUL / FFMA that come in such pairs.  Comparisons, selects, min/max, MUFU.RCP and shuffles
// stay per component.  Each f32x2 operation rounds its two halves like the scalar instruction, so the
// results are those of k_corrector_fused<2, NZP, 2> up to where the compiler contracts a*b+c there.
// The plane loop is unrolled by four: the register windows (3 and 4 planes deep) then rotate by
// renaming instead of 42 moves per iteration.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f2 lane_lo2(f2 v) { return f2{lane_lo(v.x), lane_lo(v.y)}; }
__device__ __forceinline__ f2 lane_hi2(f2 v) { return f2{lane_hi(v.x), lane_hi(v.y)}; }
__device__ __forceinline__ f2 donor2(f2 U, f2 l, f2 r) { return U * f2{U.x > 0.f ? l.x : r.x, U.y > 0.f ? l.y : r.y}; }
__device__ __forceinline__ f2 pos2(f2 v) { return f2{fmaxf(0.f, v.x), fmaxf(0.f, v.y)}; }     // max(0, v)
__device__ __forceinline__ f2 npos2(f2 v) { return f2{fminf(0.f, v.x), fminf(0.f, v.y)}; }    // min(0, v)

template <int NZP>
__global__ void __launch_bounds__(TW * (NZP / 2))
k_corrector_fused_p2(const Geom g, const OutSet outs, const __grid_constant__ FusedMaps maps, const int chunk) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int nzt = NZP / 2;
    const int tx = threadIdx.x, tz = threadIdx.y;
    const int scalar = blockIdx.x % outs.n, xt = blockIdx.x / outs.n;
    float* __restrict__ gout = outs.out[scalar];
    const int nx = g.nx, nz = g.nz, ny = g.ny;
    const int x0 = xt * TOUT - 4;
    const int i = x0 + tx;
    const bool col = (i >= 0 && i < nx);
    const bool ilast = (i == nx - 1);
    const f2 xedge = splat((i == 0 || ilast) ? 0.f : 1.f);
    const int ja = 1 + blockIdx.y * chunk;
    const int jb = min(ja + chunk, ny - 1);
    const int k0 = tz * 2;
    constexpr int box = NZP * TW;
    const f2 eps = splat(EPS), e15 = splat(E15);

    enum { R_QLO, R_QHI, R_MHLO, R_MHHI, R_MLLO, R_MLHI, R_FZHI, R_BILO, R_BOLO, R_FLHI, R_COUNT };
    float* rows = reinterpret_cast<float*>(smraw);
    constexpr int rows_floats = 2 * R_COUNT * (nzt + 2) * TW;
    float* stage0 = rows + rows_floats;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage0 + 2 * NF * box);
    auto row = [&](int buf, int kind, int z) -> float& { return rows[((buf * R_COUNT + kind) * (nzt + 2) + (z + 1)) * TW + tx]; };
    const int tid = tz * TW + tx;
    const int niter = jb - ja + 4;
    const uint32_t stage_bytes = (uint32_t)(NF * nz * TW * sizeof(float));

    auto issue = [&](int it) {
        const int P = ja - 2 + it, st = it & 1;
        float* dst = stage0 + st * NF * box;
        mbar_arrive_expect_tx(&mbar[st], stage_bytes);
        tma_load_box(dst + F_Q2 * box, &maps.q2[scalar], &mbar[st], x0, 0, P);
        tma_load_box(dst + F_L * box, &maps.l[scalar], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 0; f < 3; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 3; f < 11; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P - 2);
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    for (int t = tid; t < 2 * NF * box; t += TW * nzt) stage0[t] = 0.f;
    if (tz == 0) {
        for (int b = 0; b < 2; ++b) {
            row(b, R_QHI, -1) = 0.f; row(b, R_MHHI, -1) = -BIG; row(b, R_MLHI, -1) = BIG;
            row(b, R_FZHI, -1) = 0.f; row(b, R_FLHI, -1) = 0.f;
            row(b, R_QLO, nzt) = 0.f; row(b, R_MHLO, nzt) = -BIG; row(b, R_MLLO, nzt) = BIG;
            row(b, R_BILO, nzt) = 0.f; row(b, R_BOLO, nzt) = 0.f;
        }
    }
    __syncthreads();
    if (tid == 0) issue(0);

    // per-pair register state; .x = level k0, .y = level k0 + 1 (the windows are those of k_corrector_fused)
    f2 a0 = splat(0.f), a1 = a0, a2 = a0, a3 = a0;                  // q2 at planes P-3, P-2, P-1, P
    f2 sx0 = a0, sx1 = a0, sx2 = a0;                                // x-face pair sums at planes P-3, P-2, P-1
    f2 sz0 = a0, sz1 = a0, sz2 = a0;                                // z-face pair sums
    f2 mh0 = splat(-BIG), mh1 = mh0, mh2 = mh0;                     // max(q2, l) at planes P-2, P-1, P
    f2 ml0 = splat(BIG), ml1 = ml0, ml2 = ml0;                      // min(q2, l)
    f2 v2a = a0, fya = a0, byi = a0, byo = a0, fylp = a0;
    float qab0 = 0.f, qab1 = 0.f, qab2 = 0.f, qab3 = 0.f;           // q2 of the row above (k0+2) at planes P-3..P
    float qbl0 = 0.f, qbl1 = 0.f, qbl2 = 0.f, qbl3 = 0.f;           // q2 of the row below (k0-1)
    const int cell_off = k0 * TW + tx;
    const bool lev0 = col && (k0 < nz), lev1 = col && (k0 + 1 < nz);

#pragma unroll 4
    for (int it = 0; it < niter; ++it) {
        const int P = ja - 2 + it;
        const int buf = it & 1;
        const bool pin = (P >= 0 && P < ny);
        const int Po = P - 2;
        const bool out_iter = (Po >= ja);
        if (tid == 0 && it + 1 < niter) issue(it + 1);
        while (!mbar_try_wait(&mbar[buf], (uint32_t)((it >> 1) & 1))) {}
        const float* __restrict__ sg = stage0 + buf * NF * box + cell_off;
        auto ld2 = [&](int f) { return f2{sg[f * box], sg[f * box + TW]}; };

        // ---------------- shift the windows, take plane P from the stage
        a0 = a1; a1 = a2; a2 = a3;
        mh0 = mh1; mh1 = mh2;
        ml0 = ml1; ml1 = ml2;
        const f2 q2v = ld2(F_Q2), lv = ld2(F_L);
        const f2 cay = ld2(F_AY), cby = ld2(F_BY), ccy = ld2(F_CY);
        a3 = q2v;
        mh2 = f2{(pin && lev0) ? fmaxf(q2v.x, lv.x) : -BIG, (pin && lev1) ? fmaxf(q2v.y, lv.y) : -BIG};
        ml2 = f2{(pin && lev0) ? fminf(q2v.x, lv.x) : BIG, (pin && lev1) ? fminf(q2v.y, lv.y) : BIG};
        qab0 = qab1; qab1 = qab2; qab2 = qab3;
        qbl0 = qbl1; qbl1 = qbl2; qbl2 = qbl3;

        // ---------------- in-plane pseudo-velocities of plane Po = P-2
        f2 u2 = splat(0.f), w2 = u2, fx = u2, fz = u2, rj = u2, rjz = u2;
        if (out_iter) {
            const float sx_ab = qab1 + lane_lo(qab1);
            const float sx_bl = qbl1 + lane_lo(qbl1);
            const f2 cax = ld2(F_AX), cbx = ld2(F_BX), ccx = ld2(F_CX), caz = ld2(F_AZ), cbz = ld2(F_BZ), ccz = ld2(F_CZ);
            rj = ld2(F_RJ);
            rjz = ld2(F_RJZ);
            const f2 q = a1;
            const f2 ql = lane_lo2(q);
            const f2 qu = f2{a1.y, qab1};
            // x-face (i-1 | i)
            const f2 Ax = div2(q - ql, sx1 + eps);
            const f2 Bx = div2(sx2 - sx0, sx2 + sx0 + eps);
            const f2 sxu = f2{sx1.y, sx_ab}, sxd = f2{sx_bl, sx1.x};
            const f2 Cx = div2(sxu - sxd, sxu + sxd + eps);
            u2 = fma2(neg2(ccx), Cx, fma2(neg2(cbx), Bx, cax * Ax));
            // z-face above k
            const f2 Az = div2(qu - q, sz1 + eps);
            const f2 szr = lane_hi2(sz1), szl = lane_lo2(sz1);
            const f2 Bz = div2(szr - szl, szr + szl + eps);
            const f2 Cz = div2(sz2 - sz0, sz2 + sz0 + eps);
            w2 = fma2(neg2(ccz), Cz, fma2(neg2(cbz), Bz, caz * Az));
            fx = donor2(u2, ql, q);
            fz = donor2(w2, q, qu);
        }
        // ---------------- publish rows for the z-neighbour threads
        row(buf, R_QLO, tz) = a3.x;
        row(buf, R_QHI, tz) = a3.y;
        if (out_iter) {
            row(buf, R_MHLO, tz) = mh0.x; row(buf, R_MHHI, tz) = mh0.y;
            row(buf, R_MLLO, tz) = ml0.x; row(buf, R_MLHI, tz) = ml0.y;
            row(buf, R_FZHI, tz) = fz.y;
        }
        __syncthreads();            // (no read of the TMA stage below this line)
        qab3 = row(buf, R_QLO, tz + 1);
        qbl3 = row(buf, R_QHI, tz - 1);

        // ---------------- y-face (P-1 | P): unlimited v2, then beta_y of plane P-1
        const f2 syv = a3 + a2;
        const float sy_ab = qab3 + qab2, sy_bl = qbl3 + qbl2;
        const f2 syr = lane_hi2(syv), syl = lane_lo2(syv);
        const f2 Ay = div2(a3 - a2, syv + eps);
        const f2 By = div2(syr - syl, syr + syl + eps);
        const f2 syu = f2{syv.y, sy_ab}, syd = f2{sy_bl, syv.x};
        const f2 Cy = div2(syu - syd, syu + syd + eps);
        const f2 v2b = fma2(neg2(ccy), Cy, fma2(neg2(cby), By, cay * Ay));
        const f2 fyb = donor2(v2b, a2, a3);                    // face (P-1 | P)
        // beta_y of cell plane Pc = P-1
        f2 byi1, byo1;
        {
            const int Pc = P - 1;
            const bool ylast = (Pc == ny - 1);
            const f2 yedge = splat((Pc == 0 || ylast) ? 0.f : 1.f);
            const f2 q = a2;
            const f2 hi = f2{max3(mh0.x, ylast ? q.x : mh1.x, mh2.x), max3(mh0.y, ylast ? q.y : mh1.y, mh2.y)};
            const f2 lo = f2{min3(ml0.x, ylast ? q.x : ml1.x, ml2.x), min3(ml0.y, ylast ? q.y : ml1.y, ml2.y)};
            const f2 fin = (pos2(fya) - npos2(fyb)) * yedge;
            const f2 fout = (pos2(fyb) - npos2(fya)) * yedge;
            byi1 = div2(hi - q, fin + e15);
            byo1 = div2(q - lo, fout + e15);
        }
        // limited y flux through face (P-2 | P-1)
        const f2 limy = f2{(v2a.x > 0.f) ? min3(1.f, byi1.x, byo.x) : min3(1.f, byi.x, byo1.x),
                           (v2a.y > 0.f) ? min3(1.f, byi1.y, byo.y) : min3(1.f, byi.y, byo1.y)};
        const f2 fyl = limy * fya;

        // ---------------- x / z limiter and corrector update of plane Po
        if (out_iter) {
            const float mhb = row(buf, R_MHHI, tz - 1), mlb = row(buf, R_MLHI, tz - 1);
            const float mha = row(buf, R_MHLO, tz + 1), mla = row(buf, R_MLLO, tz + 1);
            const float fzb = row(buf, R_FZHI, tz - 1);
            const f2 q = a1;
            // x direction
            const f2 mhl = lane_lo2(mh0), mhr = lane_hi2(mh0), mll = lane_lo2(ml0), mlr = lane_hi2(ml0);
            const f2 fxp = lane_hi2(fx);
            const f2 hix = f2{max3(mhl.x, ilast ? q.x : mh0.x, mhr.x), max3(mhl.y, ilast ? q.y : mh0.y, mhr.y)};
            const f2 lox = f2{min3(mll.x, ilast ? q.x : ml0.x, mlr.x), min3(mll.y, ilast ? q.y : ml0.y, mlr.y)};
            const f2 finx = (pos2(fx) - npos2(fxp)) * xedge;
            const f2 foutx = (pos2(fxp) - npos2(fx)) * xedge;
            const f2 bxi = div2(hix - q, finx + e15);
            const f2 bxo = div2(q - lox, foutx + e15);
            // z direction (flux_is_w rules at the line ends, FCT core :75-77, :92-94)
            const bool zlast0 = (k0 == nz - 1), zlast1 = (k0 + 1 == nz - 1);
            const f2 fzm = f2{fzb, fz.x};
            const f2 fzp = f2{zlast0 ? fzm.x : fz.x, zlast1 ? fzm.y : fz.y};
            const f2 hiz = f2{max3(mhb, zlast0 ? q.x : mh0.x, mh0.y), max3(mh0.x, zlast1 ? q.y : mh0.y, mha)};
            const f2 loz = f2{min3(mlb, zlast0 ? q.x : ml0.x, ml0.y), min3(ml0.x, zlast1 ? q.y : ml0.y, mla)};
            const f2 finz = pos2(fzm) - npos2(fzp);
            const f2 foutz = pos2(fzp) - npos2(fzm);
            const f2 bzi = div2(hiz - q, finz + e15);
            const f2 bzo = div2(q - loz, foutz + e15);
            row(buf, R_BILO, tz) = bzi.x;
            row(buf, R_BOLO, tz) = bzo.x;
            __syncthreads();
            const float bia = row(buf, R_BILO, tz + 1), boa = row(buf, R_BOLO, tz + 1);
            // limited x flux through face (i-1 | i)
            const f2 bil = lane_lo2(bxi), bol = lane_lo2(bxo);
            const f2 limx = f2{(u2.x > 0.f) ? min3(1.f, bxi.x, bol.x) : min3(1.f, bil.x, bxo.x),
                               (u2.y > 0.f) ? min3(1.f, bxi.y, bol.y) : min3(1.f, bil.y, bxo.y)};
            const f2 fxl = limx * fx;
            // limited z flux through the face above k
            const f2 limz = f2{(w2.x > 0.f) ? min3(1.f, bzi.y, bzo.x) : min3(1.f, bzi.x, bzo.y),
                               (w2.y > 0.f) ? min3(1.f, bia, bzo.y) : min3(1.f, bzi.y, boa)};
            const f2 fzl = limz * fz;
            row(buf, R_FLHI, tz) = fzl.y;
            __syncthreads();
            const float fzlb = row(buf, R_FLHI, tz - 1);
            const bool store_col = col && (tx >= XH) && (tx < TW - XH) && (i >= 1) && (i <= nx - 2);
            float* __restrict__ op = gout + g.at(col ? i : 0, k0, Po);
            const f2 fxr = lane_hi2(fxl);
            const f2 fzd = f2{fzlb, fzl.x};
            f2 val = fma2(neg2((fxr - fxl) + (fyl - fylp)), rj, q);
            val = fma2(neg2(fzl - fzd), rjz, val);
            if (store_col && k0 < nz) op[0] = val.x;
            if (store_col && k0 + 1 < nz) op[g.ld] = val.y;
        }
        // ---------------- roll the y state forward
        fylp = fyl;
        v2a = v2b;
        fya = fyb;
        byi = byi1; byo = byo1;
        sx0 = sx1; sx1 = sx2;
        sz0 = sz1; sz1 = sz2;
        sx2 = a3 + lane_lo2(a3);
        sz2 = a3 + f2{a3.y, qab3};
    }
}

// ---------------------------------------------------------------------------------------------
// The packed form for NP pairs per thread (ICAR_CORR_V2=1; NOT the default: it has not run on hardware
// yet).  NP = 2 is what 41..64 levels need (four levels per thread): static count 182 instructions per cell
// and plane against 225 of k_corrector_fused<4, ..>.  NP = 1 exists to check this kernel against
// k_corrector_fused_p2, with which it must agree bit for bit.
// One difference in form, none in value: max(0, f) and min(0, f) of the limiter's in/out flux sums are
// taken as (f + |f|) and (f - |f|) -- exactly twice the value -- on packed pairs (FADD2 takes |x| and
// -|x| operands), the factor 0.5 folded into the edge mask and the "+ 1e-15" into one FFMA2.  (For NP = 1
// that trades 24 FMNMX for packed adds and moves and comes out even: 370 against 369 per plane.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f2 abs2(f2 a) { return f2{fabsf(a.x), fabsf(a.y)}; }
__device__ __forceinline__ f2 twice_pos(f2 f) { return f + abs2(f); }      // 2 max(0, f)
__device__ __forceinline__ f2 twice_neg(f2 f) { return f - abs2(f); }      // 2 min(0, f)

template <int NP, int NZP, int STAGES>
__global__ void __launch_bounds__(TW * (NZP / (2 * NP)))
k_corrector_fused_pk(const Geom g, const OutSet outs, const __grid_constant__ FusedMaps maps, const int chunk) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int CPT = 2 * NP;
    constexpr int nzt = NZP / CPT;
    const int tx = threadIdx.x, tz = threadIdx.y;
    const int scalar = blockIdx.x % outs.n, xt = blockIdx.x / outs.n;
    float* __restrict__ gout = outs.out[scalar];
    const int nx = g.nx, nz = g.nz, ny = g.ny;
    const int x0 = xt * TOUT - 4;
    const int i = x0 + tx;
    const bool col = (i >= 0 && i < nx);
    const bool ilast = (i == nx - 1);
    const f2 xedge_half = splat((i == 0 || ilast) ? 0.f : 0.5f);
    const int ja = 1 + blockIdx.y * chunk;
    const int jb = min(ja + chunk, ny - 1);
    const int k0 = tz * CPT;
    constexpr int box = NZP * TW;
    const f2 eps = splat(EPS), e15 = splat(E15), half = splat(0.5f);

    enum { R_QLO, R_QHI, R_MHLO, R_MHHI, R_MLLO, R_MLHI, R_FZHI, R_BILO, R_BOLO, R_FLHI, R_COUNT };
    float* rows = reinterpret_cast<float*>(smraw);
    constexpr int rows_floats = 2 * R_COUNT * (nzt + 2) * TW;
    float* stage0 = rows + rows_floats;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage0 + STAGES * NF * box);
    auto row = [&](int buf, int kind, int z) -> float& { return rows[((buf * R_COUNT + kind) * (nzt + 2) + (z + 1)) * TW + tx]; };
    const int tid = tz * TW + tx;
    const int niter = jb - ja + 4;
    const uint32_t stage_bytes = (uint32_t)(NF * nz * TW * sizeof(float));

    auto issue = [&](int it) {
        const int P = ja - 2 + it, st = (STAGES == 2) ? (it & 1) : 0;
        float* dst = stage0 + st * NF * box;
        mbar_arrive_expect_tx(&mbar[st], stage_bytes);
        tma_load_box(dst + F_Q2 * box, &maps.q2[scalar], &mbar[st], x0, 0, P);
        tma_load_box(dst + F_L * box, &maps.l[scalar], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 0; f < 3; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P);
#pragma unroll
        for (int f = 3; f < 11; ++f) tma_load_box(dst + (F_AY + f) * box, &maps.coef[f], &mbar[st], x0, 0, P - 2);
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    for (int t = tid; t < STAGES * NF * box; t += TW * nzt) stage0[t] = 0.f;
    if (tz == 0) {
        for (int b = 0; b < 2; ++b) {
            row(b, R_QHI, -1) = 0.f; row(b, R_MHHI, -1) = -BIG; row(b, R_MLHI, -1) = BIG;
            row(b, R_FZHI, -1) = 0.f; row(b, R_FLHI, -1) = 0.f;
            row(b, R_QLO, nzt) = 0.f; row(b, R_MHLO, nzt) = -BIG; row(b, R_MLLO, nzt) = BIG;
            row(b, R_BILO, nzt) = 0.f; row(b, R_BOLO, nzt) = 0.f;
        }
    }
    __syncthreads();
    if (tid == 0) issue(0);

    // pair p of a thread holds levels k0 + 2p (.x) and k0 + 2p + 1 (.y)
    f2 a0[NP], a1[NP], a2[NP], a3[NP];          // q2 at planes P-3, P-2, P-1, P
    f2 sx0[NP], sx1[NP], sx2[NP];               // x-face pair sums at planes P-3, P-2, P-1
    f2 sz0[NP], sz1[NP], sz2[NP];               // z-face pair sums
    f2 mh0[NP], mh1[NP], mh2[NP];               // max(q2, l) at planes P-2, P-1, P
    f2 ml0[NP], ml1[NP], ml2[NP];               // min(q2, l)
    f2 v2a[NP], fya[NP], byi[NP], byo[NP], fylp[NP];
    bool lev[NP][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        a0[p] = a1[p] = a2[p] = a3[p] = splat(0.f);
        sx0[p] = sx1[p] = sx2[p] = sz0[p] = sz1[p] = sz2[p] = splat(0.f);
        mh0[p] = mh1[p] = mh2[p] = splat(-BIG);
        ml0[p] = ml1[p] = ml2[p] = splat(BIG);
        v2a[p] = fya[p] = byi[p] = byo[p] = fylp[p] = splat(0.f);
        lev[p][0] = col && (k0 + 2 * p < nz);
        lev[p][1] = col && (k0 + 2 * p + 1 < nz);
    }
    float qab0 = 0.f, qab1 = 0.f, qab2 = 0.f, qab3 = 0.f;           // q2 of the row above (k0+CPT) at planes P-3..P
    float qbl0 = 0.f, qbl1 = 0.f, qbl2 = 0.f, qbl3 = 0.f;           // q2 of the row below (k0-1)
    const int cell_off = k0 * TW + tx;
    // {v[p].y, v[p+1].x} with `above` after the last pair; {v[p-1].y, v[p].x} with `below` before the first
    auto up_of = [&](const f2 (&v)[NP], int p, float above) { return f2{v[p].y, (p + 1 < NP) ? v[(p + 1 < NP) ? p + 1 : p].x : above}; };
    auto dn_of = [&](const f2 (&v)[NP], int p, float below) { return f2{(p > 0) ? v[(p > 0) ? p - 1 : p].y : below, v[p].x}; };

#pragma unroll(NP == 1 ? 4 : 2)
    for (int it = 0; it < niter; ++it) {
        const int P = ja - 2 + it;
        const int buf = it & 1;
        const bool pin = (P >= 0 && P < ny);
        const int Po = P - 2;
        const bool out_iter = (Po >= ja);
        if (STAGES == 2) { if (tid == 0 && it + 1 < niter) issue(it + 1); }
        else             { if (tid == 0 && it > 0) issue(it); }
        const int st = (STAGES == 2) ? buf : 0;
        while (!mbar_try_wait(&mbar[st], (uint32_t)((STAGES == 2) ? ((it >> 1) & 1) : (it & 1)))) {}
        const float* __restrict__ sg = stage0 + st * NF * box + cell_off;
        auto ld2 = [&](int f, int p) { return f2{sg[f * box + (2 * p) * TW], sg[f * box + (2 * p + 1) * TW]}; };

        // ---------------- shift the windows, take plane P from the stage
        f2 cay[NP], cby[NP], ccy[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            a0[p] = a1[p]; a1[p] = a2[p]; a2[p] = a3[p];
            mh0[p] = mh1[p]; mh1[p] = mh2[p];
            ml0[p] = ml1[p]; ml1[p] = ml2[p];
            const f2 q2v = ld2(F_Q2, p), lv = ld2(F_L, p);
            cay[p] = ld2(F_AY, p); cby[p] = ld2(F_BY, p); ccy[p] = ld2(F_CY, p);
            a3[p] = q2v;
            mh2[p] = f2{(pin && lev[p][0]) ? fmaxf(q2v.x, lv.x) : -BIG, (pin && lev[p][1]) ? fmaxf(q2v.y, lv.y) : -BIG};
            ml2[p] = f2{(pin && lev[p][0]) ? fminf(q2v.x, lv.x) : BIG, (pin && lev[p][1]) ? fminf(q2v.y, lv.y) : BIG};
        }
        qab0 = qab1; qab1 = qab2; qab2 = qab3;
        qbl0 = qbl1; qbl1 = qbl2; qbl2 = qbl3;

        // ---------------- in-plane pseudo-velocities of plane Po = P-2
        f2 u2[NP], w2[NP], fx[NP], fz[NP], rj[NP], rjz[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) u2[p] = w2[p] = fx[p] = fz[p] = rj[p] = rjz[p] = splat(0.f);
        if (out_iter) {
            const float sx_ab = qab1 + lane_lo(qab1);
            const float sx_bl = qbl1 + lane_lo(qbl1);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const f2 cax = ld2(F_AX, p), cbx = ld2(F_BX, p), ccx = ld2(F_CX, p), caz = ld2(F_AZ, p),
                         cbz = ld2(F_BZ, p), ccz = ld2(F_CZ, p);
                rj[p] = ld2(F_RJ, p);
                rjz[p] = ld2(F_RJZ, p);
                const f2 q = a1[p];
                const f2 ql = lane_lo2(q);
                const f2 qu = up_of(a1, p, qab1);
                // x-face (i-1 | i)
                const f2 Ax = div2(q - ql, sx1[p] + eps);
                const f2 Bx = div2(sx2[p] - sx0[p], sx2[p] + sx0[p] + eps);
                const f2 sxu = up_of(sx1, p, sx_ab), sxd = dn_of(sx1, p, sx_bl);
                const f2 Cx = div2(sxu - sxd, sxu + sxd + eps);
                u2[p] = fma2(neg2(ccx), Cx, fma2(neg2(cbx), Bx, cax * Ax));
                // z-face above k
                const f2 Az = div2(qu - q, sz1[p] + eps);
                const f2 szr = lane_hi2(sz1[p]), szl = lane_lo2(sz1[p]);
                const f2 Bz = div2(szr - szl, szr + szl + eps);
                const f2 Cz = div2(sz2[p] - sz0[p], sz2[p] + sz0[p] + eps);
                w2[p] = fma2(neg2(ccz), Cz, fma2(neg2(cbz), Bz, caz * Az));
                fx[p] = donor2(u2[p], ql, q);
                fz[p] = donor2(w2[p], q, qu);
            }
        }
        // ---------------- publish rows for the z-neighbour threads
        row(buf, R_QLO, tz) = a3[0].x;
        row(buf, R_QHI, tz) = a3[NP - 1].y;
        if (out_iter) {
            row(buf, R_MHLO, tz) = mh0[0].x; row(buf, R_MHHI, tz) = mh0[NP - 1].y;
            row(buf, R_MLLO, tz) = ml0[0].x; row(buf, R_MLHI, tz) = ml0[NP - 1].y;
            row(buf, R_FZHI, tz) = fz[NP - 1].y;
        }
        __syncthreads();            // (no read of the TMA stage below this line)
        qab3 = row(buf, R_QLO, tz + 1);
        qbl3 = row(buf, R_QHI, tz - 1);

        // ---------------- y-face (P-1 | P): unlimited v2, then beta_y of plane P-1, then the limited flux
        f2 v2b[NP], fyb[NP], syv[NP], byi1[NP], byo1[NP], fyl[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) syv[p] = a3[p] + a2[p];
        const float sy_ab = qab3 + qab2, sy_bl = qbl3 + qbl2;
        const int Pc = P - 1;
        const bool ylast = (Pc == ny - 1);
        const f2 yedge_half = splat((Pc == 0 || ylast) ? 0.f : 0.5f);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const f2 syr = lane_hi2(syv[p]), syl = lane_lo2(syv[p]);
            const f2 Ay = div2(a3[p] - a2[p], syv[p] + eps);
            const f2 By = div2(syr - syl, syr + syl + eps);
            const f2 syu = up_of(syv, p, sy_ab), syd = dn_of(syv, p, sy_bl);
            const f2 Cy = div2(syu - syd, syu + syd + eps);
            v2b[p] = fma2(neg2(ccy[p]), Cy, fma2(neg2(cby[p]), By, cay[p] * Ay));
            fyb[p] = donor2(v2b[p], a2[p], a3[p]);                    // face (P-1 | P)
            // beta_y of cell plane Pc = P-1
            const f2 q = a2[p];
            const f2 hi = f2{max3(mh0[p].x, ylast ? q.x : mh1[p].x, mh2[p].x), max3(mh0[p].y, ylast ? q.y : mh1[p].y, mh2[p].y)};
            const f2 lo = f2{min3(ml0[p].x, ylast ? q.x : ml1[p].x, ml2[p].x), min3(ml0[p].y, ylast ? q.y : ml1[p].y, ml2[p].y)};
            const f2 fin2 = twice_pos(fya[p]) - twice_neg(fyb[p]);
            const f2 fout2 = twice_pos(fyb[p]) - twice_neg(fya[p]);
            byi1[p] = div2(hi - q, fma2(fin2, yedge_half, e15));
            byo1[p] = div2(q - lo, fma2(fout2, yedge_half, e15));
            // limited y flux through face (P-2 | P-1)
            const f2 limy = f2{(v2a[p].x > 0.f) ? min3(1.f, byi1[p].x, byo[p].x) : min3(1.f, byi[p].x, byo1[p].x),
                               (v2a[p].y > 0.f) ? min3(1.f, byi1[p].y, byo[p].y) : min3(1.f, byi[p].y, byo1[p].y)};
            fyl[p] = limy * fya[p];
        }

        // ---------------- x / z limiter and corrector update of plane Po
        if (out_iter) {
            const float mhb = row(buf, R_MHHI, tz - 1), mlb = row(buf, R_MLHI, tz - 1);
            const float mha = row(buf, R_MHLO, tz + 1), mla = row(buf, R_MLLO, tz + 1);
            const float fzb = row(buf, R_FZHI, tz - 1);
            f2 bxi[NP], bxo[NP], bzi[NP], bzo[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const f2 q = a1[p];
                // x direction
                const f2 mhl = lane_lo2(mh0[p]), mhr = lane_hi2(mh0[p]), mll = lane_lo2(ml0[p]), mlr = lane_hi2(ml0[p]);
                const f2 fxp = lane_hi2(fx[p]);
                const f2 hix = f2{max3(mhl.x, ilast ? q.x : mh0[p].x, mhr.x), max3(mhl.y, ilast ? q.y : mh0[p].y, mhr.y)};
                const f2 lox = f2{min3(mll.x, ilast ? q.x : ml0[p].x, mlr.x), min3(mll.y, ilast ? q.y : ml0[p].y, mlr.y)};
                const f2 finx2 = twice_pos(fx[p]) - twice_neg(fxp);
                const f2 foutx2 = twice_pos(fxp) - twice_neg(fx[p]);
                bxi[p] = div2(hix - q, fma2(finx2, xedge_half, e15));
                bxo[p] = div2(q - lox, fma2(foutx2, xedge_half, e15));
                // z direction (flux_is_w rules at the line ends, FCT core :75-77, :92-94)
                const bool zlast0 = (k0 + 2 * p == nz - 1), zlast1 = (k0 + 2 * p + 1 == nz - 1);
                const f2 mhd = dn_of(mh0, p, mhb), mhu = up_of(mh0, p, mha);
                const f2 mld = dn_of(ml0, p, mlb), mlu = up_of(ml0, p, mla);
                const f2 fzm = dn_of(fz, p, fzb);
                const f2 fzp = f2{zlast0 ? fzm.x : fz[p].x, zlast1 ? fzm.y : fz[p].y};
                const f2 hiz = f2{max3(mhd.x, zlast0 ? q.x : mh0[p].x, mhu.x), max3(mhd.y, zlast1 ? q.y : mh0[p].y, mhu.y)};
                const f2 loz = f2{min3(mld.x, zlast0 ? q.x : ml0[p].x, mlu.x), min3(mld.y, zlast1 ? q.y : ml0[p].y, mlu.y)};
                const f2 finz2 = twice_pos(fzm) - twice_neg(fzp);
                const f2 foutz2 = twice_pos(fzp) - twice_neg(fzm);
                bzi[p] = div2(hiz - q, fma2(finz2, half, e15));
                bzo[p] = div2(q - loz, fma2(foutz2, half, e15));
            }
            row(buf, R_BILO, tz) = bzi[0].x;
            row(buf, R_BOLO, tz) = bzo[0].x;
            __syncthreads();
            const float bia = row(buf, R_BILO, tz + 1), boa = row(buf, R_BOLO, tz + 1);
            f2 fxl[NP], fzl[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                // limited x flux through face (i-1 | i)
                const f2 bil = lane_lo2(bxi[p]), bol = lane_lo2(bxo[p]);
                const f2 limx = f2{(u2[p].x > 0.f) ? min3(1.f, bxi[p].x, bol.x) : min3(1.f, bil.x, bxo[p].x),
                                   (u2[p].y > 0.f) ? min3(1.f, bxi[p].y, bol.y) : min3(1.f, bil.y, bxo[p].y)};
                fxl[p] = limx * fx[p];
                // limited z flux through the face above k
                const f2 biu = up_of(bzi, p, bia), bou = up_of(bzo, p, boa);
                const f2 limz = f2{(w2[p].x > 0.f) ? min3(1.f, biu.x, bzo[p].x) : min3(1.f, bzi[p].x, bou.x),
                                   (w2[p].y > 0.f) ? min3(1.f, biu.y, bzo[p].y) : min3(1.f, bzi[p].y, bou.y)};
                fzl[p] = limz * fz[p];
            }
            row(buf, R_FLHI, tz) = fzl[NP - 1].y;
            __syncthreads();
            const float fzlb = row(buf, R_FLHI, tz - 1);
            const bool store_col = col && (tx >= XH) && (tx < TW - XH) && (i >= 1) && (i <= nx - 2);
            float* __restrict__ op = gout + g.at(col ? i : 0, k0, Po);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const f2 fxr = lane_hi2(fxl[p]);
                const f2 fzd = dn_of(fzl, p, fzlb);
                f2 val = fma2(neg2((fxr - fxl[p]) + (fyl[p] - fylp[p])), rj[p], a1[p]);
                val = fma2(neg2(fzl[p] - fzd), rjz[p], val);
                if (store_col && k0 + 2 * p < nz) op[(2 * p) * g.ld] = val.x;
                if (store_col && k0 + 2 * p + 1 < nz) op[(2 * p + 1) * g.ld] = val.y;
            }
        }
        // ---------------- roll the y state forward
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            fylp[p] = fyl[p];
            v2a[p] = v2b[p];
            fya[p] = fyb[p];
            byi[p] = byi1[p]; byo[p] = byo1[p];
            sx0[p] = sx1[p]; sx1[p] = sx2[p];
            sz0[p] = sz1[p]; sz1[p] = sz2[p];
            sx2[p] = a3[p] + lane_lo2(a3[p]);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) sz2[p] = a3[p] + up_of(a3, p, qab3);
    }
}

// ---- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* p; int nx, nz, ny, ld;
    bool operator==(const MapKey& o) const { return p == o.p && nx == o.nx && nz == o.nz && ny == o.ny && ld == o.ld; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const { return std::hash<const void*>()(k.p) ^ (size_t)k.nx * 1315423911u ^ (size_t)k.ny * 2654435761u; }
};

// a (nx, nz, ny) fp32 field with row pitch ld, read in boxes of 32 x nz x 1
bool field_map(const Geom& g, const float* ptr, CUtensorMap* out) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, g.nx, g.nz, g.ny, g.ld};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.nz, (cuuint64_t)g.ny};
    cuuint64_t gstr[2] = {(cuuint64_t)g.ld * sizeof(float), (cuuint64_t)g.ld * g.nz * sizeof(float)};
    cuuint32_t bx[3] = {(cuuint32_t)TW, (cuuint32_t)g.nz, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}

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
    if (g.nz <= 20) { p.cpt = 2; p.nzp = 20; }
    else if (g.nz <= 40) { p.cpt = 2; p.nzp = 40; }
    else if (g.nz <= 48) { p.cpt = 4; p.nzp = 48; }
    else { p.cpt = 4; p.nzp = 64; p.stages = 1; }
    if (env_int("ICAR_FUSED_CPT", 0) == 4 && g.nz <= 48) { p.cpt = 4; p.nzp = 48; }
    // opt-in until it has run on hardware: 49..52 levels (BASELINE configs[3] has 50) still fit TWO stages
    // (2 x 13 x 52 x 128 B + rows = 211 KB), so the TMA loads stay one plane ahead instead of being exposed
    if (env_int("ICAR_FUSED_NZP52", 0) == 1 && g.nz > 48 && g.nz <= 52) { p.cpt = 4; p.nzp = 52; p.stages = 2; }
    const int nzt = p.nzp / p.cpt;
    p.smem = (size_t)(2 * 10 * (nzt + 2) * TW + p.stages * NF * p.nzp * TW) * sizeof(float) + 16;
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
#define LAUNCH_PK(NPAIRS, Z, S)                                                                               \
    {                                                                                                         \
        auto kern = k_corrector_fused_pk<NPAIRS, Z, S>;                                                       \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) !=        \
            cudaSuccess) { (void)cudaGetLastError(); return -1; }                                             \
        kern<<<grid, block, plan.smem, st>>>(g, outs, maps, chunk);                                           \
    }
    if (v2 && plan.cpt == 2 && plan.nzp == 20) LAUNCH_PK(1, 20, 2)
    else if (v2 && plan.cpt == 2 && plan.nzp == 40) LAUNCH_PK(1, 40, 2)
    else if (v2 && plan.nzp == 48) LAUNCH_PK(2, 48, 2)
    else if (v2 && plan.nzp == 52) LAUNCH_PK(2, 52, 2)
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
