// "Simple" (SB04) cloud microphysics on sm_100a.
//
// Reference: src/physics/mp_simple.f90 -- mp_simple_driver :595-646, mp_simple :481-566,
// mp_conversions :381-420, cloud_conversion :198-280, cloud2hydrometeor :295-315,
// phase_change :333-362, sediment :437-459, sat_mr :146-182.
//
// Parallelisation.  The reference walks one column at a time; but inside a column every
// level is independent in mp_conversions and in the re-evaporation loops, and one
// sedimentation sub-step is level-parallel too once written per level:
//     q_new(k) = (q_old(k) - out(k)) + in(k),   in(k) = fall*q_old(k+1)*rho(k+1)/(rho(k)*dz(k))
// (same operation order as the Fortran's two sequential loops, :449-457).  So a CTA owns a
// tile of 32 columns x all levels with ONE THREAD PER CELL (CPT cells per thread when
// nz > 32): a warp is 32 consecutive i at one level, so every global access is one
// coalesced 128-byte line, the whole column state lives in registers from the first load
// to the last store (each field is read once and written once: the 14 field-passes of
// SURVEY.md section 8d), and only q_old(k+1) and three column reductions go through shared
// memory.
//
// Numerics: `exp` of the Fortran is glibc's expf on the reference platform; exp_glibc()
// below evaluates the same algorithm (same table, same double-precision polynomial), so
// together with IEEE division and -fmad=false the kernel reproduces the CPU oracle bit for
// bit rather than to a tolerance.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "mp_simple.cuh"

namespace icar {

namespace {

constexpr float LH_vapor = 2.26E6f;          // mp_simple.f90:64
constexpr float dLHvdt = 2400.f;             // :65
constexpr float LH_liquid = 3.34E5f;         // :69
constexpr float heat_capacity = 1006.0f;     // :70
constexpr float SMALL_VALUE = 1E-30f;        // :71
constexpr float freezing_threshold = 273.15f;
constexpr float snow_fall_rate = 1.5f;
constexpr float rain_fall_rate = 10.0f;
constexpr float cloud_init = 0.0001f;        // snow_cloud_init == rain_cloud_init
#ifndef ICAR_MP_TX
#define ICAR_MP_TX 32
#endif
constexpr int TX = ICAR_MP_TX;               // columns per CTA (one warp = 32/TX level rows)

// 2^(i/32) table of glibc's expf (sysdeps/ieee754/flt-32/e_exp2f_data.c, EXP2F_TABLE_BITS=5):
// tab[i] = asuint64(2^(i/32)) - (i << 47).  Verified against powl() and, end to end, against
// glibc 2.39 expf on 3.2e8 arguments.
__device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// glibc expf (sysdeps/ieee754/flt-32/e_expf.c): x*N/ln2 = k + r, 2^(k/N) from the table,
// 2^(r/N) from a cubic, all in double, rounded to float once.
__device__ __forceinline__ float exp_glibc(float x, const unsigned long long* __restrict__ tab) {
    if (!(fabsf(x) < 87.0f)) return expf(x);           // overflow / underflow tails: not on this path
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N,
                 C2 = 0x1.62e42ff0c52d6p-1 / N;
    double z = InvLn2N * (double)x;
    double kd = z + SHIFT;
    unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd -= SHIFT;
    double r = z - kd;
    unsigned long long t = tab[ki & 31ULL] + (ki << 47);
    double s = __longlong_as_double((long long)t);
    double p = C0 * r + C1;
    double r2 = r * r;
    double y = C2 * r + 1.0;
    y = p * r2 + y;
    return (float)(y * s);
}

struct Ctx {
    const unsigned long long* tab;
    unsigned* flags;          // per-thread diagnostic bits, flushed with one atomic per set bit at the end
    float c2r, c2s;
};

__device__ __forceinline__ float sat_mr(float t, float p, const Ctx& cx) {            // :146-182
    float a, b;
    if (t < freezing_threshold) { a = 21.8745584f; b = 7.66f; }
    else                        { a = 17.2693882f; b = 35.86f; }
    float e_s = 610.78f * exp_glibc(a * (t - 273.16f) / (t - b), cx.tab);
    if ((p - e_s) <= 0.f) e_s = p * 0.99999f;
    return 0.6219907f * e_s / (p - e_s);
}

__device__ __forceinline__ void phase_change(float& t, float& q1, float qmax, float& q2,
                                             float Lheat, float change_rate, const Ctx& cx) {  // :333-362
    float mass2temp = Lheat / heat_capacity;
    float delta = (qmax - q2) * change_rate;
    if (delta > q1) delta = q1;
    if (delta > ((qmax - q2) * 0.99f)) delta = (qmax - q2) * 0.99f;
    q1 = q1 - delta;
    if (q1 < 0.f) {
        if ((q1 + SMALL_VALUE) < 0.f) q1 = 0.f;
        else *cx.flags |= 4u;                      // reference: print + stop
    }
    q2 = q2 + delta;
    t = t + delta * mass2temp;
}

__device__ __forceinline__ void cloud2hydrometeor(float& qc, float& q, float conversion) {  // :295-315
    float delta = (qc > cloud_init) ? qc - (qc * conversion) : 0.f;
    if (delta < qc) { qc = qc - delta; q = q + delta; }
    else            { q = q + qc; qc = 0.f; }
    qc = fmaxf(qc, 0.f);
}

__device__ __forceinline__ void mp_conversions(float p, float& t, float& qv, float& qc, float& qr,
                                               float& qs, const Ctx& cx) {                  // :381-420
    const float L_melt = -1.f * LH_liquid;
    const float L_evap = -1.f * (LH_vapor + (373.15f - t) * dLHvdt);
    const float L_subl = L_melt + L_evap;
    float qvsat = 0.f;
    {   // cloud_conversion :198-280
        const float maxerr = 1e-4f;
        int iteration = 0;
        float lastqv = qv + maxerr * 2.f;
        const float vapor2temp = (LH_vapor + (373.15f - t) * dLHvdt) / heat_capacity;
        const float pre_qc = qc, pre_t = t;
        while ((fabsf(lastqv - qv) > maxerr) && (iteration < 15)) {
            iteration = iteration + 1;
            lastqv = qv;
            qvsat = sat_mr(t, p, cx);
            if (qv > qvsat) {
                float excess = (qv - qvsat) * 0.5f;
                t = t + (excess * vapor2temp);
                qv = qv - excess;
                qc = qc + excess;
            } else if (qc > 0.f) {
                float excess = (qvsat - qv) * 0.5f;
                if (excess < qc) {
                    t = t - (excess * vapor2temp);
                    qv = qv + excess;
                    qc = qc - excess;
                } else {
                    qv = qv + qc;
                    t = t - (qc * vapor2temp);
                    qc = 0.f;
                }
            }
        }
        if (iteration == 15) {
            qv = sat_mr(pre_t, p, cx);
            t = pre_t;
            qc = pre_qc;
            *cx.flags |= 1u;
        }
        qc = fmaxf(qc, 0.f);
        if ((t > 350.f) || (qc > 0.01f) || (qvsat > 1.f)) *cx.flags |= 2u;
    }
    if ((qc + qr + qs) > SMALL_VALUE) {
        if (qc > SMALL_VALUE) {
            if (t > freezing_threshold) {
                cloud2hydrometeor(qc, qr, cx.c2r);
                if (qs > SMALL_VALUE) phase_change(t, qs, 100.f, qr, L_melt, cx.c2r, cx);
            } else {
                cloud2hydrometeor(qc, qs, cx.c2s);
            }
        }
        if (qv < qvsat) {
            if (qr > SMALL_VALUE) phase_change(t, qr, qvsat, qv, L_evap, cx.c2r / 2.f, cx);
            if (qs > SMALL_VALUE) phase_change(t, qs, qvsat, qv, L_subl, cx.c2s / 2.f, cx);
        }
    }
}

// Shared memory: two planes [nzp][TX] for q_old / rho (+1 padding level that reads as 0),
// three column-reduction rows.
template <int CPT>
#ifndef ICAR_MP_MINB
#define ICAR_MP_MINB 3
#endif
__global__ void __launch_bounds__(TX * 10, ICAR_MP_MINB)
k_mp_simple(Geom g, const float* __restrict__ gp, float* __restrict__ gth,
            const float* __restrict__ gpii, const float* __restrict__ grho,
            float* __restrict__ gqv, float* __restrict__ gqc, float* __restrict__ gqr,
            float* __restrict__ gqs, float* __restrict__ rain, float* __restrict__ snow, int ld2,
            float dt, const float* __restrict__ gdz, MpTiles tiles, int kts,
            int kte, float c2r, float c2s, unsigned long long* diag, double* __restrict__ acc_rain,
            double* __restrict__ acc_snow) {
    // blockIdx.z selects one of up to 4 disjoint tiles (the ring strips of process_halo share a launch)
    const int its = tiles.its[blockIdx.z], ite = tiles.ite[blockIdx.z];
    const int jts = tiles.jts[blockIdx.z], jte = tiles.jte[blockIdx.z];
    const bool along_y = tiles.along_y[blockIdx.z] != 0;
    // lanes run along x (coalesced rows) or, for strips a few columns wide, along y
    const int lane0 = (int)blockIdx.x * TX, row = (int)blockIdx.y;
    if (along_y ? (jts + lane0 > jte || its + row > ite) : (its + lane0 > ite || jts + row > jte)) return;
    extern __shared__ float sm[];
    const int nz = g.nz;
    const int nzt = blockDim.y;                 // z-threads; thread owns levels tz + c*nzt
    const int tx = threadIdx.x, tz = threadIdx.y;
    float* s_q = sm;                            // [(nz+1)][TX]
    float* s_rho = s_q + (nz + 1) * TX;         // [(nz+1)][TX]
    float* s_red = s_rho + (nz + 1) * TX;       // [3][TX]: max qr, max qs, max dt/dz
    const int i = along_y ? its + row : its + lane0 + tx;
    const int j = along_y ? jts + lane0 + tx : jts + row;
    const bool col_ok = (i <= ite) && (j <= jte);
    unsigned dflags = 0u;
    const Ctx cx{kExp2fTab, &dflags, c2r, c2s};
    const int top = min(kte, nz - 2);

    float T[CPT], p[CPT], rho[CPT], dz[CPT], pii[CPT], qv[CPT], qc[CPT], qr[CPT], qs[CPT];
    float qvs[CPT], Tq[CPT];                   // cached sat_mr(Tq, p)
    bool live[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int k = tz + c * nzt;
        live[c] = col_ok && (k < nz);
        T[c] = p[c] = rho[c] = pii[c] = 1.f; dz[c] = 1.f; qv[c] = qc[c] = qr[c] = qs[c] = 0.f;
        qvs[c] = 0.f; Tq[c] = -1.f;
        if (live[c]) {
            const size_t a = g.at(i, k, j);
            p[c] = gp[a]; pii[c] = gpii[a]; rho[c] = grho[a]; dz[c] = gdz[a];
            qv[c] = gqv[a]; qc[c] = gqc[a]; qr[c] = gqr[a]; qs[c] = gqs[a];
            T[c] = gth[a] * pii[c];                                       // :633 (whole column)
            if (k >= kts && k <= kte) mp_conversions(p[c], T[c], qv[c], qc[c], qr[c], qs[c], cx);  // :500-502
        }
    }
    // ---- column reductions: maxval(qr), maxval(qs), maxval(dt/dz) over the WHOLE column (:505,:511).
    // The first two are only compared with SMALL_VALUE > 0 and dt/dz is positive, so the maxima of
    // max(x, 0) serve, and for non-negative floats integer order == float order: one shared-memory
    // atomicMax per value replaces a serial walk over the levels.
    int* s_redi = reinterpret_cast<int*>(s_red);
    if (tz == 0) {
        s_q[nz * TX + tx] = 0.f;            // q_old / rho "above the model top"
        s_rho[nz * TX + tx] = 0.f;
        s_redi[tx] = 0; s_redi[TX + tx] = 0; s_redi[2 * TX + tx] = 0;
    }
    __syncthreads();
    {
        float mqr = 0.f, mqs = 0.f, mdz = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int k = tz + c * nzt;
            if (live[c]) { mqr = fmaxf(mqr, qr[c]); mqs = fmaxf(mqs, qs[c]); mdz = fmaxf(mdz, dt / dz[c]); }
            if (k < nz) s_rho[k * TX + tx] = rho[c];
        }
        atomicMax(&s_redi[tx], __float_as_int(mqr));
        atomicMax(&s_redi[TX + tx], __float_as_int(mqs));
        atomicMax(&s_redi[2 * TX + tx], __float_as_int(mdz));
    }
    __syncthreads();
    const float max_qr = s_red[tx], max_qs = s_red[TX + tx], max_dtdz = s_red[2 * TX + tx];
    // rounding is monotone, so maxval(dt/dz*v) == maxval(dt/dz)*v exactly
    int n_r = 0, n_s = 0;
    float fall_r = 0.f, fall_s = 0.f, rate_r = 0.f, rate_s = 0.f;
    if (col_ok && max_qr > SMALL_VALUE) {                                  // :505-513
        float cfl = ceilf(max_dtdz * rain_fall_rate);
        fall_r = dt * rain_fall_rate / cfl;
        n_r = (int)lroundf(cfl);
        rate_r = c2r / (float)(2 * n_r);
    }
    if (col_ok && max_qs > SMALL_VALUE) {                                  // :533-541
        float cfl = ceilf(max_dtdz * snow_fall_rate);
        fall_s = dt * snow_fall_rate / cfl;
        n_s = (int)lroundf(cfl);
        rate_s = c2s / (float)(2 * n_s);
    }
    const int nmax_r = __reduce_max_sync(0xffffffffu, n_r);               // same 32 columns in every warp
    const int nmax_s = __reduce_max_sync(0xffffffffu, n_s);
    const float L_melt = -1.f * LH_liquid;
    // accumulated in the one thread of the column that owns level kts, in the reference's order
    const size_t c2 = (size_t)(col_ok ? i : its) + (size_t)ld2 * (size_t)(col_ok ? j : jts);
    bool owner = false;
#pragma unroll
    for (int c = 0; c < CPT; ++c) owner = owner || (live[c] && (tz + c * nzt) == kts);
    // resident mode (acc_rain != null): the per-call REAL*4 precipitation array of process_subdomain starts
    // at zero (mp_driver.f90:411-413) and is added to the REAL*8 accumulator at the end (:587-595) -- done
    // here by the owning thread, which saves the clear and the separate accumulation pass
    float rain_acc = (owner && !acc_rain) ? rain[c2] : 0.f, snow_acc = (owner && !acc_rain) ? snow[c2] : 0.f;
    // (accumulators are read here, not at the end: a dependent load in a CTA's last instructions is exposed)
    const double acc_r0 = (owner && acc_rain) ? acc_rain[c2] : 0.0, acc_s0 = (owner && acc_rain) ? acc_snow[c2] : 0.0;

    // ---- sedimentation + re-evaporation sub-steps, rain then snow (:507-564, sediment :437-459)
#pragma unroll
    for (int species = 0; species < 2; ++species) {
        const int nmax = species ? nmax_s : nmax_r;
        const int n = species ? n_s : n_r;
        const float fall = species ? fall_s : fall_r;
        const float rate = species ? rate_s : rate_r;
        for (int step = 0; step < nmax; ++step) {
            __syncthreads();                    // previous readers of s_q are done
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int k = tz + c * nzt;
                if (k < nz) s_q[k * TX + tx] = species ? qs[c] : qr[c];
            }
            __syncthreads();
            if (step < n) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int k = tz + c * nzt;
                    if (!live[c] || k < kts || k > max(kte, top + 1)) continue;
                    float& q = species ? qs[c] : qr[c];
                    float nq = q;
                    if (k == kts) {
                        const float sed = fall * q * rho[c];               // :446
                        nq = q - (sed / dz[c] / rho[c]);                   // :448
                        if (species) { snow_acc = snow_acc + sed; rain_acc = rain_acc + sed; }   // :544-545
                        else rain_acc = rain_acc + sed;                                           // :516
                    } else if (k <= top + 1) {
                        nq = q - (fall * q * rho[c]) / (rho[c] * dz[c]);   // flux(k-1), :451,:455
                    }
                    if (k <= top) {
                        const float flux = fall * s_q[(k + 1) * TX + tx] * s_rho[(k + 1) * TX + tx];
                        nq = nq + flux / (rho[c] * dz[c]);                 // :454
                    }
                    if (k <= kte && nq > SMALL_VALUE) {                    // :519-528 / :549-561
                        const float L_evap = -1.f * (LH_vapor + (373.15f - T[c]) * dLHvdt);
                        // sat_mr depends on (T, p) only: reuse the last value while T has not changed
                        if (T[c] != Tq[c]) { qvs[c] = sat_mr(T[c], p[c], cx); Tq[c] = T[c]; }
                        if (qv[c] < qvs[c])
                            phase_change(T[c], nq, qvs[c], qv[c], species ? (L_melt + L_evap) : L_evap, rate, cx);
                    }
                    q = nq;
                }
            }
        }
    }
    // ---- write back: th = T/pii for the whole column (:639)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int k = tz + c * nzt;
        if (!live[c]) continue;
        const size_t a = g.at(i, k, j);
        gth[a] = T[c] / pii[c];
        gqv[a] = qv[c]; gqc[a] = qc[c]; gqr[a] = qr[c]; gqs[a] = qs[c];
    }
    if (owner && (n_r > 0 || n_s > 0)) {
        if (acc_rain) {
            acc_rain[c2] = acc_r0 + (double)rain_acc;
            acc_snow[c2] = acc_s0 + (double)snow_acc;
        } else {
            rain[c2] = rain_acc; snow[c2] = snow_acc;
        }
    }
    if (dflags) {                // events, not cells: the counters say "how many threads saw it"
        if (dflags & 1u) atomicAdd(diag + 0, 1ULL);
        if (dflags & 2u) atomicAdd(diag + 1, 1ULL);
        if (dflags & 4u) atomicAdd(diag + 2, 1ULL);
    }
}

}  // namespace

int mp_simple_launch(const Geom& g, const float* p, float* th, const float* pii, const float* rho,
                     float* qv, float* qc, float* qr, float* qs, float* rain, float* snow, int ld2,
                     float dt, const float* dz, const MpTiles& tiles, int kts, int kte, float c2r,
                     float c2s, unsigned long long* diag, cudaStream_t st, double* acc_rain, double* acc_snow) {
    int gx = 0, gy = 0;
    for (int t = 0; t < tiles.n; ++t) {
        if (tiles.ite[t] < tiles.its[t] || tiles.jte[t] < tiles.jts[t]) continue;
        const int w = tiles.ite[t] - tiles.its[t] + 1, h = tiles.jte[t] - tiles.jts[t] + 1;
        gx = std::max(gx, ((tiles.along_y[t] ? h : w) + TX - 1) / TX);
        gy = std::max(gy, tiles.along_y[t] ? w : h);
    }
    if (gx == 0 || gy == 0) return 0;
    int cpt = (g.nz + 9) / 10;                              // cells per thread: ~10 z-threads per CTA
    if (const char* e = std::getenv("ICAR_MP_CPT")) { int v = std::atoi(e); if (v >= 1 && v <= 8 && v * 10 >= g.nz) cpt = v; }
    if (cpt > 8) return -1;                                 // nz > 80 is not built
    const int nzt = (g.nz + cpt - 1) / cpt;
    dim3 block(TX, nzt);
    dim3 grid(gx, gy, tiles.n);
    size_t smem = (size_t)(2 * (g.nz + 1) + 3) * TX * sizeof(float);
#define LAUNCH(C)                                                                                 \
    k_mp_simple<C><<<grid, block, smem, st>>>(g, p, th, pii, rho, qv, qc, qr, qs, rain, snow, ld2, \
                                              dt, dz, tiles, kts, kte, c2r, c2s, diag, acc_rain, acc_snow)
    switch (cpt) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break;
        case 5: LAUNCH(5); break;
        case 6: LAUNCH(6); break;
        case 7: LAUNCH(7); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    return 1;
}

}  // namespace icar
