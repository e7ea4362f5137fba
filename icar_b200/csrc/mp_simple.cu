// "Simple" (SB04) cloud microphysics, one thread per (i,j) column; a warp covers 32
// consecutive i so every level access is one coalesced 128-byte line per field.
//
// Reference: src/physics/mp_simple.f90 -- mp_simple_driver :595-646, mp_simple :481-566,
// mp_conversions :381-420, cloud_conversion :198-280, cloud2hydrometeor :295-315,
// phase_change :333-362, sediment :437-459, sat_mr :146-182.
//
// The column is updated in place in global memory (each thread owns its column; lines
// stay in L1/L2 between the sweeps).  Temperature lives in the `th` array between the
// first sweep (T = th*pii, :633) and the last (th = T/pii, :639).
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

__device__ __forceinline__ float sat_mr(float t, float p) {            // :146-182
    float a, b;
    if (t < freezing_threshold) { a = 21.8745584f; b = 7.66f; }
    else                        { a = 17.2693882f; b = 35.86f; }
    float e_s = 610.78f * expf(a * (t - 273.16f) / (t - b));
    if ((p - e_s) <= 0.f) e_s = p * 0.99999f;
    return 0.6219907f * e_s / (p - e_s);
}

struct Diag { unsigned long long* d; };

__device__ __forceinline__ void phase_change(float& t, float& q1, float qmax, float& q2,
                                             float Lheat, float change_rate, Diag dg) {  // :333-362
    float mass2temp = Lheat / heat_capacity;
    float delta = (qmax - q2) * change_rate;
    if (delta > q1) delta = q1;
    if (delta > ((qmax - q2) * 0.99f)) delta = (qmax - q2) * 0.99f;
    q1 = q1 - delta;
    if (q1 < 0.f) {
        if ((q1 + SMALL_VALUE) < 0.f) q1 = 0.f;
        else atomicAdd(dg.d + 2, 1ULL);          // reference: print + stop
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
                                               float& qs, float c2r, float c2s, Diag dg) {  // :381-420
    const float L_melt = -1.f * LH_liquid;
    const float L_evap = -1.f * (LH_vapor + (373.15f - t) * dLHvdt);
    const float L_subl = L_melt + L_evap;
    // ---- cloud_conversion :198-280
    float qvsat = 0.f;
    {
        const float maxerr = 1e-4f;
        int iteration = 0;
        float lastqv = qv + maxerr * 2.f;
        const float vapor2temp = (LH_vapor + (373.15f - t) * dLHvdt) / heat_capacity;
        const float pre_qc = qc, pre_t = t;
        while ((fabsf(lastqv - qv) > maxerr) && (iteration < 15)) {
            iteration = iteration + 1;
            lastqv = qv;
            qvsat = sat_mr(t, p);
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
            qv = sat_mr(pre_t, p);
            t = pre_t;
            qc = pre_qc;
            atomicAdd(dg.d + 0, 1ULL);
        }
        qc = fmaxf(qc, 0.f);
        if ((t > 350.f) || (qc > 0.01f) || (qvsat > 1.f)) atomicAdd(dg.d + 1, 1ULL);
    }
    if ((qc + qr + qs) > SMALL_VALUE) {
        if (qc > SMALL_VALUE) {
            if (t > freezing_threshold) {
                cloud2hydrometeor(qc, qr, c2r);
                if (qs > SMALL_VALUE) phase_change(t, qs, 100.f, qr, L_melt, c2r, dg);
            } else {
                cloud2hydrometeor(qc, qs, c2s);
            }
        }
        if (qv < qvsat) {
            if (qr > SMALL_VALUE) phase_change(t, qr, qvsat, qv, L_evap, c2r / 2.f, dg);
            if (qs > SMALL_VALUE) phase_change(t, qs, qvsat, qv, L_subl, c2s / 2.f, dg);
        }
    }
}

// One sedimentation + re-evaporation sub-step sweep for one species (:437-459, :513-529 / :540-562).
// q: the falling species, T lives in `th`.  Returns the mass that left the bottom level.
template <bool SNOW>
__device__ __forceinline__ float sediment_sweep(float* __restrict__ q, float* __restrict__ th,
                                                float* __restrict__ qv,
                                                const float* __restrict__ p,
                                                const float* __restrict__ rho,
                                                const float* __restrict__ dz, size_t base,
                                                size_t ld, int nz, int kts, int kte, float fall,
                                                float rate, Diag dg) {
    const float L_melt = -1.f * LH_liquid;
    const int top = min(kte, nz - 2);
    const int last = max(kte, top + 1);
    size_t c = base + (size_t)kts * ld;
    float qk = q[c], rk = rho[c], dk = dz[c];
    const float sed = fall * qk * rk;
    float below = 0.f;                     // flux(k-1) = v*q_old(k)*rho(k)
    for (int k = kts; k <= last; ++k, c += ld) {
        float nq = qk;
        const bool moved = (k <= top + 1);
        if (k == kts) nq = qk - (sed / dk / rk);
        else if (moved) nq = qk - below / (rk * dk);
        float qn = 0.f, rn = 0.f, dn = 0.f;
        if (k + 1 < nz) { qn = q[c + ld]; rn = rho[c + ld]; dn = dz[c + ld]; }
        if (k <= top) {
            float flux = fall * qn * rn;
            nq = nq + flux / (rk * dk);
            below = flux;
        }
        if (k <= kte) {                    // re-evaporation / sublimation
            float t = th[c], pk = p[c], v = qv[c];
            float L_evap = -1.f * (LH_vapor + (373.15f - t) * dLHvdt);
            float qvsat = sat_mr(t, pk);
            if (v < qvsat && nq > SMALL_VALUE) {
                phase_change(t, nq, qvsat, v, SNOW ? (L_melt + L_evap) : L_evap, rate, dg);
                th[c] = t;
                qv[c] = v;
            }
        }
        if (moved || k <= kte) q[c] = nq;
        qk = qn; rk = rn; dk = dn;
    }
    return sed;
}

__global__ void __launch_bounds__(128)
k_mp_simple(Geom g, const float* __restrict__ p, float* __restrict__ th,
            const float* __restrict__ pii, const float* __restrict__ rho, float* __restrict__ qv,
            float* __restrict__ qc, float* __restrict__ qr, float* __restrict__ qs,
            float* __restrict__ rain, float* __restrict__ snow, int ld2, float dt,
            const float* __restrict__ dz, int its, int ite, int jts, int jte, int kts, int kte,
            float c2r, float c2s, unsigned long long* diag) {
    const int i = its + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = jts + blockIdx.y;
    if (i > ite || j > jte) return;
    const Diag dg{diag};
    const size_t ld = g.ld;
    const size_t base = g.at(i, 0, j);
    const int nz = g.nz;

    // sweep 1: T = th*pii (whole column, :633); conversions on kts..kte (:500-502);
    // maxval(qr), maxval(qs) and maxval(dt/dz*fall_rate) over the whole column (:505,:511)
    float max_qr = -3.4e38f, max_qs = -3.4e38f, m_r = -3.4e38f, m_s = -3.4e38f;
    size_t c = base;
    for (int k = 0; k < nz; ++k, c += ld) {
        float t = th[c] * pii[c];
        float r = qr[c], s = qs[c];
        if (k >= kts && k <= kte) {
            float v = qv[c], cw = qc[c];
            mp_conversions(p[c], t, v, cw, r, s, c2r, c2s, dg);
            qv[c] = v; qc[c] = cw; qr[c] = r; qs[c] = s;
        }
        th[c] = t;
        float dtdz = dt / dz[c];
        max_qr = fmaxf(max_qr, r);
        max_qs = fmaxf(max_qs, s);
        m_r = fmaxf(m_r, dtdz * rain_fall_rate);
        m_s = fmaxf(m_s, dtdz * snow_fall_rate);
    }
    const size_t c2 = (size_t)i + (size_t)ld2 * (size_t)j;
    float rain_acc = rain[c2];
    if (max_qr > SMALL_VALUE) {                                        // :505-530
        float cfl = ceilf(m_r);
        float fall = dt * rain_fall_rate / cfl;
        int n = (int)lroundf(cfl);
        float rate = c2r / (float)(2 * n);
        for (int s = 0; s < n; ++s)
            rain_acc = rain_acc + sediment_sweep<false>(qr, th, qv, p, rho, dz, base, ld, nz, kts,
                                                        kte, fall, rate, dg);
    }
    if (max_qs > SMALL_VALUE) {                                        // :533-564
        float snow_acc = snow[c2];
        float cfl = ceilf(m_s);
        float fall = dt * snow_fall_rate / cfl;
        int n = (int)lroundf(cfl);
        float rate = c2s / (float)(2 * n);
        for (int s = 0; s < n; ++s) {
            float f = sediment_sweep<true>(qs, th, qv, p, rho, dz, base, ld, nz, kts, kte, fall,
                                           rate, dg);
            snow_acc = snow_acc + f;
            rain_acc = rain_acc + f;
        }
        snow[c2] = snow_acc;
    }
    rain[c2] = rain_acc;
    // last sweep: th = T/pii (whole column, :639)
    c = base;
    for (int k = 0; k < nz; ++k, c += ld) th[c] = th[c] / pii[c];
}

// domain%accumulated_precipitation (REAL*8) += precipitation (mp_driver.f90:587-595)
__global__ void k_accumulate_precip(int nx, int ny, int ld2, const float* __restrict__ rain,
                                    const float* __restrict__ snow, double* __restrict__ acc_rain,
                                    double* __restrict__ acc_snow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    size_t c = (size_t)i + (size_t)ld2 * j;
    if (acc_rain) acc_rain[c] = acc_rain[c] + (double)rain[c];
    if (acc_snow) acc_snow[c] = acc_snow[c] + (double)snow[c];
}

}  // namespace

int mp_simple_launch(const Geom& g, const float* p, float* th, const float* pii, const float* rho,
                     float* qv, float* qc, float* qr, float* qs, float* rain, float* snow, int ld2,
                     float dt, const float* dz, int its, int ite, int jts, int jte, int kts,
                     int kte, float c2r, float c2s, unsigned long long* diag, cudaStream_t st) {
    if (ite < its || jte < jts) return 0;
    dim3 block(128);
    dim3 grid((ite - its + 1 + 127) / 128, jte - jts + 1);
    k_mp_simple<<<grid, block, 0, st>>>(g, p, th, pii, rho, qv, qc, qr, qs, rain, snow, ld2, dt,
                                        dz, its, ite, jts, jte, kts, kte, c2r, c2s, diag);
    return 1;
}

int accumulate_precip_launch(int nx, int ny, int ld2, const float* rain, const float* snow,
                             double* acc_rain, double* acc_snow, cudaStream_t st) {
    dim3 block(128);
    dim3 grid((nx + 127) / 128, ny);
    k_accumulate_precip<<<grid, block, 0, st>>>(nx, ny, ld2, rain, snow, acc_rain, acc_snow);
    return 1;
}

}  // namespace icar
