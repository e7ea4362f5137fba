// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of ICAR's "simple" (SB04) cloud microphysics, statement by
// statement.  PARITY PINNED TO THE REFERENCE'S OWN SOURCE (see mpdata_oracle.hpp):
// the reference holds no expected values for this path (src/tests/test_mp_simple.f90
// prints only and does not compile against the current module), so oracle/f90run
// executes src/physics/mp_simple.f90 itself; this restatement equals its outputs
// in every bit (driver on tiles and sub-tiles, kts > kms, the test_mp_simple.f90
// scenario; tests/test_reference_source.py).
//
// Reference files followed (relative to /root/reference):
//   src/physics/mp_simple.f90:64-96    constants
//   src/physics/mp_simple.f90:146-182  sat_mr
//   src/physics/mp_simple.f90:198-280  cloud_conversion
//   src/physics/mp_simple.f90:295-315  cloud2hydrometeor
//   src/physics/mp_simple.f90:333-362  phase_change
//   src/physics/mp_simple.f90:381-420  mp_conversions
//   src/physics/mp_simple.f90:437-459  sediment
//   src/physics/mp_simple.f90:481-566  mp_simple (column)
//   src/physics/mp_simple.f90:595-646  mp_simple_driver
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>
#include <algorithm>

namespace icar_oracle {

struct MpSimpleDiag {
    long long not_converged = 0;   // cloud_conversion hit 15 iterations (:248)
    long long out_of_bounds = 0;   // "mp_simple ERROR" print block (:260)
    long long negative_stop = 0;   // phase_change `stop` branch (:354-356)
};

template <class R>
struct MpSimple {
    // :64-87
    static constexpr R LH_vapor = R(2.26E6f);
    static constexpr R dLHvdt = R(2400.f);
    static constexpr R LH_liquid = R(3.34E5f);
    static constexpr R heat_capacity = R(1006.0f);
    static constexpr R SMALL_VALUE = R(1E-30f);
    static constexpr R freezing_threshold = R(273.15f);
    static constexpr R snow_fall_rate = R(1.5f);
    static constexpr R rain_fall_rate = R(10.0f);
    static constexpr R snow_cloud_init = R(0.0001f);
    static constexpr R rain_cloud_init = R(0.0001f);

    R cloud2rain, cloud2snow;   // :619-620
    MpSimpleDiag* diag;

    static R sat_mr(R temperature, R pressure) {                        // :146-182
        R a, b;
        if (temperature < freezing_threshold) { a = R(21.8745584f); b = R(7.66f); }
        else                                  { a = R(17.2693882f); b = R(35.86f); }
        R e_s = R(610.78f) * std::exp(a * (temperature - R(273.16f)) / (temperature - b));
        if ((pressure - e_s) <= R(0)) e_s = pressure * R(0.99999f);
        return R(0.6219907f) * e_s / (pressure - e_s);
    }

    void cloud_conversion(R pressure, R& temperature, R& qv, R& qc, R& qvsat) {   // :198-280
        const R maxerr = R(1e-4f);
        int iteration = 0;
        R lastqv = qv + maxerr * R(2);
        R vapor2temp = (LH_vapor + (R(373.15f) - temperature) * dLHvdt) / heat_capacity;
        R pre_qc = qc, pre_t = temperature, excess = 0;
        while ((std::abs(lastqv - qv) > maxerr) && (iteration < 15)) {             // :217
            iteration = iteration + 1;
            lastqv = qv;
            qvsat = sat_mr(temperature, pressure);
            if (qv > qvsat) {
                excess = (qv - qvsat) * R(0.5);
                temperature = temperature + (excess * vapor2temp);
                qv = qv - excess;
                qc = qc + excess;
            } else if (qc > R(0)) {
                excess = (qvsat - qv) * R(0.5);
                if (excess < qc) {
                    temperature = temperature - (excess * vapor2temp);
                    qv = qv + excess;
                    qc = qc - excess;
                } else {
                    qv = qv + qc;
                    temperature = temperature - (qc * vapor2temp);
                    excess = qc;
                    qc = R(0);
                }
            }
        }
        if (iteration == 15) {                                                      // :248-255
            qv = sat_mr(pre_t, pressure);
            temperature = pre_t;
            qc = pre_qc;
            if (diag) diag->not_converged++;
        }
        qc = std::max(qc, R(0));                                                    // :258
        if ((temperature > R(350)) || (qc > R(0.01f)) || (qvsat > R(1)))            // :260
            if (diag) diag->out_of_bounds++;
    }

    static void cloud2hydrometeor(R& qc, R& q, R conversion, R qcmin) {            // :295-315
        R delta;
        if (qc > qcmin) delta = qc - (qc * conversion);
        else            delta = 0;
        if (delta < qc) { qc = qc - delta; q = q + delta; }
        else            { q = q + qc; qc = R(0); }
        qc = std::max(qc, R(0));
    }

    void phase_change(R& temperature, R& q1, R qmax, R& q2, R Lheat, R change_rate) {   // :333-362
        R mass2temp = Lheat / heat_capacity;
        R delta = (qmax - q2) * change_rate;
        if (delta > q1) delta = q1;
        if (delta > ((qmax - q2) * R(0.99f))) delta = (qmax - q2) * R(0.99f);
        q1 = q1 - delta;
        if (q1 < R(0)) {
            if ((q1 + SMALL_VALUE) < R(0)) q1 = 0;
            else if (diag) diag->negative_stop++;            // reference: print + stop
        }
        q2 = q2 + delta;
        temperature = temperature + delta * mass2temp;
    }

    void mp_conversions(R pressure, R& temperature, R& qv, R& qc, R& qr, R& qs) {   // :381-420
        R qvsat = 0;
        R L_melt = R(-1) * LH_liquid;
        R L_evap = R(-1) * (LH_vapor + (R(373.15f) - temperature) * dLHvdt);
        R L_subl = L_melt + L_evap;
        cloud_conversion(pressure, temperature, qv, qc, qvsat);
        if ((qc + qr + qs) > SMALL_VALUE) {
            if (qc > SMALL_VALUE) {
                if (temperature > freezing_threshold) {
                    cloud2hydrometeor(qc, qr, cloud2rain, rain_cloud_init);
                    if (qs > SMALL_VALUE)
                        phase_change(temperature, qs, R(100), qr, L_melt, cloud2rain);
                } else {
                    cloud2hydrometeor(qc, qs, cloud2snow, snow_cloud_init);
                }
            }
            if (qv < qvsat) {
                if (qr > SMALL_VALUE)
                    phase_change(temperature, qr, qvsat, qv, L_evap, cloud2rain / R(2));
                if (qs > SMALL_VALUE)
                    phase_change(temperature, qs, qvsat, qv, L_subl, cloud2snow / R(2));
            }
        }
    }

    // :437-459 ; arrays are full columns (0..nz-1), kts/kte 0-based inclusive
    static R sediment(R* q, const R* v, const R* rho, const R* dz, int nz, int kts, int kte,
                      R* flux) {
        R sed = v[kts] * q[kts] * rho[kts];
        q[kts] = q[kts] - (sed / dz[kts] / rho[kts]);
        int top = std::min(kte, nz - 2);
        for (int i = kts; i <= top; ++i) flux[i] = v[i + 1] * q[i + 1] * rho[i + 1];
        for (int i = kts; i <= top; ++i) {
            q[i] = q[i] + flux[i] / (rho[i] * dz[i]);
            q[i + 1] = q[i + 1] - flux[i] / (rho[i + 1] * dz[i + 1]);
        }
        return sed;
    }

    // :481-566
    void column(const R* pressure, R* temperature, const R* rho, R* qv, R* qc, R* qr, R* qs,
                R& rain, R& snow, R dt, const R* dz, int nz, int kts, int kte, R* fall_rate,
                R* flux) {
        R L_melt = R(-1) * LH_liquid;
        for (int i = kts; i <= kte; ++i)
            mp_conversions(pressure[i], temperature[i], qv[i], qc[i], qr[i], qs[i]);

        for (int species = 0; species < 2; ++species) {
            R* qh = species == 0 ? qr : qs;
            R mx = qh[0];
            for (int i = 1; i < nz; ++i) mx = std::max(mx, qh[i]);     // maxval over full column
            if (!(mx > SMALL_VALUE)) continue;
            R rate0 = species == 0 ? rain_fall_rate : snow_fall_rate;
            R m = dt / dz[0] * rate0;
            for (int i = 1; i < nz; ++i) m = std::max(m, dt / dz[i] * rate0);
            R cfl = std::ceil(m);                                       // :511 / :538
            for (int i = 0; i < nz; ++i) fall_rate[i] = dt * rate0 / cfl;
            long ncfl = std::lround(cfl);                               // nint(cfl)
            for (long step = 1; step <= ncfl; ++step) {
                R fallen = sediment(qh, fall_rate, rho, dz, nz, kts, kte, flux);
                if (species == 0) {
                    rain = rain + fallen;                               // :516
                } else {
                    snow = snow + fallen;                               // :544-545
                    rain = rain + fallen;
                }
                for (int i = kts; i <= kte; ++i) {
                    R L_evap = R(-1) * (LH_vapor + (R(373.15f) - temperature[i]) * dLHvdt);
                    R L_subl = L_melt + L_evap;
                    R qvsat = sat_mr(temperature[i], pressure[i]);
                    if (qv[i] < qvsat) {
                        if (qh[i] > SMALL_VALUE) {
                            if (species == 0)
                                phase_change(temperature[i], qr[i], qvsat, qv[i], L_evap,
                                             cloud2rain / R(2 * ncfl));
                            else
                                phase_change(temperature[i], qs[i], qvsat, qv[i], L_subl,
                                             cloud2snow / R(2 * ncfl));
                        }
                    }
                }
            }
        }
    }
};

// mp_simple.f90:595-646.  3-D arrays (nx,nz,ny) column-major; rain/snow (nx,ny).
// its..jte are 0-based inclusive tile bounds; kts..kte likewise.
template <class R>
void mp_simple_driver(const R* pressure, R* th, const R* pii, const R* rho, R* qv, R* qc, R* qr,
                      R* qs, R* rain, R* snow, R dt, const R* dz, int nx, int nz, int ny, int its,
                      int ite, int jts, int jte, int kts, int kte, MpSimpleDiag* diag) {
    R c2s = std::exp(R(-1.0f) * (R(1) / R(2000.0f)) * dt);
    R c2r = std::exp(R(-1.0f) * (R(1) / R(500.0f)) * dt);
    MpSimpleDiag total;
#pragma omp parallel
    {
        MpSimpleDiag local;
        MpSimple<R> mp;
        mp.cloud2snow = c2s;
        mp.cloud2rain = c2r;
        mp.diag = &local;
        std::vector<R> T(nz), p(nz), r(nz), cqv(nz), cqc(nz), cqr(nz), cqs(nz), cdz(nz), fr(nz),
            fl(nz);
#pragma omp for schedule(static)
        for (int j = jts; j <= jte; ++j)
            for (int i = its; i <= ite; ++i) {
                for (int k = 0; k < nz; ++k) {
                    size_t c = (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)j);
                    T[k] = th[c] * pii[c];                              // :633
                    p[k] = pressure[c]; r[k] = rho[c]; cdz[k] = dz[c];
                    cqv[k] = qv[c]; cqc[k] = qc[c]; cqr[k] = qr[c]; cqs[k] = qs[c];
                }
                size_t c2 = (size_t)i + (size_t)nx * (size_t)j;
                mp.column(p.data(), T.data(), r.data(), cqv.data(), cqc.data(), cqr.data(),
                          cqs.data(), rain[c2], snow[c2], dt, cdz.data(), nz, kts, kte, fr.data(),
                          fl.data());
                for (int k = 0; k < nz; ++k) {
                    size_t c = (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)j);
                    th[c] = T[k] / pii[c];                              // :639
                    qv[c] = cqv[k]; qc[c] = cqc[k]; qr[c] = cqr[k]; qs[c] = cqs[k];
                }
            }
#pragma omp critical
        {
            total.not_converged += local.not_converged;
            total.out_of_bounds += local.out_of_bounds;
            total.negative_stop += local.negative_stop;
        }
    }
    if (diag) *diag = total;
}

}  // namespace icar_oracle
