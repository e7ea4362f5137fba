// Thompson et al. (2008) bulk microphysics column update on sm_100a.
//
// Reference: src/physics/mp_thompson.f90 -- mp_gt_driver :772-1044 (3-D <-> column wrapper, precipitation
// sums, water-vapour floor) and mp_thompson :1057-2844 (the column physics).
//
// Parallelisation.  The reference walks one column at a time, but almost every statement of the column
// routine is independent per level: the load/clamp block, thermodynamics, snow moments, all process
// rates, the conservation rescaling, the tendencies and the state update never look at another level.
// What couples levels is small:
//   * no_micro            -- an OR over the column                              (:1245-1345, exit :1363)
//   * graupel intercept   -- N0_min is a running minimum from the model top     (:1456-1468, :2379-2391)
//   * fall speeds         -- levels without a species inherit the speed above   (:2544-2545 ...)
//   * nstep / ksed1       -- column maxima per species                          (:2548-2555 ...)
//   * sedimentation       -- flux divergence needs sed(k+1), per sub-step       (:2660-2770)
// So, as for mp_simple, a CTA owns TX columns x all levels with ONE THREAD PER CELL (threadIdx.x = column
// within the tile, threadIdx.y = level): the cell's state and its ~55 REAL*8 process rates live in
// registers, the couplings above go through a few shared-memory rows, and each field is read once and
// written once.  Levels are processed in the Fortran's statement order within a cell, so the result
// differs from the CPU restatement only through the device's powf/expf/log10f/pow/exp.
//
// Typing follows the Fortran: REAL -> float, DOUBLE PRECISION -> double, mixed expressions promote
// where the Fortran's do.  -fmad=false (see Makefile): no contraction.
#include <algorithm>
#include <cstdlib>

#include "mp_thompson.cuh"

namespace icar {

namespace {

#ifndef ICAR_TH_TX
#define ICAR_TH_TX 8
#endif
constexpr int TX = ICAR_TH_TX;      // columns per CTA: 8 floats = one 32-byte sector per level row

constexpr float T_0 = 273.15f, PI2 = 3.1415926536f;
constexpr float rho_w = 1000.0f;
constexpr float mu_g = 0.0f, mu_i = 0.0f, mu_s = 0.6357f, Kap0 = 490.6f, Kap1 = 17.46f, Lam0 = 20.78f, Lam1 = 3.29f;
constexpr float gonv_min = 1.E4f, gonv_max = 3.E6f;
constexpr float bm_r = 3.0f, bm_s = 2.0f, bm_g = 3.0f, bm_i = 3.0f, av_r = 4854.0f, fv_r = 195.0f, bv_i = 1.0f;
constexpr float C_cube = 0.5f, R1 = 1.E-12f, R2 = 1.E-6f, eps = 1.E-15f, ATO = 0.304f;
constexpr float HGFR = 235.16f, RR2 = 287.04f, Cp2 = 1004.0f, lsub = 2.834E6f, lvap0 = 2.5E6f;
constexpr float xm0i = 1.E-12f, D0c = 1.E-6f, D0r = 50.E-6f, D0s = 200.E-6f, D0g = 250.E-6f;

__device__ const float k_sa[10] = {5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f,
                                   0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
__device__ const float k_sb[10] = {0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f,
                                   0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};

__device__ __noinline__ float rslf(float P, float T) {                  // :3776-3805
    const float C0 = .611583699E03f, C1 = .444606896E02f, C2 = .143177157E01f, C3 = .264224321E-1f,
                C4 = .299291081E-3f, C5 = .203154182E-5f, C6 = .702620698E-8f, C7 = .379534310E-11f,
                C8 = -.321582393E-13f;
    const float X = fmaxf(-80.f, T - 273.16f);
    const float ESL = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
    return .622f * ESL / (P - ESL);
}
__device__ __noinline__ float rsif(float P, float T) {                  // :3810-3835
    const float C0 = .609868993E03f, C1 = .499320233E02f, C2 = .184672631E01f, C3 = .402737184E-1f,
                C4 = .565392987E-3f, C5 = .521693933E-5f, C6 = .307839583E-7f, C7 = .105785160E-9f,
                C8 = .161444444E-12f;
    const float X = fmaxf(-80.f, T - 273.16f);
    const float ESI = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
    return .622f * ESI / (P - ESI);
}

// x**y in REAL*8.  The exponents are run-time constants of the scheme (gamma-distribution shape
// parameters: cre(), cge(), bm_r ...) and for the default options most are small integers -- x**1,
// x**3, (lamr+fv_r)**(-5) ... -- which a few multiplications evaluate to within a few ulp(double) of pow()
// (also quarter-integer exponents such as oge1 = 1/4, cre(12) = 2.5, cre(7) = 3.5, through sqrt);
// the general case stays a call (not inlined: ~250 instructions each, the kernel has 40 call sites).
// Code size matters here: with every helper inlined the kernel was 17 000 SASS instructions (272 KB) and
// "no instruction" (instruction-cache misses) was the top warp stall, 6.8 warps per issue slot.  The
// transcendental functions and the helpers that contain them are therefore real calls.
__device__ __noinline__ double dpow_general(double x, double y) { return pow(x, y); }
__device__ __noinline__ float fpow(float x, float y) { return powf(x, y); }
__device__ __noinline__ float fexp(float x) { return expf(x); }
__device__ __noinline__ float flog10(float x) { return log10f(x); }
__device__ __noinline__ double dlog(double x) { return log(x); }
__device__ __noinline__ int nint_f(float x) { return (int)lroundf(x); }
__device__ __noinline__ double dpow(double x, double y) {
    const double y4 = 4.0 * y;
    if (y4 == rint(y4) && fabs(y) <= 8.0) {        // uniform across the grid: y is a constant
        // y = n + f/4: x**n by multiplication, x**(1/2) and x**(1/4) by (correctly rounded) square roots
        const double ay = fabs(y);
        int n = (int)ay;
        const int f = (int)(4.0 * (ay - (double)n));      // 0..3
        double p = 1.0, b = x;
        while (n) { if (n & 1) p *= b; b *= b; n >>= 1; }
        if (f) {
            const double r2 = sqrt(x);
            if (f & 2) p *= r2;
            if (f & 1) p *= sqrt(r2);
        }
        return y < 0.0 ? 1.0 / p : p;
    }
    return dpow_general(x, y);
}

// 10.**n for integer n the way gfortran evaluates it (libgcc __powisf2)
__device__ __noinline__ float pow10i(int m) {
    unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
    float x = 10.f, y = (n & 1u) ? x : 1.f;
    while (n >>= 1) { x = x * x; if (n & 1u) y *= x; }
    return m < 0 ? 1.f / y : y;
}
// index into a decade table, :1562-1676
__device__ __noinline__ int decade_index(float x, int n2, int ntb) {
    const int nic = (int)lroundf(flog10(x));
    int n = nic - 1;
    for (int nn = nic - 1; nn <= nic + 1; ++nn) {
        n = nn;
        const float r = x / pow10i(nn);
        if (r >= 1.0f && r < 10.0f) break;
    }
    const int idx = (int)(x / pow10i(n)) + 10 * (n - n2) - (n - n2);
    return max(1, min(idx, ntb));
}
__device__ __noinline__ int decade_index_d(double x, int n2, int ntb) {
    const int nic = (int)lround(log10(x));
    int n = nic - 1;
    for (int nn = nic - 1; nn <= nic + 1; ++nn) {
        n = nn;
        const double r = x / pow10i(nn);
        if (r >= 1.0 && r < 10.0) break;
    }
    const int idx = (int)(x / pow10i(n)) + 10 * (n - n2) - (n - n2);
    return max(1, min(idx, ntb));
}

__device__ __forceinline__ float sa_poly(const float* s, float tc0, float x) {
    return s[0] + s[1] * tc0 + s[2] * x + s[3] * tc0 * x + s[4] * tc0 * tc0 + s[5] * x * x + s[6] * tc0 * tc0 * x +
           s[7] * tc0 * x * x + s[8] * tc0 * tc0 * tc0 + s[9] * x * x * x;
}
// a_ * smo2**b_ for the `x`-th moment (e.g. :1413-1423)
__device__ __noinline__ float snow_moment(float tc0, float x, float smo2) {
    const float loga_ = sa_poly(k_sa, tc0, x);
    const float a_ = fpow(10.0f, loga_);
    const float b_ = sa_poly(k_sb, tc0, x);
    return a_ * fpow(smo2, b_);
}

__device__ __forceinline__ size_t t4(int i, int j, int k, int m, int n1, int n2, int n3) {   // 1-based, column-major
    return (size_t)(i - 1) + (size_t)n1 * ((size_t)(j - 1) + (size_t)n2 * ((size_t)(k - 1) + (size_t)n3 * (size_t)(m - 1)));
}

// rain mass/number balance (:1283-1293, :2185-2197, :2286-2296): returns mvd_r; in the two clamped
// branches nr is recomputed and `clamped` set
__device__ __noinline__ float rain_balance(const ThConst& c, float rr, float& nr, bool& clamped) {
    double lamr = fpow((c.am_r * c.crg[2] * c.org2 * nr / rr), c.obmr);
    float mvd = (float)((3.0f + c.o.mu_r + 0.672f) / lamr);
    clamped = false;
    if (mvd > 2.5E-3f) {
        mvd = 2.5E-3f;
        lamr = (3.0f + c.o.mu_r + 0.672f) / mvd;
        nr = (float)(c.crg[1] * c.org3 * rr * dpow(lamr, (double)bm_r) / c.am_r);
        clamped = true;
    } else if (mvd < D0r * 0.75f) {
        mvd = D0r * 0.75f;
        lamr = (3.0f + c.o.mu_r + 0.672f) / mvd;
        nr = (float)(c.crg[1] * c.org3 * rr * dpow(lamr, (double)bm_r) / c.am_r);
        clamped = true;
    }
    return mvd;
}

struct Args {
    float *qv, *qc, *qr, *qi, *qs, *qg, *ni, *nr, *th;
    const float *pii, *p, *dz;
    float *rainnc, *snownc, *snowncv, *graupelnc, *graupelncv, *sr;
    const float* rainncv;
    double *acc_precip, *acc_snow, *acc_graupel;      // resident mode: REAL*8 accumulators (mp_driver.f90:587-595)
};

// Shared memory (doubles first): s_n0[nz][TX] double | s_f[7][(nz+1)][TX] float | s_i[10][TX] int
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_mp_thompson(Geom g, ThConst c, ThTables tb, Args A, int ld2, float dt, MpTiles tiles, int kts, int kte,
              unsigned long long* diag) {
    const int its = tiles.its[blockIdx.z], ite = tiles.ite[blockIdx.z];
    const int jts = tiles.jts[blockIdx.z], jte = tiles.jte[blockIdx.z];
    const bool along_y = tiles.along_y[blockIdx.z] != 0;
    const int lane0 = (int)blockIdx.x * TX, row = (int)blockIdx.y;
    if (along_y ? (jts + lane0 > jte || its + row > ite) : (its + lane0 > ite || jts + row > jte)) return;
    extern __shared__ double smem_d[];
    const int nk = kte - kts + 1;                 // levels of the column, local index kk = 0..nk-1
    const int tx = threadIdx.x, kk = threadIdx.y;
    double* s_n0 = smem_d;                                            // [nk][TX]
    float* s_f = reinterpret_cast<float*>(s_n0 + (size_t)nk * TX);    // 7 planes of [(nk+1)][TX]
    const int PL = (nk + 1) * TX;
    int* s_i = reinterpret_cast<int*>(s_f + 7 * PL);                  // [10][TX]
    const int i = along_y ? its + row : its + lane0 + tx;
    const int j = along_y ? jts + lane0 + tx : jts + row;
    const bool col_ok = (i <= ite) && (j <= jte);
    const bool live = col_ok && kk < nk;
    const int k = kts + kk;
    const size_t a = live ? g.at(i, k, j) : 0;

    // -------------------------------------------------------------- gather (:884-896)
    float t1d = 273.f, p1d = 1.e5f, dzq = 1.f, pii = 1.f;
    float qv1d = 0.f, qc1d = 0.f, qi1d = 0.f, qr1d = 0.f, qs1d = 0.f, qg1d = 0.f, ni1d = 0.f, nr1d = 0.f;
    if (live) {
        pii = A.pii[a];
        t1d = A.th[a] * pii; p1d = A.p[a]; dzq = A.dz[a];
        qv1d = A.qv[a]; qc1d = A.qc[a]; qi1d = A.qi[a]; qr1d = A.qr[a]; qs1d = A.qs[a]; qg1d = A.qg[a];
        ni1d = A.ni[a]; nr1d = A.nr[a];
    }
    if (kk == 0) { for (int n = 0; n < 10; ++n) s_i[n * TX + tx] = 0; }
    // (accumulators are read here, not at the end: a dependent load in a CTA's last instructions is exposed)
    double acc0[3] = {0.0, 0.0, 0.0};
    if (live && kk == 0 && A.acc_precip) {
        const size_t a2 = (size_t)i + (size_t)ld2 * (size_t)j;
        acc0[0] = A.acc_precip[a2]; acc0[1] = A.acc_snow[a2]; acc0[2] = A.acc_graupel[a2];
    }
    __syncthreads();

    const float dtsave = dt, odt = 1.f / dt, odts = 1.f / dtsave;
    float tten = 0.f, qvten = 0.f, qcten = 0.f, qiten = 0.f, qrten = 0.f, qsten = 0.f, qgten = 0.f, niten = 0.f, nrten = 0.f;

    // -------------------------------------------------------------- local copies, clamps (:1240-1319)
    float temp = t1d, qv = fmaxf(1.E-10f, qv1d), pres = p1d;
    float rho = 0.622f * pres / (RR2 * temp * (qv + 0.622f));
    float rc, ri, rr, rs, rg, ni, nr, mvd_r = 0.0f;
    bool L_qc, L_qi, L_qr, L_qs, L_qg, micro = false;
    if (qc1d > R1) { micro = true; rc = qc1d * rho; L_qc = true; } else { qc1d = 0.0f; rc = R1; L_qc = false; }
    if (qi1d > R1) {
        micro = true;
        ri = qi1d * rho;
        ni = fmaxf(R2, ni1d * rho);
        L_qi = true;
        double lami = fpow((c.am_i * c.cig[1] * c.oig1 * ni / ri), c.obmi);
        const double ilami = 1. / lami;
        const float xDi = (float)((bm_i + mu_i + 1.f) * ilami);
        if (xDi < 20.E-6f) {
            lami = c.cie[1] / 20.E-6f;
            ni = (float)fmin(250.E3, c.cig[0] * c.oig2 * ri / c.am_i * dpow(lami, (double)bm_i));
        } else if (xDi > 300.E-6f) {
            lami = c.cie[1] / 300.E-6f;
            ni = (float)(c.cig[0] * c.oig2 * ri / c.am_i * dpow(lami, (double)bm_i));
        }
    } else { qi1d = 0.0f; ni1d = 0.0f; ri = R1; ni = R2; L_qi = false; }
    if (qr1d > R1) {
        micro = true;
        rr = qr1d * rho;
        nr = fmaxf(R2, nr1d * rho);
        L_qr = true;
        bool cl;
        mvd_r = rain_balance(c, rr, nr, cl);
    } else { qr1d = 0.0f; nr1d = 0.0f; rr = R1; nr = R2; L_qr = false; }
    if (qs1d > R1) { micro = true; rs = qs1d * rho; L_qs = true; } else { qs1d = 0.0f; rs = R1; L_qs = false; }
    if (qg1d > R1) { micro = true; rg = qg1d * rho; L_qg = true; } else { qg1d = 0.0f; rg = R1; L_qg = false; }

    // -------------------------------------------------------------- thermodynamics (:1328-1356)
    float tempc = temp - 273.15f;
    float rhof = sqrtf(c.rho_not / rho);
    float rhof2 = sqrtf(rhof);
    float qvs = rslf(pres, temp);
    const float delQvs = fmaxf(0.0f, rslf(pres, 273.15f) - qv);
    const float qvsi = (tempc <= 0.0f) ? rsif(pres, temp) : qvs;
    float ssatw = qv / qvs - 1.f;
    float ssati = qv / qvsi - 1.f;
    if (fabsf(ssatw) < eps) ssatw = 0.0f;
    if (fabsf(ssati) < eps) ssati = 0.0f;
    if (ssati > 0.0f) micro = true;
    float diffu = 2.11E-5f * fpow((temp / 273.15f), 1.94f) * (101325.f / pres);
    float visco = (tempc >= 0.0f) ? (1.718f + 0.0049f * tempc) * 1.0E-5f
                                  : (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
    float ocp = 1.f / (Cp2 * (1.f + 0.887f * qv));
    float vsc2 = sqrtf(rho / visco);
    float lvap = lvap0 + (2106.0f - 4218.0f) * tempc;
    float tcond = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;

    if (live && micro) s_i[tx] = 1;
    __syncthreads();
    const bool active = live && (s_i[tx] != 0);        // .not. no_micro  (:1363)
    if (col_ok && kk == 0 && s_i[tx] == 0) atomicAdd(diag, 1ULL);

    // process rates (REAL*8, :1077-1098), all zero-initialised (:1173-1235)
    double prw_vcd = 0, prr_wau = 0, prr_rcw = 0, prr_rcs = 0, prr_rcg = 0, prr_sml = 0, prr_gml = 0, prr_rci = 0, prv_rev = 0,
           pnr_wau = 0, pnr_rcs = 0, pnr_rcg = 0, pnr_rci = 0, pnr_sml = 0, pnr_gml = 0, pnr_rev = 0, pnr_rcr = 0, pnr_rfz = 0;
    double pri_inu = 0, pni_inu = 0, pri_ihm = 0, pni_ihm = 0, pri_wfz = 0, pni_wfz = 0, pri_rfz = 0, pni_rfz = 0, pri_ide = 0,
           pni_ide = 0, pri_rci = 0, pni_rci = 0, pni_sci = 0, pni_iau = 0;
    double prs_iau = 0, prs_sci = 0, prs_rcs = 0, prs_scw = 0, prs_sde = 0, prs_ihm = 0, prs_ide = 0;
    double prg_scw = 0, prg_rfz = 0, prg_gde = 0, prg_gcw = 0, prg_rci = 0, prg_rcs = 0, prg_rcg = 0, prg_ihm = 0;
    double ilamr = 0, ilamg = 0, N0_r = 0, N0_g = 0;
    float smob = 0.f, smo2 = 0.f, smo1 = 0.f, smo0 = 0.f, smoc = 0.f, smod = 0.f, smoe = 0.f, smof = 0.f;
    float mvd_c = D0c, vts_boost = 1.5f;

    // -------------------------------------------------------------- snow moments (:1369-1451)
    if (active && L_qs) {
        const float tc0 = fminf(-0.1f, temp - 273.15f);
        smob = rs * c.oams;
        if (bm_s > (2.0f - 1.e-3f) && bm_s < (2.0f + 1.e-3f)) smo2 = smob;
        else {
            const float loga_ = sa_poly(k_sa, tc0, bm_s), a_ = fpow(10.0f, loga_), b_ = sa_poly(k_sb, tc0, bm_s);
            smo2 = fpow((smob / a_), (1.f / b_));
        }
        {   // 0th moment :1394-1397
            const float loga_ = k_sa[0] + k_sa[1] * tc0 + k_sa[4] * tc0 * tc0 + k_sa[8] * tc0 * tc0 * tc0;
            const float a_ = fpow(10.0f, loga_);
            const float b_ = k_sb[0] + k_sb[1] * tc0 + k_sb[4] * tc0 * tc0 + k_sb[8] * tc0 * tc0 * tc0;
            smo0 = a_ * fpow(smo2, b_);
        }
        {   // 1st moment :1400-1410
            const float loga_ = k_sa[0] + k_sa[1] * tc0 + k_sa[2] + k_sa[3] * tc0 + k_sa[4] * tc0 * tc0 + k_sa[5] +
                                k_sa[6] * tc0 * tc0 + k_sa[7] * tc0 + k_sa[8] * tc0 * tc0 * tc0 + k_sa[9];
            const float a_ = fpow(10.0f, loga_);
            const float b_ = k_sb[0] + k_sb[1] * tc0 + k_sb[2] + k_sb[3] * tc0 + k_sb[4] * tc0 * tc0 + k_sb[5] +
                             k_sb[6] * tc0 * tc0 + k_sb[7] * tc0 + k_sb[8] * tc0 * tc0 * tc0 + k_sb[9];
            smo1 = a_ * fpow(smo2, b_);
        }
        smoc = snow_moment(tc0, c.cse[0], smo2);
        smoe = snow_moment(tc0, c.cse[12], smo2);
        smof = snow_moment(tc0, c.cse[15], smo2);
    }

    // -------------------------------------------------------------- graupel / rain intercepts (:1456-1494)
    // N0_min(k) = min(gonv_max, N0_exp(k), N0_exp(k+1), ..., N0_exp(kte)): a suffix minimum (exact, so
    // the order of the comparisons is immaterial).
    auto graupel_rain_slopes = [&]() {
        double N0_exp = 0.0;
        if (active) {
            const float xslw1 = (temp < 270.65f && L_qr && mvd_r > 100.E-6f) ? 4.01f + flog10(mvd_r) : 0.01f;
            const float ygra1 = 4.31f + flog10(fmaxf(5.E-5f, rg));
            const float zans1 = 3.1f + (100.f / (300.f * xslw1 * ygra1 / (10.f / xslw1 + 1.f + 0.25f * ygra1) + 30.f + 10.f * ygra1));
            N0_exp = fpow(10.f, (zans1));
            N0_exp = fmax((double)gonv_min, fmin(N0_exp, (double)gonv_max));
        }
        __syncthreads();                       // previous readers of s_n0 are done
        if (kk < nk) s_n0[kk * TX + tx] = N0_exp;
        __syncthreads();
        if (active) {
            double N0_min = gonv_max;
#pragma unroll 1
            for (int q = kk; q < nk; ++q) N0_min = fmin(s_n0[q * TX + tx], N0_min);
            N0_exp = N0_min;
            const double lam_exp = dpow((N0_exp * c.am_g * c.cgg[0] / rg), (double)c.oge1);
            const double lamg = lam_exp * fpow((c.cgg[2] * c.ogg2 * c.ogg1), c.obmg);
            ilamg = 1. / lamg;
            N0_g = N0_exp / (c.cgg[1] * lam_exp) * dpow(lamg, (double)c.cge[1]);
            const double lamr = fpow((c.am_r * c.crg[2] * c.org2 * nr / rr), c.obmr);
            ilamr = 1. / lamr;
            mvd_r = (float)((3.0f + c.o.mu_r + 0.672f) / lamr);
            N0_r = nr * c.org2 * dpow(lamr, (double)c.cre[1]);
        }
    };
    graupel_rain_slopes();

    float rvs = 0.f, t1_subl = 0.f;
    double lamr_fv9 = -1.0;        // (lamr + fv_r)**(-cre(9)) once it has been evaluated (it is positive)
    if (active) {
        // ---------------------------------------------------------- warm rain (:1500-1544)
        if (L_qr && mvd_r > D0r) {
            const float Ef_rr = 2.0f - fexp(2300.0f * (mvd_r - 1600.0E-6f));
            pnr_rcr = Ef_rr * 4.f * nr * rr;
        }
        mvd_c = D0c;
        if (L_qc) {
            const float xDc = fmaxf(D0c * 1.E6f, (fpow((rc / (c.am_r * c.o.Nt_c)), c.obmr)) * 1.E6f);
            const double lamc = fpow((c.o.Nt_c * c.am_r * c.ccg[1] * c.ocg1 / rc), c.obmr);
            mvd_c = (float)((3.0f + c.mu_c + 0.672f) / lamc);
            if (rc > 0.01e-3f) {
                const float Dc_g = (float)((fpow((c.ccg[2] * c.ocg2), c.obmr) / lamc) * 1.E6f);
                const float Dc_b = fpow((xDc * xDc * xDc * Dc_g * Dc_g * Dc_g - xDc * xDc * xDc * xDc * xDc * xDc), (1.f / 6.f));
                const float zeta1 = 0.5f * ((6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f) + fabsf(6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f));
                const float zeta = 0.027f * rc * zeta1;
                const float taud = 0.5f * ((0.5f * Dc_b - 7.5f) + fabsf(0.5f * Dc_b - 7.5f)) + R1;
                const float tau = 3.72f / (rc * taud);
                prr_wau = zeta / tau;
                prr_wau = fmin((double)(rc * odts), prr_wau);
                pnr_wau = prr_wau / (c.am_r * c.mu_c * D0r * D0r * D0r);
            }
            if (L_qr && mvd_r > D0r && mvd_c > D0c) {
                const double lamr = 1. / ilamr;
                int idx = 1 + (int)(nbins * dlog(mvd_r / c.Dr1) / dlog(c.Drn / c.Dr1));
                idx = min(idx, nbins);
                const int jc = min((int)(mvd_c * 1.E6f), nbins);          // (upper clamp: memory safety only)
                const float Ef_rw = (float)tb.t[T_Efrw][(idx - 1) + nbins * (jc - 1)];
                lamr_fv9 = dpow((lamr + fv_r), (double)(-c.cre[8]));          // reused by pri_rci / pnr_rci (same ilamr)
                prr_rcw = rhof * c.t1_qr_qc * Ef_rw * rc * N0_r * (lamr_fv9);
                prr_rcw = fmin((double)(rc * odts), prr_rcw);
            }
        }

        // ---------------------------------------------------------- frozen species (:1550-2009)
        vts_boost = 1.5f;
        tempc = temp - 273.15f;
        const int idx_tc = max(1, min(nint_f(-tempc), 45));
        int idx_t = (int)((tempc - 2.5f) / 5.f) - 1;
        idx_t = max(1, -idx_t);
        idx_t = min(idx_t, ntb_t);
        const int idx_c = (rc > 1.e-6f) ? decade_index(rc, c.nic2, ntb_c) : 1;
        const int idx_i = (ri > 1.e-10f) ? decade_index(ri, c.nii2, ntb_i) : 1;
        const int idx_i1 = (ni > 1.0f) ? decade_index(ni, c.nii3, ntb_i1) : 1;
        int idx_r = 1, idx_r1 = ntb_r1, idx_g = 1, idx_g1 = ntb_g1;
        if (rr > 1.e-6f) {
            idx_r = decade_index(rr, c.nir2, ntb_r);
            const double lamr = 1. / ilamr;
            const double lam_exp = lamr * fpow((c.crg[2] * c.org2 * c.org1), bm_r);
            const double N0_exp = c.org1 * rr / c.am_r * dpow(lam_exp, (double)c.cre[0]);
            idx_r1 = decade_index_d(N0_exp, c.nir3, ntb_r1);
        }
        const int idx_s = (rs > 1.e-5f) ? decade_index(rs, c.nis2, ntb_s) : 1;
        if (rg > 1.e-5f) {
            idx_g = decade_index(rg, c.nig2, ntb_g);
            const double lamg = 1. / ilamg;
            const double lam_exp = lamg * fpow((c.cgg[2] * c.ogg2 * c.ogg1), bm_g);
            const double N0_exp = c.ogg1 * rg / c.am_g * dpow(lam_exp, (double)c.cge[0]);
            idx_g1 = decade_index_d(N0_exp, c.nig3, ntb_g1);
        }

        {   // deposition / sublimation prefactor (:1679-1695)
            const float otemp = 1.f / temp;
            rvs = rho * qvsi;
            const float rvs_p = rvs * otemp * (lsub * otemp * c.oRv - 1.f);
            const float rvs_pp = rvs * (otemp * (lsub * otemp * c.oRv - 1.f) * otemp * (lsub * otemp * c.oRv - 1.f) +
                                        (-2.f * lsub * otemp * otemp * otemp * c.oRv) + otemp * otemp);
            const float gamsc = lsub * diffu / tcond * rvs_p;
            float alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
            alphsc = fmaxf(1.E-9f, alphsc);
            float xsat = ssati;
            if (fabsf(xsat) < 1.E-9f) xsat = 0.f;
            t1_subl = 4.f * PI2 * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat -
                                   5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);
        }

        if (L_qc && mvd_c > D0c) {                                           // :1698-1725
            float xDs = 0.0f;
            if (L_qs) xDs = smoc / smob;
            if (xDs > D0s) {
                int idx = 1 + (int)(nbins * dlog(xDs / c.Ds1) / dlog(c.Dsn / c.Ds1));
                idx = min(idx, nbins);
                const int jc = min((int)(mvd_c * 1.E6f), nbins);
                const float Ef_sw = (float)tb.t[T_Efsw][(idx - 1) + nbins * (jc - 1)];
                prs_scw = rhof * c.t1_qs_qc * Ef_sw * rc * smoe;
            }
            if (rg >= 1.e-5f && mvd_c > D0c) {
                const float xDg = (float)((bm_g + mu_g + 1.f) * ilamg);
                const float vtg = (float)(rhof * c.o.av_g * c.cgg[5] * c.ogg3 * dpow(ilamg, (double)c.o.bv_g));
                const float stoke_g = mvd_c * mvd_c * vtg * rho_w / (9.f * visco * xDg);
                if (xDg > D0g) {
                    float Ef_gw = 0.f;
                    if (stoke_g >= 0.4f && stoke_g <= 10.f) Ef_gw = 0.55f * flog10(2.51f * stoke_g);
                    else if (stoke_g < 0.4f) Ef_gw = 0.0f;
                    else if (stoke_g > 10.f) Ef_gw = 0.77f;
                    prg_gcw = rhof * c.t1_qg_qc * Ef_gw * rc * N0_g * dpow(ilamg, (double)c.cge[8]);
                }
            }
        }

        if (rr >= 1.e-6f) {                                                  // :1730-1783
            if (rs >= 1.e-5f) {
                const size_t q = t4(idx_s, idx_t, idx_r1, idx_r, ntb_s, ntb_t, ntb_r1);
                const double tmr_racs1 = tb.t[T_tmr_racs1][q], tmr_racs2 = tb.t[T_tmr_racs2][q],
                             tcr_sacr1 = tb.t[T_tcr_sacr1][q], tcr_sacr2 = tb.t[T_tcr_sacr2][q],
                             tcs_racs1 = tb.t[T_tcs_racs1][q], tms_sacr1 = tb.t[T_tms_sacr1][q];
                if (temp < T_0) {
                    prr_rcs = -(tmr_racs2 + tcr_sacr2 + tmr_racs1 + tcr_sacr1);
                    prs_rcs = tmr_racs2 + tcr_sacr2 - tcs_racs1 - tms_sacr1;
                    prg_rcs = tmr_racs1 + tcr_sacr1 + tcs_racs1 + tms_sacr1;
                    prr_rcs = fmax((double)(-rr * odts), prr_rcs);
                    prs_rcs = fmax((double)(-rs * odts), prs_rcs);
                    prg_rcs = fmin((double)((rr + rs) * odts), prg_rcs);
                    pnr_rcs = tb.t[T_tnr_racs1][q] + tb.t[T_tnr_racs2][q] + tb.t[T_tnr_sacr1][q] + tb.t[T_tnr_sacr2][q];
                } else {
                    prs_rcs = -tcs_racs1 - tms_sacr1 + tmr_racs2 + tcr_sacr2;
                    prs_rcs = fmax((double)(-rs * odts), prs_rcs);
                    prr_rcs = -prs_rcs;
                    pnr_rcs = tb.t[T_tnr_racs2][q] + tb.t[T_tnr_sacr2][q];
                }
                pnr_rcs = fmin((double)(nr * odts), pnr_rcs);
            }
            if (rg >= 1.e-5f) {
                const size_t q = t4(idx_g1, idx_g, idx_r1, idx_r, ntb_g1, ntb_g, ntb_r1);
                if (temp < T_0) {
                    prg_rcg = tb.t[T_tmr_racg][q] + tb.t[T_tcr_gacr][q];
                    prg_rcg = fmin((double)(rr * odts), prg_rcg);
                    prr_rcg = -prg_rcg;
                    pnr_rcg = tb.t[T_tnr_racg][q] + tb.t[T_tnr_gacr][q];
                    pnr_rcg = fmin((double)(nr * odts), pnr_rcg);
                } else {
                    prr_rcg = tb.t[T_tcg_racg][q];
                    prr_rcg = fmin((double)(rg * odts), prr_rcg);
                    prg_rcg = -prr_rcg;
                }
            }
        }

        if (temp < T_0) {                                                    // :1789-1949
            vts_boost = 1.0f;
            float rate_max = (qv - qvsi) * rho * odts * 0.999f;
            if (rr > 1.e-6f) {
                const size_t q = (size_t)(idx_r - 1) + (size_t)ntb_r * ((size_t)(idx_r1 - 1) + (size_t)ntb_r1 * (size_t)(idx_tc - 1));
                prg_rfz = tb.t[T_tpg_qrfz][q] * odts;
                pri_rfz = tb.t[T_tpi_qrfz][q] * odts;
                pni_rfz = tb.t[T_tni_qrfz][q] * odts;
                pnr_rfz = tb.t[T_tnr_qrfz][q] * odts;
                pnr_rfz = fmin((double)(nr * odts), pnr_rfz);
            } else if (rr > R1 && temp < HGFR) {
                pri_rfz = rr * odts;
                pnr_rfz = nr * odts;
                pni_rfz = pnr_rfz;
            }
            if (rc > 1.e-6f) {
                const size_t q = (size_t)(idx_c - 1) + (size_t)ntb_c * (size_t)(idx_tc - 1);
                pri_wfz = tb.t[T_tpi_qcfz][q] * odts;
                pri_wfz = fmin((double)(rc * odts), pri_wfz);
                pni_wfz = tb.t[T_tni_qcfz][q] * odts;
                pni_wfz = fmin(fmin((double)(c.o.Nt_c * odts), pri_wfz / (2.f * xm0i)), pni_wfz);
            } else if (rc > R1 && temp < HGFR) {
                pri_wfz = rc * odts;
                pni_wfz = fmin(fmin((double)(c.o.Nt_c * odts), pri_wfz / (2.f * xm0i)), pni_wfz);
            }
            if ((ssati >= 0.25f) || (ssatw > eps && temp < 261.15f)) {       // :1820-1827
                const float xnc = fminf(250.E3f, c.o.TNO * fexp(ATO * (T_0 - temp)));
                const float xni = (float)(ni + (pni_rfz + pni_wfz) * dtsave);
                pni_inu = 0.5f * (xnc - xni + fabsf(xnc - xni)) * odts;
                pri_inu = fmin((double)rate_max, xm0i * pni_inu);
                pni_inu = pri_inu / xm0i;
            }
            // (the Fortran evaluates lami, xDi, xmi twice from the same ni, ri -- :1831-1835 and :1897-1901 -- once here)
            double ilami = 0.0;
            float xDi = 0.f, oxmi = 0.f;
            if (L_qi) {
                const double lami = fpow((c.am_i * c.cig[1] * c.oig1 * ni / ri), c.obmi);
                ilami = 1. / lami;
                xDi = (float)fmax((double)c.D0i, (bm_i + mu_i + 1.f) * ilami);
                const float xmi = c.am_i * fpow(xDi, bm_i);
                oxmi = 1.f / xmi;
            }
            if (L_qi) {                                                      // :1830-1863
                pri_ide = C_cube * t1_subl * diffu * ssati * rvs * c.oig1 * c.cig[4] * ni * ilami;
                const size_t q = (size_t)(idx_i - 1) + (size_t)ntb_i * (size_t)(idx_i1 - 1);
                if (pri_ide < 0.0) {
                    pri_ide = fmax(fmax((double)(-ri * odts), pri_ide), (double)rate_max);
                    pni_ide = pri_ide * oxmi;
                    pni_ide = fmax((double)(-ni * odts), pni_ide);
                } else {
                    pri_ide = fmin(pri_ide, (double)rate_max);
                    const double tp = tb.t[T_tpi_ide][q];
                    prs_ide = (1.0 - tp) * pri_ide;
                    pri_ide = tp * pri_ide;
                }
                if ((idx_i == ntb_i) || (xDi > 5.0f * D0s)) {
                    prs_iau = ri * .99f * odts;
                    pni_iau = ni * .95f * odts;
                } else if (xDi < 0.1f * D0s) {
                    prs_iau = 0.; pni_iau = 0.;
                } else {
                    prs_iau = tb.t[T_tps_iaus][q] * odts;
                    prs_iau = fmin((double)(ri * .99f * odts), prs_iau);
                    pni_iau = tb.t[T_tni_iaus][q] * odts;
                    pni_iau = fmin((double)(ni * .95f * odts), pni_iau);
                }
            }
            if (L_qs) {                                                      // :1867-1882
                float C_snow = c.o.C_sqrd + (tempc + 15.f) * (c.o.C_cubes - c.o.C_sqrd) / (-30.f + 15.f);
                C_snow = fmaxf(c.o.C_sqrd, fminf(C_snow, c.o.C_cubes));
                prs_sde = C_snow * t1_subl * diffu * ssati * rvs * (c.t1_qs_sd * smo1 + c.t2_qs_sd * rhof2 * vsc2 * smof);
                if (prs_sde < 0.) prs_sde = fmax(fmax((double)(-rs * odts), prs_sde), (double)rate_max);
                else prs_sde = fmin(prs_sde, (double)rate_max);
            }
            if (L_qg && ssati < -eps) {                                      // :1884-1893
                prg_gde = C_cube * t1_subl * diffu * ssati * rvs * N0_g *
                          (c.t1_qg_sd * dpow(ilamg, (double)c.cge[9]) + c.t2_qg_sd * vsc2 * rhof2 * dpow(ilamg, (double)c.cge[10]));
                if (prg_gde < 0.) prg_gde = fmax(fmax((double)(-rg * odts), prg_gde), (double)rate_max);
                else prg_gde = fmin(prg_gde, (double)rate_max);
            }
            if (L_qi) {                                                      // :1896-1920
                if (rs >= 1.e-5f) {
                    prs_sci = c.t1_qs_qi * rhof * c.o.Ef_si * ri * smoe;
                    pni_sci = prs_sci * oxmi;
                }
                if (rr >= 1.e-6f && mvd_r > 4.f * xDi) {
                    const double lamr = 1. / ilamr;
                    const double lf9 = (lamr_fv9 >= 0.0) ? lamr_fv9 : dpow((lamr + fv_r), (double)(-c.cre[8]));
                    pri_rci = rhof * c.t1_qr_qi * c.o.Ef_ri * ri * N0_r * (lf9);
                    pnr_rci = rhof * c.t1_qr_qi * c.o.Ef_ri * ni * N0_r * (lf9);
                    pni_rci = pri_rci * oxmi;
                    prr_rci = rhof * c.t2_qr_qi * c.o.Ef_ri * ni * N0_r * (dpow((lamr + fv_r), (double)(-c.cre[7])));
                    prr_rci = fmin((double)(rr * odts), prr_rci);
                    prg_rci = pri_rci + prr_rci;
                }
            }
            if (prg_gcw > eps && tempc > -8.0f) {                            // :1923-1936
                float tf = 0.f;
                if (tempc >= -5.0f && tempc < -3.0f) tf = 0.5f * (-3.0f - tempc);
                else if (tempc > -8.0f && tempc < -5.0f) tf = 0.33333333f * (8.0f + tempc);
                pni_ihm = 3.5E8f * tf * prg_gcw;
                pri_ihm = xm0i * pni_ihm;
                prs_ihm = prs_scw / (prs_scw + prg_gcw) * pri_ihm;
                prg_ihm = prg_gcw / (prs_scw + prg_gcw) * pri_ihm;
            }
            if (prs_scw > 5.0f * prs_sde && prs_sde > eps) {                 // :1942-1949
                const float r_frac = (float)fmin(30.0, prs_scw / prs_sde);
                const float g_frac = fminf(0.75f, 0.05f + (r_frac - 5.f) * .028f);
                vts_boost = fminf(1.5f, 1.1f + (r_frac - 5.f) * .016f);
                prg_scw = g_frac * prs_scw;
                prs_scw = (1.f - g_frac) * prs_scw;
            }
        } else {                                                             // :1951-2007
            if (L_qs) {
                prr_sml = (tempc * tcond - lvap0 * diffu * delQvs) * (c.t1_qs_me * smo1 + c.t2_qs_me * rhof2 * vsc2 * smof);
                prr_sml = prr_sml + 4218.f * c.olfus * tempc * (prr_rcs + prs_scw);
                prr_sml = fmin((double)(rs * odts), fmax(0., prr_sml));
                pnr_sml = smo0 / rs * prr_sml * fpow(10.0f, (-0.75f * tempc));
                pnr_sml = fmin((double)(smo0 * odts), pnr_sml);
                if (tempc > 3.5f || rs < 0.005E-3f) pnr_sml = 0.0;
                if (ssati < 0.f) {
                    prs_sde = c.o.C_cubes * t1_subl * diffu * ssati * rvs * (c.t1_qs_sd * smo1 + c.t2_qs_sd * rhof2 * vsc2 * smof);
                    prs_sde = fmax((double)(-rs * odts), prs_sde);
                }
            }
            if (L_qg) {
                const double ig10 = dpow(ilamg, (double)c.cge[9]), ig11 = dpow(ilamg, (double)c.cge[10]);
                prr_gml = (tempc * tcond - lvap0 * diffu * delQvs) * N0_g *
                          (c.t1_qg_me * ig10 + c.t2_qg_me * rhof2 * vsc2 * ig11);
                prr_gml = fmin((double)(rg * odts), fmax(0., prr_gml));
                pnr_gml = N0_g * c.cgg[1] * dpow(ilamg, (double)c.cge[1]) / rg * prr_gml * fpow(10.0f, (-1.5f * tempc));
                if (tempc > 7.5f || rg < 0.005E-3f) pnr_gml = 0.0;
                if (ssati < 0.f) {
                    prg_gde = C_cube * t1_subl * diffu * ssati * rvs * N0_g *
                              (c.t1_qg_sd * ig10 + c.t2_qg_sd * vsc2 * rhof2 * ig11);
                    prg_gde = fmax((double)(-rg * odts), prg_gde);
                }
            }
            if (dt > 120.f) {
                prr_rcw = prr_rcw + prs_scw + prg_gcw;
                prs_scw = 0.; prg_gcw = 0.;
            }
        }

        // ---------------------------------------------------------- conservation (:2015-2110)
        {
            float sump = (float)(pri_inu + pri_ide + prs_ide + prs_sde + prg_gde);
            float rate_max = (qv - qvsi) * odts * 0.999f;
            if ((sump > eps && sump > rate_max) || (sump < -eps && sump < rate_max)) {
                const float ratio = rate_max / sump;
                pri_inu *= ratio; pri_ide *= ratio; pni_ide *= ratio; prs_ide *= ratio; prs_sde *= ratio; prg_gde *= ratio;
            }
            sump = (float)(-prr_wau - pri_wfz - prr_rcw - prs_scw - prg_scw - prg_gcw);
            rate_max = -rc * odts;
            if (sump < rate_max && L_qc) {
                const float ratio = rate_max / sump;
                prr_wau *= ratio; pri_wfz *= ratio; prr_rcw *= ratio; prs_scw *= ratio; prg_scw *= ratio; prg_gcw *= ratio;
            }
            sump = (float)(pri_ide - prs_iau - prs_sci - pri_rci);
            rate_max = -ri * odts;
            if (sump < rate_max && L_qi) {
                const float ratio = rate_max / sump;
                pri_ide *= ratio; prs_iau *= ratio; prs_sci *= ratio; pri_rci *= ratio;
            }
            sump = (float)(-prg_rfz - pri_rfz - prr_rci + prr_rcs + prr_rcg);
            rate_max = -rr * odts;
            if (sump < rate_max && L_qr) {
                const float ratio = rate_max / sump;
                prg_rfz *= ratio; pri_rfz *= ratio; prr_rci *= ratio; prr_rcs *= ratio; prr_rcg *= ratio;
            }
            sump = (float)(prs_sde - prs_ihm - prr_sml + prs_rcs);
            rate_max = -rs * odts;
            if (sump < rate_max && L_qs) {
                const float ratio = rate_max / sump;
                prs_sde *= ratio; prs_ihm *= ratio; prr_sml *= ratio; prs_rcs *= ratio;
            }
            sump = (float)(prg_gde - prg_ihm - prr_gml + prg_rcg);
            rate_max = -rg * odts;
            if (sump < rate_max && L_qg) {
                const float ratio = rate_max / sump;
                prg_gde *= ratio; prg_ihm *= ratio; prr_gml *= ratio; prg_rcg *= ratio;
            }
            pri_ihm = prs_ihm + prg_ihm;                                     // :2100-2108
            float ratio = (float)fmin(fabs(prr_rcg), fabs(prg_rcg));
            prr_rcg = ratio * copysignf(1.0f, (float)prr_rcg);
            prg_rcg = -prr_rcg;
            if (temp > T_0) {
                ratio = (float)fmin(fabs(prr_rcs), fabs(prs_rcs));
                prr_rcs = ratio * copysignf(1.0f, (float)prr_rcs);
                prs_rcs = -prr_rcs;
            }
        }

        // ---------------------------------------------------------- tendencies (:2116-2236)
        {
            const float orho = 1.f / rho;
            const float lfus2 = lsub - lvap;
            qvten = (float)(qvten + (-pri_inu - pri_ide - prs_ide - prs_sde - prg_gde) * orho);
            qcten = (float)(qcten + (-prr_wau - pri_wfz - prr_rcw - prs_scw - prg_scw - prg_gcw) * orho);
            qiten = (float)(qiten + (pri_inu + pri_ihm + pri_wfz + pri_rfz + pri_ide - prs_iau - prs_sci - pri_rci) * orho);
            niten = (float)(niten + (pni_inu + pni_ihm + pni_wfz + pni_rfz + pni_ide - pni_iau - pni_sci - pni_rci) * orho);
            const float xri = fmaxf(R1, (qi1d + qiten * dtsave) * rho);
            float xni = fmaxf(R2, (ni1d + niten * dtsave) * rho);
            if (xri > R1) {
                double lami = fpow((c.am_i * c.cig[1] * c.oig1 * xni / xri), c.obmi);
                const double ilami = 1. / lami;
                const float xDi = (float)((bm_i + mu_i + 1.f) * ilami);
                if (xDi < 20.E-6f) {
                    lami = c.cie[1] / 20.E-6f;
                    xni = (float)fmin(250.E3, c.cig[0] * c.oig2 * xri / c.am_i * dpow(lami, (double)bm_i));
                    niten = (xni - ni1d * rho) * odts * orho;
                } else if (xDi > 300.E-6f) {
                    lami = c.cie[1] / 300.E-6f;
                    xni = (float)(c.cig[0] * c.oig2 * xri / c.am_i * dpow(lami, (double)bm_i));
                    niten = (xni - ni1d * rho) * odts * orho;
                }
            } else {
                niten = -ni1d * odts;
            }
            xni = fmaxf(0.f, (ni1d + niten * dtsave) * rho);
            if (xni > 250.E3f) niten = (250.E3f - ni1d * rho) * odts * orho;

            qrten = (float)(qrten + (prr_wau + prr_rcw + prr_sml + prr_gml + prr_rcs + prr_rcg - prg_rfz - pri_rfz - prr_rci) * orho);
            nrten = (float)(nrten + (pnr_wau + pnr_sml + pnr_gml - (pnr_rfz + pnr_rcr + pnr_rcg + pnr_rcs + pnr_rci)) * orho);
            const float xrr = fmaxf(R1, (qr1d + qrten * dtsave) * rho);
            float xnr = fmaxf(R2, (nr1d + nrten * dtsave) * rho);
            if (xrr > R1) {
                bool cl;
                mvd_r = rain_balance(c, xrr, xnr, cl);
                if (cl) nrten = (xnr - nr1d * rho) * odts * orho;
            } else {
                qrten = -qr1d * odts;
                nrten = -nr1d * odts;
            }
            qsten = (float)(qsten + (prs_iau + prs_sde + prs_sci + prs_scw + prs_rcs + prs_ide - prs_ihm - prr_sml) * orho);
            qgten = (float)(qgten + (prg_scw + prg_rfz + prg_gde + prg_rcg + prg_gcw + prg_rci + prg_rcs - prg_ihm - prr_gml) * orho);
            if (temp < T_0) {
                tten = (float)(tten + (lsub * ocp * (pri_inu + pri_ide + prs_ide + prs_sde + prg_gde) +
                                       lfus2 * ocp * (pri_wfz + pri_rfz + prg_rfz + prs_scw + prg_scw + prg_gcw + prg_rcs +
                                                      prs_rcs + prr_rci + prg_rcg)) * orho * 1);
            } else {
                tten = (float)(tten + (c.lfus * ocp * (-prr_sml - prr_gml - prr_rcg - prr_rcs) + lsub * ocp * (prs_sde + prg_gde)) * orho * 1);
            }
        }
    }

    // -------------------------------------------------------------- state at tau+1 (:2241-2318)
    float lvt2 = 0.f;
    if (active) {
        temp = t1d + dt * tten;
        const float otemp = 1.f / temp;
        tempc = temp - 273.15f;
        qv = fmaxf(1.E-10f, qv1d + dt * qvten);
        rho = 0.622f * pres / (RR2 * temp * (qv + 0.622f));
        rhof = sqrtf(c.rho_not / rho);
        rhof2 = sqrtf(rhof);
        qvs = rslf(pres, temp);
        ssatw = qv / qvs - 1.f;
        if (fabsf(ssatw) < eps) ssatw = 0.0f;
        diffu = 2.11E-5f * fpow((temp / 273.15f), 1.94f) * (101325.f / pres);
        visco = (tempc >= 0.0f) ? (1.718f + 0.0049f * tempc) * 1.0E-5f
                                : (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
        vsc2 = sqrtf(rho / visco);
        lvap = lvap0 + (2106.0f - 4218.0f) * tempc;
        tcond = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
        ocp = 1.f / (Cp2 * (1.f + 0.887f * qv));
        lvt2 = lvap * lvap * ocp * c.oRv * otemp * otemp;
        if ((qc1d + qcten * dt) > R1) { rc = (qc1d + qcten * dt) * rho; L_qc = true; } else { rc = R1; L_qc = false; }
        if ((qi1d + qiten * dt) > R1) {
            ri = (qi1d + qiten * dt) * rho;
            ni = fmaxf(R2, (ni1d + niten * dt) * rho);
            L_qi = true;
        } else { ri = R1; ni = R2; L_qi = false; }
        if ((qr1d + qrten * dt) > R1) {
            rr = (qr1d + qrten * dt) * rho;
            nr = fmaxf(R2, (nr1d + nrten * dt) * rho);
            L_qr = true;
            bool cl;
            mvd_r = rain_balance(c, rr, nr, cl);
        } else { rr = R1; nr = R2; L_qr = false; }
        if ((qs1d + qsten * dt) > R1) { rs = (qs1d + qsten * dt) * rho; L_qs = true; } else { rs = R1; L_qs = false; }
        if ((qg1d + qgten * dt) > R1) { rg = (qg1d + qgten * dt) * rho; L_qg = true; } else { rg = R1; L_qg = false; }

        if (L_qs) {                                                          // :2325-2374
            const float tc0 = fminf(-0.1f, temp - 273.15f);
            smob = rs * c.oams;
            if (bm_s > (2.0f - 1.e-3f) && bm_s < (2.0f + 1.e-3f)) smo2 = smob;
            else {
                const float loga_ = sa_poly(k_sa, tc0, bm_s), a_ = fpow(10.0f, loga_), b_ = sa_poly(k_sb, tc0, bm_s);
                smo2 = fpow((smob / a_), (1.f / b_));
            }
            smoc = snow_moment(tc0, c.cse[0], smo2);
            smod = snow_moment(tc0, c.cse[13], smo2);
        }
    }
    graupel_rain_slopes();                                                   // :2379-2408

    if (active) {
        // ---------------------------------------------------------- condensation (:2414-2440)
        if ((ssatw > eps) || (ssatw < -eps && L_qc)) {
            float clap = (qv - qvs) / (1.f + lvt2 * qvs);
            for (int n = 1; n <= 3; ++n) {
                const float fcd = qvs * fexp(lvt2 * clap) - qv + clap;
                const float dfcd = qvs * lvt2 * fexp(lvt2 * clap) + 1.f;
                clap = clap - fcd / dfcd;
            }
            const float xrc = rc + clap;
            if (xrc > 0.0f) prw_vcd = clap * odt;
            else prw_vcd = -rc / rho * odts;
            qcten = (float)(qcten + prw_vcd);
            qvten = (float)(qvten - prw_vcd);
            tten = (float)(tten + lvap * ocp * prw_vcd * 1);
            rc = fmaxf(R1, (qc1d + dt * qcten) * rho);
            qv = fmaxf(1.E-10f, qv1d + dt * qvten);
            temp = t1d + dt * tten;
            rho = 0.622f * pres / (RR2 * temp * (qv + 0.622f));
            qvs = rslf(pres, temp);
            ssatw = qv / qvs - 1.f;
        }
        // ---------------------------------------------------------- rain evaporation (:2446-2505)
        if ((ssatw < -eps) && L_qr && (!(prw_vcd > 0.))) {
            tempc = temp - 273.15f;
            const float otemp = 1.f / temp;
            rhof = sqrtf(c.rho_not / rho);
            rhof2 = sqrtf(rhof);
            diffu = 2.11E-5f * fpow((temp / 273.15f), 1.94f) * (101325.f / pres);
            visco = (tempc >= 0.0f) ? (1.718f + 0.0049f * tempc) * 1.0E-5f
                                    : (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
            vsc2 = sqrtf(rho / visco);
            lvap = lvap0 + (2106.0f - 4218.0f) * tempc;
            tcond = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
            ocp = 1.f / (Cp2 * (1.f + 0.887f * qv));
            rvs = rho * qvs;
            const float rvs_p = rvs * otemp * (lvap * otemp * c.oRv - 1.f);
            const float rvs_pp = rvs * (otemp * (lvap * otemp * c.oRv - 1.f) * otemp * (lvap * otemp * c.oRv - 1.f) +
                                        (-2.f * lvap * otemp * otemp * otemp * c.oRv) + otemp * otemp);
            const float gamsc = lvap * diffu / tcond * rvs_p;
            float alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
            alphsc = fmaxf(1.E-9f, alphsc);
            const float xsat = fminf(-1.E-9f, ssatw);
            const float t1_evap = 2.f * PI2 * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat -
                                               5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);
            const double lamr = 1. / ilamr;
            if (qv / qvs < 0.95f && rr / rho <= 1.E-8f) {
                prv_rev = rr / rho * odts;
            } else {
                prv_rev = t1_evap * diffu * (-ssatw) * N0_r * rvs *
                          (c.t1_qr_ev * dpow(ilamr, (double)c.cre[9]) + c.t2_qr_ev * vsc2 * rhof2 * (dpow((lamr + 0.5f * fv_r), (double)(-c.cre[10]))));
                const float rate_max = fminf((rr / rho * odts), (qvs - qv) * odts);
                prv_rev = fmin((double)rate_max, prv_rev / rho);
            }
            pnr_rev = fmin((double)(nr * 0.99f / rho * odts), prv_rev * nr / rr);
            qrten = (float)(qrten - prv_rev);
            qvten = (float)(qvten + prv_rev);
            nrten = (float)(nrten - pnr_rev);
            tten = (float)(tten - lvap * ocp * prv_rev * 1);
            rr = fmaxf(R1, (qr1d + dt * qrten) * rho);
            qv = fmaxf(1.E-10f, qv1d + dt * qvten);
            nr = fmaxf(R2, (nr1d + dt * nrten) * rho);
            temp = t1d + dt * tten;
            rho = 0.622f * pres / (RR2 * temp * (qv + 0.622f));
        }
    }

    // -------------------------------------------------------------- fall speeds (:2515-2650)
    // planes of s_f: 0 vtrk, 1 vtnrk, 2 vtik, 3 vtnik, 4 vtsk, 5 vtgk (each with a zero "level kte+1"),
    // 6 = validity bits of this level (rain 1, ice 2, snow 4, graupel 8) as float bit pattern
    // s_i rows: 0 active flag, 1..4 nstep per species, 5..8 ksed1 per species (1-based level)
    float vtrk = 0.f, vtnrk = 0.f, vtik = 0.f, vtnik = 0.f, vtsk = 0.f, vtgk = 0.f;
    if (active) {
        rhof = sqrtf(c.rho_not / rho);
        if (rr > R1) {
            const double lamr = fpow((c.am_r * c.crg[2] * c.org2 * nr / rr), c.obmr);
            vtrk = (float)(rhof * av_r * c.crg[5] * c.org3 * dpow(lamr, (double)c.cre[2]) * (dpow((lamr + fv_r), (double)(-c.cre[5]))));
            vtnrk = (float)(rhof * av_r * c.crg[6] / c.crg[11] * dpow(lamr, (double)c.cre[11]) * (dpow((lamr + fv_r), (double)(-c.cre[6]))));
        }
        if (ri > R1) {
            const double lami = fpow((c.am_i * c.cig[1] * c.oig1 * ni / ri), c.obmi);
            const double ilami = 1. / lami;
            vtik = (float)(rhof * c.o.av_i * c.cig[2] * c.oig2 * dpow(ilami, (double)bv_i));
            vtnik = (float)(rhof * c.o.av_i * c.cig[5] / c.cig[6] * dpow(ilami, (double)bv_i));
        }
    }
    __syncthreads();
    if (kk < nk) {
        s_f[0 * PL + kk * TX + tx] = vtrk; s_f[1 * PL + kk * TX + tx] = vtnrk;
        s_f[2 * PL + kk * TX + tx] = vtik; s_f[3 * PL + kk * TX + tx] = vtnik;
        unsigned bits = 0u;
        if (active) bits = (rr > R1 ? 1u : 0u) | (ri > R1 ? 2u : 0u) | (rs > R1 ? 4u : 0u) | (rg > R1 ? 8u : 0u);
        s_f[6 * PL + kk * TX + tx] = __uint_as_float(bits);
    }
    __syncthreads();
    // a level without the species takes the value of the nearest level above that has it (else 0)
    auto inherit = [&](int plane, unsigned bit, float own) -> float {
#pragma unroll 1
        for (int q = kk; q < nk; ++q)
            if (__float_as_uint(s_f[6 * PL + q * TX + tx]) & bit) return (q == kk) ? own : s_f[plane * PL + q * TX + tx];
        return 0.f;
    };
    if (active) {
        vtrk = inherit(0, 1u, vtrk); vtnrk = inherit(1, 1u, vtnrk);
        vtik = inherit(2, 2u, vtik); vtnik = inherit(3, 2u, vtnik);
        if (rs > R1) {                                                       // :2595-2611
            const float xDs = smoc / smob;
            const float Mrat = 1.f / xDs;
            float ils1 = 1.f / (Mrat * Lam0 + c.o.fv_s);
            float ils2 = 1.f / (Mrat * Lam1 + c.o.fv_s);
            const float t1_vts = Kap0 * c.csg[3] * fpow(ils1, c.cse[3]);
            const float t2_vts = Kap1 * fpow(Mrat, mu_s) * c.csg[9] * fpow(ils2, c.cse[9]);
            ils1 = 1.f / (Mrat * Lam0);
            ils2 = 1.f / (Mrat * Lam1);
            const float t3_vts = Kap0 * c.csg[0] * fpow(ils1, c.cse[0]);
            const float t4_vts = Kap1 * fpow(Mrat, mu_s) * c.csg[6] * fpow(ils2, c.cse[6]);
            const float vts = rhof * c.o.av_s * (t1_vts + t2_vts) / (t3_vts + t4_vts);
            vtsk = (temp > T_0) ? fmaxf(vts * vts_boost, vtrk) : vts * vts_boost;
        }
        if (rg > R1) {                                                       // :2631-2637
            const float vtg = (float)(rhof * c.o.av_g * c.cgg[5] * c.ogg3 * dpow(ilamg, (double)c.o.bv_g));
            vtgk = (temp > T_0) ? fmaxf(vtg, vtrk) : vtg;
        }
    }
    __syncthreads();
    if (kk < nk) { s_f[4 * PL + kk * TX + tx] = vtsk; s_f[5 * PL + kk * TX + tx] = vtgk; }
    __syncthreads();
    if (active) {
        vtsk = inherit(4, 4u, vtsk);
        vtgk = inherit(5, 8u, vtgk);
        // nstep and ksed1 per species: column maxima
        const float v4[4] = {fmaxf(vtrk, vtnrk), vtik, vtsk, vtgk};
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (v4[s] > 1.E-3f) {
                const float delta_tp = dzq / v4[s];
                atomicMax(&s_i[(1 + s) * TX + tx], (int)(dt / delta_tp + 1.f));
                atomicMax(&s_i[(5 + s) * TX + tx], kk + 1);
            }
    }
    __syncthreads();

    // -------------------------------------------------------------- sedimentation (:2660-2770)
    float ppt[4] = {0.f, 0.f, 0.f, 0.f};      // rain, ice, snow, graupel (accumulated by the kts thread)
    float onstep[4];
    int nstep[4], ksed1[4], nmax[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int ns = s_i[(1 + s) * TX + tx];
        onstep[s] = (ns > 0) ? 1.f / (float)ns : 1.0f;
        nstep[s] = active ? nint_f(1.f / onstep[s]) : 0;
        int ks = max(1, s_i[(5 + s) * TX + tx]);
        if (ks == nk) ks = nk - 1;
        ksed1[s] = ks;
        // uniform trip count for the CTA: the largest sub-step count of its columns
        int m = col_ok && (s_i[tx] != 0) ? nint_f(1.f / onstep[s]) : 0;
        m = __reduce_max_sync(0xffffffffu, m);             // every warp holds all TX columns (TX divides 32)
        nmax[s] = m;
    }
    const float odzq = 1.f / dzq;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        for (int n = 1; n <= nmax[s]; ++n) {
            // sed(k) from the state before this sub-step; planes 0/1 reused as sed / sed_n
            float& r_ = (s == 0) ? rr : (s == 1) ? ri : (s == 2) ? rs : rg;
            float& n_ = (s == 0) ? nr : ni;                                   // only for s < 2
            const float vm = (s == 0) ? vtrk : (s == 1) ? vtik : (s == 2) ? vtsk : vtgk;
            const float vn = (s == 0) ? vtnrk : vtnik;
            const float sed = vm * r_;
            const float sedn = (s < 2) ? vn * n_ : 0.f;
            __syncthreads();
            if (kk < nk) { s_f[0 * PL + kk * TX + tx] = sed; if (s < 2) s_f[1 * PL + kk * TX + tx] = sedn; }
            __syncthreads();
            if (active && n <= nstep[s]) {
                const float orho = 1.f / rho;
                float& qten = (s == 0) ? qrten : (s == 1) ? qiten : (s == 2) ? qsten : qgten;
                float& nten = (s == 0) ? nrten : niten;
                if (kk == nk - 1) {
                    qten = qten - sed * odzq * onstep[s] * orho;
                    if (s < 2) nten = nten - sedn * odzq * onstep[s] * orho;
                    r_ = fmaxf(R1, r_ - sed * odzq * dt * onstep[s]);
                    if (s < 2) n_ = fmaxf(R2, n_ - sedn * odzq * dt * onstep[s]);
                }
                if (kk <= ksed1[s] - 1) {
                    const float sed_up = s_f[0 * PL + (kk + 1) * TX + tx];
                    qten = qten + (sed_up - sed) * odzq * onstep[s] * orho;
                    if (s < 2) {
                        const float sedn_up = s_f[1 * PL + (kk + 1) * TX + tx];
                        nten = nten + (sedn_up - sedn) * odzq * onstep[s] * orho;
                        n_ = fmaxf(R2, n_ + (sedn_up - sedn) * odzq * dt * onstep[s]);
                    }
                    r_ = fmaxf(R1, r_ + (sed_up - sed) * odzq * dt * onstep[s]);
                }
                if (kk == 0 && r_ > R1 * 10.f) ppt[s] = ppt[s] + sed * dt * onstep[s];
            }
        }
    }

    // -------------------------------------------------------------- instant melt / freeze, apply (:2776-2842)
    if (active) {
        const float xri = fmaxf(0.0f, qi1d + qiten * dt);
        if ((temp > T_0) && (xri > 0.0f)) {
            qcten = qcten + xri * odt;
            qiten = qiten - xri * odt;
            niten = -ni1d * odt;
            tten = tten - c.lfus * ocp * xri * odt * 1;
        }
        const float xrc = fmaxf(0.0f, qc1d + qcten * dt);
        if ((temp < HGFR) && (xrc > 0.0f)) {
            const float lfus2 = lsub - lvap;
            qiten = qiten + xrc * odt;
            niten = niten + xrc / xm0i * odt;
            qcten = qcten - xrc * odt;
            tten = tten + lfus2 * ocp * xrc * odt * 1;
        }
        t1d = t1d + tten * dt;
        qv1d = fmaxf(1.E-10f, qv1d + qvten * dt);
        qc1d = qc1d + qcten * dt;
        if (qc1d <= R1) qc1d = 0.0f;
        qi1d = qi1d + qiten * dt;
        ni1d = fmaxf(R2 / rho, ni1d + niten * dt);
        if (qi1d <= R1) { qi1d = 0.0f; ni1d = 0.0f; }
        else {
            double lami = fpow((c.am_i * c.cig[1] * c.oig1 * ni1d / qi1d), c.obmi);
            const double ilami = 1. / lami;
            const float xDi = (float)((bm_i + mu_i + 1.f) * ilami);
            if (xDi < 20.E-6f) lami = c.cie[1] / 20.E-6f;
            else if (xDi > 300.E-6f) lami = c.cie[1] / 300.E-6f;
            ni1d = (float)fmin(c.cig[0] * c.oig2 * qi1d / c.am_i * dpow(lami, (double)bm_i), 250.E3 / rho);
        }
        qr1d = qr1d + qrten * dt;
        nr1d = fmaxf(R2 / rho, nr1d + nrten * dt);
        if (qr1d <= R1) { qr1d = 0.0f; nr1d = 0.0f; }
        else {
            double lamr = fpow((c.am_r * c.crg[2] * c.org2 * nr1d / qr1d), c.obmr);
            float mv = (float)((3.0f + c.o.mu_r + 0.672f) / lamr);
            if (mv > 2.5E-3f) mv = 2.5E-3f;
            else if (mv < D0r * 0.75f) mv = D0r * 0.75f;
            lamr = (3.0f + c.o.mu_r + 0.672f) / mv;
            nr1d = (float)(c.crg[1] * c.org3 * qr1d * dpow(lamr, (double)bm_r) / c.am_r);
        }
        qs1d = qs1d + qsten * dt;
        if (qs1d <= R1) qs1d = 0.0f;
        qg1d = qg1d + qgten * dt;
        if (qg1d <= R1) qg1d = 0.0f;
    }

    // -------------------------------------------------------------- scatter (:908-926, :997-1010)
    if (live) {
        // (:997-1010: every branch of the water-vapour floor ends in 1.E-7 once qv1d < 1.E-7)
        A.qv[a] = (qv1d < 1.E-7f) ? 1.E-7f : qv1d;
        A.qc[a] = qc1d; A.qi[a] = qi1d; A.qr[a] = qr1d; A.qs[a] = qs1d; A.qg[a] = qg1d; A.ni[a] = ni1d; A.nr[a] = nr1d;
        A.th[a] = t1d / pii;
        if (kk == 0) {
            const float pptrain = ppt[0], pptice = ppt[1], pptsnow = ppt[2], pptgraul = ppt[3];
            const size_t a2 = (size_t)i + (size_t)ld2 * (size_t)j;
            if (A.acc_precip) {        // the per-call arrays start at zero (mp_driver.f90:411-413): 0 + ppt..., then REAL*8 +=
                A.acc_precip[a2] = acc0[0] + (double)(0.f + pptrain + pptsnow + pptgraul + pptice);
                A.acc_snow[a2] = acc0[1] + (double)(0.f + pptsnow + pptice);
                A.acc_graupel[a2] = acc0[2] + (double)(0.f + pptgraul);
            }
            if (A.rainnc) A.rainnc[a2] = A.rainnc[a2] + pptrain + pptsnow + pptgraul + pptice;
            if (A.snownc) A.snownc[a2] = A.snownc[a2] + pptsnow + pptice;
            if (A.snowncv) A.snowncv[a2] = pptsnow + pptice;
            if (A.graupelnc) A.graupelnc[a2] = A.graupelnc[a2] + pptgraul;
            if (A.graupelncv) A.graupelncv[a2] = pptgraul;
            if (A.sr) A.sr[a2] = (pptsnow + pptgraul + pptice) / ((A.rainncv ? A.rainncv[a2] : 0.f) + 1.e-12f);
        }
    }
}

}  // namespace

int th_driver_launch(const Geom& g, const ThConst& c, const ThTables& tb, float* qv, float* qc, float* qr, float* qi,
                     float* qs, float* qg, float* ni, float* nr, float* th, const float* pii, const float* p,
                     const float* dz, float dt, float* rainnc, const float* rainncv, float* snownc, float* snowncv,
                     float* graupelnc, float* graupelncv, float* sr, int ld2, const MpTiles& tiles, int kts, int kte,
                     unsigned long long* diag, cudaStream_t st, double* acc_precip, double* acc_snow,
                     double* acc_graupel) {
    int gx = 0, gy = 0;
    for (int t = 0; t < tiles.n; ++t) {
        if (tiles.ite[t] < tiles.its[t] || tiles.jte[t] < tiles.jts[t]) continue;
        const int w = tiles.ite[t] - tiles.its[t] + 1, h = tiles.jte[t] - tiles.jts[t] + 1;
        gx = std::max(gx, ((tiles.along_y[t] ? h : w) + TX - 1) / TX);
        gy = std::max(gy, tiles.along_y[t] ? w : h);
    }
    if (gx == 0 || gy == 0) return 0;
    const int nk = kte - kts + 1;
    if (nk < 1 || nk * TX > 1024) return -1;                // one thread per level
    const int nkt = (nk * TX + 31) / 32 * 32 / TX;          // whole warps
    Args A{qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz, rainnc, snownc, snowncv, graupelnc, graupelncv, sr, rainncv,
           acc_precip, acc_snow, acc_graupel};
    dim3 block(TX, nkt), grid(gx, gy, tiles.n);
    const size_t smem = (size_t)nk * TX * sizeof(double) + (size_t)7 * (nk + 1) * TX * sizeof(float) + (size_t)10 * TX * sizeof(int);
    const int nthreads = TX * nkt;
#define TH_LAUNCH(MAXT, MINB) k_mp_thompson<MAXT, MINB><<<grid, block, smem, st>>>(g, c, tb, A, ld2, dt, tiles, kts, kte, diag)
    if (nthreads <= 320) TH_LAUNCH(320, 2);
    else if (nthreads <= 640) TH_LAUNCH(640, 1);
    else TH_LAUNCH(1024, 1);
#undef TH_LAUNCH
    return 1;
}

}  // namespace icar
