// thompson_init (src/physics/mp_thompson.f90:342-766) for the B200 path: the scalar constants and bin
// grids are computed on the host, the 29 lookup tables by device kernels -- the two collision
// tables (qr_acr_qg :2853-3007, qr_acr_qs :3014-3264; 1.4e10 bin-pair evaluations, minutes on a CPU,
// which is why the reference caches them in *_mpt.dat files) take a few hundred milliseconds here,
// so there is no cache file.
//
// Layout: every table is REAL*8 in the Fortran's column-major order, i.e. byte-compatible with the
// records of the reference's qr_acr_qg_mpt.dat / qr_acr_qs_mpt.dat / freezeH2O_mpt.dat.
//
// Numerics: each table entry is a sum over 100 x 100 bin pairs.  One CTA owns one entry; the pair
// sum is a tree reduction, so entries differ from a sequential Fortran sum by a few 1e-16 relative.
#include <cmath>

#include "mp_thompson.cuh"

namespace icar {

namespace {

constexpr float PI2 = 3.1415926536f;                 // :54
constexpr float rho_w = 1000.0f, rho_i = 890.0f;     // :57-60
constexpr float mu_g = 0.0f, mu_i = 0.0f, mu_s = 0.6357f, Kap0 = 490.6f, Kap1 = 17.46f, Lam0 = 20.78f, Lam1 = 3.29f;
constexpr float bm_r = 3.0f, bm_s = 2.0f, bm_g = 3.0f, bm_i = 3.0f, av_r = 4854.0f, bv_r = 1.0f, bv_i = 1.0f;
constexpr float C_cube = 0.5f, Sc = 0.632f, Rv = 461.5f, lsub = 2.834E6f, lvap0 = 2.5E6f;
constexpr float xm0i = 1.E-12f, D0c = 1.E-6f, D0r = 50.E-6f, D0s = 200.E-6f, D0g = 250.E-6f;

// GAMMLN / WGAMMA, :3718-3771
float gammln_h(float xx) {
    static const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                  -1.231739572450155, .1208650973866179E-2, -.5395239384953E-5};
    double x = xx, y = x, tmp = x + 5.5;
    tmp = (x + 0.5) * std::log(tmp) - tmp;
    double ser = 1.000000000190015;
    for (double c : cof) { y += 1.0; ser += c / y; }
    return (float)(tmp + std::log(2.5066282746310005 * ser / x));
}
float wgamma_h(float y) { return std::exp(gammln_h(y)); }

// value d x 10^e of a decade table as the REAL literal the Fortran source spells (:204-268)
__host__ __device__ inline float decade_value(int idx0, int first_exp) {
    // d * 10^e in double is within half an ulp(double) of the decimal value, far inside the float
    // rounding interval of every entry of these tables, so one conversion gives the literal's float.
    const int e = idx0 / 9 + first_exp, d = idx0 % 9 + 1;
    double v = (double)d;
    if (e >= 0) { for (int n = 0; n < e; ++n) v *= 10.0; }
    else { double p = 1.0; for (int n = 0; n < -e; ++n) p *= 10.0; v = v / p; }
    return (float)v;
}

}  // namespace

size_t th_table_elems(int id) {
    if (id <= T_tnr_gacr) return (size_t)ntb_g1 * ntb_g * ntb_r1 * ntb_r;
    if (id <= T_tnr_sacr2) return (size_t)ntb_s * ntb_t * ntb_r1 * ntb_r;
    if (id <= T_tni_qcfz) return (size_t)ntb_c * 45;
    if (id <= T_tnr_qrfz) return (size_t)ntb_r * ntb_r1 * 45;
    if (id <= T_tpi_ide) return (size_t)ntb_i * ntb_i1;
    return (size_t)nbins * nbins;
}

void th_make_constants(const ThOptions& o, ThConst* cp, double* bins) {
    ThConst& c = *cp;
    c.o = o;
    c.am_g = PI2 * o.rho_g / 6.0f;                     // :417
    c.am_r = PI2 * rho_w / 6.0f;                       // :97
    c.am_i = PI2 * rho_i / 6.0f;                       // :103
    c.lfus = lsub - lvap0; c.olfus = 1.f / c.lfus; c.oRv = 1.f / Rv;      // :157,:166-167
    c.rho_not = 101325.0f / (287.05f * 298.0f);        // :146
    c.mu_c = std::fmin(15.f, (1000.E6f / o.Nt_c + 2.f));                  // :424
    c.Sc3 = std::pow(Sc, (1.f / 3.f));
    c.D0i = std::pow((xm0i / c.am_i), (1.f / bm_i));
    c.xm0s = o.am_s * std::pow(D0s, bm_s);
    c.xm0g = c.am_g * std::pow(D0g, bm_g);
    const float bv_s = o.bv_s, bv_g = o.bv_g, mu_r = o.mu_r, mu_c = c.mu_c;
    // exponent lists :436-531 and their gamma functions
    const float cce[3] = {mu_c + 1.f, bm_r + mu_c + 1.f, bm_r + mu_c + 4.f};
    const float cie[7] = {mu_i + 1.f, bm_i + mu_i + 1.f, bm_i + mu_i + bv_i + 1.f, mu_i + bv_i + 1.f, mu_i + 2.f,
                          bm_i * 0.5f + mu_i + bv_i + 1.f, bm_i * 0.5f + mu_i + 1.f};
    const float cre[13] = {bm_r + 1.f, mu_r + 1.f, bm_r + mu_r + 1.f, bm_r * 2.f + mu_r + 1.f, mu_r + bv_r + 1.f,
                           bm_r + mu_r + bv_r + 1.f, bm_r * 0.5f + mu_r + bv_r + 1.f, bm_r + mu_r + bv_r + 3.f,
                           mu_r + bv_r + 3.f, mu_r + 2.f, 0.5f * (bv_r + 5.f + 2.f * mu_r), bm_r * 0.5f + mu_r + 1.f,
                           bm_r * 2.f + mu_r + bv_r + 1.f};
    float cse[18] = {bm_s + 1.f, bm_s + 2.f, bm_s * 2.f, bm_s + bv_s + 1.f, bm_s * 2.f + bv_s + 1.f, bm_s * 2.f + 1.f,
                     bm_s + mu_s + 1.f, bm_s + mu_s + 2.f, bm_s + mu_s + 3.f, bm_s + mu_s + bv_s + 1.f,
                     bm_s * 2.f + mu_s + bv_s + 1.f, bm_s * 2.f + mu_s + 1.f, bv_s + 2.f, bm_s + bv_s, mu_s + 1.f,
                     1.0f + (1.0f + bv_s) / 2.f, 0.f, bv_s + mu_s + 3.f};
    cse[16] = cse[15] + mu_s + 1.f;
    const float cge[12] = {bm_g + 1.f, mu_g + 1.f, bm_g + mu_g + 1.f, bm_g * 2.f + mu_g + 1.f,
                           bm_g * 2.f + mu_g + bv_g + 1.f, bm_g + mu_g + bv_g + 1.f, bm_g + mu_g + bv_g + 2.f,
                           bm_g + mu_g + bv_g + 3.f, mu_g + bv_g + 3.f, mu_g + 2.f, 0.5f * (bv_g + 5.f + 2.f * mu_g),
                           0.5f * (bv_g + 5.f) + mu_g};
    for (int n = 0; n < 3; ++n) { c.cce[n] = cce[n]; c.ccg[n] = wgamma_h(cce[n]); }
    for (int n = 0; n < 7; ++n) { c.cie[n] = cie[n]; c.cig[n] = wgamma_h(cie[n]); }
    for (int n = 0; n < 13; ++n) { c.cre[n] = cre[n]; c.crg[n] = wgamma_h(cre[n]); }
    for (int n = 0; n < 18; ++n) { c.cse[n] = cse[n]; c.csg[n] = wgamma_h(cse[n]); }
    for (int n = 0; n < 12; ++n) { c.cge[n] = cge[n]; c.cgg[n] = wgamma_h(cge[n]); }
    c.ocg1 = 1.f / c.ccg[0]; c.ocg2 = 1.f / c.ccg[1];
    c.oig1 = 1.f / c.cig[0]; c.oig2 = 1.f / c.cig[1]; c.obmi = 1.f / bm_i;
    c.obmr = 1.f / bm_r; c.ore1 = 1.f / c.cre[0]; c.org1 = 1.f / c.crg[0]; c.org2 = 1.f / c.crg[1]; c.org3 = 1.f / c.crg[2];
    c.oams = 1.f / o.am_s; c.obms = 1.f / bm_s; c.ocms = std::pow(c.oams, c.obms);
    c.oamg = 1.f / c.am_g; c.obmg = 1.f / bm_g; c.ocmg = std::pow(c.oamg, c.obmg);
    c.oge1 = 1.f / c.cge[0]; c.ogg1 = 1.f / c.cgg[0]; c.ogg2 = 1.f / c.cgg[1]; c.ogg3 = 1.f / c.cgg[2];
    // rate prefactors :538-568
    c.t1_qr_qc = PI2 * .25f * av_r * c.crg[8];
    c.t1_qr_qi = PI2 * .25f * av_r * c.crg[8];
    c.t2_qr_qi = PI2 * .25f * c.am_r * av_r * c.crg[7];
    c.t1_qg_qc = PI2 * .25f * o.av_g * c.cgg[8];
    c.t1_qs_qc = PI2 * .25f * o.av_s;
    c.t1_qs_qi = PI2 * .25f * o.av_s;
    c.t1_qr_ev = 0.78f * c.crg[9];
    c.t2_qr_ev = 0.308f * c.Sc3 * std::sqrt(av_r) * c.crg[10];
    c.t1_qs_sd = 0.86f;
    c.t2_qs_sd = 0.28f * c.Sc3 * std::sqrt(o.av_s);
    c.t1_qs_me = PI2 * 4.f * o.C_sqrd * c.olfus * 0.86f;
    c.t2_qs_me = PI2 * 4.f * o.C_sqrd * c.olfus * 0.28f * c.Sc3 * std::sqrt(o.av_s);
    c.t1_qg_sd = 0.86f * c.cgg[9];
    c.t2_qg_sd = 0.28f * c.Sc3 * std::sqrt(o.av_g) * c.cgg[10];
    c.t1_qg_me = PI2 * 4.f * C_cube * c.olfus * 0.86f * c.cgg[9];
    c.t2_qg_me = PI2 * 4.f * C_cube * c.olfus * 0.28f * c.Sc3 * std::sqrt(o.av_g) * c.cgg[10];
    // :571-578  NINT(ALOG10(first entry))
    c.nic2 = -6; c.nii2 = -10; c.nii3 = 0; c.nir2 = -6; c.nir3 = 6; c.nis2 = -5; c.nig2 = -5; c.nig3 = 4;
    // bins :581-634
    double *Dc = bins, *dtc = bins + 100, *Db[4] = {bins + 200, bins + 400, bins + 600, bins + 800};
    Dc[0] = D0c * 1.0; dtc[0] = D0c * 1.0;
    for (int n = 1; n < nbins; ++n) { Dc[n] = Dc[n - 1] + 1.0E-6; dtc[n] = Dc[n] - Dc[n - 1]; }
    const double lo[4] = {c.D0i * 1.0, D0r * 1.0, D0s * 1.0, D0g * 1.0}, hi[4] = {5.0 * D0s, 0.005, 0.02, 0.05};
    for (int b = 0; b < 4; ++b) {
        double x[nbins + 1];
        x[0] = lo[b]; x[nbins] = hi[b];
        for (int n = 2; n <= nbins; ++n)
            x[n - 1] = std::exp((double)(n - 1) / (double)nbins * std::log(x[nbins] / x[0]) + std::log(x[0]));
        for (int n = 0; n < nbins; ++n) { Db[b][n] = std::sqrt(x[n] * x[n + 1]); Db[b][n + 100] = x[n + 1] - x[n]; }
    }
    c.Dr1 = Db[1][0]; c.Drn = Db[1][nbins - 1]; c.Ds1 = Db[2][0]; c.Dsn = Db[2][nbins - 1];
}

namespace {

// ---- device side -----------------------------------------------------------------------------
__device__ __forceinline__ const double* bin_D(const double* bins, int which) { return bins + 200 * which; }   // 0 c,1 i,2 r,3 s,4 g
__device__ __forceinline__ const double* bin_dt(const double* bins, int which) { return bins + 200 * which + 100; }

// rain fall speed polynomial used by the table builders (:2893-2895)
__device__ __forceinline__ double vr_poly(double D) {
    return -0.1021f + 4.932E3f * D - 0.9551E6f * D * D + 0.07934E9f * D * D * D - 0.002362E12f * D * D * D * D;
}

// aux[0..99] vr, [100..] vg, [200..] vs, [300..] massr, [400..] massg, [500..] masss, [600..] massc
__global__ void k_table_aux(ThConst c, const double* __restrict__ bins, double* __restrict__ aux) {
    const int n = threadIdx.x;
    if (n >= nbins) return;
    const double Dr = bin_D(bins, 2)[n], Dg = bin_D(bins, 4)[n], Ds = bin_D(bins, 3)[n], Dc = bin_D(bins, 0)[n];
    aux[n] = vr_poly(Dr);
    aux[100 + n] = c.o.av_g * pow(Dg, (double)c.o.bv_g);                                 // :2898
    aux[200 + n] = 1.5f * c.o.av_s * pow(Ds, (double)c.o.bv_s) * exp(-c.o.fv_s * Ds);     // :3072
    aux[300 + n] = c.am_r * pow(Dr, (double)bm_r);
    aux[400 + n] = c.am_g * pow(Dg, (double)bm_g);
    aux[500 + n] = c.o.am_s * pow(Ds, (double)bm_s);
    aux[600 + n] = c.am_r * pow(Dc, (double)bm_r);
}

// exponential rain distribution of table cell (k = N0r_exp index, m = r_r index): N_r(n2), :2921-2926
__device__ __forceinline__ void rain_dsd(const ThConst& c, int k, int m, double& lamr, double& N0_r) {
    const float N0r = decade_value(k, 6), rr = decade_value(m, -6);
    double lam_exp = powf((N0r * c.am_r * c.crg[0] / rr), c.ore1);
    lamr = lam_exp * powf((c.crg[2] * c.org2 * c.org1), c.obmr);
    N0_r = N0r / (c.crg[1] * lam_exp) * pow(lamr, (double)c.cre[1]);
}

template <int NS>
__device__ __forceinline__ void block_sum(double (&v)[NS], double* s_red) {
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] += __shfl_down_sync(0xffffffffu, v[a], o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0)
#pragma unroll
        for (int a = 0; a < NS; ++a) s_red[a * 8 + w] = v[a];
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            double t = s_red[a * 8];
            for (int x = 1; x < nw; ++x) t += s_red[a * 8 + x];
            v[a] = t;
        }
}

// qr_acr_qg :2917-2975.  blockIdx.x = (j-1)*ntb_g1 + (i-1), blockIdx.y = (m-1)*ntb_r1 + (k-1)
__global__ void __launch_bounds__(128)
k_qr_acr_qg(ThConst c, const double* __restrict__ bins, const double* __restrict__ aux, double* tcg_racg,
            double* tmr_racg, double* tcr_gacr, double* tmg_gacr, double* tnr_racg, double* tnr_gacr) {
    __shared__ double s_Nr[nbins], s_Ng[nbins], s_par[4], s_red[6 * 8];
    const int i = blockIdx.x % ntb_g1, j = blockIdx.x / ntb_g1, k = blockIdx.y % ntb_r1, m = blockIdx.y / ntb_r1;
    const double *Dr = bin_D(bins, 2), *dtr = bin_dt(bins, 2), *Dg = bin_D(bins, 4), *dtg = bin_dt(bins, 4);
    if (threadIdx.x == 0) {
        rain_dsd(c, k, m, s_par[0], s_par[1]);
        const float N0g = decade_value(i, 4), rg = decade_value(j, -5);
        double lam_exp = powf((N0g * c.am_g * c.cgg[0] / rg), c.oge1);                          // :2930-2932
        double lamg = lam_exp * powf((c.cgg[2] * c.ogg2 * c.ogg1), c.obmg);
        s_par[2] = lamg;
        s_par[3] = N0g / (c.cgg[1] * lam_exp) * pow(lamg, (double)c.cge[1]);
    }
    __syncthreads();
    if (threadIdx.x < nbins) {
        const int n = threadIdx.x;
        s_Nr[n] = s_par[1] * pow(Dr[n], (double)c.o.mu_r) * exp(-s_par[0] * Dr[n]) * dtr[n];
        s_Ng[n] = s_par[3] * pow(Dg[n], (double)mu_g) * exp(-s_par[2] * Dg[n]) * dtg[n];
    }
    __syncthreads();
    const float pef = PI2 * .25f * c.o.Ef_rg;
    double acc[6] = {0, 0, 0, 0, 0, 0};      // t1, z1, y1, t2, y2, z2
    for (int p = threadIdx.x; p < nbins * nbins; p += blockDim.x) {
        const int n2 = p / nbins, n = p % nbins;
        const double vr = aux[n2], vg = aux[100 + n], massr = aux[300 + n2], massg = aux[400 + n];
        const double dvg = 0.5 * ((vr - vg) + fabs(vr - vg));
        const double dvr = 0.5 * ((vg - vr) + fabs(vg - vr));
        const double cc = pef * (Dg[n] + Dr[n2]) * (Dg[n] + Dr[n2]);
        const double NN_g = s_Ng[n], NN_r = s_Nr[n2];
        acc[0] += cc * dvg * massg * NN_g * NN_r;
        acc[1] += cc * dvg * massr * NN_g * NN_r;
        acc[2] += cc * dvg * NN_g * NN_r;
        acc[3] += cc * dvr * massr * NN_g * NN_r;
        acc[4] += cc * dvr * NN_g * NN_r;
        acc[5] += cc * dvr * massg * NN_g * NN_r;
    }
    block_sum<6>(acc, s_red);
    if (threadIdx.x == 0) {
        const size_t a = (size_t)i + (size_t)ntb_g1 * ((size_t)j + (size_t)ntb_g * ((size_t)k + (size_t)ntb_r1 * (size_t)m));
        tcg_racg[a] = acc[0];
        tmr_racg[a] = fmin(acc[1], decade_value(m, -6) * 1.0);
        tcr_gacr[a] = acc[3];
        tmg_gacr[a] = acc[5];
        tnr_racg[a] = acc[2];
        tnr_gacr[a] = acc[4];
    }
}

// Field et al. (2005) moment regression in REAL arithmetic, the Fortran's term order (:3113-3123)
__device__ __forceinline__ float sa_poly(const float* s, float tc0, float x) {
    return s[0] + s[1] * tc0 + s[2] * x + s[3] * tc0 * x + s[4] * tc0 * tc0 + s[5] * x * x + s[6] * tc0 * tc0 * x +
           s[7] * tc0 * x * x + s[8] * tc0 * tc0 * tc0 + s[9] * x * x * x;
}
__device__ const float k_sa[10] = {5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f,
                                   0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
__device__ const float k_sb[10] = {0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f,
                                   0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};
__device__ const float k_Tc[ntb_t] = {-0.01f, -5.f, -10.f, -15.f, -20.f, -25.f, -30.f, -35.f, -40.f};

// qr_acr_qs :3093-3220.  blockIdx.x = (j-1)*ntb_s + (i-1), blockIdx.y = (m-1)*ntb_r1 + (k-1)
struct RacsOut { double* t[12]; };
__global__ void __launch_bounds__(128)
k_qr_acr_qs(ThConst c, const double* __restrict__ bins, const double* __restrict__ aux, RacsOut out) {
    __shared__ double s_Nr[nbins], s_Ns[nbins], s_par[6], s_red[12 * 8];
    const int i = blockIdx.x % ntb_s, j = blockIdx.x / ntb_s, k = blockIdx.y % ntb_r1, m = blockIdx.y / ntb_r1;
    const double *Dr = bin_D(bins, 2), *dtr = bin_dt(bins, 2), *Ds = bin_D(bins, 3), *dts = bin_dt(bins, 3);
    if (threadIdx.x == 0) {
        rain_dsd(c, k, m, s_par[0], s_par[1]);
        const float tcj = k_Tc[j];
        double M2 = decade_value(i, -5) * c.oams * 1.0, second;
        if (bm_s > 2.0f - 1.E-3f && bm_s < 2.0f + 1.E-3f) {          // :3112 (sic: the branch taken for bm_s = 2)
            double loga_ = sa_poly(k_sa, tcj, bm_s);
            double a_ = pow(10.0, loga_);
            double b_ = sa_poly(k_sb, tcj, bm_s);
            second = pow((M2 / a_), (1. / b_));
        } else {
            second = M2;
        }
        double loga_ = sa_poly(k_sa, tcj, c.cse[0]);
        double a_ = pow(10.0, loga_);
        double b_ = sa_poly(k_sb, tcj, c.cse[0]);
        double M3 = a_ * pow(second, b_);
        double oM3 = 1. / M3;
        s_par[2] = M2 * (M2 * oM3) * (M2 * oM3) * (M2 * oM3);       // Mrat
        s_par[3] = pow((M2 * oM3), (double)mu_s);                   // M0
        s_par[4] = M2 * oM3 * Lam0;                                 // slam1
        s_par[5] = M2 * oM3 * Lam1;                                 // slam2
    }
    __syncthreads();
    if (threadIdx.x < nbins) {
        const int n = threadIdx.x;
        s_Nr[n] = s_par[1] * pow(Dr[n], (double)c.o.mu_r) * exp(-s_par[0] * Dr[n]) * dtr[n];
        s_Ns[n] = s_par[2] * (Kap0 * exp(-s_par[4] * Ds[n]) + Kap1 * s_par[3] * pow(Ds[n], (double)mu_s) * exp(-s_par[5] * Ds[n])) * dts[n];
    }
    __syncthreads();
    const float pef = PI2 * .25f * c.o.Ef_rs;
    double acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // t1 z1 y1 t2 y2 z2 | t3 z3 y3 t4 y4 z4
    for (int p = threadIdx.x; p < nbins * nbins; p += blockDim.x) {
        const int n2 = p / nbins, n = p % nbins;
        const double vr = aux[n2], vs = aux[200 + n], massr = aux[300 + n2], masss = aux[500 + n];
        const double dvs = 0.5 * ((vr - vs) + fabs(vr - vs));
        const double dvr = 0.5 * ((vs - vr) + fabs(vs - vr));
        const double cc = pef * (Ds[n] + Dr[n2]) * (Ds[n] + Dr[n2]);
        const double NN_s = s_Ns[n], NN_r = s_Nr[n2];
        const double a0 = cc * dvs * masss * NN_s * NN_r, a1 = cc * dvs * massr * NN_s * NN_r, a2 = cc * dvs * NN_s * NN_r;
        const double a3 = cc * dvr * massr * NN_s * NN_r, a4 = cc * dvr * NN_s * NN_r, a5 = cc * dvr * masss * NN_s * NN_r;
        if (massr > 1.5f * masss) { acc[0] += a0; acc[1] += a1; acc[2] += a2; acc[3] += a3; acc[4] += a4; acc[5] += a5; }
        else { acc[6] += a0; acc[7] += a1; acc[8] += a2; acc[9] += a3; acc[10] += a4; acc[11] += a5; }
    }
    block_sum<12>(acc, s_red);
    if (threadIdx.x == 0) {
        const size_t a = (size_t)i + (size_t)ntb_s * ((size_t)j + (size_t)ntb_t * ((size_t)k + (size_t)ntb_r1 * (size_t)m));
        out.t[0][a] = acc[0];                                       // tcs_racs1 = t1
        out.t[1][a] = fmin(acc[1], decade_value(m, -6) * 1.0);      // tmr_racs1 = min(z1, r_r)
        out.t[2][a] = acc[6];                                       // tcs_racs2 = t3
        out.t[3][a] = acc[7];                                       // tmr_racs2 = z3
        out.t[4][a] = acc[3];                                       // tcr_sacr1 = t2
        out.t[5][a] = acc[5];                                       // tms_sacr1 = z2
        out.t[6][a] = acc[9];                                       // tcr_sacr2 = t4
        out.t[7][a] = acc[11];                                      // tms_sacr2 = z4
        out.t[8][a] = acc[2];                                       // tnr_racs1 = y1
        out.t[9][a] = acc[8];                                       // tnr_racs2 = y3
        out.t[10][a] = acc[4];                                      // tnr_sacr1 = y2
        out.t[11][a] = acc[10];                                     // tnr_sacr2 = y4
    }
}

// freezeH2O :3323-3381: the bin loop runs from the largest drops down and EXITs once the frozen mass
// reaches the available mass, so it stays a sequential loop -- one thread per table entry.
__global__ void k_freezeH2O(ThConst c, const double* __restrict__ bins, const double* __restrict__ aux,
                            double* tpi_qrfz, double* tni_qrfz, double* tpg_qrfz, double* tnr_qrfz,
                            double* tpi_qcfz, double* tni_qcfz) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrain = ntb_r * ntb_r1 * 45;
    const double orho_w = 1.f / rho_w;
    if (t < nrain) {
        const int i = t % ntb_r, j = (t / ntb_r) % ntb_r1, k = t / (ntb_r * ntb_r1) + 1;
        const double Texp = exp((double)k - c.o.t_adjust * 1.0) - 1.0;
        const double *Dr = bin_D(bins, 2), *dtr = bin_dt(bins, 2);
        double lamr, N0_r;
        rain_dsd(c, j, i, lamr, N0_r);
        const float rr = decade_value(i, -6);
        double sum1 = 0, sum2 = 0, sumn1 = 0, sumn2 = 0;
        for (int n2 = nbins - 1; n2 >= 0; --n2) {
            const double N_r = N0_r * pow(Dr[n2], (double)c.o.mu_r) * exp(-lamr * Dr[n2]) * dtr[n2];
            const double massr = aux[300 + n2];
            const double vol = massr * orho_w;
            double prob = 1.0 - exp(-120.0 * vol * 5.2E-4 * Texp);
            prob = fmax(prob, 0.0);
            if (massr < c.xm0g) { sumn1 = sumn1 + prob * N_r; sum1 = sum1 + prob * N_r * massr; }
            else { sumn2 = sumn2 + prob * N_r; sum2 = sum2 + prob * N_r * massr; }
            if ((sum1 + sum2) >= rr) break;
        }
        tpi_qrfz[t] = sum1; tni_qrfz[t] = sumn1; tpg_qrfz[t] = sum2; tnr_qrfz[t] = sumn2;
    } else if (t < nrain + ntb_c * 45) {
        const int u = t - nrain, i = u % ntb_c, k = u / ntb_c + 1;
        const double Texp = exp((double)k - c.o.t_adjust * 1.0) - 1.0;
        const double *Dc = bin_D(bins, 0), *dtc = bin_dt(bins, 0);
        const float rc = decade_value(i, -6);
        const double lamc = 1.0E-6 * powf((c.o.Nt_c * c.am_r * c.ccg[1] * c.ocg1 / rc), c.obmr);
        const double N0_c = 1.0E-18 * c.o.Nt_c * c.ocg1 * pow(lamc, (double)c.cce[0]);
        double sum1 = 0, sumn2 = 0;
        for (int n = nbins - 1; n >= 0; --n) {
            const double y = Dc[n] * 1.0E6;
            const double massc = aux[600 + n];
            const double vol = massc * orho_w;
            double prob = 1.0 - exp(-120.0 * vol * 5.2E-4 * Texp);
            prob = fmax(prob, 0.0);
            double N_c = N0_c * pow(y, (double)c.mu_c) * exp(-lamc * y) * dtc[n];
            N_c = 1.0E24 * N_c;
            sumn2 = sumn2 + prob * N_c;
            sum1 = sum1 + prob * N_c * massc;
            if (sum1 >= rc) break;
        }
        tpi_qcfz[u] = sum1; tni_qcfz[u] = sumn2;
    }
}

// GAMMLN / GSER / GCF / GAMMP in the Fortran's REAL arithmetic, :3650-3761
__device__ float gammln_d(float xx) {
    const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                           -1.231739572450155, .1208650973866179E-2, -.5395239384953E-5};
    double x = xx, y = x, tmp = x + 5.5;
    tmp = (x + 0.5) * log(tmp) - tmp;
    double ser = 1.000000000190015;
    for (int j = 0; j < 6; ++j) { y = y + 1.0; ser = ser + cof[j] / y; }
    return (float)(tmp + log(2.5066282746310005 * ser / x));
}
__device__ float gammp_d(float a, float x) {
    const float gEPS = 3.E-7f, FPMIN = 1.E-30f;
    if ((x < 0.f) || (a <= 0.f)) return 0.f;
    const float gln = gammln_d(a);
    if (x < a + 1.f) {                                    // GSER
        if (x <= 0.f) return 0.f;
        float ap = a, sum = 1.f / a, del = sum;
        for (int n = 1; n <= 100; ++n) {
            ap = ap + 1.f;
            del = del * x / ap;
            sum = sum + del;
            if (fabsf(del) < fabsf(sum) * gEPS) break;
        }
        return sum * expf(-x + a * logf(x) - gln);
    }
    float b = x + 1.f - a, cc = 1.f / FPMIN, d = 1.f / b, h = d;   // GCF
    for (int i = 1; i <= 100; ++i) {
        const float an = -(float)i * ((float)i - a);
        b = b + 2.f;
        d = an * d + b;
        if (fabsf(d) < FPMIN) d = FPMIN;
        cc = b + an / cc;
        if (fabsf(cc) < FPMIN) cc = FPMIN;
        d = 1.f / d;
        const float del = d * cc;
        h = h * del;
        if (fabsf(del - 1.f) < gEPS) break;
    }
    return 1.f - expf(-x + a * logf(x) - gln) * h;
}

// qi_aut_qs :3425-3454, one thread per entry
__global__ void k_qi_aut_qs(ThConst c, const double* __restrict__ bins, double* tps_iaus, double* tni_iaus,
                            double* tpi_ide) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntb_i * ntb_i1) return;
    const int i = t % ntb_i, j = t / ntb_i;
    const float ri = decade_value(i, -10), Nti = decade_value(j, 0);
    const double *Di = bin_D(bins, 1), *dti = bin_dt(bins, 1);
    const double lami = powf((c.am_i * c.cig[1] * c.oig1 * Nti / ri), c.obmi);
    const double Di_mean = (bm_i + mu_i + 1.f) / lami;
    const double N0_i = Nti * c.oig1 * pow(lami, (double)c.cie[0]);
    double t1 = 0, t2 = 0, ide;
    if ((float)Di_mean > 5.f * D0s) { t1 = ri; t2 = Nti; ide = 0.0; }
    else if ((float)Di_mean < c.D0i) { t1 = 0; t2 = 0; ide = 1.0; }
    else {
        const float xlimit_intg = (float)(lami * D0s);
        ide = gammp_d(mu_i + 2.0f, xlimit_intg) * 1.0;
        for (int n2 = 0; n2 < nbins; ++n2) {
            const double N_i = N0_i * pow(Di[n2], (double)mu_i) * exp(-lami * Di[n2]) * dti[n2];
            if (Di[n2] >= D0s) { t1 = t1 + N_i * c.am_i * pow(Di[n2], (double)bm_i); t2 = t2 + N_i; }
        }
    }
    tps_iaus[t] = t1; tni_iaus[t] = t2; tpi_ide[t] = ide;
}

// table_Efrw :3473-3523 and table_Efsw :3542-3576, one thread per entry; which = 0 rain, 1 snow
__global__ void k_table_Ef(ThConst c, const double* __restrict__ bins, double* t_Efrw, double* t_Efsw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * nbins * nbins) return;
    const int which = t / (nbins * nbins), u = t % (nbins * nbins), i = u % nbins, j = u / nbins;
    const double Dcj = bin_D(bins, 0)[j];
    if (which == 0) {
        const double Dri = bin_D(bins, 2)[i];
        double Ef_rw = 0.0;
        const double p = Dcj / Dri;
        if (Dri < 50.E-6f || Dcj < 3.E-6f) {
            // t_Efrw = 0
        } else if (p > 0.25f) {
            const double X = Dcj * 1.E6;
            if (Dri < 75.e-6f) Ef_rw = 0.026794f * X - 0.20604f;
            else if (Dri < 125.e-6f) Ef_rw = -0.00066842f * X * X + 0.061542f * X - 0.37089f;
            else if (Dri < 175.e-6f)
                Ef_rw = 4.091e-06f * X * X * X * X - 0.00030908f * X * X * X + 0.0066237f * X * X - 0.0013687f * X - 0.073022f;
            else if (Dri < 250.e-6f) Ef_rw = 9.6719e-5f * X * X * X - 0.0068901f * X * X + 0.17305f * X - 0.65988f;
            else if (Dri < 350.e-6f) Ef_rw = 9.0488e-5f * X * X * X - 0.006585f * X * X + 0.16606f * X - 0.56125f;
            else Ef_rw = 0.00010721f * X * X * X - 0.0072962f * X * X + 0.1704f * X - 0.46929f;
        } else {
            const double vtr = vr_poly(Dri);
            const double stokes = Dcj * Dcj * vtr * rho_w / (9.f * 1.718E-5f * Dri);
            const double reynolds = 9.f * stokes / (p * p * rho_w);
            const double F = log(reynolds);
            const double G = -0.1007 - 0.358 * F + 0.0261 * F * F;
            const double K0 = exp(G);
            const double z = log(stokes / (K0 + 1.E-15));
            const double H = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
            const double yc0 = 2.0 / PI2 * atan(H);
            Ef_rw = (yc0 + p) * (yc0 + p) / ((1.f + p) * (1.f + p));
        }
        double v = fmaxf(0.0f, fminf((float)Ef_rw, 0.95f));
        if (c.o.Ef_rw_l && Ef_rw != 0.0) v = 1.0;
        t_Efrw[u] = v;
    } else {
        const double Dsi = bin_D(bins, 3)[i];
        const double vtc = 1.19E4 * (1.0E4 * Dcj * Dcj * 0.25);
        const double vts = c.o.av_s * pow(Dsi, (double)c.o.bv_s) * exp(-c.o.fv_s * Dsi) - vtc;
        const double Ds_m = pow((c.o.am_s * pow(Dsi, (double)bm_s) / c.am_r), (double)c.obmr);
        const double p = Dcj / Ds_m;
        double v = 0.0;
        if (!(p > 0.25f || Dsi < D0s || Dcj < 6.E-6f || vts < 1.E-3f)) {
            const double stokes = Dcj * Dcj * vts * rho_w / (9.f * 1.718E-5f * Ds_m);
            const double reynolds = 9.f * stokes / (p * p * rho_w);
            const double F = log(reynolds);
            const double G = -0.1007 - 0.358 * F + 0.0261 * F * F;
            const double K0 = exp(G);
            const double z = log(stokes / (K0 + 1.E-15));
            const double H = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
            const double yc0 = 2.0 / PI2 * atan(H);
            const double Ef_sw = (yc0 + p) * (yc0 + p) / ((1.f + p) * (1.f + p));
            v = fmaxf(0.0f, fminf((float)Ef_sw, 0.95f));
            if (c.o.Ef_sw_l && Ef_sw != 0.0) v = 1.0;
        }
        t_Efsw[u] = v;
    }
}

}  // namespace

int th_build_tables(const ThConst& c, const double* d_bins, double* d_aux, double* const* tb, cudaStream_t st) {
    int n = 0;
    k_table_aux<<<1, 128, 0, st>>>(c, d_bins, d_aux); ++n;
    k_table_Ef<<<(2 * nbins * nbins + 127) / 128, 128, 0, st>>>(c, d_bins, tb[T_Efrw], tb[T_Efsw]); ++n;
    k_qr_acr_qg<<<dim3(ntb_g1 * ntb_g, ntb_r1 * ntb_r), 128, 0, st>>>(
        c, d_bins, d_aux, tb[T_tcg_racg], tb[T_tmr_racg], tb[T_tcr_gacr], tb[T_tmg_gacr], tb[T_tnr_racg], tb[T_tnr_gacr]); ++n;
    RacsOut ro;
    for (int a = 0; a < 12; ++a) ro.t[a] = tb[T_tcs_racs1 + a];
    k_qr_acr_qs<<<dim3(ntb_s * ntb_t, ntb_r1 * ntb_r), 128, 0, st>>>(c, d_bins, d_aux, ro); ++n;
    const int nf = ntb_r * ntb_r1 * 45 + ntb_c * 45;
    k_freezeH2O<<<(nf + 127) / 128, 128, 0, st>>>(c, d_bins, d_aux, tb[T_tpi_qrfz], tb[T_tni_qrfz], tb[T_tpg_qrfz],
                                                  tb[T_tnr_qrfz], tb[T_tpi_qcfz], tb[T_tni_qcfz]); ++n;
    k_qi_aut_qs<<<(ntb_i * ntb_i1 + 127) / 128, 128, 0, st>>>(c, d_bins, tb[T_tps_iaus], tb[T_tni_iaus], tb[T_tpi_ide]); ++n;
    return n;
}

}  // namespace icar
