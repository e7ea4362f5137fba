// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of ICAR's Thompson et al. (2008) bulk microphysics (mp = 1), statement by
// statement, keeping the Fortran's REAL (float) / DOUBLE PRECISION (double) typing of every
// variable so that mixed-precision expressions promote exactly where the Fortran's do.
// PARITY PINNED TO THE REFERENCE'S OWN SOURCE: the reference holds no expected values for this path and
// cannot be compiled here (no Fortran compiler), so oracle/f90run executes src/physics/mp_thompson.f90 itself.
// Equal in every bit to its outputs: thompson_init's 130 scalar constants and bin grids, the tables of freezeH2O /
// qi_aut_qs / table_Efrw / table_Efsw, three (k, m) slabs of each rain-collection table from qr_acr_qg / qr_acr_qs,
// and mp_gt_driver + mp_thompson over six calls with every species present (tests/test_reference_source.py,
// tests/golden/reference_source_thompson.npz).  Not executed: the other 1366 slabs of each of the two collection tables
// (1.4e10 inner iterations); those have an independent NumPy / SciPy check.  DESIGN.md section 5.
//
// Reference files followed (relative to /root/reference):
//   src/physics/mp_thompson.f90:50-331     constants, table axes
//   src/physics/mp_thompson.f90:342-766    thompson_init
//   src/physics/mp_thompson.f90:772-1044   mp_gt_driver
//   src/physics/mp_thompson.f90:1057-2844  mp_thompson (one column)
//   src/physics/mp_thompson.f90:2853-3007  qr_acr_qg
//   src/physics/mp_thompson.f90:3014-3264  qr_acr_qs
//   src/physics/mp_thompson.f90:3273-3399  freezeH2O
//   src/physics/mp_thompson.f90:3413-3456  qi_aut_qs
//   src/physics/mp_thompson.f90:3464-3578  table_Efrw, table_Efsw
//   src/physics/mp_thompson.f90:3650-3835  GCF, GSER, GAMMLN, GAMMP, WGAMMA, RSLF, RSIF
//   src/objects/opt_types.f90:31-45        mp_options_type
// Intrinsics: `x**y` REAL -> powf, DOUBLE -> pow; `10.**n` (integer n) -> libgcc __powisf2, which is
// what gfortran lowers real**integer to; NINT -> lround; INT -> truncation; ALOG10 -> log10f.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

namespace icar_oracle {

struct ThompsonOptions {          // mp_options_type, opt_types.f90:31-45 (defaults options_obj.f90:1259-1280)
    float Nt_c = 100.e6f, TNO = 5.0f, am_s = 0.069f, rho_g = 500.0f;
    float av_s = 40.0f, bv_s = 0.55f, fv_s = 100.0f, av_i = 1847.5f;
    float av_g = 442.0f, bv_g = 0.89f;
    float Ef_si = 0.05f, Ef_rs = 0.95f, Ef_rg = 0.75f, Ef_ri = 0.95f;
    float C_cubes = 0.5f, C_sqrd = 0.3f, mu_r = 0.f, t_adjust = 0.f;
    int Ef_rw_l = 0, Ef_sw_l = 0;
};

namespace th {
// :50-193
constexpr float T_0 = 273.15f, PI2 = 3.1415926536f;
constexpr float rho_w = 1000.0f, rho_s = 100.0f, rho_i = 890.0f;
constexpr float mu_g = 0.0f, mu_i = 0.0f;
constexpr float mu_s = 0.6357f, Kap0 = 490.6f, Kap1 = 17.46f, Lam0 = 20.78f, Lam1 = 3.29f;
constexpr float gonv_min = 1.E4f, gonv_max = 3.E6f;
constexpr float bm_r = 3.0f, bm_s = 2.0f, bm_g = 3.0f, bm_i = 3.0f;
constexpr float av_r = 4854.0f, bv_r = 1.0f, fv_r = 195.0f, bv_i = 1.0f;
constexpr float C_cube = 0.5f;
constexpr float R1 = 1.E-12f, R2 = 1.E-6f, eps = 1.E-15f;
constexpr float ATO = 0.304f;
constexpr float Sc = 0.632f, HGFR = 235.16f, Rv = 461.5f, RR2 = 287.04f, Cp2 = 1004.0f;
constexpr float lsub = 2.834E6f, lvap0 = 2.5E6f;
constexpr float xm0i = 1.E-12f, D0c = 1.E-6f, D0r = 50.E-6f, D0s = 200.E-6f, D0g = 250.E-6f;
constexpr int nbins = 100, nbc = nbins, nbi = nbins, nbr = nbins, nbs = nbins, nbg = nbins;
constexpr int ntb_c = 37, ntb_i = 64, ntb_r = 37, ntb_s = 28, ntb_g = 28, ntb_g1 = 28, ntb_r1 = 37,
              ntb_i1 = 55, ntb_t = 9;

inline float powi_f(float x, int m) {       // libgcc __powisf2
    unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
    float y = (n % 2) ? x : 1.f;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return m < 0 ? 1.f / y : y;
}
inline int nint_f(float x) { return (int)std::lround(x); }
inline int nint_d(double x) { return (int)std::lround(x); }

// decade tables :204-268: entry (1-based) idx = d + 9*(n - n_first) holds d * 10**n, d = 1..9
inline float decade(int idx1, int n_first) {
    int n = (idx1 - 1) / 9, d = (idx1 - 1) % 9 + 1;
    static const char* fmt = "%de%d";
    char b[32]; std::snprintf(b, sizeof b, fmt, d, n + n_first);
    float v; std::sscanf(b, "%f", &v);       // the literal "2.e-6" as a REAL constant
    return v;
}
}  // namespace th

struct Thompson {
    // ---- module state (:196-201, :288-330)
    double Dc[th::nbc], dtc[th::nbc], Di[th::nbi], dti[th::nbi], Dr[th::nbr], dtr[th::nbr], Ds[th::nbs],
        dts[th::nbs], Dg[th::nbg], dtg[th::nbg];
    float r_c[th::ntb_c], r_i[th::ntb_i], r_r[th::ntb_r], r_g[th::ntb_g], r_s[th::ntb_s], N0r_exp[th::ntb_r1],
        N0g_exp[th::ntb_g1], Nt_i[th::ntb_i1];
    const float sa[10] = {5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f,
                          0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
    const float sb[10] = {0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f,
                          0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};
    const float Tc[th::ntb_t] = {-0.01f, -5.f, -10.f, -15.f, -20.f, -25.f, -30.f, -35.f, -40.f};
    // tables, Fortran column-major, fp64
    std::vector<double> tcg_racg, tmr_racg, tcr_gacr, tmg_gacr, tnr_racg, tnr_gacr;           // (g1,g,r1,r)
    std::vector<double> tcs_racs1, tmr_racs1, tcs_racs2, tmr_racs2, tcr_sacr1, tms_sacr1, tcr_sacr2,
        tms_sacr2, tnr_racs1, tnr_racs2, tnr_sacr1, tnr_sacr2;                                // (s,t,r1,r)
    std::vector<double> tpi_qcfz, tni_qcfz;                                                   // (c,45)
    std::vector<double> tpi_qrfz, tpg_qrfz, tni_qrfz, tnr_qrfz;                               // (r,r1,45)
    std::vector<double> tps_iaus, tni_iaus, tpi_ide;                                          // (i,i1)
    std::vector<double> t_Efrw, t_Efsw;                                                       // (nbr|nbs, nbc)
    float mu_c, Sc3, D0i, xm0s, xm0g;
    float cce[3], ccg[3], ocg1, ocg2;
    float cie[7], cig[7], oig1, oig2, obmi;
    float cre[13], crg[13], ore1, org1, org2, org3, obmr;
    float cse[18], csg[18], oams, obms, ocms;
    float cge[12], cgg[12], oge1, ogg1, ogg2, ogg3, oamg, obmg, ocmg;
    float t1_qr_qc, t1_qr_qi, t2_qr_qi, t1_qg_qc, t1_qs_qc, t1_qs_qi, t1_qr_ev, t2_qr_ev, t1_qs_sd, t2_qs_sd,
        t1_qg_sd, t2_qg_sd, t1_qs_me, t2_qs_me, t1_qg_me, t2_qg_me;
    int nic2, nii2, nii3, nir2, nir3, nis2, nig2, nig3;
    float Nt_c, TNO, am_s, rho_g, av_s, bv_s, fv_s, av_g, bv_g, av_i, Ef_si, Ef_rs, Ef_rg, Ef_ri, C_cubes, C_sqrd,
        mu_r, t_adjust, am_g;
    bool Ef_rw_l, Ef_sw_l;
    float am_r, am_i, lfus, olfus, oRv, rho_not;

    // ---- :3718-3771
    static float GAMMLN(float XX) {
        const double STP = 2.5066282746310005;
        const double COF[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                               -1.231739572450155, .1208650973866179E-2, -.5395239384953E-5};
        double X = XX, Y = X, TMP = X + 5.5;
        TMP = (X + 0.5) * std::log(TMP) - TMP;
        double SER = 1.000000000190015;
        for (int J = 0; J < 6; ++J) { Y = Y + 1.0; SER = SER + COF[J] / Y; }
        return (float)(TMP + std::log(STP * SER / X));
    }
    static float WGAMMA(float y) { return std::exp(GAMMLN(y)); }
    static float GSER(float A, float X) {                                   // :3686-3715
        const float gEPS = 3.E-7f;
        float GLN = GAMMLN(A);
        if (X <= 0.f) return 0.f;
        float AP = A, SUM = 1.f / A, DEL = SUM;
        for (int N = 1; N <= 100; ++N) {
            AP = AP + 1.f;
            DEL = DEL * X / AP;
            SUM = SUM + DEL;
            if (std::fabs(DEL) < std::fabs(SUM) * gEPS) break;
        }
        return SUM * std::exp(-X + A * std::log(X) - GLN);
    }
    static float GCF(float A, float X) {                                    // :3650-3683
        const float gEPS = 3.E-7f, FPMIN = 1.E-30f;
        float GLN = GAMMLN(A);
        float B = X + 1.f - A, C = 1.f / FPMIN, D = 1.f / B, H = D;
        for (int I = 1; I <= 100; ++I) {
            float AN = -(float)I * ((float)I - A);
            B = B + 2.f;
            D = AN * D + B;
            if (std::fabs(D) < FPMIN) D = FPMIN;
            C = B + AN / C;
            if (std::fabs(C) < FPMIN) C = FPMIN;
            D = 1.f / D;
            float DEL = D * C;
            H = H * DEL;
            if (std::fabs(DEL - 1.f) < gEPS) break;
        }
        return std::exp(-X + A * std::log(X) - GLN) * H;
    }
    static float GAMMP(float A, float X) {                                  // :3743-3761
        if ((X < 0.f) || (A <= 0.f)) return 0.f;
        if (X < A + 1.f) return GSER(A, X);
        return 1.f - GCF(A, X);
    }
    static float RSLF(float P, float T) {                                   // :3776-3805
        const float C0 = .611583699E03f, C1 = .444606896E02f, C2 = .143177157E01f, C3 = .264224321E-1f,
                    C4 = .299291081E-3f, C5 = .203154182E-5f, C6 = .702620698E-8f, C7 = .379534310E-11f,
                    C8 = -.321582393E-13f;
        float X = std::max(-80.f, T - 273.16f);
        float ESL = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
        return .622f * ESL / (P - ESL);
    }
    static float RSIF(float P, float T) {                                   // :3810-3835
        const float C0 = .609868993E03f, C1 = .499320233E02f, C2 = .184672631E01f, C3 = .402737184E-1f,
                    C4 = .565392987E-3f, C5 = .521693933E-5f, C6 = .307839583E-7f, C7 = .105785160E-9f,
                    C8 = .161444444E-12f;
        float X = std::max(-80.f, T - 273.16f);
        float ESI = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
        return .622f * ESI / (P - ESI);
    }

    static size_t i4(int i, int j, int k, int m, int n1, int n2, int n3) {   // 0-based, column-major
        return (size_t)i + (size_t)n1 * ((size_t)j + (size_t)n2 * ((size_t)k + (size_t)n3 * (size_t)m));
    }
    static size_t i3(int i, int j, int k, int n1, int n2) { return (size_t)i + (size_t)n1 * ((size_t)j + (size_t)n2 * (size_t)k); }
    static size_t i2(int i, int j, int n1) { return (size_t)i + (size_t)n1 * (size_t)j; }

    // snow-moment regression (Field et al. 2005) for the `cse`-th moment, REAL arithmetic, the
    // Fortran's term order (e.g. :1413-1422)
    float sa_poly(const float* s, float tc0, float c) const {
        return s[0] + s[1] * tc0 + s[2] * c + s[3] * tc0 * c + s[4] * tc0 * tc0 + s[5] * c * c +
               s[6] * tc0 * tc0 * c + s[7] * tc0 * c * c + s[8] * tc0 * tc0 * tc0 + s[9] * c * c * c;
    }

    // ------------------------------------------------------------------ thompson_init :342-766
    // tables_mask: bit 0 = the two big collision tables (qr_acr_qg, qr_acr_qs); all others always.
    void init(const ThompsonOptions& o, bool big_tables = true) {
        using namespace th;
        Nt_c = o.Nt_c; TNO = o.TNO; am_s = o.am_s; rho_g = o.rho_g; av_s = o.av_s; bv_s = o.bv_s; fv_s = o.fv_s;
        av_g = o.av_g; bv_g = o.bv_g; av_i = o.av_i; Ef_rs = o.Ef_rs; Ef_rg = o.Ef_rg; Ef_si = o.Ef_si;
        Ef_ri = o.Ef_ri; C_cubes = o.C_cubes; C_sqrd = o.C_sqrd; mu_r = o.mu_r; t_adjust = o.t_adjust;
        Ef_rw_l = o.Ef_rw_l != 0; Ef_sw_l = o.Ef_sw_l != 0;
        am_g = PI2 * rho_g / 6.0f;                         // :417
        am_r = PI2 * rho_w / 6.0f;                         // :97
        am_i = PI2 * rho_i / 6.0f;                         // :103
        lfus = lsub - lvap0; olfus = 1.f / lfus; oRv = 1.f / Rv;          // :157,:166-167
        rho_not = 101325.0f / (287.05f * 298.0f);          // :146
        for (int n = 1; n <= ntb_c; ++n) r_c[n - 1] = decade(n, -6);
        for (int n = 1; n <= ntb_i; ++n) r_i[n - 1] = decade(n, -10);
        for (int n = 1; n <= ntb_r; ++n) r_r[n - 1] = decade(n, -6);
        for (int n = 1; n <= ntb_g; ++n) r_g[n - 1] = decade(n, -5);
        for (int n = 1; n <= ntb_s; ++n) r_s[n - 1] = decade(n, -5);
        for (int n = 1; n <= ntb_r1; ++n) N0r_exp[n - 1] = decade(n, 6);
        for (int n = 1; n <= ntb_g1; ++n) N0g_exp[n - 1] = decade(n, 4);
        for (int n = 1; n <= ntb_i1; ++n) Nt_i[n - 1] = decade(n, 0);

        mu_c = std::min(15.f, (1000.E6f / Nt_c + 2.f));    // :424
        Sc3 = std::pow(Sc, (1.f / 3.f));                   // :427
        D0i = std::pow((xm0i / am_i), (1.f / bm_i));       // :430
        xm0s = am_s * std::pow(D0s, bm_s);
        xm0g = am_g * std::pow(D0g, bm_g);

        cce[0] = mu_c + 1.f; cce[1] = bm_r + mu_c + 1.f; cce[2] = bm_r + mu_c + 4.f;       // :436-443
        for (int n = 0; n < 3; ++n) ccg[n] = WGAMMA(cce[n]);
        ocg1 = 1.f / ccg[0]; ocg2 = 1.f / ccg[1];

        cie[0] = mu_i + 1.f; cie[1] = bm_i + mu_i + 1.f; cie[2] = bm_i + mu_i + bv_i + 1.f;  // :445-461
        cie[3] = mu_i + bv_i + 1.f; cie[4] = mu_i + 2.f; cie[5] = bm_i * 0.5f + mu_i + bv_i + 1.f;
        cie[6] = bm_i * 0.5f + mu_i + 1.f;
        for (int n = 0; n < 7; ++n) cig[n] = WGAMMA(cie[n]);
        oig1 = 1.f / cig[0]; oig2 = 1.f / cig[1]; obmi = 1.f / bm_i;

        cre[0] = bm_r + 1.f; cre[1] = mu_r + 1.f; cre[2] = bm_r + mu_r + 1.f; cre[3] = bm_r * 2.f + mu_r + 1.f;  // :463-483
        cre[4] = mu_r + bv_r + 1.f; cre[5] = bm_r + mu_r + bv_r + 1.f; cre[6] = bm_r * 0.5f + mu_r + bv_r + 1.f;
        cre[7] = bm_r + mu_r + bv_r + 3.f; cre[8] = mu_r + bv_r + 3.f; cre[9] = mu_r + 2.f;
        cre[10] = 0.5f * (bv_r + 5.f + 2.f * mu_r); cre[11] = bm_r * 0.5f + mu_r + 1.f;
        cre[12] = bm_r * 2.f + mu_r + bv_r + 1.f;
        for (int n = 0; n < 13; ++n) crg[n] = WGAMMA(cre[n]);
        obmr = 1.f / bm_r; ore1 = 1.f / cre[0]; org1 = 1.f / crg[0]; org2 = 1.f / crg[1]; org3 = 1.f / crg[2];

        cse[0] = bm_s + 1.f; cse[1] = bm_s + 2.f; cse[2] = bm_s * 2.f; cse[3] = bm_s + bv_s + 1.f;   // :485-508
        cse[4] = bm_s * 2.f + bv_s + 1.f; cse[5] = bm_s * 2.f + 1.f; cse[6] = bm_s + mu_s + 1.f;
        cse[7] = bm_s + mu_s + 2.f; cse[8] = bm_s + mu_s + 3.f; cse[9] = bm_s + mu_s + bv_s + 1.f;
        cse[10] = bm_s * 2.f + mu_s + bv_s + 1.f; cse[11] = bm_s * 2.f + mu_s + 1.f; cse[12] = bv_s + 2.f;
        cse[13] = bm_s + bv_s; cse[14] = mu_s + 1.f; cse[15] = 1.0f + (1.0f + bv_s) / 2.f;
        cse[16] = cse[15] + mu_s + 1.f; cse[17] = bv_s + mu_s + 3.f;
        for (int n = 0; n < 18; ++n) csg[n] = WGAMMA(cse[n]);
        oams = 1.f / am_s; obms = 1.f / bm_s; ocms = std::pow(oams, obms);

        cge[0] = bm_g + 1.f; cge[1] = mu_g + 1.f; cge[2] = bm_g + mu_g + 1.f; cge[3] = bm_g * 2.f + mu_g + 1.f;  // :510-531
        cge[4] = bm_g * 2.f + mu_g + bv_g + 1.f; cge[5] = bm_g + mu_g + bv_g + 1.f;
        cge[6] = bm_g + mu_g + bv_g + 2.f; cge[7] = bm_g + mu_g + bv_g + 3.f; cge[8] = mu_g + bv_g + 3.f;
        cge[9] = mu_g + 2.f; cge[10] = 0.5f * (bv_g + 5.f + 2.f * mu_g); cge[11] = 0.5f * (bv_g + 5.f) + mu_g;
        for (int n = 0; n < 12; ++n) cgg[n] = WGAMMA(cge[n]);
        oamg = 1.f / am_g; obmg = 1.f / bm_g; ocmg = std::pow(oamg, obmg);
        oge1 = 1.f / cge[0]; ogg1 = 1.f / cgg[0]; ogg2 = 1.f / cgg[1]; ogg3 = 1.f / cgg[2];

        t1_qr_qc = PI2 * .25f * av_r * crg[8];             // :538-568
        t1_qr_qi = PI2 * .25f * av_r * crg[8];
        t2_qr_qi = PI2 * .25f * am_r * av_r * crg[7];
        t1_qg_qc = PI2 * .25f * av_g * cgg[8];
        t1_qs_qc = PI2 * .25f * av_s;
        t1_qs_qi = PI2 * .25f * av_s;
        t1_qr_ev = 0.78f * crg[9];
        t2_qr_ev = 0.308f * Sc3 * std::sqrt(av_r) * crg[10];
        t1_qs_sd = 0.86f;
        t2_qs_sd = 0.28f * Sc3 * std::sqrt(av_s);
        t1_qs_me = PI2 * 4.f * C_sqrd * olfus * 0.86f;
        t2_qs_me = PI2 * 4.f * C_sqrd * olfus * 0.28f * Sc3 * std::sqrt(av_s);
        t1_qg_sd = 0.86f * cgg[9];
        t2_qg_sd = 0.28f * Sc3 * std::sqrt(av_g) * cgg[10];
        t1_qg_me = PI2 * 4.f * C_cube * olfus * 0.86f * cgg[9];
        t2_qg_me = PI2 * 4.f * C_cube * olfus * 0.28f * Sc3 * std::sqrt(av_g) * cgg[10];

        nic2 = nint_f(std::log10(r_c[0])); nii2 = nint_f(std::log10(r_i[0])); nii3 = nint_f(std::log10(Nt_i[0]));  // :571-578
        nir2 = nint_f(std::log10(r_r[0])); nir3 = nint_f(std::log10(N0r_exp[0])); nis2 = nint_f(std::log10(r_s[0]));
        nig2 = nint_f(std::log10(r_g[0])); nig3 = nint_f(std::log10(N0g_exp[0]));

        Dc[0] = D0c * 1.0; dtc[0] = D0c * 1.0;             // :581-586
        for (int n = 1; n < nbc; ++n) { Dc[n] = Dc[n - 1] + 1.0E-6; dtc[n] = (Dc[n] - Dc[n - 1]); }
        auto bins = [](double lo, double hi, int nb, double* D, double* dt) {   // :589-634
            std::vector<double> xDx(nb + 1);
            xDx[0] = lo; xDx[nb] = hi;
            for (int n = 2; n <= nb; ++n)
                xDx[n - 1] = std::exp((double)(n - 1) / (double)nb * std::log(xDx[nb] / xDx[0]) + std::log(xDx[0]));
            for (int n = 0; n < nb; ++n) { D[n] = std::sqrt(xDx[n] * xDx[n + 1]); dt[n] = xDx[n + 1] - xDx[n]; }
        };
        bins(D0i * 1.0, 5.0 * D0s, nbi, Di, dti);
        bins(D0r * 1.0, 0.005, nbr, Dr, dtr);
        bins(D0s * 1.0, 0.02, nbs, Ds, dts);
        bins(D0g * 1.0, 0.05, nbg, Dg, dtg);

        table_Efrw();
        table_Efsw();
        const size_t ng = (size_t)ntb_g1 * ntb_g * ntb_r1 * ntb_r, ns = (size_t)ntb_s * ntb_t * ntb_r1 * ntb_r;
        for (auto* t : {&tcg_racg, &tmr_racg, &tcr_gacr, &tmg_gacr, &tnr_racg, &tnr_gacr}) t->assign(ng, 0.0);
        for (auto* t : {&tcs_racs1, &tmr_racs1, &tcs_racs2, &tmr_racs2, &tcr_sacr1, &tms_sacr1, &tcr_sacr2, &tms_sacr2,
                        &tnr_racs1, &tnr_racs2, &tnr_sacr1, &tnr_sacr2}) t->assign(ns, 0.0);
        if (big_tables) {
#pragma omp parallel for schedule(dynamic)
            for (int km = 0; km < ntb_r * ntb_r1; ++km) qr_acr_qg_km(km, 0, ntb_g1 * ntb_g);
#pragma omp parallel for schedule(dynamic)
            for (int km = 0; km < ntb_r * ntb_r1; ++km) qr_acr_qs_km(km, 0, ntb_s * ntb_t);
        }
        freezeH2O();
        qi_aut_qs();
    }

    // rain fall speed polynomial of the table builders (:2893-2895, :3066-3068, :3499-3501)
    static double vr_poly(double D) {
        return -0.1021f + 4.932E3f * D - 0.9551E6f * D * D + 0.07934E9f * D * D * D - 0.002362E12f * D * D * D * D;
    }

    // ------------------------------------------------------------------ qr_acr_qg :2853-3007
    // one (k, m) slab [km = (m-1)*ntb_r1 + (k-1)], cells ij in [ij0, ij1) with ij = (j-1)*ntb_g1 + (i-1)
    void qr_acr_qg_km(int km, int ij0, int ij1) {
        using namespace th;
        double vr[nbr], vg[nbg], N_r[nbr], N_g[nbg], massr[nbr], massg[nbg];
        for (int n2 = 0; n2 < nbr; ++n2) vr[n2] = vr_poly(Dr[n2]);
        for (int n = 0; n < nbg; ++n) vg[n] = av_g * std::pow(Dg[n], (double)bv_g);
        for (int n2 = 0; n2 < nbr; ++n2) massr[n2] = am_r * std::pow(Dr[n2], (double)bm_r);
        for (int n = 0; n < nbg; ++n) massg[n] = am_g * std::pow(Dg[n], (double)bm_g);
        const int m = km / ntb_r1, k = km % ntb_r1;
        double lam_exp = std::pow((N0r_exp[k] * am_r * crg[0] / r_r[m]), ore1);
        double lamr = lam_exp * std::pow((crg[2] * org2 * org1), obmr);
        double N0_r = N0r_exp[k] / (crg[1] * lam_exp) * std::pow(lamr, (double)cre[1]);
        for (int n2 = 0; n2 < nbr; ++n2)
            N_r[n2] = N0_r * std::pow(Dr[n2], (double)mu_r) * std::exp(-lamr * Dr[n2]) * dtr[n2];
        const float pef = PI2 * .25f * Ef_rg;
        for (int ij = ij0; ij < ij1; ++ij) {
            const int j = ij / ntb_g1, i = ij % ntb_g1;
            lam_exp = std::pow((N0g_exp[i] * am_g * cgg[0] / r_g[j]), oge1);
            double lamg = lam_exp * std::pow((cgg[2] * ogg2 * ogg1), obmg);
            double N0_g = N0g_exp[i] / (cgg[1] * lam_exp) * std::pow(lamg, (double)cge[1]);
            for (int n = 0; n < nbg; ++n)
                N_g[n] = N0_g * std::pow(Dg[n], (double)mu_g) * std::exp(-lamg * Dg[n]) * dtg[n];
            double t1 = 0, t2 = 0, z1 = 0, z2 = 0, y1 = 0, y2 = 0;
            for (int n2 = 0; n2 < nbr; ++n2) {
                for (int n = 0; n < nbg; ++n) {
                    double dvg = 0.5 * ((vr[n2] - vg[n]) + std::fabs(vr[n2] - vg[n]));
                    double dvr = 0.5 * ((vg[n] - vr[n2]) + std::fabs(vg[n] - vr[n2]));
                    double c = pef * (Dg[n] + Dr[n2]) * (Dg[n] + Dr[n2]);
                    t1 = t1 + c * dvg * massg[n] * N_g[n] * N_r[n2];
                    z1 = z1 + c * dvg * massr[n2] * N_g[n] * N_r[n2];
                    y1 = y1 + c * dvg * N_g[n] * N_r[n2];
                    t2 = t2 + c * dvr * massr[n2] * N_g[n] * N_r[n2];
                    y2 = y2 + c * dvr * N_g[n] * N_r[n2];
                    z2 = z2 + c * dvr * massg[n] * N_g[n] * N_r[n2];
                }
            }
            const size_t a = i4(i, j, k, m, ntb_g1, ntb_g, ntb_r1);
            tcg_racg[a] = t1;
            tmr_racg[a] = std::min(z1, r_r[m] * 1.0);
            tcr_gacr[a] = t2;
            tmg_gacr[a] = z2;
            tnr_racg[a] = y1;
            tnr_gacr[a] = y2;
        }
    }

    // ------------------------------------------------------------------ qr_acr_qs :3014-3264
    void qr_acr_qs_km(int km, int ij0, int ij1) {
        using namespace th;
        double vr[nbr], vs[nbs], N_r[nbr], N_s[nbs], massr[nbr], masss[nbs];
        for (int n2 = 0; n2 < nbr; ++n2) vr[n2] = vr_poly(Dr[n2]);
        for (int n = 0; n < nbs; ++n) vs[n] = 1.5f * av_s * std::pow(Ds[n], (double)bv_s) * std::exp(-fv_s * Ds[n]);
        for (int n2 = 0; n2 < nbr; ++n2) massr[n2] = am_r * std::pow(Dr[n2], (double)bm_r);
        for (int n = 0; n < nbs; ++n) masss[n] = am_s * std::pow(Ds[n], (double)bm_s);
        const int m = km / ntb_r1, k = km % ntb_r1;
        double lam_exp = std::pow((N0r_exp[k] * am_r * crg[0] / r_r[m]), ore1);
        double lamr = lam_exp * std::pow((crg[2] * org2 * org1), obmr);
        double N0_r = N0r_exp[k] / (crg[1] * lam_exp) * std::pow(lamr, (double)cre[1]);
        for (int n2 = 0; n2 < nbr; ++n2)
            N_r[n2] = N0_r * std::pow(Dr[n2], (double)mu_r) * std::exp(-lamr * Dr[n2]) * dtr[n2];
        const float pef = PI2 * .25f * Ef_rs;
        for (int ij = ij0; ij < ij1; ++ij) {
            const int j = ij / ntb_s, i = ij % ntb_s;
            double M2 = r_s[i] * oams * 1.0, second;
            if (bm_s > 2.0f - 1.E-3f && bm_s < 2.0f + 1.E-3f) {          // :3112 (sic: taken when bm_s ~ 2)
                double loga_ = sa_poly(sa, Tc[j], bm_s);
                double a_ = std::pow(10.0, loga_);
                double b_ = sa_poly(sb, Tc[j], bm_s);
                second = std::pow((M2 / a_), (1. / b_));
            } else {
                second = M2;
            }
            double loga_ = sa_poly(sa, Tc[j], cse[0]);
            double a_ = std::pow(10.0, loga_);
            double b_ = sa_poly(sb, Tc[j], cse[0]);
            double M3 = a_ * std::pow(second, b_);
            double oM3 = 1. / M3;
            double Mrat = M2 * (M2 * oM3) * (M2 * oM3) * (M2 * oM3);
            double M0 = std::pow((M2 * oM3), (double)mu_s);
            double slam1 = M2 * oM3 * Lam0;
            double slam2 = M2 * oM3 * Lam1;
            for (int n = 0; n < nbs; ++n)
                N_s[n] = Mrat * (Kap0 * std::exp(-slam1 * Ds[n]) +
                                 Kap1 * M0 * std::pow(Ds[n], (double)mu_s) * std::exp(-slam2 * Ds[n])) * dts[n];
            double t1 = 0, t2 = 0, t3 = 0, t4 = 0, z1 = 0, z2 = 0, z3 = 0, z4 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0;
            for (int n2 = 0; n2 < nbr; ++n2) {
                for (int n = 0; n < nbs; ++n) {
                    double dvs = 0.5 * ((vr[n2] - vs[n]) + std::fabs(vr[n2] - vs[n]));
                    double dvr = 0.5 * ((vs[n] - vr[n2]) + std::fabs(vs[n] - vr[n2]));
                    double c = pef * (Ds[n] + Dr[n2]) * (Ds[n] + Dr[n2]);
                    if (massr[n2] > 1.5f * masss[n]) {
                        t1 = t1 + c * dvs * masss[n] * N_s[n] * N_r[n2];
                        z1 = z1 + c * dvs * massr[n2] * N_s[n] * N_r[n2];
                        y1 = y1 + c * dvs * N_s[n] * N_r[n2];
                        t2 = t2 + c * dvr * massr[n2] * N_s[n] * N_r[n2];
                        y2 = y2 + c * dvr * N_s[n] * N_r[n2];
                        z2 = z2 + c * dvr * masss[n] * N_s[n] * N_r[n2];
                    } else {
                        t3 = t3 + c * dvs * masss[n] * N_s[n] * N_r[n2];
                        z3 = z3 + c * dvs * massr[n2] * N_s[n] * N_r[n2];
                        y3 = y3 + c * dvs * N_s[n] * N_r[n2];
                        t4 = t4 + c * dvr * massr[n2] * N_s[n] * N_r[n2];
                        y4 = y4 + c * dvr * N_s[n] * N_r[n2];
                        z4 = z4 + c * dvr * masss[n] * N_s[n] * N_r[n2];
                    }
                }
            }
            const size_t a = i4(i, j, k, m, ntb_s, ntb_t, ntb_r1);
            tcs_racs1[a] = t1;
            tmr_racs1[a] = std::min(z1, r_r[m] * 1.0);
            tcs_racs2[a] = t3;
            tmr_racs2[a] = z3;
            tcr_sacr1[a] = t2;
            tms_sacr1[a] = z2;
            tcr_sacr2[a] = t4;
            tms_sacr2[a] = z4;
            tnr_racs1[a] = y1;
            tnr_racs2[a] = y3;
            tnr_sacr1[a] = y2;
            tnr_sacr2[a] = y4;
        }
    }

    // ------------------------------------------------------------------ freezeH2O :3273-3399
    void freezeH2O() {
        using namespace th;
        tpi_qrfz.assign((size_t)ntb_r * ntb_r1 * 45, 0.0); tni_qrfz = tpi_qrfz; tpg_qrfz = tpi_qrfz; tnr_qrfz = tpi_qrfz;
        tpi_qcfz.assign((size_t)ntb_c * 45, 0.0); tni_qcfz = tpi_qcfz;
        double orho_w = 1.f / rho_w;
        double massr[nbr], massc[nbc], N_r[nbr], N_c[nbc];
        for (int n2 = 0; n2 < nbr; ++n2) massr[n2] = am_r * std::pow(Dr[n2], (double)bm_r);
        for (int n = 0; n < nbc; ++n) massc[n] = am_r * std::pow(Dc[n], (double)bm_r);
        for (int k = 1; k <= 45; ++k) {
            double Texp = std::exp((double)k - t_adjust * 1.0) - 1.0;
            for (int j = 0; j < ntb_r1; ++j)
                for (int i = 0; i < ntb_r; ++i) {
                    double lam_exp = std::pow((N0r_exp[j] * am_r * crg[0] / r_r[i]), ore1);
                    double lamr = lam_exp * std::pow((crg[2] * org2 * org1), obmr);
                    double N0_r = N0r_exp[j] / (crg[1] * lam_exp) * std::pow(lamr, (double)cre[1]);
                    double sum1 = 0, sum2 = 0, sumn1 = 0, sumn2 = 0;
                    for (int n2 = nbr - 1; n2 >= 0; --n2) {
                        N_r[n2] = N0_r * std::pow(Dr[n2], (double)mu_r) * std::exp(-lamr * Dr[n2]) * dtr[n2];
                        double vol = massr[n2] * orho_w;
                        double prob = 1.0 - std::exp(-120.0 * vol * 5.2E-4 * Texp);
                        prob = std::max(prob, 0.0);
                        if (massr[n2] < xm0g) {
                            sumn1 = sumn1 + prob * N_r[n2];
                            sum1 = sum1 + prob * N_r[n2] * massr[n2];
                        } else {
                            sumn2 = sumn2 + prob * N_r[n2];
                            sum2 = sum2 + prob * N_r[n2] * massr[n2];
                        }
                        if ((sum1 + sum2) >= r_r[i]) break;
                    }
                    const size_t a = i3(i, j, k - 1, ntb_r, ntb_r1);
                    tpi_qrfz[a] = sum1; tni_qrfz[a] = sumn1; tpg_qrfz[a] = sum2; tnr_qrfz[a] = sumn2;
                }
            for (int i = 0; i < ntb_c; ++i) {
                double lamc = 1.0E-6 * std::pow((Nt_c * am_r * ccg[1] * ocg1 / r_c[i]), obmr);
                double N0_c = 1.0E-18 * Nt_c * ocg1 * std::pow(lamc, (double)cce[0]);
                double sum1 = 0, sumn2 = 0;
                for (int n = nbc - 1; n >= 0; --n) {
                    double y = Dc[n] * 1.0E6;
                    double vol = massc[n] * orho_w;
                    double prob = 1.0 - std::exp(-120.0 * vol * 5.2E-4 * Texp);
                    prob = std::max(prob, 0.0);
                    N_c[n] = N0_c * std::pow(y, (double)mu_c) * std::exp(-lamc * y) * dtc[n];
                    N_c[n] = 1.0E24 * N_c[n];
                    sumn2 = sumn2 + prob * N_c[n];
                    sum1 = sum1 + prob * N_c[n] * massc[n];
                    if (sum1 >= r_c[i]) break;
                }
                tpi_qcfz[i2(i, k - 1, ntb_c)] = sum1;
                tni_qcfz[i2(i, k - 1, ntb_c)] = sumn2;
            }
        }
    }

    // ------------------------------------------------------------------ qi_aut_qs :3413-3456
    void qi_aut_qs() {
        using namespace th;
        tps_iaus.assign((size_t)ntb_i * ntb_i1, 0.0); tni_iaus = tps_iaus; tpi_ide = tps_iaus;
        double N_i[nbi];
        for (int j = 0; j < ntb_i1; ++j)
            for (int i = 0; i < ntb_i; ++i) {
                double lami = std::pow((am_i * cig[1] * oig1 * Nt_i[j] / r_i[i]), obmi);
                double Di_mean = (bm_i + mu_i + 1.f) / lami;
                double N0_i = Nt_i[j] * oig1 * std::pow(lami, (double)cie[0]);
                double t1 = 0, t2 = 0;
                const size_t a = i2(i, j, ntb_i);
                if ((float)Di_mean > 5.f * D0s) {
                    t1 = r_i[i]; t2 = Nt_i[j]; tpi_ide[a] = 0.0;
                } else if ((float)Di_mean < D0i) {
                    t1 = 0; t2 = 0; tpi_ide[a] = 1.0;
                } else {
                    float xlimit_intg = (float)(lami * D0s);
                    tpi_ide[a] = GAMMP(mu_i + 2.0f, xlimit_intg) * 1.0;
                    for (int n2 = 0; n2 < nbi; ++n2) {
                        N_i[n2] = N0_i * std::pow(Di[n2], (double)mu_i) * std::exp(-lami * Di[n2]) * dti[n2];
                        if (Di[n2] >= D0s) {
                            t1 = t1 + N_i[n2] * am_i * std::pow(Di[n2], (double)bm_i);
                            t2 = t2 + N_i[n2];
                        }
                    }
                }
                tps_iaus[a] = t1;
                tni_iaus[a] = t2;
            }
    }

    // ------------------------------------------------------------------ table_Efrw :3464-3525
    void table_Efrw() {
        using namespace th;
        t_Efrw.assign((size_t)nbr * nbc, 0.0);
        for (int j = 0; j < nbc; ++j)
            for (int i = 0; i < nbr; ++i) {
                double Ef_rw = 0.0;
                double p = Dc[j] / Dr[i];
                if (Dr[i] < 50.E-6f || Dc[j] < 3.E-6f) {
                    t_Efrw[i2(i, j, nbr)] = 0.0;
                } else if (p > 0.25f) {
                    double X = Dc[j] * 1.E6;
                    if (Dr[i] < 75.e-6f) Ef_rw = 0.026794f * X - 0.20604f;
                    else if (Dr[i] < 125.e-6f) Ef_rw = -0.00066842f * X * X + 0.061542f * X - 0.37089f;
                    else if (Dr[i] < 175.e-6f)
                        Ef_rw = 4.091e-06f * X * X * X * X - 0.00030908f * X * X * X + 0.0066237f * X * X - 0.0013687f * X - 0.073022f;
                    else if (Dr[i] < 250.e-6f) Ef_rw = 9.6719e-5f * X * X * X - 0.0068901f * X * X + 0.17305f * X - 0.65988f;
                    else if (Dr[i] < 350.e-6f) Ef_rw = 9.0488e-5f * X * X * X - 0.006585f * X * X + 0.16606f * X - 0.56125f;
                    else Ef_rw = 0.00010721f * X * X * X - 0.0072962f * X * X + 0.1704f * X - 0.46929f;
                } else {
                    double vtr = vr_poly(Dr[i]);
                    double stokes = Dc[j] * Dc[j] * vtr * rho_w / (9.f * 1.718E-5f * Dr[i]);
                    double reynolds = 9.f * stokes / (p * p * rho_w);
                    double F = std::log(reynolds);
                    double G = -0.1007 - 0.358 * F + 0.0261 * F * F;
                    double K0 = std::exp(G);
                    double z = std::log(stokes / (K0 + 1.E-15));
                    double H = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
                    double yc0 = 2.0 / PI2 * std::atan(H);
                    Ef_rw = (yc0 + p) * (yc0 + p) / ((1.f + p) * (1.f + p));
                }
                t_Efrw[i2(i, j, nbr)] = std::max(0.0f, std::min((float)Ef_rw, 0.95f));
                if (Ef_rw_l && Ef_rw != 0.0) t_Efrw[i2(i, j, nbr)] = 1.0;
            }
    }

    // ------------------------------------------------------------------ table_Efsw :3533-3578
    void table_Efsw() {
        using namespace th;
        t_Efsw.assign((size_t)nbs * nbc, 0.0);
        for (int j = 0; j < nbc; ++j) {
            double vtc = 1.19E4 * (1.0E4 * Dc[j] * Dc[j] * 0.25);
            for (int i = 0; i < nbs; ++i) {
                double vts = av_s * std::pow(Ds[i], (double)bv_s) * std::exp(-fv_s * Ds[i]) - vtc;
                double Ds_m = std::pow((am_s * std::pow(Ds[i], (double)bm_s) / am_r), (double)obmr);
                double p = Dc[j] / Ds_m;
                if (p > 0.25f || Ds[i] < D0s || Dc[j] < 6.E-6f || vts < 1.E-3f) {
                    t_Efsw[i2(i, j, nbs)] = 0.0;
                } else {
                    double stokes = Dc[j] * Dc[j] * vts * rho_w / (9.f * 1.718E-5f * Ds_m);
                    double reynolds = 9.f * stokes / (p * p * rho_w);
                    double F = std::log(reynolds);
                    double G = -0.1007 - 0.358 * F + 0.0261 * F * F;
                    double K0 = std::exp(G);
                    double z = std::log(stokes / (K0 + 1.E-15));
                    double H = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
                    double yc0 = 2.0 / PI2 * std::atan(H);
                    double Ef_sw = (yc0 + p) * (yc0 + p) / ((1.f + p) * (1.f + p));
                    t_Efsw[i2(i, j, nbs)] = std::max(0.0f, std::min((float)Ef_sw, 0.95f));
                    if (Ef_sw_l && Ef_sw != 0.0) t_Efsw[i2(i, j, nbs)] = 1.0;
                }
            }
        }
    }

    // lookup index into a decade table (:1562-1676).  x REAL version / N0_exp DOUBLE version.
    static int decade_index(float x, int n2, int ntb) {
        int nic = th::nint_f(std::log10(x)), n = nic - 1;
        for (int nn = nic - 1; nn <= nic + 1; ++nn) {
            n = nn;
            if ((x / th::powi_f(10.f, nn)) >= 1.0f && (x / th::powi_f(10.f, nn)) < 10.0f) break;
        }
        int idx = (int)(x / th::powi_f(10.f, n)) + 10 * (n - n2) - (n - n2);
        return std::max(1, std::min(idx, ntb));
    }
    static int decade_index_d(double x, int n2, int ntb) {
        int nic = th::nint_d(std::log10(x)), n = nic - 1;
        for (int nn = nic - 1; nn <= nic + 1; ++nn) {
            n = nn;
            if ((x / th::powi_f(10.f, nn)) >= 1.0 && (x / th::powi_f(10.f, nn)) < 10.0) break;
        }
        int idx = (int)(x / th::powi_f(10.f, n)) + 10 * (n - n2) - (n - n2);
        return std::max(1, std::min(idx, ntb));
    }

    // ------------------------------------------------------------------ mp_thompson :1057-2844
    // One column; arrays are kts..kte (nk levels, index 0 = kts).  Returns true when the column took
    // the `no_micro` early exit (:1363).
    bool column(float* qv1d, float* qc1d, float* qi1d, float* qr1d, float* qs1d, float* qg1d, float* ni1d,
                float* nr1d, float* t1d, const float* p1d, const float* dzq, float& pptrain, float& pptsnow,
                float& pptgraul, float& pptice, int nk, float dt) const {
        using namespace th;
        const int kts = 0, kte = nk - 1;
        std::vector<float> F((size_t)nk * 60 + 8, 0.f);
        std::vector<double> D((size_t)nk * 52, 0.0);
        int fo = 0, dd = 0;
        auto fa = [&](int extra = 0) { float* p = F.data() + fo; fo += nk + extra; return p; };
        auto da = [&]() { double* p = D.data() + dd; dd += nk; return p; };
        float *tten = fa(), *qvten = fa(), *qcten = fa(), *qiten = fa(), *qrten = fa(), *qsten = fa(), *qgten = fa(),
              *niten = fa(), *nrten = fa();
        double* prw_vcd = da();
        double *prr_wau = da(), *prr_rcw = da(), *prr_rcs = da(), *prr_rcg = da(), *prr_sml = da(), *prr_gml = da(),
               *prr_rci = da(), *prv_rev = da(), *pnr_wau = da(), *pnr_rcs = da(), *pnr_rcg = da(), *pnr_rci = da(),
               *pnr_sml = da(), *pnr_gml = da(), *pnr_rev = da(), *pnr_rcr = da(), *pnr_rfz = da();
        double *pri_inu = da(), *pni_inu = da(), *pri_ihm = da(), *pni_ihm = da(), *pri_wfz = da(), *pni_wfz = da(),
               *pri_rfz = da(), *pni_rfz = da(), *pri_ide = da(), *pni_ide = da(), *pri_rci = da(), *pni_rci = da(),
               *pni_sci = da(), *pni_iau = da();
        double *prs_iau = da(), *prs_sci = da(), *prs_rcs = da(), *prs_scw = da(), *prs_sde = da(), *prs_ihm = da(),
               *prs_ide = da();
        double *prg_scw = da(), *prg_rfz = da(), *prg_gde = da(), *prg_gcw = da(), *prg_rci = da(), *prg_rcs = da(),
               *prg_rcg = da(), *prg_ihm = da();
        float *temp = fa(), *pres = fa(), *qv = fa();
        float *rc = fa(), *ri = fa(), *rr = fa(), *rs = fa(), *rg = fa(), *ni = fa(), *nr = fa();
        float *rho = fa(), *rhof = fa(), *rhof2 = fa();
        float *qvs = fa(), *qvsi = fa(), *delQvs = fa();
        float *satw = fa(), *sati = fa(), *ssatw = fa(), *ssati = fa();
        float *diffu = fa(), *visco = fa(), *vsc2 = fa(), *tcond = fa(), *lvap = fa(), *ocp = fa(), *lvt2 = fa();
        double *ilamr = da(), *ilamg = da(), *N0_r = da(), *N0_g = da();
        float *mvd_r = fa(), *mvd_c = fa();
        float *smob = fa(), *smo2 = fa(), *smo1 = fa(), *smo0 = fa(), *smoc = fa(), *smod = fa(), *smoe = fa(), *smof = fa();
        float *sed_r = fa(), *sed_s = fa(), *sed_g = fa(), *sed_i = fa(), *sed_n = fa();
        float *vtik = fa(1), *vtnik = fa(1), *vtrk = fa(1), *vtnrk = fa(1), *vtsk = fa(1), *vtgk = fa(1);
        float* vts_boost = fa();
        std::vector<char> Lb((size_t)nk * 5, 0);
        char *L_qc = Lb.data(), *L_qi = L_qc + nk, *L_qr = L_qi + nk, *L_qs = L_qr + nk, *L_qg = L_qs + nk;

        float orho, lfus2, onstep[4], xDc, Dc_b, Dc_g, xDi, xDs, xDg, zeta1, zeta, taud, tau, stoke_g, vti, vtr, vts, vtg;
        double N0_exp, N0_min, lam_exp, lamc, lamr, lamg, lami, ilami;
        float Mrat, ils1, ils2, t1_vts, t2_vts, t3_vts, t4_vts, C_snow, a_, b_, loga_, tf, tempc, tc0;
        float xnc, xri, xni, xmi, oxmi, xrc, xrr, xnr, xsat, rate_max, sump, ratio, clap, fcd, dfcd;
        float otemp, rvs = 0, rvs_p, rvs_pp, gamsc, alphsc, t1_evap, t1_subl = 0, r_frac, g_frac;
        float Ef_rw, Ef_sw, Ef_gw = 0, Ef_rr, dtsave, odts, odt, odzq, xslw1, ygra1, zans1, delta_tp;
        int idx_tc, idx_t, idx_s, idx_g1, idx_g, idx_r1, idx_r, idx_i1, idx_i, idx_c, idx, nstep, ksed1[4];
        bool no_micro = true;
        dtsave = dt; odt = 1.f / dt; odts = 1.f / dtsave;

        // :1240-1319
        for (int k = kts; k <= kte; ++k) {
            temp[k] = t1d[k];
            qv[k] = std::max(1.E-10f, qv1d[k]);
            pres[k] = p1d[k];
            rho[k] = 0.622f * pres[k] / (RR2 * temp[k] * (qv[k] + 0.622f));
            if (qc1d[k] > R1) { no_micro = false; rc[k] = qc1d[k] * rho[k]; L_qc[k] = 1; }
            else { qc1d[k] = 0.0f; rc[k] = R1; L_qc[k] = 0; }
            if (qi1d[k] > R1) {
                no_micro = false;
                ri[k] = qi1d[k] * rho[k];
                ni[k] = std::max(R2, ni1d[k] * rho[k]);
                L_qi[k] = 1;
                lami = std::pow((am_i * cig[1] * oig1 * ni[k] / ri[k]), obmi);
                ilami = 1. / lami;
                xDi = (float)((bm_i + mu_i + 1.f) * ilami);
                if (xDi < 20.E-6f) {
                    lami = cie[1] / 20.E-6f;
                    ni[k] = (float)std::min(250.E3, cig[0] * oig2 * ri[k] / am_i * std::pow(lami, (double)bm_i));
                } else if (xDi > 300.E-6f) {
                    lami = cie[1] / 300.E-6f;
                    ni[k] = (float)(cig[0] * oig2 * ri[k] / am_i * std::pow(lami, (double)bm_i));
                }
            } else { qi1d[k] = 0.0f; ni1d[k] = 0.0f; ri[k] = R1; ni[k] = R2; L_qi[k] = 0; }
            mvd_r[k] = 0.0f;
            if (qr1d[k] > R1) {
                no_micro = false;
                rr[k] = qr1d[k] * rho[k];
                nr[k] = std::max(R2, nr1d[k] * rho[k]);
                L_qr[k] = 1;
                lamr = std::pow((am_r * crg[2] * org2 * nr[k] / rr[k]), obmr);
                mvd_r[k] = (float)((3.0f + mu_r + 0.672f) / lamr);
                if (mvd_r[k] > 2.5E-3f) {
                    mvd_r[k] = 2.5E-3f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    nr[k] = (float)(crg[1] * org3 * rr[k] * std::pow(lamr, (double)bm_r) / am_r);
                } else if (mvd_r[k] < D0r * 0.75f) {
                    mvd_r[k] = D0r * 0.75f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    nr[k] = (float)(crg[1] * org3 * rr[k] * std::pow(lamr, (double)bm_r) / am_r);
                }
            } else { qr1d[k] = 0.0f; nr1d[k] = 0.0f; rr[k] = R1; nr[k] = R2; L_qr[k] = 0; }
            if (qs1d[k] > R1) { no_micro = false; rs[k] = qs1d[k] * rho[k]; L_qs[k] = 1; }
            else { qs1d[k] = 0.0f; rs[k] = R1; L_qs[k] = 0; }
            if (qg1d[k] > R1) { no_micro = false; rg[k] = qg1d[k] * rho[k]; L_qg[k] = 1; }
            else { qg1d[k] = 0.0f; rg[k] = R1; L_qg[k] = 0; }
        }

        // :1328-1356
        for (int k = kts; k <= kte; ++k) {
            tempc = temp[k] - 273.15f;
            rhof[k] = std::sqrt(rho_not / rho[k]);
            rhof2[k] = std::sqrt(rhof[k]);
            qvs[k] = RSLF(pres[k], temp[k]);
            delQvs[k] = std::max(0.0f, RSLF(pres[k], 273.15f) - qv[k]);
            if (tempc <= 0.0f) qvsi[k] = RSIF(pres[k], temp[k]);
            else qvsi[k] = qvs[k];
            satw[k] = qv[k] / qvs[k];
            sati[k] = qv[k] / qvsi[k];
            ssatw[k] = satw[k] - 1.f;
            ssati[k] = sati[k] - 1.f;
            if (std::fabs(ssatw[k]) < eps) ssatw[k] = 0.0f;
            if (std::fabs(ssati[k]) < eps) ssati[k] = 0.0f;
            if (no_micro && ssati[k] > 0.0f) no_micro = false;
            diffu[k] = 2.11E-5f * std::pow((temp[k] / 273.15f), 1.94f) * (101325.f / pres[k]);
            if (tempc >= 0.0f) visco[k] = (1.718f + 0.0049f * tempc) * 1.0E-5f;
            else visco[k] = (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
            ocp[k] = 1.f / (Cp2 * (1.f + 0.887f * qv[k]));
            vsc2[k] = std::sqrt(rho[k] / visco[k]);
            lvap[k] = lvap0 + (2106.0f - 4218.0f) * tempc;
            tcond[k] = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
        }

        if (no_micro) return true;                                         // :1363

        // snow moments :1369-1451 (first pass: smo0, smo1, smoc, smoe, smof)
        for (int k = kts; k <= kte; ++k) {
            if (!L_qs[k]) continue;
            tc0 = std::min(-0.1f, temp[k] - 273.15f);
            smob[k] = rs[k] * oams;
            if (bm_s > (2.0f - 1.e-3f) && bm_s < (2.0f + 1.e-3f)) {
                smo2[k] = smob[k];
            } else {
                loga_ = sa_poly(sa, tc0, bm_s);
                a_ = std::pow(10.0f, loga_);
                b_ = sa_poly(sb, tc0, bm_s);
                smo2[k] = std::pow((smob[k] / a_), (1.f / b_));
            }
            loga_ = sa[0] + sa[1] * tc0 + sa[4] * tc0 * tc0 + sa[8] * tc0 * tc0 * tc0;                  // :1394
            a_ = std::pow(10.0f, loga_);
            b_ = sb[0] + sb[1] * tc0 + sb[4] * tc0 * tc0 + sb[8] * tc0 * tc0 * tc0;
            smo0[k] = a_ * std::pow(smo2[k], b_);
            loga_ = sa[0] + sa[1] * tc0 + sa[2] + sa[3] * tc0 + sa[4] * tc0 * tc0 + sa[5] + sa[6] * tc0 * tc0 +
                    sa[7] * tc0 + sa[8] * tc0 * tc0 * tc0 + sa[9];                                      // :1400
            a_ = std::pow(10.0f, loga_);
            b_ = sb[0] + sb[1] * tc0 + sb[2] + sb[3] * tc0 + sb[4] * tc0 * tc0 + sb[5] + sb[6] * tc0 * tc0 +
                 sb[7] * tc0 + sb[8] * tc0 * tc0 * tc0 + sb[9];
            smo1[k] = a_ * std::pow(smo2[k], b_);
            loga_ = sa_poly(sa, tc0, cse[0]); a_ = std::pow(10.0f, loga_); b_ = sa_poly(sb, tc0, cse[0]);
            smoc[k] = a_ * std::pow(smo2[k], b_);
            loga_ = sa_poly(sa, tc0, cse[12]); a_ = std::pow(10.0f, loga_); b_ = sa_poly(sb, tc0, cse[12]);
            smoe[k] = a_ * std::pow(smo2[k], b_);
            loga_ = sa_poly(sa, tc0, cse[15]); a_ = std::pow(10.0f, loga_); b_ = sa_poly(sb, tc0, cse[15]);
            smof[k] = a_ * std::pow(smo2[k], b_);
        }

        // graupel intercept / slope :1456-1482, rain :1489-1494 (shared with :2379-2408)
        auto graupel_rain_slopes = [&]() {
            N0_min = gonv_max;
            for (int k = kte; k >= kts; --k) {
                if (temp[k] < 270.65f && L_qr[k] && mvd_r[k] > 100.E-6f) xslw1 = 4.01f + std::log10(mvd_r[k]);
                else xslw1 = 0.01f;
                ygra1 = 4.31f + std::log10(std::max(5.E-5f, rg[k]));
                zans1 = 3.1f + (100.f / (300.f * xslw1 * ygra1 / (10.f / xslw1 + 1.f + 0.25f * ygra1) + 30.f + 10.f * ygra1));
                N0_exp = std::pow(10.f, (zans1));
                N0_exp = std::max((double)gonv_min, std::min(N0_exp, (double)gonv_max));
                N0_min = std::min(N0_exp, N0_min);
                N0_exp = N0_min;
                lam_exp = std::pow((N0_exp * am_g * cgg[0] / rg[k]), (double)oge1);
                lamg = lam_exp * std::pow((cgg[2] * ogg2 * ogg1), obmg);
                ilamg[k] = 1. / lamg;
                N0_g[k] = N0_exp / (cgg[1] * lam_exp) * std::pow(lamg, (double)cge[1]);
            }
            for (int k = kte; k >= kts; --k) {
                lamr = std::pow((am_r * crg[2] * org2 * nr[k] / rr[k]), obmr);
                ilamr[k] = 1. / lamr;
                mvd_r[k] = (float)((3.0f + mu_r + 0.672f) / lamr);
                N0_r[k] = nr[k] * org2 * std::pow(lamr, (double)cre[1]);
            }
        };
        graupel_rain_slopes();

        // warm rain :1500-1544
        for (int k = kts; k <= kte; ++k) {
            if (L_qr[k] && mvd_r[k] > D0r) {
                Ef_rr = 2.0f - std::exp(2300.0f * (mvd_r[k] - 1600.0E-6f));
                pnr_rcr[k] = Ef_rr * 4.f * nr[k] * rr[k];
            }
            mvd_c[k] = D0c;
            if (!L_qc[k]) continue;
            xDc = std::max(D0c * 1.E6f, (std::pow((rc[k] / (am_r * Nt_c)), obmr)) * 1.E6f);
            lamc = std::pow((Nt_c * am_r * ccg[1] * ocg1 / rc[k]), obmr);
            mvd_c[k] = (float)((3.0f + mu_c + 0.672f) / lamc);
            if (rc[k] > 0.01e-3f) {
                Dc_g = (float)((std::pow((ccg[2] * ocg2), obmr) / lamc) * 1.E6f);
                Dc_b = std::pow((xDc * xDc * xDc * Dc_g * Dc_g * Dc_g - xDc * xDc * xDc * xDc * xDc * xDc), (1.f / 6.f));
                zeta1 = 0.5f * ((6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f) + std::fabs(6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f));
                zeta = 0.027f * rc[k] * zeta1;
                taud = 0.5f * ((0.5f * Dc_b - 7.5f) + std::fabs(0.5f * Dc_b - 7.5f)) + R1;
                tau = 3.72f / (rc[k] * taud);
                prr_wau[k] = zeta / tau;
                prr_wau[k] = std::min((double)(rc[k] * odts), prr_wau[k]);
                pnr_wau[k] = prr_wau[k] / (am_r * mu_c * D0r * D0r * D0r);
            }
            if (L_qr[k] && mvd_r[k] > D0r && mvd_c[k] > D0c) {
                lamr = 1. / ilamr[k];
                idx = 1 + (int)(nbr * std::log(mvd_r[k] / Dr[0]) / std::log(Dr[nbr - 1] / Dr[0]));
                idx = std::min(idx, nbr);
                Ef_rw = (float)t_Efrw[i2(idx - 1, std::min((int)(mvd_c[k] * 1.E6f), nbc) - 1, nbr)];   // (clamp: memory safety only)
                prr_rcw[k] = rhof[k] * t1_qr_qc * Ef_rw * rc[k] * N0_r[k] * (std::pow((lamr + fv_r), (double)(-cre[8])));
                prr_rcw[k] = std::min((double)(rc[k] * odts), prr_rcw[k]);
            }
        }

        // frozen processes :1550-2009
        for (int k = kts; k <= kte; ++k) {
            vts_boost[k] = 1.5f;
            tempc = temp[k] - 273.15f;
            idx_tc = std::max(1, std::min(nint_f(-tempc), 45));
            idx_t = (int)((tempc - 2.5f) / 5.f) - 1;
            idx_t = std::max(1, -idx_t);
            idx_t = std::min(idx_t, ntb_t);
            idx_c = (rc[k] > r_c[0]) ? decade_index(rc[k], nic2, ntb_c) : 1;
            idx_i = (ri[k] > r_i[0]) ? decade_index(ri[k], nii2, ntb_i) : 1;
            idx_i1 = (ni[k] > Nt_i[0]) ? decade_index(ni[k], nii3, ntb_i1) : 1;
            if (rr[k] > r_r[0]) {
                idx_r = decade_index(rr[k], nir2, ntb_r);
                lamr = 1. / ilamr[k];
                lam_exp = lamr * std::pow((crg[2] * org2 * org1), bm_r);
                N0_exp = org1 * rr[k] / am_r * std::pow(lam_exp, (double)cre[0]);
                idx_r1 = decade_index_d(N0_exp, nir3, ntb_r1);
            } else { idx_r = 1; idx_r1 = ntb_r1; }
            idx_s = (rs[k] > r_s[0]) ? decade_index(rs[k], nis2, ntb_s) : 1;
            if (rg[k] > r_g[0]) {
                idx_g = decade_index(rg[k], nig2, ntb_g);
                lamg = 1. / ilamg[k];
                lam_exp = lamg * std::pow((cgg[2] * ogg2 * ogg1), bm_g);
                N0_exp = ogg1 * rg[k] / am_g * std::pow(lam_exp, (double)cge[0]);
                idx_g1 = decade_index_d(N0_exp, nig3, ntb_g1);
            } else { idx_g = 1; idx_g1 = ntb_g1; }

            otemp = 1.f / temp[k];                                           // :1679-1695
            rvs = rho[k] * qvsi[k];
            rvs_p = rvs * otemp * (lsub * otemp * oRv - 1.f);
            rvs_pp = rvs * (otemp * (lsub * otemp * oRv - 1.f) * otemp * (lsub * otemp * oRv - 1.f) +
                            (-2.f * lsub * otemp * otemp * otemp * oRv) + otemp * otemp);
            gamsc = lsub * diffu[k] / tcond[k] * rvs_p;
            alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
            alphsc = std::max(1.E-9f, alphsc);
            xsat = ssati[k];
            if (std::fabs(xsat) < 1.E-9f) xsat = 0.f;
            t1_subl = 4.f * PI2 * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat -
                                   5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);

            if (L_qc[k] && mvd_c[k] > D0c) {                                 // :1698-1725
                xDs = 0.0f;
                if (L_qs[k]) xDs = smoc[k] / smob[k];
                if (xDs > D0s) {
                    idx = 1 + (int)(nbs * std::log(xDs / Ds[0]) / std::log(Ds[nbs - 1] / Ds[0]));
                    idx = std::min(idx, nbs);
                    Ef_sw = (float)t_Efsw[i2(idx - 1, std::min((int)(mvd_c[k] * 1.E6f), nbc) - 1, nbs)];
                    prs_scw[k] = rhof[k] * t1_qs_qc * Ef_sw * rc[k] * smoe[k];
                }
                if (rg[k] >= r_g[0] && mvd_c[k] > D0c) {
                    xDg = (float)((bm_g + mu_g + 1.f) * ilamg[k]);
                    vtg = (float)(rhof[k] * av_g * cgg[5] * ogg3 * std::pow(ilamg[k], (double)bv_g));
                    stoke_g = mvd_c[k] * mvd_c[k] * vtg * rho_w / (9.f * visco[k] * xDg);
                    if (xDg > D0g) {
                        if (stoke_g >= 0.4f && stoke_g <= 10.f) Ef_gw = 0.55f * std::log10(2.51f * stoke_g);
                        else if (stoke_g < 0.4f) Ef_gw = 0.0f;
                        else if (stoke_g > 10.f) Ef_gw = 0.77f;
                        prg_gcw[k] = rhof[k] * t1_qg_qc * Ef_gw * rc[k] * N0_g[k] * std::pow(ilamg[k], (double)cge[8]);
                    }
                }
            }

            if (rr[k] >= r_r[0]) {                                           // :1730-1783
                if (rs[k] >= r_s[0]) {
                    const size_t a = i4(idx_s - 1, idx_t - 1, idx_r1 - 1, idx_r - 1, ntb_s, ntb_t, ntb_r1);
                    if (temp[k] < T_0) {
                        prr_rcs[k] = -(tmr_racs2[a] + tcr_sacr2[a] + tmr_racs1[a] + tcr_sacr1[a]);
                        prs_rcs[k] = tmr_racs2[a] + tcr_sacr2[a] - tcs_racs1[a] - tms_sacr1[a];
                        prg_rcs[k] = tmr_racs1[a] + tcr_sacr1[a] + tcs_racs1[a] + tms_sacr1[a];
                        prr_rcs[k] = std::max((double)(-rr[k] * odts), prr_rcs[k]);
                        prs_rcs[k] = std::max((double)(-rs[k] * odts), prs_rcs[k]);
                        prg_rcs[k] = std::min((double)((rr[k] + rs[k]) * odts), prg_rcs[k]);
                        pnr_rcs[k] = tnr_racs1[a] + tnr_racs2[a] + tnr_sacr1[a] + tnr_sacr2[a];
                    } else {
                        prs_rcs[k] = -tcs_racs1[a] - tms_sacr1[a] + tmr_racs2[a] + tcr_sacr2[a];
                        prs_rcs[k] = std::max((double)(-rs[k] * odts), prs_rcs[k]);
                        prr_rcs[k] = -prs_rcs[k];
                        pnr_rcs[k] = tnr_racs2[a] + tnr_sacr2[a];
                    }
                    pnr_rcs[k] = std::min((double)(nr[k] * odts), pnr_rcs[k]);
                }
                if (rg[k] >= r_g[0]) {
                    const size_t a = i4(idx_g1 - 1, idx_g - 1, idx_r1 - 1, idx_r - 1, ntb_g1, ntb_g, ntb_r1);
                    if (temp[k] < T_0) {
                        prg_rcg[k] = tmr_racg[a] + tcr_gacr[a];
                        prg_rcg[k] = std::min((double)(rr[k] * odts), prg_rcg[k]);
                        prr_rcg[k] = -prg_rcg[k];
                        pnr_rcg[k] = tnr_racg[a] + tnr_gacr[a];
                        pnr_rcg[k] = std::min((double)(nr[k] * odts), pnr_rcg[k]);
                    } else {
                        prr_rcg[k] = tcg_racg[a];
                        prr_rcg[k] = std::min((double)(rg[k] * odts), prr_rcg[k]);
                        prg_rcg[k] = -prr_rcg[k];
                    }
                }
            }

            if (temp[k] < T_0) {                                             // :1789-1949
                vts_boost[k] = 1.0f;
                rate_max = (qv[k] - qvsi[k]) * rho[k] * odts * 0.999f;
                if (rr[k] > r_r[0]) {
                    const size_t a = i3(idx_r - 1, idx_r1 - 1, idx_tc - 1, ntb_r, ntb_r1);
                    prg_rfz[k] = tpg_qrfz[a] * odts;
                    pri_rfz[k] = tpi_qrfz[a] * odts;
                    pni_rfz[k] = tni_qrfz[a] * odts;
                    pnr_rfz[k] = tnr_qrfz[a] * odts;
                    pnr_rfz[k] = std::min((double)(nr[k] * odts), pnr_rfz[k]);
                } else if (rr[k] > R1 && temp[k] < HGFR) {
                    pri_rfz[k] = rr[k] * odts;
                    pnr_rfz[k] = nr[k] * odts;
                    pni_rfz[k] = pnr_rfz[k];
                }
                if (rc[k] > r_c[0]) {
                    const size_t a = i2(idx_c - 1, idx_tc - 1, ntb_c);
                    pri_wfz[k] = tpi_qcfz[a] * odts;
                    pri_wfz[k] = std::min((double)(rc[k] * odts), pri_wfz[k]);
                    pni_wfz[k] = tni_qcfz[a] * odts;
                    pni_wfz[k] = std::min(std::min((double)(Nt_c * odts), pri_wfz[k] / (2.f * xm0i)), pni_wfz[k]);
                } else if (rc[k] > R1 && temp[k] < HGFR) {
                    pri_wfz[k] = rc[k] * odts;
                    pni_wfz[k] = std::min(std::min((double)(Nt_c * odts), pri_wfz[k] / (2.f * xm0i)), pni_wfz[k]);
                }
                if ((ssati[k] >= 0.25f) || (ssatw[k] > eps && temp[k] < 261.15f)) {                     // :1820-1827
                    xnc = std::min(250.E3f, TNO * std::exp(ATO * (T_0 - temp[k])));
                    xni = (float)(ni[k] + (pni_rfz[k] + pni_wfz[k]) * dtsave);
                    pni_inu[k] = 0.5f * (xnc - xni + std::fabs(xnc - xni)) * odts;
                    pri_inu[k] = std::min((double)rate_max, xm0i * pni_inu[k]);
                    pni_inu[k] = pri_inu[k] / xm0i;
                }
                if (L_qi[k]) {                                                                          // :1830-1863
                    lami = std::pow((am_i * cig[1] * oig1 * ni[k] / ri[k]), obmi);
                    ilami = 1. / lami;
                    xDi = (float)std::max((double)D0i, (bm_i + mu_i + 1.f) * ilami);
                    xmi = am_i * std::pow(xDi, bm_i);
                    oxmi = 1.f / xmi;
                    pri_ide[k] = C_cube * t1_subl * diffu[k] * ssati[k] * rvs * oig1 * cig[4] * ni[k] * ilami;
                    if (pri_ide[k] < 0.0) {
                        pri_ide[k] = std::max(std::max((double)(-ri[k] * odts), pri_ide[k]), (double)rate_max);
                        pni_ide[k] = pri_ide[k] * oxmi;
                        pni_ide[k] = std::max((double)(-ni[k] * odts), pni_ide[k]);
                    } else {
                        pri_ide[k] = std::min(pri_ide[k], (double)rate_max);
                        const size_t a = i2(idx_i - 1, idx_i1 - 1, ntb_i);
                        prs_ide[k] = (1.0 - tpi_ide[a]) * pri_ide[k];
                        pri_ide[k] = tpi_ide[a] * pri_ide[k];
                    }
                    if ((idx_i == ntb_i) || (xDi > 5.0f * D0s)) {
                        prs_iau[k] = ri[k] * .99f * odts;
                        pni_iau[k] = ni[k] * .95f * odts;
                    } else if (xDi < 0.1f * D0s) {
                        prs_iau[k] = 0.; pni_iau[k] = 0.;
                    } else {
                        const size_t a = i2(idx_i - 1, idx_i1 - 1, ntb_i);
                        prs_iau[k] = tps_iaus[a] * odts;
                        prs_iau[k] = std::min((double)(ri[k] * .99f * odts), prs_iau[k]);
                        pni_iau[k] = tni_iaus[a] * odts;
                        pni_iau[k] = std::min((double)(ni[k] * .95f * odts), pni_iau[k]);
                    }
                }
                if (L_qs[k]) {                                                                          // :1867-1882
                    C_snow = C_sqrd + (tempc + 15.f) * (C_cubes - C_sqrd) / (-30.f + 15.f);
                    C_snow = std::max(C_sqrd, std::min(C_snow, C_cubes));
                    prs_sde[k] = C_snow * t1_subl * diffu[k] * ssati[k] * rvs *
                                 (t1_qs_sd * smo1[k] + t2_qs_sd * rhof2[k] * vsc2[k] * smof[k]);
                    if (prs_sde[k] < 0.) prs_sde[k] = std::max(std::max((double)(-rs[k] * odts), prs_sde[k]), (double)rate_max);
                    else prs_sde[k] = std::min(prs_sde[k], (double)rate_max);
                }
                if (L_qg[k] && ssati[k] < -eps) {                                                       // :1884-1893
                    prg_gde[k] = C_cube * t1_subl * diffu[k] * ssati[k] * rvs * N0_g[k] *
                                 (t1_qg_sd * std::pow(ilamg[k], (double)cge[9]) +
                                  t2_qg_sd * vsc2[k] * rhof2[k] * std::pow(ilamg[k], (double)cge[10]));
                    if (prg_gde[k] < 0.) prg_gde[k] = std::max(std::max((double)(-rg[k] * odts), prg_gde[k]), (double)rate_max);
                    else prg_gde[k] = std::min(prg_gde[k], (double)rate_max);
                }
                if (L_qi[k]) {                                                                          // :1896-1920
                    lami = std::pow((am_i * cig[1] * oig1 * ni[k] / ri[k]), obmi);
                    ilami = 1. / lami;
                    xDi = (float)std::max((double)D0i, (bm_i + mu_i + 1.f) * ilami);
                    xmi = am_i * std::pow(xDi, bm_i);
                    oxmi = 1.f / xmi;
                    if (rs[k] >= r_s[0]) {
                        prs_sci[k] = t1_qs_qi * rhof[k] * Ef_si * ri[k] * smoe[k];
                        pni_sci[k] = prs_sci[k] * oxmi;
                    }
                    if (rr[k] >= r_r[0] && mvd_r[k] > 4.f * xDi) {
                        lamr = 1. / ilamr[k];
                        pri_rci[k] = rhof[k] * t1_qr_qi * Ef_ri * ri[k] * N0_r[k] * (std::pow((lamr + fv_r), (double)(-cre[8])));
                        pnr_rci[k] = rhof[k] * t1_qr_qi * Ef_ri * ni[k] * N0_r[k] * (std::pow((lamr + fv_r), (double)(-cre[8])));
                        pni_rci[k] = pri_rci[k] * oxmi;
                        prr_rci[k] = rhof[k] * t2_qr_qi * Ef_ri * ni[k] * N0_r[k] * (std::pow((lamr + fv_r), (double)(-cre[7])));
                        prr_rci[k] = std::min((double)(rr[k] * odts), prr_rci[k]);
                        prg_rci[k] = pri_rci[k] + prr_rci[k];
                    }
                }
                if (prg_gcw[k] > eps && tempc > -8.0f) {                                                // :1923-1936
                    tf = 0.f;
                    if (tempc >= -5.0f && tempc < -3.0f) tf = 0.5f * (-3.0f - tempc);
                    else if (tempc > -8.0f && tempc < -5.0f) tf = 0.33333333f * (8.0f + tempc);
                    pni_ihm[k] = 3.5E8f * tf * prg_gcw[k];
                    pri_ihm[k] = xm0i * pni_ihm[k];
                    prs_ihm[k] = prs_scw[k] / (prs_scw[k] + prg_gcw[k]) * pri_ihm[k];
                    prg_ihm[k] = prg_gcw[k] / (prs_scw[k] + prg_gcw[k]) * pri_ihm[k];
                }
                if (prs_scw[k] > 5.0f * prs_sde[k] && prs_sde[k] > eps) {                               // :1942-1949
                    r_frac = (float)std::min(30.0, prs_scw[k] / prs_sde[k]);
                    g_frac = std::min(0.75f, 0.05f + (r_frac - 5.f) * .028f);
                    vts_boost[k] = std::min(1.5f, 1.1f + (r_frac - 5.f) * .016f);
                    prg_scw[k] = g_frac * prs_scw[k];
                    prs_scw[k] = (1.f - g_frac) * prs_scw[k];
                }
            } else {                                                                                    // :1951-2007
                if (L_qs[k]) {
                    prr_sml[k] = (tempc * tcond[k] - lvap0 * diffu[k] * delQvs[k]) *
                                 (t1_qs_me * smo1[k] + t2_qs_me * rhof2[k] * vsc2[k] * smof[k]);
                    prr_sml[k] = prr_sml[k] + 4218.f * olfus * tempc * (prr_rcs[k] + prs_scw[k]);
                    prr_sml[k] = std::min((double)(rs[k] * odts), std::max(0., prr_sml[k]));
                    pnr_sml[k] = smo0[k] / rs[k] * prr_sml[k] * std::pow(10.0f, (-0.75f * tempc));
                    pnr_sml[k] = std::min((double)(smo0[k] * odts), pnr_sml[k]);
                    if (tempc > 3.5f || rs[k] < 0.005E-3f) pnr_sml[k] = 0.0;
                    if (ssati[k] < 0.f) {
                        prs_sde[k] = C_cubes * t1_subl * diffu[k] * ssati[k] * rvs *
                                     (t1_qs_sd * smo1[k] + t2_qs_sd * rhof2[k] * vsc2[k] * smof[k]);
                        prs_sde[k] = std::max((double)(-rs[k] * odts), prs_sde[k]);
                    }
                }
                if (L_qg[k]) {
                    prr_gml[k] = (tempc * tcond[k] - lvap0 * diffu[k] * delQvs[k]) * N0_g[k] *
                                 (t1_qg_me * std::pow(ilamg[k], (double)cge[9]) +
                                  t2_qg_me * rhof2[k] * vsc2[k] * std::pow(ilamg[k], (double)cge[10]));
                    prr_gml[k] = std::min((double)(rg[k] * odts), std::max(0., prr_gml[k]));
                    pnr_gml[k] = N0_g[k] * cgg[1] * std::pow(ilamg[k], (double)cge[1]) / rg[k] * prr_gml[k] *
                                 std::pow(10.0f, (-1.5f * tempc));
                    if (tempc > 7.5f || rg[k] < 0.005E-3f) pnr_gml[k] = 0.0;
                    if (ssati[k] < 0.f) {
                        prg_gde[k] = C_cube * t1_subl * diffu[k] * ssati[k] * rvs * N0_g[k] *
                                     (t1_qg_sd * std::pow(ilamg[k], (double)cge[9]) +
                                      t2_qg_sd * vsc2[k] * rhof2[k] * std::pow(ilamg[k], (double)cge[10]));
                        prg_gde[k] = std::max((double)(-rg[k] * odts), prg_gde[k]);
                    }
                }
                if (dt > 120.f) {
                    prr_rcw[k] = prr_rcw[k] + prs_scw[k] + prg_gcw[k];
                    prs_scw[k] = 0.; prg_gcw[k] = 0.;
                }
            }
        }

        // conservation :2015-2110
        for (int k = kts; k <= kte; ++k) {
            sump = (float)(pri_inu[k] + pri_ide[k] + prs_ide[k] + prs_sde[k] + prg_gde[k]);
            rate_max = (qv[k] - qvsi[k]) * odts * 0.999f;
            if ((sump > eps && sump > rate_max) || (sump < -eps && sump < rate_max)) {
                ratio = rate_max / sump;
                pri_inu[k] *= ratio; pri_ide[k] *= ratio; pni_ide[k] *= ratio; prs_ide[k] *= ratio;
                prs_sde[k] *= ratio; prg_gde[k] *= ratio;
            }
            sump = (float)(-prr_wau[k] - pri_wfz[k] - prr_rcw[k] - prs_scw[k] - prg_scw[k] - prg_gcw[k]);
            rate_max = -rc[k] * odts;
            if (sump < rate_max && L_qc[k]) {
                ratio = rate_max / sump;
                prr_wau[k] *= ratio; pri_wfz[k] *= ratio; prr_rcw[k] *= ratio; prs_scw[k] *= ratio;
                prg_scw[k] *= ratio; prg_gcw[k] *= ratio;
            }
            sump = (float)(pri_ide[k] - prs_iau[k] - prs_sci[k] - pri_rci[k]);
            rate_max = -ri[k] * odts;
            if (sump < rate_max && L_qi[k]) {
                ratio = rate_max / sump;
                pri_ide[k] *= ratio; prs_iau[k] *= ratio; prs_sci[k] *= ratio; pri_rci[k] *= ratio;
            }
            sump = (float)(-prg_rfz[k] - pri_rfz[k] - prr_rci[k] + prr_rcs[k] + prr_rcg[k]);
            rate_max = -rr[k] * odts;
            if (sump < rate_max && L_qr[k]) {
                ratio = rate_max / sump;
                prg_rfz[k] *= ratio; pri_rfz[k] *= ratio; prr_rci[k] *= ratio; prr_rcs[k] *= ratio; prr_rcg[k] *= ratio;
            }
            sump = (float)(prs_sde[k] - prs_ihm[k] - prr_sml[k] + prs_rcs[k]);
            rate_max = -rs[k] * odts;
            if (sump < rate_max && L_qs[k]) {
                ratio = rate_max / sump;
                prs_sde[k] *= ratio; prs_ihm[k] *= ratio; prr_sml[k] *= ratio; prs_rcs[k] *= ratio;
            }
            sump = (float)(prg_gde[k] - prg_ihm[k] - prr_gml[k] + prg_rcg[k]);
            rate_max = -rg[k] * odts;
            if (sump < rate_max && L_qg[k]) {
                ratio = rate_max / sump;
                prg_gde[k] *= ratio; prg_ihm[k] *= ratio; prr_gml[k] *= ratio; prg_rcg[k] *= ratio;
            }
            pri_ihm[k] = prs_ihm[k] + prg_ihm[k];                                                       // :2100-2108
            ratio = (float)std::min(std::fabs(prr_rcg[k]), std::fabs(prg_rcg[k]));
            prr_rcg[k] = ratio * std::copysign(1.0f, (float)prr_rcg[k]);
            prg_rcg[k] = -prr_rcg[k];
            if (temp[k] > T_0) {
                ratio = (float)std::min(std::fabs(prr_rcs[k]), std::fabs(prs_rcs[k]));
                prr_rcs[k] = ratio * std::copysign(1.0f, (float)prr_rcs[k]);
                prs_rcs[k] = -prr_rcs[k];
            }
        }

        // tendencies :2116-2236
        for (int k = kts; k <= kte; ++k) {
            orho = 1.f / rho[k];
            lfus2 = lsub - lvap[k];
            qvten[k] = (float)(qvten[k] + (-pri_inu[k] - pri_ide[k] - prs_ide[k] - prs_sde[k] - prg_gde[k]) * orho);
            qcten[k] = (float)(qcten[k] + (-prr_wau[k] - pri_wfz[k] - prr_rcw[k] - prs_scw[k] - prg_scw[k] - prg_gcw[k]) * orho);
            qiten[k] = (float)(qiten[k] + (pri_inu[k] + pri_ihm[k] + pri_wfz[k] + pri_rfz[k] + pri_ide[k] - prs_iau[k] -
                                           prs_sci[k] - pri_rci[k]) * orho);
            niten[k] = (float)(niten[k] + (pni_inu[k] + pni_ihm[k] + pni_wfz[k] + pni_rfz[k] + pni_ide[k] - pni_iau[k] -
                                           pni_sci[k] - pni_rci[k]) * orho);
            xri = std::max(R1, (qi1d[k] + qiten[k] * dtsave) * rho[k]);                                 // :2145-2165
            xni = std::max(R2, (ni1d[k] + niten[k] * dtsave) * rho[k]);
            if (xri > R1) {
                lami = std::pow((am_i * cig[1] * oig1 * xni / xri), obmi);
                ilami = 1. / lami;
                xDi = (float)((bm_i + mu_i + 1.f) * ilami);
                if (xDi < 20.E-6f) {
                    lami = cie[1] / 20.E-6f;
                    xni = (float)std::min(250.E3, cig[0] * oig2 * xri / am_i * std::pow(lami, (double)bm_i));
                    niten[k] = (xni - ni1d[k] * rho[k]) * odts * orho;
                } else if (xDi > 300.E-6f) {
                    lami = cie[1] / 300.E-6f;
                    xni = (float)(cig[0] * oig2 * xri / am_i * std::pow(lami, (double)bm_i));
                    niten[k] = (xni - ni1d[k] * rho[k]) * odts * orho;
                }
            } else {
                niten[k] = -ni1d[k] * odts;
            }
            xni = std::max(0.f, (ni1d[k] + niten[k] * dtsave) * rho[k]);
            if (xni > 250.E3f) niten[k] = (250.E3f - ni1d[k] * rho[k]) * odts * orho;

            qrten[k] = (float)(qrten[k] + (prr_wau[k] + prr_rcw[k] + prr_sml[k] + prr_gml[k] + prr_rcs[k] + prr_rcg[k] -
                                           prg_rfz[k] - pri_rfz[k] - prr_rci[k]) * orho);
            nrten[k] = (float)(nrten[k] + (pnr_wau[k] + pnr_sml[k] + pnr_gml[k] -
                                           (pnr_rfz[k] + pnr_rcr[k] + pnr_rcg[k] + pnr_rcs[k] + pnr_rci[k])) * orho);
            xrr = std::max(R1, (qr1d[k] + qrten[k] * dtsave) * rho[k]);                                 // :2182-2201
            xnr = std::max(R2, (nr1d[k] + nrten[k] * dtsave) * rho[k]);
            if (xrr > R1) {
                lamr = std::pow((am_r * crg[2] * org2 * xnr / xrr), obmr);
                mvd_r[k] = (float)((3.0f + mu_r + 0.672f) / lamr);
                if (mvd_r[k] > 2.5E-3f) {
                    mvd_r[k] = 2.5E-3f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    xnr = (float)(crg[1] * org3 * xrr * std::pow(lamr, (double)bm_r) / am_r);
                    nrten[k] = (xnr - nr1d[k] * rho[k]) * odts * orho;
                } else if (mvd_r[k] < D0r * 0.75f) {
                    mvd_r[k] = D0r * 0.75f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    xnr = (float)(crg[1] * org3 * xrr * std::pow(lamr, (double)bm_r) / am_r);
                    nrten[k] = (xnr - nr1d[k] * rho[k]) * odts * orho;
                }
            } else {
                qrten[k] = -qr1d[k] * odts;
                nrten[k] = -nr1d[k] * odts;
            }
            qsten[k] = (float)(qsten[k] + (prs_iau[k] + prs_sde[k] + prs_sci[k] + prs_scw[k] + prs_rcs[k] + prs_ide[k] -
                                           prs_ihm[k] - prr_sml[k]) * orho);
            qgten[k] = (float)(qgten[k] + (prg_scw[k] + prg_rfz[k] + prg_gde[k] + prg_rcg[k] + prg_gcw[k] + prg_rci[k] +
                                           prg_rcs[k] - prg_ihm[k] - prr_gml[k]) * orho);
            if (temp[k] < T_0) {
                tten[k] = (float)(tten[k] + (lsub * ocp[k] * (pri_inu[k] + pri_ide[k] + prs_ide[k] + prs_sde[k] + prg_gde[k]) +
                                             lfus2 * ocp[k] * (pri_wfz[k] + pri_rfz[k] + prg_rfz[k] + prs_scw[k] + prg_scw[k] +
                                                               prg_gcw[k] + prg_rcs[k] + prs_rcs[k] + prr_rci[k] + prg_rcg[k])) *
                                                orho * (1 - 0));
            } else {
                tten[k] = (float)(tten[k] + (lfus * ocp[k] * (-prr_sml[k] - prr_gml[k] - prr_rcg[k] - prr_rcs[k]) +
                                             lsub * ocp[k] * (prs_sde[k] + prg_gde[k])) * orho * (1 - 0));
            }
        }

        // update :2241-2318
        for (int k = kts; k <= kte; ++k) {
            temp[k] = t1d[k] + dt * tten[k];
            otemp = 1.f / temp[k];
            tempc = temp[k] - 273.15f;
            qv[k] = std::max(1.E-10f, qv1d[k] + dt * qvten[k]);
            rho[k] = 0.622f * pres[k] / (RR2 * temp[k] * (qv[k] + 0.622f));
            rhof[k] = std::sqrt(rho_not / rho[k]);
            rhof2[k] = std::sqrt(rhof[k]);
            qvs[k] = RSLF(pres[k], temp[k]);
            ssatw[k] = qv[k] / qvs[k] - 1.f;
            if (std::fabs(ssatw[k]) < eps) ssatw[k] = 0.0f;
            diffu[k] = 2.11E-5f * std::pow((temp[k] / 273.15f), 1.94f) * (101325.f / pres[k]);
            if (tempc >= 0.0f) visco[k] = (1.718f + 0.0049f * tempc) * 1.0E-5f;
            else visco[k] = (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
            vsc2[k] = std::sqrt(rho[k] / visco[k]);
            lvap[k] = lvap0 + (2106.0f - 4218.0f) * tempc;
            tcond[k] = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
            ocp[k] = 1.f / (Cp2 * (1.f + 0.887f * qv[k]));
            lvt2[k] = lvap[k] * lvap[k] * ocp[k] * oRv * otemp * otemp;
            if ((qc1d[k] + qcten[k] * dt) > R1) { rc[k] = (qc1d[k] + qcten[k] * dt) * rho[k]; L_qc[k] = 1; }
            else { rc[k] = R1; L_qc[k] = 0; }
            if ((qi1d[k] + qiten[k] * dt) > R1) {
                ri[k] = (qi1d[k] + qiten[k] * dt) * rho[k];
                ni[k] = std::max(R2, (ni1d[k] + niten[k] * dt) * rho[k]);
                L_qi[k] = 1;
            } else { ri[k] = R1; ni[k] = R2; L_qi[k] = 0; }
            if ((qr1d[k] + qrten[k] * dt) > R1) {
                rr[k] = (qr1d[k] + qrten[k] * dt) * rho[k];
                nr[k] = std::max(R2, (nr1d[k] + nrten[k] * dt) * rho[k]);
                L_qr[k] = 1;
                lamr = std::pow((am_r * crg[2] * org2 * nr[k] / rr[k]), obmr);
                mvd_r[k] = (float)((3.0f + mu_r + 0.672f) / lamr);
                if (mvd_r[k] > 2.5E-3f) {
                    mvd_r[k] = 2.5E-3f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    nr[k] = (float)(crg[1] * org3 * rr[k] * std::pow(lamr, (double)bm_r) / am_r);
                } else if (mvd_r[k] < D0r * 0.75f) {
                    mvd_r[k] = D0r * 0.75f;
                    lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                    nr[k] = (float)(crg[1] * org3 * rr[k] * std::pow(lamr, (double)bm_r) / am_r);
                }
            } else { rr[k] = R1; nr[k] = R2; L_qr[k] = 0; }
            if ((qs1d[k] + qsten[k] * dt) > R1) { rs[k] = (qs1d[k] + qsten[k] * dt) * rho[k]; L_qs[k] = 1; }
            else { rs[k] = R1; L_qs[k] = 0; }
            if ((qg1d[k] + qgten[k] * dt) > R1) { rg[k] = (qg1d[k] + qgten[k] * dt) * rho[k]; L_qg[k] = 1; }
            else { rg[k] = R1; L_qg[k] = 0; }
        }

        // snow moments again :2325-2374 (smoc, smod)
        for (int k = kts; k <= kte; ++k) {
            if (!L_qs[k]) continue;
            tc0 = std::min(-0.1f, temp[k] - 273.15f);
            smob[k] = rs[k] * oams;
            if (bm_s > (2.0f - 1.e-3f) && bm_s < (2.0f + 1.e-3f)) {
                smo2[k] = smob[k];
            } else {
                loga_ = sa_poly(sa, tc0, bm_s);
                a_ = std::pow(10.0f, loga_);
                b_ = sa_poly(sb, tc0, bm_s);
                smo2[k] = std::pow((smob[k] / a_), (1.f / b_));
            }
            loga_ = sa_poly(sa, tc0, cse[0]); a_ = std::pow(10.0f, loga_); b_ = sa_poly(sb, tc0, cse[0]);
            smoc[k] = a_ * std::pow(smo2[k], b_);
            loga_ = sa_poly(sa, tc0, cse[13]); a_ = std::pow(10.0f, loga_); b_ = sa_poly(sb, tc0, cse[13]);
            smod[k] = a_ * std::pow(smo2[k], b_);
        }
        graupel_rain_slopes();                                                                          // :2379-2408

        // condensation / evaporation of cloud water :2414-2440
        for (int k = kts; k <= kte; ++k) {
            if ((ssatw[k] > eps) || (ssatw[k] < -eps && L_qc[k])) {
                clap = (qv[k] - qvs[k]) / (1.f + lvt2[k] * qvs[k]);
                for (int n = 1; n <= 3; ++n) {
                    fcd = qvs[k] * std::exp(lvt2[k] * clap) - qv[k] + clap;
                    dfcd = qvs[k] * lvt2[k] * std::exp(lvt2[k] * clap) + 1.f;
                    clap = clap - fcd / dfcd;
                }
                xrc = rc[k] + clap;
                if (xrc > 0.0f) prw_vcd[k] = clap * odt;
                else prw_vcd[k] = -rc[k] / rho[k] * odts;
                qcten[k] = (float)(qcten[k] + prw_vcd[k]);
                qvten[k] = (float)(qvten[k] - prw_vcd[k]);
                tten[k] = (float)(tten[k] + lvap[k] * ocp[k] * prw_vcd[k] * (1 - 0));
                rc[k] = std::max(R1, (qc1d[k] + dt * qcten[k]) * rho[k]);
                qv[k] = std::max(1.E-10f, qv1d[k] + dt * qvten[k]);
                temp[k] = t1d[k] + dt * tten[k];
                rho[k] = 0.622f * pres[k] / (RR2 * temp[k] * (qv[k] + 0.622f));
                qvs[k] = RSLF(pres[k], temp[k]);
                ssatw[k] = qv[k] / qvs[k] - 1.f;
            }
        }

        // rain evaporation :2446-2505
        for (int k = kts; k <= kte; ++k) {
            if ((ssatw[k] < -eps) && L_qr[k] && (!(prw_vcd[k] > 0.))) {
                tempc = temp[k] - 273.15f;
                otemp = 1.f / temp[k];
                rhof[k] = std::sqrt(rho_not / rho[k]);
                rhof2[k] = std::sqrt(rhof[k]);
                diffu[k] = 2.11E-5f * std::pow((temp[k] / 273.15f), 1.94f) * (101325.f / pres[k]);
                if (tempc >= 0.0f) visco[k] = (1.718f + 0.0049f * tempc) * 1.0E-5f;
                else visco[k] = (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
                vsc2[k] = std::sqrt(rho[k] / visco[k]);
                lvap[k] = lvap0 + (2106.0f - 4218.0f) * tempc;
                tcond[k] = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
                ocp[k] = 1.f / (Cp2 * (1.f + 0.887f * qv[k]));
                rvs = rho[k] * qvs[k];
                rvs_p = rvs * otemp * (lvap[k] * otemp * oRv - 1.f);
                rvs_pp = rvs * (otemp * (lvap[k] * otemp * oRv - 1.f) * otemp * (lvap[k] * otemp * oRv - 1.f) +
                                (-2.f * lvap[k] * otemp * otemp * otemp * oRv) + otemp * otemp);
                gamsc = lvap[k] * diffu[k] / tcond[k] * rvs_p;
                alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
                alphsc = std::max(1.E-9f, alphsc);
                xsat = std::min(-1.E-9f, ssatw[k]);
                t1_evap = 2.f * PI2 * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat -
                                       5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);
                lamr = 1. / ilamr[k];
                if (qv[k] / qvs[k] < 0.95f && rr[k] / rho[k] <= 1.E-8f) {
                    prv_rev[k] = rr[k] / rho[k] * odts;
                } else {
                    prv_rev[k] = t1_evap * diffu[k] * (-ssatw[k]) * N0_r[k] * rvs *
                                 (t1_qr_ev * std::pow(ilamr[k], (double)cre[9]) +
                                  t2_qr_ev * vsc2[k] * rhof2[k] * (std::pow((lamr + 0.5f * fv_r), (double)(-cre[10]))));
                    rate_max = std::min((rr[k] / rho[k] * odts), (qvs[k] - qv[k]) * odts);
                    prv_rev[k] = std::min((double)rate_max, prv_rev[k] / rho[k]);
                }
                pnr_rev[k] = std::min((double)(nr[k] * 0.99f / rho[k] * odts), prv_rev[k] * nr[k] / rr[k]);
                qrten[k] = (float)(qrten[k] - prv_rev[k]);
                qvten[k] = (float)(qvten[k] + prv_rev[k]);
                nrten[k] = (float)(nrten[k] - pnr_rev[k]);
                tten[k] = (float)(tten[k] - lvap[k] * ocp[k] * prv_rev[k] * (1 - 0));
                rr[k] = std::max(R1, (qr1d[k] + dt * qrten[k]) * rho[k]);
                qv[k] = std::max(1.E-10f, qv1d[k] + dt * qvten[k]);
                nr[k] = std::max(R2, (nr1d[k] + dt * nrten[k]) * rho[k]);
                temp[k] = t1d[k] + dt * tten[k];
                rho[k] = 0.622f * pres[k] / (RR2 * temp[k] * (qv[k] + 0.622f));
            }
        }

        // fall speeds :2515-2650
        nstep = 0;
        for (int n = 0; n < 4; ++n) { onstep[n] = 1.0f; ksed1[n] = 1; }      // ksed1 is a 1-based level (kts = 1)
        for (int k = kte + 1; k >= kts; --k) { vtrk[k] = 0.f; vtnrk[k] = 0.f; vtik[k] = 0.f; vtnik[k] = 0.f; vtsk[k] = 0.f; vtgk[k] = 0.f; }
        for (int k = kte; k >= kts; --k) {
            vtr = 0.f;
            rhof[k] = std::sqrt(rho_not / rho[k]);
            if (rr[k] > R1) {
                lamr = std::pow((am_r * crg[2] * org2 * nr[k] / rr[k]), obmr);
                vtr = (float)(rhof[k] * av_r * crg[5] * org3 * std::pow(lamr, (double)cre[2]) *
                              (std::pow((lamr + fv_r), (double)(-cre[5]))));
                vtrk[k] = vtr;
                vtr = (float)(rhof[k] * av_r * crg[6] / crg[11] * std::pow(lamr, (double)cre[11]) *
                              (std::pow((lamr + fv_r), (double)(-cre[6]))));
                vtnrk[k] = vtr;
            } else {
                vtrk[k] = vtrk[k + 1];
                vtnrk[k] = vtnrk[k + 1];
            }
            if (std::max(vtrk[k], vtnrk[k]) > 1.E-3f) {
                ksed1[0] = std::max(ksed1[0], k + 1);
                delta_tp = dzq[k] / (std::max(vtrk[k], vtnrk[k]));
                nstep = std::max(nstep, (int)(dt / delta_tp + 1.f));
            }
        }
        if (ksed1[0] == kte + 1) ksed1[0] = kte;
        if (nstep > 0) onstep[0] = 1.f / (float)nstep;

        nstep = 0;
        for (int k = kte; k >= kts; --k) {
            vti = 0.f;
            if (ri[k] > R1) {
                lami = std::pow((am_i * cig[1] * oig1 * ni[k] / ri[k]), obmi);
                ilami = 1. / lami;
                vti = (float)(rhof[k] * av_i * cig[2] * oig2 * std::pow(ilami, (double)bv_i));
                vtik[k] = vti;
                vti = (float)(rhof[k] * av_i * cig[5] / cig[6] * std::pow(ilami, (double)bv_i));
                vtnik[k] = vti;
            } else {
                vtik[k] = vtik[k + 1];
                vtnik[k] = vtnik[k + 1];
            }
            if (vtik[k] > 1.E-3f) {
                ksed1[1] = std::max(ksed1[1], k + 1);
                delta_tp = dzq[k] / vtik[k];
                nstep = std::max(nstep, (int)(dt / delta_tp + 1.f));
            }
        }
        if (ksed1[1] == kte + 1) ksed1[1] = kte;
        if (nstep > 0) onstep[1] = 1.f / (float)nstep;

        nstep = 0;
        for (int k = kte; k >= kts; --k) {
            vts = 0.f;
            if (rs[k] > R1) {
                xDs = smoc[k] / smob[k];
                Mrat = 1.f / xDs;
                ils1 = 1.f / (Mrat * Lam0 + fv_s);
                ils2 = 1.f / (Mrat * Lam1 + fv_s);
                t1_vts = Kap0 * csg[3] * std::pow(ils1, cse[3]);
                t2_vts = Kap1 * std::pow(Mrat, mu_s) * csg[9] * std::pow(ils2, cse[9]);
                ils1 = 1.f / (Mrat * Lam0);
                ils2 = 1.f / (Mrat * Lam1);
                t3_vts = Kap0 * csg[0] * std::pow(ils1, cse[0]);
                t4_vts = Kap1 * std::pow(Mrat, mu_s) * csg[6] * std::pow(ils2, cse[6]);
                vts = rhof[k] * av_s * (t1_vts + t2_vts) / (t3_vts + t4_vts);
                if (temp[k] > T_0) vtsk[k] = std::max(vts * vts_boost[k], vtrk[k]);
                else vtsk[k] = vts * vts_boost[k];
            } else {
                vtsk[k] = vtsk[k + 1];
            }
            if (vtsk[k] > 1.E-3f) {
                ksed1[2] = std::max(ksed1[2], k + 1);
                delta_tp = dzq[k] / vtsk[k];
                nstep = std::max(nstep, (int)(dt / delta_tp + 1.f));
            }
        }
        if (ksed1[2] == kte + 1) ksed1[2] = kte;
        if (nstep > 0) onstep[2] = 1.f / (float)nstep;

        nstep = 0;
        for (int k = kte; k >= kts; --k) {
            vtg = 0.f;
            if (rg[k] > R1) {
                vtg = (float)(rhof[k] * av_g * cgg[5] * ogg3 * std::pow(ilamg[k], (double)bv_g));
                if (temp[k] > T_0) vtgk[k] = std::max(vtg, vtrk[k]);
                else vtgk[k] = vtg;
            } else {
                vtgk[k] = vtgk[k + 1];
            }
            if (vtgk[k] > 1.E-3f) {
                ksed1[3] = std::max(ksed1[3], k + 1);
                delta_tp = dzq[k] / vtgk[k];
                nstep = std::max(nstep, (int)(dt / delta_tp + 1.f));
            }
        }
        if (ksed1[3] == kte + 1) ksed1[3] = kte;
        if (nstep > 0) onstep[3] = 1.f / (float)nstep;

        // sedimentation :2660-2770.  Loop `do k = ksed1, kts, -1` with 1-based ksed1 -> 0-based ksed1-1.
        // sed_x(k+1) for k = kte is never read because ksed1 <= kte-1 (1-based).
        nstep = nint_f(1.f / onstep[0]);
        for (int n = 1; n <= nstep; ++n) {
            for (int k = kte; k >= kts; --k) { sed_r[k] = vtrk[k] * rr[k]; sed_n[k] = vtnrk[k] * nr[k]; }
            int k = kte;
            odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
            qrten[k] = qrten[k] - sed_r[k] * odzq * onstep[0] * orho;
            nrten[k] = nrten[k] - sed_n[k] * odzq * onstep[0] * orho;
            rr[k] = std::max(R1, rr[k] - sed_r[k] * odzq * dt * onstep[0]);
            nr[k] = std::max(R2, nr[k] - sed_n[k] * odzq * dt * onstep[0]);
            for (k = ksed1[0] - 1; k >= kts; --k) {
                odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
                qrten[k] = qrten[k] + (sed_r[k + 1] - sed_r[k]) * odzq * onstep[0] * orho;
                nrten[k] = nrten[k] + (sed_n[k + 1] - sed_n[k]) * odzq * onstep[0] * orho;
                rr[k] = std::max(R1, rr[k] + (sed_r[k + 1] - sed_r[k]) * odzq * dt * onstep[0]);
                nr[k] = std::max(R2, nr[k] + (sed_n[k + 1] - sed_n[k]) * odzq * dt * onstep[0]);
            }
            if (rr[kts] > R1 * 10.f) pptrain = pptrain + sed_r[kts] * dt * onstep[0];
        }
        nstep = nint_f(1.f / onstep[1]);
        for (int n = 1; n <= nstep; ++n) {
            for (int k = kte; k >= kts; --k) { sed_i[k] = vtik[k] * ri[k]; sed_n[k] = vtnik[k] * ni[k]; }
            int k = kte;
            odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
            qiten[k] = qiten[k] - sed_i[k] * odzq * onstep[1] * orho;
            niten[k] = niten[k] - sed_n[k] * odzq * onstep[1] * orho;
            ri[k] = std::max(R1, ri[k] - sed_i[k] * odzq * dt * onstep[1]);
            ni[k] = std::max(R2, ni[k] - sed_n[k] * odzq * dt * onstep[1]);
            for (k = ksed1[1] - 1; k >= kts; --k) {
                odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
                qiten[k] = qiten[k] + (sed_i[k + 1] - sed_i[k]) * odzq * onstep[1] * orho;
                niten[k] = niten[k] + (sed_n[k + 1] - sed_n[k]) * odzq * onstep[1] * orho;
                ri[k] = std::max(R1, ri[k] + (sed_i[k + 1] - sed_i[k]) * odzq * dt * onstep[1]);
                ni[k] = std::max(R2, ni[k] + (sed_n[k + 1] - sed_n[k]) * odzq * dt * onstep[1]);
            }
            if (ri[kts] > R1 * 10.f) pptice = pptice + sed_i[kts] * dt * onstep[1];
        }
        nstep = nint_f(1.f / onstep[2]);
        for (int n = 1; n <= nstep; ++n) {
            for (int k = kte; k >= kts; --k) sed_s[k] = vtsk[k] * rs[k];
            int k = kte;
            odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
            qsten[k] = qsten[k] - sed_s[k] * odzq * onstep[2] * orho;
            rs[k] = std::max(R1, rs[k] - sed_s[k] * odzq * dt * onstep[2]);
            for (k = ksed1[2] - 1; k >= kts; --k) {
                odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
                qsten[k] = qsten[k] + (sed_s[k + 1] - sed_s[k]) * odzq * onstep[2] * orho;
                rs[k] = std::max(R1, rs[k] + (sed_s[k + 1] - sed_s[k]) * odzq * dt * onstep[2]);
            }
            if (rs[kts] > R1 * 10.f) pptsnow = pptsnow + sed_s[kts] * dt * onstep[2];
        }
        nstep = nint_f(1.f / onstep[3]);
        for (int n = 1; n <= nstep; ++n) {
            for (int k = kte; k >= kts; --k) sed_g[k] = vtgk[k] * rg[k];
            int k = kte;
            odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
            qgten[k] = qgten[k] - sed_g[k] * odzq * onstep[3] * orho;
            rg[k] = std::max(R1, rg[k] - sed_g[k] * odzq * dt * onstep[3]);
            for (k = ksed1[3] - 1; k >= kts; --k) {
                odzq = 1.f / dzq[k]; orho = 1.f / rho[k];
                qgten[k] = qgten[k] + (sed_g[k + 1] - sed_g[k]) * odzq * onstep[3] * orho;
                rg[k] = std::max(R1, rg[k] + (sed_g[k + 1] - sed_g[k]) * odzq * dt * onstep[3]);
            }
            if (rg[kts] > R1 * 10.f) pptgraul = pptgraul + sed_g[kts] * dt * onstep[3];
        }

        // instant melt / freeze :2777-2794
        for (int k = kts; k <= kte; ++k) {
            xri = std::max(0.0f, qi1d[k] + qiten[k] * dt);
            if ((temp[k] > T_0) && (xri > 0.0f)) {
                qcten[k] = qcten[k] + xri * odt;
                qiten[k] = qiten[k] - xri * odt;
                niten[k] = -ni1d[k] * odt;
                tten[k] = tten[k] - lfus * ocp[k] * xri * odt * (1 - 0);
            }
            xrc = std::max(0.0f, qc1d[k] + qcten[k] * dt);
            if ((temp[k] < HGFR) && (xrc > 0.0f)) {
                lfus2 = lsub - lvap[k];
                qiten[k] = qiten[k] + xrc * odt;
                niten[k] = niten[k] + xrc / xm0i * odt;
                qcten[k] = qcten[k] - xrc * odt;
                tten[k] = tten[k] + lfus2 * ocp[k] * xrc * odt * (1 - 0);
            }
        }

        // apply :2800-2842
        for (int k = kts; k <= kte; ++k) {
            t1d[k] = t1d[k] + tten[k] * dt;
            qv1d[k] = std::max(1.E-10f, qv1d[k] + qvten[k] * dt);
            qc1d[k] = qc1d[k] + qcten[k] * dt;
            if (qc1d[k] <= R1) qc1d[k] = 0.0f;
            qi1d[k] = qi1d[k] + qiten[k] * dt;
            ni1d[k] = std::max(R2 / rho[k], ni1d[k] + niten[k] * dt);
            if (qi1d[k] <= R1) { qi1d[k] = 0.0f; ni1d[k] = 0.0f; }
            else {
                lami = std::pow((am_i * cig[1] * oig1 * ni1d[k] / qi1d[k]), obmi);
                ilami = 1. / lami;
                xDi = (float)((bm_i + mu_i + 1.f) * ilami);
                if (xDi < 20.E-6f) lami = cie[1] / 20.E-6f;
                else if (xDi > 300.E-6f) lami = cie[1] / 300.E-6f;
                ni1d[k] = (float)std::min(cig[0] * oig2 * qi1d[k] / am_i * std::pow(lami, (double)bm_i), 250.E3 / rho[k]);
            }
            qr1d[k] = qr1d[k] + qrten[k] * dt;
            nr1d[k] = std::max(R2 / rho[k], nr1d[k] + nrten[k] * dt);
            if (qr1d[k] <= R1) { qr1d[k] = 0.0f; nr1d[k] = 0.0f; }
            else {
                lamr = std::pow((am_r * crg[2] * org2 * nr1d[k] / qr1d[k]), obmr);
                mvd_r[k] = (float)((3.0f + mu_r + 0.672f) / lamr);
                if (mvd_r[k] > 2.5E-3f) mvd_r[k] = 2.5E-3f;
                else if (mvd_r[k] < D0r * 0.75f) mvd_r[k] = D0r * 0.75f;
                lamr = (3.0f + mu_r + 0.672f) / mvd_r[k];
                nr1d[k] = (float)(crg[1] * org3 * qr1d[k] * std::pow(lamr, (double)bm_r) / am_r);
            }
            qs1d[k] = qs1d[k] + qsten[k] * dt;
            if (qs1d[k] <= R1) qs1d[k] = 0.0f;
            qg1d[k] = qg1d[k] + qgten[k] * dt;
            if (qg1d[k] <= R1) qg1d[k] = 0.0f;
        }
        return false;
    }

    // ------------------------------------------------------------------ mp_gt_driver :772-1044
    // 3-D arrays in the reference layout (i fastest, then k, then j); tile bounds 0-based inclusive;
    // ide0 / jde0 = 0-based index of the last domain cell relative to this memory block
    // (i_end = MIN(ite, ide-1), :821-822).  Optional outputs may be null.
    void driver(float* qv, float* qc, float* qr, float* qi, float* qs, float* qg, float* ni, float* nr, float* th,
                const float* pii, const float* p, const float* dz, float dt_in, float* RAINNC, const float* RAINNCV,
                float* SNOWNC, float* GRAUPELNC, float* SR, int nx, int nz, int ny, int its, int ite, int jts, int jte,
                int kts, int kte, int ide0, int jde0, long long* n_no_micro = nullptr) const {
        const int i_end = std::min(ite, ide0 - 1), j_end = std::min(jte, jde0 - 1);
        const int nk = kte - kts + 1;
        long long nm = 0;
#pragma omp parallel for schedule(guided) reduction(+ : nm)
        for (int j = jts; j <= j_end; ++j) {
            std::vector<float> c((size_t)nk * 11);
            float *qv1d = c.data(), *qc1d = qv1d + nk, *qi1d = qc1d + nk, *qr1d = qi1d + nk, *qs1d = qr1d + nk,
                  *qg1d = qs1d + nk, *ni1d = qg1d + nk, *nr1d = ni1d + nk, *t1d = nr1d + nk, *p1d = t1d + nk,
                  *dz1d = p1d + nk;
            for (int i = its; i <= i_end; ++i) {
                float pptrain = 0.f, pptsnow = 0.f, pptgraul = 0.f, pptice = 0.f;
                for (int k = kts; k <= kte; ++k) {
                    const size_t a = (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)j);
                    const int kk = k - kts;
                    t1d[kk] = th[a] * pii[a]; p1d[kk] = p[a]; dz1d[kk] = dz[a]; qv1d[kk] = qv[a]; qc1d[kk] = qc[a];
                    qi1d[kk] = qi[a]; qr1d[kk] = qr[a]; qs1d[kk] = qs[a]; qg1d[kk] = qg[a]; ni1d[kk] = ni[a];
                    nr1d[kk] = nr[a];
                }
                if (column(qv1d, qc1d, qi1d, qr1d, qs1d, qg1d, ni1d, nr1d, t1d, p1d, dz1d, pptrain, pptsnow, pptgraul,
                           pptice, nk, dt_in))
                    ++nm;
                const size_t a2 = (size_t)i + (size_t)nx * (size_t)j;
                RAINNC[a2] = RAINNC[a2] + pptrain + pptsnow + pptgraul + pptice;                         // :908
                if (SNOWNC) SNOWNC[a2] = SNOWNC[a2] + pptsnow + pptice;
                if (GRAUPELNC) GRAUPELNC[a2] = GRAUPELNC[a2] + pptgraul;
                if (SR) SR[a2] = (pptsnow + pptgraul + pptice) / ((RAINNCV ? RAINNCV[a2] : 0.f) + 1.e-12f);   // :915
                for (int k = kts; k <= kte; ++k) {
                    const size_t a = (size_t)i + (size_t)nx * ((size_t)k + (size_t)nz * (size_t)j);
                    const int kk = k - kts;
                    qv[a] = qv1d[kk]; qc[a] = qc1d[kk]; qi[a] = qi1d[kk]; qr[a] = qr1d[kk]; qs[a] = qs1d[kk];
                    qg[a] = qg1d[kk]; ni[a] = ni1d[kk]; nr[a] = nr1d[kk];
                    th[a] = t1d[kk] / pii[a];
                    if (qv1d[kk] < 1.E-7f) {                                                             // :997-1010
                        if (k < kte - 2 && k > kts + 1) {
                            const size_t am = a - (size_t)nx, ap = a + (size_t)nx;
                            qv[a] = 0.5f * (qv[am] + qv[ap]);
                            if (qv1d[kk] < 1.E-7f) qv[a] = 1.E-7f;
                        } else {
                            qv[a] = 1.E-7f;
                        }
                    }
                }
            }
        }
        if (n_no_micro) *n_no_micro = nm;
    }
};

}  // namespace icar_oracle
