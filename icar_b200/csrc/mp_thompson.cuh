// Thompson et al. (2008) bulk microphysics (options%physics%microphysics = kMP_THOMPSON = 1) on sm_100a.
// Reference: src/physics/mp_thompson.f90 (thompson_init :342-766, mp_gt_driver :772-1044,
// mp_thompson :1057-2844, table builders :2853-3578).
#pragma once
#include "common.cuh"
#include "mp_simple.cuh"      // MpTiles

namespace icar {

constexpr int TH_NTABLES = 29;
// table ids, in the allocation order of thompson_init (mp_thompson.f90:353-390)
enum ThTableId {
    T_tcg_racg = 0, T_tmr_racg, T_tcr_gacr, T_tmg_gacr, T_tnr_racg, T_tnr_gacr,
    T_tcs_racs1, T_tmr_racs1, T_tcs_racs2, T_tmr_racs2, T_tcr_sacr1, T_tms_sacr1,
    T_tcr_sacr2, T_tms_sacr2, T_tnr_racs1, T_tnr_racs2, T_tnr_sacr1, T_tnr_sacr2,
    T_tpi_qcfz, T_tni_qcfz, T_tpi_qrfz, T_tpg_qrfz, T_tni_qrfz, T_tnr_qrfz,
    T_tps_iaus, T_tni_iaus, T_tpi_ide, T_Efrw, T_Efsw
};

// lookup-table dimensions (mp_thompson.f90:179-193)
constexpr int nbins = 100;
constexpr int ntb_c = 37, ntb_i = 64, ntb_r = 37, ntb_s = 28, ntb_g = 28, ntb_g1 = 28, ntb_r1 = 37, ntb_i1 = 55,
              ntb_t = 9;

// mp_options_type (src/objects/opt_types.f90:31-45) -- same layout as icar_mp_options in the C ABI
struct ThOptions {
    float Nt_c, TNO, am_s, rho_g, av_s, bv_s, fv_s, av_i, av_g, bv_g, Ef_si, Ef_rs, Ef_rg, Ef_ri, C_cubes, C_sqrd,
        mu_r, t_adjust;
    int Ef_rw_l, Ef_sw_l;
};

// Everything thompson_init derives from the options (module variables of mp_thompson.f90:305-330);
// computed once on the host (th_make_constants) and passed to the kernels by value.
struct ThConst {
    ThOptions o;
    float am_g, am_r, am_i, mu_c, Sc3, D0i, xm0s, xm0g, lfus, olfus, oRv, rho_not;
    float cce[3], ccg[3], ocg1, ocg2;
    float cie[7], cig[7], oig1, oig2, obmi;
    float cre[13], crg[13], ore1, org1, org2, org3, obmr;
    float cse[18], csg[18], oams, obms, ocms;
    float cge[12], cgg[12], oge1, ogg1, ogg2, ogg3, oamg, obmg, ocmg;
    float t1_qr_qc, t1_qr_qi, t2_qr_qi, t1_qg_qc, t1_qs_qc, t1_qs_qi, t1_qr_ev, t2_qr_ev, t1_qs_sd, t2_qs_sd, t1_qg_sd,
        t2_qg_sd, t1_qs_me, t2_qs_me, t1_qg_me, t2_qg_me;
    int nic2, nii2, nii3, nir2, nir3, nis2, nig2, nig3;
    double Dr1, Drn, Ds1, Dsn;          // Dr(1), Dr(nbr), Ds(1), Ds(nbs): the efficiency-table index needs them
};

struct ThTables { const double* t[TH_NTABLES]; };

// element count of table `id`
size_t th_table_elems(int id);

// thompson_init's scalar part (:397-578) and bin grids (:581-634): host code.  bins = 10 x 100 doubles:
// Dc, dtc, Di, dti, Dr, dtr, Ds, dts, Dg, dtg.
void th_make_constants(const ThOptions& o, ThConst* c, double* bins);

// The table builders (qr_acr_qg, qr_acr_qs, freezeH2O, qi_aut_qs, table_Efrw, table_Efsw) on the device.
// d_bins = the 1000 doubles above (device), d_aux = scratch of 800 doubles (device), tables = TH_NTABLES
// device buffers of th_table_elems(id) doubles.  Returns the number of kernels queued.
int th_build_tables(const ThConst& c, const double* d_bins, double* d_aux, double* const* tables, cudaStream_t st);

// mp_gt_driver (:772-1044) on up to four 0-based inclusive tiles; 2-D arrays have row pitch ld2 and may be
// null (the OPTIONAL dummies / unused RAINNCV).  diag[0] counts columns that took the no_micro exit.
int th_driver_launch(const Geom& g, const ThConst& c, const ThTables& tb, float* qv, float* qc, float* qr, float* qi,
                     float* qs, float* qg, float* ni, float* nr, float* th, const float* pii, const float* p,
                     const float* dz, float dt, float* rainnc, const float* rainncv, float* snownc, float* snowncv,
                     float* graupelnc, float* graupelncv, float* sr, int ld2, const MpTiles& tiles, int kts, int kte,
                     unsigned long long* diag, cudaStream_t st, double* acc_precip = nullptr, double* acc_snow = nullptr,
                     double* acc_graupel = nullptr);
// (acc_* non-null: resident mode -- the call's precipitation / snowfall / graupel are added straight to the
//  REAL*8 accumulators of domain_t, mp_driver.f90:587-595; rainnc / snownc / graupelnc are then null)

bool th_driver_preload();     // see preload_kernel (common.cuh)

}  // namespace icar
