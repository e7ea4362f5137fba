"""The Thompson oracle (oracle/mp_thompson_oracle.hpp) on the CPU: frozen vectors, an independent
NumPy evaluation of lookup-table entries, and physical invariants of the column routine.

Parity pinned elsewhere (tests/test_reference_source.py: mp_thompson.f90 executed by oracle/f90run); the reference has no expected values for mp_thompson and cannot be compiled here;
these tests pin the restatement against regressions and against a second, independent evaluation
of the table integrals (mp_thompson.f90:2917-2975, :3093-3220, :3323-3381, :3425-3454)."""
import math
import os

import numpy as np
import pytest
from scipy import special

from icar_b200 import ideal

PI2 = np.float32(3.1415926536)


@pytest.fixture(scope="module")
def thompson(oracle):
    """Small tables only (the two collision tables are filled slab by slab where a test needs them)."""
    return oracle.Thompson(big_tables=False)


@pytest.fixture(scope="module")
def thompson_full(oracle):
    return oracle.Thompson(big_tables=True)


def test_constants_are_gamma_functions(thompson):
    c = thompson.constants()
    mu_c, Sc3, D0i, xm0s, xm0g, am_r, am_i, am_g = c[:8]
    assert mu_c == 12.0                                    # MIN(15, 1000e6/Nt_c + 2), mp_thompson.f90:424
    assert abs(am_r - np.pi * 1000.0 / 6.0) < 1e-3
    assert abs(D0i - (1e-12 / am_i) ** (1 / 3)) / D0i < 1e-6
    assert abs(xm0g - am_g * 250e-6 ** 3) / xm0g < 1e-6
    # exponent lists and WGAMMA = exp(gammln): cce/ccg (3), cie/cig (7), cre/crg (13), cse/csg (18), cge/cgg (12)
    o = 8
    for n in (3, 7, 13, 18, 12):
        e, g = c[o:o + n].astype(np.float64), c[o + n:o + 2 * n].astype(np.float64)
        np.testing.assert_allclose(g, special.gamma(e), rtol=2e-6)
        o += 2 * n
    assert o + 16 == c.size


def test_bins_are_geometric(thompson):
    b = thompson.bins()
    np.testing.assert_allclose(np.diff(b["Dc"]), 1e-6, rtol=1e-9)         # :581-586
    for k, lo, hi in (("Dr", 50e-6, 5e-3), ("Ds", 200e-6, 2e-2), ("Dg", 250e-6, 5e-2)):
        r = b[k][1:] / b[k][:-1]
        np.testing.assert_allclose(r, r[0], rtol=1e-9)                     # log-spaced, :600-634
        assert lo < b[k][0] < b[k][-1] < hi


def test_small_tables_are_sane(thompson):
    for n in ("t_Efrw", "t_Efsw"):
        t = thompson.table(n)
        assert t.min() == 0.0 and abs(t.max() - 0.95) < 1e-7              # MAX(0, MIN(Ef, 0.95)), :3514
    ide = thompson.table("tpi_ide")
    assert ide.min() >= 0.0 and ide.max() <= 1.0
    # Bigg freezing: colder (larger k) freezes at least as much, and never more than is there
    for n in ("tpi_qcfz", "tni_qcfz"):
        t = thompson.table(n)                                              # shape (45, 37)
        assert (np.diff(t, axis=0) >= -1e-12 * np.abs(t[1:])).all()
    # rain: the bin loop EXITs once the frozen mass reaches r_r (:3352), so at -45 C everything freezes
    frozen = thompson.table("tpi_qrfz") + thompson.table("tpg_qrfz")       # (45, 37 N0, 37 r_r)
    assert np.isfinite(frozen).all() and (frozen >= 0).all()
    r_r = np.array([float(f"{m % 9 + 1}e{m // 9 - 6}") for m in range(37)])
    assert (frozen[44] >= 0.9 * r_r[None, :]).mean() > 0.5 and (frozen <= 1.5 * r_r[None, None, :]).all()


def _rain_dsd(th, k, m):
    """N_r(n2) for table cell (N0r_exp(k), r_r(m)), float64 everywhere (:2921-2926)."""
    c = th.constants().astype(np.float64)
    am_r = c[5]
    cre, crg = c[8 + 6 + 14:8 + 6 + 14 + 13], c[8 + 6 + 14 + 13:8 + 6 + 14 + 26]
    N0r = float(10.0 ** (k // 9 + 6) * (k % 9 + 1))
    rr = float(np.float32(f"{m % 9 + 1}e{m // 9 - 6}"))
    lam_exp = (N0r * am_r * crg[0] / rr) ** (1 / cre[0])
    lamr = lam_exp * (crg[2] / crg[1] / crg[0]) ** (1 / 3)
    N0 = N0r / (crg[1] * lam_exp) * lamr ** cre[1]
    b = th.bins()
    dtr = _dt_of(b["Dr"], 50e-6, 5e-3)
    return N0 * np.exp(-lamr * b["Dr"]) * dtr, rr


def _dt_of(D, lo, hi):
    n = D.size
    lo = float(np.float32(lo))             # D0r, D0s, D0g are REAL parameters (:173-175)
    x = np.exp(np.arange(n + 1) / n * np.log(hi / lo) + np.log(lo))
    np.testing.assert_allclose(np.sqrt(x[:-1] * x[1:]), D, rtol=1e-12)
    return np.diff(x)


@pytest.mark.parametrize("k,m,i,j", [(0, 0, 0, 0), (17, 20, 13, 9), (36, 36, 27, 27), (30, 5, 2, 25)])
def test_qr_acr_qg_against_numpy(thompson, k, m, i, j):
    """One entry of the six rain-graupel tables evaluated with whole-array NumPy in float64."""
    th = thompson
    th.big_slab(0, m * 37 + k)
    c = th.constants().astype(np.float64)
    am_r, am_g = c[5], c[7]
    o = 8 + 6 + 14 + 26 + 36
    cge, cgg = c[o:o + 12], c[o + 12:o + 24]
    b = th.bins()
    N_r, rr = _rain_dsd(th, k, m)
    N0g = float(10.0 ** (i // 9 + 4) * (i % 9 + 1))
    rg = float(np.float32(f"{j % 9 + 1}e{j // 9 - 5}"))
    lam_exp = (N0g * am_g * cgg[0] / rg) ** (1 / cge[0])
    lamg = lam_exp * (cgg[2] / cgg[1] / cgg[0]) ** (1 / 3)
    N0_g = N0g / (cgg[1] * lam_exp) * lamg ** cge[1]
    Dr, Dg = b["Dr"], b["Dg"]
    N_g = N0_g * np.exp(-lamg * Dg) * _dt_of(Dg, 250e-6, 5e-2)
    f = lambda x: np.float64(np.float32(x))
    vr = f(-0.1021) + f(4.932E3) * Dr - f(0.9551E6) * Dr ** 2 + f(0.07934E9) * Dr ** 3 - f(0.002362E12) * Dr ** 4
    vg = 442.0 * Dg ** f(0.89)
    dvg = np.maximum(vr[:, None] - vg[None, :], 0)
    dvr = np.maximum(vg[None, :] - vr[:, None], 0)
    kern = f(PI2 * np.float32(.25) * np.float32(0.75)) * (Dg[None, :] + Dr[:, None]) ** 2 * N_g[None, :] * N_r[:, None]
    massr, massg = am_r * Dr ** 3, am_g * Dg ** 3
    want = dict(tcg_racg=(kern * dvg * massg[None, :]).sum(),
                tmr_racg=min((kern * dvg * massr[:, None]).sum(), rr),
                tcr_gacr=(kern * dvr * massr[:, None]).sum(),
                tmg_gacr=(kern * dvr * massg[None, :]).sum(),
                tnr_racg=(kern * dvg).sum(), tnr_gacr=(kern * dvr).sum())
    for name, w in want.items():
        got = th.table(name)[m, k, j, i]
        assert abs(got - w) <= 2e-5 * abs(w) + 1e-300, (name, got, w)      # float-rounded am_*, lam exponents


def test_golden_table_samples(thompson_full, golden_dir):
    g = np.load(os.path.join(golden_dir, "thompson_table_samples.npz"))
    np.testing.assert_array_equal(thompson_full.constants(), g["constants"])
    from oracle import oracle as O
    for name in O.THOMPSON_TABLES:
        np.testing.assert_allclose(thompson_full.table(name).ravel()[g[name + "_idx"]], g[name], rtol=1e-12,
                                   err_msg=name)


def test_golden_column(thompson_full, golden_dir):
    g = np.load(os.path.join(golden_dir, "thompson_column_f32.npz"))
    nz = g["columns"].shape[2]
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=3, nz=nz, ny=3), moisten=3.0)
    for call in range(g["columns"].shape[0]):
        thompson_full.driver(st, 60.0, 1, 1, 1, 1, 0, nz - 1)
        got = np.stack([st[k][1, :, 1] for k in ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr")])
        np.testing.assert_allclose(got, g["columns"][call], rtol=1e-6, atol=1e-12, err_msg=f"call {call}")
    # every category was exercised by the frozen scenario
    assert (g["columns"][:, 2:7].max(axis=(0, 2)) > 1e-6).all()
    np.testing.assert_allclose([st[k][1, 1] for k in ("rainnc", "snownc", "graupelnc")], g["ppt"], rtol=1e-6)


def test_no_micro_column_is_left_alone(thompson_full):
    """Warm, dry, cloud-free column: mp_thompson returns at :1363; only th = (th*pii)/pii rounds."""
    nz = 10
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=3, nz=nz, ny=3))
    st["th"] += np.float32(25.0)
    st["qv"][:] = np.float32(1e-4)
    before = ideal.copy_state(st)
    n_exit = thompson_full.driver(st, 60.0, 1, 1, 1, 1, 0, nz - 1)
    assert n_exit == 1
    for k in ("qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "rainnc"):
        np.testing.assert_array_equal(st[k], before[k])
    np.testing.assert_allclose(st["th"], before["th"], rtol=2e-7)


def test_water_budget_closes(thompson_full):
    """Total water of a column (using the scheme's own air density) + fallen precipitation is
    conserved by mp_thompson up to the change of air density with latent heating: mixing ratios are
    per kg of air, sedimentation moves rho*q with the density of the moment (:2660-2770), so the
    budget closes to a few per cent, not to rounding."""
    nx, nz, ny = 6, 24, 4
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)

    def column_water(s):
        T = s["th"].astype(np.float64) * s["pii"]
        qv = s["qv"].astype(np.float64)
        rho = 0.622 * s["pressure"] / (287.04 * T * (qv + 0.622))
        q = sum(s[k].astype(np.float64) for k in ("qv", "qc", "qr", "qi", "qs", "qg"))
        return (q * rho * s["dz_mass"])[1:-1, :, 1:-1].sum()

    w0 = column_water(st)
    for _ in range(20):
        thompson_full.driver(st, 60.0, 1, nx - 2, 1, ny - 2, 0, nz - 1)
    fallen = st["rainnc"].astype(np.float64)[1:-1, 1:-1].sum()
    assert fallen > 0.01 * w0                                  # it did rain
    assert abs((column_water(st) + fallen) - w0) < 0.03 * w0
    for k in ("qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr"):
        assert np.isfinite(st[k]).all() and (st[k] >= 0).all()


def test_driver_clips_tile_to_domain_edge(thompson_full):
    """i_end = MIN(ite, ide-1) (mp_thompson.f90:821-822): a tile that reaches the last domain cell
    leaves that column untouched."""
    nx, nz, ny = 6, 12, 5
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
    before = ideal.copy_state(st)
    thompson_full.driver(st, 60.0, 0, nx - 1, 0, ny - 1, 0, nz - 1)
    assert (st["qc"][:-1, :, :-1] != before["qc"][:-1, :, :-1]).any()
    for k in ("th", "qv", "qc"):
        np.testing.assert_array_equal(st[k][:, :, -1], before[k][:, :, -1])
        np.testing.assert_array_equal(st[k][-1], before[k][-1])


# ------------------------------------------------------------------ more independent table evaluations
def _snow_dsd(th, i, j):
    """N_s(n) of table cell (r_s(i), Tc(j)) in float64, Field et al. (2005) as in mp_thompson.f90:3111-3150."""
    c = th.constants().astype(np.float64)
    f = lambda x: np.float64(np.float32(x))
    sa = [f(v) for v in (5.065339, -0.062659, -3.032362, 0.029469, -0.000285, 0.31255, 0.000204, 0.003199, 0.0, -0.015952)]
    sb = [f(v) for v in (0.476221, -0.015896, 0.165977, 0.007468, -0.000141, 0.060366, 0.000079, 0.000594, 0.0, -0.003577)]
    Tc = f((-0.01, -5., -10., -15., -20., -25., -30., -35., -40.)[j])
    am_s = f(0.069)
    rs = float(np.float32(f"{i % 9 + 1}e{i // 9 - 5}"))

    def poly(s, x):       # evaluated in REAL by the Fortran; float64 here (differences ~1e-7 relative)
        return (s[0] + s[1] * Tc + s[2] * x + s[3] * Tc * x + s[4] * Tc * Tc + s[5] * x * x + s[6] * Tc * Tc * x
                + s[7] * Tc * x * x + s[8] * Tc ** 3 + s[9] * x ** 3)

    M2 = rs / am_s
    second = (M2 / 10.0 ** poly(sa, 2.0)) ** (1.0 / poly(sb, 2.0))          # the bm_s ~ 2 branch (:3112-3124)
    M3 = 10.0 ** poly(sa, 3.0) * second ** poly(sb, 3.0)
    r = M2 / M3
    Mrat, M0, slam1, slam2 = M2 * r ** 3, r ** f(0.6357), r * f(20.78), r * f(3.29)
    Ds = th.bins()["Ds"]
    return Mrat * (f(490.6) * np.exp(-slam1 * Ds) + f(17.46) * M0 * Ds ** f(0.6357) * np.exp(-slam2 * Ds)) * _dt_of(Ds, 200e-6, 2e-2)


@pytest.mark.parametrize("k,m,i,j", [(3, 4, 0, 0), (20, 18, 14, 4), (36, 36, 27, 8), (9, 30, 22, 2)])
def test_qr_acr_qs_against_numpy(thompson, k, m, i, j):
    """The twelve rain-snow tables (mp_thompson.f90:3093-3220), including the massr > 1.5 masss split."""
    th = thompson
    th.big_slab(1, m * 37 + k)
    c = th.constants().astype(np.float64)
    am_r = c[5]
    b = th.bins()
    N_r, rr = _rain_dsd(th, k, m)
    N_s = _snow_dsd(th, i, j)
    Dr, Ds = b["Dr"], b["Ds"]
    f = lambda x: np.float64(np.float32(x))
    vr = f(-0.1021) + f(4.932E3) * Dr - f(0.9551E6) * Dr ** 2 + f(0.07934E9) * Dr ** 3 - f(0.002362E12) * Dr ** 4
    vs = f(np.float32(1.5) * np.float32(40.0)) * Ds ** f(0.55) * np.exp(-f(100.0) * Ds)
    dvs = np.maximum(vr[:, None] - vs[None, :], 0)
    dvr = np.maximum(vs[None, :] - vr[:, None], 0)
    kern = f(PI2 * np.float32(.25) * np.float32(0.95)) * (Ds[None, :] + Dr[:, None]) ** 2 * N_s[None, :] * N_r[:, None]
    massr, masss = (am_r * Dr ** 3)[:, None], (f(0.069) * Ds ** 2)[None, :]
    big = massr > f(1.5) * masss
    want = dict(
        tcs_racs1=(kern * dvs * masss)[big].sum(), tmr_racs1=min((kern * dvs * massr)[big].sum(), rr),
        tcs_racs2=(kern * dvs * masss)[~big].sum(), tmr_racs2=(kern * dvs * massr)[~big].sum(),
        tcr_sacr1=(kern * dvr * massr)[big].sum(), tms_sacr1=(kern * dvr * masss)[big].sum(),
        tcr_sacr2=(kern * dvr * massr)[~big].sum(), tms_sacr2=(kern * dvr * masss)[~big].sum(),
        tnr_racs1=(kern * dvs)[big].sum(), tnr_racs2=(kern * dvs)[~big].sum(),
        tnr_sacr1=(kern * dvr)[big].sum(), tnr_sacr2=(kern * dvr)[~big].sum())
    for name, w in want.items():
        got = th.table(name)[m, k, j, i]
        # the snow-moment regression is REAL arithmetic in the Fortran (and the oracle), float64 here
        assert abs(got - w) <= 2e-4 * abs(w) + 1e-300, (name, got, w)


def test_freezeH2O_against_numpy(thompson):
    """Bigg (1953) freezing tables (mp_thompson.f90:3323-3381) with the largest-drop-first early EXIT."""
    th = thompson
    c = th.constants().astype(np.float64)
    am_r, xm0g = c[5], c[4]
    b = th.bins()
    Dr, Dc = b["Dr"], b["Dc"]
    massr, massc = am_r * Dr ** 3, am_r * Dc ** 3
    rng = np.random.default_rng(1)
    for _ in range(40):
        i, j, k = int(rng.integers(37)), int(rng.integers(37)), int(rng.integers(1, 46))
        Texp = np.exp(float(k)) - 1.0
        N_r, rr = _rain_dsd(th, j, i)
        prob = np.maximum(1.0 - np.exp(-120.0 * (massr / 1000.0) * 5.2e-4 * Texp), 0.0)
        s1 = s2 = n1 = n2 = 0.0
        for n in range(99, -1, -1):
            if massr[n] < xm0g:
                n1 += prob[n] * N_r[n]; s1 += prob[n] * N_r[n] * massr[n]
            else:
                n2 += prob[n] * N_r[n]; s2 += prob[n] * N_r[n] * massr[n]
            if s1 + s2 >= rr:
                break
        got = [th.table(t)[k - 1, j, i] for t in ("tpi_qrfz", "tni_qrfz", "tpg_qrfz", "tnr_qrfz")]
        np.testing.assert_allclose(got, [s1, n1, s2, n2], rtol=2e-5, atol=1e-300)
    # cloud water: all 45 temperatures of a few r_c entries
    mu_c, Nt_c = c[0], 100.e6
    o = 8
    cce, ccg = c[o:o + 3], c[o + 3:o + 6]
    for i in (0, 12, 36):
        rc = float(np.float32(f"{i % 9 + 1}e{i // 9 - 6}"))
        lamc = 1e-6 * (Nt_c * am_r * ccg[1] / ccg[0] / rc) ** (1 / 3)
        N0_c = 1e-18 * Nt_c / ccg[0] * lamc ** cce[0]
        y = Dc * 1e6
        N_c = 1e24 * N0_c * y ** mu_c * np.exp(-lamc * y) * 1e-6          # dtc = 1e-6 (D0c for the first bin too)
        for k in (1, 7, 20, 45):
            # at k = 1 the exponent is below 1e-16 for the small bins, so 1 - exp(-x) is pure rounding of exp:
            # libm's exp (what the Fortran and the oracle call), not NumPy's vector one
            x = -120.0 * (massc / 1000.0) * 5.2e-4 * (math.exp(float(k)) - 1.0)
            prob = np.maximum(1.0 - np.array([math.exp(v) for v in x]), 0.0)
            s1 = sn = 0.0
            for n in range(99, -1, -1):
                sn += prob[n] * N_c[n]; s1 += prob[n] * N_c[n] * massc[n]
                if s1 >= rc:
                    break
            np.testing.assert_allclose([th.table("tpi_qcfz")[k - 1, i], th.table("tni_qcfz")[k - 1, i]], [s1, sn], rtol=2e-5)


def test_qi_aut_qs_against_scipy(thompson):
    """Ice -> snow conversion tables (mp_thompson.f90:3425-3454): the incomplete gamma function and the bin sums."""
    th = thompson
    c = th.constants().astype(np.float64)
    am_i, D0i = c[6], c[2]
    Di = th.bins()["Di"]
    x = np.exp(np.arange(101) / 100 * np.log(5.0 * float(np.float32(200e-6)) / D0i) + np.log(D0i))
    np.testing.assert_allclose(np.sqrt(x[:-1] * x[1:]), Di, rtol=1e-12)
    dti = np.diff(x)
    D0s = float(np.float32(200e-6))
    rng = np.random.default_rng(2)
    checked = 0
    for _ in range(200):
        i, j = int(rng.integers(64)), int(rng.integers(55))
        ri = float(np.float32(f"{i % 9 + 1}e{i // 9 - 10}"))
        Nt = float(10.0 ** (j // 9) * (j % 9 + 1))
        lami = (am_i * 6.0 * Nt / ri) ** (1 / 3)          # cig(2) = gamma(4) = 6, oig1 = 1
        Dm = 4.0 / lami
        got = [th.table(t)[j, i] for t in ("tps_iaus", "tni_iaus", "tpi_ide")]
        if Dm > 5.0 * D0s * (1 + 1e-5):
            want = [ri, Nt, 0.0]
        elif Dm < D0i * (1 - 1e-5):
            want = [0.0, 0.0, 1.0]
        elif abs(Dm / (5.0 * D0s) - 1) < 1e-5 or abs(Dm / D0i - 1) < 1e-5:
            continue                                       # on a branch boundary in REAL arithmetic
        else:
            N_i = Nt * lami * np.exp(-lami * Di) * dti
            big = Di >= D0s
            want = [(N_i * am_i * Di ** 3)[big].sum(), N_i[big].sum(), special.gammainc(2.0, lami * D0s)]
        np.testing.assert_allclose(got, want, rtol=3e-5, atol=1e-300)
        checked += 1
    assert checked > 150
