"""The oracle pinned to the REFERENCE'S OWN SOURCE TEXT.

The reference is Fortran and no Fortran compiler exists in this image, so oracle/f90run executes
src/physics/adv_mpdata.f90 (+ adv_mpdata_FCT_core.f90) and src/physics/mp_simple.f90 themselves: read where they lie,
translated statement by statement to Python/NumPy, run with IEEE binary32 arithmetic and glibc's expf.

  * test_golden_*: the committed outputs of that execution (tests/golden/reference_source_*.npz, made by
    tests/golden/make_reference_source_golden.py) against the C++ oracle -- equal in every bit.  These run anywhere.
  * test_live_*: the same, executed live on fresh inputs, wherever /root/reference exists (this container).
"""
import importlib.util
import os
import re

import numpy as np
import pytest

from icar_b200 import ideal

REF_ROOT = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REF_ROOT, "src", "physics", "adv_mpdata.f90")),
                                     reason="the reference source tree is not on this machine")


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ------------------------------------------------------------------ committed outputs of the reference source
def test_golden_hill_case_oracle_equals_reference_source(oracle, golden_dir):
    g = load(golden_dir, "reference_source_hill_24x10x20_3steps_f32.npz")
    st = ideal.make_ideal_case(nx=24, nz=10, ny=20, terrain=True)
    st["qv"] *= np.float32(3.0)
    dt = np.float32(st["dt"])
    assert float(dt) == float(g["dt"])
    for _ in range(3):
        oracle.step_simple(st, dt, dt, 2, True)
    for k in ("th", "qv", "qc", "qr", "qs", "rain", "snow"):
        assert np.array_equal(st[k], g[k]), k
    assert g["qc"].max() > 1e-4 and g["rain"].max() > 0            # the microphysics was active
    # ... and the file frozen from the oracle in round 1 holds the same numbers
    old = load(golden_dir, "hill_24x10x20_3steps_f32.npz")
    for k in ("th", "qv", "qc", "qr", "qs", "rain", "snow"):
        assert np.array_equal(old[k], g[k]), k


def test_golden_reference_test_scenarios(oracle, golden_dir):
    """src/tests/test_mpdata.f90:51-111 and src/tests/test_mp_simple.f90:4-33, run through the reference's own
    advect3d / mp_simple: the oracle reproduces both"""
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = load(golden_dir, "reference_source_mpdata_scenario_f32.npz")
    rows, line = mg.mpdata_scenario(nloops=2)
    assert np.array_equal(rows, g["rows"]) and np.array_equal(line, g["final_line"])
    assert g["rows"][:, -2].max() <= 2.0 and g["rows"][:, -1].min() >= 1.0          # FCT: no new extrema
    t = load(golden_dir, "reference_source_mp_simple_scenario_f32.npz")["table"]
    assert np.array_equal(mg.mp_simple_scenario(), t)
    assert t[-1, 5] > 0.05 and t[-1, 6] > 0.01                                      # it rained and snowed


def test_golden_random_fields(oracle, golden_dir):
    r = load(golden_dir, "reference_source_random_f32.npz")
    for order, fct in ((1, True), (2, True), (2, False), (3, True)):
        q = r["q"].copy()
        oracle.advect3d(q, None, r["jaco"], r["dz"], r["U"], r["V"], r["W"], order, fct)
        assert np.array_equal(q, r[f"advect3d_order{order}_fct{int(fct)}"]), (order, fct)
    mp = {k: r["mp_in_" + k].copy() for k in ("pressure", "th", "pii", "rho", "dz_mass", "qv", "qc", "qr", "qs")}
    ny, nz, nx = mp["th"].shape
    rain = np.zeros((ny, nx), np.float32)
    snow = np.zeros((ny, nx), np.float32)
    t = [int(x) - 1 for x in r["mp_tile"]]
    for _ in range(3):
        oracle.mp_simple_driver(mp["pressure"], mp["th"], mp["pii"], mp["rho"], mp["qv"], mp["qc"], mp["qr"], mp["qs"],
                                rain, snow, float(r["mp_dt"]), mp["dz_mass"], *t)
    for k in ("th", "qv", "qc", "qr", "qs"):
        assert np.array_equal(mp[k], r["mp_out_" + k]), k
    assert np.array_equal(rain, r["mp_out_rain"]) and np.array_equal(snow, r["mp_out_snow"])


# ------------------------------------------------------------------ the reference source executed live
@pytest.fixture(scope="module")
def ref():
    from oracle import f90run
    return f90run.Reference(REF_ROOT)


@needs_reference
def test_live_translation_is_complete(ref):
    """every routine on the path was translated from the files where they lie, nothing was skipped"""
    for name in ("flux1", "upwind_advection", "mpdata_fluxes", "flux_limiter", "advect3d", "mpdata"):
        assert callable(getattr(ref.adv, name))
    for name in ("sat_mr", "cloud_conversion", "cloud2hydrometeor", "phase_change", "mp_conversions", "sediment",
                 "mp_simple", "mp_simple_driver"):
        assert callable(getattr(ref.mp, name))
    assert "adv_mpdata_FCT_core" not in ref.adv.source and "beta_in_i2" in ref.adv.source       # the include was inlined
    # constants come from the source text, not from this repository
    assert float(ref.mp.M.lh_vapor) == 2.26e6 and float(ref.mp.M.snow_fall_rate) == 1.5
    assert ref.mp.M.rain_formation_time_const == np.float32(1) / np.float32(500.0)


@needs_reference
@pytest.mark.parametrize("order,fct", [(1, True), (2, True), (2, False), (3, True)])
def test_live_advect3d(ref, oracle, order, fct):
    nx, nz, ny = 11, 6, 8
    rng = np.random.default_rng(order * 10 + fct)
    f = np.float32
    q = rng.uniform(0, 2, (ny, nz, nx)).astype(f)
    q[rng.random(q.shape) < 0.3] = 0
    U = rng.uniform(-0.3, 0.3, (ny, nz, nx - 1)).astype(f)
    V = rng.uniform(-0.3, 0.3, (ny - 1, nz, nx)).astype(f)
    W = rng.uniform(-20, 20, (ny, nz, nx)).astype(f)
    jaco = rng.uniform(0.7, 1.0, (ny, nz, nx)).astype(f)
    dz = rng.uniform(300, 400, (ny, nz, nx)).astype(f)
    rho = rng.uniform(0.6, 1.2, (ny, nz, nx)).astype(f)
    ref.set_winds(U, V, W)
    for density in (None, rho):
        want = ref.advect3d(q.copy(), jaco, dz, order, fct, rho=density)
        got = q.copy()
        oracle.advect3d(got, density, jaco, dz, U, V, W, order, fct)
        assert np.array_equal(got, want)
        assert np.abs(want - q).max() > 1e-3


@needs_reference
def test_live_mpdata_driver(ref, oracle):
    """mpdata(domain, options, dt) itself on a mock domain: the wind normalisation of adv_mpdata.f90:500-506 and the
    order of the scalar list, against oracle.wind_normalise + advect3d"""
    st = ideal.make_ideal_case(nx=14, nz=8, ny=9, terrain=True, perturb=True)
    rng = np.random.default_rng(2)
    st["v"] = (st["v"] + rng.uniform(-4, 4, st["v"].shape)).astype(np.float32)
    dt = np.float32(st["dt"])
    a = {k: st[k].copy() for k in ("qv", "th")}
    b = {k: st[k].copy() for k in ("qv", "th")}
    ref.mpdata({"water_vapor": b["qv"], "potential_temperature": b["th"]}, st["u"], st["v"], st["w"], st["jacobian_u"],
               st["jacobian_v"], st["jacobian_w"], st["jacobian"], st["advection_dz"], dt, st["dx"])
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"],
                                    None, dt, np.float32(st["dx"]))
    M = ref.adv.M
    assert np.array_equal(M.u_m.a.T, U) and np.array_equal(M.v_m.a.T, V) and np.array_equal(M.w_m.a.T, W)
    for k in ("qv", "th"):
        oracle.advect3d(a[k], None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
        assert np.array_equal(a[k], b[k]), k


@needs_reference
def test_live_mp_simple_driver(ref, oracle):
    st = ideal.make_ideal_case(nx=10, nz=9, ny=5, terrain=True, perturb=True)
    st["qv"] *= np.float32(3.0)
    names = ("pressure", "th", "pii", "rho", "qv", "qc", "qr", "qs")
    a = {k: st[k].copy() for k in names}
    b = {k: st[k].copy() for k in names}
    ra = np.zeros((5, 10), np.float32); sa = ra.copy(); rb = ra.copy(); sb = ra.copy()
    for call in range(3):
        dt = np.float32(30.0 + 10 * call)
        oracle.mp_simple_driver(*[a[k] for k in names], ra, sa, dt, st["dz_mass"], 1, 8, 0, 3, 0, 8)
        ref.mp_simple_driver(*[b[k] for k in names], rb, sb, dt, st["dz_mass"], tile=(2, 9, 1, 4, 1, 9))
        for k in names:
            assert np.array_equal(a[k], b[k]), (call, k)
        assert np.array_equal(ra, rb) and np.array_equal(sa, sb)
    assert ra.max() > 0 and a["qc"].max() > 1e-4


@needs_reference
def test_live_kts_above_kms_reallocates_temperature(ref, oracle):
    """`temperature = th(i,:,j) * pii(i,:,j)` reallocates the kts:kte array to the whole column (mp_simple.f90:633),
    so conversions run on kts:kte while maxval(qr) and sedimentation see the whole column: same in the oracle"""
    st = ideal.make_ideal_case(nx=6, nz=8, ny=4, terrain=True, perturb=True)
    st["qv"] *= np.float32(3.0)
    rng = np.random.default_rng(4)
    st["qr"] = rng.uniform(0, 1e-3, st["qr"].shape).astype(np.float32)
    names = ("pressure", "th", "pii", "rho", "qv", "qc", "qr", "qs")
    a = {k: st[k].copy() for k in names}
    b = {k: st[k].copy() for k in names}
    ra = np.zeros((4, 6), np.float32); sa = ra.copy(); rb = ra.copy(); sb = ra.copy()
    oracle.mp_simple_driver(*[a[k] for k in names], ra, sa, 40.0, st["dz_mass"], 0, 5, 0, 3, 1, 6)
    ref.mp_simple_driver(*[b[k] for k in names], rb, sb, 40.0, st["dz_mass"], tile=(1, 6, 1, 4, 2, 7))
    for k in names:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(ra, rb)


# ------------------------------------------------------------------ Thompson: src/physics/mp_thompson.f90 executed
THOMPSON_KEYS = ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "rainnc", "snownc", "graupelnc", "sr")


@pytest.fixture(scope="module")
def thompson_full(oracle):
    return oracle.Thompson(big_tables=True)


def thompson_state(nx=6, nz=16, ny=4):
    return ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True), moisten=3.0)


def test_golden_thompson_init(oracle, golden_dir, thompson_full):
    """thompson_init as the reference's text computes it -- 130 scalar constants, the five bin grids, the freezing /
    ice-to-snow / collision-efficiency tables, three slabs of each rain-collection table from qr_acr_qg / qr_acr_qs run
    for those slabs -- against the C++ oracle: equal in every bit"""
    g = load(golden_dir, "reference_source_thompson.npz")
    assert np.array_equal(thompson_full.constants(), g["constants"])
    bins = thompson_full.bins()
    for n in ("Dc", "Di", "Dr", "Ds", "Dg"):
        assert np.array_equal(bins[n], g["bins_" + n.lower()]), n
    for n in oracle.THOMPSON_TABLES[18:]:
        flat = thompson_full.table(n).ravel()                   # the view is in the Fortran's storage order
        want = g["table_" + n]
        got = flat[g["table_" + n + "_idx"]] if ("table_" + n + "_idx") in g.files else flat
        assert np.array_equal(got, want), n
        assert (want != 0).any()
    from oracle.f90run import Reference
    # THOMPSON_SLABS of make_reference_source_golden.py: the first, an inner and the last (k, m) slab of each builder
    for which, km in ((0, 0), (0, 5 * 37 + 11), (0, 36 * 37 + 36), (1, 0), (1, 20 * 37 + 7), (1, 36 * 37 + 36)):
        m, k = divmod(km, 37)
        names = Reference.BIG_QG if which == 0 else Reference.BIG_QS
        for n in names:
            assert np.array_equal(thompson_full.table(n)[m, k].T, g[f"slab{km}_{n}"]), (n, km)
        # (in some rain-snow slabs the `sacr1` / `racs2` tables are zero: that branch of the collection is not taken there)
        assert sum(int((g[f"slab{km}_{n}"] != 0).sum()) for n in names) >= 1000, km


def test_golden_thompson_driver(golden_dir, thompson_full):
    """six calls of mp_gt_driver + mp_thompson on a moist hill state in which cloud water, rain, ice, snow and graupel all
    form: the oracle against the reference's text, every field of every call equal in every bit"""
    g = load(golden_dir, "reference_source_thompson.npz")
    st = thompson_state()
    ny, nz, nx = st["th"].shape
    for call in range(1, 7):
        thompson_full.driver(st, 60.0, 0, nx - 1, 0, ny - 1, 0, nz - 1)
        for k in THOMPSON_KEYS:
            assert np.array_equal(st[k], g[f"call{call}_{k}"]), (call, k)
    for k in ("qc", "qr", "qi", "qs", "qg"):
        assert g["call6_" + k].max() > 1e-6, k
    assert g["call6_rainnc"].max() > 0.1


@needs_reference
def test_live_thompson(ref, oracle, thompson_full):
    """mp_thompson.f90 translated and run live: thompson_init (without freezeH2O's two-minute loop, whose tables the
    golden file covers) and two driver calls on fresh inputs"""
    tabs = {n: thompson_full.table(n).T for n in oracle.THOMPSON_TABLES}
    ref.thompson_init(tabs, skip_freezeh2o=True)
    assert np.array_equal(ref.thompson_constants(), thompson_full.constants())
    M = ref.th.M
    for n in ("tps_iaus", "tni_iaus", "tpi_ide", "t_Efrw", "t_Efsw"):
        assert np.array_equal(getattr(M, n.lower()).a, thompson_full.table(n).T), n
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=4, nz=14, ny=3, perturb=True), moisten=3.5)
    rng = np.random.default_rng(8)
    for k, top in (("qi", 5e-5), ("qs", 2e-4), ("qg", 2e-4), ("qr", 5e-4)):
        st[k] = (rng.uniform(0, top, st[k].shape) * (rng.random(st[k].shape) < 0.6)).astype(np.float32)
    st["ni"] = (st["qi"] * np.float32(1e9)).astype(np.float32)
    st["nr"] = (st["qr"] * np.float32(1e8)).astype(np.float32)
    a, b = ideal.copy_state(st), ideal.copy_state(st)
    for call in range(2):
        thompson_full.driver(a, 45.0, 0, 3, 0, 2, 0, 13)
        ref.mp_gt_driver(b, 45.0)
        for k in THOMPSON_KEYS:
            assert np.array_equal(a[k], b[k]), (call, k)
    assert np.abs(a["th"] - st["th"]).max() > 0.1


# ------------------------------------------------------------------ the rows around the hot path
@needs_reference
def test_live_upwind_advection_scheme(ref, oracle):
    """advection = 1: upwind(domain, options, dt) of src/physics/advect.f90 (setup_module_winds :304-336, whose factors are
    multiplied in another order than mpdata's, and its own advect3d :106-178) against the oracle's scheme 1 + order 1"""
    st = ideal.make_ideal_case(nx=14, nz=8, ny=9, terrain=True, perturb=True)
    rng = np.random.default_rng(3)
    st["v"] = (st["v"] + rng.uniform(-4, 4, st["v"].shape)).astype(np.float32)
    dt = np.float32(st["dt"])
    a = {k: st[k].copy() for k in ("qv", "th")}
    b = {k: st[k].copy() for k in ("qv", "th")}
    ref.upwind({"water_vapor": b["qv"], "potential_temperature": b["th"]}, st["u"], st["v"], st["w"], st["jacobian_u"],
               st["jacobian_v"], st["jacobian_w"], st["jacobian"], st["advection_dz"], dt, st["dx"])
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None,
                                    dt, np.float32(st["dx"]), 1)
    M = ref.upw.M
    assert np.array_equal(M.u_m.a.T, U) and np.array_equal(M.v_m.a.T, V) and np.array_equal(M.w_m.a.T, W)
    for k in ("qv", "th"):
        oracle.advect3d(a[k], None, st["jacobian"], st["advection_dz"], U, V, W, 1, True)
        assert np.array_equal(a[k], b[k]), k
        assert np.abs(a[k] - st[k]).max() > 0


@needs_reference
def test_live_balance_uvw(ref):
    """balance_uvw + calc_divergence(horz_only) of src/physics/wind.f90:81-228 against the restatement the GPU test of
    icar_balance_uvw uses (icar_b200/ideal.py)"""
    st = ideal.make_ideal_case(nx=14, nz=8, ny=9, terrain=True, perturb=True)
    rng = np.random.default_rng(3)
    st["v"] = (st["v"] + rng.uniform(-4, 4, st["v"].shape)).astype(np.float32)
    w = np.full_like(st["w"], 7.0)                           # balance_uvw starts from w = 0 (:104)
    ref.balance_uvw(st["u"], st["v"], w, st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], st["advection_dz"], st["dx"],
                    st["jacobian"])
    want = ideal.balance_uvw(st["u"], st["v"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], st["advection_dz"], st["dx"])
    assert np.array_equal(w, want) and np.abs(want).max() > 1.0


@needs_reference
@pytest.mark.parametrize("strictness", [1, 2, 3, 4, 5])
def test_live_compute_dt(ref, strictness):
    """compute_dt of src/main/time_step.f90:217-330 (all five cfl_strictness levels, use_density = .false.) against the
    restatement the GPU test of icar_compute_dt uses; where the reference stops ("time step too small") both must"""
    from oracle import np_restatement as NP
    from oracle.f90run.runtime import FortranStop
    st = ideal.make_ideal_case(nx=13, nz=7, ny=9, perturb=True)
    rng = np.random.default_rng(strictness)
    for scale in (1.0, 0.01):
        u = (st["u"] * rng.uniform(0.2, 1.5, st["u"].shape) * scale).astype(np.float32)
        v = (st["v"] * scale).astype(np.float32)
        w = (rng.uniform(-2, 2, st["w"].shape) * scale).astype(np.float32)
        try:
            want = NP.compute_dt_np(st["dx"], u, v, w, st["dz_levels"], 0.9, strictness)
        except FloatingPointError:
            want = "stop"
        try:
            got = min(float(ref.compute_dt(st["dx"], u, v, w, st["dz_levels"], 0.9, strictness)), 120.0)   # update_dt's cap, :421
        except FortranStop:
            got = "stop"
        assert got == want, (strictness, scale)


@needs_reference
@pytest.mark.parametrize("strictness", [1, 2, 3, 4, 5])
def test_live_update_dt(ref, strictness):
    """update_dt of src/main/time_step.f90:375-423 executed as written (the branch a build without __INTEL_COMPILER takes):
    compute_dt with use_density = .false. on the domain's winds, co_min over the images, then the 120 s cap -- the order
    icar_compute_dt keeps (minimum first, cap after) and the value its restatement gives."""
    from oracle import np_restatement as NP
    st = ideal.make_ideal_case(nx=24, nz=8, ny=10, perturb=True)
    for scale in (0.5, 1.0, 3.0):
        u = (st["u"] * np.float32(scale)).astype(np.float32)
        try:
            own = NP.compute_dt_np(st["dx"], u, st["v"], st["w"], st["dz_levels"], 0.9, strictness)      # capped at 120
        except FloatingPointError:           # levels 1, 2 and 5 take the winds in m/s: "time step too small", both must stop
            from oracle.f90run.runtime import FortranStop
            with pytest.raises(FortranStop):
                ref.update_dt(st["dx"], u, st["v"], st["w"], st["dz_levels"], 0.9, strictness)
            continue
        assert float(ref.update_dt(st["dx"], u, st["v"], st["w"], st["dz_levels"], 0.9, strictness)) == own
        for others in ((300.0, 7.5), (500.0,), (119.0, 121.0)):
            got = float(ref.update_dt(st["dx"], u, st["v"], st["w"], st["dz_levels"], 0.9, strictness, other_images=others))
            assert got == min(own, min(others), 120.0), (scale, others)


@needs_reference
def test_live_domain_halo_calls(ref):
    """domain%halo_send / halo_retrieve / halo_exchange of domain_obj.f90:109-155 executed on a domain that has the scalars
    of each scheme: every one of them is sent once and retrieved once, potential_temperature's retrieve is the one that
    synchronises (the others pass no_sync), and the order of the sends is the order of the product's exchange lists
    (kExchangedSimple / kExchangedThompson in icar_b200.cu, which icar_halo_send packs in)."""
    member = {"WATER_VAPOR": "water_vapor", "POTENTIAL_TEMPERATURE": "potential_temperature", "CLOUD_WATER": "cloud_water_mass",
              "CLOUD_ICE": "cloud_ice_mass", "ICE_NUMBER": "cloud_ice_number", "RAIN_IN_AIR": "rain_mass",
              "RAIN_NUMBER": "rain_number", "SNOW_IN_AIR": "snow_mass", "GRAUPEL_IN_AIR": "graupel_mass"}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "icar_b200", "csrc", "icar_b200.cu")).read()
    for table, count in (("kExchangedSimple", 5), ("kExchangedThompson", 9)):
        m = re.search(table + r"\[\d+\]\s*=\s*\{([^}]*)\}", src)
        ours = [member[v.strip().replace("ICAR_VAR_", "")] for v in m.group(1).split(",")]
        assert len(ours) == count
        log = ref.domain_halo_calls(set(ours))
        sends = [n for n, what, _ in log if what == "send"]
        assert sends == ours                                                    # same variables, same order
        retrieves = [(n, no_sync) for n, what, no_sync in log if what == "retrieve"]
        assert sorted(n for n, _ in retrieves) == sorted(ours)
        assert retrieves[0] == ("potential_temperature", False) and all(ns for _, ns in retrieves[1:])
        assert log == ref.domain_halo_calls(set(ours), "halo_send") + ref.domain_halo_calls(set(ours), "halo_retrieve")
    assert ref.domain_halo_calls(set()) == []


@needs_reference
def test_live_update_winds_order(ref):
    """update_winds of src/physics/wind.f90:289-369 executed with its callees recorded: for windtype = kWIND_LINEAR the
    once-per-forcing-interval wind update is  rotate to the grid -> linear_perturb(reverse = .False.) -> balance_uvw,
    on u / v / w at start-up (update = .False.) and on the next forcing step's winds afterwards (update = .True.) --
    the order `bench.py --linear-winds` and the shim's b200_linear_perturb keep (icar_linear_perturb, then
    icar_balance_uvw; the rotation is the identity on the ideal grids and stays on the host)."""
    first = ref.update_winds_calls("kwind_linear", True)
    assert first == [("init_winds", None, None), ("make_winds_grid_relative", "u", None), ("linear_perturb", None, False),
                     ("balance_uvw", "u", None)]
    later = ref.update_winds_calls("kwind_linear", False)
    nxt = "u of the next forcing step"
    assert later == [("make_winds_grid_relative", nxt, None), ("linear_perturb", None, True), ("balance_uvw", nxt, None)]
    # without a wind scheme only the balance remains; every variant ends with it
    assert [c[0] for c in ref.update_winds_calls(0, True)] == ["init_winds", "make_winds_grid_relative", "balance_uvw"]
    for wt in ("kconserve_mass", "kiterative_winds", "klinear_iterative_winds"):
        for f in (True, False):
            assert ref.update_winds_calls(wt, f)[-1][0] == "balance_uvw"
    bench = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    assert bench.index("dom.linear_perturb(") < bench.index("dom.balance_uvw()")


@needs_reference
def test_live_advect_dispatch(ref):
    """advect() of advection_driver.f90:51-66 and the scheme numbers of icar_constants.f90:347-372, executed: which
    scheme runs for which options%physics%advection -- the numbering icar_options uses (include/icar_b200.h) and the
    "nothing for any other value" that icar_advect keeps."""
    assert ref.advect_dispatch(1) == [("upwind", 120.0)]
    assert ref.advect_dispatch(2) == [("mpdata", 120.0)]
    assert ref.advect_dispatch(0) == [] and ref.advect_dispatch(3) == []
    c = ref.constants
    assert (int(c.kadv_upwind), int(c.kadv_mpdata), int(c.kmp_thompson), int(c.kmp_sb04)) == (1, 2, 1, 2)
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "icar_b200.h")).read()
    assert "1 kADV_UPWIND, 2 kADV_MPDATA" in hdr and "1 kMP_THOMPSON" in hdr


def _passes(log):
    """The log of a step loop cut into passes: (dt handed to the physics or None, dt of apply_forcing, enforce_limits ran),
    after checking that every pass is exactly the source's sequence of calls."""
    full = ["update_dt", "diagnostic_update", "rad", "lsm", "pbl", "convect", "mp", "halo_send", "mp", "halo_retrieve",
            "advect", "apply_forcing"]
    out, i = [], 0
    while i < len(log):
        assert log[i] == ("update_dt",) and log[i + 1] == ("diagnostic_update",)
        j = i + 2
        while j < len(log) and log[j][0] != "update_dt":
            j += 1
        names = [e[0] for e in log[i:j]]
        if names == ["update_dt", "diagnostic_update"]:
            out.append((None, None, False))                             # a pass that only advances the clock
        else:
            limits = names[-1] == "enforce_limits"
            assert names[:len(names) - limits] == full, names
            body = log[i + 2:j - limits]
            assert body[4][2] == ("halo", 1) and body[6][2] == ("subset", 1)
            rdt = {e[1] for e in body[:-1] if len(e) > 1}
            assert len(rdt) == 1                                          # every physics call gets the same real(dt)
            out.append((rdt.pop(), body[-1][1], limits))
        i = j
    return out


@needs_reference
@pytest.mark.parametrize("start,end,dts", [
    (1000.0, 1290.0, [120.0] * 4),                       # two full steps and a clipped one
    (0.0, 240.0005, [120.0] * 4),                        # the remainder is half a millisecond: no physics for it (:484)
    (3600.0, 3960.0, [120.0] * 5),                       # dt divides the interval
    (86400.0 * 400, 86400.0 * 400 + 3600.0, [37.3] * 120),   # many steps, a dt that no binary fraction of a day holds
    (50.0, 60.0, [120.0] * 2),                           # one clipped step
    (0.0, 600.0, [120.0, 80.0, 33.25, 120.0, 120.0, 120.0, 120.0, 120.0]),    # dt changes from step to step
])
def test_live_step_loop(ref, start, end, dts):
    """step(domain, end_time, options) of src/main/time_step.f90:438-549 executed as written (physics replaced by recorders,
    model time in the reference's REAL*8 days) against the control flow of icar_step: the same calls in the same order in
    every pass, the same real(dt) for the physics, apply_forcing's dt and the final time equal to within the rounding of
    a day count (1e-6 s), enforce_limits in the same passes.  Passes that only advance the clock (dt <= 1e-3 s: what a
    clock in days leaves over, :484) may differ in number; nothing runs in them."""
    got_log, got_t = ref.step_loop(start, end, dts)
    from oracle import np_restatement as NP
    want_log, want_t = NP.step_loop_np(start, end, dts)
    got, want = _passes(got_log), _passes(want_log)
    physics = lambda ps: [p for p in ps if p[0] is not None]            # noqa: E731
    assert len(physics(got)) == len(physics(want)) >= 1
    t0, ties = start, 0
    for g, w in zip(physics(got), physics(want)):
        assert abs(g[0] - w[0]) <= 1e-5 and abs(g[1] - w[1]) <= 1e-5, (g, w)
        # `(end_time - model_time) < 2 dt` (:540): where the two sides are EQUAL in seconds the reference's clock in days
        # decides by its rounding (3600 s -> 3960 s in steps of 120 s: it clamps one update earlier); a clock in
        # seconds has no such rounding and says "not less"
        if abs((end - t0) - 2 * w[1]) <= 1e-5:
            ties += 1
            assert not w[2]
        else:
            assert g[2] == w[2], (g, w)
        t0 += w[1]
    assert ties == (1 if (start, end) == (3600.0, 3960.0) else 0)
    assert abs(got_t - want_t) <= 1e-5 and abs(want_t - end) <= 1e-9
    if end == 240.0005:
        assert got[-1][0] is None and want[-1][0] is None and len(physics(want)) == 2


@needs_reference
def test_live_diagnostic_update(ref):
    """diagnostic_update of src/main/time_step.f90:49-198 on a mock domain (exner through exner_function of
    atm_utilities.f90 with Rd, cp read from icar_constants.f90) against the restatements the GPU tests of
    icar_diagnostic_update / icar_diagnostic_output use"""
    from oracle import np_restatement as NP
    st = ideal.make_ideal_case(nx=14, nz=8, ny=9, terrain=True, perturb=True)
    rng = np.random.default_rng(5)
    st["v"] = (st["v"] + rng.uniform(-3, 3, st["v"].shape)).astype(np.float32)
    dzdx, dzdy = ideal.terrain_slopes(st)
    dzdy = (dzdy + rng.uniform(-0.05, 0.05, dzdy.shape)).astype(np.float32)
    got = ref.diagnostic_update(st["pressure"], st["th"], st["u"], st["v"], st["w"], st["jacobian"], dzdx, dzdy)
    want = NP.diagnostic_output_np(st["pressure"], st["th"], st["u"], st["v"], st["w"], st["jacobian"], dzdx, dzdy)
    want["exner"], want["density"] = NP.diagnostic_update_np(st["pressure"], st["th"])
    assert float(ref.constants.rd) == float(NP.RD) and float(ref.constants.cp) == float(NP.CP)
    for k in got:
        assert np.array_equal(got[k], want[k]), k
    assert np.abs(want["w_real"]).max() > 0.1


@needs_reference
def test_live_mp_driver_schedule(ref):
    """mp(domain, options, dt_in, halo, subset) of src/physics/mp_driver.f90:673-772 executed with its scheme calls
    replaced by recorders: the update_interval rule, the mp_dt lag (mp_dt = model_time - last_model_time, first call
    max(update_interval, dt_in)) and the halo / subset tile arithmetic are what icar_mp implements -- the schedule below
    is the one tests/test_gpu_parity.py::test_mp_update_interval_and_dt_lag holds the device to."""
    g = dict(ids=1, ide=20, jds=1, jde=16, kds=1, kde=10, ims=1, ime=20, jms=1, jme=16, kms=1, kme=10,
             its=2, ite=19, jts=2, jte=15, kts=1, kte=10, nx=20, ny=16, nz=10)
    log = ref.mp_schedule([(0.0, 30.0, None, None), (30.0, 10.0, None, None), (40.0, 30.0, None, None)], 50, g)
    assert log == [[("process_subdomain", 50.0, 2, 19, 2, 15, 1, 10)], [], [("process_subdomain", 40.0, 2, 19, 2, 15, 1, 10)]]
    # step(): mp(halo=1) then mp(subset=1) at the same model time: the halo call must not advance last_model_time
    log = ref.mp_schedule([(0.0, 30.0, 1, None), (0.0, 30.0, None, 1), (30.0, 30.0, 1, None), (30.0, 30.0, None, 1)], 0, g)
    # process_halo (:609-658) itself: west and east columns over all rows, then south and north rows without the corners --
    # the four strips icar_mp queues on its auxiliary stream (icar_b200.cu, `has_halo`)
    ring = [("process_subdomain", 30.0, 2, 2, 2, 15, 1, 10), ("process_subdomain", 30.0, 19, 19, 2, 15, 1, 10),
            ("process_subdomain", 30.0, 3, 18, 2, 2, 1, 10), ("process_subdomain", 30.0, 3, 18, 15, 15, 1, 10)]
    assert log == [ring, [("process_subdomain", 30.0, 3, 18, 3, 14, 1, 10)], ring, [("process_subdomain", 30.0, 3, 18, 3, 14, 1, 10)]]
    wide = ref.mp_schedule([(0.0, 20.0, 3, None)], 0, g)[0]
    assert [t[2:6] for t in wide] == [(2, 4, 2, 15), (17, 19, 2, 15), (5, 16, 2, 4), (5, 16, 13, 15)]


@needs_reference
@pytest.mark.parametrize("n,nimg,halo", [(512, 8, 3), (512, 8, 1), (100, 3, 1), (1024, 4, 3), (37, 5, 2)])
def test_live_grid_bounds(ref, n, nimg, halo):
    """grid_t%set_grid_dimensions of src/objects/grid_obj.f90:144-255 (my_n, my_start, update_with_halos) executed for every
    image of a slab decomposition (ximages = nimages, yimages = 1 and the transpose -- the choice BASELINE.json prescribes in
    place of domain_decomposition's most-square split) against icar_grid_xslab / icar_grid_yslab (pure host functions)"""
    import icar_b200
    keys = ("ims", "ime", "kms", "kme", "jms", "jme", "its", "ite", "kts", "kte", "jts", "jte", "ids", "ide", "kds", "kde",
            "jds", "jde", "halo_size")
    for img in range(1, nimg + 1):
        want = ref.grid_dimensions(n, 64, 20, nimg, img, halo, "x")
        g = icar_b200.Grid.xslab(n, 20, 64, nimg, img, halo)
        assert {k: getattr(g, k) for k in keys} == {k: want[k] for k in keys}, ("x", img)
        want = ref.grid_dimensions(64, n, 20, nimg, img, halo, "y")
        g = icar_b200.Grid.yslab(64, 20, n, nimg, img, halo)
        assert {k: getattr(g, k) for k in keys} == {k: want[k] for k in keys}, ("y", img)
    # domain_decomposition as written picks the most square split instead: 8 images of a square domain are 2 x 4
    if (n, nimg) == (512, 8):
        sq = ref.grid_dimensions(512, 512, 40, 8, 3, 1, "reference")
        assert (sq["ximages"], sq["yimages"], sq["ximg"], sq["yimg"]) == (2, 4, 1, 2)


@needs_reference
def test_live_linear_winds(ref):
    """src/physics/linear_winds.f90 executed (complex double arithmetic; FFTW's entry points bound to numpy.fft; the
    reference's own fftshift.f90 with its single-precision temporary): initialize_linear_theory_data, the per-height
    solution with l /= 0, the layer means over constant and terrain-following bounds, add_buffer_topo -- against
    oracle/linear_winds_np.py, the restatement the cuFFT path is tested against"""
    from oracle import linear_winds_np as LW
    from oracle.f90run.runtime import unallocated
    lw = ref.linear_winds()
    nx, ny, dx = 24, 20, 2000.0
    lt = ref.lt_data(nx, ny, dx)
    L = LW.LinearTheoryData(nx, ny, dx)
    assert np.array_equal(lt.k.a, L.k) and np.array_equal(lt.l.a, L.l) and np.array_equal(lt.kl.a, L.kl)
    rng = np.random.default_rng(0)
    f = np.float32
    terrain = (500 * np.exp(-(((np.arange(nx)[:, None] - 12) / 4.0) ** 2 + ((np.arange(ny)[None, :] - 9) / 3.0) ** 2))).astype(f)
    ft = np.asfortranarray(np.fft.fft2(terrain.astype(np.complex128)) / (nx * ny))
    ft_np = LW.fftshift2cc(ft.copy())
    lw.shift.fftshift2cc(ft)
    assert np.array_equal(ft, ft_np)
    for (U, V, N2, z) in ((f(10), f(3), f(1e-4), f(800)), (f(-5), f(7), f(3e-5), f(2500)), (f(0), f(0), f(1e-4), f(500))):
        lw.linear_perturbation_at_height(U, V, N2, z, ft, lt)
        u, v = LW.linear_perturbation_at_height(U, V, N2, z, ft_np, L)
        assert np.array_equal(lt.u_perturb.a, u) and np.array_equal(lt.v_perturb.a, v)
    assert np.abs(u).max() == 0                                      # calm background: no perturbation (:192-196)
    lw.linear_perturbation_constz(f(10), f(3), f(1e-4), f(400), f(1200), f(100), ft, lt)
    u, v = LW.linear_perturbation_constz(f(10), f(3), f(1e-4), f(400), f(1200), f(100), ft_np, L)
    assert np.array_equal(lt.u_perturb.a, u) and np.array_equal(lt.v_perturb.a, v) and np.abs(u).max() > 0.1
    buffer = 3
    lw.M.buffer = buffer
    zb = rng.uniform(300, 500, (nx - 2 * buffer, ny - 2 * buffer)).astype(f)
    zt = (zb + rng.uniform(150, 400, zb.shape)).astype(f)
    lw.linear_perturbation_varyingz(f(10), f(3), f(1e-4), np.asfortranarray(zb), np.asfortranarray(zt), f(100), ft, lt)
    u, v = LW.linear_perturbation_varyingz(f(10), f(3), f(1e-4), zb, zt, f(100), ft_np, L, buffer)
    assert np.array_equal(lt.u_perturb.a, u) and np.array_equal(lt.v_perturb.a, v)
    for smooth_window, buf in ((5, 4), (0, 2), (2, 6)):
        bt = unallocated(np.complex128, 2)
        lw.add_buffer_topo(np.asfortranarray(terrain), bt, smooth_window, buf, None)
        want = LW.add_buffer_topo(terrain, smooth_window, buf)
        # (the smoothing window is a SUM over up to 11 x 11 cells: the Fortran adds them in array order, NumPy pairwise)
        assert np.abs(bt.a - want).max() <= 1e-15 * np.abs(want).max()
        if smooth_window == 0:
            assert np.array_equal(bt.a, want)


@needs_reference
def test_live_apply_forcing_and_enforce_limits(ref):
    """domain%apply_forcing (src/objects/domain_obj.f90:2383-2448: `data + dqdt * dt%seconds()` with a REAL*8 seconds, i.e. a
    REAL*8 product and sum rounded once on assignment; boundary cells only for force_boundaries variables; w everywhere)
    and domain%enforce_limits (:2228-2243) executed on a mock domain, against the restatements icar_apply_forcing /
    icar_enforce_limits are tested against"""
    from oracle import np_restatement as NP
    rng = np.random.default_rng(0)
    f = np.float32
    mk = lambda: rng.uniform(-1, 1, (7, 5, 9)).astype(f)                # noqa: E731
    q1, t1, q2, t2, w, tw = mk(), mk() * f(1e-3), mk(), mk() * f(1e-3), mk(), mk() * f(1e-3)
    sec = 37.123456789
    for edges in (dict(west=True, east=False, south=True, north=True), dict(west=False, east=True, south=False, north=False)):
        a = [x.copy() for x in (q1, q2, w)]
        b = [x.copy() for x in (q1, q2, w)]
        ref.apply_forcing([(b[0], t1, True), (b[1], t2, False)], (b[2], tw), sec, **edges)
        NP.apply_forcing_np(a[0], t1, sec, True, **edges)
        NP.apply_forcing_np(a[1], t2, sec, False)
        NP.apply_forcing_np(a[2], tw, sec, False)
        for x, y, z in zip(a, b, (q1, q2, w)):
            assert np.array_equal(x, y) and not np.array_equal(x, z)
        # a REAL product would differ: the REAL*8 evaluation is what the source text says
        assert not np.array_equal(b[1], q2 + t2 * f(sec))
    fields = {"water_vapor": mk(), "rain_mass": mk(), "potential_temperature": mk() + f(300)}
    c = {k: v.copy() for k, v in fields.items()}
    ref.enforce_limits(fields)
    NP.enforce_limits_np(list(c.values()))
    assert all(np.array_equal(fields[k], c[k]) for k in c) and min(v.min() for v in fields.values()) == 0


@needs_reference
@pytest.mark.parametrize("halo", [1, 3])
def test_live_halo_exchange(ref, halo):
    """exchangeable_t%send / %retrieve of src/objects/exchangeable_obj.f90:138-162, 249-356 executed for three images in a
    row along x (the coarray PUT `halo_west_in(...)[east_neighbor] = data_3d(...)` lands in the neighbour's object): every
    ghost column ends up holding the neighbour's owned column of the global field, and the columns moved are exactly the
    ones icar_halo_plan names (what k_halo_push stores into the neighbours' ghost cells)"""
    import icar_b200
    nxg, nz, ny, nimg = 37, 4, 6, 3
    F = np.random.default_rng(0).uniform(0, 1, (ny, nz, nxg)).astype(np.float32)
    grids = [icar_b200.Grid.xslab(nxg, nz, ny, nimg, p, halo) for p in range(1, nimg + 1)]
    tiles = [np.ascontiguousarray(F[:, :, g.ims - 1:g.ime]) for g in grids]
    want = [t.copy() for t in tiles]
    for g, t in zip(grids, tiles):                         # ghost columns hold rubbish before the exchange
        if g.ims > 1:
            t[:, :, :halo] = -1
        if g.ime < nxg:
            t[:, :, -halo:] = -1
    got = [t.copy() for t in tiles]
    ref.halo_exchange_x(got, halo)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    mine = [t.copy() for t in tiles]
    plans = [icar_b200.halo_plan(g, p, nimg) for p, g in enumerate(grids, 1)]
    for p, pl in enumerate(plans):
        w = pl["width"]
        assert w == halo
        if pl["has_hi"]:
            mine[p + 1][:, :, plans[p + 1]["recv_lo"]:plans[p + 1]["recv_lo"] + w] = tiles[p][:, :, pl["send_hi"]:pl["send_hi"] + w]
        if pl["has_lo"]:
            mine[p - 1][:, :, plans[p - 1]["recv_hi"]:plans[p - 1]["recv_hi"] + w] = tiles[p][:, :, pl["send_lo"]:pl["send_lo"] + w]
    for a, b in zip(mine, got):
        assert np.array_equal(a, b)


@needs_reference
@pytest.mark.parametrize("smooth", [False, True])
def test_live_spatial_winds(ref, smooth):
    """spatial_winds of src/physics/linear_winds.f90:840-1127 executed on a mock domain, with calc_stability / calc_direction
    / calc_speed of atm_utilities.f90 and calc_weight / smooth_array / linear_space of array_utilities.f90 executed as
    well: per-cell stability and its two smoothings, column-mean wind, table positions and weights, the trilinear
    interpolation between eight table entries, the relaxation -- against oracle/spatial_winds_np.py (the restatement
    icar_linear_perturb_spatial is tested against) with a RANDOM look-up table handed to both: u, v and the stored
    perturbations equal in every bit.  (What initialize_spatial_winds puts INTO the table -- linear_perturbation for the
    node's wind and stability, staggered -- is covered by test_live_linear_winds, the loop around it is not executed.)"""
    from oracle import spatial_winds_np as SW
    nx, nz, ny = 9, 6, 7
    st0 = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, terrain=True, perturb=True)
    rng = np.random.default_rng(1)
    f = np.float32
    st0["u"] = (st0["u"] + rng.uniform(-6, 6, st0["u"].shape)).astype(f)
    st0["v"] = (st0["v"] + rng.uniform(-6, 6, st0["v"].shape)).astype(f)
    st0["qc"] = rng.uniform(0, 2e-7, st0["qv"].shape).astype(f)          # both branches of calc_stability (qc < 1e-7)
    o = SW.LTOptions()
    o.smooth_nsq, o.vert_smooth, o.stability_window_size, o.linear_update_fraction = smooth, 2, 2, f(0.6)
    lut_u = rng.uniform(-3, 3, (o.n_spd_values, o.n_dir_values, o.n_nsq_values, nx + 1, nz, ny)).astype(f)
    lut_v = rng.uniform(-3, 3, (o.n_spd_values, o.n_dir_values, o.n_nsq_values, nx, nz, ny + 1)).astype(f)

    class Table:
        dirs, nsqs, spds = SW.table_axes(o)

        def get(self, s, d, n, z):
            return np.ascontiguousarray(lut_u[s, d, n, :, z, :].T), np.ascontiguousarray(lut_v[s, d, n, :, z, :].T)
    copy = lambda st: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}      # noqa: E731
    a, b = copy(st0), copy(st0)
    up_a = rng.uniform(-1, 1, a["u"].shape).astype(f)
    vp_a = rng.uniform(-1, 1, a["v"].shape).astype(f)
    up_b, vp_b = up_a.copy(), vp_a.copy()
    lognsq = SW.spatial_winds(a, Table(), o, up_a, vp_a)
    nsq = ref.spatial_winds(b, lut_u, lut_v, o, up_b, vp_b)
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["v"], b["v"])
    assert np.array_equal(up_a, up_b) and np.array_equal(vp_a, vp_b)
    assert np.array_equal(SW.expf(lognsq), nsq)                          # the routine leaves exp(log N^2) in domain%nsquared
    assert np.abs(b["u"] - st0["u"]).max() > 1.0
    M = ref._sw[0].M
    assert np.array_equal(M.dir_values.a, Table.dirs) and np.array_equal(M.nsq_values.a, Table.nsqs) and \
        np.array_equal(M.spd_values.a, Table.spds)                       # linear_space (array_utilities.f90:233-235)


@needs_reference
@pytest.mark.parametrize("varying", [True, False])
def test_live_spatial_winds_table(ref, varying):
    """initialize_spatial_winds (src/physics/linear_winds.f90:596-838) executed on one image: its own loop over the table
    nodes -- the ijk -> (direction, stability, speed) mapping, calc_u / calc_v, exp(nsq), the layer bounds, the generic
    linear_perturbation, the staggering, copy_data_to_remote (a coarray PUT) -- fills the whole look-up table; every entry
    equals what oracle/spatial_winds_np.py's EntryCache computes on demand (which is what icar_linear_perturb_spatial
    evaluates on the device), terrain-following and constant layer bounds"""
    from oracle import spatial_winds_np as SW
    f = np.float32
    nx, ny, nz = 7, 6, 2
    o = SW.LTOptions()
    o.buffer, o.n_dir_values, o.n_spd_values, o.n_nsq_values = 3, 5, 3, 3
    terrain = (400 * np.exp(-(((np.arange(nx)[:, None] - 3) / 2.0) ** 2 + ((np.arange(ny)[None, :] - 2) / 1.5) ** 2))).astype(f)
    rng = np.random.default_rng(2)
    if varying:
        zi = np.zeros((nx, nz, ny), f)
        dzi = np.zeros((nx, nz, ny), f)
        zi[:, 0, :] = terrain
        dzi[:, 0, :] = rng.uniform(250, 350, (nx, ny))
        zi[:, 1, :] = zi[:, 0, :] + dzi[:, 0, :]
        dzi[:, 1, :] = rng.uniform(400, 600, (nx, ny))
        lu, lv = ref.spatial_winds_table(terrain, 2000.0, o, z_interface=zi, dz_interface=dzi)

        def bounds(z):
            zb = zi[:, z, :] - terrain
            return zb, zb + dzi[:, z, :]
    else:
        dzl = np.array([300, 500], f)
        lh = np.array([150, 550], f)                                  # mid-level heights above the terrain: sum(dz) - dz/2
        lu, lv = ref.spatial_winds_table(terrain, 2000.0, o, layer_height=lh, dz_levels=dzl)

        def bounds(z):
            return lh[z] - dzl[z] / f(2), lh[z] + dzl[z] / f(2)
    cache = SW.EntryCache(terrain, 2000.0, o, bounds)
    assert lu.shape == (3, 5, 3, nx + 1, nz, ny) and lv.shape == (3, 5, 3, nx, nz, ny + 1)
    for s_ in range(3):
        for d in range(5):
            for n in range(3):
                for z in range(nz):
                    tu, tv = cache.get(s_, d, n, z)
                    assert np.array_equal(tu.T, lu[s_, d, n, :, z, :]) and np.array_equal(tv.T, lv[s_, d, n, :, z, :]), (s_, d, n, z)
    assert np.abs(lu).max() > 1.0 and (lu[0] == 0).all()             # speed node 0 is calm: no perturbation
