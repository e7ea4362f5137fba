"""The "next" rows around the hot path: diagnostic_update's exner / density,
update_dt / compute_dt and enforce_limits on the device-resident fields, against the NumPy restatement."""
import numpy as np
import pytest

from icar_b200 import ideal
from oracle import np_restatement as NP

pytestmark = pytest.mark.gpu


def make(gpu_lib, st, **opt):
    nx, nz, ny = st["nx"], st["nz"], st["ny"]
    dom = gpu_lib.Domain(gpu_lib.Grid.single_image(nx, nz, ny), gpu_lib.Options(dx=st["dx"], **opt))
    dom.upload_state(st)
    return dom


def test_diagnostic_update_matches_time_step_f90(gpu_lib):
    st = ideal.make_ideal_case(nx=70, nz=15, ny=11, perturb=True)
    dom = make(gpu_lib, st)
    dom.upload("exner", np.zeros_like(st["pii"]))
    dom.upload("density", np.zeros_like(st["rho"]))
    dom.diagnostic_update()
    ex, rho = NP.diagnostic_update_np(st["pressure"], st["th"])
    # powf of the device vs glibc: <= 2 ulp
    np.testing.assert_allclose(dom.download("exner"), ex, rtol=3e-7)
    np.testing.assert_allclose(dom.download("density"), rho, rtol=5e-7)
    # and the ideal-case generator agrees with both
    np.testing.assert_allclose(ex, st["pii"], rtol=3e-7)


def test_diagnostic_output_matches_time_step_f90(gpu_lib):
    """The output-only fields of diagnostic_update (time_step.f90:88-108, 165-194): temperature, the interface fields,
    psfc, u_mass, v_mass, w_real -- the same REAL operations in the same order, so equal bit for bit given exner."""
    st = ideal.make_ideal_case(nx=70, nz=15, ny=11, perturb=True)
    rng = np.random.default_rng(5)
    st["v"] = (st["v"] + rng.uniform(-3, 3, st["v"].shape)).astype(np.float32)
    dzdx, dzdy = ideal.terrain_slopes(st)
    dzdy = (dzdy + rng.uniform(-0.05, 0.05, dzdy.shape)).astype(np.float32)        # the hill is uniform in y
    dom = make(gpu_lib, st)
    dom.diagnostic_update()
    got = dom.diagnostic_output(dzdx, dzdy)
    want = NP.diagnostic_output_np(st["pressure"], st["th"], st["u"], st["v"], st["w"], st["jacobian"], dzdx, dzdy)
    for k in ("pressure_interface", "surface_pressure", "u_mass", "v_mass", "w_real"):
        assert np.array_equal(got[k], want[k]), k
    assert np.abs(want["w_real"]).max() > 0.1
    # temperature goes through the device's powf (exner): <= 2 ulp; from the device's own exner it is exact
    ex = dom.download("exner")
    t = st["th"] * ex
    assert np.array_equal(got["temperature"], t)
    ti = np.empty_like(t)
    ti[:, 1:, :] = (t[:, :-1, :] + t[:, 1:, :]) / np.float32(2)
    ti[:, 0, :] = t[:, 0, :] + (t[:, 0, :] - t[:, 1, :]) / np.float32(2)
    assert np.array_equal(got["temperature_interface"], ti)
    np.testing.assert_allclose(got["temperature"], want["temperature"], rtol=3e-7)
    # a subset of the outputs, and w_real without the slopes is refused
    only = dom.diagnostic_output(want=("u_mass",))
    assert list(only) == ["u_mass"] and np.array_equal(only["u_mass"], want["u_mass"])
    with pytest.raises(gpu_lib.IcarError):
        dom.diagnostic_output(want=("w_real",))


@pytest.mark.parametrize("strictness", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("scale", [1.0, 0.01])
def test_compute_dt_all_strictness_levels(gpu_lib, strictness, scale):
    """cfl_strictness 1, 5 (and the 1-D branch of 2) divide the CFL factor by a wind SPEED, not by
    speed/dx (time_step.f90:241-258,296-305), so with ordinary winds the reference stops with "time step
    too small"; the library must fail the same way, and agree exactly where the reference succeeds."""
    st = ideal.make_ideal_case(nx=53, nz=12, ny=17, perturb=True)
    rng = np.random.default_rng(strictness)
    st["u"] = (st["u"] * rng.uniform(0.2, 1.5, st["u"].shape) * scale).astype(np.float32)
    st["v"] = (st["v"] * scale).astype(np.float32)
    st["w"] = (rng.uniform(-2, 2, st["w"].shape) * scale).astype(np.float32)
    dom = make(gpu_lib, st)
    try:
        want = NP.compute_dt_np(st["dx"], st["u"], st["v"], st["w"], st["dz_levels"], 0.9, strictness)
    except FloatingPointError:
        with pytest.raises(gpu_lib.IcarError) as e:
            dom.compute_dt(st["dz_levels"], 0.9, strictness)
        assert e.value.code == gpu_lib.ERR_PHYSICS
        return
    got = dom.compute_dt(st["dz_levels"], 0.9, strictness)
    assert got == want, (got, want)          # maxima are exact, the arithmetic is the same REAL sequence


def test_compute_dt_cap_and_failure(gpu_lib):
    st = ideal.make_ideal_case(nx=20, nz=8, ny=9)
    st["u"][:] = np.float32(0.5); st["v"][:] = 0; st["w"][:] = 0
    dom = make(gpu_lib, st)
    assert dom.compute_dt(st["dz_levels"]) == 120.0                      # min(seconds, 120), time_step.f90:421
    st["u"][:] = np.float32(5e5)
    dom.upload("u", st["u"])
    with pytest.raises(gpu_lib.IcarError) as e:                          # "ERROR time step too small", :322-328
        dom.compute_dt(st["dz_levels"])
    assert e.value.code == gpu_lib.ERR_PHYSICS and "time step too small" in str(e.value)


@pytest.mark.parametrize("mp", [2, 1])
def test_enforce_limits_clamps_every_scheme_scalar(gpu_lib, mp):
    st = ideal.make_ideal_case(nx=40, nz=8, ny=10, perturb=True)
    names = ["th", "qv", "qc", "qr", "qs"]
    if mp == 1:
        ideal.add_thompson_fields(st)
        names += ["qi", "qg", "ni", "nr"]
    rng = np.random.default_rng(0)
    for k in names:
        st[k] = (st[k] + rng.normal(0, 1e-3, st[k].shape)).astype(np.float32) if k != "th" else st[k]
        st[k][rng.random(st[k].shape) < 0.2] *= np.float32(-1)
    dom = make(gpu_lib, st, microphysics=mp)
    dom.enforce_limits()
    want = [st[k].copy() for k in names]
    NP.enforce_limits_np(want)
    for k, w in zip(names, want):
        np.testing.assert_array_equal(dom.download(k), w, err_msg=k)


def test_apply_forcing_diagnosed_fields_and_scalar_boundaries(gpu_lib):
    st = ideal.make_ideal_case(nx=37, nz=9, ny=13, perturb=True)
    rng = np.random.default_rng(5)
    dom = make(gpu_lib, st)
    names = ("u", "v", "w", "pressure", "th", "qv")
    tend = {k: (rng.normal(0, 1, st[k].shape) * np.abs(st[k]).max() * 1e-4).astype(np.float32) for k in names}
    for k in names:
        dom.upload_tendency(k, tend[k])
    want = {k: st[k].copy() for k in names}
    for seconds in (117.3, 36.0):
        dom.apply_forcing(seconds)
        for k in names:
            NP.apply_forcing_np(want[k], tend[k], seconds, boundaries_only=k in ("th", "qv"))
    for k in names:
        np.testing.assert_array_equal(dom.download(k), want[k], err_msg=k)
    # interior of a boundary-forced scalar is untouched, its edges are not
    assert np.array_equal(want["qv"][1:-1, :, 1:-1], st["qv"][1:-1, :, 1:-1])
    assert not np.array_equal(want["qv"][0], st["qv"][0])
    # a scalar without a tendency is left alone
    np.testing.assert_array_equal(dom.download("qc"), st["qc"])


def test_balance_uvw_matches_wind_f90(gpu_lib):
    """icar_balance_uvw against the NumPy restatement of wind.f90:81-228 the ideal-case generator uses."""
    st = ideal.make_ideal_case(nx=45, nz=14, ny=11, perturb=True)
    rng = np.random.default_rng(2)
    st["u"] = (st["u"] * rng.uniform(0.5, 1.5, st["u"].shape)).astype(np.float32)
    st["v"] = (st["v"] * rng.uniform(0.5, 1.5, st["v"].shape)).astype(np.float32)
    dom = make(gpu_lib, st)
    dom.upload("w", np.full_like(st["w"], 7.0))            # balance_uvw starts from w = 0 (:104)
    dom.balance_uvw()
    want = ideal.balance_uvw(st["u"], st["v"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"],
                             st["advection_dz"], st["dx"])
    got = dom.download("w")
    # same REAL operations in the same order; -fmad=false on the device, NumPy does not contract either
    np.testing.assert_array_equal(got, want)


def test_device_resident_loop_body(gpu_lib, oracle):
    """The loop body of step() (time_step.f90:464-547 with rad/lsm/pbl off): update_dt -> diagnostic_update ->
    mp -> advect -> apply_forcing -> enforce_limits entirely on the device, a few steps, against the same
    sequence on the CPU (exact arithmetic mode)."""
    st = ideal.make_ideal_case(nx=48, nz=16, ny=20, perturb=True)
    st["qv"] *= np.float32(3.0)
    ref = ideal.copy_state(st)
    dom = make(gpu_lib, st, exact_arithmetic=1)
    rng = np.random.default_rng(11)
    tend = {"u": (rng.normal(0, 2e-3, st["u"].shape)).astype(np.float32),
            "pressure": (rng.normal(0, 1e-2, st["pressure"].shape)).astype(np.float32),
            "th": (rng.normal(0, 1e-4, st["th"].shape)).astype(np.float32)}
    for k, v in tend.items():
        dom.upload_tendency(k, v)
    for _ in range(3):
        dt = dom.compute_dt(st["dz_levels"])
        dom.diagnostic_update()
        dom.physics_step(dt)
        dom.apply_forcing(dt)
        dom.enforce_limits()
        dt_ref = NP.compute_dt_np(ref["dx"], ref["u"], ref["v"], ref["w"], ref["dz_levels"], 0.9, 3)
        assert dt == dt_ref
        ref["pii"], ref["rho"] = NP.diagnostic_update_np(ref["pressure"], ref["th"])
        oracle.step_simple(ref, np.float32(dt_ref), np.float32(dt_ref), 2, True)
        for k, v in tend.items():
            NP.apply_forcing_np(ref[k], v, dt_ref, boundaries_only=(k == "th"))
        NP.enforce_limits_np([ref[k] for k in ("th", "qv", "qc", "qr", "qs")])
    for k in ("th", "qv", "qc", "qr", "qs"):
        a, b = dom.download(k).astype(np.float64), ref[k].astype(np.float64)
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), k         # exner via device powf feeds mp_simple


def test_step_is_the_loop_of_time_step_f90(gpu_lib):
    """icar_step == the do-while loop of step() (time_step.f90:464-549) written out with the calls it is made of: update_dt,
    the clip to end_time, diagnostic_update, the physics of the step with real(dt), apply_forcing(dt), enforce_limits in
    the last two updates, model_time += dt in REAL*8 -- bit for bit and in the same number of passes; a pass whose dt
    is below 1e-3 s only advances the clock (:484)."""
    st = ideal.make_ideal_case(nx=48, nz=16, ny=20, perturb=True)
    st["qv"] *= np.float32(3.0)
    st["qc"][:, 10:, :] = np.float32(-1e-5)             # something for enforce_limits to clamp
    rng = np.random.default_rng(11)
    tend = {"u": (rng.normal(0, 2e-3, st["u"].shape)).astype(np.float32),
            "pressure": (rng.normal(0, 1e-2, st["pressure"].shape)).astype(np.float32),
            "th": (rng.normal(0, 1e-4, st["th"].shape)).astype(np.float32)}
    a, b = make(gpu_lib, st), make(gpu_lib, st)
    for d in (a, b):
        for k, v in tend.items():
            d.upload_tendency(k, v)
        d.model_time = 1000.0
    passes = []                                          # (dt, physics ran, enforce_limits ran) of the loop written out

    def by_hand(d, end):
        n = 0
        while d.model_time < end:
            dt = d.compute_dt(st["dz_levels"])
            if d.model_time + dt > end:
                dt = end - d.model_time
            d.diagnostic_update()
            t0 = d.model_time
            physics = limits = False
            if dt > 1e-3:
                d.physics_step(dt)                       # real(dt%seconds())
                d.model_time = t0
                d.apply_forcing(dt)
                physics = True
                if (end - t0) < dt * 2:
                    d.enforce_limits()
                    limits = True
            d.model_time = t0 + dt
            passes.append((dt, physics, limits))
            n += 1
        return n

    def same_state():
        assert a.model_time == b.model_time
        for k in ("th", "qv", "qc", "qr", "qs", "u", "pressure", "exner", "density"):
            np.testing.assert_array_equal(a.download(k), b.download(k), err_msg=k)

    dt0 = b.compute_dt(st["dz_levels"])
    end = 1000.0 + 2.4 * dt0                             # two full steps and a clipped one
    n = a.step(end, st["dz_levels"])
    assert n == by_hand(b, end) and n >= 3
    assert abs(a.model_time - end) < 1e-9
    same_state()
    assert sum(p[2] for p in passes) >= 2 and not passes[0][2]      # enforce_limits: only in the last updates
    assert (a.download("qc") >= 0).all()
    # one more full step and half a millisecond: the last pass must not run any physics
    end = a.model_time + b.compute_dt(st["dz_levels"]) + 0.0005
    first = len(passes)
    n = a.step(end, st["dz_levels"])
    assert n == by_hand(b, end) and n == 2
    same_state()
    assert passes[first][1] and not passes[-1][1] and passes[-1][0] < 1e-3


def test_step_host_derives_exner_and_density_on_the_device(gpu_lib):
    """icar_step_host == upload; icar_diagnostic_update; icar_physics_step; download -- bit for bit -- and it
    refuses host exner / density."""
    st = ideal.make_ideal_case(nx=48, nz=16, ny=20, perturb=True)
    st["qv"] *= np.float32(3.0)
    for k in ("precipitation", "snowfall"):
        st[k] = np.zeros((st["ny"], st["nx"]), np.float64)
    dt = float(st["dt"])
    ref = make(gpu_lib, st)
    ref.diagnostic_update()
    ref.physics_step(dt)
    dom = make(gpu_lib, st)
    host = {k: (np.ascontiguousarray(v.copy()) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    with pytest.raises(gpu_lib.IcarError):
        dom.step_host(dt, dom.host_domain(host))                 # exner / density present
    host = {k: v for k, v in host.items() if k not in ("pii", "rho")}
    dom.step_host(dt, dom.host_domain(host))
    for k in ("th", "qv", "qc", "qr", "qs", "precipitation"):
        np.testing.assert_array_equal(host[k], ref.download(k), err_msg=k)


@pytest.mark.parametrize("bands,derive", [(3, False), (5, True), (1, False)])
def test_host_pipeline_equals_the_undivided_host_step(gpu_lib, bands, derive):
    """icar_b200.pipeline.HostPipeline: bands of rows as stand-alone tiles with 4 ghost rows, stepped
    concurrently -- the rows each band owns must equal the undivided domain's, bit for bit."""
    from icar_b200.pipeline import HostPipeline
    st = ideal.make_ideal_case(nx=64, nz=12, ny=70, perturb=True)
    st["qv"] *= np.float32(3.0)
    for k in ("precipitation", "snowfall"):
        st[k] = np.zeros((st["ny"], st["nx"]), np.float64)
    dt = float(st["dt"])
    per_step = ("u", "v", "w", "pressure", "pii", "rho", "th", "qv", "qc", "qr", "qs", "precipitation", "snowfall")
    out_keys = ("th", "qv", "qc", "qr", "qs", "precipitation", "snowfall")
    # reference: one domain, in-place host arrays
    ref_host = {k: np.ascontiguousarray(st[k].copy()) for k in per_step}
    ref = make(gpu_lib, st)
    hd = ref.host_domain(ref_host)
    # pipeline: two sets of host arrays, swapped every step
    a = {k: np.ascontiguousarray(st[k].copy()) for k in per_step}
    b = {k: np.ascontiguousarray(st[k].copy()) for k in per_step}
    pipe = HostPipeline(st, gpu_lib.Options(dx=st["dx"]), bands=bands, derive_diagnostics=derive)
    for step in range(3):
        if derive:
            ref.upload_state({k: ref_host[k] for k in ("pressure", "th")})
            ref.diagnostic_update()
            ref_in = {k: v for k, v in ref_host.items() if k not in ("pii", "rho")}
            ref.physics_step_host(dt, ref.host_domain(ref_in))
        else:
            ref.physics_step_host(dt, hd)
        src, dst = (a, b) if step % 2 == 0 else (b, a)
        pipe.step(dt, src, {k: dst[k] for k in out_keys})
        for k in out_keys:
            np.testing.assert_array_equal(dst[k], ref_host[k], err_msg=f"step {step} {k}")
    with pytest.raises(ValueError):
        pipe.step(dt, a, {k: a[k] for k in out_keys})
    pipe.close()
