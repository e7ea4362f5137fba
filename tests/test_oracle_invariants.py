"""Size-independent properties of the algorithm, checked on the oracle (CPU).  The same
properties are re-checked on the CUDA path in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from icar_b200 import ideal


def closed_domain_case(nx, nz, ny, dtype, seed=0):
    """Set-A metric (jacobian 1, uniform dz): zero normal wind on every domain wall, horizontally
    non-divergent-free random interior wind; mass sum(q*dz) over the interior telescopes."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(0.5, 2.0, (ny, nz, nx)).astype(dtype)
    U = rng.uniform(-0.2, 0.2, (ny, nz, nx - 1)).astype(dtype)
    V = rng.uniform(-0.2, 0.2, (ny - 1, nz, nx)).astype(dtype)
    W = rng.uniform(-20, 20, (ny, nz, nx)).astype(dtype)
    # no flux between the boundary-condition ring and the interior, none through the top
    U[:, :, 0] = 0; U[:, :, nx - 2] = 0
    V[0] = 0; V[ny - 2] = 0
    W[:, nz - 1, :] = 0
    jaco = np.ones((ny, nz, nx), dtype)
    dz = np.full((ny, nz, nx), 400.0, dtype)
    return q, U, V, W, jaco, dz


def interior_mass(q, dz):
    return float((q[1:-1, :, 1:-1].astype(np.float64) * dz[1:-1, :, 1:-1]).sum())


@pytest.mark.parametrize("order,fct", [(1, True), (2, True), (2, False)])
def test_mass_conservation_f64(oracle, order, fct):
    q, U, V, W, jaco, dz = closed_domain_case(14, 9, 12, np.float64)
    m0 = interior_mass(q, dz)
    for _ in range(5):
        oracle.advect3d(q, None, jaco, dz, U, V, W, order, fct)
    assert abs(interior_mass(q, dz) - m0) / m0 < 1e-12


def test_mass_conservation_f32(oracle):
    q, U, V, W, jaco, dz = closed_domain_case(14, 9, 12, np.float32)
    m0 = interior_mass(q, dz)
    for _ in range(5):
        oracle.advect3d(q, None, jaco, dz, U, V, W, 2, True)
    assert abs(interior_mass(q, dz) - m0) / m0 < 5e-6


def test_fct_monotone(oracle):
    """1-D y advection of a step (the test_mpdata.f90 scenario): no new extrema with FCT."""
    nx, nz, ny = 3, 3, 60
    q = np.ones((ny, nz, nx), np.float32); q[ny // 2:] = 2
    U = np.zeros((ny, nz, nx - 1), np.float32); V = np.full((ny - 1, nz, nx), 0.25, np.float32)
    W = np.zeros((ny, nz, nx), np.float32)
    one = np.ones_like(q)
    qf = q.copy()
    for _ in range(40):
        oracle.advect3d(qf, None, one, one, U, V, W, 2, True)
    assert qf.max() <= 2.0 and qf.min() >= 1.0
    qn = q.copy()
    for _ in range(40):
        oracle.advect3d(qn, None, one, one, U, V, W, 2, False)
    assert qn.max() > 2.0 or qn.min() < 1.0      # without FCT the corrector over/undershoots


def test_order1_is_upwind(oracle):
    rng = np.random.default_rng(3)
    st = ideal.make_ideal_case(nx=16, nz=8, ny=12)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"],
                                    st["jacobian_w"], None, 30.0, st["dx"], scheme=2)
    U1, V1, W1 = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"],
                                       st["jacobian_w"], None, 30.0, st["dx"], scheme=1)
    # rho == 1: adv_mpdata.f90:500 and advect.f90:344 differ only in factor order
    assert np.array_equal(U, U1) and np.array_equal(V, V1) and np.array_equal(W, W1)
    q = rng.uniform(0, 1, st["th"].shape).astype(np.float32)
    a = q.copy()
    oracle.advect3d(a, None, st["jacobian"], st["advection_dz"], U, V, W, 1, True)
    b = oracle.upwind_advection(q, U, V, W, None, st["advection_dz"], st["jacobian"])
    assert np.array_equal(a, b)


def test_balanced_wind_keeps_uniform_field(oracle):
    """With w from balance_uvw the normalised wind is discretely non-divergent, so a uniform
    scalar stays uniform under the donor-cell pass (up to rounding) on the hill case."""
    st = ideal.make_ideal_case(nx=20, nz=8, ny=10)
    dt = st["dt"]
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"],
                                    st["jacobian_w"], None, dt, st["dx"])
    q = np.full(st["th"].shape, 3.0, np.float32)
    oracle.advect3d(q, None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
    assert np.abs(q[1:-1, :-1, 1:-1] - 3.0).max() < 2e-5


def test_mp_simple_water_budget(oracle):
    """Column water (vapour + cloud + rain + snow, mass weighted) + fallen precipitation is
    conserved by mp_simple; dry columns are left untouched."""
    st = ideal.make_ideal_case(nx=12, nz=10, ny=8)
    st["qv"] *= np.float32(3.0)
    nx, nz, ny = st["nx"], st["nz"], st["ny"]
    w0 = ((st["qv"] + st["qc"] + st["qr"] + st["qs"]).astype(np.float64) * st["rho"] * st["dz_mass"]).sum(axis=1)
    th0 = st["th"].copy()
    diag = oracle.mp_simple_driver(st["pressure"], st["th"], st["pii"], st["rho"], st["qv"], st["qc"],
                                   st["qr"], st["qs"], st["rain"], st["snow"], 60.0, st["dz_mass"],
                                   1, nx - 2, 1, ny - 2, 0, nz - 1)
    assert diag == (0, 0, 0)
    w1 = ((st["qv"] + st["qc"] + st["qr"] + st["qs"]).astype(np.float64) * st["rho"] * st["dz_mass"]).sum(axis=1)
    assert st["qc"].max() > 0 and st["rain"].max() > 0          # the case is microphysically active
    lhs = w1[1:-1, 1:-1] + st["rain"][1:-1, 1:-1]
    assert np.abs(lhs - w0[1:-1, 1:-1]).max() / w0.max() < 2e-5
    # boundary-condition ring is never processed (its = ims+1 ..., grid_obj.f90:241-250)
    assert np.array_equal(st["th"][0], th0[0]) and np.array_equal(st["th"][:, :, 0], th0[:, :, 0])


def test_compute_dt_restatement_matches_the_ideal_case_generator():
    """oracle/np_restatement.compute_dt_np (time_step.f90:217-330, all strictness levels) against the
    strictness-3 formula icar_b200.ideal uses to pick the benchmark's dt."""
    from oracle import np_restatement as NP
    st = ideal.make_ideal_case(nx=30, nz=10, ny=12, perturb=True)
    want = ideal.compute_dt(st["u"], st["v"], st["w"], st["dx"], st["dz_levels"])
    assert NP.compute_dt_np(st["dx"], st["u"], st["v"], st["w"], st["dz_levels"], 0.9, 3) == pytest.approx(want, rel=1e-7)
    # strictness 4 = strictness 3 / (sqrt(3) * 1.001); 2 <= 3 (mean instead of sum, but at least the 1-D limit)
    d3 = NP.compute_dt_np(st["dx"], st["u"] * 1e-3, st["v"] * 1e-3, st["w"] * 1e-3, st["dz_levels"], 0.9, 3)
    d4 = NP.compute_dt_np(st["dx"], st["u"] * 1e-3, st["v"] * 1e-3, st["w"] * 1e-3, st["dz_levels"], 0.9, 4)
    assert d3 == 120.0 and d4 == 120.0          # capped (time_step.f90:421)
    with pytest.raises(FloatingPointError):     # strictness 1 divides by a speed, not speed/dx (:241-249)
        NP.compute_dt_np(st["dx"], st["u"], st["v"], st["w"], st["dz_levels"], 0.9, 1)
    ex, rho = NP.diagnostic_update_np(st["pressure"], st["th"])
    np.testing.assert_allclose(ex, st["pii"], rtol=3e-7)
    np.testing.assert_allclose(rho, st["rho"], rtol=5e-7)
