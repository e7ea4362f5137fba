"""oracle/linear_winds_np.py (linear-theory wind perturbation) on the CPU.
The restatement equals the reference's own linear_winds.f90 executed by oracle/f90run (tests/test_reference_source.py);
these tests check it against the conventions it must reproduce, against the analytic hydrostatic mountain-wave
solution and against the reference's Python solver (no expected values exist in the reference)."""
import numpy as np
import pytest

from oracle import linear_winds_np as LW


@pytest.mark.parametrize("n", [8, 9, 15, 16])
def test_fftshift_pair_moves_by_half_rounded_up_and_inverts(n):
    """fftshift.f90: element i goes to mod(i + (n+1)/2, n): a rotation by (n+1)/2 -- MATLAB's fftshift for an
    even length, its ifftshift for an odd one -- and ifftshift undoes it."""
    rng = np.random.default_rng(n)
    a = (rng.standard_normal((n, n + 3)) + 1j * rng.standard_normal((n, n + 3))).astype(np.complex64).astype(np.complex128)
    want = np.roll(np.roll(a, (n + 1) // 2, axis=0), (n + 4) // 2, axis=1)
    np.testing.assert_array_equal(LW.fftshift2cc(a), want)
    if n % 2 == 0:
        np.testing.assert_array_equal(LW.fftshift2cc(a[:, :n]), np.fft.fftshift(a[:, :n]))
    np.testing.assert_array_equal(LW.ifftshift2cc(LW.fftshift2cc(a)), a)
    # the shifts go through a single-precision temporary (fftshift.f90:91): double input loses its low bits
    assert np.array_equal(LW.fftshift2cc(a + 1e-12), want)


def test_wavenumber_grids():
    lt = LW.LinearTheoryData(21, 11, 4000.0)
    assert lt.k.dtype == np.float32 and lt.k.shape == (21, 11)
    np.testing.assert_allclose(lt.k[0, :], -np.pi / 4000.0, rtol=1e-6)               # -pi/dx .. +pi/dx inclusive
    np.testing.assert_allclose(lt.k[-1, :], np.pi / 4000.0, rtol=1e-6)
    np.testing.assert_allclose(lt.l[:, 0], -np.pi / 4000.0, rtol=1e-6)
    assert abs(lt.k[10, 0]) < 1e-9 and abs(lt.l[0, 5]) < 1e-9
    assert lt.kl[10, 5] == LW.SMALL_VALUE and (lt.kl > 0).all()


def test_buffered_terrain_is_periodic_and_keeps_the_interior():
    rng = np.random.default_rng(3)
    x = np.arange(40)[:, None]
    terrain = (1000 * np.exp(-((x - 20) / 6.0) ** 2) + 50 * rng.random((40, 30)) + 10 * x).astype(np.float32)
    buf = 6
    b = LW.add_buffer_topo(terrain, 0, buf)
    assert b.shape == (52, 42)
    np.testing.assert_array_equal(b[buf:-buf, buf:-buf].real, terrain)
    # the two halves of the rim meet half way between the opposite edges: no jump across the wrap-around
    np.testing.assert_allclose(b[0, buf:-buf].real, 0.5 * (terrain[0] + terrain[-1]), rtol=1e-6)
    np.testing.assert_allclose(b[-1, buf:-buf].real, 0.5 * (terrain[0] + terrain[-1]), rtol=1e-6)
    np.testing.assert_allclose(b[:, 0].real, b[:, -1].real, rtol=1e-6)
    smooth = LW.add_buffer_topo(terrain, 5, buf)
    np.testing.assert_array_equal(smooth[buf:-buf, buf:-buf].real, terrain)          # only rim cells are smoothed
    rough = lambda a: np.abs(np.diff(a[:, :buf].real, axis=0)).mean()
    assert rough(smooth) < rough(b)
    tf, total = LW.terrain_frequency(terrain, buf)
    assert total == buf + 2 and tf.shape == (56, 46)
    centre = tf[np.ix_(LW._shift_index(56), LW._shift_index(46))][0, 0]              # the mean, moved to the middle
    second = LW.add_buffer_topo(LW.add_buffer_topo(terrain, 5, buf).real.astype(np.float32), 0, 2)
    assert abs(centre - second.mean()) < 1e-3 * abs(second.mean())


def ridge_case(nx=64, ny=16, dx=2000.0, h0=100.0, waves=4):
    """A sinusoidal ridge h = h0 cos(2 pi waves x / nx), uniform in y, and its spectrum prepared the way
    setup_linwinds does it (forward FFT / N, fftshift) -- without the buffer, the field is already periodic."""
    lt = LW.LinearTheoryData(nx, ny, dx)
    x = np.arange(nx)[:, None]
    terrain = h0 * np.cos(2 * np.pi * waves * x / nx) * np.ones((1, ny))
    ft = LW.fftshift2cc(np.fft.fft2(terrain) / (nx * ny))
    return lt, ft, terrain


def test_no_wind_no_perturbation_and_linearity():
    lt, ft, _ = ridge_case(nx=65, ny=17)
    u, v = LW.linear_perturbation_at_height(0.0, 0.0, 1e-4, 500.0, ft, lt)
    assert not u.any() and not v.any()
    u1, v1 = LW.linear_perturbation_at_height(10.0, 3.0, 1e-4, 500.0, ft, lt)
    u2, v2 = LW.linear_perturbation_at_height(10.0, 3.0, 1e-4, 500.0, 2.5 * ft, lt)
    np.testing.assert_allclose(u2, 2.5 * u1, rtol=1e-5, atol=1e-9)                    # linear in the terrain
    np.testing.assert_allclose(v2, 2.5 * v1, rtol=1e-5, atol=1e-9)


def test_hydrostatic_mountain_wave_amplitude_and_phase_lines():
    """For h = h0 cos(k0 x), wind U along x and l = 0 the solution is u' = N h0 sin(k0 x + m z), m = N / U
    (Barstad & Gronas 2006, eq. A1-A4 with f = 0): amplitude N h0 at every height, phase tilting with height.
    The reference's wavenumber grid sits half a cell off the FFT's (k, l run over -pi/dx..pi/dx INCLUSIVE, and
    fftshift puts the mean where l = +gain/2), so a ridge that is uniform in y still sees l / k = 0.5 / 5.5 and 0.5 / 4.5 here:
    amplitude and vertical wavenumber are off by l^2 / k^2 ~ 1 % and a v' of l / k times u' appears.  The
    restatement has to reproduce that, so the tolerances are those of the quirk, not of the arithmetic."""
    nx, ny, dx, h0, U, Nsq = 64, 64, 2000.0, 100.0, 10.0, 1e-4
    lt, ft, terrain = ridge_case(nx, ny, dx, h0, waves=5)
    N = np.sqrt(Nsq)
    amps, phases = [], []
    for z in (0.0, 250.0, 500.0, 1000.0):
        u, v = LW.linear_perturbation_at_height(U, 0.0, Nsq, z, ft, lt)
        line = u[:, 3].real
        assert np.abs(u.imag).max() < 0.02 * N * h0                                   # an (almost) real field comes back
        assert np.abs(v).max() < (0.5 / 4.5) * N * h0                                 # the half-cell offset in l (bins +5.5, -4.5)
        np.testing.assert_allclose(u, u[:, :1] * np.ones((1, ny)), atol=1e-9)         # uniform in y like the ridge
        amps.append(np.abs(line).max())
        phases.append(np.angle(np.fft.fft(line)[5]))
    np.testing.assert_allclose(amps, N * h0, rtol=2e-2)
    d = np.diff(np.unwrap(phases))                                                    # m dz = 0.25, 0.25, 0.5 rad
    np.testing.assert_allclose(np.abs(d), [0.25, 0.25, 0.5], rtol=3e-2)
    assert (np.sign(d) == np.sign(d[0])).all()
    # at the ground the perturbation is in quadrature with the terrain: u' ~ sin where h ~ cos
    u0 = LW.linear_perturbation_at_height(U, 0.0, Nsq, 0.0, ft, lt)[0][:, 3].real
    assert abs(np.dot(u0, terrain[:, 3])) < 0.05 * np.linalg.norm(u0) * np.linalg.norm(terrain[:, 3])


def test_evanescent_modes_decay_with_height():
    """Short waves in weak stability (N^2 kl / sig^2 ... msq < 0 is forced with a negative N^2): the reference
    takes m = +i sqrt(-msq) for either sign of sig, so the response decays upward."""
    lt, ft, _ = ridge_case(nx=65, ny=17, waves=20)
    a0 = np.abs(LW.linear_perturbation_at_height(10.0, 0.0, -1e-4, 0.0, ft, lt)[0]).max()
    a1 = np.abs(LW.linear_perturbation_at_height(10.0, 0.0, -1e-4, 2000.0, ft, lt)[0]).max()
    a2 = np.abs(LW.linear_perturbation_at_height(-10.0, 0.0, -1e-4, 2000.0, ft, lt)[0]).max()
    assert a1 < 0.5 * a0 and abs(a2 - a1) < 1e-6 * a0


def test_layer_means():
    lt, ft, _ = ridge_case(nx=65, ny=17, waves=5)
    args = (10.0, 2.0, 1e-4)
    u, v = LW.linear_perturbation_constz(*args, 200.0, 800.0, 100.0, ft, lt)
    samples = [LW.linear_perturbation_at_height(*args, z, ft, lt) for z in (250.0, 350.0, 450.0, 550.0, 650.0, 750.0)]
    np.testing.assert_allclose(u, sum(s[0] for s in samples) / 6, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(v, sum(s[1] for s in samples) / 6, rtol=1e-6, atol=1e-12)
    # a layer whose bounds do not vary in space: varyingz equals constz on the columns the bounds were given for
    zb = np.full((57, 9), 200.0, np.float32)
    zt = np.full((57, 9), 800.0, np.float32)
    uv, vv = LW.linear_perturbation_varyingz(*args, zb, zt, 100.0, ft, lt, buffer=4)
    np.testing.assert_allclose(uv, u, rtol=1e-5, atol=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# Pinned against the reference itself: tests/golden/linear_theory_1d.npz holds u_hat of the reference's own Python
# implementation of the same solution, helpers/ideal_linear.py:solve(), run in this container by
# tests/golden/make_linear_theory_golden.py.  It is the x-z form (l = 0) of linear_perturbation_at_height.

def one_dimensional(golden, n):
    """The case as linear_perturbation_at_height sees it: an (Nx, 1) plane with l = 0."""
    f = np.float32
    Fzs, U, Nsq, dx = golden[f"Fzs_{n}"], float(golden[f"U_{n}"]), float(golden[f"Ndsq_{n}"]), float(golden["dx"])
    Nx = Fzs.shape[0]
    lt = LW.LinearTheoryData.__new__(LW.LinearTheoryData)
    k1 = LW.LinearTheoryData(Nx, 2, dx).k[:, :1]                     # the Fortran's REAL k grid (:447-451)
    np.testing.assert_allclose(k1[:, 0], np.linspace(-np.pi / dx, np.pi / dx, Nx), rtol=2e-6, atol=1e-10)   # == solve()'s k
    lt.k, lt.l = k1, np.zeros((Nx, 1), f)
    lt.kl = (k1 ** 2).astype(f)
    lt.kl[lt.kl == 0] = LW.SMALL_VALUE
    lt.nx, lt.ny = Nx, 1
    return lt, Fzs[:, None].astype(np.complex128), U, Nsq


@pytest.fixture(scope="module")
def golden_1d():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linear_theory_1d.npz"))


def test_restatement_reproduces_the_references_python_solver(golden_1d):
    g = golden_1d
    for n in range(int(g["ncases"])):
        lt, ft, U, Nsq = one_dimensional(g, n)
        for iz, z in enumerate(g["zlevels"]):
            u, v = LW.linear_perturbation_at_height(U, 0.0, Nsq, z, ft, lt)
            want = g[f"u_hat_{n}"][iz]
            scale = np.abs(want).max()
            assert scale > 0
            # REAL k, sig and the single-precision temporary of ifftshift against float64 throughout in solve()
            assert np.abs(u[:, 0].real - want).max() <= 2e-7 * scale, (n, z)        # measured 4.5e-8
            assert not v.any()                                         # l = 0: no v'
