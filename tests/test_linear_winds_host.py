"""Linear-theory winds without a GPU: the arithmetic the kernels of
icar_b200/csrc/linear_winds.cu run -- lw_math.cuh, compiled here for the host by tests/cpp/lw_math_host.cpp --
and the host-side set-up the library does itself (icar_add_buffer_topo) against the NumPy restatement
oracle/linear_winds_np.py.  The transforms are numpy's here and cuFFT's on the device (tests/test_gpu_linear_winds.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import linear_winds_np as LW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lw") / "liblw_math_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Werror",
           os.path.join(ROOT, "tests", "cpp", "lw_math_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def fptr(a): return a.ctypes.data_as(C.POINTER(C.c_float))
def dptr(a): return a.ctypes.data_as(C.POINTER(C.c_double))


def f_order(a):
    """(nx, ny) indexed [i, j] -> the Fortran memory image (first index fastest) as a C array [j][i]."""
    return np.ascontiguousarray(a.T)


def spectrum(host, tf, lt, U, V, Nsq, z0, step, nsteps):
    fnx, fny = tf.shape
    kx = np.ascontiguousarray(lt.k[:, 0]); ly = np.ascontiguousarray(lt.l[0, :])
    uh = np.empty((fny, fnx), np.complex128); vh = np.empty((fny, fnx), np.complex128)
    tfm = f_order(tf)
    host.lw_host_spectrum(dptr(tfm.view(np.float64)), fptr(kx), fptr(ly), fnx, fny, C.c_float(U), C.c_float(V),
                          C.c_float(Nsq), C.c_float(z0), C.c_float(step), nsteps, dptr(uh.view(np.float64)),
                          dptr(vh.view(np.float64)))
    return uh.T, vh.T                      # back to [i, j]


def hill(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    x = np.arange(nx)[:, None]; y = np.arange(ny)[None, :]
    return (800 * np.exp(-((x - nx / 2) / 5.0) ** 2 - ((y - ny / 3) / 7.0) ** 2) + 30 * rng.random((nx, ny)) + 4 * x).astype(np.float32)


@pytest.mark.parametrize("smooth,buffer", [(0, 3), (5, 6), (5, 2), (2, 9)])
def test_add_buffer_topo_of_the_library_matches_the_restatement(smooth, buffer):
    import icar_b200
    terrain = hill(23, 17, seed=buffer)
    want = LW.add_buffer_topo(terrain, smooth, buffer)
    got = np.empty((17 + 2 * buffer, 23 + 2 * buffer), np.complex128)
    icar_b200.lib().icar_add_buffer_topo(fptr(f_order(terrain)), 23, 17, smooth, buffer, dptr(got.view(np.float64)))
    # numpy sums the smoothing windows pairwise, the Fortran (and the library) sequentially: last-bit differences
    np.testing.assert_allclose(got.T, want, rtol=1e-13, atol=0)
    if smooth == 0:
        assert np.array_equal(got.T, want)


def test_terrain_frequency_shift_and_normalisation(host):
    terrain = hill(20, 14)
    second = LW.add_buffer_topo(LW.add_buffer_topo(terrain, 5, 4).real.astype(np.float32), 0, 2)
    fnx, fny = second.shape
    raw = f_order(np.fft.fft2(second))
    out = np.empty((fny, fnx), np.complex128)
    host.lw_host_shift_norm(dptr(raw.view(np.float64)), dptr(out.view(np.float64)), fnx, fny)
    want, total = LW.terrain_frequency(terrain, 4)
    assert total == 6
    assert np.array_equal(out.T, want)


@pytest.mark.parametrize("U,V,Nsq,z", [(10.0, 3.0, 1e-4, 500.0), (-7.0, 2.5, 3e-5, 1250.0), (4.0, -9.0, 1e-4, 0.0),
                                       (10.0, 0.0, -1e-4, 800.0), (0.0, 6.0, 2e-4, 300.0)])
def test_one_height_matches_the_restatement(host, U, V, Nsq, z):
    terrain = hill(30, 22, seed=3)
    tf, _ = LW.terrain_frequency(terrain, 5)
    fnx, fny = tf.shape
    lt = LW.LinearTheoryData(fnx, fny, 2000.0)
    uh, vh = spectrum(host, tf, lt, U, V, Nsq, z, 0.0, 1)
    u = np.fft.ifft2(uh) * uh.size
    v = np.fft.ifft2(vh) * vh.size
    wu, wv = LW.linear_perturbation_at_height(U, V, Nsq, z, tf, lt)
    scale = max(np.abs(wu).max(), np.abs(wv).max())
    assert scale > 0
    # numpy divides complex numbers through a reciprocal, gfortran (and lw_math.cuh) by Smith's formula: the two
    # differ in the last bits of a double, and very rarely that flips the single-precision rounding of one
    # spectral coefficient (6e-8 of that coefficient)
    assert np.abs(u - wu).max() <= 1e-9 * scale and np.abs(v - wv).max() <= 1e-9 * scale


def test_layer_mean_taken_in_wavenumber_space(host):
    """linear_perturbation_constz averages the fields of n sub-layer heights; the kernel averages their spectra
    and transforms once."""
    terrain = hill(30, 22, seed=4)
    tf, _ = LW.terrain_frequency(terrain, 5)
    fnx, fny = tf.shape
    lt = LW.LinearTheoryData(fnx, fny, 4000.0)
    U, V, Nsq, zb, zt, minstep = 12.0, -4.0, 1e-4, 200.0, 830.0, 100.0
    n_steps = max(1, int(np.ceil((np.float32(zt) - np.float32(zb)) / np.float32(minstep))))
    assert n_steps == 7
    step = (np.float32(zt) - np.float32(zb)) / np.float32(n_steps)
    z0 = np.float32(zb) + step / np.float32(2)
    uh, vh = spectrum(host, tf, lt, U, V, Nsq, float(z0), float(step), n_steps)
    u = np.fft.ifft2(uh) * uh.size / n_steps
    v = np.fft.ifft2(vh) * vh.size / n_steps
    wu, wv = LW.linear_perturbation_constz(U, V, Nsq, zb, zt, minstep, tf, lt)
    scale = max(np.abs(wu).max(), np.abs(wv).max())
    assert np.abs(u - wu).max() <= 1e-9 * scale and np.abs(v - wv).max() <= 1e-9 * scale


def test_layer_fraction_is_the_restatements(host):
    rng = np.random.default_rng(7)
    zb = (rng.random(4000) * 900).astype(np.float32)
    zt = (zb + 50 + rng.random(4000) * 400).astype(np.float32)
    f = np.float32
    for cz, step in ((f(25.0), f(50.0)), (f(512.5), f(75.0)), (f(1300.0), f(100.0))):
        got = np.empty(4000, np.float32)
        host.lw_host_layer_fraction(fptr(zb), fptr(zt), 4000, C.c_float(cz), C.c_float(step), fptr(got))
        half = step / f(2)
        want = np.maximum(f(0), np.minimum(half, cz - zb) + np.minimum(f(0), zt - cz)
                          + np.minimum(half, zt - cz) + np.minimum(f(0), cz - zb)) / step
        assert np.array_equal(got, want.astype(np.float32))
        assert (got >= 0).all() and (got <= 1).all()


# ---------------------------------------------------------------------------------------------------------------
# The kernels themselves (icar_b200/csrc/lw_kernels.cuh) run on the CPU by tests/cpp/lw_kernels_host.cpp, with the
# host logic of icar_linear_winds_init / icar_linear_perturb (icar_b200/csrc/linear_winds.cu, icar_b200.cu) restated
# here in Python and numpy.fft where the product calls cuFFT: a whole linear_perturb against the restatement.

@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lwk") / "liblw_kernels_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Werror",
           os.path.join(ROOT, "tests", "cpp", "lw_kernels_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def cptr(a): return a.ctypes.data_as(C.c_void_p)


class EmulatedLinearWinds:
    """lw_create / lw_perturbation_constz / lw_perturbation_varyingz / lw_apply_level with emulated kernels.
    Arrays are C-order images of the device buffers: complex planes [fny][fnx], u [ny][nz][ldu], v [ny+1][nz][ld]."""

    def __init__(self, emu, terrain_yx, buffer, dx):
        import icar_b200
        self.emu = emu
        ny, nx = terrain_yx.shape
        self.nx, self.ny = nx, ny
        lib = icar_b200.lib()
        first = np.empty((ny + 2 * buffer, nx + 2 * buffer), np.complex128)
        lib.icar_add_buffer_topo(fptr(np.ascontiguousarray(terrain_yx, np.float32)), nx, ny, 5, buffer, cptr(first))
        first_real = np.ascontiguousarray(first.real.astype(np.float32))
        self.fnx, self.fny, self.buffer = nx + 2 * buffer + 4, ny + 2 * buffer + 4, buffer + 2
        second = np.empty((self.fny, self.fnx), np.complex128)
        lib.icar_add_buffer_topo(fptr(first_real), nx + 2 * buffer, ny + 2 * buffer, 0, 2, cptr(second))
        f = np.float32
        offset = f(3.1415927) / f(dx)
        gain = f(2) * offset / f(self.fnx - 1)
        self.kx = (np.arange(self.fnx, dtype=f) * gain - offset).astype(f)
        gain = f(2) * offset / f(self.fny - 1)
        self.ly = (np.arange(self.fny, dtype=f) * gain - offset).astype(f)
        raw = np.ascontiguousarray(np.fft.fft2(second))                     # CUFFT_FORWARD
        self.tf = np.empty_like(raw)
        emu.emu_shift_norm(cptr(raw), cptr(self.tf), self.fnx, self.fny)
        self.up = self.vp = None
        self.upert = self.vpert = None

    def _inverse(self, uhat, vhat):
        n = uhat.size
        self.up = np.ascontiguousarray(np.fft.ifft2(uhat) * n)             # CUFFT_INVERSE is unnormalised
        self.vp = np.ascontiguousarray(np.fft.ifft2(vhat) * n)

    def _spectrum(self, U, V, Nsq, z0, step, nsteps):
        uhat = np.empty((self.fny, self.fnx), np.complex128)
        vhat = np.empty_like(uhat)
        self.emu.emu_spectrum(cptr(self.tf), fptr(self.kx), fptr(self.ly), self.fnx, self.fny, C.c_float(U), C.c_float(V),
                              C.c_float(Nsq), C.c_float(z0), C.c_float(step), nsteps, cptr(uhat), cptr(vhat))
        return uhat, vhat

    def constz(self, U, V, Nsq, zb, zt, minimum_step):
        f = np.float32
        zb, zt, minimum_step = f(zb), f(zt), f(minimum_step)
        n_steps = max(1, int(np.ceil((zt - zb) / minimum_step)))
        step = (zt - zb) / f(n_steps)
        z0 = zb + step / f(2)
        self._inverse(*self._spectrum(U, V, Nsq, z0, step, n_steps))
        if n_steps > 1:
            self.emu.emu_divide(cptr(self.up), cptr(self.vp), C.c_longlong(self.up.size), C.c_double(n_steps))

    def varyingz(self, U, V, Nsq, zb_yx, zt_yx, minimum_step):
        f = np.float32
        zb = np.ascontiguousarray(zb_yx, f); zt = np.ascontiguousarray(zt_yx, f)
        start_z, end_z, thinnest = zb.min(), zt.max(), (zt - zb).min()
        step = min(f(minimum_step), thinnest)
        n = self.fnx * self.fny
        ibot = np.empty((self.fny, self.fnx), f); itop = np.empty_like(ibot)
        self.emu.emu_layer_bounds(fptr(zb), fptr(zt), self.nx, self.ny, self.buffer, C.c_float(start_z), C.c_float(end_z),
                                  fptr(ibot), fptr(itop), self.fnx, self.fny)
        ua = np.zeros((self.fny, self.fnx), np.complex128); va = np.zeros_like(ua)
        count = np.zeros((self.fny, self.fnx), f)
        cz = f(start_z + step / f(2))
        while cz < end_z:
            self._inverse(*self._spectrum(U, V, Nsq, cz, 0.0, 1))
            self.emu.emu_accumulate(cptr(self.up), cptr(self.vp), fptr(ibot), fptr(itop), C.c_float(cz), C.c_float(step),
                                    cptr(ua), cptr(va), fptr(count), C.c_longlong(n))
            cz = f(cz + step)
        self.up = np.empty_like(ua); self.vp = np.empty_like(va)
        self.emu.emu_divide_by_count(cptr(ua), cptr(va), fptr(count), cptr(self.up), cptr(self.vp), C.c_longlong(n))

    def apply_level(self, u_dev, v_dev, nx, ny, nz, ldu, ld, i0, j0, k, contribution, fraction):
        if self.upert is None:
            self.upert = np.zeros_like(u_dev); self.vpert = np.zeros_like(v_dev)
        self.emu.emu_apply_u(cptr(self.up), self.fnx, self.buffer, nx, ny, nz, ldu, i0, j0, k, fptr(u_dev), fptr(self.upert),
                             C.c_float(contribution), C.c_float(fraction))
        self.emu.emu_apply_v(cptr(self.vp), self.fnx, self.buffer, nx, ny, nz, ld, i0, j0, k, fptr(v_dev), fptr(self.vpert),
                             C.c_float(contribution), C.c_float(fraction))


@pytest.mark.parametrize("varying,tile", [(False, None), (True, None), (True, (9, 5, 21, 13))])
def test_whole_linear_perturb_with_emulated_kernels(emu, varying, tile):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from lw_expected import expected_winds
    from icar_b200 import ideal
    st = ideal.make_ideal_case(nx=40, nz=7, ny=24, perturb=True)
    nxg, nyg, nz = 40, 24, 7
    buffer, U, V, Nsq, frac, contrib = 7, 10.0, 4.0, 1e-4, 0.2, 0.7
    i0, j0, nx, ny = tile or (0, 0, nxg, nyg)
    if tile:                                                               # an image's tile of the global domain
        st["u_tile"] = np.ascontiguousarray(st["u"][j0:j0 + ny, :, i0:i0 + nx + 1])
        st["v_tile"] = np.ascontiguousarray(st["v"][j0:j0 + ny + 1, :, i0:i0 + nx])
    u0, v0 = (st["u_tile"], st["v_tile"]) if tile else (st["u"], st["v"])
    want_u, want_v, lh, zi = expected_winds(st, buffer, U, V, Nsq, frac, contrib, varying, calls=2, tile=tile)
    # device images with padded rows (Geom: ld, ldu multiples of 32), the padding poisoned
    ld, ldu = -(-nx // 32) * 32, -(-(nx + 1) // 32) * 32
    u_dev = np.full((ny, nz, ldu), np.float32(1e30)); u_dev[:, :, :nx + 1] = u0
    v_dev = np.full((ny + 1, nz, ld), np.float32(1e30)); v_dev[:, :, :nx] = v0
    lw = EmulatedLinearWinds(emu, st["terrain"], buffer, st["dx"])
    want_tf = LW.terrain_frequency(np.ascontiguousarray(st["terrain"].T), buffer)[0]
    assert np.abs(lw.tf.T - want_tf).max() <= 1e-7 * np.abs(want_tf).max()   # single-precision roundings may flip
    f = np.float32
    for _ in range(2):
        for k in range(nz):                                                # icar_linear_perturb's level loop
            if varying:
                zb = zi[:, k, :] - st["terrain"]
                lw.varyingz(U, V, Nsq, zb, zb + st["dz_interface"][:, k, :], 100.0)
            else:
                half = st["dz_levels"][k] / f(2)
                lw.constz(U, V, Nsq, lh[k] - half, lh[k] + half, 100.0)
            lw.apply_level(u_dev, v_dev, nx, ny, nz, ldu, ld, i0, j0, k, contrib, frac)
    assert (u_dev[:, :, nx + 1:] == np.float32(1e30)).all() and (v_dev[:, :, nx:] == np.float32(1e30)).all()
    du, dv = u_dev[:, :, :nx + 1] - u0, v_dev[:, :, :nx] - v0
    wu, wv = want_u - u0, want_v - v0
    scale = max(np.abs(wu).max(), np.abs(wv).max())
    assert scale > 1e-3
    assert np.abs(du - wu).max() <= 2e-6 * max(scale, 10.0) and np.abs(dv - wv).max() <= 2e-6 * max(scale, 10.0)


def test_kernel_arithmetic_reproduces_the_references_python_solver(host):
    """lw_math.cuh (what k_lw_spectrum runs per wavenumber) against u_hat of the reference's own Python solver,
    helpers/ideal_linear.py:solve(), frozen in tests/golden/linear_theory_1d.npz: the x-z case, l = 0."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "linear_theory_1d.npz"))
    for n in range(int(g["ncases"])):
        Fzs, U, Nsq, dx = g[f"Fzs_{n}"], float(g[f"U_{n}"]), float(g[f"Ndsq_{n}"]), float(g["dx"])
        Nx = Fzs.shape[0]
        kx = np.ascontiguousarray(LW.LinearTheoryData(Nx, 2, dx).k[:, 0])
        ly = np.zeros(1, np.float32)
        tf = np.ascontiguousarray(Fzs.astype(np.complex128))
        for iz, z in enumerate(g["zlevels"]):
            uh = np.empty(Nx, np.complex128); vh = np.empty(Nx, np.complex128)
            host.lw_host_spectrum(dptr(tf.view(np.float64)), fptr(kx), fptr(ly), Nx, 1, C.c_float(U), C.c_float(0.0),
                                  C.c_float(Nsq), C.c_float(z), C.c_float(0.0), 1, dptr(uh.view(np.float64)),
                                  dptr(vh.view(np.float64)))
            u = (np.fft.ifft(uh) * Nx).real
            want = g[f"u_hat_{n}"][iz]
            assert np.abs(u - want).max() <= 2e-7 * np.abs(want).max(), (n, z)      # measured 4.5e-8
            assert not vh.any()
