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
