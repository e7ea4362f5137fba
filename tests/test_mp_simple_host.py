"""k_mp_simple run ON THE CPU (tests/cpp/mp_simple_host.cpp: one OS thread per CUDA thread of a CTA) against the
oracle's mp_simple_driver: BIT FOR BIT, like on the device (the kernel evaluates glibc's expf itself and keeps the
Fortran's operation order), call after call, for every cells-per-thread instantiation a level count can select,
with the lanes along y, and on a tile that stops below the model top."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from icar_b200 import ideal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f = np.float32
FIELDS = ("th", "qv", "qc", "qr", "qs", "rain", "snow")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(tempfile.mkdtemp(prefix="mps"), "libmp_simple_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w",
           os.path.join(ROOT, "tests", "cpp", "mp_simple_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def emulated_driver(emu, st, dt, its, ite, jts, jte, kts, kte, along_y=0, cpt=0):
    ny, nz, nx = st["th"].shape
    ld = -(-nx // 32) * 32

    def pad(a):
        out = np.zeros(a.shape[:-1] + (ld,), a.dtype); out[..., :nx] = a
        return out
    names = ("pressure", "th", "pii", "rho", "qv", "qc", "qr", "qs", "rain", "snow")
    dev = {k: pad(st[k]) for k in names}
    dz = pad(st["dz_mass"])
    diag = (C.c_ulonglong * 3)(0, 0, 0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    # cloud2snow / cloud2rain of mp_simple_driver (mp_simple.f90:619-620), as icar_mp forms them: exp(-1 * (1/2000) * dt) in REAL
    # (glibc's expf, which is what std::exp(float) calls in the library and in the oracle; NumPy's differs by an ulp here)
    libm = C.CDLL("libm.so.6")
    libm.expf.restype, libm.expf.argtypes = C.c_float, [C.c_float]
    c2s = libm.expf(float((f(-1) * (f(1) / f(2000.0))) * f(dt)))
    c2r = libm.expf(float((f(-1) * (f(1) / f(500.0))) * f(dt)))
    rc = emu.emu_mp_simple(nx, nz, ny, ld, ld, *[P(dev[k]) for k in names], C.c_float(dt), P(dz), its, ite, jts, jte, kts, kte,
                           along_y, C.c_float(c2r), C.c_float(c2s), cpt, diag)
    assert rc == 1
    for k in FIELDS:
        st[k][...] = dev[k][..., :nx]
    return tuple(diag)


def state(nx, nz, ny):
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    st["qv"] *= f(3.0)                                            # saturated aloft: cloud, rain, snow within a few calls
    return st


@pytest.mark.parametrize("dims,cpt", [((40, 20, 6), 0), ((36, 40, 5), 0), ((20, 16, 5), 1), ((12, 50, 4), 0), ((34, 7, 5), 0)])
def test_emulated_kernel_is_the_oracle_bit_for_bit(emu, oracle, dims, cpt):
    nx, nz, ny = dims
    ref = state(nx, nz, ny)
    dt = 60.0
    for call in range(8):
        st = ideal.copy_state(ref)
        emulated_driver(emu, st, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1, cpt=cpt)
        oracle.mp_simple_driver(ref["pressure"], ref["th"], ref["pii"], ref["rho"], ref["qv"], ref["qc"], ref["qr"], ref["qs"],
                                ref["rain"], ref["snow"], f(dt), ref["dz_mass"], 1, nx - 2, 1, ny - 2, 0, nz - 1)
        for k in FIELDS:
            assert np.array_equal(st[k], ref[k]), (call, k, float(np.abs(st[k].astype(np.float64) - ref[k]).max()))
    assert ref["qc"].max() > 0 and ref["rain"].max() > 0         # the scenario rained


def test_lanes_along_y_and_a_tile_below_the_top(emu, oracle):
    nx, nz, ny = 9, 24, 40
    ref = state(nx, nz, ny)
    for _ in range(4):
        oracle.mp_simple_driver(ref["pressure"], ref["th"], ref["pii"], ref["rho"], ref["qv"], ref["qc"], ref["qr"], ref["qs"],
                                ref["rain"], ref["snow"], f(60.0), ref["dz_mass"], 1, nx - 2, 1, ny - 2, 0, nz - 1)
    st = ideal.copy_state(ref)
    emulated_driver(emu, st, 60.0, 2, 3, 1, ny - 2, 0, nz - 5, along_y=1)
    oracle.mp_simple_driver(ref["pressure"], ref["th"], ref["pii"], ref["rho"], ref["qv"], ref["qc"], ref["qr"], ref["qs"],
                            ref["rain"], ref["snow"], f(60.0), ref["dz_mass"], 2, 3, 1, ny - 2, 0, nz - 5)
    for k in FIELDS:
        assert np.array_equal(st[k], ref[k]), k
