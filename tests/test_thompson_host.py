"""k_mp_thompson run ON THE CPU (tests/cpp/thompson_host.cpp: one OS thread per CUDA thread of a CTA, pthread
barriers, host atomics, the host's libm) against the oracle's mp_gt_driver, from identical inputs, call after call.

The kernel runs on hardware (tests/test_gpu_thompson.py); this emulation exists so that the next versions of it --
it is the slowest kernel of the repository, DESIGN.md section 3.4 -- can be checked before they cost GPU time.  With
the same libm on both sides the kernel's statements are the oracle's: the only differences come from the three shapes
of x**y the kernel evaluates with cheaper routines (10.**y through exp10f, x**3 by multiplication, x**(1/3 in REAL)
through a refined cbrtf; mp_thompson_kernel.inc), and they stay far inside the parity bar -- every cell within 1e-6 of
the field maximum, relative L2 <= 1e-7 per field, most fields of most calls unchanged in every bit (measured <= 4e-8;
the GPU test's bars are 2e-5 and 1e-6).  With plain powf in their place the results were bit-identical (round 1)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from icar_b200 import ideal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f = np.float32
FIELDS = ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr")


def build(*defines):
    import icar_b200
    icar_b200.lib()
    so = os.path.join(tempfile.mkdtemp(prefix="thk"), "libthompson_host.so")
    libdir = os.path.join(ROOT, "icar_b200")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w", *defines,
           os.path.join(ROOT, "tests", "cpp", "thompson_host.cpp"), "-L", libdir, "-licar_b200", f"-Wl,-rpath,{libdir}", "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emu():
    return build()


@pytest.fixture(scope="module")
def oracle_thompson(oracle):
    return oracle.Thompson(big_tables=True)


def emulated_driver(emu, oth, st, dt, its, ite, jts, jte, kts, kte, along_y=0):
    """mp_gt_driver through the emulated kernel, in place on the state dict; 0-based inclusive bounds."""
    from oracle import oracle as O
    ny, nz, nx = st["th"].shape
    ld = -(-nx // 32) * 32

    def pad(a):
        out = np.zeros(a.shape[:-1] + (ld,), a.dtype); out[..., :nx] = a
        return out
    names = ("qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "th", "pii", "pressure", "dz_mass")
    dev = {k: pad(st[k]) for k in names}
    two = {k: pad(st[k]) for k in ("rainnc", "snownc", "graupelnc", "sr")}
    tables = [np.ascontiguousarray(oth.table(n)) for n in O.THOMPSON_TABLES]
    tp = (C.c_void_p * 29)(*[t.ctypes.data for t in tables])
    diag = (C.c_ulonglong * 1)(0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu_thompson_driver(nx, nz, ny, ld, ld, C.byref(oth.options), tp, *[P(dev[k]) for k in names], C.c_float(dt),
                                 P(two["rainnc"]), P(two["snownc"]), P(two["graupelnc"]), P(two["sr"]),
                                 its, ite, jts, jte, kts, kte, along_y, diag)
    assert rc == 1
    for k in FIELDS:
        st[k][...] = dev[k][..., :nx]
    for k in two:
        st[k][...] = two[k][..., :nx]
    return int(diag[0])


def compare(got, ref, what):
    """Returns the number of fields that are equal in every bit."""
    exact = 0
    for k in FIELDS + ("rainnc", "snownc", "graupelnc", "sr"):
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        assert np.isfinite(a).all(), (what, k)
        scale = max(np.abs(b).max(), 1e-30)
        assert np.abs(a - b).max() <= 1e-6 * scale, (what, k, np.abs(a - b).max() / scale)
        assert np.sqrt(((a - b) ** 2).sum()) <= 1e-7 * max(np.sqrt((b ** 2).sum()), 1e-30), (what, k)
        exact += int(np.array_equal(got[k], ref[k]))
    return exact


@pytest.mark.parametrize("dims,dt,calls", [((20, 20, 6), 60.0, 14), ((12, 40, 5), 120.0, 8)])
def test_emulated_kernel_matches_the_oracle_call_after_call(emu, oracle_thompson, dims, dt, calls):
    nx, nz, ny = dims
    ref = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
    seen = np.zeros(5)
    exact = 0
    for call in range(calls):
        st = ideal.copy_state(ref)
        emulated_driver(emu, oracle_thompson, st, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
        oracle_thompson.driver(ref, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
        exact += compare(st, ref, f"call {call}")
        np.testing.assert_array_equal(st["qv"][0], ref["qv"][0])               # untouched outside the tile
        np.testing.assert_array_equal(st["th"][:, :, 0], ref["th"][:, :, 0])
        seen = np.maximum(seen, [ref[k].max() for k in ("qc", "qr", "qi", "qs", "qg")])
    assert exact >= 0.6 * calls * 13                       # most fields of most calls do not change by a bit
    if nz == 20:
        assert (seen > 1e-6).all(), f"scenario did not exercise every species: {seen}"


def test_lanes_along_y_and_a_sub_tile(emu, oracle_thompson):
    """The strips of mp(halo) run with the lanes along y; a tile that stops short of the model top (top_mp_level)."""
    nx, nz, ny = 14, 24, 19
    ref = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
    for _ in range(5):
        oracle_thompson.driver(ref, 60.0, 1, nx - 2, 1, ny - 2, 0, nz - 1)
    st = ideal.copy_state(ref)
    emulated_driver(emu, oracle_thompson, st, 60.0, 3, 4, 2, 16, 0, nz - 4, along_y=1)
    oracle_thompson.driver(ref, 60.0, 3, 4, 2, 16, 0, nz - 4)
    compare(st, ref, "strip")
