"""MPDATA pass 1 as the product runs it (k_upwind_all_v4: all scalars in one launch, four cells per thread, quotients from
one correctly rounded reciprocal per divisor) with its kernel body run ON THE CPU (tests/cpp/upwind_host.cpp over
icar_b200/csrc/upwind_all.cuh and host_emu.h) against the oracle's upwind_advection (adv_mpdata.f90:44-105), BIT FOR BIT:
bottom / top levels, a one-level grid, the boundary ring, a ragged last group of four cells, tiny and zero mixing
ratios, both signs of every wind component."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from icar_b200 import ideal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f = np.float32


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(tempfile.mkdtemp(prefix="upw"), "libupwind_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Werror", "-Wno-unknown-pragmas",
           "-Wno-unused-function", os.path.join(ROOT, "tests", "cpp", "upwind_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def device_image(a, ld, rows_y=None, fill=0.0):
    """(ny, nz, nx) -> (rows_y or ny, nz, ld) with the padding (and any extra row) set to `fill`."""
    ny, nz, nx = a.shape
    out = np.full((rows_y or ny, nz, ld), f(fill))
    out[:ny, :, :nx] = a
    return out


@pytest.mark.parametrize("dims", [(150, 23, 19), (64, 1, 9), (33, 2, 7), (40, 9, 12), (37, 8, 5)])
def test_pass1_equals_the_oracle_bit_for_bit(emu, oracle, dims):
    nx, nz, ny = dims
    st = ideal.make_ideal_case(nx=nx, nz=max(nz, 8), ny=ny, perturb=True)
    for k, v in list(st.items()):
        if isinstance(v, np.ndarray) and v.ndim == 3:
            st[k] = np.ascontiguousarray(v[:, :nz, :])
    rng = np.random.default_rng(8)
    shape = st["qv"].shape
    st["qc"] = (rng.random(shape) ** 8 * 1e-3).astype(f)
    st["qr"] = (10.0 ** rng.uniform(-30, -3, shape)).astype(f)     # tiny, but fluxes stay normal numbers: the shared-reciprocal
    #                                   quotient is the IEEE one outside the subnormal range (tests/test_fp_identities.py)
    st["qr"][rng.random(shape) < 0.3] = 0
    st["w"] = (st["w"] + rng.uniform(-0.3, 0.3, shape)).astype(f)
    st["u"] = (st["u"] * rng.uniform(-1.2, 1.2, st["u"].shape)).astype(f)       # both signs of every wind component
    st["v"] = (st["v"] + rng.uniform(-6, 6, st["v"].shape)).astype(f)
    dt = f(float(st["dt"]) * 0.5)
    ones = np.ones(shape, f)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], ones, dt,
                                    f(st["dx"]), 2)
    names = ("qv", "qc", "qr", "qs", "th")
    want = {k: oracle.upwind_advection(st[k], U, V, W, ones, st["advection_dz"], st["jacobian"]) for k in names}
    # device layout (mpdata_staged.cu): U[i] = face between i-1 and i, V[j] likewise (ny + 1 rows), W[k] = face above k
    ld = -(-nx // 32) * 32
    Ud = np.zeros((ny, nz, ld), f); Ud[:, :, 1:nx] = U
    Vd = np.zeros((ny + 1, nz, ld), f); Vd[1:ny, :, :nx] = V
    Wd = device_image(W, ld)
    dz, jaco = device_image(st["advection_dz"], ld, fill=1.0), device_image(st["jacobian"], ld, fill=1.0)
    qin = [device_image(st[k], ld, rows_y=ny + 1) for k in names]
    q2 = [np.full((ny + 1, nz, ld), f(7e7)) for _ in names]
    out = [np.full((ny + 1, nz, ld), f(7e7)) for _ in names]
    P = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    for n in (5, 2):
        for a in q2 + out:
            a[...] = f(7e7)
        rc = emu.emu_upwind_all(nx, nz, ny, ld, n, P(qin[:n]), P(q2[:n]), P(out[:n]), fp(Ud), fp(Vd), fp(Wd), fp(dz), fp(jaco))
        assert rc == 0
        for s in range(n):
            got = q2[s][:ny, :, :nx]
            assert np.array_equal(got, want[names[s]]), (names[s], np.abs(got - want[names[s]]).max())
            # the boundary ring also goes to `out` (it is never touched by the corrector)
            ring = np.ones((ny, nz, nx), bool); ring[1:-1, :, 1:-1] = False
            assert np.array_equal(out[s][:ny, :, :nx][ring], st[names[s]][ring])
            assert (out[s][:ny, :, :nx][~ring] == f(7e7)).all()
            assert (q2[s][ny] == f(7e7)).all()                                   # nothing written past the last row
