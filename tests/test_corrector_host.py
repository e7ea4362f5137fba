"""The fused MPDATA corrector (k_face_coefficients + k_corrector_fused) run ON THE CPU (tests/cpp/corrector_host.cpp: one OS
thread per CUDA thread of a CTA, pthread barriers for __syncthreads and warp shuffles, an emulated mbarrier / TMA box copy)
against the oracle's advect3d (order 2, FCT): BIT FOR BIT, for every instantiation the product launches.

The kernel evaluates the reference's operations in the reference's order (corrector_fused.inc), so this is an equality test;
it covers arithmetic, indexing, the register windows, the z-exchange rows and their barriers (the OS threads of the
emulation really race: a missing barrier shows up as a wrong or varying result), ragged tiles and y-chunks.  What it cannot
cover -- the device's division sequence, launch geometry, TMA descriptors -- is covered by tests/test_gpu_parity.py."""
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
    so = os.path.join(tempfile.mkdtemp(prefix="corr"), "libcorrector_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-psabi",
           os.path.join(ROOT, "tests", "cpp", "corrector_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def image(a, ld, rows_y=None, fill=0.0):
    ny, nz, nx = a.shape
    out = np.full((rows_y or ny, nz, ld), f(fill))
    out[:ny, :, :nx] = a
    return out


def P(arrs): return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
def fp(a): return a.ctypes.data_as(C.c_void_p)


def case(nx, nz, ny):
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True, dz_level=400.0 if nz >= 7 else 800.0)
    rng = np.random.default_rng(4)
    for k in ("qc", "qr", "qs"):
        st[k] = (rng.uniform(0, 1e-3, st[k].shape) * (rng.random(st[k].shape) < 0.5)).astype(f)
    st["w"] = (st["w"] + rng.uniform(-0.3, 0.3, st["w"].shape)).astype(f)
    return st


def emulated_advect(emu, oracle, st, dt, np_, nzp, stages, chunk=None, names=("th", "qc", "qr")):
    """fused_advect_all: face coefficients (emulated kernel), pass 1 (the oracle's: tests/test_upwind_host.py shows the
    device's is the same bit for bit), corrector (emulated kernel).  Returns {name: field after one advect call}."""
    nx, nz, ny = st["nx"], st["nz"], st["ny"]
    ld = -(-nx // 32) * 32
    ones = np.ones(st["qv"].shape, f)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    Ud = np.zeros((ny + 1, nz, ld), f); Ud[:ny, :, 1:nx] = U
    Vd = np.zeros((ny + 1, nz, ld), f); Vd[1:ny, :, :nx] = V
    Wd = image(W, ld, rows_y=ny + 1)
    dz, jaco = image(st["advection_dz"], ld, rows_y=ny + 1, fill=1.0), image(st["jacobian"], ld, rows_y=ny + 1, fill=1.0)
    Wn = (Wd / dz).astype(f)                                             # k_wind_normalise_v4: W / dz
    coef = [np.full((ny + 1, nz, ld), f(3e33)) for _ in range(10)]        # poison: every cell a box reads must be written
    emu.emu_face_coefficients(nx, nz, ny, ld, fp(Ud), fp(Vd), fp(Wn), fp(jaco), P(coef))
    q = [image(st[k], ld, rows_y=ny + 1) for k in names]
    q2 = [image(oracle.upwind_advection(st[k], U, V, W, ones, st["advection_dz"], st["jacobian"]), ld, rows_y=ny + 1) for k in names]
    out = [image(st[k], ld, rows_y=ny + 1) for k in names]               # staged_upwind_all copies the ring into `out`
    for o, k in zip(out, names):
        o[1:ny - 1, :, 1:nx - 1] = f(-7e7)                                 # the interior must come from the corrector
    rc = emu.emu_corrector(np_, nzp, stages, nx, nz, ny, ld, len(names), P(q2), P(q), P(out), P(coef), fp(jaco), fp(dz),
                           chunk or (ny - 2))
    assert rc == 0, {1: "no such instantiation", 2: "more shared memory than a CTA has"}[rc]
    return {k: o[:ny, :, :nx] for k, o in zip(names, out)}


def check(emu, oracle, dims, np_, nzp, stages, chunk=None):
    nx, nz, ny = dims
    st = case(nx, nz, ny)
    dt = float(st["dt"]) * 0.8
    got = emulated_advect(emu, oracle, st, dt, np_, nzp, stages, chunk)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    for k, a in got.items():
        want = st[k].copy()
        oracle.advect3d(want, None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
        assert np.isfinite(a).all(), k
        assert np.array_equal(a, want), (k, int((a != want).sum()), float(np.abs(a - want).max()))
    return got


# (dims, pairs per thread, level slots, TMA stages): every instantiation fused_plan can choose
INSTANTIATIONS = [((45, 7, 7), 1, 20, 2), ((40, 20, 7), 1, 20, 2), ((40, 33, 6), 1, 40, 2), ((36, 48, 7), 2, 48, 2),
                  ((40, 50, 7), 1, 52, 2), ((30, 55, 6), 1, 56, 2), ((33, 60, 7), 2, 60, 2), ((33, 62, 6), 2, 64, 1)]


@pytest.mark.parametrize("dims,np_,nzp,stages", INSTANTIATIONS)
def test_corrector_equals_the_oracle_bit_for_bit(emu, oracle, dims, np_, nzp, stages):
    check(emu, oracle, dims, np_, nzp, stages)


def test_y_chunks_and_repeated_runs_do_not_change_a_bit(emu, oracle):
    """Cutting the march along y into chunks (each CTA re-reads four warm-up planes) must not change a bit, and neither
    may the scheduling of the racing OS threads (one set of z-exchange rows, three barriers per output plane)."""
    dims = (40, 20, 9)
    a = check(emu, oracle, dims, 1, 20, 2)
    b = check(emu, oracle, dims, 1, 20, 2, chunk=4)
    c = check(emu, oracle, dims, 1, 20, 2, chunk=3)
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k


def blob_case(nx, nz, ny):
    """Fields that are zero over most of the domain: blocks of zero and non-zero cells meet inside a warp's columns, between
    the level groups of neighbouring warps, between x tiles and from one plane to the next."""
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    rng = np.random.default_rng(11)
    st["w"] = (st["w"] + rng.uniform(-0.3, 0.3, st["w"].shape)).astype(f)
    z = lambda: np.zeros((ny, nz, nx), f)
    qc, qr, qs = z(), z(), z()
    qc[4:9, 3:9, 20:50] = rng.uniform(1e-5, 1e-3, (5, 6, 30))          # one block in the middle, crossing two x tiles
    qr[:, 0:2, :] = rng.uniform(1e-5, 1e-3, (ny, 2, nx))                # the lowest two levels only
    qr[10:, 12:, 60:] = rng.uniform(1e-5, 1e-3, qr[10:, 12:, 60:].shape)   # and a corner up to the last plane / level / column
    qs[7, 10, 33] = f(1e-3)                                             # a single cell
    st["qc"], st["qr"], st["qs"] = qc, qr, qs
    return st


@pytest.mark.parametrize("np_,nzp,nz,chunk", [(1, 20, 20, None), (1, 20, 17, 5), (2, 48, 44, None)])
def test_fields_that_are_mostly_zero(emu, oracle, np_, nzp, nz, chunk):
    """Hydrometeor fields as a model run has them -- zero almost everywhere, with sharp edges: the result must be the
    oracle's bit for bit where zero and non-zero blocks meet (0 / (0 + 1e-10) quotients, limiter factors of 0 / 1e-15)."""
    nx, ny = 72, 14
    st = blob_case(nx, nz, ny)
    dt = float(st["dt"]) * 0.8
    names = ("qc", "qr", "qs", "th")
    got = emulated_advect(emu, oracle, st, dt, np_, nzp, 2, chunk, names=names)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    for k in names:
        want = st[k].copy()
        oracle.advect3d(want, None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
        assert np.array_equal(got[k], want), (k, int((got[k] != want).sum()), float(np.abs(got[k] - want).max()))
    assert (got["qs"] != 0).sum() > 1 and (got["qc"] == 0).mean() > 0.5


def test_more_shared_memory_than_a_cta_has_is_refused(emu, oracle):
    st = case(30, 64, 6)
    with pytest.raises(AssertionError, match="more shared memory"):
        emulated_advect(emu, oracle, st, float(st["dt"]) * 0.8, 2, 64, 2)          # two stages of 64 levels: 241 KB
