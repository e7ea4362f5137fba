"""The fused MPDATA corrector kernels run ON THE CPU (tests/cpp/corrector_host.cpp: one OS thread per CUDA thread
of a CTA, pthread barriers for __syncthreads and warp shuffles, an emulated mbarrier / TMA box copy) against the
oracle's advect3d (order 2, FCT), at the bar of the GPU test (tests/test_gpu_parity.py: RTOL_FUSED = 1e-6 of the field
maximum per call, per unit of 1 / jacobian).

Two purposes.  (1) The instantiations that have run on hardware -- the scalar forms and the default f32x2 form -- pass
here too, which validates the emulation.  (2) The candidates that have NOT run on hardware -- the packed form for more
than 40 levels (ICAR_CORR_V2=1, k_corrector_fused_pk) and the double-buffered 49..52-level instantiation
(ICAR_FUSED_NZP52=1) -- are checked before they cost GPU time; any later variant of the dominant kernel can be, too."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from icar_b200 import ideal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f = np.float32
RTOL_FUSED = 1e-6


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(tempfile.mkdtemp(prefix="corr"), "libcorrector_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
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


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def case(nx, nz, ny):
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True, dz_level=400.0 if nz >= 7 else 800.0)
    rng = np.random.default_rng(4)
    for k in ("qc", "qr", "qs"):
        st[k] = (rng.uniform(0, 1e-3, st[k].shape) * (rng.random(st[k].shape) < 0.5)).astype(f)
    st["w"] = (st["w"] + rng.uniform(-0.3, 0.3, st["w"].shape)).astype(f)
    return st


def emulated_advect(emu, oracle, st, dt, kind, cpt, nzp, stages, chunk=None, v4=1, names=("th", "qc", "qr"), rowbufs=2, nsync=0):
    """fused_advect_all: face coefficients (emulated kernel), pass 1 (the oracle's: bit-identical to the device's),
    corrector (emulated kernel).  Returns {name: field after one advect call}."""
    nx, nz, ny = st["nx"], st["nz"], st["ny"]
    ld = -(-nx // 32) * 32
    ones = np.ones(st["qv"].shape, f)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    Ud = np.zeros((ny + 1, nz, ld), f); Ud[:ny, :, 1:nx] = U
    Vd = np.zeros((ny + 1, nz, ld), f); Vd[1:ny, :, :nx] = V
    Wd = image(W, ld, rows_y=ny + 1)
    dz, jaco = image(st["advection_dz"], ld, rows_y=ny + 1, fill=1.0), image(st["jacobian"], ld, rows_y=ny + 1, fill=1.0)
    Wn = (Wd / dz).astype(f)                                             # k_wn_G / k_wind_normalise_v4: W / dz
    coef = [np.full((ny + 1, nz, ld), f(3e33)) for _ in range(11)]        # poison: every cell a box reads must be written
    emu.emu_face_coefficients(v4, nx, nz, ny, ld, fp(Ud), fp(Vd), fp(Wn), fp(jaco), fp(dz), P(coef))
    q = [image(st[k], ld, rows_y=ny + 1) for k in names]
    q2 = [image(oracle.upwind_advection(st[k], U, V, W, ones, st["advection_dz"], st["jacobian"]), ld, rows_y=ny + 1) for k in names]
    out = [image(st[k], ld, rows_y=ny + 1) for k in names]               # staged_upwind_all copies the ring into `out`
    for o, k in zip(out, names):
        o[1:ny - 1, :, 1:nx - 1] = f(-7e7)                                 # the interior must come from the corrector
    rc = emu.emu_corrector(kind, cpt, nzp, stages, rowbufs, nsync, nx, nz, ny, ld, len(names), P(q2), P(q), P(out), P(coef), chunk or (ny - 2))
    assert rc == 0, "no such instantiation"
    return {k: o[:ny, :, :nx] for k, o in zip(names, out)}


def check(emu, oracle, dims, kind, cpt, nzp, stages, chunk=None, v4=1, rowbufs=2, nsync=0):
    nx, nz, ny = dims
    st = case(nx, nz, ny)
    dt = float(st["dt"]) * 0.8
    got = emulated_advect(emu, oracle, st, dt, kind, cpt, nzp, stages, chunk, v4, rowbufs=rowbufs, nsync=nsync)
    U, V, W = oracle.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    bar = RTOL_FUSED * max(1.0, float(1.0 / st["jacobian"].min()))
    worst = 0.0
    for k, a in got.items():
        want = st[k].copy()
        oracle.advect3d(want, None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
        assert np.isfinite(a).all(), k
        e = rel_err(a, want)
        worst = max(worst, e)
        assert e <= bar, (k, e)
        assert np.array_equal(a[0], st[k][0]) and np.array_equal(a[:, :, 0], st[k][:, :, 0])     # the ring is never advected
    return got, worst


# (dims, kind, cpt, nzp, stages): kind 0 = k_corrector_fused, 1 = _p2 (default for nz <= 40), 2 = _pk (candidate)
VALIDATED = [((45, 7, 7), 0, 2, 20, 2), ((40, 20, 7), 1, 2, 20, 2), ((40, 33, 6), 1, 2, 40, 2), ((36, 48, 7), 0, 4, 48, 2),
             ((40, 50, 7), 0, 4, 64, 1)]
CANDIDATES = [((40, 20, 7), 2, 2, 20, 2), ((40, 33, 6), 2, 2, 40, 2), ((36, 48, 7), 2, 4, 48, 2), ((40, 50, 7), 2, 4, 52, 2),
              ((40, 50, 7), 0, 4, 52, 2), ((33, 60, 7), 2, 4, 64, 1)]


@pytest.mark.parametrize("dims,kind,cpt,nzp,stages", VALIDATED)
def test_instantiations_that_have_run_on_hardware_pass_the_emulation(emu, oracle, dims, kind, cpt, nzp, stages):
    check(emu, oracle, dims, kind, cpt, nzp, stages)


@pytest.mark.parametrize("dims,kind,cpt,nzp,stages", CANDIDATES)
def test_candidates_that_have_not_run_on_hardware(emu, oracle, dims, kind, cpt, nzp, stages):
    check(emu, oracle, dims, kind, cpt, nzp, stages)


def test_packed_forms_agree_bit_for_bit_and_y_chunks_do_not_matter(emu, oracle):
    """k_corrector_fused_pk<1, ..> must equal k_corrector_fused_p2 bit for bit (its header says so), and cutting the
    march along y into chunks (each CTA re-reads four warm-up planes) must not change a bit either."""
    dims = (40, 20, 9)
    a, _ = check(emu, oracle, dims, 1, 2, 20, 2)
    b, _ = check(emu, oracle, dims, 2, 2, 20, 2)
    c, _ = check(emu, oracle, dims, 1, 2, 20, 2, chunk=4)
    d, _ = check(emu, oracle, dims, 1, 2, 20, 2, v4=0)                    # one-cell-per-thread coefficient kernel
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
        assert np.array_equal(a[k], d[k]), k


@pytest.mark.parametrize("dims,chunk", [((33, 60, 7), None), ((30, 55, 7), 3)])
def test_two_stage_candidate_for_53_to_60_levels(emu, oracle, dims, chunk):
    """ICAR_FUSED_NZP60=1 + ICAR_CORR_V2=1: k_corrector_fused_pk<2, 60, 2, ROWBUFS = 1> -- BASELINE configs[4] has 60
    levels.  One set of exchange rows makes room for the second TMA stage (221 KB); the warm-up planes take one extra
    barrier.  The OS threads of the emulation really race, so a missing barrier shows up as a wrong or varying
    result: the kernel is run twice and must give the same bits, which must be those of the single-stage kernel."""
    a, worst = check(emu, oracle, dims, 2, 4, 60, 2, chunk=chunk, rowbufs=1)
    b, _ = check(emu, oracle, dims, 2, 4, 60, 2, chunk=chunk, rowbufs=1) if not chunk else (a, 0)
    c, _ = check(emu, oracle, dims, 2, 4, 64, 1, chunk=chunk)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k


def test_more_shared_memory_than_a_cta_has_is_refused(emu, oracle):
    st = case(30, 64, 6)
    with pytest.raises(AssertionError, match="no such instantiation"):
        emulated_advect(emu, oracle, st, float(st["dt"]) * 0.8, 2, 4, 64, 2)          # two stages of 64 levels: 236 KB


def test_neighbour_only_synchronisation_experiment(emu, oracle):
    """ICAR_CORR_NSYNC=1 (with ICAR_CORR_V2=1): k_corrector_fused_pk<.., NSYNC = 1> has no CTA-wide barrier in its plane
    loop -- a warp synchronises with the warps tz - 1 and tz + 1 only, through monotonically increasing flags in shared
    memory, and a TMA stage is released through an "empty" mbarrier.  The first version used one mbarrier per warp and
    exchange and DEADLOCKED here (a warp publishes plane P + 1 before it waits, so a neighbour still waiting for plane P
    sees the parity it started with) -- on the CPU, not on a GPU box.  Same bits as the barrier version, with the OS
    threads of the emulation really drifting apart.  Statically the variant is 5688 SASS instructions against 3016
    (warp-level reconvergence around the polls), so it is an experiment to be timed, not a candidate default."""
    dims = (40, 20, 8)
    a, _ = check(emu, oracle, dims, 2, 2, 20, 2)
    b, _ = check(emu, oracle, dims, 2, 2, 20, 2, nsync=1)
    c, _ = check(emu, oracle, dims, 2, 2, 20, 2, nsync=1, chunk=3)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
