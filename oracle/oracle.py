"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes handle on oracle/_build/libicar_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product (icar_b200/) never does.

PARITY PINNED to the reference's own source text, executed by oracle/f90run (the reference holds no
golden vectors and cannot be compiled in this image): see oracle/mpdata_oracle.hpp and tests/test_reference_source.py.

All arrays are numpy arrays in the reference's memory order: Fortran (i,k,j)
column-major == numpy C-order shape (ny, nz, nx).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libicar_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C++ restatement (g++ -O2 -ffp-contract=off -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("oracle_api.cpp", "mpdata_oracle.hpp", "mp_simple_oracle.hpp", "mp_thompson_oracle.hpp",
             "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
FAST_FLAGS = "-O2 -mfma -ftree-vectorize -funroll-loops -march=native (the reference's own, src/makefile:159), OpenMP"
STRICT_FLAGS = "-O2 -ffp-contract=off -fno-fast-math, OpenMP"


def use_fast_build() -> str:
    """FOR TIMING ONLY (bench.py's cpu_baseline / --impl reference legs): rebuilds the restatement on THIS host with the
    flags the reference builds itself with (oracle/Makefile, target `fast`) and routes every later call of this module
    through it.  Results are then no longer bit-identical to the strict build, which is the checker."""
    global _lib, _SO
    fast = os.path.join(_HERE, "_build", "libicar_oracle_fast.so")
    subprocess.check_call(["make", "-C", _HERE, "-B", "fast"], stdout=subprocess.DEVNULL)
    _SO = fast
    _lib = None
    return FAST_FLAGS


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _suffix(dtype) -> tuple[str, type]:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(dtype)


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _dims(a):
    ny, nz, nx = a.shape
    return nx, nz, ny


def set_threads(n: int) -> None:
    lib().oracle_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def wind_normalise(u, v, w, jac_u, jac_v, jac_w, rho, dt, dx, scheme=2):
    """adv_mpdata.f90:500-506.  u (ny,nz,nx+1), v (ny+1,nz,nx), w (ny,nz,nx).
    Returns compact U_m (ny,nz,nx-1), V_m (ny-1,nz,nx), W_m (ny,nz,nx)."""
    nx, nz, ny = _dims(w)
    suf, ct = _suffix(w.dtype)
    U = np.zeros((ny, nz, nx - 1), w.dtype)
    V = np.zeros((ny - 1, nz, nx), w.dtype)
    W = np.zeros((ny, nz, nx), w.dtype)
    getattr(lib(), f"oracle_wind_normalise_{suf}")(
        _p(u), _p(v), _p(w), _p(jac_u), _p(jac_v), _p(jac_w), _p(rho), ct(dt), ct(dx),
        nx, nz, ny, int(scheme), _p(U), _p(V), _p(W))
    return U, V, W


def upwind_advection(qin, U, V, W, rho, dz, jaco):
    nx, nz, ny = _dims(qin)
    suf, _ = _suffix(qin.dtype)
    q = np.empty_like(qin)
    getattr(lib(), f"oracle_upwind_advection_{suf}")(
        _p(qin), _p(U), _p(V), _p(W), _p(rho), _p(q), _p(dz), nx, nz, ny, _p(jaco))
    return q


def mpdata_fluxes(q, U, V, Wn, G):
    nx, nz, ny = _dims(q)
    suf, _ = _suffix(q.dtype)
    u2 = np.empty_like(U)
    v2 = np.empty_like(V)
    w2 = np.empty_like(q)
    getattr(lib(), f"oracle_mpdata_fluxes_{suf}")(
        _p(q), _p(U), _p(V), _p(Wn), _p(u2), _p(v2), _p(w2), nx, nz, ny, _p(G))
    return u2, v2, w2


def flux_limiter(q, q2, u2, v2, w2):
    """In place on u2, v2, w2 (adv_mpdata.f90:257-354)."""
    nx, nz, ny = _dims(q)
    suf, _ = _suffix(q.dtype)
    getattr(lib(), f"oracle_flux_limiter_{suf}")(_p(q), _p(q2), _p(u2), _p(v2), _p(w2), nx, nz, ny)


def advect3d(q, rho, jaco, dz, U, V, W, order=2, fct=True):
    """In place on q (adv_mpdata.f90:356-418)."""
    nx, nz, ny = _dims(q)
    suf, _ = _suffix(q.dtype)
    getattr(lib(), f"oracle_advect3d_{suf}")(
        _p(q), _p(rho), nx, nz, ny, _p(jaco), _p(dz), _p(U), _p(V), _p(W), int(order), int(bool(fct)))


def mp_simple_driver(p, th, pii, rho, qv, qc, qr, qs, rain, snow, dt, dz,
                     its, ite, jts, jte, kts, kte):
    """In place (mp_simple.f90:595-646).  Tile bounds are 0-based inclusive.
    Returns (not_converged, out_of_bounds, negative_stop) diagnostic counters."""
    nx, nz, ny = _dims(th)
    suf, ct = _suffix(th.dtype)
    diag = (C.c_longlong * 3)()
    getattr(lib(), f"oracle_mp_simple_driver_{suf}")(
        _p(p), _p(th), _p(pii), _p(rho), _p(qv), _p(qc), _p(qr), _p(qs), _p(rain), _p(snow),
        ct(dt), _p(dz), nx, nz, ny, int(its), int(ite), int(jts), int(jte), int(kts), int(kte), diag)
    return tuple(int(x) for x in diag)


def mp_simple_column(p, T, rho, qv, qc, qr, qs, rain, snow, dt, dz, kts, kte):
    """One column on temperature (not theta), as src/tests/test_mp_simple.f90 drives it.
    rain/snow are 1-element arrays updated in place."""
    nz = T.shape[0]
    suf, ct = _suffix(T.dtype)
    getattr(lib(), f"oracle_mp_simple_column_{suf}")(
        _p(p), _p(T), _p(rho), _p(qv), _p(qc), _p(qr), _p(qs), _p(rain), _p(snow), ct(dt), _p(dz),
        nz, int(kts), int(kte))


def step_simple(st, mp_dt, dt, order=2, fct=True, scratch=None):
    """One mp_simple + MPDATA model step on a dict of float32 arrays (the CPU baseline)."""
    nx, nz, ny = _dims(st["th"])
    if scratch is None:
        scratch = (np.empty((ny, nz, nx - 1), np.float32), np.empty((ny - 1, nz, nx), np.float32),
                   np.empty((ny, nz, nx), np.float32))
    f = C.c_float
    lib().oracle_step_simple_f32(
        _p(st["pressure"]), _p(st["th"]), _p(st["pii"]), _p(st["rho"]), _p(st["qv"]), _p(st["qc"]),
        _p(st["qr"]), _p(st["qs"]), _p(st["rain"]), _p(st["snow"]), _p(st["dz_mass"]),
        _p(st["u"]), _p(st["v"]), _p(st["w"]), _p(st["jacobian_u"]), _p(st["jacobian_v"]),
        _p(st["jacobian_w"]), _p(st["jacobian"]), _p(st["advection_dz"]), f(mp_dt), f(dt),
        f(st["dx"]), nx, nz, ny, int(order), int(bool(fct)), _p(scratch[0]), _p(scratch[1]),
        _p(scratch[2]))
    return scratch


# ------------------------------------------------------------------ Thompson (mp = 1)
class ThompsonOptions(C.Structure):
    """mp_options_type (src/objects/opt_types.f90:31-45), defaults of options_obj.f90:1259-1280."""
    _fields_ = [(n, C.c_float) for n in (
        "Nt_c", "TNO", "am_s", "rho_g", "av_s", "bv_s", "fv_s", "av_i", "av_g", "bv_g",
        "Ef_si", "Ef_rs", "Ef_rg", "Ef_ri", "C_cubes", "C_sqrd", "mu_r", "t_adjust")] + \
        [("Ef_rw_l", C.c_int), ("Ef_sw_l", C.c_int)]

    def __init__(self, **kw):
        super().__init__(100.e6, 5.0, 0.069, 500.0, 40.0, 0.55, 100.0, 1847.5, 442.0, 0.89,
                         0.05, 0.95, 0.75, 0.95, 0.5, 0.3, 0.0, 0.0, 0, 0)
        for k, v in kw.items():
            setattr(self, k, v)


# order of oracle_thompson_table ids == allocation order in thompson_init (mp_thompson.f90:353-390)
THOMPSON_TABLES = (
    "tcg_racg", "tmr_racg", "tcr_gacr", "tmg_gacr", "tnr_racg", "tnr_gacr",
    "tcs_racs1", "tmr_racs1", "tcs_racs2", "tmr_racs2", "tcr_sacr1", "tms_sacr1",
    "tcr_sacr2", "tms_sacr2", "tnr_racs1", "tnr_racs2", "tnr_sacr1", "tnr_sacr2",
    "tpi_qcfz", "tni_qcfz", "tpi_qrfz", "tpg_qrfz", "tni_qrfz", "tnr_qrfz",
    "tps_iaus", "tni_iaus", "tpi_ide", "t_Efrw", "t_Efsw")
# Fortran shapes (column-major); numpy views are the reversed shape in C order
THOMPSON_SHAPES = dict(
    [(n, (28, 28, 37, 37)) for n in THOMPSON_TABLES[0:6]] +
    [(n, (28, 9, 37, 37)) for n in THOMPSON_TABLES[6:18]] +
    [(n, (37, 45)) for n in THOMPSON_TABLES[18:20]] +
    [(n, (37, 37, 45)) for n in THOMPSON_TABLES[20:24]] +
    [(n, (64, 55)) for n in THOMPSON_TABLES[24:27]] +
    [(n, (100, 100)) for n in THOMPSON_TABLES[27:29]])


class Thompson:
    """thompson_init + mp_gt_driver of the CPU restatement (oracle/mp_thompson_oracle.hpp).
    big_tables=False skips qr_acr_qg / qr_acr_qs (1.4e10 inner iterations); single (k,m) slabs of
    them can then be filled with big_slab()."""

    def __init__(self, options: ThompsonOptions | None = None, big_tables: bool = True):
        l = lib()
        l.oracle_thompson_create.restype = C.c_void_p
        l.oracle_thompson_table.restype = C.c_longlong
        l.oracle_thompson_driver.restype = C.c_longlong
        self.options = options or ThompsonOptions()
        self._h = C.c_void_p(l.oracle_thompson_create(C.byref(self.options), int(big_tables)))

    def __del__(self):
        try:
            lib().oracle_thompson_destroy(self._h)
        except Exception:
            pass

    def table(self, name):
        """numpy VIEW (C order, reversed Fortran shape) on the oracle's table."""
        p = C.POINTER(C.c_double)()
        n = lib().oracle_thompson_table(self._h, THOMPSON_TABLES.index(name), C.byref(p))
        return np.ctypeslib.as_array(p, (n,)).reshape(THOMPSON_SHAPES[name][::-1])

    def big_slab(self, which, km, ij0=0, ij1=None):
        """Fill cells [ij0, ij1) of slab km = (m-1)*37 + (k-1) of qr_acr_qg (0) or qr_acr_qs (1)."""
        if ij1 is None:
            ij1 = 28 * 28 if which == 0 else 28 * 9
        lib().oracle_thompson_big_slab(self._h, int(which), int(km), int(ij0), int(ij1))

    def constants(self):
        out = (C.c_float * 256)()
        n = lib().oracle_thompson_constants(self._h, out)
        return np.array(out[:n], np.float32)

    def bins(self):
        out = np.empty((5, 100), np.float64)
        lib().oracle_thompson_bins(self._h, _p(out))
        return dict(zip(("Dc", "Di", "Dr", "Ds", "Dg"), out))

    def driver(self, st, dt, its, ite, jts, jte, kts, kte, ide0=None, jde0=None):
        """mp_gt_driver in place on a state dict with th, qv, qc, qr, qi, qs, qg, ni, nr, pii,
        pressure, dz_mass and 2-D rainnc, snownc, graupelnc, sr (float32).  0-based inclusive bounds.
        Returns the number of columns that took the no_micro early exit."""
        nx, nz, ny = _dims(st["th"])
        ide0 = nx - 1 if ide0 is None else ide0
        jde0 = ny - 1 if jde0 is None else jde0
        return int(lib().oracle_thompson_driver(
            self._h, _p(st["qv"]), _p(st["qc"]), _p(st["qr"]), _p(st["qi"]), _p(st["qs"]), _p(st["qg"]),
            _p(st["ni"]), _p(st["nr"]), _p(st["th"]), _p(st["pii"]), _p(st["pressure"]), _p(st["dz_mass"]),
            C.c_float(dt), _p(st["rainnc"]), _p(st.get("rainncv")), _p(st.get("snownc")),
            _p(st.get("graupelnc")), _p(st.get("sr")), nx, nz, ny, int(its), int(ite), int(jts), int(jte),
            int(kts), int(kte), int(ide0), int(jde0)))

    def column(self, col, dt):
        """mp_thompson on one column dict (qv,qc,qi,qr,qs,qg,ni,nr,T,p,dz float32 1-D); returns
        (no_micro, ppt[rain, snow, graupel, ice])."""
        ppt = np.zeros(4, np.float32)
        nk = col["T"].shape[0]
        r = lib().oracle_thompson_column(
            self._h, _p(col["qv"]), _p(col["qc"]), _p(col["qi"]), _p(col["qr"]), _p(col["qs"]), _p(col["qg"]),
            _p(col["ni"]), _p(col["nr"]), _p(col["T"]), _p(col["p"]), _p(col["dz"]), _p(ppt), nk, C.c_float(dt))
        return bool(r), ppt

    def step(self, st, mp_dt, dt, order=2, fct=True, scratch=None):
        """One Thompson + MPDATA model step on a state dict (the CPU baseline for BASELINE configs[2]).
        st needs the 9 scalars, the metric / wind fields and float64 precipitation, snowfall, graupel."""
        nx, nz, ny = _dims(st["th"])
        if scratch is None:
            scratch = (np.empty((ny, nz, nx - 1), np.float32), np.empty((ny - 1, nz, nx), np.float32),
                       np.empty((ny, nz, nx), np.float32))
        f = C.c_float
        lib().oracle_step_thompson_f32(
            self._h, _p(st["pressure"]), _p(st["th"]), _p(st["pii"]), _p(st["qv"]), _p(st["qc"]), _p(st["qr"]),
            _p(st["qi"]), _p(st["qs"]), _p(st["qg"]), _p(st["ni"]), _p(st["nr"]), _p(st["precipitation"]),
            _p(st["snowfall"]), _p(st["graupel"]), _p(st["sr"]), _p(st["dz_mass"]), _p(st["u"]), _p(st["v"]),
            _p(st["w"]), _p(st["jacobian_u"]), _p(st["jacobian_v"]), _p(st["jacobian_w"]), _p(st["jacobian"]),
            _p(st["advection_dz"]), f(mp_dt), f(dt), f(st["dx"]), nx, nz, ny, int(order), int(bool(fct)),
            _p(scratch[0]), _p(scratch[1]), _p(scratch[2]))
