"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes handle on oracle/_build/libicar_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product (icar_b200/) never does.

PARITY UNPINNED: see oracle/mpdata_oracle.hpp -- the reference (NCAR/icar) holds no
golden vectors for the MPDATA / mp_simple path and cannot be compiled in this image.

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
            ("oracle_api.cpp", "mpdata_oracle.hpp", "mp_simple_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


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
