"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Second, independent restatement in NumPy.

The reference Fortran is written as whole-array slice expressions; this file
transcribes those slice expressions one-to-one (1-based inclusive Fortran
slices -> 0-based NumPy slices on arrays of shape (ny, nz, nx)), whereas
oracle/mpdata_oracle.hpp was written as per-cell loops.  The two were written
separately; tests/test_oracle_cross.py requires them to agree BIT FOR BIT in
fp32, which catches index-translation mistakes in either.  The flux limiter
is written here in its "independent faces" form rather
than the Fortran's carried-variable loop, which also proves that the carried
temporaries are pure reuse -- the property the CUDA kernels rely on.

Parity: the C++ oracle this file must agree with is pinned to the reference source executed by oracle/f90run
(see mpdata_oracle.hpp); no reference golden vectors exist.

Follows /root/reference/src/physics/adv_mpdata.f90 and adv_mpdata_FCT_core.f90.
"""
from __future__ import annotations

import numpy as np


def flux1(l, r, U):
    """adv_mpdata.f90:41"""
    return ((U + np.abs(U)) * l + (U - np.abs(U)) * r) / 2


def upwind_advection(qin, u, v, w, rho, dz, jaco):
    """adv_mpdata.f90:44-105.  u (ny,nz,nx-1) v (ny-1,nz,nx) w (ny,nz,nx)."""
    ny, nz, nx = qin.shape
    q = qin.copy()
    J = slice(1, ny - 1)                      # do i=jms+1,jme-1
    X = slice(1, nx - 1)                      # ims+1:ime-1
    f1 = flux1(qin[J, :, 0:nx - 1], qin[J, :, 1:nx], u[J, :, :])
    f3 = flux1(qin[J, :, X], qin[2:ny, :, X], v[1:ny - 1, :, X])      # v(:,:,i+1)
    f4 = flux1(qin[0:ny - 2, :, X], qin[J, :, X], v[0:ny - 2, :, X])  # v(:,:,i)
    f5 = flux1(qin[J, 0:nz - 1, X], qin[J, 1:nz, X], w[J, 0:nz - 1, X])
    q[J, :, X] = q[J, :, X] - ((f1[:, :, 1:nx - 1] - f1[:, :, 0:nx - 2]) + (f3 - f4)) \
        / (jaco[J, :, X] * rho[J, :, X])
    q[J, 1:nz - 1, X] = q[J, 1:nz - 1, X] - (f5[:, 1:nz - 1, :] - f5[:, 0:nz - 2, :]) \
        / (dz[J, 1:nz - 1, X] * jaco[J, 1:nz - 1, X] * rho[J, 1:nz - 1, X])
    q[J, 0, X] = q[J, 0, X] - f5[:, 0, :] / (dz[J, 0, X] * jaco[J, 0, X] * rho[J, 0, X])
    q[J, nz - 1, X] = q[J, nz - 1, X] - (qin[J, nz - 1, X] * w[J, nz - 1, X] - f5[:, nz - 2, :]) \
        / (dz[J, nz - 1, X] * jaco[J, nz - 1, X] * rho[J, nz - 1, X])
    return q


def mpdata_fluxes(q, u, v, w, G):
    """adv_mpdata.f90:107-255; w is W_m/dz.  Loops over (j,k) like the Fortran,
    slice expressions in x."""
    ny, nz, nx = q.shape
    dt = q.dtype
    u2 = np.zeros_like(u)
    v2 = np.zeros_like(v)
    w2 = np.zeros_like(w)
    quarter = dt.type(1) / dt.type(4.0)
    half = dt.type(0.5)
    eps = dt.type(np.float32(1e-10))
    for i in range(ny):            # Fortran i (y)
        for j in range(nz):        # Fortran j (z)
            # ---- U
            rx = q[i, j, 1:nx]
            lx = q[i, j, 0:nx - 1]
            denomx = rx + lx + eps
            GG = G[i, j, 1:nx] + G[i, j, 0:nx - 1]
            uu = u[i, j, :]
            t = np.abs(uu) * (1 - np.abs(uu) / (half * GG))
            t = t * (rx - lx) / denomx
            if 0 < i < ny - 1:
                eq = (q[i + 1, j, 1:nx] - q[i - 1, j, 1:nx] + q[i + 1, j, 0:nx - 1] - q[i - 1, j, 0:nx - 1]) / \
                     (q[i + 1, j, 1:nx] + q[i - 1, j, 1:nx] + q[i + 1, j, 0:nx - 1] + q[i - 1, j, 0:nx - 1] + eps)
                # v(ims+1:ime,j,i) -> compact row i-1 ; v(..,i+1) -> row i
                ev = quarter * (v[i - 1, j, 1:nx] + v[i, j, 1:nx] + v[i - 1, j, 0:nx - 1] + v[i, j, 0:nx - 1])
                t = t - half * uu * ev * eq / GG
            if 0 < j < nz - 1:
                eq = (q[i, j + 1, 1:nx] - q[i, j - 1, 1:nx] + q[i, j + 1, 0:nx - 1] - q[i, j - 1, 0:nx - 1]) / \
                     (q[i, j + 1, 1:nx] + q[i, j - 1, 1:nx] + q[i, j + 1, 0:nx - 1] + q[i, j - 1, 0:nx - 1] + eps)
                ev = quarter * (w[i, j, 1:nx] + w[i, j - 1, 1:nx] + w[i, j, 0:nx - 1] + w[i, j - 1, 0:nx - 1])
                t = t - half * uu * ev * eq / GG
            u2[i, j, :] = t
            # ---- V  (compact row i-1 is Fortran v(:,j,i))
            if i > 0:
                r = q[i, j, :]
                l = q[i - 1, j, :]
                denom = r + l + eps
                GG = G[i, j, :] + G[i - 1, j, :]
                vv = v[i - 1, j, :]
                t = np.abs(vv) * (1 - np.abs(vv) / (half * GG))
                t = t * (r - l) / denom
                edge_q = np.zeros(nx, dt)
                edge_v = np.zeros(nx, dt)
                edge_q[1:nx - 1] = (q[i - 1, j, 2:nx] - q[i, j, 0:nx - 2] + q[i, j, 2:nx] - q[i - 1, j, 0:nx - 2]) / \
                                   (q[i, j, 2:nx] + q[i - 1, j, 2:nx] + q[i, j, 0:nx - 2] + q[i - 1, j, 0:nx - 2] + eps)
                # u(ims+2:ime) -> compact 1:nx-1 ; u(ims+1:ime-1) -> compact 0:nx-2
                edge_v[1:nx - 1] = quarter * (u[i, j, 1:nx - 1] + u[i - 1, j, 1:nx - 1] + u[i, j, 0:nx - 2] + u[i - 1, j, 0:nx - 2])
                t = t - half * vv * edge_v * edge_q / GG
                if 0 < j < nz - 1:
                    eq = (q[i - 1, j + 1, :] - q[i, j - 1, :] + q[i, j + 1, :] - q[i - 1, j - 1, :]) / \
                         (q[i - 1, j + 1, :] + q[i, j - 1, :] + q[i, j + 1, :] + q[i - 1, j - 1, :] + eps)
                    ev = quarter * (w[i, j, :] + w[i, j - 1, :] + w[i - 1, j, :] + w[i - 1, j - 1, :])
                    t = t - half * vv * ev * eq / GG
                v2[i - 1, j, :] = t
            # ---- W
            if j == nz - 1:
                w2[i, j, :] = 0
            else:
                r = q[i, j + 1, :]
                l = q[i, j, :]
                denom = r + l + eps
                GG = G[i, j + 1, :] + G[i, j, :]
                ww = w[i, j, :]
                t = np.abs(ww) * (1 - np.abs(ww) / (half * GG))
                t = t * (r - l) / denom
                edge_q = np.zeros(nx, dt)
                edge_v = np.zeros(nx, dt)
                edge_q[1:nx - 1] = (q[i, j + 1, 2:nx] - q[i, j, 0:nx - 2] + q[i, j, 2:nx] - q[i, j + 1, 0:nx - 2]) / \
                                   (q[i, j, 2:nx] + q[i, j + 1, 2:nx] + q[i, j, 0:nx - 2] + q[i, j + 1, 0:nx - 2] + eps)
                edge_v[1:nx - 1] = quarter * (u[i, j, 1:nx - 1] + u[i, j + 1, 1:nx - 1] + u[i, j, 0:nx - 2] + u[i, j + 1, 0:nx - 2])
                t = t - half * ww * edge_v * edge_q / GG
                if 0 < i < ny - 1:
                    eq = (q[i + 1, j + 1, :] - q[i - 1, j, :] + q[i + 1, j, :] - q[i - 1, j + 1, :]) / \
                         (q[i + 1, j, :] + q[i - 1, j + 1, :] + q[i + 1, j + 1, :] + q[i - 1, j, :] + eps)
                    # v(:,j,i)->row i-1, v(:,j+1,i)->row i-1, v(:,j,i+1)->row i, v(:,j+1,i+1)->row i
                    ev = quarter * (v[i - 1, j, :] + v[i - 1, j + 1, :] + v[i, j, :] + v[i, j + 1, :])
                    t = t - half * ww * ev * eq / GG
                w2[i, j, :] = t
    return u2, v2, w2


def fct_lines(q1, l, U2, is_w, axis):
    """Independent-faces form of adv_mpdata_FCT_core.f90:47-116 applied along `axis`
    of whole arrays.  q1, l have n cells along axis; U2 has n-1 faces.  Returns limited U2."""
    q1 = np.moveaxis(q1, axis, -1)
    l = np.moveaxis(l, axis, -1)
    U2 = np.moveaxis(U2, axis, -1)
    dt = q1.dtype
    n = q1.shape[-1]
    f = flux1(q1[..., :n - 1], q1[..., 1:], U2)
    hi = np.empty_like(q1)
    lo = np.empty_like(q1)
    m_hi = np.maximum(q1, l)
    m_lo = np.minimum(q1, l)
    hi[..., 1:n - 1] = np.maximum(np.maximum(m_hi[..., 0:n - 2], m_hi[..., 1:n - 1]), m_hi[..., 2:n])
    lo[..., 1:n - 1] = np.minimum(np.minimum(m_lo[..., 0:n - 2], m_lo[..., 1:n - 1]), m_lo[..., 2:n])
    hi[..., 0] = np.maximum(m_hi[..., 0], m_hi[..., 1])
    lo[..., 0] = np.minimum(m_lo[..., 0], m_lo[..., 1])
    hi[..., n - 1] = np.maximum(np.maximum(q1[..., n - 2], q1[..., n - 1]), l[..., n - 2])
    lo[..., n - 1] = np.minimum(np.minimum(q1[..., n - 2], q1[..., n - 1]), l[..., n - 2])
    zero = dt.type(0)
    fin = np.zeros_like(q1)
    fout = np.zeros_like(q1)
    fin[..., 1:n - 1] = np.maximum(zero, f[..., 0:n - 2]) - np.minimum(zero, f[..., 1:n - 1])
    fout[..., 1:n - 1] = np.maximum(zero, f[..., 1:n - 1]) - np.minimum(zero, f[..., 0:n - 2])
    if is_w:
        fin[..., 0] = zero - np.minimum(zero, f[..., 0])
        fout[..., 0] = np.maximum(zero, f[..., 0])
        fin[..., n - 1] = np.maximum(zero, f[..., n - 2]) - np.minimum(zero, f[..., n - 2])
        fout[..., n - 1] = fin[..., n - 1]
    e15 = dt.type(np.float32(1e-15))
    one = dt.type(1)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        beta_in = (hi - q1) / (fin + e15)
        beta_out = (q1 - lo) / (fout + e15)
    pos = np.minimum(np.minimum(one, beta_in[..., 1:]), beta_out[..., :n - 1]) * U2
    neg = np.minimum(np.minimum(one, beta_in[..., :n - 1]), beta_out[..., 1:]) * U2
    out = np.where(U2 > 0, pos, np.where(U2 < 0, neg, U2))
    return np.moveaxis(out, -1, axis)


def flux_limiter(q, q2, u, v, w):
    """adv_mpdata.f90:257-354 coverage rules.  Returns new (u, v, w)."""
    ny, nz, nx = q.shape
    u = u.copy()
    v = v.copy()
    w = w.copy()
    J = slice(1, ny - 1)
    u[J] = fct_lines(q2[J], q[J], u[J], False, axis=2)                       # x lines, j interior, all k
    X = slice(1, nx - 1)
    w[J, 0:nz - 1, X] = fct_lines(q2[J, :, X], q[J, :, X], w[J, 0:nz - 1, X], True, axis=1)
    w[J, nz - 1, X] = 0
    v[:] = fct_lines(q2, q, v, False, axis=0)                                # y lines, ALL i, k
    return u, v, w


def advect3d(q, rho, jaco, dz, U_m, V_m, W_m, order=2, fct=True):
    """adv_mpdata.f90:356-418; returns the new q."""
    q = q.copy()
    q2 = None
    for iord in range(1, order + 1):
        if iord == 1:
            q2 = upwind_advection(q, U_m, V_m, W_m, rho, dz, jaco)
        else:
            u2, v2, w2 = mpdata_fluxes(q2, U_m, V_m, W_m / dz, jaco * rho)
            u2 = u2 * q.dtype.type(0.5)
            v2 = v2 * q.dtype.type(0.5)
            w2 = w2 * q.dtype.type(0.5) * dz
            if fct:
                u2, v2, w2 = flux_limiter(q, q2, u2, v2, w2)
            q = upwind_advection(q2, u2, v2, w2, rho, dz, jaco)
        if iord != order:
            if iord > 1:
                q2 = q.copy()
        elif iord == 1:
            q = q2.copy()
    return q


# ------------------------------------------------------------------ the rest of step()'s loop body
RD = np.float32(287.058)     # src/constants/icar_constants.f90:391
CP = np.float32(1012.0)      # src/constants/icar_constants.f90:393


def powf(x, y):
    """x ** y on REAL operands as the compiled Fortran evaluates it: glibc's powf, element by element (NumPy's own
    vectorised float32 power differs from it in the last bit now and then)"""
    import ctypes
    import ctypes.util
    lm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    lm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    lm.powf.restype = ctypes.c_float
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    fo, fx, yy = out.reshape(-1), x.reshape(-1), float(np.float32(y))
    for n in range(fx.size):
        fo[n] = lm.powf(float(fx[n]), yy)
    return out


def diagnostic_update_np(pressure, th):
    """exner and density as diagnostic_update computes them (src/main/time_step.f90:85,96,101;
    exner_function, src/utilities/atm_utilities.f90:682-691).  float32 throughout."""
    exner = powf(pressure / np.float32(100000), RD / CP)
    temperature = th * exner
    density = pressure / (RD * temperature)
    return exner.astype(np.float32), density.astype(np.float32)


def diagnostic_output_np(pressure, th, u, v, w, jacobian, dzdx, dzdy):
    """The rest of diagnostic_update (src/main/time_step.f90:88-108, 165-194), array expressions as written there.
    Arrays are (y, z, x): u (ny,nz,nx+1), v (ny+1,nz,nx), dzdx like u, dzdy like v.  float32 throughout."""
    f = np.float32
    exner = powf(pressure / f(100000), RD / CP)
    o = {}
    pi = np.empty_like(pressure)
    pi[:, 1:, :] = (pressure[:, :-1, :] + pressure[:, 1:, :]) / f(2)                          # :88
    pi[:, 0, :] = pressure[:, 0, :] + (pressure[:, 0, :] - pressure[:, 1, :]) / f(2)           # :89
    o["pressure_interface"] = pi
    o["surface_pressure"] = pi[:, 0, :].copy()                                                 # :93
    t = th * exner                                                                             # :96
    o["temperature"] = t
    ti = np.empty_like(t)
    ti[:, 1:, :] = (t[:, :-1, :] + t[:, 1:, :]) / f(2)                                         # :97
    ti[:, 0, :] = t[:, 0, :] + (t[:, 0, :] - t[:, 1, :]) / f(2)                                # :98
    o["temperature_interface"] = ti
    o["u_mass"] = (u[:, :, 1:] + u[:, :, :-1]) / f(2)                                          # :104
    o["v_mass"] = (v[1:, :, :] + v[:-1, :, :]) / f(2)                                          # :107
    ny, nz, nx = w.shape
    wr = np.zeros_like(w)
    lastw = np.zeros((ny - 2, nx - 2), f)
    for z in range(nz):                                                                        # :165-194
        uw = u[1:-1, z, 1:nx] * dzdx[1:-1, z, 1:nx]            # u(ims+1:ime, z, jms+1:jme-1) * dzdx(...)
        vw = v[1:ny, z, 1:-1] * dzdy[1:ny, z, 1:-1]            # v(ims+1:ime-1, z, jms+1:jme) * dzdy(...)
        currw = w[1:-1, z, 1:-1]
        wr[1:-1, z, 1:-1] = ((uw[:, :-1] + uw[:, 1:]) * f(0.5) + (vw[:-1, :] + vw[1:, :]) * f(0.5)
                             + jacobian[1:-1, z, 1:-1] * (lastw + currw) * f(0.5))
        lastw = currw
    o["w_real"] = wr
    return o


def step_loop_np(start_seconds, end_seconds, dts):
    """The control flow of icar_step (icar_b200/csrc/icar_b200.cu) = step(), time_step.f90:438-549, with the model time
    in REAL*8 seconds: update_dt hands out the scripted `dts`; returns (log, final time) in the format of
    oracle.f90run.Reference.step_loop -- the same routine names in the same order -- so that the two can be compared."""
    log, script = [], list(dts)
    t = float(start_seconds)
    end = float(end_seconds)
    while t < end:                                                       # :464
        log.append(("update_dt",))                                       # :466
        dt = float(script.pop(0))
        if t + dt > end:                                                 # :469-471
            dt = end - t
        log.append(("diagnostic_update",))                               # :474
        if dt > 1e-3:                                                    # :484
            r = float(np.float32(dt))                                    # real(dt%seconds())
            for name in ("rad", "lsm", "pbl", "convect"):                # :489-509 (0 on this path; the order is the source's)
                log.append((name, r))
            log.append(("mp", r, ("halo", 1)))                           # :512
            log.append(("halo_send",))                                   # :515
            log.append(("mp", r, ("subset", 1)))                         # :523
            log.append(("halo_retrieve",))                               # :526
            log.append(("advect", r))                                    # :529
            log.append(("apply_forcing", dt))                            # :537
            if (end - t) < dt * 2:                                       # :540-542
                log.append(("enforce_limits",))
        t = t + dt                                                       # :547
    return log, t


def compute_dt_np(dx, u, v, w, dz_levels, cfl, cfl_strictness):
    """compute_dt (src/main/time_step.f90:217-330) with use_density = .false., then update_dt's 120 s cap
    (:421).  u (ny,nz,nx+1), v (ny+1,nz,nx), w (ny,nz,nx); REAL arithmetic.  Raises where the reference
    stops ("ERROR time step too small")."""
    f = np.float32
    au, av, aw = np.abs(u), np.abs(v), np.abs(w)
    sqrt3 = f(np.sqrt(f(3.0))) * f(1.001)
    three_d_cfl = f(0.577350269)
    if cfl_strictness == 1:
        maxwind1d = max(au.max(), av.max(), aw.max())
        maxwind3d = f(maxwind1d) * sqrt3
    elif cfl_strictness == 5:
        maxwind3d = f(au.max()) + f(av.max()) + f(aw.max())
    else:
        wl = aw.copy()
        wl[:, 1:, :] = np.maximum(aw[:, 1:, :], aw[:, :-1, :])          # zoffset = -1 for k > 1
        cur = (np.maximum(au[:, :, :-1], au[:, :, 1:]) / f(dx)
               + np.maximum(av[:-1], av[1:]) / f(dx)
               + wl / np.asarray(dz_levels, f).reshape(1, -1, 1))
        maxwind3d = f(cur.max())
        if cfl_strictness == 2:
            maxwind3d = maxwind3d * three_d_cfl
            maxwind3d = max(f(max(au.max(), av.max(), aw.max())), maxwind3d)
        elif cfl_strictness == 4:
            maxwind3d = maxwind3d * sqrt3
    dt = f(cfl) / f(maxwind3d)
    if dt < f(1e-1):
        raise FloatingPointError("ERROR time step too small")
    return min(float(dt), 120.0)


def enforce_limits_np(fields):
    """domain%enforce_limits (src/objects/domain_obj.f90:2228-2243), in place."""
    for a in fields:
        a[a < 0] = 0


def apply_forcing_np(q, dqdt, seconds, boundaries_only, west=True, east=True, south=True, north=True):
    """domain%apply_forcing for one 3-D variable (src/objects/domain_obj.f90:2404-2426), in place:
    q = q + dqdt * dt%seconds() with the product and sum in REAL*8, rounded once on assignment."""
    new = (q.astype(np.float64) + dqdt.astype(np.float64) * float(seconds)).astype(np.float32)
    if not boundaries_only:
        q[...] = new
        return
    if west:
        q[1:-1, :, 0] = new[1:-1, :, 0]
    if east:
        q[1:-1, :, -1] = new[1:-1, :, -1]
    if south:
        q[0] = new[0]
    if north:
        q[-1] = new[-1]
