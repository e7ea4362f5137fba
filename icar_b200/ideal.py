"""Synthetic ideal "hill" case: the inputs of BASELINE.json's configs.

Restates (as input generation only -- nothing here is on the hot path):
  * helpers/gen_init_ideal.py:26-102      terrain, sounding (th, qv, p), u=10 m/s
  * src/objects/domain_obj.f90:1200-1316  setup_simple_z (Gal-Chen terrain-following levels)
  * src/objects/domain_obj.f90:1368-1385  jacobian_u / jacobian_v / jacobian_w
  * src/physics/wind.f90:81-228,528-534   balance_uvw / calc_divergence, advection_dz
  * src/main/time_step.f90:85-101         exner, temperature, density
  * src/main/time_step.f90:217-330,421    compute_dt (cfl_strictness 3, factor 0.9, cap 120 s)
  * src/utilities/atm_utilities.f90:682   exner_function

All arrays are float32 in the reference's memory order: Fortran (i,k,j) column-major
== NumPy C-order shape (ny, nz, nx).  u is (ny,nz,nx+1), v is (ny+1,nz,nx).
"""
from __future__ import annotations

import numpy as np

RD = np.float32(287.058)   # src/constants/icar_constants.f90:391
CP = np.float32(1012.0)    # src/constants/icar_constants.f90:393

F = np.float32


def balance_uvw(u, v, jac_u, jac_v, jac_w, dz, dx):
    """wind.f90:81-147 + calc_divergence(horz_only) :173-228 -> w that balances u, v."""
    ny, nz, nx = jac_w.shape
    u_met = u * jac_u
    v_met = v * jac_v
    div = ((u_met[:, :, 1:] - u_met[:, :, :-1]) + (v_met[1:, :, :] - v_met[:-1, :, :])) / F(dx)
    w = np.zeros((ny, nz, nx), np.float32)
    for k in range(nz):
        if k == 0:
            w[:, k, :] = F(0) - div[:, k, :] * dz[:, k, :] / jac_w[:, k, :]
        else:
            w[:, k, :] = (w[:, k - 1, :] * jac_w[:, k - 1, :] - div[:, k, :] * dz[:, k, :]) / jac_w[:, k, :]
    return w


def compute_dt(u, v, w, dx, dz_levels, cfl=0.9, cap=120.0):
    """time_step.f90:264-318 with cfl_strictness=3, then the 120 s cap of :421."""
    ny, nz, nx = w.shape
    au, av, aw = np.abs(u), np.abs(v), np.abs(w)
    wl = aw.copy()
    wl[:, 1:, :] = np.maximum(aw[:, 1:, :], aw[:, :-1, :])
    cur = (np.maximum(au[:, :, :-1], au[:, :, 1:]) / F(dx)
           + np.maximum(av[:-1, :, :], av[1:, :, :]) / F(dx)
           + wl / dz_levels.reshape(1, nz, 1))
    return float(min(F(cfl) / cur.max(), cap))


def make_ideal_case(nx=100, nz=20, ny=100, dx=4000.0, dz_level=400.0, terrain=True,
                    u0=10.0, v0=0.0, perturb=False, seed=1234, rows=None):
    """Returns a dict of float32 arrays + scalars describing one ICAR domain state.

    terrain=False is "Set A" (flat metric, jacobians == 1, space_varying_dz=.false.);
    terrain=True is "Set B" (the hill, space_varying_dz=.true., flat_z_height=-1).
    perturb=True adds the reproducible noise + v wind used for the throughput runs
    (SURVEY.md section 8d).
    rows=(j0, j1): only rows j0..j1-1 of the (ny-row) domain are built -- the same values the whole state has there
    (the case is uniform in y except for the noise, whose random stream is advanced to the band) -- so that a rank of a
    large decomposed run never holds the global state (2048 x 2048 x 60: 1 GB per field)."""
    ny_global = ny
    j0, j1 = (0, ny) if rows is None else rows
    if not (0 <= j0 < j1 <= ny_global):
        raise ValueError("rows outside the domain")
    ny = j1 - j0
    dzl = np.full(nz, dz_level, np.float32)
    i = np.arange(nx, dtype=np.float64)
    hgt1 = (1000.0 * (np.cos(i / nx * 2 * np.pi + np.pi) + 1)).astype(np.float32)
    if not terrain:
        hgt1[:] = 0
    hgt = np.broadcast_to(hgt1.reshape(1, nx), (ny, nx)).copy()

    # --- setup_simple_z
    jac = np.ones((ny, nz, nx), np.float32)
    if terrain:
        max_level = nz - 1                      # find_flat_model_level with flat_z_height=-1
        H = F(dzl[:max_level].sum())
        if not H > hgt.max():
            raise ValueError(f"terrain ({hgt.max()} m) reaches the flat level ({H} m): use more levels")
        jac[:, :max_level, :] = ((H - hgt) / H).reshape(ny, 1, nx)
    dz_interface = dzl.reshape(1, nz, 1) * jac
    dz_mass = np.empty_like(jac)
    dz_mass[:, 0, :] = dzl[0] / F(2) * jac[:, 0, :]
    for k in range(1, nz):
        dz_mass[:, k, :] = dzl[k] / F(2) * jac[:, k, :] + dzl[k - 1] / F(2) * jac[:, k - 1, :]
    z = np.empty_like(jac)
    z[:, 0, :] = hgt + dz_mass[:, 0, :]
    for k in range(1, nz):
        z[:, k, :] = z[:, k - 1, :] + dz_mass[:, k, :]

    # --- staggered jacobians (domain_obj.f90:1368-1385)
    jac_u = np.empty((ny, nz, nx + 1), np.float32)
    jac_u[:, :, 0] = jac[:, :, 0]
    jac_u[:, :, nx] = jac[:, :, nx - 1]
    jac_u[:, :, 1:nx] = (jac[:, :, 1:] + jac[:, :, :-1]) / F(2)
    jac_v = np.empty((ny + 1, nz, nx), np.float32)
    jac_v[0] = jac[0]
    jac_v[ny] = jac[ny - 1]
    jac_v[1:ny] = (jac[1:] + jac[:-1]) / F(2)
    jac_w = np.ones((ny, nz, nx), np.float32)
    jac_w[:, :nz - 1, :] = (jac[:, :nz - 1, :] + jac[:, 1:, :]) / F(2)

    advection_dz = dz_interface.copy()          # wind.f90:533 (fixed_dz_advection=.false.)

    # --- sounding (gen_init_ideal.py)
    p = (F(1e5) * (F(1) - F(2.25577e-5) * z) ** F(5.25588)).astype(np.float32)
    th1 = (273.0 + 27.0 * np.arange(nz) / nz).astype(np.float32)
    th = np.broadcast_to(th1.reshape(1, nz, 1), (ny, nz, nx)).copy()
    qv = np.full((ny, nz, nx), 0.0013, np.float32)

    u = np.full((ny, nz, nx + 1), u0, np.float32)
    v = np.full((ny + 1, nz, nx), v0, np.float32)
    if perturb:
        rng = np.random.default_rng(seed)
        row, cells = nz * nx, ny_global * nz * nx            # one 64-bit draw per uniform double
        rng.bit_generator.advance(j0 * row)
        qv *= (1 + 0.1 * rng.uniform(-1, 1, qv.shape)).astype(np.float32)
        rng.bit_generator.advance(cells - j1 * row + j0 * row)
        th += (0.5 * rng.uniform(-1, 1, th.shape)).astype(np.float32)
        v[:] = F(5.0)
    w = balance_uvw(u, v, jac_u, jac_v, jac_w, advection_dz, dx)

    pii = ((p / F(1e5)) ** (RD / CP)).astype(np.float32)      # exner_function
    temperature = th * pii
    rho = (p / (RD * temperature)).astype(np.float32)

    zeros = lambda: np.zeros((ny, nz, nx), np.float32)
    st = dict(
        nx=nx, nz=nz, ny=ny, dx=float(dx), dz_levels=dzl, rows=(j0, j1), ny_global=ny_global,
        jacobian=jac, jacobian_u=jac_u, jacobian_v=jac_v, jacobian_w=jac_w,
        advection_dz=advection_dz, dz_mass=dz_mass, dz_interface=dz_interface, z=z, terrain=hgt,
        pressure=p, th=th, pii=pii, rho=rho, qv=qv, qc=zeros(), qr=zeros(), qs=zeros(),
        u=u, v=v, w=w,
        rain=np.zeros((ny, nx), np.float32), snow=np.zeros((ny, nx), np.float32),
    )
    st["dt"] = compute_dt(u, v, w, dx, dzl)
    return st


def add_thompson_fields(st, moisten=1.0):
    """Adds the extra Thompson species (qi, qg, ni, nr = 0) and the 2-D per-call precipitation arrays of
    mp_gt_driver to an ideal state; `moisten` scales qv (3.0 makes every process of the scheme fire
    within a few steps)."""
    if moisten != 1.0:
        st["qv"] = (st["qv"] * F(moisten)).astype(np.float32)
    for k in ("qi", "qg", "ni", "nr"):
        st[k] = np.zeros_like(st["qv"])
    for k in ("rainnc", "rainncv", "snownc", "snowncv", "graupelnc", "graupelncv", "sr"):
        st[k] = np.zeros(st["rain"].shape, np.float32)
    return st


def copy_state(st):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}


def slab_of(st, grid):
    """The part of a global state dict one image holds (ghost cells included) for a slab `grid`
    (x-slab or y-slab: whichever of ims..ime / jms..jme is narrower than the domain)."""
    ilo, ihi = grid.ims - 1, grid.ime
    jlo, jhi = grid.jms - 1, grid.jme
    out = {}
    for k, v in st.items():
        if not isinstance(v, np.ndarray) or v.ndim < 2:
            out[k] = v
        elif v.ndim == 2:                                   # (ny, nx)
            out[k] = np.ascontiguousarray(v[jlo:jhi, ilo:ihi])
        elif k in ("u", "jacobian_u"):                      # (ny, nz, nx+1)
            out[k] = np.ascontiguousarray(v[jlo:jhi, :, ilo:ihi + 1])
        elif k in ("v", "jacobian_v"):                      # (ny+1, nz, nx)
            out[k] = np.ascontiguousarray(v[jlo:jhi + 1, :, ilo:ihi])
        else:
            out[k] = np.ascontiguousarray(v[jlo:jhi, :, ilo:ihi])
    return out


def terrain_slopes(st):
    """domain%dzdx (ny, nz, nx+1) and domain%dzdy (ny+1, nz, nx) of setup_dzdxy (domain_obj.f90:1417-1451):
    differences of the mass-level heights over dx on the interior faces, 0 on the outermost faces."""
    z, dx = st["z"], np.float32(st["dx"])
    ny, nz, nx = z.shape
    dzdx = np.zeros((ny, nz, nx + 1), np.float32)
    dzdy = np.zeros((ny + 1, nz, nx), np.float32)
    dzdx[:, :, 1:nx] = (z[:, :, 1:] - z[:, :, :-1]) / dx
    dzdy[1:ny, :, :] = (z[1:, :, :] - z[:-1, :, :]) / dx
    return dzdx, dzdy

