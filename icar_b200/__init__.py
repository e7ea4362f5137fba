"""icar_b200 -- B200-native (sm_100a) MPDATA advection + cloud microphysics for ICAR.

This package is a thin ctypes binding over the C ABI in include/icar_b200.h
(libicar_b200.so, built from icar_b200/csrc by `python -c "import __graft_entry__ as g; g.build()"`).
All compute happens in the CUDA library; there is NO CPU fallback: importing works without a
GPU (so that symbols can be inspected), but creating a context raises IcarError.

The Python surface mirrors the reference's plugin call surface:
    advect(domain, options, dt)              src/physics/advection_driver.f90:51
    mp(domain, options, dt, halo, subset)    src/physics/mp_driver.f90:673
    domain%halo_send / halo_retrieve         src/objects/domain_obj.f90:104-144
with `Domain` playing domain_t (device-resident fields) and `Options` options_t.

Arrays are NumPy float32 in the reference's memory order: Fortran (i,k,j) == C-order (ny,nz,nx).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicar_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_PHYSICS, ERR_COMM = 0, 1, 2, 3, 4

# enum icar_var (include/icar_b200.h)
VAR = dict(
    u=0, v=1, w=2, pressure=3, exner=4, density=5, potential_temperature=6, water_vapor=7,
    cloud_water=8, rain_in_air=9, snow_in_air=10, jacobian=11, jacobian_u=12, jacobian_v=13,
    jacobian_w=14, advection_dz=15, dz_mass=16, precipitation=17, snowfall=18,
    cloud_ice=19, ice_number=20, rain_number=21, graupel_in_air=22, graupel=23)
# names used by icar_b200.ideal -> domain_t member names
ALIASES = dict(th="potential_temperature", qv="water_vapor", qc="cloud_water", qr="rain_in_air",
               qs="snow_in_air", pii="exner", rho="density", qi="cloud_ice", ni="ice_number",
               nr="rain_number", qg="graupel_in_air")

ADVECTED_SIMPLE = ("water_vapor", "cloud_water", "rain_in_air", "snow_in_air", "potential_temperature")


class IcarError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"icar_b200 error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    """grid_t: 1-based inclusive bounds (src/objects/grid_obj.f90:144-255)."""
    _fields_ = [(n, C.c_int) for n in (
        "ims", "ime", "kms", "kme", "jms", "jme", "its", "ite", "kts", "kte", "jts", "jte",
        "ids", "ide", "kds", "kde", "jds", "jde", "halo_size")]

    @classmethod
    def single_image(cls, nx, nz, ny, halo_size=1):
        """One image owning the whole domain: tile = everything but the outermost BC cell."""
        return cls(1, nx, 1, nz, 1, ny, 2, nx - 1, 1, nz, 2, ny - 1, 1, nx, 1, nz, 1, ny, halo_size)

    @classmethod
    def xslab(cls, nx_global, nz, ny_global, nimages, image, halo_size):
        g = cls()
        rc = lib().icar_grid_xslab(nx_global, nz, ny_global, nimages, image, halo_size, C.byref(g))
        if rc:
            raise IcarError(rc, "icar_grid_xslab: invalid decomposition")
        return g

    @classmethod
    def yslab(cls, nx_global, nz, ny_global, nimages, image, halo_size):
        g = cls()
        rc = lib().icar_grid_yslab(nx_global, nz, ny_global, nimages, image, halo_size, C.byref(g))
        if rc:
            raise IcarError(rc, "icar_grid_yslab: invalid decomposition")
        return g

    @property
    def nx(self): return self.ime - self.ims + 1
    @property
    def nz(self): return self.kme - self.kms + 1
    @property
    def ny(self): return self.jme - self.jms + 1


class Options(C.Structure):
    """The members of options_t this path reads (see icar_options in the header)."""
    _fields_ = [("advection", C.c_int), ("microphysics", C.c_int), ("mpdata_order", C.c_int),
                ("flux_corrected_transport", C.c_int), ("advect_density", C.c_int),
                ("update_interval", C.c_int), ("top_mp_level", C.c_int), ("dx", C.c_float),
                ("exact_arithmetic", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        lib().icar_default_options(C.byref(self))
        for k, v in kw.items():
            if k not in dict(self._fields_):
                raise TypeError(f"unknown option {k}")
            setattr(self, k, v)


class MpOptions(C.Structure):
    """mp_options_type (src/objects/opt_types.f90:31-45); defaults options_obj.f90:1259-1280."""
    _fields_ = [(n, C.c_float) for n in (
        "Nt_c", "TNO", "am_s", "rho_g", "av_s", "bv_s", "fv_s", "av_i", "av_g", "bv_g",
        "Ef_si", "Ef_rs", "Ef_rg", "Ef_ri", "C_cubes", "C_sqrd", "mu_r", "t_adjust")] + \
        [("Ef_rw_l", C.c_int), ("Ef_sw_l", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        lib().icar_default_mp_options(C.byref(self))
        for k, v in kw.items():
            if k not in dict(self._fields_):
                raise TypeError(f"unknown mp option {k}")
            setattr(self, k, v)


class LtOptions(C.Structure):
    """lt_parameters as spatial_winds reads them (options_obj.f90:1401-1530); defaults :1448-1479."""
    _fields_ = [("variable_N", C.c_int), ("smooth_nsq", C.c_int), ("vert_smooth", C.c_int), ("stability_window_size", C.c_int),
                ("max_stability", C.c_float), ("min_stability", C.c_float), ("N_squared", C.c_float),
                ("linear_contribution", C.c_float), ("linear_update_fraction", C.c_float),
                ("dirmax", C.c_float), ("dirmin", C.c_float), ("spdmax", C.c_float), ("spdmin", C.c_float),
                ("n_dir_values", C.c_int), ("n_nsq_values", C.c_int), ("n_spd_values", C.c_int),
                ("minimum_layer_size", C.c_float)]

    def __init__(self, **kw):
        super().__init__()
        lib().icar_default_lt_options(C.byref(self))
        for k, v in kw.items():
            if k not in dict(self._fields_):
                raise TypeError(f"unknown lt option {k}")
            setattr(self, k, v)


# lookup tables in the allocation order of thompson_init (mp_thompson.f90:353-390) + Fortran shapes
THOMPSON_TABLES = (
    "tcg_racg", "tmr_racg", "tcr_gacr", "tmg_gacr", "tnr_racg", "tnr_gacr",
    "tcs_racs1", "tmr_racs1", "tcs_racs2", "tmr_racs2", "tcr_sacr1", "tms_sacr1",
    "tcr_sacr2", "tms_sacr2", "tnr_racs1", "tnr_racs2", "tnr_sacr1", "tnr_sacr2",
    "tpi_qcfz", "tni_qcfz", "tpi_qrfz", "tpg_qrfz", "tni_qrfz", "tnr_qrfz",
    "tps_iaus", "tni_iaus", "tpi_ide", "t_Efrw", "t_Efsw")
THOMPSON_SHAPES = dict(
    [(n, (28, 28, 37, 37)) for n in THOMPSON_TABLES[0:6]] +
    [(n, (28, 9, 37, 37)) for n in THOMPSON_TABLES[6:18]] +
    [(n, (37, 45)) for n in THOMPSON_TABLES[18:20]] +
    [(n, (37, 37, 45)) for n in THOMPSON_TABLES[20:24]] +
    [(n, (64, 55)) for n in THOMPSON_TABLES[24:27]] +
    [(n, (100, 100)) for n in THOMPSON_TABLES[27:29]])


class HostDomain(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "u", "v", "w", "pressure", "exner", "density", "potential_temperature", "water_vapor",
        "cloud_water", "rain_in_air", "snow_in_air", "accumulated_precipitation",
        "accumulated_snowfall", "cloud_ice", "ice_number", "rain_number", "graupel_in_air",
        "accumulated_graupel")]


_lib = None

# every symbol include/icar_b200.h declares
SYMBOLS = (
    "icar_default_options", "icar_device_count", "icar_grid_xslab", "icar_grid_yslab", "icar_create", "icar_destroy", "icar_last_error",
    "icar_upload", "icar_download", "icar_device_ptr", "icar_set_model_time", "icar_get_model_time",
    "icar_advect", "icar_mp", "icar_halo_send", "icar_halo_retrieve", "icar_halo_exchange",
    "icar_physics_step", "icar_physics_steps", "icar_synchronize", "icar_mp_diagnostics", "icar_kernel_launches",
    "icar_selftest_division", "icar_fused_kernel_name",
    "icar_timer_start", "icar_timer_stop",
    "icar_halo_plan", "icar_comm_unique_id", "icar_comm_init", "icar_comm_init_local", "icar_mpdata_set_winds", "icar_mpdata_advect3d",
    "icar_mpdata_set_winds_f64", "icar_mpdata_advect3d_f64", "icar_mp_simple_driver",
    "icar_physics_step_host", "icar_step_host", "icar_physics_step_host_io", "icar_physics_step_host_io_async", "icar_host_register", "icar_host_unregister",
    "icar_default_mp_options", "icar_thompson_init", "icar_thompson_table", "icar_mp_gt_driver",
    "icar_profile_enable", "icar_profile_read", "icar_diagnostic_update", "icar_diagnostic_output", "icar_compute_dt", "icar_enforce_limits", "icar_step",
    "icar_upload_tendency", "icar_apply_forcing", "icar_balance_uvw",
    "icar_add_buffer_topo", "icar_linear_winds_init", "icar_linear_winds_dims", "icar_linear_winds_terrain_frequency",
    "icar_linear_perturbation", "icar_linear_perturb", "icar_default_lt_options", "icar_linear_winds_set_z",
    "icar_linear_perturb_spatial", "icar_restart_info", "icar_write_restart", "icar_read_restart")


def lib() -> C.CDLL:
    """Loads libicar_b200.so; raises loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IcarError(ERR_CUDA, f"{LIB_PATH} is missing: build it with "
                                      "`python -c 'import __graft_entry__ as g; g.build()'` "
                                      "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        l.icar_last_error.restype = C.c_char_p
        l.icar_last_error.argtypes = [C.c_void_p]
        l.icar_kernel_launches.restype = C.c_longlong
        l.icar_kernel_launches.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def restart_info(path):
    """Header of a restart file without a GPU: (Grid, microphysics, model_time, [var ids])."""
    g, mp, t, n = Grid(), C.c_int(), C.c_double(), C.c_int()
    ids = (C.c_int * len(VAR))()
    rc = lib().icar_restart_info(os.fsencode(path), C.byref(g), C.byref(mp), C.byref(t), C.byref(n), ids)
    if rc:
        raise IcarError(rc, lib().icar_last_error(None).decode())
    return g, mp.value, t.value, list(ids[:n.value])


def _f32(a, shape=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Domain:
    """domain_t for the hot path: one image's tile, fields resident on one B200."""

    def __init__(self, grid: Grid, options: Options | None = None, device: int = 0):
        self.grid = grid
        self.options = options or Options()
        self._h = C.c_void_p()
        rc = lib().icar_create(int(device), C.byref(self.grid), C.byref(self.options), C.byref(self._h))
        if rc:
            raise IcarError(rc, lib().icar_last_error(None).decode())

    # -- plumbing
    def _check(self, rc):
        if rc:
            raise IcarError(rc, lib().icar_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().icar_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, name):
        g = self.grid
        if name in ("u", "jacobian_u"):
            return (g.ny, g.nz, g.nx + 1), np.float32
        if name in ("v", "jacobian_v"):
            return (g.ny + 1, g.nz, g.nx), np.float32
        if name in ("precipitation", "snowfall", "graupel"):
            return (g.ny, g.nx), np.float64
        return (g.ny, g.nz, g.nx), np.float32

    # -- resident state
    def upload(self, name, array):
        name = ALIASES.get(name, name)
        shape, dt = self._shape(name)
        a = np.ascontiguousarray(array, dtype=dt)
        if a.shape != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        self._check(lib().icar_upload(self._h, VAR[name], _ptr(a)))

    def download(self, name):
        name = ALIASES.get(name, name)
        shape, dt = self._shape(name)
        a = np.empty(shape, dt)
        self._check(lib().icar_download(self._h, VAR[name], _ptr(a)))
        return a

    def upload_state(self, st):
        """Uploads every member of an icar_b200.ideal state dict the device side knows."""
        for k, v in st.items():
            n = ALIASES.get(k, k)
            if n in VAR and isinstance(v, np.ndarray):
                self.upload(n, v)

    def device_ptr(self, name):
        p = C.c_void_p()
        pitch = C.c_longlong()
        self._check(lib().icar_device_ptr(self._h, VAR[ALIASES.get(name, name)], C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    @property
    def model_time(self):
        s = C.c_double()
        self._check(lib().icar_get_model_time(self._h, C.byref(s)))
        return s.value

    @model_time.setter
    def model_time(self, seconds):
        self._check(lib().icar_set_model_time(self._h, C.c_double(seconds)))

    # -- plugin call surface
    def advect(self, dt):
        self._check(lib().icar_advect(self._h, C.c_float(dt)))

    def mp(self, dt, halo=None, subset=None):
        self._check(lib().icar_mp(self._h, C.c_float(dt), -1 if halo is None else int(halo),
                                  -1 if subset is None else int(subset)))

    def halo_send(self): self._check(lib().icar_halo_send(self._h))
    def halo_retrieve(self): self._check(lib().icar_halo_retrieve(self._h))
    def halo_exchange(self): self._check(lib().icar_halo_exchange(self._h))

    def physics_step(self, dt):
        self._check(lib().icar_physics_step(self._h, C.c_float(dt)))

    def physics_steps(self, dt, nsteps, graph=True):
        """nsteps passes of step()'s loop body at a fixed dt (time_step.f90:464-549); graph=True replays
        pairs of steps from a CUDA graph, with results identical to nsteps physics_step calls."""
        self._check(lib().icar_physics_steps(self._h, C.c_float(dt), int(nsteps), 1 if graph else 0))

    # -- the rest of step()'s loop body
    def diagnostic_update(self):
        """exner and density from pressure and potential temperature (time_step.f90:85-101)."""
        self._check(lib().icar_diagnostic_update(self._h))

    def diagnostic_output(self, dzdx=None, dzdy=None, want=("temperature", "pressure_interface", "temperature_interface",
                                                            "surface_pressure", "u_mass", "v_mass", "w_real")):
        """The output-only fields of diagnostic_update (time_step.f90:88-108, 165-194) as host arrays
        (ny, nz, nx) / (ny, nx); dzdx (ny, nz, nx+1) and dzdy (ny+1, nz, nx) are needed for w_real."""
        g = self.grid
        names = ("temperature", "pressure_interface", "temperature_interface", "surface_pressure", "u_mass", "v_mass", "w_real")
        out = {n: np.zeros((g.ny, g.nx) if n == "surface_pressure" else (g.ny, g.nz, g.nx), np.float32)
               for n in names if n in want}
        sx = _f32(dzdx, (g.ny, g.nz, g.nx + 1), "dzdx") if dzdx is not None else None
        sy = _f32(dzdy, (g.ny + 1, g.nz, g.nx), "dzdy") if dzdy is not None else None
        self._check(lib().icar_diagnostic_output(self._h, _ptr(sx) if sx is not None else None,
                                                 _ptr(sy) if sy is not None else None,
                                                 *[_ptr(out[n]) if n in out else None for n in names]))
        return out

    def compute_dt(self, dz_levels, cfl_reduction_factor=0.9, cfl_strictness=3):
        """update_dt (time_step.f90:375-423): CFL time step, min over images, capped at 120 s."""
        dz = _f32(dz_levels, (self.grid.nz,), "dz_levels")
        out = C.c_double()
        self._check(lib().icar_compute_dt(self._h, _ptr(dz), C.c_float(cfl_reduction_factor), int(cfl_strictness),
                                          C.byref(out)))
        return out.value

    def upload_tendency(self, name, dqdt):
        """var%dqdt_3d of a forced variable (per second), same shape as the variable."""
        name = ALIASES.get(name, name)
        shape, _ = self._shape(name)
        self._check(lib().icar_upload_tendency(self._h, VAR[name], _ptr(_f32(dqdt, shape, name))))

    def apply_forcing(self, dt_seconds):
        """domain%apply_forcing(dt), domain_obj.f90:2383-2448."""
        self._check(lib().icar_apply_forcing(self._h, C.c_double(dt_seconds)))

    def balance_uvw(self):
        """balance_uvw (wind.f90:81-172): w from the horizontal divergence of the resident u, v."""
        self._check(lib().icar_balance_uvw(self._h))

    # -- linear-theory winds (linear_winds.f90); 2-D arrays are (ny, nx) C-order = the Fortran (nx, ny) image
    def linear_winds_init(self, global_terrain, buffer=50):
        """setup_linwinds (linear_winds.f90:1180-1222): terrain_frequency on the device.  buffer: lt_parameters default 50."""
        g = self.grid
        t = _f32(global_terrain, (g.jde - g.jds + 1, g.ide - g.ids + 1), "global_terrain")
        self._check(lib().icar_linear_winds_init(self._h, _ptr(t), t.shape[1], t.shape[0], int(buffer)))

    def linear_winds_dims(self):
        """(fft_nx, fft_ny, total buffer) of the buffered grid."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().icar_linear_winds_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def terrain_frequency(self):
        """domain%terrain_frequency as a (fft_ny, fft_nx) complex128 array."""
        fnx, fny, _ = self.linear_winds_dims()
        out = np.empty((fny, fnx), np.complex128)
        self._check(lib().icar_linear_winds_terrain_frequency(self._h, _ptr(out)))
        return out

    def linear_perturbation(self, U, V, Nsq, z_bottom, z_top, minimum_step=100.0):
        """linear_perturbation (linear_winds.f90:239-344): scalars -> constz, (ny, nx) fields -> varyingz.
        Returns lt_data%u_perturb, v_perturb as (fft_ny, fft_nx) complex128 arrays."""
        fnx, fny, _ = self.linear_winds_dims()
        u = np.empty((fny, fnx), np.complex128)
        v = np.empty((fny, fnx), np.complex128)
        if np.ndim(z_bottom) == 0:
            rc = lib().icar_linear_perturbation(self._h, C.c_float(U), C.c_float(V), C.c_float(Nsq), C.c_float(z_bottom),
                                                C.c_float(z_top), None, None, C.c_float(minimum_step), _ptr(u), _ptr(v))
        else:
            g = self.grid
            shape = (g.jde - g.jds + 1, g.ide - g.ids + 1)
            zb, zt = _f32(z_bottom, shape, "z_bottom"), _f32(z_top, shape, "z_top")
            rc = lib().icar_linear_perturbation(self._h, C.c_float(U), C.c_float(V), C.c_float(Nsq), C.c_float(0), C.c_float(0),
                                                _ptr(zb), _ptr(zt), C.c_float(minimum_step), _ptr(u), _ptr(v))
        self._check(rc)
        return u, v

    def linear_perturb(self, U, V, Nsq, global_z_interface=None, global_dz_interface=None, layer_height=None,
                       dz_levels=None, minimum_layer_size=100.0, linear_contribution=1.0, linear_update_fraction=0.2):
        """linear_perturb (linear_winds.f90:1317-1366) evaluated directly for the background (U, V, Nsq): adds the
        perturbation of every level to the resident u, v.  Defaults: options_obj.f90:1459-1479."""
        g = self.grid
        zi = dzi = lh = dz = None
        if global_z_interface is not None:
            shape = (g.jde - g.jds + 1, g.nz, g.ide - g.ids + 1)
            zi, dzi = _f32(global_z_interface, shape, "global_z_interface"), _f32(global_dz_interface, shape, "global_dz_interface")
        else:
            lh, dz = _f32(layer_height, (g.nz,), "layer_height"), _f32(dz_levels, (g.nz,), "dz_levels")
        self._check(lib().icar_linear_perturb(self._h, C.c_float(U), C.c_float(V), C.c_float(Nsq), _ptr(zi), _ptr(dzi),
                                              _ptr(lh), _ptr(dz), C.c_float(minimum_layer_size),
                                              C.c_float(linear_contribution), C.c_float(linear_update_fraction)))

    def linear_winds_set_z(self, z):
        """domain%z%data_3d of the tile, (ny, nz, nx): calc_stability reads it when variable_N."""
        g = self.grid
        self._check(lib().icar_linear_winds_set_z(self._h, _ptr(_f32(z, (g.ny, g.nz, g.nx), "z"))))

    def linear_perturb_spatial(self, options=None, global_z_interface=None, global_dz_interface=None, layer_height=None,
                               dz_levels=None):
        """spatial_winds (linear_winds.f90:840-1127) with the reference's table semantics, the entries evaluated on
        demand.  Returns the number of (entry, level) evaluations."""
        g = self.grid
        zi = dzi = lh = dz = None
        if global_z_interface is not None:
            shape = (g.jde - g.jds + 1, g.nz, g.ide - g.ids + 1)
            zi, dzi = _f32(global_z_interface, shape, "global_z_interface"), _f32(global_dz_interface, shape, "global_dz_interface")
        else:
            lh, dz = _f32(layer_height, (g.nz,), "layer_height"), _f32(dz_levels, (g.nz,), "dz_levels")
        opt = options or LtOptions()
        n = C.c_longlong()
        self._check(lib().icar_linear_perturb_spatial(self._h, C.byref(opt), _ptr(zi), _ptr(dzi), _ptr(lh), _ptr(dz), C.byref(n)))
        return n.value

    # -- restart / output files (src/io/output_obj.f90 save_file, src/io/restart.f90:11-90)
    def write_restart(self, path):
        """Every resident variable + model time to one self-describing file (format: icar_b200/restart.py)."""
        self._check(lib().icar_write_restart(self._h, os.fsencode(path)))

    def read_restart(self, path):
        """restart_model: the variables and the model time of a file written for this tile and scheme."""
        self._check(lib().icar_read_restart(self._h, os.fsencode(path)))

    def enforce_limits(self):
        """domain%enforce_limits (domain_obj.f90:2228-2243)."""
        self._check(lib().icar_enforce_limits(self._h))

    def step(self, end_time, dz_levels, cfl_reduction_factor=0.9, cfl_strictness=3):
        """step(domain, end_time, options) (time_step.f90:438-549): internal time steps until model_time reaches
        end_time (seconds); returns how many were taken."""
        dz = _f32(dz_levels, (self.grid.nz,), "dz_levels")
        n = C.c_int()
        self._check(lib().icar_step(self._h, C.c_double(end_time), _ptr(dz), C.c_float(cfl_reduction_factor),
                                    int(cfl_strictness), C.byref(n)))
        return n.value

    def synchronize(self):
        self._check(lib().icar_synchronize(self._h))

    def mp_diagnostics(self):
        out = (C.c_longlong * 3)()
        self._check(lib().icar_mp_diagnostics(self._h, out))
        return tuple(int(x) for x in out)

    @property
    def kernel_launches(self):
        return int(lib().icar_kernel_launches(self._h))

    @property
    def fused_kernel_name(self) -> str:
        buf = C.create_string_buffer(64)
        self._check(lib().icar_fused_kernel_name(self._h, buf))
        return buf.value.decode()

    def selftest_division(self, a, b):
        """(div2x(a, b), a / b) evaluated on the device for two float32 arrays of the same even size."""
        a = np.ascontiguousarray(a, np.float32).ravel()
        b = np.ascontiguousarray(b, np.float32).ravel()
        packed, ieee = np.empty_like(a), np.empty_like(a)
        fp = lambda x: x.ctypes.data_as(C.c_void_p)
        self._check(lib().icar_selftest_division(self._h, fp(a), fp(b), fp(packed), fp(ieee), C.c_longlong(a.size)))
        return packed, ieee

    def timer_start(self):
        self._check(lib().icar_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(lib().icar_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def profile_enable(self, on=True):
        self._check(lib().icar_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        """{kernel family: (milliseconds, launches)} since the last read (CUDA events per launch)."""
        names = ((C.c_char * 32) * 16)()
        ms = (C.c_float * 16)()
        n_l = (C.c_longlong * 16)()
        n = C.c_int()
        self._check(lib().icar_profile_read(self._h, 16, names, ms, n_l, C.byref(n)))
        return {names[i].value.decode(): (float(ms[i]), int(n_l[i])) for i in range(n.value)}

    # -- multi-image
    def comm_init(self, image, nimages, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(lib().icar_comm_init(self._h, int(image), int(nimages), buf))

    # -- host-pointer leaf calls
    def mpdata_set_winds(self, u, v, w, jac_u, jac_v, jac_w, rho, dt, dx, scheme=2):
        dt_ = np.float64 if np.asarray(w).dtype == np.float64 else np.float32
        arrs = [np.ascontiguousarray(a, dtype=dt_) if a is not None else None
                for a in (u, v, w, jac_u, jac_v, jac_w, rho)]
        if dt_ == np.float64:
            self._check(lib().icar_mpdata_set_winds_f64(self._h, *[_ptr(a) for a in arrs],
                                                        C.c_double(dt), C.c_double(dx), int(scheme)))
        else:
            self._check(lib().icar_mpdata_set_winds(self._h, *[_ptr(a) for a in arrs],
                                                    C.c_float(dt), C.c_float(dx), int(scheme)))

    def mpdata_advect3d(self, q, rho, jaco, dz, order=2, fct=True):
        """advect3d on host arrays; returns the advected copy of q."""
        g = self.grid
        dt_ = np.float64 if np.asarray(q).dtype == np.float64 else np.float32
        q = np.array(q, dtype=dt_, order="C", copy=True)
        rho = None if rho is None else np.ascontiguousarray(rho, dtype=dt_)
        jaco = np.ascontiguousarray(jaco, dtype=dt_)
        dz = np.ascontiguousarray(dz, dtype=dt_)
        fn = lib().icar_mpdata_advect3d_f64 if dt_ == np.float64 else lib().icar_mpdata_advect3d
        self._check(fn(self._h, _ptr(q), _ptr(rho), g.ims, g.ime, g.kms, g.kme, g.jms, g.jme,
                       _ptr(jaco), _ptr(dz), int(order), int(bool(fct))))
        return q

    def mp_simple_driver(self, pressure, th, pii, rho, qv, qc, qr, qs, rain, snow, dt, dz,
                         its, ite, jts, jte, kts, kte):
        """mp_simple_driver on host arrays (updated in place); bounds are Fortran 1-based."""
        g = self.grid
        for a in (th, qv, qc, qr, qs, rain, snow):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        self._check(lib().icar_mp_simple_driver(
            self._h, _ptr(_f32(pressure)), _ptr(th), _ptr(_f32(pii)), _ptr(_f32(rho)), _ptr(qv),
            _ptr(qc), _ptr(qr), _ptr(qs), _ptr(rain), _ptr(snow), C.c_float(dt), _ptr(_f32(dz)),
            g.ims, g.ime, g.jms, g.jme, g.kms, g.kme, int(its), int(ite), int(jts), int(jte),
            int(kts), int(kte)))

    # -- Thompson microphysics
    def thompson_init(self, mp_options: "MpOptions | None" = None):
        """thompson_init(mp_options), mp_thompson.f90:342 -- builds the lookup tables on the device."""
        self._check(lib().icar_thompson_init(self._h, C.byref(mp_options) if mp_options is not None else None))

    def thompson_table(self, name):
        """One lookup table as float64, C-order view of the Fortran array (reversed shape)."""
        n = C.c_longlong()
        tid = THOMPSON_TABLES.index(name)
        self._check(lib().icar_thompson_table(self._h, tid, None, C.byref(n)))
        a = np.empty(n.value, np.float64)
        self._check(lib().icar_thompson_table(self._h, tid, _ptr(a), C.byref(n)))
        return a.reshape(THOMPSON_SHAPES[name][::-1])

    def mp_gt_driver(self, st, dt, its, ite, jts, jte, kts, kte, ide=None, jde=None):
        """mp_gt_driver on host arrays of a state dict (updated in place); bounds are Fortran
        1-based.  st: th, qv, qc, qr, qi, qs, qg, ni, nr, pii, pressure, dz_mass (3-D float32) and
        rainnc, sr [, rainncv, snownc, snowncv, graupelnc, graupelncv] (2-D float32)."""
        g = self.grid
        for k in ("qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "th", "rainnc", "sr"):
            assert st[k].dtype == np.float32 and st[k].flags["C_CONTIGUOUS"], k
        o = lambda k: _ptr(st[k]) if st.get(k) is not None else None
        self._check(lib().icar_mp_gt_driver(
            self._h, _ptr(st["qv"]), _ptr(st["qc"]), _ptr(st["qr"]), _ptr(st["qi"]), _ptr(st["qs"]),
            _ptr(st["qg"]), _ptr(st["ni"]), _ptr(st["nr"]), _ptr(st["th"]), _ptr(_f32(st["pii"])),
            _ptr(_f32(st["pressure"])), _ptr(_f32(st["dz_mass"])), C.c_float(dt), 1,
            _ptr(st["rainnc"]), o("rainncv"), o("snownc"), o("snowncv"), o("graupelnc"), o("graupelncv"),
            _ptr(st["sr"]), g.ids, g.ide if ide is None else ide, g.jds, g.jde if jde is None else jde,
            g.kds, g.kde, g.ims, g.ime, g.jms, g.jme, g.kms, g.kme,
            int(its), int(ite), int(jts), int(jte), int(kts), int(kte)))

    # -- host-domain step (end-to-end path)
    def host_domain(self, st):
        """icar_host_domain over a state dict of contiguous NumPy arrays (kept alive by caller)."""
        hd = HostDomain()
        m = dict(u="u", v="v", w="w", pressure="pressure", exner="pii", density="rho",
                 potential_temperature="th", water_vapor="qv", cloud_water="qc", rain_in_air="qr",
                 snow_in_air="qs", accumulated_precipitation="precipitation",
                 accumulated_snowfall="snowfall", cloud_ice="qi", ice_number="ni", rain_number="nr",
                 graupel_in_air="qg", accumulated_graupel="graupel")
        for field, key in m.items():
            a = st.get(key)
            if a is not None:
                assert a.flags["C_CONTIGUOUS"]
                setattr(hd, field, a.ctypes.data)
        return hd

    def physics_step_host(self, dt, hd):
        self._check(lib().icar_physics_step_host(self._h, C.c_float(dt), C.byref(hd)))

    def step_host(self, dt, hd):
        """diagnostic_update + physics step on host arrays; hd.exner and hd.density must be NULL."""
        self._check(lib().icar_step_host(self._h, C.c_float(dt), C.byref(hd)))

    def physics_step_host_io(self, dt, hd_in, hd_out, row0, row1, derive_diagnostics=False):
        """Host step with separate input / output arrays; only the tile's rows row0..row1 come back."""
        self._check(lib().icar_physics_step_host_io(self._h, C.c_float(dt), C.byref(hd_in), C.byref(hd_out),
                                                    int(row0), int(row1), 1 if derive_diagnostics else 0))

    def physics_step_host_io_async(self, dt, hd_in, hd_out, row0, row1, derive_diagnostics=False):
        """physics_step_host_io queued only (pinned host arrays); wait with synchronize()."""
        self._check(lib().icar_physics_step_host_io_async(self._h, C.c_float(dt), C.byref(hd_in), C.byref(hd_out),
                                                          int(row0), int(row1), 1 if derive_diagnostics else 0))

    def host_register(self, hd): self._check(lib().icar_host_register(self._h, C.byref(hd)))
    def host_unregister(self, hd): self._check(lib().icar_host_unregister(self._h, C.byref(hd)))


def halo_plan(grid: Grid, image: int, nimages: int):
    """icar_halo_plan: first local indices (0-based, along the exchange axis) sent to / received
    from the lower (west/south) and upper (east/north) neighbour."""
    plan = (C.c_int * 8)()
    rc = lib().icar_halo_plan(C.byref(grid), int(image), int(nimages), plan)
    if rc:
        raise IcarError(rc, "icar_halo_plan: invalid arguments")
    return dict(send_lo=plan[0], send_hi=plan[1], recv_lo=plan[2], recv_hi=plan[3],
                has_lo=bool(plan[4]), has_hi=bool(plan[5]), axis=("x", "y")[plan[6]], width=plan[7])


def comm_init_local(domains) -> None:
    """icar_comm_init_local: connects Domains that live in this process (image order) so that they exchange halos."""
    arr = (C.c_void_p * len(domains))(*[d._h.value for d in domains])
    rc = lib().icar_comm_init_local(arr, len(domains))
    if rc:
        raise IcarError(rc, "; ".join(lib().icar_last_error(d._h).decode() for d in domains))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().icar_comm_unique_id(buf)
    if rc:
        raise IcarError(rc, lib().icar_last_error(None).decode())
    return buf.raw


# -- the reference's call surface as free functions -------------------------------------------
def advect(domain: Domain, options: Options | None, dt: float) -> None:
    """advect(domain, options, dt) -- advection_driver.f90:51 (options live in the domain's context)."""
    domain.advect(dt)


def mp(domain: Domain, options: Options | None, dt: float, halo=None, subset=None) -> None:
    """mp(domain, options, dt_in, halo, subset) -- mp_driver.f90:673."""
    domain.mp(dt, halo=halo, subset=subset)
