"""TEST INFRASTRUCTURE -- the reference's own Fortran, executed.

    from oracle import f90run
    ref = f90run.Reference("/root/reference")          # translates src/physics/adv_mpdata.f90 and mp_simple.f90 in place
    ref.advect3d(q, U_m, V_m, W_m, jaco, dz, mpdata_order=2, fct=True)          # adv_mpdata.f90:356 on (ny, nz, nx) arrays
    ref.mp_simple_driver(p, th, pii, rho, qv, qc, qr, qs, rain, snow, dt, dz)  # mp_simple.f90:595

Nothing under /root/reference is copied: the source files are read where they lie, translated statement by statement
(translate.py) and run on runtime.py (IEEE binary32, glibc expf).  Only tests/ and scripts that make golden vectors
import this package; it exists so that the hand-written restatements in oracle/ are pinned to outputs of the
reference's own source text (tests/test_reference_source.py, tests/golden/reference_source_*.npz).
"""
import os
import types

import numpy as np

from .runtime import FArr, f32, f_exp, unallocated
from .translate import Module, TranslateError  # noqa: F401


def _f(a):
    """(ny, nz, nx) C-ordered float32 -> the Fortran array (nx, nz, ny) over the SAME memory"""
    a = np.ascontiguousarray(a, np.float32)
    return a, a.T


class Reference:
    def __init__(self, root="/root/reference"):
        phys = os.path.join(root, "src", "physics")
        self.adv_path = os.path.join(phys, "adv_mpdata.f90")
        self.mp_path = os.path.join(phys, "mp_simple.f90")
        for p in (self.adv_path, self.mp_path, os.path.join(phys, "adv_mpdata_FCT_core.f90")):
            if not os.path.exists(p):
                raise FileNotFoundError(p)
        # kVARS (src/constants/icar_constants.f90): only the identity of an index matters to mpdata's scalar list
        self.kvars = types.SimpleNamespace(**{n: k + 1 for k, n in enumerate(self.ADVECTED)})
        self.adv = Module(self.adv_path, names={"kvars": self.kvars}).build(["advect3d", "upwind_advection", "mpdata_fluxes", "flux_limiter", "flux1", "mpdata"])
        self.mp = Module(self.mp_path).build(["mp_simple_driver"])
        self.upw = Module(os.path.join(phys, "advect.f90"), names={"kvars": self.kvars}).build(["upwind"])
        self.wind_path = os.path.join(phys, "wind.f90")
        self.wind = Module(self.wind_path).build(["balance_uvw", "calc_divergence"])
        self._time_step_path = os.path.join(root, "src", "main", "time_step.f90")
        self.time_step = Module(self._time_step_path).build(["compute_dt"])
        # icar_constants (Rd, cp, ...) for exner_function: the values come from the source text as well
        self.constants = Module(os.path.join(root, "src", "constants", "icar_constants.f90")).build([]).M
        self.atm = Module(os.path.join(root, "src", "utilities", "atm_utilities.f90"),
                          names={"rd": self.constants.rd, "cp": self.constants.cp}).build(["exner_function"])

    # ---- mp_thompson.f90
    # thompson_init runs as written except for the two rain-collection tables, whose builders (qr_acr_qg, qr_acr_qs;
    # 1.4e10 inner iterations) take the entries from `big_tables` -- the C++ oracle's tables, which have their own
    # independent NumPy / SciPy check (tests/test_oracle_thompson.py).  Every scalar constant, the bin grids, the four
    # other table builders (freezeH2O, qi_aut_qs, table_Efrw, table_Efsw), mp_gt_driver and mp_thompson are the
    # reference's text, executed.
    BIG_QG = ("tcg_racg", "tmr_racg", "tcr_gacr", "tmg_gacr", "tnr_racg", "tnr_gacr")
    BIG_QS = ("tcs_racs1", "tmr_racs1", "tcs_racs2", "tmr_racs2", "tcr_sacr1", "tms_sacr1", "tcr_sacr2", "tms_sacr2",
              "tnr_racs1", "tnr_racs2", "tnr_sacr1", "tnr_sacr2")
    MP_OPTIONS = dict(nt_c=100.e6, tno=5.0, am_s=0.069, rho_g=500.0, av_s=40.0, bv_s=0.55, fv_s=100.0, av_i=1847.5,
                      av_g=442.0, bv_g=0.89, ef_si=0.05, ef_rs=0.95, ef_rg=0.75, ef_ri=0.95, c_cubes=0.5, c_sqrd=0.3,
                      mu_r=0.0, t_adjust=0.0, ef_rw_l=False, ef_sw_l=False)      # options_obj.f90:1259-1280

    FREEZE = ("tpi_qcfz", "tni_qcfz", "tpi_qrfz", "tpg_qrfz", "tni_qrfz", "tnr_qrfz")

    def thompson_build(self, tables, skip_freezeh2o=False):
        """translate mp_thompson.f90; tables: name -> array in the Fortran's shape for the builders that are not run
        (qr_acr_qg, qr_acr_qs always; freezeH2O -- 6e6 inner iterations, two minutes here -- on request)"""
        def fill(names):
            def f(M):
                for n in names:
                    getattr(M, n).a[...] = tables[n]
            return lambda M, *a, **k: f(M)
        path = os.path.join(os.path.dirname(self.mp_path), "mp_thompson.f90")
        over = {"qr_acr_qg": fill(self.BIG_QG), "qr_acr_qs": fill(self.BIG_QS)}
        if skip_freezeh2o:
            over["freezeh2o"] = fill(self.FREEZE)
        mod = Module(path, externals={"wrf_debug": None, "wrf_error_fatal": None}, overrides=over)
        self.th = mod.build(["thompson_init", "mp_gt_driver", "mp_thompson", "rslf", "rsif", "wgamma", "gammp", "gammln"])
        return self.th

    def thompson_init(self, tables, skip_freezeh2o=False, **options):
        """thompson_init(mp_options), mp_thompson.f90:342 (about three minutes; half a minute without freezeH2O)"""
        self.thompson_build(tables, skip_freezeh2o)
        o = dict(self.MP_OPTIONS)
        o.update(options)
        self.th.thompson_init(types.SimpleNamespace(**{k: (v if isinstance(v, bool) else f32(v)) for k, v in o.items()}))
        return self.th

    def thompson_big_slab(self, which, km):
        """One (k, m) slab, km = (m-1)*37 + (k-1), of the rain-graupel (which = 0: qr_acr_qg) or rain-snow (1: qr_acr_qs)
        collection tables, computed by the reference's own builder: the routine runs as written with the km range
        [km, km] -- the range wrf_dm_decomp1d hands one rank of a distributed run (mp_thompson.f90:2903-2907, 3075-3079;
        the serial build assigns 0 .. ntb_r*ntb_r1-1 there).  Needs thompson_init first.  Returns name -> (28, 28 | 9) slab."""
        name = ("qr_acr_qg", "qr_acr_qs")[which]
        path = os.path.join(os.path.dirname(self.mp_path), "mp_thompson.f90")
        mod = Module(path, externals={"wrf_debug": None, "wrf_error_fatal": None},
                     patches={name: [(r"^km_s\s*=", f"km_s = {int(km)}"), (r"^km_e\s*=", f"km_e = {int(km)}")]})
        ns = mod.build([name])
        # the module state the builder reads: everything thompson_init has set
        for k, v in vars(self.th.M).items():
            if isinstance(v, FArr) and k in (self.BIG_QG + self.BIG_QS):
                setattr(ns.M, k, FArr(np.zeros_like(v.a), v.lb, v.dtype, True))
            else:
                setattr(ns.M, k, v)
        getattr(ns, name)()
        k, m = km % 37 + 1, km // 37 + 1
        return {n: getattr(ns.M, n).a[:, :, k - 1, m - 1].copy() for n in (self.BIG_QG if which == 0 else self.BIG_QS)}

    # the order of oracle.Thompson.constants()
    def thompson_constants(self):
        M = self.th.M
        out = [M.mu_c, M.sc3, M.d0i, M.xm0s, M.xm0g, M.am_r, M.am_i, M.am_g]
        for a in ("cce", "ccg", "cie", "cig", "cre", "crg", "cse", "csg", "cge", "cgg"):
            out += list(getattr(M, a).a)
        out += [getattr(M, n) for n in ("t1_qr_qc", "t1_qr_qi", "t2_qr_qi", "t1_qg_qc", "t1_qs_qc", "t1_qs_qi", "t1_qr_ev",
                                        "t2_qr_ev", "t1_qs_sd", "t2_qs_sd", "t1_qg_sd", "t2_qg_sd", "t1_qs_me", "t2_qs_me",
                                        "t1_qg_me", "t2_qg_me")]
        return np.array(out, np.float32)

    def mp_gt_driver(self, st, dt, tile=None):
        """mp_gt_driver (mp_thompson.f90:772) on a state dict of (ny, nz, nx) float32 arrays (th, qv, qc, qr, qi, qs, qg,
        ni, nr, pii, pressure, dz_mass) and (ny, nx) rainnc, snownc, graupelnc, sr, updated in place; tile 1-based.
        The domain is the memory tile (ide = nx, jde = ny), so the routine's own clip to ide-1 / jde-1 (:821-822) applies."""
        ny, nz, nx = st["th"].shape
        its, ite, jts, jte, kts, kte = tile or (1, nx, 1, ny, 1, nz)
        T = lambda k: st[k].T                                   # noqa: E731
        for k in ("qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "th", "pii", "pressure", "dz_mass", "rainnc", "snownc", "graupelnc", "sr"):
            if not (st[k].flags.c_contiguous and st[k].dtype == np.float32):
                raise ValueError(k + " must be C-contiguous float32")
        rainncv = np.zeros((ny, nx), np.float32)
        snowncv = np.zeros((ny, nx), np.float32)
        graupelncv = np.zeros((ny, nx), np.float32)
        self.th.mp_gt_driver(T("qv"), T("qc"), T("qr"), T("qi"), T("qs"), T("qg"), T("ni"), T("nr"), T("th"), T("pii"),
                             T("pressure"), T("dz_mass"), f32(dt), 1, T("rainnc"), rainncv.T, T("snownc"), snowncv.T,
                             T("graupelnc"), graupelncv.T, T("sr"),
                             1, nx, 1, ny, 1, nz, 1, nx, 1, ny, 1, nz, its, ite, jts, jte, kts, kte)

    exp4 = staticmethod(f_exp)          # exp of a REAL as the translated code evaluates it (glibc expf)

    # the order of mpdata's advect3d calls (adv_mpdata.f90:512-522): kVARS name -> domain component
    ADVECTED = ("water_vapor", "cloud_water", "rain_in_air", "snow_in_air", "potential_temperature", "cloud_ice",
                "graupel_in_air", "ice_number_concentration", "rain_number_concentration", "snow_number_concentration",
                "graupel_number_concentration")
    COMPONENT = {"water_vapor": "water_vapor", "cloud_water": "cloud_water_mass", "rain_in_air": "rain_mass",
                 "snow_in_air": "snow_mass", "potential_temperature": "potential_temperature",
                 "cloud_ice": "cloud_ice_mass", "graupel_in_air": "graupel_mass",
                 "ice_number_concentration": "cloud_ice_number", "rain_number_concentration": "rain_number",
                 "snow_number_concentration": "snow_number", "graupel_number_concentration": "graupel_number"}

    def _mock_domain(self, fields, u, v, w, jacobian_u, jacobian_v, jacobian_w, jacobian, advection_dz, dx, mpdata_order, fct):
        ns = types.SimpleNamespace
        ny, nz, nx = w.shape
        fa = lambda a: FArr(_f(a)[1], (1, 1, 1))          # noqa: E731
        dom = ns(grid=ns(ims=1, ime=nx, kms=1, kme=nz, jms=1, jme=ny), ims=1, ime=nx, kms=1, kme=nz, jms=1, jme=ny,
                 dx=f32(dx), advection_dz=fa(advection_dz), density=ns(data_3d=fa(np.ones_like(w))),
                 u=ns(data_3d=fa(u)), v=ns(data_3d=fa(v)), w=ns(data_3d=fa(w)), jacobian_u=fa(jacobian_u),
                 jacobian_v=fa(jacobian_v), jacobian_w=fa(jacobian_w), jacobian=fa(jacobian))
        flags = {}
        for name in self.ADVECTED:
            arr = fields.get(name)
            flags[getattr(self.kvars, name)] = 1 if arr is not None else 0
            if arr is not None:
                if not (arr.flags.c_contiguous and arr.dtype == np.float32):
                    raise ValueError("fields must be C-contiguous float32 (advected in place)")
                setattr(dom, self.COMPONENT[name], ns(data_3d=FArr(arr.T, (1, 1, 1))))
        opts = ns(parameters=ns(advect_density=False, debug=False, dz_levels=None), vars_to_advect=flags,
                  adv_options=ns(mpdata_order=int(mpdata_order), flux_corrected_transport=bool(fct)))
        return dom, opts

    def upwind(self, fields, u, v, w, jacobian_u, jacobian_v, jacobian_w, jacobian, advection_dz, dt, dx):
        """upwind(domain, options, dt), advect.f90:361-388 (advection = 1): setup_module_winds :304-336 + advect3d :106-178"""
        dom, opts = self._mock_domain(fields, u, v, w, jacobian_u, jacobian_v, jacobian_w, jacobian, advection_dz, dx, 1, False)
        M = self.upw.M
        for n in ("u_m", "v_m", "w_m", "lastqv_m"):
            getattr(M, n).a = None
        self.upw.upwind(dom, opts, f32(dt))

    def balance_uvw(self, u, v, w, jacobian_u, jacobian_v, jacobian_w, dz, dx, jacobian):
        """balance_uvw (wind.f90:81-172): w (ny, nz, nx) is overwritten in place"""
        for a in (u, v, w):
            if not (a.flags.c_contiguous and a.dtype == np.float32):
                raise ValueError("u, v, w must be C-contiguous float32")
        self.wind.balance_uvw(u.T, v.T, w.T, _f(jacobian_u)[1], _f(jacobian_v)[1], _f(jacobian_w)[1], _f(dz)[1], f32(dx),
                              _f(jacobian)[1], None)

    def compute_dt(self, dx, u, v, w, dz_levels, cfl, cfl_strictness):
        """compute_dt (time_step.f90:217-330) with use_density = .false.; raises runtime.FortranStop where it stops"""
        return self.time_step.compute_dt(f32(dx), _f(u)[1], _f(v)[1], _f(w)[1], _f(np.ones_like(w))[1],
                                         np.ascontiguousarray(dz_levels, np.float32), f32(cfl), int(cfl_strictness), False)

    def update_dt(self, dx, u, v, w, dz_levels, cfl, cfl_strictness, other_images=()):
        """update_dt(dt, options, domain, end_time), time_step.f90:375-423, as written (a build without __INTEL_COMPILER):
        compute_dt on the domain's winds with use_density = .false., `co_min` over this image's value and those of
        `other_images`, and dt%set(seconds = min(seconds, 120)).  Returns the seconds dt was set to."""
        ns = types.SimpleNamespace
        if not hasattr(self, "_update_dt"):
            self._co_min_others = []
            self._update_dt = Module(self._time_step_path,
                                     externals={"co_min": (lambda x: min([x] + list(self._co_min_others)), [0])}
                                     ).build(["update_dt"])
        self._co_min_others = [np.float64(x) for x in other_images]
        fa = lambda a: FArr(_f(a)[1], (1,) * a.ndim)          # noqa: E731
        dom = ns(dx=f32(dx), u=ns(data_3d=fa(u)), v=ns(data_3d=fa(v)), w=ns(data_3d=fa(w)),
                 density=ns(data_3d=fa(np.ones_like(w))))
        opts = ns(parameters=ns(dz_levels=np.ascontiguousarray(dz_levels, np.float32), cfl_reduction_factor=f32(cfl),
                                cfl_strictness=int(cfl_strictness)))
        got = []
        dt = ns(set=lambda seconds: got.append(seconds))
        self._update_dt.update_dt(dt, opts, dom, None)
        assert len(got) == 1
        return got[0]

    def diagnostic_update(self, pressure, th, u, v, w, jacobian, dzdx, dzdy):
        """diagnostic_update(domain, options), time_step.f90:49-198, on a mock domain that has the fields of this path
        (exner, density, temperature, the interface fields, psfc, u_mass, v_mass, w_real) and none of the others
        (ivt / iwv / roughness / 10 m winds are not associated).  Returns name -> (ny, nz, nx) / (ny, nx) arrays."""
        ns = types.SimpleNamespace
        ny, nz, nx = w.shape
        if not hasattr(self, "_diag"):
            self._diag = Module(self._time_step_path, names={"karman": self.constants.karman, "rd": self.constants.rd},
                                externals={"exner_function": self.atm.exner_function, "compute_ivt": None,
                                           "compute_iq": None}).build(["diagnostic_update"])
        out = {k: np.zeros((ny, nz, nx), np.float32) for k in ("exner", "density", "temperature", "pressure_interface",
                                                                "temperature_interface", "u_mass", "v_mass", "w_real")}
        out["surface_pressure"] = np.zeros((ny, nx), np.float32)
        fa = lambda a: FArr(_f(a)[1], (1,) * a.ndim)          # noqa: E731
        var3 = lambda a: ns(data_3d=fa(a))                    # noqa: E731
        none3, none2 = ns(data_3d=None), ns(data_2d=None)
        dom = ns(ims=1, ime=nx, kms=1, kme=nz, jms=1, jme=ny, exner=var3(out["exner"]), pressure=var3(pressure),
                 pressure_interface=var3(out["pressure_interface"]), dz_interface=none3,
                 surface_pressure=ns(data_2d=FArr(out["surface_pressure"].T, (1, 1))), density=var3(out["density"]),
                 water_vapor=none3, cloud_water_mass=none3, rain_mass=none3, cloud_ice_mass=none3, snow_mass=none3,
                 graupel_mass=none3, temperature=var3(out["temperature"]),
                 temperature_interface=var3(out["temperature_interface"]), u=var3(u), v=var3(v), w=var3(w),
                 w_real=var3(out["w_real"]), u_mass=var3(out["u_mass"]), v_mass=var3(out["v_mass"]),
                 potential_temperature=var3(th), ivt=none2, iwv=none2, iwl=none2, iwi=none2, roughness_z0=none2,
                 u_10m=none2, v_10m=none2, ustar=None, z=none3, terrain=none2, dzdx=fa(dzdx), dzdy=fa(dzdy),
                 jacobian=fa(jacobian))
        M = self._diag.M
        for n in ("lastw", "currw", "uw", "vw"):            # `if (.not.allocated(lastw))`: a new grid needs new module arrays
            getattr(M, n).a = None
        self._diag.diagnostic_update(dom, ns())
        return out

    def advect_dispatch(self, advection):
        """advect(domain, options, dt), src/physics/advection_driver.f90:51-66, with upwind / mpdata replaced by recorders
        and kADV_* read from icar_constants.f90: returns the list of schemes it called for options%physics%advection."""
        ns = types.SimpleNamespace
        called = []
        if not hasattr(self, "_adv_driver"):
            root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
            self._adv_driver = Module(os.path.join(root, "src", "physics", "advection_driver.f90"),
                                      names={"kadv_upwind": self.constants.kadv_upwind, "kadv_mpdata": self.constants.kadv_mpdata},
                                      externals={"upwind": lambda d, o, dt: d.append(("upwind", float(dt))),
                                                 "mpdata": lambda d, o, dt: d.append(("mpdata", float(dt)))}).build(["advect"])
        self._adv_driver.advect(called, ns(physics=ns(advection=int(advection))), f32(120.0))
        return called

    def update_winds_calls(self, windtype, first):
        """update_winds(domain, options), src/physics/wind.f90:289-369, with every routine it calls replaced by a recorder.
        first: the call at start-up (domain%sintheta not allocated yet: works on u / v / w themselves) or a later one (works on
        the next forcing step's winds held in %meta_data%dqdt_3d, update = .True.).  windtype: the name of the constant in
        icar_constants.f90 ("kwind_linear", ...) or an integer.  Returns [(routine, which arrays, update flag or None)]."""
        ns = types.SimpleNamespace
        log = []
        tags = ["?"]

        def arrays(name):                                   # a one-cell array whose value says which array it is
            tags.append(name)
            return FArr(np.full((1, 1, 1), len(tags) - 1, np.float32), (1, 1, 1))

        def which(a):
            a = getattr(a, "a", a)
            return tags[int(np.asarray(a).reshape(-1)[0])]
        u, v, w = (ns(data_3d=arrays(n), meta_data=ns(dqdt_3d=arrays(n + " of the next forcing step"))) for n in "uvw")
        dom = ns(sintheta=unallocated(np.float32, 2) if first else arrays("sintheta"), costheta=arrays("costheta"), u=u, v=v, w=w,
                 jacobian_u=arrays("jacobian_u"), jacobian_v=arrays("jacobian_v"), jacobian_w=arrays("jacobian_w"),
                 advection_dz=arrays("advection_dz"), dx=f32(4000.0), jacobian=arrays("jacobian"),
                 zfr_u=arrays("zfr_u"), zfr_v=arrays("zfr_v"), zr_u=arrays("zr_u"), zr_v=arrays("zr_v"))
        c = self.constants
        wt = int(getattr(c, windtype)) if isinstance(windtype, str) else int(windtype)
        opts = ns(physics=ns(windtype=wt), lt_options=ns(vert_smooth=2),
                  parameters=ns(advect_density=False, use_terrain_difference=False))
        if not hasattr(self, "_update_winds"):
            self._update_winds = Module(
                self.wind_path, names={n: int(getattr(c, n)) for n in ("kwind_linear", "kconserve_mass", "kiterative_winds",
                                                                         "klinear_iterative_winds")},
                overrides={
                    "init_winds": lambda M, d, o: self._uw_log.append(("init_winds", None, None)),
                    "make_winds_grid_relative": lambda M, uu, vv, ww, st, ct: self._uw_log.append(
                        ("make_winds_grid_relative", self._uw_which(uu), None)),
                    "mass_conservative_acceleration": lambda M, uu, vv, zu, zv: self._uw_log.append(
                        ("mass_conservative_acceleration", self._uw_which(uu), None)),
                    "iterative_winds": lambda M, d, o, update_in=False: self._uw_log.append(("iterative_winds", None, bool(update_in))),
                    "balance_uvw": lambda M, uu, vv, ww, ju, jv, jw, dz, dx, jaco, o: self._uw_log.append(
                        ("balance_uvw", self._uw_which(uu), None))},
                externals={"linear_perturb": lambda d, o, vsmooth, reverse, use_density, update=False: self._uw_log.append(
                    ("linear_perturb", "reverse" if reverse else None, bool(update)))}).build(["update_winds"])
        self._uw_log, self._uw_which = log, which
        self._update_winds.update_winds(dom, opts)
        return log

    def step_loop(self, start_seconds, end_seconds, dts):
        """step(domain, end_time, options), time_step.f90:438-549, executed as written on a mock domain whose physics are
        recorders: update_dt hands out the scripted `dts` (seconds) in turn; rad / lsm / pbl / convect / mp / advect,
        domain%halo_send / halo_retrieve / apply_forcing / enforce_limits and diagnostic_update append to the log.
        Model time is the reference's: a Time_type holds days in REAL*8 (time_obj.f90: `addition` :892-902 adds dt%days(),
        `difference` :861-884 multiplies the day difference by 86400, the comparisons :656-831 compare days) and a
        time_delta_t holds days too (time_delta_obj.f90: set :28-44 divides seconds by 86400, seconds() multiplies).
        The debug / interactive branches (domain_check, print_progress) are patched to .false..
        Returns (log, final model time in seconds); log entries are (name, *arguments)."""
        ns = types.SimpleNamespace
        log = []

        class Delta:                                     # time_delta_t
            def __init__(self, seconds=0.0): self.delta = np.float64(seconds) / np.float64(86400.0)
            def set(self, seconds): self.delta = np.float64(seconds) / np.float64(86400.0)
            def days(self): return self.delta
            def seconds(self): return self.delta * np.float64(86400.0)
            def __mul__(self, k): d = Delta(); d.delta = self.delta * np.float64(k); return d

        class Time:                                      # Time_type, same calendar and reference date on both sides
            def __init__(self, days): self.current_date_time = np.float64(days)
            def mjd(self): return np.longdouble(self.current_date_time)
            def seconds(self): return np.longdouble(self.current_date_time) * np.float64(86400.0)
            def __add__(self, dt): return Time(np.float64(self.mjd() + dt.days()))
            def __sub__(self, other):
                d = Delta(); d.set(np.float64(self.mjd() - other.mjd()) * np.float64(86400.0)); return d
            def __lt__(self, o): return not (self.current_date_time >= o.current_date_time)
            def __gt__(self, o): return self.current_date_time > o.current_date_time

        script = list(dts)

        def update_dt(M, dt, options, domain, end_time):
            log.append(("update_dt",))
            # dt%set(seconds=...), time_step.f90:421; `dt` is the caller's local (a plain record until step() assigns a
            # difference of two times to it), so it is given time_delta_t's state and accessors here
            dt.delta = np.float64(script.pop(0)) / np.float64(86400.0)
            dt.days = lambda d=dt: d.delta
            dt.seconds = lambda d=dt: d.delta * np.float64(86400.0)

        def recorder(name):
            def f(domain, options, dt, halo=None, subset=None, **k):
                log.append((name, float(dt)) + ((("halo", int(halo)),) if halo is not None else ())
                           + ((("subset", int(subset)),) if subset is not None else ()))
            return f

        dom = ns(grid=ns(ims=1, jms=1, kms=1), model_time=Time(np.float64(start_seconds) / np.float64(86400.0)))
        dom.halo_send = lambda: log.append(("halo_send",))
        dom.halo_retrieve = lambda: log.append(("halo_retrieve",))
        dom.apply_forcing = lambda dt: log.append(("apply_forcing", float(dt.seconds())))
        dom.enforce_limits = lambda: log.append(("enforce_limits",))
        false_if = [(r"^if\s*\(options%parameters%interactive.*then$", "if (.false.) then")]
        mod = Module(self._time_step_path,
                     overrides={"update_dt": update_dt,
                                "diagnostic_update": lambda M, domain, options: log.append(("diagnostic_update",)),
                                "print_progress": lambda M, *a: None},
                     externals=dict({n: recorder(n) for n in ("rad", "lsm", "pbl", "convect", "mp", "advect")},
                                    domain_check=None, str=lambda *a, **k: ""),
                     patches={"step": false_if}).build(["step"])
        mod.step(dom, Time(np.float64(end_seconds) / np.float64(86400.0)), ns(parameters=ns(interactive=False, debug=False)))
        return log, float(dom.model_time.seconds())

    def mp_schedule(self, calls, update_interval, grid):
        """mp(domain, options, dt_in, halo, subset), mp_driver.f90:673-772, with process_subdomain / process_halo replaced
        by recorders: for each (model_time, dt_in, halo, subset) of `calls` the list of (routine, mp_dt, its, ite, jts,
        jte, kts, kte) it would run -- the update_interval test, the mp_dt lag and the tile arithmetic, as written.
        grid: dict with ids..kde, ims..kme, its..kte, nx, ny, nz."""
        ns = types.SimpleNamespace
        log = []

        def rec(name):
            def f(M, domain, options, mp_dt, *a, **k):
                names = ("its", "ite", "jts", "jte", "kts", "kte")
                b = tuple(int(k[n]) for n in names) if k else tuple(int(x) for x in a[:6])     # keyword or positional call
                log[-1].append((name, float(mp_dt)) + b)
            return f
        path = os.path.join(os.path.dirname(self.mp_path), "mp_driver.f90")
        # process_halo (:609-658) runs as written: its four ring strips arrive at the process_subdomain recorder
        mod = Module(path, overrides={"process_subdomain": rec("process_subdomain"),
                                      "allocate_module_variables": lambda M, *a: None})
        drv = mod.build(["mp"])
        drv.M.update_interval = f32(update_interval) if isinstance(drv.M.update_interval, np.float32) else int(update_interval)
        drv.M.last_model_time = type(drv.M.last_model_time)(-999)
        opts = ns(mp_options=ns(top_mp_level=0))
        for t, dt, halo, subset in calls:
            log.append([])
            dom = ns(grid=ns(**grid), model_time=ns(seconds=lambda t=t: np.float64(t)))
            drv.mp(dom, opts, f32(dt), halo, subset)
        return log

    def grid_dimensions(self, nx, ny, nz, nimages, image, halo, decomposition="x"):
        """grid_t%set_grid_dimensions (src/objects/grid_obj.f90:144-255: my_n, my_start, update_with_halos) for image
        `image` of `nimages`.  decomposition = "x": the slabs along x that BASELINE.json prescribes (ximages = nimages,
        yimages = 1 -- domain_decomposition's search over splits replaced by that choice), "y": slabs along y,
        "reference": domain_decomposition as written (the most square split).  Returns the grid_t integers."""
        ns = types.SimpleNamespace

        def slabs(M, this, gnx, gny, nimg, for_image=None, ratio=None):
            this.ximages, this.yimages = (int(nimg), 1) if decomposition == "x" else (1, int(nimg))
            this.ximg, this.yimg = (int(for_image), 1) if decomposition == "x" else (1, int(for_image))
        root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
        mod = Module(os.path.join(root, "src", "objects", "grid_obj.f90"), externals={"assert": None},
                     names={"assertions": False, "kdefault_halo_size": 1, "_fn_num_images": lambda: int(nimages),
                            "_fn_this_image": lambda: int(image)},
                     overrides={} if decomposition == "reference" else {"domain_decomposition": slabs})
        mod.component_dtype = {"dimensions": "object"}
        g = mod.build(["set_grid_dimensions"])
        this = ns(dimensions=None)
        g.set_grid_dimensions(this, int(nx), int(ny), int(nz), None, None, int(halo), int(image))
        return {k: v for k, v in vars(this).items() if isinstance(v, int) and not isinstance(v, bool)}

    # ---- linear_winds.f90 (+ fftshift.f90): the per-height solution, the layer means, the buffered terrain
    def linear_winds(self):
        """Translates the routines of src/physics/linear_winds.f90 that the cuFFT path restates and binds FFTW's entry
        points to numpy.fft (an unnormalised backward / forward 2-D DFT, what fftw_execute_dft computes).  Returns the
        namespace; `lt = ref.lt_data(nx, ny, dx)` then makes a linear_theory_type through initialize_linear_theory_data."""
        if hasattr(self, "_lw"):
            return self._lw
        root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
        shift = Module(os.path.join(root, "src", "utilities", "fftshift.f90")).build(["ifftshift2cc", "fftshift2cc"])

        def execute_dft(plan, src, dst):
            a = np.asarray(src)
            dst[...] = (np.fft.ifft2(a) * a.size) if plan == "backward" else np.fft.fft2(a)
        mod = Module(os.path.join(self.mp_path.rsplit("/", 1)[0], "linear_winds.f90"),
                     externals={"ifftshift": shift.ifftshift2cc, "fftshift": shift.fftshift2cc, "fftw_execute_dft": execute_dft,
                                "fftw_alloc_complex": lambda n: np.zeros(int(n), np.complex128),
                                "fftw_plan_dft_2d": lambda n0, n1, a, b, sign, flags: "backward" if sign > 0 else "forward",
                                "smooth_array": None, "io_write": None},
                     names={"pi": self.constants.pi, "fftw_backward": 1, "fftw_forward": -1, "fftw_measure": 0,
                            "fftw_estimate": 64})
        mod.component_dtype = {k: "_np.complex128" for k in ("denom", "m", "msq", "mimag", "ineta")}
        self._lw = mod.build(["initialize_linear_theory_data", "linear_perturbation_at_height", "linear_perturbation_constz",
                              "linear_perturbation_varyingz", "add_buffer_topo"])
        self._lw.shift = shift
        return self._lw

    def lt_data(self, nx, ny, dx):
        lw = self.linear_winds()
        lt = types.SimpleNamespace()
        lw.initialize_linear_theory_data(lt, int(nx), int(ny), f32(dx))
        return lt

    # ---- domain_obj.f90: apply_forcing, enforce_limits (the rest of step()'s loop body)
    def _domain_module(self):
        if not hasattr(self, "_dom"):
            root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
            self._dom = Module(os.path.join(root, "src", "objects", "domain_obj.f90")).build(
                ["apply_forcing", "enforce_limits", "halo_send", "halo_retrieve", "halo_exchange"])
        return self._dom

    def apply_forcing(self, variables, w, seconds, west=True, east=True, south=True, north=True):
        """domain%apply_forcing(dt), domain_obj.f90:2383-2448.  variables: [(data (ny, nz, nx) float32, dqdt, force_boundaries)]
        in the order of the variables_to_force dictionary; w: (data, dqdt) of this%w%meta_data; updated in place."""
        ns = types.SimpleNamespace
        fa = lambda a: FArr(a.T, (1, 1, 1))                    # noqa: E731

        class Dictionary:
            def __init__(self, items):
                self.items, self.pos = items, 0

            def reset_iterator(self):
                self.pos = 0

            def has_more_elements(self):
                return self.pos < len(self.items)

            def next(self):
                self.pos += 1
                return self.items[self.pos - 1]
        items = [ns(two_d=False, three_d=True, force_boundaries=bool(fb), data_3d=fa(d), dqdt_3d=fa(t), data_2d=None, dqdt_2d=None)
                 for d, t, fb in variables]
        this = ns(variables_to_force=Dictionary(items), west_boundary=west, east_boundary=east, south_boundary=south,
                  north_boundary=north, w=ns(meta_data=ns(data_3d=fa(w[0]), dqdt_3d=fa(w[1]))),
                  external_precipitation=ns(data_2d=None), accumulated_precipitation=ns(data_2d=None, data_2dd=None))
        self._domain_module().apply_forcing(this, ns(seconds=lambda: np.float64(seconds)))

    def enforce_limits(self, fields):
        """domain%enforce_limits(), domain_obj.f90:2228-2243; fields: component name -> (ny, nz, nx) float32 array, in place"""
        ns = types.SimpleNamespace
        names = ("water_vapor", "potential_temperature", "cloud_water_mass", "cloud_number", "cloud_ice_mass",
                 "cloud_ice_number", "rain_mass", "rain_number", "snow_mass", "snow_number", "graupel_mass", "graupel_number")
        this = ns(**{n: ns(data_3d=FArr(fields[n].T, (1, 1, 1)) if n in fields else None) for n in names})
        self._domain_module().enforce_limits(this)

    def domain_halo_calls(self, allocated, routine="halo_exchange"):
        """domain%halo_send / halo_retrieve / halo_exchange, domain_obj.f90:109-155, on a mock domain whose members in
        `allocated` have their data_3d associated and record what is called on them: [(member, "send" | "retrieve",
        no_sync)] in the order of the calls."""
        ns = types.SimpleNamespace
        names = ("water_vapor", "potential_temperature", "cloud_water_mass", "cloud_number", "cloud_ice_mass",
                 "cloud_ice_number", "rain_mass", "rain_number", "snow_mass", "snow_number", "graupel_mass", "graupel_number")
        unknown = set(allocated) - set(names)
        if unknown:
            raise KeyError(f"not an exchanged member of domain_t: {sorted(unknown)}")
        log = []

        def member(n):
            return ns(data_3d=FArr(np.zeros((1, 1, 1), np.float32), (1, 1, 1)) if n in allocated else None,
                      send=lambda: log.append((n, "send", False)),
                      retrieve=lambda no_sync=False: log.append((n, "retrieve", bool(no_sync))))
        this = ns(**{n: member(n) for n in names})
        getattr(self._domain_module(), routine)(this)
        return log

    # ---- exchangeable_obj.f90: the halo exchange (coarray PUTs into the neighbours' halo buffers, then copies home)
    def halo_exchange_x(self, tiles, halo):
        """exchangeable_t%send then %retrieve (src/objects/exchangeable_obj.f90:138-162, 249-356) for a row of images along
        x.  tiles: one (ny, nz, nx_local) float32 array per image, ghost columns included, updated in place.  Every image's
        module variables (halo_size, the neighbour numbers, :4-7) are set before its calls, as each image holds its own."""
        ns = types.SimpleNamespace
        root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
        n = len(tiles)
        objs = {}
        mod = Module(os.path.join(root, "src", "objects", "exchangeable_obj.f90"), names={"_remote": lambda this, img: objs[int(img)]})
        ex = mod.build(["send", "retrieve"])
        for p, a in enumerate(tiles, 1):
            ny, nz, nxl = a.shape
            buf = lambda shape: FArr(np.full(shape, np.nan, np.float32, order="F"), (1, 1, 1))     # noqa: E731
            objs[p] = ns(data_3d=FArr(a.T, (1, 1, 1)), north_boundary=True, south_boundary=True, west_boundary=(p == 1),
                         east_boundary=(p == n), halo_west_in=buf((halo + 1, nz, ny + 2 * halo)),
                         halo_east_in=buf((halo + 1, nz, ny + 2 * halo)), halo_north_in=None, halo_south_in=None)

        def as_image(p):
            ex.M.halo_size, ex.M.west_neighbor, ex.M.east_neighbor = int(halo), p - 1, p + 1
            ex.M.north_neighbor = ex.M.south_neighbor = 0
        for p in range(1, n + 1):
            as_image(p)
            ex.send(objs[p])
        for p in range(1, n + 1):                            # after `sync images`: every PUT has landed
            as_image(p)
            ex.retrieve(objs[p], None)

    # ---- linear_winds.f90: spatial_winds, the routine linear_perturb runs every forcing interval (table look-up per cell)
    def spatial_winds(self, st, lut_u, lut_v, o, u_pert, v_pert):
        """spatial_winds(domain, reverse=.False., vsmooth, winsz, update=.False.), linear_winds.f90:840-1127, on a mock domain.
        st: u (ny, nz, nx+1), v (ny+1, nz, nx), th, pii, z, qv [, qc, qi, qr, qs] (ny, nz, nx), updated in place (u, v).
        lut_u (n_spd, n_dir, n_nsq, nx+1, nz, ny), lut_v (n_spd, n_dir, n_nsq, nx, nz, ny+1): hi_u_LUT / hi_v_LUT in the
        Fortran's index order; u_pert (ny, nz, nx+1), v_pert (ny+1, nz, nx): hi_u/v_perturbation, in place.  o: the
        lt_options (oracle.spatial_winds_np.LTOptions names).  calc_stability / calc_direction / calc_speed come from
        atm_utilities.f90, calc_weight / smooth_array / linear_space from array_utilities.f90, translated as well.
        Returns log N^2 -> N^2 as the routine leaves it in domain%nsquared, (ny, nz, nx)."""
        ns = types.SimpleNamespace
        root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
        if not hasattr(self, "_sw"):
            consts = {k: v for k, v in vars(self.constants).items() if isinstance(v, (np.floating, int))}
            atm = Module(os.path.join(root, "src", "utilities", "atm_utilities.f90"), names=consts).build(
                ["calc_stability", "calc_direction", "calc_speed", "calc_u", "calc_v"])
            arr = Module(os.path.join(root, "src", "utilities", "array_utilities.f90")).build(
                ["calc_weight", "smooth_array_3d", "linear_space"])
            mod = Module(os.path.join(self.mp_path.rsplit("/", 1)[0], "linear_winds.f90"),
                         externals={"calc_stability": atm.calc_stability, "calc_direction": atm.calc_direction,
                                    "calc_speed": atm.calc_speed, "calc_weight": (arr.calc_weight, [2]),
                                    "smooth_array": arr.smooth_array_3d},
                         names={"pi": self.constants.pi})
            self._sw = (mod.build(["spatial_winds"]), atm, arr)
        sw, atm, arr = self._sw
        ny, nz, nxu = st["u"].shape
        nx, nyv = nxu - 1, ny + 1
        fa = lambda a: FArr(a.T, (1,) * a.ndim)                # noqa: E731
        var = lambda a: ns(data_3d=fa(a) if a is not None else None)    # noqa: E731
        nsq = np.zeros((ny, nz, nx), np.float32)
        dom = ns(u=ns(data_3d=fa(st["u"]), meta_data=ns(dqdt_3d=None)), v=ns(data_3d=fa(st["v"]), meta_data=ns(dqdt_3d=None)),
                 nsquared=var(nsq), latitude=ns(data_2d=FArr(np.zeros((nx, ny), np.float32, order="F"), (1, 1))),
                 cloud_water_mass=var(st.get("qc")), cloud_ice_mass=var(st.get("qi")), rain_mass=var(st.get("qr")),
                 snow_mass=var(st.get("qs")), potential_temperature=var(st["th"]), exner=var(st["pii"]), z=var(st["z"]),
                 water_vapor=var(st["qv"]))
        M = sw.M
        M.variable_n, M.smooth_nsq = bool(o.variable_N), bool(o.smooth_nsq)
        M.n_squared, M.min_stability, M.max_stability = f32(o.N_squared), f32(o.min_stability), f32(o.max_stability)
        M.linear_update_fraction, M.linear_contribution = f32(o.linear_update_fraction), f32(o.linear_contribution)
        M.n_dir_values, M.n_nsq_values, M.n_spd_values = int(o.n_dir_values), int(o.n_nsq_values), int(o.n_spd_values)
        for name, lo, hi, n in (("dir_values", o.dirmin, o.dirmax, o.n_dir_values), ("nsq_values", o.nsqmin, o.nsqmax, o.n_nsq_values),
                                ("spd_values", o.spdmin, o.spdmax, o.n_spd_values)):
            axis = FArr(None, (1,), np.float32, True)                   # an unallocated allocatable, as at :656-660
            arr.linear_space(axis, f32(lo), f32(hi), int(n))
            setattr(M, name, axis)
        M.hi_u_lut = FArr(np.asfortranarray(lut_u), (1,) * 6)
        M.hi_v_lut = FArr(np.asfortranarray(lut_v), (1,) * 6)
        M.hi_u_perturbation, M.hi_v_perturbation = fa(u_pert), fa(v_pert)
        M.lo_u_perturbation = M.lo_v_perturbation = None
        sw.spatial_winds(dom, False, int(o.vert_smooth), int(o.stability_window_size), False)
        return nsq

    def spatial_winds_table(self, terrain_xy, dx, o, z_interface=None, dz_interface=None, layer_height=None, dz_levels=None):
        """initialize_spatial_winds (linear_winds.f90:596-838) on ONE image: the whole look-up table hi_u_LUT / hi_v_LUT
        (n_spd, n_dir, n_nsq, nx+1 | nx, nz, ny | ny+1) filled by the routine's own loop -- the ijk -> (direction, stability,
        speed) mapping, calc_u / calc_v, exp(nsq), the layer bounds (terrain-following from z_interface / dz_interface
        (nx, nz, ny), or constant from layer_height / dz_levels), linear_perturbation, the staggering, copy_data_to_remote.
        Replaced by the harness: setup_remote_grids (one image owns the whole domain; the grid algebra has its own test),
        the file-name string and LUT file I/O (not taken), FFTW (numpy.fft).  terrain_xy (nx, ny); the terrain frequency is
        the buffered, transformed terrain of setup_linwinds (:1201-1222), taken from the restatement that equals it."""
        from oracle import linear_winds_np as LW
        ns = types.SimpleNamespace
        root = os.path.dirname(os.path.dirname(os.path.dirname(self.mp_path)))
        consts = {k: v for k, v in vars(self.constants).items() if isinstance(v, (np.floating, int))}
        atm = Module(os.path.join(root, "src", "utilities", "atm_utilities.f90"), names=consts).build(["calc_u", "calc_v"])
        arr = Module(os.path.join(root, "src", "utilities", "array_utilities.f90")).build(["linear_space"])
        shift = Module(os.path.join(root, "src", "utilities", "fftshift.f90")).build(["ifftshift2cc"])
        nx, ny = terrain_xy.shape
        varying = z_interface is not None
        nz = z_interface.shape[1] if varying else len(dz_levels)
        built = {}

        def execute_dft(plan, src, dst):
            a = np.asarray(src)
            dst[...] = np.fft.ifft2(a) * a.size

        def grids(M, u_grids, v_grids, terrain, nz_):
            u_grids[1] = ns(nx=nx + 1, ny=ny, ims=1, ime=nx + 1, jms=1, jme=ny)
            v_grids[1] = ns(nx=nx, ny=ny + 1, ims=1, ime=nx, jms=1, jme=ny + 1)

        def perturbation(M, U, V, Nsq, zb, zt, step, ft, lt):           # the generic interface :71-74: by the rank of the bounds
            fn = built["ns"].linear_perturbation_varyingz if np.ndim(zb) else built["ns"].linear_perturbation_constz
            fn(U, V, Nsq, np.asfortranarray(zb) if np.ndim(zb) else zb, np.asfortranarray(zt) if np.ndim(zt) else zt, step, ft, lt)
        lut = {}
        mod = Module(os.path.join(self.mp_path.rsplit("/", 1)[0], "linear_winds.f90"),
                     externals={"calc_u": atm.calc_u, "calc_v": atm.calc_v, "linear_space": arr.linear_space,
                                "read_lut": lambda *a: 1, "write_lut": lambda *a: 0, "ifftshift": shift.ifftshift2cc,
                                "fftw_execute_dft": execute_dft, "fftw_alloc_complex": lambda n: np.zeros(int(n), np.complex128),
                                "fftw_plan_dft_2d": lambda *a: "backward"},
                     overrides={"setup_remote_grids": grids, "linear_perturbation": perturbation,
                                "destroy_linear_theory_data": lambda M, *a: None},
                     names={"pi": self.constants.pi, "fftw_backward": 1, "fftw_measure": 0,
                            "_remote": lambda table, img: table, "_fn_num_images": lambda: 1, "_fn_this_image": lambda: 1},
                     patches={"initialize_spatial_winds": [(r"^lut_file\s*=", 'lut_file = "unused"')]})
        mod.component_dtype = {k: "_np.complex128" for k in ("denom", "m", "msq", "mimag", "ineta")}
        built["ns"] = sw = mod.build(["initialize_spatial_winds", "linear_perturbation_constz", "linear_perturbation_varyingz",
                                      "initialize_linear_theory_data", "copy_data_to_remote"])
        tf, total = LW.terrain_frequency(terrain_xy, o.buffer)
        M = sw.M
        M.buffer, M.lt_data_m = int(total), ns()
        M.n_dir_values, M.n_nsq_values, M.n_spd_values = int(o.n_dir_values), int(o.n_nsq_values), int(o.n_spd_values)
        M.dirmin, M.dirmax, M.nsqmin, M.nsqmax = f32(o.dirmin), f32(o.dirmax), f32(o.nsqmin), f32(o.nsqmax)
        M.spdmin, M.spdmax, M.minimum_layer_size = f32(o.spdmin), f32(o.spdmax), f32(o.minimum_layer_size)
        f3 = lambda a: FArr(np.asfortranarray(a, np.float32), (1,) * np.ndim(a))            # noqa: E731
        dom = ns(z=ns(data_3d=f3(np.zeros((nx, nz, ny), np.float32) if layer_height is None else
                                  np.broadcast_to((np.asarray(layer_height, np.float32))[None, :, None], (nx, nz, ny)).copy())),
                 u=ns(data_3d=f3(np.zeros((nx + 1, nz, ny), np.float32))), global_terrain=f3(terrain_xy),
                 terrain=ns(data_2d=f3(np.zeros((nx, ny), np.float32))), terrain_frequency=FArr(np.asfortranarray(tf), (1, 1)),
                 dx=f32(dx), global_z_interface=f3(z_interface) if varying else None,
                 global_dz_interface=f3(dz_interface) if varying else None)
        opts = ns(lt_options=ns(u_lut_filename="", read_lut=False, write_lut=False),
                  parameters=ns(space_varying_dz=bool(varying), interactive=False,
                                dz_levels=FArr(np.asarray(dz_levels if dz_levels is not None else np.zeros(nz), np.float32), (1,))))
        sw.initialize_spatial_winds(dom, opts, False)
        return M.hi_u_lut.a, M.hi_v_lut.a

    def exner_function(self, pressure):
        """exner_function (atm_utilities.f90:682-691), elemental, with Rd and cp of icar_constants.f90"""
        return self.atm.exner_function(np.ascontiguousarray(pressure, np.float32))

    def mpdata(self, fields, u, v, w, jacobian_u, jacobian_v, jacobian_w, jacobian, advection_dz, dt, dx,
               mpdata_order=2, fct=True):
        """mpdata(domain, options, dt), adv_mpdata.f90:463-524, on a mock domain: `fields` maps kVARS names
        (ADVECTED) to (ny, nz, nx) float32 arrays, advected in place in the reference's order; u (ny, nz, nx+1),
        v (ny+1, nz, nx).  The module arrays U_m, V_m, W_m are (re)allocated by the routine itself."""
        dom, opts = self._mock_domain(fields, u, v, w, jacobian_u, jacobian_v, jacobian_w, jacobian, advection_dz, dx,
                                      mpdata_order, fct)
        M = self.adv.M
        for n in ("u_m", "v_m", "w_m"):                  # `if (.not.allocated(U_m))`: a new grid needs new module arrays
            getattr(M, n).a = None
        self.adv.mpdata(dom, opts, f32(dt))

    def set_winds(self, U_m, V_m, W_m):
        """the module arrays U_m (ims+1:ime, :, :), V_m (:, :, jms+1:jme), W_m, given as (ny, nz, nx-1), (ny-1, nz, nx),
        (ny, nz, nx) arrays (adv_mpdata.f90:492-495 bounds, ims = kms = jms = 1)"""
        M = self.adv.M
        M.u_m = FArr(np.asfortranarray(np.asarray(U_m, np.float32).T), (2, 1, 1))
        M.v_m = FArr(np.asfortranarray(np.asarray(V_m, np.float32).T), (1, 1, 2))
        M.w_m = FArr(np.asfortranarray(np.asarray(W_m, np.float32).T), (1, 1, 1))

    def advect3d(self, q, jaco, dz, mpdata_order=2, fct=True, rho=None):
        """advect3d(q, rho, ims, ime, kms, kme, jms, jme, jaco, dz, options), adv_mpdata.f90:356; q (ny, nz, nx) is
        updated in place and returned"""
        qc, qf = _f(q)
        ny, nz, nx = qc.shape
        rho = np.ones_like(qc) if rho is None else rho
        opts = types.SimpleNamespace(adv_options=types.SimpleNamespace(mpdata_order=int(mpdata_order),
                                                                       flux_corrected_transport=bool(fct)))
        self.adv.advect3d(qf, _f(rho)[1], 1, nx, 1, nz, 1, ny, _f(jaco)[1], _f(dz)[1], opts)
        return qc

    def upwind_advection(self, qin, u, v, w, rho, dz, jaco):
        """upwind_advection (adv_mpdata.f90:44): returns q; u (ny, nz, nx-1), v (ny-1, nz, nx), w (ny, nz, nx)"""
        qc, qf = _f(qin)
        ny, nz, nx = qc.shape
        out = np.zeros_like(qc)
        self.adv.upwind_advection(qf, _f(u)[1], _f(v)[1], _f(w)[1], _f(rho)[1], out.T, _f(dz)[1], 1, nx, 1, nz, 1, ny, _f(jaco)[1])
        return out

    # ---- mp_simple.f90
    def mp_simple_driver(self, pressure, th, pii, rho, qv, qc, qr, qs, rain, snow, dt, dz, tile=None):
        """mp_simple_driver (mp_simple.f90:595) on (ny, nz, nx) / (ny, nx) float32 arrays, updated in place.
        tile = (its, ite, jts, jte, kts, kte), 1-based inclusive; default the whole memory tile."""
        arrs = [np.ascontiguousarray(a, np.float32) for a in (pressure, th, pii, rho, qv, qc, qr, qs)]
        for a, b in zip(arrs, (pressure, th, pii, rho, qv, qc, qr, qs)):
            if a is not b:
                raise ValueError("mp_simple_driver needs C-contiguous float32 arrays (updated in place)")
        ny, nz, nx = arrs[0].shape
        its, ite, jts, jte, kts, kte = tile or (1, nx, 1, ny, 1, nz)
        self.mp.mp_simple_driver(*[a.T for a in arrs], rain.T, snow.T, f32(dt), _f(dz)[1],
                                 1, nx, 1, ny, 1, nz, its, ite, jts, jte, kts, kte)
