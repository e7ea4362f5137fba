"""Golden vectors made by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCE (oracle/f90run): src/physics/adv_mpdata.f90 with
the included adv_mpdata_FCT_core.f90, and src/physics/mp_simple.f90, read where they lie under /root/reference,
translated statement by statement to Python/NumPy with IEEE binary32 arithmetic and glibc's expf, and run.

Unlike make_golden.py (frozen from the hand-written oracle), these files pin the oracle -- and through it the CUDA
path -- to outputs of the reference's text:

  reference_source_hill_24x10x20_3steps_f32.npz   three coupled model steps (mp_simple_driver on the interior tile, then
        mpdata(domain, options, dt) itself: wind normalisation :500-506 and the scalar list :512-522) of the reduced
        BASELINE config-1 hill case
  reference_source_mpdata_scenario_f32.npz        the scenario of src/tests/test_mpdata.f90:51-111 (two loops)
  reference_source_mp_simple_scenario_f32.npz     the scenario of src/tests/test_mp_simple.f90:4-33 (100 calls of mp_simple)
  reference_source_thompson.npz                   src/physics/mp_thompson.f90: thompson_init's 130 scalar constants, the bin
        grids, the tables of freezeH2O / qi_aut_qs / table_Efrw / table_Efsw (2-D ones whole, 3-D ones sampled), three (k, m)
        slabs of each rain-collection table from qr_acr_qg / qr_acr_qs themselves, and mp_gt_driver + mp_thompson: six calls
        on a moist 6 x 16 x 4 hill state in which every species forms.  (The remaining slabs of the two collection tables
        -- 1.4e10 inner iterations -- are taken from the C++ oracle for the run.)
  reference_source_random_f32.npz                 advect3d (orders 1-3, FCT on / off) and mp_simple_driver (sub-tile) on
        random fields of a ragged 13 x 7 x 9 grid, inputs included

Re-run (needs /root/reference; about a quarter of an hour):  python tests/golden/make_reference_source_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import f90run  # noqa: E402
from icar_b200 import ideal  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCALARS = {"water_vapor": "qv", "cloud_water": "qc", "rain_in_air": "qr", "snow_in_air": "qs", "potential_temperature": "th"}


def reference_step(ref, st, dt):
    """mp(domain) on the interior tile (the tile of a single image stops one cell short of every domain edge,
    grid_obj.f90:241-250) followed by advect(domain) = mpdata(domain, options, dt)"""
    ny, nz, nx = st["w"].shape
    ref.mp_simple_driver(st["pressure"], st["th"], st["pii"], st["rho"], st["qv"], st["qc"], st["qr"], st["qs"],
                         st["rain"], st["snow"], dt, st["dz_mass"], tile=(2, nx - 1, 2, ny - 1, 1, nz))
    ref.mpdata({k: st[v] for k, v in SCALARS.items()}, st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"],
               st["jacobian_w"], st["jacobian"], st["advection_dz"], dt, st["dx"])


def hill_case(ref, nx=24, nz=10, ny=20, steps=3):
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, terrain=True)
    st["qv"] *= np.float32(3.0)
    dt = np.float32(st["dt"])
    for _ in range(steps):
        reference_step(ref, st, dt)
    return {k: st[k] for k in ("th", "qv", "qc", "qr", "qs", "rain", "snow")}, float(dt)


def mpdata_scenario(ref, nloops=2):
    """src/tests/test_mpdata.f90:51-111 (test_3d), as make_golden.mpdata_scenario states it"""
    nx, nz, ny = 3, 3, 100
    cfl = 0.25
    f = np.float32
    q = np.ones((ny, nz, nx), f)
    q[ny // 2:, :, :] = 2
    ref.set_winds(np.zeros((ny, nz, nx - 1), f), np.full((ny - 1, nz, nx), cfl, f), np.zeros((ny, nz, nx), f))
    dz = np.ones((ny, nz, nx), f)
    jaco = np.ones((ny, nz, nx), f)
    q[0] = q[ny - 2]
    q[ny - 1] = q[1]
    rows = []
    for loop in range(nloops):
        for _ in range(int((ny - 2) / cfl)):
            ref.advect3d(q, jaco, dz, 2, True)
            q[0] = q[ny - 2]
            q[ny - 1] = q[1]
        line = q[:, 0, 1]
        rows.append(np.concatenate([line[::ny // 10], [line.max(), line.min()]]))
    return np.array(rows, f), q[:, 0, 1].copy()


def mp_simple_scenario(ref):
    """src/tests/test_mp_simple.f90:4-33: a 5-level column cooled by 0.1 K per call, 100 calls of mp_simple"""
    nz = 5
    f = np.float32
    p = np.full(nz, 80000.0, f); T = np.full(nz, 280.0, f); rho = np.ones(nz, f); qv = np.full(nz, 0.005, f)
    qc = np.zeros(nz, f); qr = np.zeros(nz, f); qs = np.zeros(nz, f); dz = np.full(nz, 200.0, f)
    rain = f(0); snow = f(0)
    M = ref.mp.M
    dt = f(20.0)
    # what mp_simple_driver sets before it calls mp_simple (mp_simple.f90:619-622)
    M.cloud2snow = ref.exp4(f(-1.0) * M.snow_formation_time_const * dt)
    M.cloud2rain = ref.exp4(f(-1.0) * M.rain_formation_time_const * dt)
    M.snow_evap = ref.exp4(f(-1.0) * M.snow_evap_time_const * dt)
    M.rain_evap = ref.exp4(f(-1.0) * M.rain_evap_time_const * dt)
    out = []
    for _ in range(100):
        rain, snow = ref.mp.mp_simple(p, T, rho, qv, qc, qr, qs, rain, snow, dt, dz, 1, nz, 1, nz)
        out.append([T[0], qv[0], qc[0], qr[0], qs[0], rain, snow])
        T -= f(0.1)
    return np.array(out, f)


def random_cases(ref, nx=13, nz=7, ny=9, seed=3):
    rng = np.random.default_rng(seed)
    f = np.float32
    rnd = lambda shape, lo, hi: rng.uniform(lo, hi, shape).astype(f)      # noqa: E731
    out = {}
    q = rnd((ny, nz, nx), 0.0, 2.0)
    q[rng.random(q.shape) < 0.3] = 0
    out.update(q=q, U=rnd((ny, nz, nx - 1), -0.3, 0.3), V=rnd((ny - 1, nz, nx), -0.3, 0.3), W=rnd((ny, nz, nx), -20, 20),
               jaco=rnd((ny, nz, nx), 0.7, 1.0), dz=rnd((ny, nz, nx), 300, 400))
    ref.set_winds(out["U"], out["V"], out["W"])
    for order, fct in ((1, True), (2, True), (2, False), (3, True)):
        out[f"advect3d_order{order}_fct{int(fct)}"] = ref.advect3d(q.copy(), out["jaco"], out["dz"], order, fct)
    # microphysics: a moist, partly supersaturated random state, three calls on a sub-tile
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, terrain=True, perturb=True)
    mp = {k: st[k].copy() for k in ("pressure", "th", "pii", "rho", "dz_mass")}
    mp["qv"] = (st["qv"] * rnd((ny, nz, nx), 0.5, 4.0)).astype(f)
    mp["qc"] = rnd((ny, nz, nx), 0, 2e-3); mp["qr"] = rnd((ny, nz, nx), 0, 1e-3); mp["qs"] = rnd((ny, nz, nx), 0, 1e-3)
    mp["qr"][rng.random(q.shape) < 0.5] = 0
    mp["qs"][rng.random(q.shape) < 0.5] = 0
    for k, v in mp.items():
        out["mp_in_" + k] = v.copy()
    rain = np.zeros((ny, nx), f); snow = np.zeros((ny, nx), f)
    tile = (2, nx - 2, 3, ny - 1, 1, nz)
    out["mp_tile"] = np.array(tile)
    out["mp_dt"] = f(45.0)
    for _ in range(3):
        ref.mp_simple_driver(mp["pressure"], mp["th"], mp["pii"], mp["rho"], mp["qv"], mp["qc"], mp["qr"], mp["qs"], rain, snow,
                             out["mp_dt"], mp["dz_mass"], tile=tile)
    for k in ("th", "qv", "qc", "qr", "qs"):
        out["mp_out_" + k] = mp[k]
    out["mp_out_rain"], out["mp_out_snow"] = rain, snow
    return out


THOMPSON_KEYS = ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "rainnc", "snownc", "graupelnc", "sr")
# (table family, km): first, an inner and the last (k, m) slab of each builder
THOMPSON_SLABS = ((0, 0), (0, 5 * 37 + 11), (0, 36 * 37 + 36), (1, 0), (1, 20 * 37 + 7), (1, 36 * 37 + 36))


def thompson_state(nx=6, nz=16, ny=4):
    return ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True), moisten=3.0)


def thompson_case(ref):
    from oracle import oracle as O
    T = O.Thompson(big_tables=True)
    big = {n: T.table(n).T for n in ref.BIG_QG + ref.BIG_QS}
    ref.thompson_init(big)                                      # everything else is the reference's text, executed
    M = ref.th.M
    out = {"constants": ref.thompson_constants()}
    for n in ("dc", "di", "dr", "ds", "dg"):
        out["bins_" + n] = getattr(M, n).a.copy()
    rng = np.random.default_rng(7)
    for n in O.THOMPSON_TABLES[18:]:
        a = getattr(M, n.lower()).a
        flat = a.reshape(-1, order="F")                         # the Fortran's storage order
        if a.ndim == 2:
            out["table_" + n] = flat.copy()
        else:
            idx = np.sort(rng.choice(flat.size, size=2048, replace=False))
            out["table_" + n + "_idx"] = idx
            out["table_" + n] = flat[idx].copy()
    for which, km in THOMPSON_SLABS:
        for n, v in ref.thompson_big_slab(which, km).items():
            out[f"slab{km}_{n}"] = v
    st = thompson_state()
    for call in range(1, 7):
        ref.mp_gt_driver(st, 60.0)
        for k in THOMPSON_KEYS:
            out[f"call{call}_{k}"] = st[k].copy()
    return out


if __name__ == "__main__":
    ref = f90run.Reference("/root/reference")
    if "--thompson-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "reference_source_thompson.npz"), **thompson_case(ref))
        print("written")
        sys.exit(0)
    fields, dt = hill_case(ref)
    np.savez_compressed(os.path.join(HERE, "reference_source_hill_24x10x20_3steps_f32.npz"), dt=dt, **fields)
    rows, line = mpdata_scenario(ref)
    np.savez_compressed(os.path.join(HERE, "reference_source_mpdata_scenario_f32.npz"), rows=rows, final_line=line)
    np.savez_compressed(os.path.join(HERE, "reference_source_mp_simple_scenario_f32.npz"), table=mp_simple_scenario(ref))
    np.savez_compressed(os.path.join(HERE, "reference_source_random_f32.npz"), **random_cases(ref))
    np.savez_compressed(os.path.join(HERE, "reference_source_thompson.npz"), **thompson_case(ref))
    print("written")
