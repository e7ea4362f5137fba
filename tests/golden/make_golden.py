"""Generates tests/golden/*.npz from the CPU oracle (oracle/).

The reference (NCAR/icar) ships no expected values for this path, only two scenarios without
asserts (src/tests/test_mpdata.f90:51-111, src/tests/test_mp_simple.f90:4-33), and it cannot be
compiled in this image (no Fortran compiler).  The vectors frozen here therefore come from the
oracle restatement, NOT from a run of the reference: they pin the oracle against regressions and
give the CUDA path fixed targets.  (Round 2: make_reference_source_golden.py makes the same scenarios by EXECUTING the
reference's Fortran source; tests/test_reference_source.py holds the two sets of files to equality.)
Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from icar_b200 import ideal  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mpdata_scenario(dtype=np.float32, nloops=10):
    """src/tests/test_mpdata.f90:51-111 (test_3d): step function advected in y at CFL 0.25 on a
    3 x 3 x 100 grid with wrap-around boundary cells, mpdata_order=2, FCT on."""
    nx, nz, ny = 3, 3, 100
    cfl = 0.25
    q = np.ones((ny, nz, nx), dtype)
    q[ny // 2:, :, :] = 2            # do i=1,ny: if (i>ny/2) q(:,:,i)=2   (1-based i -> rows 50..99)
    U = np.zeros((ny, nz, nx - 1), dtype)
    V = np.full((ny - 1, nz, nx), cfl, dtype)
    W = np.zeros((ny, nz, nx), dtype)
    dz = np.ones((ny, nz, nx), dtype)
    jaco = np.ones((ny, nz, nx), dtype)
    q[0] = q[ny - 2]
    q[ny - 1] = q[1]
    rows = []
    for loop in range(nloops):
        for _ in range(int((ny - 2) / cfl)):
            O.advect3d(q, None, jaco, dz, U, V, W, 2, True)
            q[0] = q[ny - 2]
            q[ny - 1] = q[1]
        line = q[:, 0, 1]
        rows.append(np.concatenate([line[::ny // 10], [line.max(), line.min()]]))
    return np.array(rows, dtype), q[:, 0, 1].copy()


def mp_simple_scenario():
    """src/tests/test_mp_simple.f90:4-33: 5-level column cooled by 0.1 K per call, 100 calls."""
    nz = 5
    f = np.float32
    p = np.full(nz, 80000.0, f)
    T = np.full(nz, 280.0, f)
    rho = np.ones(nz, f)
    qv = np.full(nz, 0.005, f)
    qc = np.zeros(nz, f)
    qr = np.zeros(nz, f)
    qs = np.zeros(nz, f)
    dz = np.full(nz, 200.0, f)
    rain = np.zeros(1, f)
    snow = np.zeros(1, f)
    out = []
    for _ in range(100):
        O.mp_simple_column(p, T, rho, qv, qc, qr, qs, rain, snow, 20.0, dz, 0, nz - 1)
        out.append([T[0], qv[0], qc[0], qr[0], qs[0], rain[0], snow[0]])
        T -= f(0.1)
    return np.array(out, f)


def hill_case(nx=24, nz=10, ny=20, steps=3):
    """Config-1 style ideal hill case, reduced: mp_simple + MPDATA(2, FCT) model steps."""
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, terrain=True)
    # moisten so that the microphysics is active from the first step
    st["qv"] *= np.float32(3.0)
    dt = np.float32(st["dt"])
    for _ in range(steps):
        O.step_simple(st, dt, dt, 2, True)
    return {k: st[k] for k in ("th", "qv", "qc", "qr", "qs", "rain", "snow")}, float(dt)


def thompson_column_case(nz=24, ncall=12, dt=60.0):
    """A single moist column (the ideal sounding, qv x 3) pushed through mp_thompson `ncall` times:
    cloud -> ice -> snow -> graupel -> rain all fire.  Frozen from the oracle (the oracle itself is pinned by reference_source_thompson.npz)."""
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=3, nz=nz, ny=3), moisten=3.0)
    T = O.Thompson(big_tables=True)
    out = []
    for _ in range(ncall):
        T.driver(st, dt, 1, 1, 1, 1, 0, nz - 1)
        out.append(np.stack([st[k][1, :, 1] for k in ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr")]))
    ppt = np.array([st[k][1, 1] for k in ("rainnc", "snownc", "graupelnc")], np.float32)
    return np.array(out, np.float32), ppt


def thompson_table_samples():
    """A fixed sample of entries of every lookup table (the full tables are 88 MB)."""
    T = O.Thompson(big_tables=True)
    rng = np.random.default_rng(7)
    out = {}
    for name in O.THOMPSON_TABLES:
        flat = T.table(name).ravel()
        idx = rng.choice(flat.size, size=min(64, flat.size), replace=False)
        out[name + "_idx"] = idx
        out[name] = flat[idx].copy()
    out["constants"] = T.constants()
    return out


if __name__ == "__main__":
    O.build()
    rows, line = mpdata_scenario()
    np.savez_compressed(os.path.join(HERE, "mpdata_scenario_f32.npz"), rows=rows, final_line=line)
    np.savez_compressed(os.path.join(HERE, "mp_simple_scenario_f32.npz"), table=mp_simple_scenario())
    fields, dt = hill_case()
    np.savez_compressed(os.path.join(HERE, "hill_24x10x20_3steps_f32.npz"), dt=dt, **fields)
    cols, ppt = thompson_column_case()
    np.savez_compressed(os.path.join(HERE, "thompson_column_f32.npz"), columns=cols, ppt=ppt)
    np.savez_compressed(os.path.join(HERE, "thompson_table_samples.npz"), **thompson_table_samples())
    print("written")
