"""Per-call comparison from IDENTICAL inputs: the oracle state is advanced, the GPU gets a copy each call."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import icar_b200
from icar_b200 import ideal
from oracle import oracle as O

nx, nz, ny = 64, 20, 24
ref = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
dom = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(microphysics=1))
dom.thompson_init(); T = O.Thompson()
dt = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
F = ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr")
for it in range(40):
    st = ideal.copy_state(ref)
    dom.mp_gt_driver(st, dt, 2, nx - 1, 2, ny - 1, 1, nz)
    T.driver(ref, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
    line = []
    badcols = np.zeros((ny, nx), bool)
    for k in F:
        a, b = st[k].astype(np.float64), ref[k].astype(np.float64)
        mx = max(np.abs(b).max(), 1e-30)
        d = np.abs(a - b) / mx
        bad = d > 1e-5
        badcols |= bad.any(axis=1)
        line.append(f"{k} {d.max():.1e}({int(bad.sum())})")
    print(it, " ".join(line), "badcols", int(badcols.sum()))
    if badcols.any() and it > 3 and not os.environ.get("QUIET"):
        j, i = np.argwhere(badcols)[0]
        for k in F:
            a, b = st[k][j, :, i], ref[k][j, :, i]
            kk = np.argmax(np.abs(a.astype(np.float64) - b))
            print(f"   col({i},{j}) {k}: worst level {kk}: gpu {a[kk]:.9e} ref {b[kk]:.9e}")
        os.environ["QUIET"] = "1"
