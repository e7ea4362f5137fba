"""Exploratory comparison of the CUDA Thompson path with the CPU oracle (run on the GPU box)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import icar_b200
from icar_b200 import ideal
from oracle import oracle as O

nx, nz, ny = 64, 20, 24
st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
dom = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(microphysics=1))
t0 = time.time(); dom.thompson_init(); print("gpu thompson_init %.3f s" % (time.time() - t0))
t0 = time.time(); T = O.Thompson(); print("oracle thompson_init %.1f s" % (time.time() - t0))
for name in icar_b200.THOMPSON_TABLES:
    a, b = dom.thompson_table(name), T.table(name)
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-300)
    print(f"{name:10s} max|b| {np.abs(b).max():.3e}  max abs {d.max():.3e}  max rel {rel[np.abs(b) > 0].max() if (np.abs(b) > 0).any() else 0:.3e}  zeros a/b {(a == 0).sum()}/{(b == 0).sum()}")
ref = ideal.copy_state(st)
dt = 60.0
for it in range(12):
    dom.mp_gt_driver(st, dt, 2, nx - 1, 2, ny - 1, 1, nz)
    T.driver(ref, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
    line = []
    for k in ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr", "rainnc", "snownc", "graupelnc", "sr"):
        a, b = st[k].astype(np.float64), ref[k].astype(np.float64)
        mx = max(np.abs(b).max(), 1e-30)
        l2 = np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))
        line.append(f"{k} {np.abs(a - b).max() / mx:.1e}/{l2:.1e}")
    print(it, " ".join(line))
print({k: float(ref[k].max()) for k in ("qc", "qr", "qi", "qs", "qg", "ni", "nr", "rainnc")})
