"""Find the worst cell of a per-call comparison and test whether the oracle itself is discontinuous there
(re-run the oracle on inputs perturbed by 1 ulp)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import icar_b200
from icar_b200 import ideal
from oracle import oracle as O

nx, nz, ny, dt = 37, 40, 9, 120.0
ref = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
dom = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(microphysics=1))
T = O.Thompson()
F = ("th", "qv", "qc", "qr", "qi", "qs", "qg", "ni", "nr")
for call in range(8):
    inp = ideal.copy_state(ref)
    st = ideal.copy_state(ref)
    dom.mp_gt_driver(st, dt, 2, nx - 1, 2, ny - 1, 1, nz)
    T.driver(ref, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
    d = np.abs(st["qg"].astype(np.float64) - ref["qg"]) / max(ref["qg"].max(), 1e-30)
    j, k, i = np.unravel_index(np.argmax(d), d.shape)
    print(call, "worst qg", d.max(), "at", (i, k, j), "n cells > 1e-4:", int((d > 1e-4).sum()))
    if d.max() > 1e-3:
        print("column (i,j)=", i, j)
        for f in F:
            print(f"  {f:3s} in  ", " ".join(f"{x:.4e}" for x in inp[f][j, max(0,k-3):k+3, i]))
            print(f"  {f:3s} gpu ", " ".join(f"{x:.4e}" for x in st[f][j, max(0,k-3):k+3, i]))
            print(f"  {f:3s} ref ", " ".join(f"{x:.4e}" for x in ref[f][j, max(0,k-3):k+3, i]))
        # oracle sensitivity: perturb inputs by 1 ulp
        for name in ("th", "qv", "qr", "qg", "nr"):
            for sgn in (+1, -1):
                p = ideal.copy_state(inp)
                p[name] = np.nextafter(p[name], p[name] + np.float32(sgn) * np.abs(p[name]) - np.float32(0) if False else (np.float32(np.inf) * sgn)).astype(np.float32)
                T.driver(p, dt, 1, nx - 2, 1, ny - 2, 0, nz - 1)
                dd = np.abs(p["qg"].astype(np.float64) - ref["qg"]) / max(ref["qg"].max(), 1e-30)
                print(f"  oracle with {name} {'+' if sgn > 0 else '-'}1ulp: qg at worst cell changes by {dd[j, k, i]:.3e}; max anywhere {dd.max():.3e}")
        break
