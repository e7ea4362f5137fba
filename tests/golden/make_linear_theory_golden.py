"""Generates tests/golden/linear_theory_1d.npz by RUNNING THE REFERENCE's own Python implementation of the
linear mountain-wave solution, helpers/ideal_linear.py:solve() (the 2-D x-z form of what
src/physics/linear_winds.f90:181-237 does in 3-D: same k grid from -pi/dx to +pi/dx inclusive, same
hydrostatic m with sign(sig), same evanescent branch, same  u_hat = -m sig k i eta / k^2).

    python tests/golden/make_linear_theory_golden.py        # needs /root/reference; the .npz is committed

The file is imported unmodified.  Three things around it are shimmed, none of which touches the arithmetic:
  * `import matplotlib.pyplot`, `import Nio` at its top (plotting / NetCDF output, not installed here) are stubbed;
  * `np.complex`, an alias NumPy removed in 1.24, is restored as the builtin `complex` it always was;
  * get_params() returns a dict while solve() reads attributes: it is wrapped in the reference's own Bunch
    (helpers/lib/bunch.py), as helpers/wrf/wind_compare.py's caller needs as well."""
import os
import sys
import types

import numpy as np

REF = "/root/reference/helpers"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "linear_theory_1d.npz")


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "Nio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "complex"):
        np.complex = complex
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "lib"))
    import ideal_linear
    from bunch import Bunch

    dx = 2000.0
    zlevels = [0.0, 250.0, 500.0, 1000.0, 2000.0, 4000.0]
    out = {"dx": dx, "zlevels": np.array(zlevels)}
    cases = [  # Nx (even: np.fft.fftshift == fftshift.f90 there), mountain height, half width, U, N^2
        (256, 1000.0, 10.0, 10.0, 3.6e-5),
        (256, 1000.0, 10.0, -8.0, 1.0e-4),
        (420, 1800.0, 30.0, 15.0, 3.0e-5),
        (128, 500.0, 3.1, 5.0, -1.0e-5),          # negative N^2: every mode evanescent
    ]
    for n, (Nx, hm, am, U, Ndsq) in enumerate(cases):
        x = np.arange(Nx, dtype=np.float64)
        zs = hm / (1.0 + ((x - Nx / 2.0) / am) ** 2)             # setup_experiment's mountain ("eqn from Trude")
        Fzs, params = ideal_linear.get_params(268.0, U, Ndsq, zs, -0.004)
        Pt, w, U3d, z3d = ideal_linear.solve(Fzs, U, dx, Bunch(**params), zlevels=zlevels)
        out[f"zs_{n}"] = zs
        out[f"Fzs_{n}"] = Fzs
        out[f"U_{n}"] = U
        out[f"Ndsq_{n}"] = Ndsq
        out[f"u_hat_{n}"] = U3d - U
        out[f"w_{n}"] = w
    out["ncases"] = len(cases)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.shape(v) for k, v in out.items() if k.endswith("_0")})


if __name__ == "__main__":
    main()
