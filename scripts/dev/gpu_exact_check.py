"""Development check (GPU box): the bit-exact fused corrector against the oracle, the packed division against `/`."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import icar_b200
from icar_b200 import ideal
from oracle import oracle as O

f = np.float32
def dom_for(nx, nz, ny, dx, **opt):
    return icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(dx=dx, **opt))

def division():
    d = dom_for(16, 4, 8, 4000.0)
    rng = np.random.default_rng(7)
    n = 1 << 24
    worst = 0
    for rep, (alo, ahi, blo, bhi) in enumerate([(-30, 8, -34, 10), (-60, 20, -50, 12), (-10, 10, -10, 10), (-1, 1, -1, 1)]):
        a = (rng.uniform(1, 2, n) * np.exp2(rng.integers(alo, ahi, n)) * rng.choice([-1, 1], n)).astype(f)
        b = (rng.uniform(1, 2, n) * np.exp2(rng.integers(blo, bhi, n))).astype(f)
        a[rng.random(n) < 0.05] = 0
        # quotients close to midpoints: a = b * (m + 0.5 ulp) for a random 24-bit m
        m = rng.uniform(1, 2, n // 4).astype(f)
        a[: n // 4] = (b[: n // 4].astype(np.float64) * (m.astype(np.float64) + 2.0 ** -24)).astype(f)
        p, q = d.selftest_division(a, b)
        host = a / b
        bad = int((p.view(np.uint32) != q.view(np.uint32)).sum() - ((p == 0) & (q == 0) & (p.view(np.uint32) != q.view(np.uint32))).sum())
        badh = int((q != host).sum())
        worst += bad
        print(f"division set {rep}: packed != device '/' in {bad} of {n}; device '/' != host '/' in {badh}", flush=True)
    d.close()
    return worst

def advect_bitwise(dims, steps=3):
    nx, nz, ny = dims
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True, dz_level=400.0 if nz >= 7 else 800.0)
    rng = np.random.default_rng(4)
    for k in ("qc", "qr", "qs"):
        st[k] = (rng.uniform(0, 1e-3, st[k].shape) * (rng.random(st[k].shape) < 0.5)).astype(f)
    st["w"] = (st["w"] + rng.uniform(-0.3, 0.3, st["w"].shape)).astype(f)
    dt = float(st["dt"]) * 0.8
    dom = dom_for(nx, nz, ny, st["dx"])
    dom.upload_state(st)
    U, V, W = O.wind_normalise(st["u"], st["v"], st["w"], st["jacobian_u"], st["jacobian_v"], st["jacobian_w"], None, f(dt), st["dx"])
    tot = 0
    for step in range(steps):
        dom.advect(dt)
        for k in ("th", "qv", "qc", "qr", "qs"):
            O.advect3d(st[k], None, st["jacobian"], st["advection_dz"], U, V, W, 2, True)
            got = dom.download(k)
            nbad = int((got != st[k]).sum())
            tot += nbad
            if nbad:
                err = np.abs(got.astype(np.float64) - st[k]).max() / np.abs(st[k]).max()
                print(f"  {dims} step {step} {k}: {nbad} cells differ, rel {err:.2e}", flush=True)
    print(f"advect {dims}: {tot} cells differ over {steps} calls x 5 scalars", flush=True)
    dom.close()
    return tot

def config1(steps=20):
    st = ideal.make_ideal_case(nx=100, nz=20, ny=100)
    ref = ideal.copy_state(st)
    dt = float(st["dt"])
    dom = dom_for(100, 20, 100, st["dx"])
    dom.upload_state(st)
    for _ in range(steps):
        dom.physics_step(dt)
        O.step_simple(ref, f(dt), f(dt), 2, True)
    tot = 0
    for k in ("th", "qv", "qc", "qr", "qs"):
        got = dom.download(k)
        nbad = int((got != ref[k]).sum())
        err = np.abs(got.astype(np.float64) - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-30)
        print(f"config1 {steps} steps {k}: {nbad} cells differ, rel-to-max {err:.2e}", flush=True)
        tot += nbad
    dom.close()
    return tot

if __name__ == "__main__":
    bad = division()
    for dims in [(72, 12, 40), (100, 20, 100), (45, 7, 23), (61, 33, 19), (30, 64, 12), (40, 50, 14), (33, 60, 11), (36, 48, 9), (130, 40, 37)]:
        bad += advect_bitwise(dims)
    bad += config1()
    print("TOTAL DIFFERING:", bad)
