"""Exact-arithmetic checks (no GPU) of the two floating-point identities the bit-exact pass-1 kernels rely on
(icar_b200/csrc/mpdata_staged.cu):

* div_shared: with y = RN(1/b) the Markstein sequence q0 = RN(a*y), r = a - b*q0 (exact in an FMA),
  q = RN(q0 + r*y) is the IEEE quotient RN(a/b);
* the reference's donor flux U+ * l + U- * r with U+- = (U +- |U|) / 2 (adv_mpdata.f90:41), one of whose two products
  is always an exact zero, equals the select form U * (U > 0 ? l : r) every kernel uses (donor() in common.cuh).

fp32 operations are emulated with rational arithmetic and one correct rounding per operation."""
from fractions import Fraction

import numpy as np
import pytest

F = np.float32


def rn(x: Fraction) -> np.float32:
    """Round a rational to the nearest float32, ties to even (normal range)."""
    if x == 0:
        return F(0.0)
    c = F(float(x))                                   # within one ulp of the answer (two roundings)
    cands = {c, np.nextafter(c, F(np.inf), dtype=F), np.nextafter(c, F(-np.inf), dtype=F)}
    best = sorted(cands, key=lambda f: (abs(Fraction(float(f)) - x), int(np.float32(f).view(np.uint32)) & 1))
    return F(best[0])


def fr(f) -> Fraction:
    return Fraction(float(f))


def fma(a, b, c) -> np.float32:
    return rn(fr(a) * fr(b) + fr(c))


@pytest.mark.parametrize("seed", [0, 1])
def test_markstein_quotient_with_a_correctly_rounded_reciprocal_is_the_ieee_quotient(seed):
    rng = np.random.default_rng(seed)
    n = 4000
    a = (rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 3, n)).astype(F)
    b = (rng.uniform(0.05, 1.0, n) * 10.0 ** rng.uniform(0, 3, n)).astype(F)      # jacobian, dz * jacobian
    a[: n // 20] = 0
    for x, d in zip(a, b):
        y = rn(1 / fr(d))                              # __frcp_rn
        q0 = rn(fr(x) * fr(y))
        r = fma(-d, q0, x)
        q = fma(r, y, q0)
        want = rn(fr(x) / fr(d))                       # a / b
        assert q == want, (x, d, q, want)


def test_donor_flux_in_the_reference_form_equals_the_select_form():
    rng = np.random.default_rng(5)
    n = 3000
    U = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0, n)).astype(F)
    l = (rng.uniform(0, 1, n) * 10.0 ** rng.uniform(-8, 2.5, n)).astype(F)
    r = (rng.uniform(0, 1, n) * 10.0 ** rng.uniform(-8, 2.5, n)).astype(F)
    U[:100] = 0
    l[100:200] = 0
    r[200:300] = 0
    for u, ql, qr in zip(U, l, r):
        up = rn((fr(u) + abs(fr(u))) * Fraction(1, 2))       # (U + |U|) * 0.5: both steps exact
        um = rn((fr(u) - abs(fr(u))) * Fraction(1, 2))
        got = fma(um, qr, rn(fr(up) * fr(ql)))
        want = rn(fr(u) * fr(ql if u > 0 else qr))
        assert got == want, (u, ql, qr, got, want)          # +0 == -0 here, as in the kernels' bitwise tests
