"""oracle/f90run (the Fortran-subset translator the oracle's parity rests on) held to the Fortran rules on small inline
programs whose results are known by hand: integer division, DO trip counts, sections with lower bounds, views as actual
arguments, scalars copied back, pointer assignment, (re)allocation on assignment, labelled DO / GOTO, optional and keyword
arguments, REAL rounding of literals and of mixed-kind expressions, constant folding, WHERE, complex kinds."""
import os

import numpy as np
import pytest

from oracle.f90run.runtime import FArr, FortranStop, lit4, unallocated
from oracle.f90run.translate import Module, TranslateError


def build(tmp_path, text, entries, **kw):
    p = os.path.join(tmp_path, "m.f90")
    with open(p, "w") as f:
        f.write(text)
    return Module(p, **kw).build(entries)


SRC = """
module m
    implicit none
    real, parameter :: third = 1 / 3.0          ! REAL division, rounded once to binary32
    real, parameter :: d0 = 200.e-6, two = 2.0
    real :: folded = d0**two                    ! constant ** constant: folded correctly rounded
    integer :: calls = 0
contains
    integer function idiv(a, b)
        integer, intent(in) :: a, b
        idiv = a / b                            ! truncates toward zero
    end function
    subroutine trip(lo, hi, step, n, last)
        integer, intent(in) :: lo, hi, step
        integer, intent(out) :: n, last
        integer :: i
        n = 0
        last = -99
        do i = lo, hi, step
            n = n + 1
            last = i
        enddo
    end subroutine
    subroutine sections(a, ims, ime, s)
        integer, intent(in) :: ims, ime
        real, intent(inout) :: a(ims:ime)
        real, intent(out) :: s
        a(ims+1:ime) = a(ims:ime-1) + a(ims+1:ime)      ! the right side is evaluated before anything is stored
        s = a(ime)
        call bump(a(ims+1:ims+2), s)                    ! a section is updated in place, a scalar copied back
    end subroutine
    subroutine bump(x, t)
        real, intent(inout) :: x(:), t
        x = x * 2
        t = t + x(1)                                    ! assumed shape: lower bound 1
        calls = calls + 1
    end subroutine
    subroutine pointers(out)
        real, intent(out) :: out(4)
        real, target :: a(0:3), b(5:8)
        real, pointer :: p(:)
        a = 1
        b = 2
        p => a
        p = p + 1                                       ! through the pointer: a = 2
        p => b
        p(5) = 9                                        ! the target's bounds
        out(1) = a(0); out(2) = a(3); out(3) = b(5); out(4) = b(8)
    end subroutine
    subroutine realloc(n, out)
        integer, intent(in) :: n
        real, intent(out) :: out
        real, allocatable :: t(:)
        real :: big(10)
        integer :: i
        do i = 1, 10
            big(i) = i
        enddo
        allocate(t(2:3))
        t = big(1:n)                                    ! reallocated to n elements, lower bound 1
        out = t(n) + size(t)
    end subroutine
    integer function find(v, x)
        real, intent(in) :: v(:), x
        integer :: i, n
        n = 0
        do 10 i = 1, size(v)
            n = i
            if (v(i) >= x) goto 20
10      continue
20      continue
        find = n
    end function
    real function opt(a, b, c)
        real, intent(in) :: a
        real, intent(in), optional :: b, c
        opt = a
        if (present(b)) opt = opt + 10 * b
        if (present(c)) opt = opt + 100 * c
    end function
    real function caller()
        caller = opt(1.0) + opt(1.0, c=2.0) + opt(1.0, 3.0)
    end function
    double precision function mixed(x, y)
        real, intent(in) :: x, y
        double precision :: d
        d = 1.d0
        mixed = x * y + d                               ! x*y is a REAL product, then promoted
    end function
    subroutine masked(a)
        real, intent(inout) :: a(:)
        where (a < 0) a = 0
    end subroutine
    subroutine cplx(z, w)
        complex(kind=8), intent(inout) :: z(:)
        complex, intent(out) :: w
        complex, parameter :: j = (0, 1)
        z = z * j
        w = j * j
    end subroutine
    subroutine halt(x)
        real, intent(in) :: x
        if (x < 0) stop
    end subroutine
end module
"""


@pytest.fixture(scope="module")
def m(tmp_path_factory):
    return build(str(tmp_path_factory.mktemp("f90")), SRC,
                 ["idiv", "trip", "sections", "pointers", "realloc", "find", "caller", "mixed", "masked", "cplx", "halt"])


def test_integer_division_truncates_toward_zero(m):
    assert [m.idiv(7, 2), m.idiv(-7, 2), m.idiv(7, -2), m.idiv(-7, -2), m.idiv(1, 3)] == [3, -3, -3, 3, 0]


def test_do_trip_counts(m):
    assert m.trip(1, 5, 1, 0, 0) == (5, 5)
    assert m.trip(5, 1, 1, 0, 0) == (0, -99)            # zero-trip loop
    assert m.trip(5, 1, -2, 0, 0) == (3, 1)
    assert m.trip(1, 6, 2, 0, 0) == (3, 5)


def test_sections_lower_bounds_views_and_copy_back(m):
    a = np.array([1, 2, 3, 4], np.float32)
    (s,) = m.sections(a, 5, 8, np.float32(0))
    # a(6:8) = a(5:7) + a(6:8) -> 1 3 5 7 ; s = 7 ; bump doubles a(6:7) -> 1 6 10 7 and adds a(6) = 6
    assert a.tolist() == [1, 6, 10, 7] and s == 13 and m.M.calls == 1
    with pytest.raises(IndexError):
        FArr(a, (5,))[9]                                # subscripts are checked against the declared bounds


def test_pointer_assignment_and_targets_bounds(m):
    out = np.zeros(4, np.float32)
    m.pointers(out)
    assert out.tolist() == [2, 2, 9, 2]


def test_reallocation_on_assignment(m):
    assert m.realloc(4, 0)[0] == 4 + 4 and m.realloc(7, 0)[0] == 7 + 7


def test_labelled_do_and_goto_out_of_the_loop(m):
    v = np.array([1, 4, 9, 16], np.float32)
    assert m.find(v, np.float32(5)) == 3 and m.find(v, np.float32(1)) == 1
    assert m.find(v, np.float32(99)) == 4               # the loop ran out: n = last i


def test_goto_that_skips_executable_statements_is_refused(tmp_path):
    src = SRC.replace("10      continue\n", "10      continue\n        n = -1\n")
    assert src != SRC
    # `n = -1` between the loop end and label 20 is executable: a GOTO over it is outside the subset -> refused loudly
    with pytest.raises(TranslateError):
        build(str(tmp_path), src, ["find"])


def test_optional_and_keyword_arguments(m):
    assert m.caller() == np.float32(1 + 201 + 31)


def test_real_arithmetic_rounds_once_per_operation(m):
    f = np.float32
    assert m.M.third == f(1) / f(3.0)
    assert m.M.folded == f(3.9999996e-08) and m.M.folded != f(4e-8)      # what powf(200e-6, 2) returns at run time
    x, y = f(1.0000001), f(1.0000001)
    assert m.mixed(x, y) == np.float64(f(x * y)) + 1.0                   # not the double product
    assert lit4("0.1") == f(0.1) and lit4("1e-10") == f(1e-10) and lit4("16777217") == f(16777216)


def test_where_and_complex_kinds(m):
    a = np.array([-1, 2, -3], np.float32)
    m.masked(a)
    assert a.tolist() == [0, 2, 0]
    z = np.array([1 + 2j], np.complex128)
    (w,) = m.cplx(z, 0)
    assert z[0] == complex(-2, 1) and w == complex(-1, 0) and isinstance(w, np.complexfloating)


def test_stop_raises(m):
    m.halt(np.float32(1))
    with pytest.raises(FortranStop):
        m.halt(np.float32(-1))


def test_unsupported_statements_are_errors_not_silence(tmp_path):
    bad = "module b\ncontains\nsubroutine s(a)\nreal :: a(:)\nforall (i=1:3) a(i) = i\nend subroutine\nend module\n"
    with pytest.raises(TranslateError):
        build(str(tmp_path), bad, ["s"])
    assert unallocated(np.float32, 2).a is None


def test_conditional_compilation_harness_methods_with_keywords_and_external_out_arguments(tmp_path):
    """#ifdef / #ifndef / #else / #endif resolve as in a build without -D switches; CALL obj%method(key = value) reaches a
    harness object with the keyword; an external subroutine declared as (function, [positions]) sets its scalar argument."""
    src = """
module m
contains
    subroutine s(obj, x)
        type(clock_t), intent(inout) :: obj
        double precision, intent(inout) :: x
#ifndef __INTEL_COMPILER
        call shrink(x)
#endif
#ifdef __INTEL_COMPILER
        x = -1
#else
        x = x + 1
#endif
        call obj%set(seconds=min(x, 120.0D0))
    end subroutine
end module
"""
    got = []
    obj = type("O", (), {"set": lambda self, seconds: got.append(seconds)})()
    mod = build(str(tmp_path), src, ["s"], externals={"shrink": (lambda x: x / 2, [0])})
    assert mod.s(obj, np.float64(10.0))[0] == 6.0 and got == [6.0]
    mod.s(obj, np.float64(1000.0))
    assert got[-1] == 120.0
    with pytest.raises(TranslateError, match="preprocessor"):
        build(str(tmp_path), "#if defined(X)\nmodule m\nend module\n#endif\n", [])
