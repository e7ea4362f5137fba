"""TEST INFRASTRUCTURE -- run-time support for Fortran source executed through oracle/f90run/translate.py.

Semantics that matter for parity and where they come from:
  * REAL is IEEE binary32: every value is a numpy.float32 scalar or a float32 array, so each +, -, *, / rounds once
    to binary32 exactly like the compiled Fortran (no -ffast-math in the reference's makefile, src/makefile:159);
    DOUBLE PRECISION is numpy.float64; INTEGER is a Python int (integer / integer truncates toward zero).
  * exp / log / x**y on REAL scalars call glibc's expf / logf / powf through ctypes: the functions a gfortran build
    of the reference links against on this platform.  x**n with an INTEGER n multiplies (libgcc __powisf2 order).
  * arrays keep their declared lower bounds (FArr); sections are numpy views, so a section passed to a routine is
    updated in place like a Fortran actual argument.
"""
import ctypes
import ctypes.util
import math

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("expf", "logf", "log10f", "sqrtf", "atanf", "tanhf", "sinf", "cosf"):
    getattr(_libm, _n).argtypes = [ctypes.c_float]
    getattr(_libm, _n).restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.powf.restype = ctypes.c_float

f32 = np.float32
f64 = np.float64


class FortranStop(Exception):
    """`stop` / `error stop` reached."""


class Sl:
    """A subscript triplet lo:hi:step in Fortran index space (None = the declared bound)."""
    __slots__ = ("lo", "hi", "step")

    def __init__(self, lo=None, hi=None, step=None):
        self.lo, self.hi, self.step = lo, hi, step


ALL = Sl()


class FArr:
    """An array with Fortran lower bounds.  `a` is the numpy array (None while an allocatable is unallocated)."""
    __slots__ = ("a", "lb", "dtype", "allocatable")

    def __init__(self, a, lb, dtype=None, allocatable=False):
        self.a = a
        self.lb = tuple(lb)
        self.dtype = dtype if dtype is not None else (a.dtype if a is not None else f32)
        self.allocatable = allocatable

    def _ix(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) != self.a.ndim:
            raise IndexError(f"rank {self.a.ndim} array indexed with {len(key)} subscripts")
        out = []
        for k, l, n in zip(key, self.lb, self.a.shape):
            if isinstance(k, Sl):
                lo = l if k.lo is None else int(k.lo)
                hi = l + n - 1 if k.hi is None else int(k.hi)
                st = 1 if k.step is None else int(k.step)
                if st != 1:
                    raise NotImplementedError("strided sections")
                if hi < lo:
                    out.append(slice(0, 0))
                    continue
                if lo < l or hi > l + n - 1:
                    raise IndexError(f"section {lo}:{hi} outside the bounds {l}:{l + n - 1}")
                out.append(slice(lo - l, hi - l + 1))
            else:
                i = int(k) - l
                if i < 0 or i >= n:
                    raise IndexError(f"subscript {int(k)} outside the bounds {l}:{l + n - 1}")
                out.append(i)
        return tuple(out)

    def __getitem__(self, key):
        return self.a[self._ix(key)]

    def __setitem__(self, key, val):
        self.a[self._ix(key)] = val

    def assign(self, val):
        """whole-array assignment `x = expr` (with reallocation on assignment for an allocatable)."""
        val = A(val)
        if self.allocatable and isinstance(val, np.ndarray) and (self.a is None or self.a.shape != val.shape):
            # Fortran 2003: the left side is (re)allocated with the shape of the right side, lower bounds 1 unless the
            # right side is a whole array (whose bounds are kept); sections have lower bound 1
            self.a = np.array(val, dtype=self.dtype, order="F")
            self.lb = (1,) * val.ndim
            return
        if self.a is None:
            raise ValueError("assignment to an unallocated array")
        self.a[...] = val


def A(x):
    """the numpy array of a whole-array reference, anything else unchanged"""
    return x.a if isinstance(x, FArr) else x


def alloc(dtype, bounds, allocatable=False):
    shape = tuple(max(int(hi) - int(lo) + 1, 0) for lo, hi in bounds)
    return FArr(np.zeros(shape, dtype=dtype, order="F"), [int(lo) for lo, _ in bounds], dtype, allocatable)


def alloc_into(obj, dtype, bounds):
    new = alloc(dtype, bounds, True)
    if isinstance(obj, FArr):
        obj.a, obj.lb, obj.dtype, obj.allocatable = new.a, new.lb, new.dtype, True
        return obj
    return new


def unallocated(dtype, rank):
    return FArr(None, (1,) * rank, dtype, True)


def wrap(x, bounds, dtype=f32):
    """A dummy array argument: the actual argument's storage under the dummy's declared bounds.
    bounds: [(lo, hi) | (lo, None) for assumed shape].  Shapes must agree (no sequence association)."""
    a = A(x)
    if not isinstance(a, np.ndarray):
        raise TypeError(f"array dummy argument got {type(x).__name__}")
    if a.ndim != len(bounds):
        raise ValueError(f"rank mismatch: actual {a.ndim}, dummy {len(bounds)}")
    for n, (lo, hi) in zip(a.shape, bounds):
        if hi is not None and max(int(hi) - int(lo) + 1, 0) != n:
            raise ValueError(f"shape mismatch: actual extent {n}, dummy {lo}:{hi}")
    if a.dtype != dtype:
        raise TypeError(f"dummy argument of type {np.dtype(dtype)} got {a.dtype}")
    return FArr(a, [int(lo) for lo, _ in bounds], dtype)


def is_allocated(x):
    return isinstance(x, np.ndarray) or (isinstance(x, FArr) and x.a is not None)


# ------------------------------------------------------------------ scalar conversions (assignment to a typed variable)
def r4(x):
    if isinstance(x, np.ndarray):                 # an elemental procedure applied to an array
        return x if x.dtype == np.float32 else x.astype(np.float32)
    return x if type(x) is np.float32 else np.float32(x)


def r8(x):
    return x if type(x) is np.float64 else np.float64(x)


def i4(x):
    if isinstance(x, (int, np.integer)):
        return int(x)
    return int(math.trunc(float(x)))


def lit4(text):
    """A REAL literal: the decimal string rounded ONCE to binary32 (via exact rational arithmetic)."""
    from fractions import Fraction
    t = text.lower().replace("d", "e")
    exact = Fraction(t)
    c = np.float32(float(t))
    best = c
    for cand in (np.nextafter(c, np.float32(-np.inf)), np.nextafter(c, np.float32(np.inf))):
        if abs(Fraction(float(cand)) - exact) < abs(Fraction(float(best)) - exact):
            best = cand
    return np.float32(best)


# ------------------------------------------------------------------ operators that need Fortran's rules
def div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, bool):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


def _powi(x, n):
    """x**n, INTEGER n, by repeated squaring in x's own precision (the order of libgcc's __powi?f2)."""
    m = -n if n < 0 else n
    one = x * 0 + 1
    y = x if (m & 1) else one
    m >>= 1
    while m:
        x = x * x
        if m & 1:
            y = y * x
        m >>= 1
    return (one / y) if n < 0 else y


def power(x, y):
    if isinstance(y, (int, np.integer)):
        if isinstance(x, (int, np.integer)):
            return int(x) ** int(y)
        return _powi(x, int(y))
    if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
        xx, yy = np.broadcast_arrays(x, y)
        if xx.dtype == np.float32 and np.result_type(xx, yy) == np.float32:
            out = np.empty(xx.shape, np.float32)
            fo, fx, fy = out.reshape(-1), xx.reshape(-1), yy.astype(np.float32).reshape(-1)
            for n in range(fx.size):
                fo[n] = _libm.powf(float(fx[n]), float(fy[n]))
            return out
        return np.power(xx.astype(np.float64), yy.astype(np.float64))
    if type(x) is np.float64 or type(y) is np.float64 or isinstance(x, float) or isinstance(y, float):
        return np.float64(math.pow(float(x), float(y)))
    return np.float32(_libm.powf(float(x), float(y)))


def cpow(x, y):
    """constant ** constant as the compiler folds it: the correctly rounded power in the type of the result"""
    import decimal
    from fractions import Fraction
    if isinstance(y, (int, np.integer)):
        return power(x, y)
    dt = np.float64 if (type(x) is np.float64 or type(y) is np.float64) else np.float32
    with decimal.localcontext() as ctx:
        ctx.prec = 80
        exact = Fraction(decimal.Decimal(float(x)) ** decimal.Decimal(float(y)))
    c = dt(float(exact)) if dt is np.float64 else np.float32(float(exact))
    best = c
    for cand in (np.nextafter(c, dt(-np.inf)), np.nextafter(c, dt(np.inf))):
        if abs(Fraction(float(cand)) - exact) < abs(Fraction(float(best)) - exact):
            best = cand
    return dt(best)


def _unary_libm(name, dname):
    cf = getattr(_libm, name)
    df = getattr(math, dname)

    def f(x):
        if np.iscomplexobj(x):
            return getattr(np, dname)(x)
        if isinstance(x, np.ndarray):
            if x.dtype == np.float32:
                out = np.empty(x.shape, np.float32)
                fo, fx = out.reshape(-1), x.reshape(-1)
                for n in range(fx.size):
                    fo[n] = cf(float(fx[n]))
                return out
            return np.vectorize(df, otypes=[np.float64])(x)
        if type(x) is np.float32:
            return np.float32(cf(float(x)))
        return np.float64(df(float(x)))
    return f


f_exp = _unary_libm("expf", "exp")
f_log = _unary_libm("logf", "log")
f_log10 = _unary_libm("log10f", "log10")


def f_sqrt(x):
    return np.sqrt(x)           # REAL / DOUBLE: correctly rounded in every implementation; COMPLEX: numpy's csqrt


def f_abs(x):
    return abs(x) if isinstance(x, (int, float)) else np.abs(x)


def _fold(op, pyop, args):
    args = [A(a) for a in args]
    if all(isinstance(a, (int, np.integer)) for a in args):
        return pyop(int(a) for a in args)
    r = args[0]
    for a in args[1:]:
        r = op(r, a)
    return r


def f_max(*args):
    return _fold(np.maximum, max, args)


def f_min(*args):
    return _fold(np.minimum, min, args)


def f_maxval(x):
    return A(x).max()


def f_minval(x):
    return A(x).min()


def f_sum(x):
    a = A(x)
    return np.cumsum(a.reshape(-1, order="F"))[-1] if a.size else a.dtype.type(0)    # sequential, like the loop


def f_size(x, dim=None):
    a = A(x)
    return int(a.size) if dim is None else int(a.shape[int(dim) - 1])


def f_lbound(x, dim):
    return int(x.lb[int(dim) - 1]) if isinstance(x, FArr) else 1


def f_ubound(x, dim):
    return f_lbound(x, dim) + int(A(x).shape[int(dim) - 1]) - 1


def f_ceiling(x):
    return int(math.ceil(float(x)))


def f_floor(x):
    return int(math.floor(float(x)))


def f_nint(x):
    v = float(x)
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def f_int(x):
    return i4(x)


def f_real(x, kind=None):
    a = A(x)
    if np.iscomplexobj(a):                        # REAL(z) has the kind of z
        return a.real.copy() if isinstance(a, np.ndarray) else a.real
    if isinstance(a, np.ndarray):
        return a.astype(np.float32)
    return np.float32(a)


def c64(x):
    a = A(x)
    return a.astype(np.complex64) if isinstance(a, np.ndarray) else np.complex64(a)


def c128(x):
    a = A(x)
    return a.astype(np.complex128) if isinstance(a, np.ndarray) else np.complex128(a)


def where_assign(obj, mask, val):
    a = obj.a if isinstance(obj, FArr) else obj
    mask = A(mask)
    val = A(val)
    a[mask] = val[mask] if isinstance(val, np.ndarray) else val


def alloc_like(src, dtype):
    a = A(src)
    return FArr(np.array(a, dtype=dtype, order="F"), (1,) * a.ndim, dtype, True)


def c_f_pointer(buffer, shape):
    """c_f_pointer(cptr, fptr, shape): the Fortran pointer over the C buffer (here: the harness's numpy buffer)"""
    shape = tuple(int(n) for n in A(shape))
    n = int(np.prod(shape))
    return FArr(buffer[:n].reshape(shape, order="F"), (1,) * len(shape), buffer.dtype)


def f_dble(x):
    a = A(x)
    return a.astype(np.float64) if isinstance(a, np.ndarray) else np.float64(a)


def f_mod(a, p):
    if isinstance(a, (int, np.integer)) and isinstance(p, (int, np.integer)):
        return int(math.fmod(int(a), int(p)))
    return np.fmod(a, p)


def f_sign(a, b):
    return np.copysign(np.abs(a), b) if not isinstance(a, (int, np.integer)) else (abs(int(a)) if b >= 0 else -abs(int(a)))


def f_merge(t, f, mask):
    return np.where(mask, t, f) if isinstance(mask, np.ndarray) else (t if mask else f)


class Obj:
    """a local variable of derived type: components appear when they are assigned"""


class ObjArray(dict):
    """a local array of derived type, 1-based: a(i) is an Obj, a%comp the array of that component over the elements"""

    def __missing__(self, key):
        self[key] = Obj()
        return self[key]

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return np.array([getattr(self[k], name) for k in sorted(self)])


def arr(items, dtype):
    return np.array([A(x) for x in items], dtype=dtype, order="F")


f_tanh = _unary_libm("tanhf", "tanh")
f_atan = _unary_libm("atanf", "atan")
f_sin = _unary_libm("sinf", "sin")
f_cos = _unary_libm("cosf", "cos")


def _as_double(fn):
    return lambda x: np.float64(fn(float(x)))


def f_this_image():
    return 1


def f_num_images():
    return 1


def do_range(lo, hi, step=1):
    lo, hi, step = int(lo), int(hi), int(step)
    return range(lo, hi + (1 if step > 0 else -1), step)


NAMES = {
    "_FArr": FArr, "_Sl": Sl, "_ALL": ALL, "_A": A, "_alloc": alloc, "_alloc_into": alloc_into, "_unallocated": unallocated, "_wrap": wrap,
    "_is_allocated": is_allocated, "_r4": r4, "_r8": r8, "_i4": i4, "_div": div, "_pow": power, "_cpow": cpow, "_do": do_range,
    "_f32": np.float32, "_f64": np.float64, "_i32": np.int32, "_np": np, "_Stop": FortranStop,
    "_fn_exp": f_exp, "_fn_log": f_log, "_fn_alog": f_log, "_fn_log10": f_log10, "_fn_alog10": f_log10,
    "_fn_sqrt": f_sqrt, "_fn_abs": f_abs,
    "_fn_max": f_max, "_fn_min": f_min, "_fn_amax1": f_max, "_fn_amin1": f_min, "_fn_maxval": f_maxval,
    "_fn_minval": f_minval, "_fn_sum": f_sum, "_fn_size": f_size, "_fn_ceiling": f_ceiling, "_fn_floor": f_floor,
    "_fn_nint": f_nint, "_fn_int": f_int, "_fn_real": f_real, "_fn_float": f_real, "_fn_dble": f_dble, "_fn_mod": f_mod,
    "_fn_sign": f_sign, "_fn_merge": f_merge, "_fn_this_image": f_this_image, "_fn_num_images": f_num_images,
    "_fn_allocated": is_allocated, "_fn_associated": lambda x, *a: x is not None, "_fn_present": lambda x: x is not None,
    "_c64": c64, "_c128": c128, "_where_assign": where_assign, "_alloc_like": alloc_like, "_c_f_pointer": c_f_pointer,
    "_fn_cmplx": lambda a, b=0, kind=None: np.complex64(complex(a, b)), "_fn_aimag": lambda z: A(z).imag,
    "_fn_conjg": lambda z: np.conj(A(z)),
    "_Obj": Obj, "_ObjArray": ObjArray, "_fn_product": lambda x: int(np.prod(A(x))), "_fn_shape": lambda x: np.array(A(x).shape),
    "_arr": arr, "_fn_lbound": f_lbound, "_fn_ubound": f_ubound, "_fn_dmax1": f_max, "_fn_dmin1": f_min, "_fn_dlog": _as_double(math.log), "_fn_dlog10": _as_double(math.log10),
    "_fn_dexp": _as_double(math.exp), "_fn_dsqrt": _as_double(math.sqrt), "_fn_dabs": f_abs, "_fn_dfloat": f_dble,
    "_fn_sngl": f_real, "_fn_tanh": f_tanh, "_fn_atan": f_atan, "_fn_sin": f_sin, "_fn_cos": f_cos, "_fn_iabs": f_abs, "_fn_trim": lambda x: x,
}
