"""TEST INFRASTRUCTURE -- a mechanical translator from a subset of Fortran 90 to Python/NumPy.

Purpose: the reference (NCAR/icar) is Fortran and no Fortran compiler exists in this image, so its own source files
cannot be compiled here.  This module EXECUTES them instead: it reads a module source file where it lies (never copied
into this repository), translates every statement one to one -- same operations, same order, same types -- and the
result runs on oracle/f90run/runtime.py, whose REAL arithmetic is IEEE binary32 with glibc's expf/powf.  The outputs pin
the hand-written restatements (oracle/*.hpp, oracle/np_restatement.py) to the reference's own text.

The subset is what src/physics/adv_mpdata.f90 (+ the included adv_mpdata_FCT_core.f90) and src/physics/mp_simple.f90 use:
modules with parameters / module variables / contained subroutines and functions; INTEGER, REAL, DOUBLE PRECISION,
LOGICAL; explicit-shape, assumed-shape, automatic, allocatable and pointer arrays with arbitrary lower bounds; array
sections and whole-array expressions; IF / ELSE IF / DO / DO WHILE / EXIT / CYCLE / CALL / ALLOCATE / INCLUDE /
ASSOCIATE; derived-type components (read through Python attributes).  Anything else raises TranslateError: nothing
is skipped silently inside a routine that is asked for (PRINT is a no-op, STOP raises).
"""
import os
import re
import types

from . import runtime as R


class TranslateError(Exception):
    pass


PY_KEYWORDS = {"in", "is", "lambda", "from", "pass", "del", "not", "and", "or", "if", "else", "elif", "for", "while", "def",
               "class", "return", "global", "import", "as", "with", "try", "except", "raise", "yield", "assert", "None",
               "True", "False", "print", "exec", "nonlocal", "finally", "break", "continue", "async", "await"}


def py_name(n):
    return n + "_" if n in PY_KEYWORDS else n


# ============================================================================================ source -> statements
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return "".join(out)


def read_statements(path, _depth=0, defined=()):
    """[(file, first line number, statement text)]: comments stripped, continuations joined, INCLUDEs inlined, and the
    conditional-compilation lines of the C preprocessor (#ifdef / #ifndef / #else / #endif on a macro name) resolved with
    the macros in `defined` -- none by default, i.e. a build with neither __INTEL_COMPILER nor any -D switch."""
    if _depth > 5:
        raise TranslateError("include depth")
    stmts = []
    with open(path) as f:
        lines = f.read().split("\n")
    cur, cur_no = "", 0
    active = []                                            # one entry per open #if: is its current branch taken?
    for no, raw in enumerate(lines, 1):
        if raw.lstrip().startswith("#"):
            d = raw.strip()[1:].split()
            if d and d[0] in ("ifdef", "ifndef") and len(d) == 2:
                active.append((d[1] in defined) == (d[0] == "ifdef"))
            elif d and d[0] == "else" and active:
                active[-1] = not active[-1]
            elif d and d[0] == "endif" and active:
                active.pop()
            elif d and d[0] in ("define", "undef", "include"):
                if all(active):
                    raise TranslateError(f"{os.path.basename(path)}:{no}: preprocessor directive outside the subset: {raw.strip()}")
            else:
                raise TranslateError(f"{os.path.basename(path)}:{no}: preprocessor directive outside the subset: {raw.strip()}")
            continue
        if not all(active):
            continue
        line = _strip_comment(raw)
        if not line.strip():
            continue
        text = line.strip()
        if cur:
            if text.startswith("&"):
                text = text[1:].lstrip()
            cur += " " + text
        else:
            cur, cur_no = text, no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        for piece in _split_semicolons(cur):
            piece = _lower_outside_strings(piece.strip())
            if not piece:
                continue
            piece = _array_constructors(piece)
            m = re.match(r"include\s+['\"]([^'\"]+)['\"]$", piece)
            if m:
                inc = os.path.join(os.path.dirname(path), m.group(1))
                stmts.extend(read_statements(inc, _depth + 1, defined))
            else:
                stmts.append((os.path.basename(path), cur_no, piece))
        cur = ""
    return stmts


def _array_constructors(s):
    """(/ a, b /) -> [ a, b ] outside strings"""
    if "(/" not in s:
        return s
    out, q, k = [], None, 0
    while k < len(s):
        ch = s[k]
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif s[k:k + 2] == "(/":
            out.append("[")
            k += 1
        elif s[k:k + 2] == "/)":
            out.append("]")
            k += 1
        else:
            out.append(ch)
        k += 1
    return "".join(out)


def split_label(t):
    """'141 continue' -> ('141', 'continue'); 'name: do i = ...' -> construct name dropped"""
    m = re.match(r"^(\d+)\s+(.*)$", t)
    label = None
    if m:
        label, t = m.group(1), m.group(2)
    m = re.match(r"^(\w+)\s*:\s*((?:do|if)\b.*)$", t)
    if m and m.group(1) not in ("else",):
        t = m.group(2)
    return label, t


def _split_semicolons(s):
    out, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


# ============================================================================================ tokens
TOKEN_RE = re.compile(r"""
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<num>(?:\d+\.(?![a-z]+\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
  | (?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[a-z]\w*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|[-+*/<>=(),:%\[\]])
  | (?P<ws>\s+)
""", re.X)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise TranslateError(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append((kind, m.group()))
    toks.append(("end", ""))
    return toks


REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
       ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

INTRINSICS = {"product", "shape", "cmplx", "aimag", "conjg", "lbound", "ubound", "abs", "max", "min", "amax1", "amin1", "dmax1", "dmin1", "maxval", "minval", "sum", "size", "exp", "log",
              "alog", "log10", "alog10", "sqrt", "ceiling", "floor", "nint", "int", "real", "float", "dble", "mod", "sign",
              "merge", "this_image", "num_images", "allocated", "associated", "present", "dlog", "dlog10", "dexp", "dsqrt",
              "dabs", "dfloat", "sngl", "tanh", "atan", "sin", "cos", "iabs", "trim", "gamma"}
INT_RESULT = {"lbound", "ubound", "size", "ceiling", "floor", "nint", "int", "this_image", "num_images", "iabs"}
REAL_RESULT = {"real", "float", "sngl", "alog", "alog10", "amax1", "amin1"}
DOUBLE_RESULT = {"dble", "dfloat", "dlog", "dlog10", "dexp", "dsqrt", "dabs", "dmax1", "dmin1"}


class Var:
    def __init__(self, name, kind, dims=None, intent=None, parameter=False, allocatable=False, pointer=False,
                 init=None, dummy=False, module=False, derived=False, optional=False):
        self.optional = optional
        self.name, self.kind, self.dims = name, kind, dims
        self.intent, self.parameter, self.allocatable, self.pointer = intent, parameter, allocatable, pointer
        self.init, self.dummy, self.module, self.derived = init, dummy, module, derived

    @property
    def is_array(self):
        return self.dims is not None


DTYPE = {"r": "_f32", "d": "_f64", "i": "_i32", "l": "bool", "c": "object", "?": "_f32", "x": "_np.complex64", "z": "_np.complex128"}


class Scope:
    """Names visible in one routine: its own declarations over the module's."""

    def __init__(self, module, routine=None):
        self.module, self.routine = module, routine
        self.locals = {}
        self.aliases = {}          # associate names -> python expression

    def lookup(self, name):
        if name in self.locals:
            return self.locals[name]
        if name in self.module.vars:
            return self.module.vars[name]
        if name in self.module.names:                # an entity of another module, supplied by the harness
            v = self.module.names[name]
            import numpy as _np
            kind = "r" if isinstance(v, _np.float32) else "d" if isinstance(v, (_np.float64, float)) else \
                   "i" if isinstance(v, int) else None
            if kind:
                return Var(name, kind, parameter=True)
        return None

    def ref(self, name):
        """python expression of a variable reference"""
        if name in self.aliases:
            return self.aliases[name]
        if name in self.locals:
            return py_name(name)
        if name in self.module.vars:
            return "_M." + py_name(name)
        if name in self.module.names:
            return py_name(name)
        raise TranslateError(f"undeclared name {name!r} in {self.routine}")


class ExprParser:
    """Recursive descent over Fortran's operator precedence; emits a python expression and a kind tag
    ('i' integer, 'r' real, 'd' double, 'l' logical, '?' unknown / derived component)."""

    def __init__(self, toks, scope, consts):
        self.t, self.p, self.scope, self.consts = toks, 0, scope, consts
        self.constant = set()      # codes of compile-time constant subexpressions (literals, PARAMETERs and their arithmetic)

    def peek(self):
        return self.t[self.p]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def accept(self, kind, val=None):
        k, v = self.t[self.p]
        if k == kind and (val is None or v == val):
            self.p += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.accept(kind, val):
            raise TranslateError(f"expected {val or kind}, got {self.t[self.p]} in {' '.join(v for _, v in self.t)}")

    # ---- precedence levels
    def expr(self):
        return self.or_()

    def or_(self):
        c, k = self.and_()
        while self.peek() == ("dot", ".or."):
            self.next()
            c2, _ = self.and_()
            c, k = f"({c} or {c2})", "l"
        return c, k

    def and_(self):
        c, k = self.not_()
        while self.peek() == ("dot", ".and."):
            self.next()
            c2, _ = self.not_()
            c, k = f"({c} and {c2})", "l"
        return c, k

    def not_(self):
        if self.peek() == ("dot", ".not."):
            self.next()
            c, _ = self.not_()
            return f"(not {c})", "l"
        return self.rel()

    def rel(self):
        c, k = self.add()
        kind, v = self.peek()
        if (kind == "op" and v in REL) or (kind == "dot" and v in REL):
            self.next()
            c2, _ = self.add()
            return f"({c} {REL[v]} {c2})", "l"
        return c, k

    @staticmethod
    def _join(k1, k2):
        if "?" in (k1, k2):
            return "?"
        if "z" in (k1, k2) or ("x" in (k1, k2) and "d" in (k1, k2)):
            return "z"
        if "x" in (k1, k2):
            return "x"
        if "d" in (k1, k2):
            return "d"
        if "r" in (k1, k2):
            return "r"
        return "i"

    def add(self):
        kind, v = self.peek()
        if kind == "op" and v in "+-":
            self.next()
            c, k = self.mul()
            c = f"(-{c})" if v == "-" else c
        else:
            c, k = self.mul()
        while True:
            kind, v = self.peek()
            if kind == "op" and v in ("+", "-"):
                self.next()
                c2, k2 = self.mul()
                both = c in self.constant and c2 in self.constant
                c, k = f"({c} {v} {c2})", self._join(k, k2)
                if both:
                    self.constant.add(c)
            else:
                return c, k

    def mul(self):
        c, k = self.pow_()
        while True:
            kind, v = self.peek()
            if kind == "op" and v == "*":
                self.next()
                c2, k2 = self.pow_()
                both = c in self.constant and c2 in self.constant
                c, k = f"({c} * {c2})", self._join(k, k2)
                if both:
                    self.constant.add(c)
            elif kind == "op" and v == "/":
                self.next()
                c2, k2 = self.pow_()
                kk = self._join(k, k2)
                both = c in self.constant and c2 in self.constant
                # integer / integer truncates; decided at run time unless an operand is known to be real
                c = f"({c} / {c2})" if kk in ("r", "d") else f"_div({c}, {c2})"
                k = kk
                if both:
                    self.constant.add(c)
            else:
                return c, k

    def pow_(self):
        c, k = self.primary()
        if self.peek() == ("op", "**"):
            self.next()
            # right associative; a unary minus in the exponent is allowed
            kind, v = self.peek()
            if kind == "op" and v in "+-":
                self.next()
                c2, k2 = self.pow_()
                c2 = f"(-{c2})" if v == "-" else c2
            else:
                c2, k2 = self.pow_()
            # constant ** constant is folded by the compiler (gfortran: MPFR, correctly rounded); anything else is
            # powf / pow / repeated multiplication at run time
            fn = "_cpow" if (c in self.constant and c2 in self.constant and k2 != "i") else "_pow"
            both = c in self.constant and c2 in self.constant
            c = f"{fn}({c}, {c2})"
            if both:
                self.constant.add(c)
            return c, (k if k2 == "i" else self._join(k, k2))
        return c, k

    def literal(self, text):
        t = text
        kindsuffix = None
        if "_" in t:
            t, kindsuffix = t.split("_", 1)
        if re.fullmatch(r"\d+", t):
            c = t.lstrip("0") or "0"
            self.constant.add(c)
            return c, "i"
        if "d" in t or kindsuffix in ("dp", "8", "real64", "r8"):
            c = f"_f64({float(t.replace('d', 'e'))!r})"
            self.constant.add(c)
            return c, "d"
        idx = self.consts.setdefault(t, len(self.consts))
        self.constant.add(f"_C[{idx}]")
        return f"_C[{idx}]", "r"

    def args(self):
        """after '(' : list of ('expr', code, kind) | ('sec', lo, hi, step) up to ')'"""
        out = []
        if self.accept("op", ")"):
            return out
        while True:
            out.append(self.arg())
            if self.accept("op", ","):
                continue
            self.expect("op", ")")
            return out

    def arg(self):
        lo = hi = st = None
        if self.peek()[0] == "name" and self.t[self.p + 1] == ("op", "="):
            kw = self.next()[1]
            self.next()
            c, k = self.expr()
            return ("kw", c, k, kw)
        if self.peek() != ("op", ":"):
            # keyword argument name=expr (only for intrinsics like real(x, kind=...)): not needed in the subset
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return ("expr", lo[0], lo[1])
        self.expect("op", ":")
        if self.peek()[1] not in (",", ")", ":"):
            hi = self.expr()
        if self.accept("op", ":"):
            st = self.expr()
        return ("sec", lo[0] if lo else None, hi[0] if hi else None, st[0] if st else None)

    @staticmethod
    def subscript(args):
        parts = []
        for a in args:
            if a[0] == "expr":
                parts.append(a[1])
            elif a[1] is None and a[2] is None and a[3] is None:
                parts.append("_ALL")
            else:
                parts.append(f"_Sl({a[1]}, {a[2]}, {a[3]})")
        return ", ".join(parts)

    def primary(self):
        kind, v = self.next()
        if kind == "num":
            return self.literal(v)
        if kind == "dot" and v in (".true.", ".false."):
            return ("True" if v == ".true." else "False"), "l"
        if kind == "str":
            return repr(v[1:-1]), "?"
        if kind == "op" and v == "(":
            c, k = self.expr()
            if self.accept("op", ","):              # a complex literal (re, im)
                c2, k2 = self.expr()
                self.expect("op", ")")
                kk = "z" if "d" in (k, k2) else "x"
                return f"{DTYPE[kk]}(complex({c}, {c2}))", kk
            self.expect("op", ")")
            if c in self.constant:
                self.constant.add(f"({c})")
            return f"({c})", k
        if kind == "op" and v == "[":
            items, k = [], None
            while True:
                c, kk = self.expr()
                items.append(c)
                k = kk if k is None else self._join(k, kk)
                if self.accept("op", ","):
                    continue
                self.expect("op", "]")
                break
            return f"_arr([{', '.join(items)}], {DTYPE.get(k, '_f32')})", k
        if kind != "name":
            raise TranslateError(f"unexpected token {v!r} in {' '.join(x for _, x in self.t)}")
        name = v
        # ---- derived-type component chain: read through python attributes
        if self.peek() == ("op", "%"):
            code = self.scope.ref(name) if (name in self.scope.locals or name in self.scope.module.vars or
                                             name in self.scope.aliases) else py_name(name)
            while self.accept("op", "%"):
                _, comp = self.next()
                code += "." + py_name(comp)
                if self.peek() == ("op", "("):
                    self.next()
                    a = self.args()
                    if not a:
                        code += "()"                # a type-bound function without arguments, e.g. model_time%seconds()
                    else:
                        code += f"[{self.subscript(a)}]"
                    if self.peek() != ("op", "%"):
                        return code, "?"
            return f"_A({code})", "?"
        var = self.scope.lookup(name) if name not in self.scope.aliases else None
        if name in self.scope.aliases:
            code = self.scope.aliases[name]
            if self.accept("op", "("):
                a = self.args()
                return f"{code}[{self.subscript(a)}]", "?"
            return f"_A({code})", "?"
        if self.peek() == ("op", "("):
            self.next()
            if var is not None and var.is_array:
                a = self.args()
                code = f"{self.scope.ref(name)}[{self.subscript(a)}]"
                if var.derived and self.peek() == ("op", "%"):
                    while self.accept("op", "%"):
                        code += "." + py_name(self.next()[1])
                        if self.peek() == ("op", "("):
                            self.next()
                            aa = self.args()
                            code += "()" if not aa else f"[{self.subscript(aa)}]"
                    return f"_A({code})", "?"
                return code, var.kind
            if var is not None and not var.is_array and self.scope.routine and name == self.scope.routine.result:
                pass        # fall through: a recursive call is outside the subset, handled as a function below
            if name in ("lbound", "ubound", "size") and self.peek()[0] == "name" and \
                    self.scope.lookup(self.peek()[1]) is not None and self.scope.lookup(self.peek()[1]).is_array and \
                    self.t[self.p + 1][1] in (",", ")"):
                obj = self.scope.ref(self.next()[1])          # the array object itself: its bounds are asked for
                rest = []
                if self.accept("op", ","):
                    rest = self.args()
                else:
                    self.expect("op", ")")
                extra = "".join(", " + x[1] for x in rest)
                return f"_fn_{name}({obj}{extra})", "i"
            a = self.args()
            if any(x[0] not in ("expr", "kw") for x in a):
                raise TranslateError(f"{name}(...) with a section but {name} is not an array")
            argc = ", ".join(x[1] for x in a if x[0] == "expr")
            if name in self.scope.module.routines:
                r = self.scope.module.routines[name]
                if not r.is_function:
                    raise TranslateError(f"subroutine {name} used as a function")
                self.scope.module.need(name)
                if r.out_positions():
                    raise TranslateError(f"function {name} sets scalar arguments: only `x = {name}(...)` is in the subset")
                kws = [(x[3], x[1]) for x in a if x[0] == "kw"]
                for kw, _ in kws:
                    if kw not in r.args:
                        raise TranslateError(f"{name}: no dummy argument {kw!r}")
                argc = ", ".join([x[1] for x in a if x[0] == "expr"] + [f"{py_name(kw)}={c}" for kw, c in kws])
                return f"_F_{name}({argc})", r.result_kind
            if any(x[0] == "kw" for x in a):
                raise TranslateError(f"{name}: keyword arguments of intrinsics are outside the subset")
            if self.scope.module.externals.get(name) is not None:       # a function of another module (harness-supplied)
                if isinstance(self.scope.module.externals[name], tuple):
                    raise TranslateError(f"function {name} sets scalar arguments: only `x = {name}(...)` is in the subset")
                return f"_X_{name}({argc})", "?"
            if name in INTRINSICS:
                if name in INT_RESULT:
                    k = "i"
                elif name in REAL_RESULT:
                    k = "r"
                elif name in DOUBLE_RESULT:
                    k = "d"
                elif name in ("allocated", "associated", "present"):
                    k = "l"
                else:
                    ks = [x[2] for x in a]
                    k = ks[0] if ks else "?"
                    for kk in ks[1:]:
                        k = self._join(k, kk)
                return f"_fn_{name}({argc})", k
            raise TranslateError(f"unknown function or array {name!r}")
        if var is None:
            if name in INTRINSICS:
                raise TranslateError(f"intrinsic {name} without arguments")
            raise TranslateError(f"undeclared name {name!r} in {self.scope.routine.name if self.scope.routine else 'module'}")
        if var.is_array:
            if var.derived:
                return self.scope.ref(name), "?"                # an array of objects is its own container
            return f"{self.scope.ref(name)}.a", var.kind
        if var.parameter:
            self.constant.add(self.scope.ref(name))
        return self.scope.ref(name), var.kind


# ============================================================================================ declarations
DECL_RE = re.compile(r"^(real|integer|logical|double\s+precision|complex|character|type\s*\(|class\s*\()")
SPEC_NO_COLONS = re.compile(r"^(real|integer|logical|double\s+precision|complex|character)\s*(\([^)]*\)|\*\s*\d+)?\s+(.*)$")


def split_top(s, sep=","):
    out, cur, depth, q = [], [], 0, None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def parse_declaration(stmt):
    """-> (kind, attrs dict, [(name, dims or None, init or None)]) or None if not a declaration"""
    if not DECL_RE.match(stmt) or re.match(r"^(real|integer|logical|double\s+precision)\s+function\b", stmt):
        return None
    if "::" in stmt:
        left, right = stmt.split("::", 1)
    else:
        m = SPEC_NO_COLONS.match(stmt)
        if not m:
            return None
        left, right = m.group(1) + (m.group(2) or ""), m.group(3)
    parts = split_top(left.strip())
    spec = parts[0].replace(" ", "")
    derived = False
    if spec.startswith("doubleprecision") or spec in ("real(8)", "real*8", "real(kind=8)", "real(dp)", "real(real64)",
                                                        "real(kind=r8size)", "real(r8size)", "real(kind=dp)"):
        kind = "d"
    elif spec.startswith("character"):
        kind = "c"
    elif spec.startswith("complex"):
        kind = "z" if ("double" in spec or "8" in spec or "16" in spec or "dp" in spec) else "x"
    elif spec.startswith("real"):
        kind = "r"
    elif spec.startswith("integer"):
        kind = "i"
    elif spec.startswith("logical"):
        kind = "l"
    elif spec.startswith("type(") or spec.startswith("class("):
        kind, derived = "?", True
    else:
        kind = "?"
    attrs = {"derived": derived}
    for a in parts[1:]:
        a = a.strip()
        al = a.replace(" ", "")
        if al.startswith("dimension("):
            attrs["dims"] = a[a.index("(") + 1:a.rindex(")")]
        elif al.startswith("intent("):
            attrs["intent"] = al[7:-1]
        elif al in ("parameter", "allocatable", "pointer", "target", "optional", "save", "private", "public", "value"):
            attrs[al] = True
        else:
            raise TranslateError(f"declaration attribute {a!r} outside the subset: {stmt}")
    ents = []
    for e in split_top(right.strip()):
        if "=" not in e:                                          # a codimension [*] / [:]: the local part is an ordinary array
            e = re.sub(r"\[[^\]]*\]\s*$", "", e.strip())
        init = None
        m = re.match(r"^(\w+)\s*(\((.*?)\))?\s*(?:\*\s*\d+)?\s*(=\s*(.*))?$", e)
        if not m:
            raise TranslateError(f"cannot parse entity {e!r} in {stmt}")
        if "=>" in e:
            raise TranslateError(f"pointer initialisation outside the subset: {stmt}")
        name, dims, init = m.group(1), m.group(3), m.group(5)
        if kind == "c" and dims is None and re.match(r"^\w+\s*\*\s*\d+$", e):
            name = re.match(r"^(\w+)", e).group(1)
        ents.append((name, dims if dims is not None else attrs.get("dims"), init))
    return kind, attrs, ents


# ============================================================================================ module / routine model
class Routine:
    def __init__(self, name, is_function, args, result, prefix_kind):
        self.name, self.is_function, self.args = name, is_function, args
        self.result = result or (name if is_function else None)
        self.result_kind = prefix_kind
        self.body = []                      # [(file, line, text)]
        self.decls = {}                     # name -> Var
        self.decl_order = []
        self.exec_start = 0
        self.py = None

    def out_positions(self):
        """positions of scalar dummies a call has to copy back (anything not intent(in))"""
        out = []
        for n, a in enumerate(self.args):
            v = self.decls.get(a)
            if v is None:
                raise TranslateError(f"dummy argument {a} of {self.name} is not declared")
            if not v.is_array and not v.derived and v.intent != "in":
                out.append(n)
        return out


class Module:
    """One Fortran module file, translated lazily routine by routine."""

    def __init__(self, path, externals=None, names=None, overrides=None, patches=None):
        self.path = path
        # routine name -> [(regex, replacement statement)]: statements the harness replaces, each one documented where
        # it is used (the only use: the km_s / km_e range that wrf_dm_decomp1d would hand a rank of a parallel run)
        self.patches = patches or {}
        self.component_dtype = {}            # derived-type component -> dtype name for ALLOCATE (default REAL)
        self.overrides = overrides or {}     # routine name -> python callable(module namespace) run INSTEAD of the routine
        self.names = names or {}             # entities of other modules the source refers to (e.g. kVARS)
        self.stmts = read_statements(path)
        self.vars = {}
        self.var_order = []
        self.routines = {}
        self.consts = {}
        self.needed, self.done = [], set()
        self.externals = externals or {}     # routine name -> python callable (routines of other modules)
        self._split()
        self._parse_module_decls()
        for r in self.routines.values():
            try:
                self._parse_routine_decls(r)
            except TranslateError as e:          # reported when (and only when) that routine is asked for
                r.decl_error = e

    # ---- structure
    def _split(self):
        cur, in_contains, self.spec = None, False, []
        for st in self.stmts:
            _, _, t = st
            m = re.match(r"^(?:(real|integer|logical|double\s+precision)\s+)?(?:(?:pure|elemental|recursive|module)\s+)*"
                         r"(subroutine|function)\s+(\w+)\s*(?:\((.*?)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?$", t)
            if m and cur is None and in_contains:
                pk = {"real": "r", "integer": "i", "logical": "l"}.get((m.group(1) or "").split()[0] if m.group(1) else "", None)
                if m.group(1) and m.group(1).startswith("double"):
                    pk = "d"
                args = [a.strip() for a in (m.group(4) or "").split(",") if a.strip()]
                cur = Routine(m.group(3), m.group(2) == "function", args, m.group(5), pk)
                continue
            if cur is not None:
                if re.match(r"^end\s*(subroutine|function)?(\s+\w+)?$", t) and not re.match(r"^end\s*(if|do|associate|where)", t):
                    self.routines[cur.name] = cur
                    cur = None
                else:
                    cur.body.append(st)
                continue
            if t == "contains":
                in_contains = True
                continue
            if not in_contains:
                self.spec.append(st)

    def _parse_module_decls(self):
        for f, no, t in self.spec:
            try:
                d = parse_declaration(t)
            except TranslateError:
                continue        # e.g. a coarray: the entity stays undeclared, a routine that uses it does not translate
            if d is None:
                continue
            kind, attrs, ents = d
            for name, dims, init in ents:
                self.vars[name] = Var(name, kind, dims, None, attrs.get("parameter", False), attrs.get("allocatable", False),
                                      attrs.get("pointer", False), init, module=True, derived=attrs["derived"])
                self.var_order.append(name)

    def _parse_routine_decls(self, r):
        n = 0
        for n, (f, no, t) in enumerate(r.body):
            if re.match(r"^use\b", t) or t.startswith("implicit ") or t.startswith("!$"):
                continue
            d = parse_declaration(t)
            if d is None:
                break
            kind, attrs, ents = d
            for name, dims, init in ents:
                r.decls[name] = Var(name, kind, dims, attrs.get("intent"), attrs.get("parameter", False),
                                    attrs.get("allocatable", False), attrs.get("pointer", False), init,
                                    dummy=name in r.args, derived=attrs["derived"], optional=attrs.get("optional", False))
                r.decl_order.append(name)
        else:
            n = len(r.body)
        r.exec_start = n
        if r.is_function:
            if r.result in r.decls:
                r.result_kind = r.decls[r.result].kind
            elif r.result_kind is None:
                raise TranslateError(f"function {r.name}: result type unknown")
            else:
                r.decls[r.result] = Var(r.result, r.result_kind)
                r.decl_order.append(r.result)

    def need(self, name):
        if name not in self.done and name not in self.needed:
            self.needed.append(name)

    # ---- expressions
    def expr(self, text, scope):
        p = ExprParser(tokenize(text), scope, self.consts)
        c, k = p.expr()
        if p.peek()[0] != "end":
            raise TranslateError(f"trailing tokens in expression {text!r}")
        return c, k

    def bounds(self, dims, scope):
        """'ims:ime, kms:kme' -> python list literal of (lo, hi); ':' -> (1, None)"""
        out, deferred = [], False
        for d in split_top(dims):
            d = d.strip()
            if d == ":":
                out.append("(1, None)")
                deferred = True
                continue
            if ":" in d:
                lo, hi = split_top(d, ":")
                lo = lo.strip() or "1"
                if hi.strip() in ("", "*"):
                    out.append(f"({self.expr(lo, scope)[0]}, None)")
                    deferred = True
                else:
                    out.append(f"({self.expr(lo, scope)[0]}, {self.expr(hi, scope)[0]})")
            else:
                out.append(f"(1, {self.expr(d, scope)[0]})")
        return "[" + ", ".join(out) + "]", deferred, len(out)

    # ---- statements
    def translate(self, name):
        r = self.routines[name]
        if getattr(r, "decl_error", None) is not None:
            raise r.decl_error
        scope = Scope(self, r)
        scope.locals = dict(r.decls)
        L = []                                             # python lines
        ind = 1

        def emit(s, at=None):
            L.append("    " * (ind if at is None else at) + s)

        sig, seen_optional = [], False
        for a in r.args:
            seen_optional = seen_optional or (r.decls.get(a) is not None and r.decls[a].optional)
            sig.append(py_name(a) + ("=None" if seen_optional else ""))
        emit(f"def _F_{name}({', '.join(sig)}):", 0)
        # prologue: dummies under their declared bounds, locals allocated / initialised
        for vn in r.decl_order:
            v = r.decls[vn]
            pn = py_name(vn)
            if v.derived:
                if not v.dummy:          # a local object / array of objects: components are created by assignment
                    emit(f"{pn} = {'_ObjArray()' if v.is_array else '_Obj()'}")
                continue
            if v.kind == "c":
                if not v.dummy:
                    emit(f"{pn} = ''")
                continue
            if v.parameter:
                if v.is_array:
                    b, _, _ = self.bounds(v.dims, scope)
                    emit(f"{pn} = _FArr({self.expr(v.init, scope)[0]}, [lo for lo, hi in {b}])")
                else:
                    emit(f"{pn} = {self.convert(self.expr(v.init, scope)[0], v.kind)}")
                continue
            if v.dummy:
                if v.is_array and v.allocatable:
                    emit(f"assert isinstance({pn}, _FArr), 'an allocatable dummy argument needs the array object'")
                elif v.is_array:
                    b, _, _ = self.bounds(v.dims, scope)
                    if v.optional:
                        emit(f"{pn} = _wrap({pn}, {b}, {DTYPE[v.kind]}) if {pn} is not None else None")
                    else:
                        emit(f"{pn} = _wrap({pn}, {b}, {DTYPE[v.kind]})")
                elif v.optional:
                    emit(f"{pn} = {self.convert(pn, v.kind)} if {pn} is not None else None")
                elif v.intent == "in" or True:
                    emit(f"{pn} = {self.convert(pn, v.kind)}")
                continue
            if v.is_array:
                b, deferred, rank = self.bounds(v.dims, scope)
                if v.pointer:
                    emit(f"{pn} = None")
                elif v.allocatable or deferred:
                    emit(f"{pn} = _unallocated({DTYPE[v.kind]}, {rank})")
                else:
                    emit(f"{pn} = _alloc({DTYPE[v.kind]}, {b})")
            else:
                if v.init is not None:
                    # an initialised local is SAVEd; that is indistinguishable from a constant iff it is never assigned
                    if any(re.match(rf"^(if\s*\(.*\)\s*)?{re.escape(vn)}\s*(\(.*\))?\s*=[^=]", t) for _, _, t in r.body[r.exec_start:]):
                        raise TranslateError(f"{name}: local {vn} has an initialiser (SAVE) and is assigned: outside the subset")
                    emit(f"{pn} = {self.convert(self.expr(v.init, scope)[0], v.kind)}")
                else:
                    emit(f"{pn} = {'False' if v.kind == 'l' else self.convert('0', v.kind)}")
        # executable part
        block = []                                         # stack of open constructs
        body = [(f, no) + split_label(t) for f, no, t in r.body[r.exec_start:]]
        for rx, repl in self.patches.get(name, []):
            hits = [n for n, (f, no, lab, t) in enumerate(body) if re.match(rx, t)]
            if len(hits) != 1:
                raise TranslateError(f"{name}: patch {rx!r} matches {len(hits)} statements")
            body[hits[0]] = body[hits[0]][:3] + (repl,)
        self._labels = {lab: n for n, (f, no, lab, t) in enumerate(body) if lab}
        self._body, self._do_labels = body, []
        for n, (f, no, lab, t) in enumerate(body):
            self._pos = n
            try:
                ind = self._statement(t, scope, r, emit, block, ind)
                # a labelled DO ends with the statement that carries its label
                while lab and self._do_labels and self._do_labels[-1] == lab:
                    self._do_labels.pop()
                    if not block or block.pop() != "do":
                        raise TranslateError("labelled DO ends inside another construct")
                    ind -= 1
            except TranslateError as e:
                raise TranslateError(f"{f}:{no}: {e}\n    {t}") from None
        if block:
            raise TranslateError(f"{name}: unterminated {block[-1]}")
        ind = 1
        emit(self._return(r))
        return "\n".join(L)

    def convert(self, code, kind):
        return {"r": f"_r4({code})", "d": f"_r8({code})", "i": f"_i4({code})", "l": f"bool({code})",
                "x": f"_c64({code})", "z": f"_c128({code})"}.get(kind, code)

    def _return(self, r):
        if r.is_function:
            res = self.convert(py_name(r.result), r.result_kind)
            outs = r.out_positions()
            if outs:        # a function that also sets scalar arguments: (result, those arguments); see _function_with_outs
                return "return (" + res + ", " + ", ".join(py_name(r.args[n]) for n in outs) + ",)"
            return f"return {res}"
        outs = r.out_positions()
        if not outs:
            return "return ()"
        return "return (" + ", ".join(py_name(r.args[n]) for n in outs) + ",)"

    def _lvalue(self, text, scope):
        """-> (kind of target, python target or None, var, is_whole_array, python object expr)"""
        text = text.strip()
        m = re.match(r"^(\w+)\s*(\(.*\))?$", text)
        if m and "%" not in text:
            name = m.group(1)
            if name in scope.aliases:
                obj = scope.aliases[name]
                if m.group(2):
                    p = ExprParser(tokenize(m.group(2)), scope, self.consts)
                    p.expect("op", "(")
                    return "elem", f"{obj}[{ExprParser.subscript(p.args())}]", None, False, obj
                return "whole?", None, None, True, obj
            v = scope.lookup(name)
            if v is None:
                raise TranslateError(f"assignment to undeclared {name!r}")
            ref = scope.ref(name)
            if m.group(2):
                if not v.is_array:
                    raise TranslateError(f"{name} is not an array")
                p = ExprParser(tokenize(m.group(2)), scope, self.consts)
                p.expect("op", "(")
                return "elem", f"{ref}[{ExprParser.subscript(p.args())}]", v, False, ref
            if v.is_array:
                return "whole", None, v, True, ref
            return "scalar", ref, v, False, ref
        # component chain
        p = ExprParser(tokenize(text), scope, self.consts)
        kind, v0 = p.next()
        code = scope.ref(v0) if (v0 in scope.locals or v0 in self.vars or v0 in scope.aliases) else py_name(v0)
        last_sub = False
        while p.accept("op", "%"):
            _, comp = p.next()
            code += "." + py_name(comp)
            last_sub = False
            if p.peek() == ("op", "("):
                p.next()
                code += f"[{ExprParser.subscript(p.args())}]"
                last_sub = True
        if p.peek()[0] != "end":
            raise TranslateError(f"cannot parse assignment target {text!r}")
        if last_sub:
            return "elem", code, None, False, code
        return "whole?", None, None, True, code

    def _assign(self, lhs, rhs, scope, emit):
        co = re.match(r"^(.*\))\s*\[([^\]]+)\]\s*$", lhs.strip())
        if co:
            # a coindexed object  x%comp(section)[image] = ... : the same component of ANOTHER image's object, which the
            # harness hands out through _remote(object, image)  (a coarray PUT)
            root = re.match(r"^(\w+)", co.group(1)).group(1)
            img = self.expr(co.group(2), scope)[0]
            saved = scope.aliases.get(root)
            scope.aliases[root] = f"_remote({scope.ref(root)}, {img})"
            try:
                kind, target, v, whole, obj = self._lvalue(co.group(1), scope)
            finally:
                if saved is None:
                    del scope.aliases[root]
                else:
                    scope.aliases[root] = saved
            if kind != "elem":
                raise TranslateError("coindexed assignment to a whole array is outside the subset")
            emit(f"{target} = {self.expr(rhs, scope)[0]}")
            return
        kind, target, v, whole, obj = self._lvalue(lhs, scope)
        c, _ = self.expr(rhs, scope)
        if kind == "scalar":
            emit(f"{target} = {self.convert(c, v.kind)}")
        elif kind == "elem":
            emit(f"{target} = {c}")
        elif kind == "whole":
            emit(f"{obj}.assign({c})")
        else:       # a derived-type component or an associate name without subscripts: array or scalar, known at run time
            emit(f"_t = {c}")
            if "." in obj and not obj.endswith("]"):
                parent, attr = obj.rsplit(".", 1)
                emit(f"if isinstance(getattr({parent}, {attr!r}, None), _FArr): {obj}.assign(_t)")
            else:
                emit(f"if isinstance({obj}, _FArr): {obj}.assign(_t)")
            emit(f"else: {obj} = _t")

    def _function_with_outs(self, lhs, rhs, scope, emit):
        """`x = f(a, b, c)` where f also sets some of its scalar arguments (intent(inout) / (out)): result and arguments
        come back together and are copied into the actual arguments.  Returns False if rhs is not such a call."""
        m = re.match(r"^\s*(\w+)\s*\((.*)\)\s*$", rhs)
        if not m:
            return False
        name = m.group(1)
        if name in self.routines and self.routines[name].is_function and name not in scope.locals:
            outs, fn = self.routines[name].out_positions(), f"_F_{name}"
            if outs:
                self.need(name)
        elif isinstance(self.externals.get(name), tuple):
            outs, fn = list(self.externals[name][1]), f"_X_{name}"
        else:
            return False
        if not outs:
            return False
        actuals = split_top(m.group(2))
        emit(f"_ret = {fn}({', '.join(self.expr(a, scope)[0] for a in actuals)})")
        kind, target, v, whole, obj = self._lvalue(lhs, scope)
        if kind == "scalar":
            emit(f"{target} = {self.convert('_ret[0]', v.kind)}")
        elif kind == "elem":
            emit(f"{target} = _ret[0]")
        else:
            raise TranslateError("array-valued function with scalar out arguments")
        for k, pos in enumerate(outs, 1):
            kind, target, v, whole, obj = self._lvalue(actuals[pos], scope)
            emit(f"{target} = {self.convert(f'_ret[{k}]', v.kind) if kind == 'scalar' else f'_ret[{k}]'}")
        return True

    def _call(self, t, scope, emit):
        m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", t)
        mb = re.match(r"^call\s+(\w+)\s*%\s*(\w+)\s*(\((.*)\))?$", t)
        if mb and (mb.group(2) in self.routines or mb.group(2) in self.overrides):
            # a type-bound procedure whose implementation is a routine of this module: the object is the first argument
            name, argtext = mb.group(2), mb.group(1) + ("," + mb.group(4) if (mb.group(4) or "").strip() else "")
        elif not m:
            mm = re.match(r"^call\s+([\w%\s]+?)\s*(\((.*)\))?$", t)
            if not mm or "%" not in mm.group(1):
                raise TranslateError("cannot parse call")
            # a type-bound procedure of an object the harness supplies: a method call on that object
            p = ExprParser(tokenize(mm.group(1).strip()), scope, self.consts)
            _, v0 = p.next()
            code = scope.ref(v0) if (v0 in scope.locals or v0 in self.vars or v0 in scope.aliases) else py_name(v0)
            while p.accept("op", "%"):
                code += "." + py_name(p.next()[1])
            args = []
            for a in (split_top(mm.group(3)) if (mm.group(3) or "").strip() else []):
                mk = re.match(r"^(\w+)\s*=(?!=)\s*(.*)$", a.strip())              # keyword = value
                args.append(f"{py_name(mk.group(1))}={self.expr(mk.group(2), scope)[0]}" if mk else self.expr(a, scope)[0])
            emit(f"{code}({', '.join(args)})")
            return
        else:
            name, argtext = m.group(1), m.group(3) or ""
        actuals = split_top(argtext) if argtext.strip() else []
        if name in self.externals and self.externals[name] is None:
            emit("pass")
            return
        kwargs = []
        for a in list(actuals):
            mk = re.match(r"^(\w+)\s*=(?!=)\s*(.*)$", a)
            if mk:
                kwargs.append((mk.group(1), self.expr(mk.group(2), scope)[0]))
                actuals.remove(a)
        codes = [self.expr(a, scope)[0] for a in actuals]
        if name in self.overrides or name in self.externals:
            for n, a in enumerate(actuals):
                v = scope.lookup(a.strip()) if re.fullmatch(r"\w+", a.strip()) else None
                if v is not None and v.is_array and v.allocatable and not v.derived:
                    codes[n] = scope.ref(a.strip())          # an allocatable array goes as the array OBJECT (it may be allocated there)
        if name in self.overrides:
            emit(f"_O_{name}({', '.join(['_M'] + codes + [py_name(k) + '=' + c for k, c in kwargs])})")
            return
        if name in self.routines:
            r = self.routines[name]
            if r.is_function:
                raise TranslateError(f"function {name} called as a subroutine")
            self.need(name)
            outs = r.out_positions()
            given = set(r.args[:len(actuals)]) | {k for k, _ in kwargs}
            if len(actuals) > len(r.args) or any(k not in r.args for k, _ in kwargs) or \
                    any(not r.decls[a].optional for a in r.args if a not in given):
                raise TranslateError(f"call {name}: actual arguments do not match the {len(r.args)} dummies")
            if any(r.args.index(k) in outs for k, _ in kwargs):
                raise TranslateError(f"call {name}: keyword argument for a scalar that is copied back")
            emit(f"_ret = _F_{name}({', '.join(codes + [py_name(k) + '=' + c for k, c in kwargs])})")
            for k, pos in enumerate(outs):
                if pos >= len(actuals):
                    continue
                a = actuals[pos].strip()
                if re.match(r"^\w+\s*(\(.*\))?$", a) or "%" in a:
                    try:
                        kind, target, v, whole, obj = self._lvalue(a, scope)
                    except TranslateError:
                        continue
                    if kind == "scalar":
                        emit(f"{target} = {self.convert(f'_ret[{k}]', v.kind)}")
                    elif kind == "elem":
                        emit(f"{target} = _ret[{k}]")
            return
        if name in self.externals:
            if self.externals[name] is None:
                emit("pass")                      # declared a no-op by the harness (logging)
            elif isinstance(self.externals[name], tuple):
                # (function, positions of the scalar arguments it sets): it returns their new values, in that order
                emit(f"_ret = _X_{name}({', '.join(codes + [py_name(k) + '=' + c for k, c in kwargs])})")
                outs = list(self.externals[name][1])
                for k, pos in enumerate(outs):
                    kind, target, v, whole, obj = self._lvalue(actuals[pos].strip(), scope)
                    val = "_ret" if len(outs) == 1 else f"_ret[{k}]"
                    emit(f"{target} = {self.convert(val, v.kind) if kind == 'scalar' else val}")
            else:
                emit(f"_X_{name}({', '.join(codes + [py_name(k) + '=' + c for k, c in kwargs])})")
            return
        raise TranslateError(f"call to {name!r}: not a routine of this module")

    def _statement(self, t, scope, r, emit_at, block, ind):
        def emit(s):
            emit_at(s, ind)

        if t.startswith("!$") or re.match(r"^use\b", t) or t.startswith("implicit "):
            return ind
        # ---- block ends / middles
        if re.match(r"^end\s*if$", t):
            if not block or block.pop() != "if":
                raise TranslateError("endif without if")
            return ind - 1
        if re.match(r"^end\s*do(\s+\w+)?$", t):
            if not block or block.pop() != "do":
                raise TranslateError("enddo without do")
            return ind - 1
        if re.match(r"^end\s*associate$", t):
            if not block or block.pop() != "associate":
                raise TranslateError("end associate without associate")
            return ind
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", t)
        if m:
            emit_at(f"elif {self.expr(m.group(1), scope)[0]}:", ind - 1)
            emit("pass")
            return ind
        if t == "else":
            emit_at("else:", ind - 1)
            emit("pass")
            return ind
        # ---- block starts
        m = re.match(r"^if\s*\((.*)\)\s*then$", t)
        if m:
            emit(f"if {self.expr(m.group(1), scope)[0]}:")
            block.append("if")
            emit_at("pass", ind + 1)
            return ind + 1
        m = re.match(r"^do\s+while\s*\((.*)\)$", t)
        if m:
            emit(f"while {self.expr(m.group(1), scope)[0]}:")
            block.append("do")
            emit_at("pass", ind + 1)
            return ind + 1
        m = re.match(r"^do\s+(\d+)\s+(\w+)\s*=\s*(.*)$", t)
        if m:
            self._do_labels.append(m.group(1))
            t = f"do {m.group(2)} = {m.group(3)}"
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", t)
        if m:
            var = m.group(1)
            parts = [self.expr(x, scope)[0] for x in split_top(m.group(2))]
            emit(f"for {scope.ref(var)} in _do({', '.join(parts)}):")
            block.append("do")
            emit_at("pass", ind + 1)
            return ind + 1
        if t == "do":
            emit("while True:")
            block.append("do")
            emit_at("pass", ind + 1)
            return ind + 1
        m = re.match(r"^associate\s*\((.*)\)$", t)
        if m:
            for item in split_top(m.group(1)):
                an, ex = item.split("=>", 1)
                p = ExprParser(tokenize(ex.strip()), scope, self.consts)
                c, _ = p.expr()
                if c.startswith("_A(") and c.endswith(")"):
                    c = c[3:-1]                     # the object itself, not its array: the alias may be indexed / assigned
                scope.aliases[an.strip()] = c
            block.append("associate")
            return ind
        # ---- one-line where (mask) lhs = rhs
        if re.match(r"^where\s*\(", t):
            depth = 0
            for k in range(t.index("("), len(t)):
                depth += t[k] == "("
                depth -= t[k] == ")"
                if depth == 0:
                    break
            mask, rest = t[t.index("(") + 1:k], t[k + 1:].strip()
            if not rest:
                raise TranslateError("WHERE construct (block form) is outside the subset")
            lhs, rhs = rest.split("=", 1)
            kind, target, v, whole, obj = self._lvalue(lhs, scope)
            if not whole:
                raise TranslateError("WHERE on a section is outside the subset")
            emit(f"_where_assign({obj}, {self.expr(mask, scope)[0]}, {self.expr(rhs, scope)[0]})")
            return ind
        m = re.match(r"^call\s+c_f_pointer\s*\((.*)\)$", t)
        if m:
            a = split_top(m.group(1))
            _, _, _, _, obj = self._lvalue(a[1], scope)
            emit(f"{obj} = _c_f_pointer({self.expr(a[0], scope)[0]}, {self.expr(a[2], scope)[0]})")
            return ind
        # ---- one-line if
        if re.match(r"^if\s*\(", t):
            depth, k = 0, t.index("(")
            for k in range(t.index("("), len(t)):
                depth += t[k] == "("
                depth -= t[k] == ")"
                if depth == 0:
                    break
            cond, rest = t[t.index("(") + 1:k], t[k + 1:].strip()
            emit(f"if {self.expr(cond, scope)[0]}:")
            return self._statement(rest, scope, r, emit_at, block, ind + 1) - 1
        # ---- simple statements
        if t == "exit" or t == "cycle" or re.match(r"^(exit|cycle)\s+\w+$", t):
            if " " in t and block.count("do") != 1 and False:
                raise TranslateError("exit / cycle of an outer loop")
            emit("break" if t.startswith("exit") else "continue")
            return ind
        if t == "continue" or t.startswith("sync "):        # SYNC IMAGES / SYNC ALL: the harness runs the images in turn
            emit("pass")
            return ind
        m = re.match(r"^go\s*to\s+(\d+)$", t)
        if m:
            # supported: a forward jump out of the innermost DO to a label that follows the loop with nothing but
            # no-ops (PRINT / CONTINUE) in between -- i.e. EXIT
            lab = m.group(1)
            if lab not in self._labels or self._labels[lab] <= self._pos or "do" not in block:
                raise TranslateError("GOTO other than a forward jump out of a loop")
            depth, k = 0, self._pos + 1
            # find the end of the innermost enclosing DO
            while k < len(self._body):
                tt = self._body[k][3]
                if re.match(r"^do\b", tt):
                    depth += 1
                elif re.match(r"^end\s*do", tt) or (self._body[k][2] and self._do_labels and self._body[k][2] == self._do_labels[-1]):
                    if depth == 0:
                        break
                    depth -= 1
                k += 1
            for kk in range(k + 1, self._labels[lab]):
                tt = self._body[kk][3]
                if not (tt.startswith("print") or tt == "continue" or tt.startswith("write")):
                    raise TranslateError(f"GOTO {lab} skips executable statements: outside the subset")
            emit("break")
            return ind
        if re.match(r"^inquire\s*\(", t):
            m = re.search(r"(?:exist|opened)\s*=\s*(\w+)", t)
            if not m:
                raise TranslateError("INQUIRE without EXIST= / OPENED=")
            emit(f"{scope.ref(m.group(1))} = False        # no cache files: tables are computed")
            return ind
        if re.match(r"^(open|close|read|rewind)\s*\(", t):
            emit("pass")
            return ind
        if t == "return":
            emit(self._return(r))
            return ind
        if t.startswith("print") or t.startswith("write") or t.startswith("flush") or t.startswith("pause"):
            emit("pass")
            return ind
        if t == "stop" or t.startswith("stop ") or t.startswith("error stop"):
            emit(f"raise _Stop({t!r})")
            return ind
        if t.startswith("call "):
            self._call(t, scope, emit)
            return ind
        m = re.match(r"^allocate\s*\((.*)\)$", t)
        if m:
            items = split_top(m.group(1))
            source = None
            if items and re.match(r"^source\s*=", items[-1]):
                source = self.expr(items.pop().split("=", 1)[1], scope)[0]
            for item in items:
                item = re.sub(r"\[[^\]]*\]\s*$", "", item.strip())      # allocate(x(...)[*]): the local part
                mm = re.match(r"^(.*?)\((.*)\)$", item.strip())
                if not mm:
                    if source is None:
                        raise TranslateError("allocate without bounds")
                    v = scope.lookup(item.strip())
                    emit(f"{scope.ref(item.strip())} = _alloc_like({source}, {DTYPE[v.kind]})")
                    continue
                target, dims = mm.group(1).strip(), mm.group(2)
                b, _, _ = self.bounds(dims, scope)
                if "%" in target:
                    _, _, _, _, obj = self._lvalue(target, scope)
                    comp = target.split("%")[-1].strip()
                    emit(f"{obj} = _alloc({self.component_dtype.get(comp, '_f32')}, {b}, True)")
                else:
                    v = scope.lookup(target)
                    # into the SAME array object: an allocatable dummy argument must be seen allocated by the caller
                    emit(f"{scope.ref(target)} = _alloc_into({scope.ref(target)}, {DTYPE[v.kind]}, {b})")
                if source is not None:
                    _, _, _, _, obj = self._lvalue(target, scope)
                    emit(f"{obj}.assign({source})")
            return ind
        m = re.match(r"^deallocate\s*\((.*)\)$", t)
        if m:
            for item in split_top(m.group(1)):
                if "%" in item:
                    _, _, _, _, obj = self._lvalue(item.strip(), scope)
                    emit(f"{obj} = None")
                    continue
                v = scope.lookup(item.strip())
                emit(f"{scope.ref(item.strip())} = _unallocated({DTYPE[v.kind]}, {len(split_top(v.dims))})")
            return ind
        # ---- assignments
        depth, q = 0, None
        for k, ch in enumerate(t):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if t[k:k + 2] == "=>":
                    lhs, rhs = t[:k].strip(), t[k + 2:].strip()
                    kind, target, v, whole, obj = self._lvalue(lhs, scope)
                    rv = scope.lookup(rhs)
                    if whole and rv is not None and rv.is_array:
                        emit(f"{obj} = {scope.ref(rhs)}")
                        return ind
                    if whole and re.fullmatch(r"[\w%\s]+", rhs) and "%" in rhs:
                        # pointer => a whole array component of an object of the harness: the array object itself
                        _, _, _, rwhole, robj = self._lvalue(rhs, scope)
                        emit(f"{obj} = {robj}")
                        return ind
                    raise TranslateError("pointer assignment other than `pointer => whole array`")
                if t[k:k + 2] == "==" or (k > 0 and t[k - 1] in "<>/="):
                    continue
                if self._function_with_outs(t[:k], t[k + 1:], scope, emit):
                    return ind
                self._assign(t[:k], t[k + 1:], scope, emit)
                return ind
        raise TranslateError("statement outside the subset")

    # ---- build
    def build(self, entry_points):
        """Translate the entry points and everything they call; returns a namespace object with the python callables
        (attribute = routine name), `M` (module variables and parameters) and `source` (the generated python)."""
        for e in entry_points:
            self.need(e)
        chunks = []
        while self.needed:
            n = self.needed.pop(0)
            if n in self.done:
                continue
            self.done.add(n)
            chunks.append(self.translate(n))
        # module variables and parameters
        mscope = Scope(self, None)
        init = ["def _init_module(_M):"]
        for vn in self.var_order:
            v = self.vars[vn]
            if v.derived:
                continue
            pn = py_name(vn)
            if v.is_array:
                _, deferred, rank = self.bounds(v.dims, mscope) if v.dims else ("", True, 0)
                if v.allocatable or v.pointer or deferred:
                    init.append(f"    _M.{pn} = _unallocated({DTYPE.get(v.kind, '_f32')}, {rank})")
                else:
                    b, _, _ = self.bounds(v.dims, mscope)
                    if v.init is not None:
                        init.append(f"    _M.{pn} = _FArr(_np.asarray(_A({self.expr(v.init, mscope)[0]}), {DTYPE[v.kind]}), [lo for lo, hi in {b}])")
                    else:
                        init.append(f"    _M.{pn} = _alloc({DTYPE[v.kind]}, {b})")
            elif v.init is not None:
                try:
                    init.append(f"    _M.{pn} = {self.convert(self.expr(v.init, mscope)[0], v.kind)}")
                except TranslateError:
                    pass        # outside the subset: the name stays undefined, and any use of it fails loudly at run time
            elif v.kind in "rdi":
                init.append(f"    _M.{pn} = {self.convert('0', v.kind)}")
            elif v.kind == "l":
                init.append(f"    _M.{pn} = False")
        init.append("    return _M")
        src = "\n\n".join(["\n".join(init)] + chunks) + "\n"
        consts = [None] * len(self.consts)
        for text, idx in self.consts.items():
            consts[idx] = R.lit4(text)
        ns = dict(R.NAMES)
        ns["_C"] = consts
        ns["_M"] = types.SimpleNamespace()
        for xn, fn in self.externals.items():
            ns["_X_" + xn] = fn[0] if isinstance(fn, tuple) else fn
        for nn, val in self.names.items():
            ns[py_name(nn)] = val
        for on, fn in self.overrides.items():
            ns["_O_" + on] = fn
        exec(compile(src, f"<f90run:{os.path.basename(self.path)}>", "exec"), ns)
        ns["_init_module"](ns["_M"])
        out = types.SimpleNamespace(M=ns["_M"], source=src, constants=dict(self.consts))
        for n in self.done:
            setattr(out, n, ns["_F_" + n])
        return out
