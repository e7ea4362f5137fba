"""A structural lint of the Fortran shim (icar_b200/fortran/icar_b200_mod.f90), which no compiler in this image can build:
every statement tokenises, every block closes with its own END, every CALL / function reference resolves to a procedure
the module defines, binds (interface) or inherits from the reference's modules, and every procedure it promises (PUBLIC) exists.
Uses the statement reader and tokenizer of oracle/f90run (test infrastructure)."""
import os
import re

from oracle.f90run.translate import read_statements, split_label, tokenize, INTRINSICS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "icar_b200", "fortran", "icar_b200_mod.f90")

# procedures and functions that come from iso_c_binding / the language / ICAR's own modules (used, not defined, by the shim)
EXTERNAL = {"c_loc", "c_associated", "c_f_pointer", "c_null_ptr", "c_null_char", "c_funloc", "present", "associated", "allocated",
            "trim", "len_trim", "adjustl", "real", "int", "size", "lbound", "ubound", "mod", "max", "min", "this_image",
            "num_images", "co_broadcast", "co_min", "co_max", "co_sum", "achar", "char", "transfer", "huge", "merge", "any", "all", "shape", "reshape", "nint", "dble"}


def statements():
    return [(no,) + split_label(t) for _, no, t in read_statements(SHIM)]


def test_every_statement_tokenises():
    for no, _, t in statements():
        tokenize(t)                      # raises TranslateError on an unknown character sequence


def test_blocks_are_balanced_and_named_consistently():
    stack = []
    openers = (("module", r"^module\s+(\w+)$"), ("interface", r"^(abstract\s+)?interface\b"), ("type", r"^type\s*(,\s*bind\(c\))?\s*(::)?\s*\w+$"),
               ("subroutine", r"^(?:\w+\s+)*subroutine\s+(\w+)"), ("function", r"^(?:[\w()=*]+\s+)*function\s+(\w+)"),
               ("if", r"^if\s*\(.*\)\s*then$"), ("do", r"^do(\s|$)"), ("select", r"^select\s+case"), ("associate", r"^associate\s*\("))
    for no, _, t in statements():
        m = re.match(r"^end\s*(module|interface|type|subroutine|function|if|do|select|associate)?\b\s*(\w+)?", t)
        if m and (t.startswith("end")):
            kind = m.group(1)
            assert stack, f"line {no}: END without an open block"
            top_kind, top_name = stack.pop()
            if kind:
                assert kind == top_kind, f"line {no}: END {kind} closes a {top_kind} ({top_name})"
            if m.group(2) and top_name and kind in ("module", "subroutine", "function"):
                assert m.group(2) == top_name, f"line {no}: END {kind} {m.group(2)} closes {top_name}"
            continue
        if re.match(r"^(else|elseif|else\s+if|case|contains)\b", t):
            continue
        for kind, rx in openers:
            mm = re.match(rx, t)
            if mm:
                if kind == "type" and "(" in t.replace("bind(c)", ""):
                    break                                     # a declaration type(x) :: y, not a definition
                name = mm.group(mm.lastindex) if mm.lastindex else None
                stack.append((kind, name))
                break
    assert not stack, f"unclosed blocks: {stack}"


def test_calls_and_public_names_resolve():
    st = statements()
    defined, bound, public = set(), set(), set()
    in_interface = 0
    for no, _, t in st:
        if re.match(r"^(abstract\s+)?interface\b", t):
            in_interface += 1
        elif re.match(r"^end\s*interface", t):
            in_interface -= 1
        m = re.match(r"^(?:[\w()=*]+\s+)*(subroutine|function)\s+(\w+)", t)
        if m and not t.startswith("end"):
            (bound if in_interface else defined).add(m.group(2))
        m = re.match(r"^public\s*::\s*(.*)$", t)
        if m:
            public |= {n.strip() for n in m.group(1).split(",")}
    assert len(defined) >= 15 and len(bound) >= 20
    missing = {n for n in public if n not in defined and not re.match(r"^(var_|icar_)", n)}
    assert not missing, f"PUBLIC names without a definition: {sorted(missing)}"
    imported = set()                     # use <module>, only : a, b => c
    for no, _, t in st:
        m = re.match(r"^use\b.*\bonly\s*:\s*(.*)$", t)
        if m:
            imported |= {n.split("=>")[0].strip() for n in m.group(1).split(",")}
    known = defined | bound | imported | EXTERNAL | INTRINSICS
    for no, _, t in st:
        for m in re.finditer(r"\bcall\s+(\w+)", t):
            assert m.group(1) in known, f"line {no}: call to unknown procedure {m.group(1)}"
        for m in re.finditer(r"\b(icar_\w+)\s*\(", t):
            assert m.group(1) in bound or m.group(1) in defined, f"line {no}: {m.group(1)} is used but has no interface"
