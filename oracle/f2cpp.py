#!/usr/bin/env python3
"""f2cpp.py -- mechanical translator, Fortran-90 subset -> C++ (TEST INFRASTRUCTURE ONLY).

Why: the image has no Fortran compiler, so the reference's HYDROLOGY.f90 / GROW.f90 cannot be
compiled as they are.  This script reads the reference's OWN source files where they lie
(/root/reference/SOURCE, never copied into the repo), translates the statements of the hot path
one for one into C++ and writes the result to oracle/_ref/h9_ref_gen.h (git-ignored).  g++ then
builds oracle/_ref/libh9ref.so from that header plus oracle/ref_harness.cpp.  The translation is
syntactic: no statement is re-ordered, merged, simplified or re-associated, every REAL stays a
32-bit float, and every intrinsic maps onto the libm function gfortran itself would call
(EXP -> expf, LOG -> logf, real ** real -> powf, real ** integer -> repeated multiplication).
libh9ref.so is therefore "the reference's own text, compiled", and it is what pins the
hand-written restatement oracle/h9_oracle.cpp (tests/test_ref_vs_oracle.py).

What is translated (FRAGMENTS below): module CONTROL and module SHARED (every declaration),
SUBROUTINE HYDROLOGY and SUBROUTINE GROW (whole files), and line ranges of INIT.f90 and
HYBRID9.f90 (allocation, layer geometry, initial state, calendar, the cell/year/day loop nest).
Line ranges are used for INIT and PROGRAM H9 because the rest of those units is MPI, netCDF and
file I/O (and HYBRID9.f90:2 holds a stray token that no compiler would accept).

Supported subset: free-form source, `&` continuations, `!` comments; REAL / INTEGER / LOGICAL
declarations with PARAMETER, DIMENSION, ALLOCATABLE, explicit bounds; CHARACTER entities are
dropped (only I/O uses them); assignment (scalar, element, whole-array and `:` section);
DO (optional step, optional construct name) / EXIT / CYCLE; block IF / ELSE IF / ELSE, logical
IF; CALL of a translated unit; ALLOCATE; STOP (-> H9Ref::f2c_stop, which throws);
WRITE / PRINT / OPEN / CLOSE are dropped (diagnostics only).  Expressions: Fortran precedence
and associativity (`**` right-associative and above unary minus), integer vs real typing with
the standard's promotion rules, integer division, relational and logical operators, intrinsics
EXP LOG ABS SQRT MAX MIN SUM MINVAL MAXVAL FLOAT REAL INT NINT MOD.
Anything outside the subset raises an error instead of being guessed at.

Usage: f2cpp.py <reference SOURCE dir> <output header>
"""
from __future__ import annotations

import re
import sys
from dataclasses import dataclass

# ------------------------------------------------------------------------------------------
# what to translate: (function name, file, first line, last line) -- 1-based, inclusive
# ------------------------------------------------------------------------------------------
MODULES = [("control", "CONTROL.f90"), ("shared", "SHARED.f90")]
SUBROUTINES = [("hydrology", "HYDROLOGY.f90"), ("grow", "GROW.f90")]
PROGRAM_DECLS = ("h9", "HYBRID9.f90", 46, 73)  # locals of PROGRAM H9 (the annual sums)
FRAGMENTS = [
    # INIT.f90 -----------------------------------------------------------------------------
    ("init_alloc_scratch", "INIT.f90", 53, 136, "init"),   # sla, theta_sum, HYDROLOGY work vectors
    ("init_gpt", "INIT.f90", 154, 154, "init"),            # sla (1) = 23.0E-3
    ("init_dt", "INIT.f90", 214, 214, "init"),             # dt = 86400.0 / FLOAT (NISURF)
    ("init_layers", "INIT.f90", 252, 263, "init"),         # dz, zc, zc_o
    ("init_alloc_grid_a", "INIT.f90", 301, 355, "init"),   # state, axy_*, soil_tex, Fmax ...
    ("init_alloc_grid_b", "INIT.f90", 374, 395, "init"),   # theta_s, hksat, lambda, bsw, psi_s
    ("init_alloc_l1", "INIT.f90", 360, 370, "init"),       # 30 arc-second input tiles and their block means
    ("init_fills", "INIT.f90", 402, 414, "init"),          # NaN / zero fills of axy_*
    ("init_regrid_layer", "INIT.f90", 575, 632, "init"),   # 60x60 block means -> layer I of the soil fields
    ("init_state", "INIT.f90", 711, 811, "init"),          # initial state of every land cell
    ("init_time_boy", "INIT.f90", 844, 859, "init"),       # calendar
    ("alloc_forcing", "INIT.f90", 901, 907, "init"),       # same shapes as READ_PGF.f90:33-108
    # HYBRID9.f90 --------------------------------------------------------------------------
    ("decade_years", "HYBRID9.f90", 103, 113, "h9"),       # syr, eyr
    ("decade_loop", "HYBRID9.f90", 120, 295, "h9"),        # the loop nest, verbatim
    ("cell_years", "HYBRID9.f90", 126, 292, "h9"),         # body of the land IF (one cell)
    ("day_derive", "HYBRID9.f90", 156, 189, "h9"),         # forcing derivation of one day
    ("day_substeps", "HYBRID9.f90", 193, 211, "h9"),       # NISURF x CALL HYDROLOGY
]

CXX_KEYWORDS = {"int", "float", "double", "do", "if", "else", "for", "while", "new", "delete",
                "class", "struct", "this", "char", "long", "short", "signed", "unsigned", "void",
                "const", "static", "return", "switch", "case", "default", "break", "continue",
                "goto", "union", "enum", "template", "typename", "namespace", "operator", "and",
                "or", "not", "xor", "true", "false", "bool", "auto", "register", "volatile"}
IGNORED_ENTITIES = {"status"}  # INTEGER :: status (MPI_STATUS_SIZE), CONTROL.f90:83: MPI only
IO_STATEMENTS = ("write", "print", "open", "close", "read")


class TranslateError(Exception):
    pass


# ------------------------------------------------------------------------------------------
# source reading: comments, continuations, case folding
# ------------------------------------------------------------------------------------------
def strip_comment(line: str, inq):
    """Returns (code without comment, quote state at end of line)."""
    out = []
    for ch in line:
        if inq:
            out.append(ch)
            if ch == inq:
                inq = None
        elif ch in "'\"":
            inq = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out), inq


def fold_case(s: str) -> str:
    out, inq = [], None
    for ch in s:
        if inq:
            out.append(ch)
            if ch == inq:
                inq = None
        else:
            if ch in "'\"":
                inq = ch
            out.append(ch.lower())
    return "".join(out)


def logical_lines(path: str, first: int = 1, last: int | None = None):
    """[(line number of first physical line, statement text)] with continuations joined."""
    with open(path, "r", errors="replace") as f:
        phys = f.read().split("\n")
    if last is None:
        last = len(phys)
    res, cur, cur_no, inq, cont = [], "", None, None, False
    for no in range(first, last + 1):
        raw = phys[no - 1].rstrip("\r")
        code, inq2 = strip_comment(raw, inq if cont else None)
        code = code.strip()
        if not code:
            continue  # comment / blank lines may sit between continuation lines
        if cont and code.startswith("&"):
            code = code[1:].lstrip() if not inq else code[1:]
        if cur_no is None:
            cur_no = no
        if code.endswith("&"):
            cur += code[:-1].rstrip() + (" " if not inq2 else "")
            cont, inq = True, inq2
            continue
        cur += code
        res.append((cur_no, fold_case(cur)))
        cur, cur_no, inq, cont = "", None, None, False
    if cur:
        raise TranslateError(f"{path}:{cur_no}: dangling continuation")
    return res


# ------------------------------------------------------------------------------------------
# tokens and expression parser (Fortran precedence)
# ------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<real>(?:\d+\.\d*|\.\d+)(?:[ed][+-]?\d+)?|\d+[ed][+-]?\d+)
  | (?P<int>\d+)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|==|/=|<=|>=|//|\(/|/\)|::|[-+*/<>=(),:%])
""", re.X)


def tokenize(s: str, where: str):
    pos, toks = 0, []
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise TranslateError(f"{where}: cannot tokenize at '{s[pos:pos + 20]}'")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append((kind, m.group()))
    return toks


DOT_REL = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}
REL_OPS = {"==", "/=", "<", "<=", ">", ">="}


class Parser:
    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "dotop"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise TranslateError(f"{self.where}: expected '{val}', found '{self.peek()[1]}'")

    def at_end(self):
        return self.i >= len(self.t)

    # level 8 (lowest) .or.
    def expr(self):
        a = self.p_and()
        while self.accept(".or."):
            a = ("bin", ".or.", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.accept(".and."):
            a = ("bin", ".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.accept(".not."):
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        k, v = self.peek()
        if k == "dotop" and v in DOT_REL:
            self.next()
            return ("bin", DOT_REL[v], a, self.p_add())
        if k == "op" and v in REL_OPS:
            self.next()
            return ("bin", v, a, self.p_add())
        return a

    def p_add(self):
        # a leading sign applies to the whole first term:  -a*b  ==  -(a*b)
        if self.accept("-"):
            a = ("un", "-", self.p_mul())
        elif self.accept("+"):
            a = self.p_mul()
        else:
            a = self.p_mul()
        while True:
            if self.accept("+"):
                a = ("bin", "+", a, self.p_mul())
            elif self.accept("-"):
                a = ("bin", "-", a, self.p_mul())
            else:
                return a

    def p_mul(self):
        a = self.p_pow()
        while True:
            if self.accept("*"):
                a = ("bin", "*", a, self.p_pow())
            elif self.accept("/"):
                a = ("bin", "/", a, self.p_pow())
            else:
                return a

    def p_pow(self):
        a = self.p_primary()
        if self.accept("**"):
            # right-associative; a signed exponent needs parentheses in standard Fortran
            return ("bin", "**", a, self.p_pow())
        return a

    def p_primary(self):
        k, v = self.next()
        if k == "int":
            return ("num", v, "int")
        if k == "real":
            return ("num", v, "double" if "d" in v else "real")
        if k == "dotop" and v in (".true.", ".false."):
            return ("logical", v == ".true.")
        if k == "str":
            return ("str", v)
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if k == "name":
            if self.peek() == ("op", "("):
                self.next()
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.p_subscript())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return ("call", v, args)
            return ("name", v)
        raise TranslateError(f"{self.where}: unexpected token '{v}'")

    def p_subscript(self):
        lo = hi = None
        if self.peek() == ("op", ":"):
            self.next()
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return ("range", None, hi)
        lo = self.expr()
        if self.accept(":"):
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return ("range", lo, hi)
        return lo


def parse_expr(text: str, where: str):
    p = Parser(tokenize(text, where), where)
    e = p.expr()
    if not p.at_end():
        raise TranslateError(f"{where}: trailing tokens after expression in '{text}'")
    return e


def split_top(s: str, sep: str = ","):
    """split at top-level separators (outside parentheses and quotes)"""
    parts, depth, cur, inq = [], 0, [], None
    for ch in s:
        if inq:
            cur.append(ch)
            if ch == inq:
                inq = None
            continue
        if ch in "'\"":
            inq = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def match_paren(s: str, start: int) -> int:
    """index of the ')' matching the '(' at s[start]"""
    depth, inq = 0, None
    for i in range(start, len(s)):
        ch = s[i]
        if inq:
            if ch == inq:
                inq = None
            continue
        if ch in "'\"":
            inq = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return i
    raise TranslateError(f"unbalanced parentheses in '{s}'")


# ------------------------------------------------------------------------------------------
# symbols
# ------------------------------------------------------------------------------------------
@dataclass
class Sym:
    name: str           # Fortran name (lower case)
    cname: str          # C++ member name
    typ: str            # 'real' | 'int' | 'logical'
    rank: int = 0
    dims: list | None = None      # [(lo_ast|None, hi_ast)] for explicit shape; None if allocatable
    param: object = None          # AST of a PARAMETER's value
    where: str = ""
    scope: str = ""


class Scope:
    def __init__(self, name, parents=()):
        self.name, self.syms, self.parents = name, {}, list(parents)
        self.dropped = set()  # CHARACTER entities and other ignored names

    def lookup(self, n):
        if n in self.syms:
            return self.syms[n]
        for p in self.parents:
            s = p.lookup(n)
            if s:
                return s
        return None

    def is_dropped(self, n):
        return n in self.dropped or any(p.is_dropped(n) for p in self.parents)


DECL_RE = re.compile(r"^(real|integer|logical|character|double\s+precision)\b")


def parse_declaration(stmt: str, where: str, scope: Scope, prefix: str):
    """REAL[, attrs] :: entity-list  -> Sym entries in scope"""
    m = DECL_RE.match(stmt)
    base = m.group(1)
    rest = stmt[m.end():].lstrip()
    if rest.startswith("("):  # kind / len selector
        j = match_paren(rest, 0)
        rest = rest[j + 1:].lstrip()
    if "::" in rest:
        attr_text, ents = rest.split("::", 1)
    else:
        attr_text, ents = "", rest
    attrs = [a.strip() for a in split_top(attr_text.strip().lstrip(","))] if attr_text.strip() else []
    is_param = any(a == "parameter" for a in attrs)
    is_alloc = any(a == "allocatable" for a in attrs)
    dim_attr = None
    for a in attrs:
        if a.startswith("dimension"):
            dim_attr = a[a.index("(") + 1: match_paren(a, a.index("("))]
        elif a not in ("parameter", "allocatable", "save") and not a.startswith("intent"):
            raise TranslateError(f"{where}: unsupported attribute '{a}'")
    if base.startswith("double"):
        raise TranslateError(f"{where}: DOUBLE PRECISION is outside the subset")
    typ = {"real": "real", "integer": "int", "logical": "logical", "character": None}[base]
    for ent in split_top(ents):
        if not ent:
            continue
        init = None
        if "=" in ent:
            k = _top_level_eq(ent)
            if k >= 0:
                ent, init = ent[:k].strip(), ent[k + 1:].strip()
        m2 = re.match(r"^([a-z_][a-z0-9_]*)\s*(\(.*\))?$", ent)
        if not m2:
            raise TranslateError(f"{where}: cannot parse entity '{ent}'")
        name, shape = m2.group(1), m2.group(2)
        if typ is None or name in IGNORED_ENTITIES:
            scope.dropped.add(name)
            continue
        shape_text = shape[1:-1] if shape else dim_attr
        dims, rank = None, 0
        if shape_text is not None:
            parts = split_top(shape_text)
            rank = len(parts)
            if all(p == ":" for p in parts):
                if not is_alloc:
                    raise TranslateError(f"{where}: deferred shape without ALLOCATABLE: {name}")
            else:
                dims = []
                for p in parts:
                    sub = split_top(p, ":")
                    if len(sub) == 1:
                        dims.append((None, parse_expr(sub[0], where)))
                    else:
                        dims.append((parse_expr(sub[0], where), parse_expr(sub[1], where)))
        cname = prefix + name + "_"   # trailing underscore: no clash with libm (Fmax -> fmax) or C++ names
        if cname in CXX_KEYWORDS:
            raise TranslateError(f"{where}: '{name}' collides with a C++ keyword")
        if is_param and init is None:
            raise TranslateError(f"{where}: PARAMETER without value: {name}")
        if init is not None and not is_param:
            raise TranslateError(f"{where}: initialised non-PARAMETER entity '{name}' is outside the subset")
        scope.syms[name] = Sym(name, cname, typ, rank, dims,
                               parse_expr(init, where) if init is not None else None, where,
                               scope.name)


def _top_level_eq(s: str) -> int:
    depth = 0
    for i, ch in enumerate(s):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if i + 1 < len(s) and s[i + 1] == "=":
                continue
            if i > 0 and s[i - 1] in "=/<>":
                continue
            return i
    return -1


# ------------------------------------------------------------------------------------------
# expression emitter
# ------------------------------------------------------------------------------------------
CTYPE = {"real": "float", "int": "int", "logical": "bool"}
SUBROUTINE_NAMES = {n for n, _ in SUBROUTINES}


def real_literal(text: str) -> str:
    t = text.replace("d", "e")
    if "." not in t and "e" not in t:
        t += ".0"
    return t + "f"


class Emitter:
    """Turns ASTs into C++ text with Fortran typing rules."""

    def __init__(self, scope: Scope, where: str):
        self.scope, self.where = scope, where
        self.secvars = None  # [(loop var, lhs_lo_code)] while emitting a section assignment

    def err(self, msg):
        raise TranslateError(f"{self.where}: {msg}")

    def to_real(self, code, typ):
        if typ == "real":
            return code
        if typ == "int":
            return f"(float)({code})"
        self.err(f"cannot convert {typ} to real: {code}")

    def unify(self, a, ta, b, tb):
        if ta == tb:
            return a, b, ta
        if {ta, tb} == {"int", "real"}:
            return self.to_real(a, ta), self.to_real(b, tb), "real"
        self.err(f"type mismatch {ta} vs {tb}: {a} ; {b}")

    def expr(self, e):
        """-> (code, type)"""
        k = e[0]
        if k == "num":
            if e[2] == "int":
                return e[1], "int"
            if e[2] == "double":
                self.err("double-precision literal is outside the subset")
            return real_literal(e[1]), "real"
        if k == "logical":
            return ("true" if e[1] else "false"), "logical"
        if k == "paren":
            c, t = self.expr(e[1])
            return f"({c})", t
        if k == "name":
            return self.name(e[1])
        if k == "call":
            return self.call(e[1], e[2])
        if k == "un":
            c, t = self.expr(e[2])
            if e[1] == "-":
                if t not in ("int", "real"):
                    self.err("unary minus on non-numeric")
                return f"(-{c})", t
            if t != "logical":
                self.err(".not. on non-logical")
            return f"(!{c})", "logical"
        if k == "bin":
            return self.binop(e[1], e[2], e[3])
        if k == "str":
            self.err("character expression outside I/O")
        self.err(f"unsupported expression node {k}")

    def name(self, n):
        if self.scope.is_dropped(n) and not self.scope.lookup(n):
            self.err(f"use of dropped (CHARACTER / MPI) entity '{n}'")
        s = self.scope.lookup(n)
        if not s:
            self.err(f"undeclared name '{n}'")
        if s.rank > 0:
            if self.secvars is None:
                self.err(f"whole-array reference '{n}' outside an array assignment")
            return self.index(s, [("range", None, None)] * s.rank)
        return s.cname, s.typ

    def index(self, s: Sym, subs):
        if len(subs) != s.rank:
            self.err(f"rank mismatch for '{s.name}': {len(subs)} subscripts, rank {s.rank}")
        codes, nsec = [], 0
        for d, sub in enumerate(subs):
            if sub[0] == "range":
                if self.secvars is None:
                    self.err(f"array section of '{s.name}' outside an array assignment / reduction")
                if nsec >= len(self.secvars):
                    self.err(f"section rank mismatch in reference to '{s.name}'")
                var, lhs_lo = self.secvars[nsec]
                lo = self.expr(sub[1])[0] if sub[1] is not None else f"{s.cname}.lb({d})"
                codes.append(f"({var} - ({lhs_lo}) + ({lo}))")
                nsec += 1
            else:
                c, t = self.expr(sub)
                if t != "int":
                    self.err(f"non-integer subscript of '{s.name}'")
                codes.append(c)
        if self.secvars is not None and nsec not in (0, len(self.secvars)):
            self.err(f"section rank mismatch in reference to '{s.name}'")
        return f"{s.cname}({', '.join(codes)})", s.typ

    def reduction(self, fname, args):
        """SUM / MINVAL / MAXVAL of one array (section) argument, elements in index order."""
        if len(args) != 1:
            self.err(f"{fname} with DIM/MASK is outside the subset")
        a = args[0]
        if a[0] == "name":
            s = self.scope.lookup(a[1])
            subs = [("range", None, None)] * (s.rank if s else 0)
        elif a[0] == "call":
            s, subs = self.scope.lookup(a[1]), a[2]
        else:
            self.err(f"{fname} of an expression is outside the subset")
        if not s or s.rank == 0:
            self.err(f"{fname} of a non-array")
        ranges = [(d, sub) for d, sub in enumerate(subs) if sub[0] == "range"]
        if len(ranges) != 1:
            self.err(f"{fname} over {len(ranges)} section dimensions is outside the subset")
        d, sub = ranges[0]
        lo = self.expr(sub[1])[0] if sub[1] is not None else f"{s.cname}.lb({d})"
        hi = self.expr(sub[2])[0] if sub[2] is not None else f"{s.cname}.ub({d})"
        saved = self.secvars
        self.secvars = [("_r", lo)]
        elem, typ = self.index(s, subs)
        self.secvars = saved
        ct = CTYPE[typ]
        if fname == "sum":
            body = f"{ct} _a = 0; for (int _r = ({lo}); _r <= ({hi}); ++_r) _a = _a + {elem}; return _a;"
        else:
            cmp = "<" if fname == "minval" else ">"
            body = (f"int _r = ({lo}); {ct} _a = {elem}; for (_r = ({lo}) + 1; _r <= ({hi}); ++_r) "
                    f"{{ const {ct} _v = {elem}; if (_v {cmp} _a) _a = _v; }} return _a;")
        return f"([&]() -> {ct} {{ {body} }})()", typ

    def call(self, n, args):
        s = self.scope.lookup(n)
        if s and s.rank > 0:
            return self.index(s, args)
        if s and s.rank == 0:
            self.err(f"'{n}' is a scalar but is subscripted / called")
        if any(a[0] == "range" for a in args) and n not in ("sum", "minval", "maxval"):
            self.err(f"section argument to '{n}'")
        if n in ("sum", "minval", "maxval"):
            return self.reduction(n, args)
        ev = [self.expr(a) for a in args]
        if n in ("exp", "log", "sqrt"):
            (c, t), = ev
            if t != "real":
                self.err(f"{n} of non-real")
            return f"{n}f({c})", "real"
        if n == "abs":
            (c, t), = ev
            return (f"fabsf({c})", "real") if t == "real" else (f"abs({c})", "int")
        if n in ("max", "min"):
            if len(ev) < 2:
                self.err(f"{n} needs two or more arguments")
            c, t = ev[0]
            for c2, t2 in ev[1:]:
                a, b, t = self.unify(c, t, c2, t2)
                c = f"f2c_{n}({a}, {b})"
            return c, t
        if n in ("float", "real"):
            (c, t), = ev
            return self.to_real(c, t), "real"
        if n == "int":
            (c, t), = ev
            return f"(int)({c})", "int"
        if n == "nint":
            (c, t), = ev
            return f"(int)lroundf({c})", "int"
        if n == "mod":
            (a, ta), (b, tb) = ev
            if ta == "int" and tb == "int":
                return f"(({a}) % ({b}))", "int"
            a, b, t = self.unify(a, ta, b, tb)
            return f"fmodf({a}, {b})", "real"
        self.err(f"unknown function or undeclared array '{n}'")

    def binop(self, op, l, r):
        a, ta = self.expr(l)
        b, tb = self.expr(r)
        if op in ("+", "-", "*", "/"):
            if ta not in ("int", "real") or tb not in ("int", "real"):
                self.err(f"arithmetic on non-numeric operands ({op})")
            a, b, t = self.unify(a, ta, b, tb)
            return f"({a} {op} {b})", t   # int / int is truncating division in both languages
        if op == "**":
            if tb == "int":
                if ta == "real":
                    return f"f2c_powi({a}, {b})", "real"
                if ta == "int":
                    return f"f2c_ipow({a}, {b})", "int"
            if ta in ("int", "real") and tb == "real":
                return f"powf({self.to_real(a, ta)}, {b})", "real"
            self.err("unsupported operand types for **")
        if op in REL_OPS:
            a, b, _ = self.unify(a, ta, b, tb)
            return f"({a} {'!=' if op == '/=' else op} {b})", "logical"
        if op in (".and.", ".or."):
            if ta != "logical" or tb != "logical":
                self.err(f"{op} on non-logical operands")
            return f"({a} {'&&' if op == '.and.' else '||'} {b})", "logical"
        self.err(f"unsupported operator {op}")


# ------------------------------------------------------------------------------------------
# statement translation
# ------------------------------------------------------------------------------------------
class Unit:
    """One translated function body."""

    def __init__(self, name, fname, scope, lines):
        self.name, self.fname, self.scope, self.lines = name, fname, scope, lines
        self.out, self.ind = [], 1
        self.blocks = []     # stack of ('do', label) / ('if',)
        self.nstop = 0
        self.tmp = 0
        self.sync = None     # name of the write-back macro of a whole-subroutine unit

    def emit(self, s):
        self.out.append("  " * self.ind + s)

    def translate(self, stop_base=0):
        self.nstop = stop_base
        for no, stmt in self.lines:
            where = f"{self.fname}:{no}"
            try:
                self.statement(stmt, where, no)
            except TranslateError:
                raise
            except Exception as ex:  # parser bugs should name the line
                raise TranslateError(f"{where}: {type(ex).__name__}: {ex} in '{stmt}'")
        if self.blocks:
            raise TranslateError(f"{self.fname}: unterminated block {self.blocks[-1]} in {self.name}")
        return self.out

    def statement(self, stmt, where, no):
        s = stmt.strip()
        em = Emitter(self.scope, where)
        # ---- structural lines of a whole-file unit
        if re.match(r"^(subroutine|end\s*subroutine|use|implicit|save|contains|return)\b", s):
            return
        if DECL_RE.match(s) and ("::" in s or re.match(r"^(real|integer|logical)\s+[a-z_]", s)):
            return  # declarations were collected in the first pass
        self.emit(f"/* {where} */")
        # ---- construct name
        label = None
        m = re.match(r"^([a-z_][a-z0-9_]*)\s*:\s*(do\b.*)$", s)
        if m:
            label, s = m.group(1), m.group(2)
        # ---- DO
        m = re.match(r"^do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", s)
        if m:
            var = self.scope.lookup(m.group(1))
            if not var or var.typ != "int" or var.rank:
                em.err(f"DO variable '{m.group(1)}' is not an integer scalar")
            parts = split_top(m.group(2))
            if len(parts) not in (2, 3):
                em.err("malformed DO control")
            lo, tlo = em.expr(parse_expr(parts[0], where))
            hi, thi = em.expr(parse_expr(parts[1], where))
            if tlo != "int" or thi != "int":
                em.err("non-integer DO bounds")
            self.tmp += 1
            e, st = f"_e{self.tmp}", f"_s{self.tmp}"
            v = var.cname
            self.emit("{")
            self.ind += 1
            # bounds and step are evaluated once, before the loop (Fortran 2003 8.1.6.4.1)
            self.emit(f"const int _b{self.tmp} = {lo}; const int {e} = {hi};")
            if len(parts) == 3:
                sc, ts = em.expr(parse_expr(parts[2], where))
                if ts != "int":
                    em.err("non-integer DO step")
                self.emit(f"const int {st} = {sc};")
                self.emit(f"for ({v} = _b{self.tmp}; ({st} > 0) ? ({v} <= {e}) : ({v} >= {e}); {v} += {st}) {{")
            else:
                self.emit(f"for ({v} = _b{self.tmp}; {v} <= {e}; {v} += 1) {{")
            self.ind += 1
            self.blocks.append(("do", label))
            return
        if re.match(r"^end\s*do\b", s):
            if not self.blocks or self.blocks[-1][0] != "do":
                em.err("END DO without DO")
            _, lab = self.blocks.pop()
            self.ind -= 1
            self.emit("}")
            self.ind -= 1
            self.emit("}")
            if lab:
                self.emit(f"{lab}_exit: ;")
            return
        # ---- IF family
        if re.match(r"^else\s*if\b", s):
            j = s.index("(")
            k = match_paren(s, j)
            if s[k + 1:].strip() != "then":
                em.err("ELSE IF without THEN")
            c, t = em.expr(parse_expr(s[j + 1:k], where))
            if t != "logical":
                em.err("non-logical IF condition")
            self.ind -= 1
            self.emit(f"}} else if ({c}) {{")
            self.ind += 1
            return
        if re.match(r"^else$", s):
            self.ind -= 1
            self.emit("} else {")
            self.ind += 1
            return
        if re.match(r"^end\s*if\b", s):
            if not self.blocks or self.blocks[-1][0] != "if":
                em.err("END IF without IF")
            self.blocks.pop()
            self.ind -= 1
            self.emit("}")
            return
        if re.match(r"^if\s*\(", s):
            j = s.index("(")
            k = match_paren(s, j)
            c, t = em.expr(parse_expr(s[j + 1:k], where))
            if t != "logical":
                em.err("non-logical IF condition")
            rest = s[k + 1:].strip()
            if rest == "then":
                self.emit(f"if ({c}) {{")
                self.ind += 1
                self.blocks.append(("if",))
            else:  # logical IF: one action statement
                self.emit(f"if ({c}) {{")
                self.ind += 1
                self.simple(rest, where, em)
                self.ind -= 1
                self.emit("}")
            return
        self.simple(s, where, em)

    def simple(self, s, where, em):
        """action statements"""
        first = re.match(r"^[a-z_]+", s)
        kw = first.group() if first else ""
        if kw in IO_STATEMENTS and re.match(r"^(write|print|open|close|read)\s*[(*]", s):
            self.emit("; /* I/O statement dropped (diagnostic output only) */")
            return
        if re.match(r"^stop\b", s):
            self.nstop += 1
            if self.sync:
                self.emit(f"{self.sync};")
            self.emit(f"f2c_stop({self.nstop}, {where.split(':')[1]});")
            return
        m = re.match(r"^exit(?:\s+([a-z_][a-z0-9_]*))?$", s)
        if m:
            if m.group(1):
                if not any(b[0] == "do" and b[1] == m.group(1) for b in self.blocks):
                    em.err(f"EXIT from unknown construct '{m.group(1)}'")
                self.emit(f"goto {m.group(1)}_exit;")
            else:
                self.emit("break;")
            return
        if s == "cycle":
            self.emit("continue;")
            return
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*(\(.*\))?$", s)
        if m:
            if m.group(1) not in SUBROUTINE_NAMES or m.group(2):
                em.err(f"CALL of '{m.group(1)}' is outside the translated set")
            self.emit(f"{m.group(1)}();")
            return
        m = re.match(r"^allocate\s*\(", s)
        if m:
            j = s.index("(")
            k = match_paren(s, j)
            for item in split_top(s[j + 1:k]):
                m2 = re.match(r"^([a-z_][a-z0-9_]*)\s*\((.*)\)$", item)
                if not m2:
                    em.err(f"cannot parse ALLOCATE item '{item}'")
                sym = self.scope.lookup(m2.group(1))
                if not sym or sym.dims is not None or sym.rank == 0:
                    em.err(f"ALLOCATE of non-allocatable '{m2.group(1)}'")
                dims = split_top(m2.group(2))
                if len(dims) != sym.rank:
                    em.err(f"ALLOCATE rank mismatch for '{sym.name}'")
                los, his = [], []
                for d in dims:
                    sub = split_top(d, ":")
                    lo = "1" if len(sub) == 1 else em.expr(parse_expr(sub[0], where))[0]
                    hi = em.expr(parse_expr(sub[-1], where))[0]
                    los.append(lo)
                    his.append(hi)
                self.emit(f"{sym.cname}.alloc({{{', '.join(los)}}}, {{{', '.join(his)}}});")
            return
        # ---- assignment
        k = _top_level_eq(s)
        if k < 0:
            em.err(f"unsupported statement '{s}'")
        lhs_t, rhs_t = s[:k].strip(), s[k + 1:].strip()
        lhs = parse_expr(lhs_t, where)
        rhs = parse_expr(rhs_t, where)
        if lhs[0] == "name":
            sym = self.scope.lookup(lhs[1])
            subs = [("range", None, None)] * sym.rank if sym and sym.rank else None
        elif lhs[0] == "call":
            sym, subs = self.scope.lookup(lhs[1]), lhs[2]
        else:
            em.err(f"invalid assignment target '{lhs_t}'")
        if not sym:
            em.err(f"assignment to undeclared '{lhs_t}'")
        if sym.param is not None:
            em.err(f"assignment to PARAMETER '{sym.name}'")
        if sym.rank == 0:
            if subs:
                em.err(f"subscripted scalar '{sym.name}'")
            c, t = em.expr(rhs)
            self.emit(f"{sym.cname} = {self.convert(c, t, sym.typ, em)};")
            return
        ranges = [(d, sub) for d, sub in enumerate(subs) if sub[0] == "range"]
        if not ranges:
            tgt, _ = em.index(sym, subs)
            c, t = em.expr(rhs)
            self.emit(f"{tgt} = {self.convert(c, t, sym.typ, em)};")
            return
        # array (section) assignment: elementwise, first section dimension innermost
        secvars, loops = [], []
        for n, (d, sub) in enumerate(ranges):
            lo = em.expr(sub[1])[0] if sub[1] is not None else f"{sym.cname}.lb({d})"
            hi = em.expr(sub[2])[0] if sub[2] is not None else f"{sym.cname}.ub({d})"
            var = f"_i{n}"
            secvars.append((var, lo))
            loops.append(f"for (int {var} = ({lo}); {var} <= ({hi}); ++{var})")
        em.secvars = secvars
        tgt, _ = em.index(sym, subs)
        c, t = em.expr(rhs)
        em.secvars = None
        self.emit(" ".join(reversed(loops)) + f" {tgt} = {self.convert(c, t, sym.typ, em)};")

    @staticmethod
    def convert(code, t, target, em):
        if t == target:
            return code
        if target == "real" and t == "int":
            return f"(float)({code})"
        if target == "int" and t == "real":
            return f"(int)({code})"   # truncation toward zero, as Fortran's intrinsic assignment
        em.err(f"cannot assign {t} to {target}")


# ------------------------------------------------------------------------------------------
# driver
# ------------------------------------------------------------------------------------------
def collect_declarations(lines, fname, scope, prefix):
    for no, stmt in lines:
        if DECL_RE.match(stmt) and ("::" in stmt or re.match(r"^(real|integer|logical)\s+[a-z_]", stmt)):
            parse_declaration(stmt, f"{fname}:{no}", scope, prefix)


def member_declarations(scope: Scope):
    """C++ member declarations for every symbol of a scope, in declaration order."""
    decl, ctor = [], []
    for s in scope.syms.values():
        em = Emitter(scope, s.where)
        if s.param is not None:
            c, t = em.expr(s.param)
            decl.append(f"  const {CTYPE[s.typ]} {s.cname} = {Unit.convert(c, t, s.typ, em)}; /* {s.where} */")
        elif s.rank == 0:
            decl.append(f"  {CTYPE[s.typ]} {s.cname} = 0; /* {s.where} */")
        else:
            decl.append(f"  FArr<{CTYPE[s.typ]}, {s.rank}> {s.cname}{{\"{s.name}\"}}; /* {s.where} */")
            if s.dims is not None:
                los = [em.expr(lo)[0] if lo is not None else "1" for lo, _ in s.dims]
                his = [em.expr(hi)[0] for _, hi in s.dims]
                ctor.append(f"    {s.cname}.alloc({{{', '.join(los)}}}, {{{', '.join(his)}}});")
    return decl, ctor


def main(src_dir: str, out_path: str):
    import os
    scopes = {}
    decl_lines, ctor_lines, bodies, protos = [], [], [], []
    manifest = []
    # modules
    mods = []
    for mname, fname in MODULES:
        path = os.path.join(src_dir, fname)
        lines = logical_lines(path)
        sc = Scope(mname, parents=list(mods))
        body = [(no, s) for no, s in lines
                if not re.match(r"^(module|end\s*module|use|implicit|save)\b", s)]
        for no, s in body:
            if not DECL_RE.match(s):
                raise TranslateError(f"{fname}:{no}: executable statement in a module: '{s}'")
        collect_declarations(body, fname, sc, "")
        d, c = member_declarations(sc)
        decl_lines += [f"  /* ---- MODULE {mname.upper()} ({fname}) ---- */"] + d
        ctor_lines += c
        mods.append(sc)
        scopes[mname] = sc
        manifest.append(f"{fname}: whole module ({len(sc.syms)} entities, {len(sc.dropped)} dropped)")
    # PROGRAM H9 locals
    pname, pfile, p0, p1 = PROGRAM_DECLS
    psc = Scope(pname, parents=mods)
    collect_declarations(logical_lines(os.path.join(src_dir, pfile), p0, p1), pfile, psc, "h9__")
    d, c = member_declarations(psc)
    decl_lines += [f"  /* ---- locals of PROGRAM H9 ({pfile}:{p0}-{p1}) ---- */"] + d
    ctor_lines += c
    scopes[pname] = psc
    # INIT locals (only `decay` is used by the translated ranges)
    isc = Scope("init", parents=mods)
    init_lines = logical_lines(os.path.join(src_dir, "INIT.f90"), 16, 20)
    collect_declarations(init_lines, "INIT.f90", isc, "init__")
    d, c = member_declarations(isc)
    decl_lines += ["  /* ---- locals of SUBROUTINE INIT (INIT.f90:16-20) ---- */"] + d
    ctor_lines += c
    scopes["init"] = isc
    # subroutines
    for sname, fname in SUBROUTINES:
        path = os.path.join(src_dir, fname)
        lines = logical_lines(path)
        sc = Scope(sname, parents=mods)
        collect_declarations(lines, fname, sc, sname + "__")
        d, c = member_declarations(sc)
        decl_lines += [f"  /* ---- locals of SUBROUTINE {sname.upper()} ({fname}) ---- */"] + d
        ctor_lines += c
        u = Unit(sname, fname, sc, lines)
        u.sync = f"F2C_SYNC_{sname}"
        body = u.translate()
        protos.append(f"  void {sname}();")
        # -DF2C_LOCALS: the subroutine's scalar locals are automatic variables (as in Fortran, so
        # the compiler may keep them in registers) and are written back to the members of the
        # same name on exit and before STOP, for the harness's diagnostics.  Default: members.
        scal = [v for v in sc.syms.values() if v.rank == 0 and v.param is None]
        loc = "".join(f"  {CTYPE[v.typ]} {v.cname} = 0;\n" for v in scal)
        wb = " ".join(f"this->{v.cname} = {v.cname};" for v in scal)
        pre = (f"#ifdef F2C_LOCALS\n{loc}#define {u.sync} do {{ {wb} }} while (0)\n"
               f"#else\n#define {u.sync} do {{ }} while (0)\n#endif\n")
        bodies.append(f"void H9Ref::{sname}() {{ /* {fname}, whole file, {u.nstop} STOP statements */\n"
                      + pre + "\n".join(body) + f"\n  {u.sync};\n}}\n#undef {u.sync}\n")
        manifest.append(f"{fname}: SUBROUTINE {sname.upper()}, {len(lines)} statements")
    # fragments
    for name, fname, a, b, scname in FRAGMENTS:
        lines = logical_lines(os.path.join(src_dir, fname), a, b)
        u = Unit(name, fname, scopes[scname], lines)
        body = u.translate(stop_base=100)
        protos.append(f"  void {name}(); /* {fname}:{a}-{b} */")
        bodies.append(f"void H9Ref::{name}() {{ /* {fname}:{a}-{b} */\n" + "\n".join(body) + "\n}\n")
        manifest.append(f"{fname}:{a}-{b} -> {name}(), {len(lines)} statements")
    with open(out_path, "w") as f:
        f.write("/* GENERATED by oracle/f2cpp.py from the reference's Fortran sources -- do not edit,\n"
                " * do not commit (oracle/_ref/ is git-ignored).  Translated units:\n")
        for m in manifest:
            f.write(f" *   {m}\n")
        f.write(" */\n#pragma once\n#include \"../f2c_rt.h\"\n\nstruct H9Ref {\n")
        f.write("\n".join(decl_lines))
        f.write("\n\n  H9Ref() {\n" + "\n".join(ctor_lines) + "\n  }\n")
        f.write("  [[noreturn]] void f2c_stop(int ordinal, int line);\n")
        f.write("\n".join(protos))
        f.write("\n};\n\n")
        f.write("\n".join(bodies))
    return manifest


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    try:
        for line in main(sys.argv[1], sys.argv[2]):
            print("f2cpp:", line)
    except TranslateError as ex:
        sys.exit(f"f2cpp: {ex}")
