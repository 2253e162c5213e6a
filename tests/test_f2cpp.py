"""Unit tests of oracle/f2cpp.py, the Fortran-subset -> C++ translator that builds the pin
(oracle/_ref).  Small Fortran fragments whose values follow from the Fortran standard are
translated, compiled with g++ and run; what is checked is exactly what a wrong translation of
HYDROLOGY.f90 would get wrong silently: operator precedence and associativity (`**` above unary
minus and right-associative, a leading sign covering the whole first term), integer division
and int/real promotion, real**integer by multiplication, DO semantics (bounds evaluated once,
negative step, value of the index after the loop, EXIT from a named construct), array sections
with explicit lower bounds, MIN/MAX/SUM/MINVAL, logical IF, and that constructs outside the
subset are refused instead of guessed at."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import f2cpp  # noqa: E402

DECLS = """
real :: a, b, c, r1, r2, r3, r4, r5, r6, r7, r8
integer :: i, j, k, n, m, i1, i2, i3, i4, i5, i6
real :: v (0:4), w (3)
real, allocatable :: q (:,:)
real, parameter :: two = 2.0
logical :: flag
"""


def translate(body: str):
    """-> C++ source of a program that prints the listed variables"""
    scope = f2cpp.Scope("t")
    lines = [(n + 1, f2cpp.fold_case(s.strip())) for n, s in enumerate(DECLS.strip().split("\n"))]
    f2cpp.collect_declarations(lines, "decl", scope, "")
    decl, ctor = f2cpp.member_declarations(scope)
    stmts = []
    for n, raw in enumerate(body.strip().split("\n")):
        code = f2cpp.strip_comment(raw, None)[0].strip()
        if code:
            stmts.append((100 + n, f2cpp.fold_case(code)))
    unit = f2cpp.Unit("run", "snippet", scope, stmts)
    out = unit.translate()
    return ("#include \"%s\"\n#include <cstdio>\nstruct T {\n%s\n  T() {\n%s\n  }\n"
            "  [[noreturn]] void f2c_stop(int o, int) { std::printf(\"STOP %%d\\n\", o); std::exit(0); }\n"
            "  void run() {\n%s\n  }\n};\n" % (os.path.join(ROOT, "oracle", "f2c_rt.h"), "\n".join(decl),
                                             "\n".join(ctor), "\n".join(out)))


def run(body: str, show: str):
    src = translate(body) + "int main() { T t; t.run(); std::printf(\"" + \
        " ".join("%.9g" for _ in show.split()) + "\\n\", " + \
        ", ".join(f"(double)t.{v}_" for v in show.split()) + "); }\n"
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DF2C_BOUNDS", "-Wno-unused-label",
                        "-o", exe, cpp], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip()
    return out if out.startswith("STOP") else [float(x) for x in out.split()]


def test_precedence_and_associativity():
    got = run("""
        a = 3.0
        b = 2.0
        r1 = -a ** 2              ! -(a**2) = -9
        r2 = 2.0 ** 3 ** 2        ! 2**(3**2) = 512
        r3 = -a * b + 1.0         ! (-(a*b)) + 1 = -5
        r4 = a - b - 1.0          ! (a-b)-1 = 0
        r5 = a / b / 4.0          ! (a/b)/4 = 0.375
        r6 = -1.0 / b             ! -(1/b) = -0.5
    """, "r1 r2 r3 r4 r5 r6")
    assert got == [-9.0, 512.0, -5.0, 0.0, 0.375, -0.5]


def test_integer_division_promotion_and_conversion():
    got = run("""
        i1 = 7 / 2                ! 3
        i2 = -7 / 2               ! -3 (truncation toward zero)
        r1 = 7 / 2                ! integer division first, then conversion: 3.0
        r2 = 7 / 2.0              ! 3.5
        r3 = 1 / 3 * 3.0          ! (1/3 = 0) * 3.0 = 0
        i3 = 2.9                  ! truncation on assignment: 2
        i4 = -2.9                 ! -2
        r4 = FLOAT (7) / FLOAT (2)
        i5 = MOD (-7, 3)          ! -1: sign of the dividend
        i6 = NINT (2.5) + NINT (-2.5)   ! 3 + (-3)
    """, "i1 i2 r1 r2 r3 i3 i4 r4 i5 i6")
    assert got == [3, -3, 3.0, 3.5, 0.0, 2, -2, 3.5, -1, 0]


def test_real_to_integer_power_is_multiplication():
    import numpy as np
    got = run("""
        a = 1.1
        r1 = a ** 2
        r2 = a ** 4
        r3 = a ** 3
        r4 = a ** (-2)
        r5 = a ** 2.0
        i1 = 3 ** 4
    """, "r1 r2 r3 r4 r5 i1")
    x = np.float32(1.1)
    x2 = np.float32(x * x)
    got32 = [np.float32(v) for v in got]  # %.9g round-trips a float32
    assert got32[0] == x2
    assert got32[1] == np.float32(x2 * x2)            # (x*x)*(x*x), as libgcc's __powisf2
    assert got32[2] == np.float32(x * x2)
    assert got32[3] == np.float32(np.float32(1.0) / x2)
    assert abs(got[4] - 1.21) < 1e-6 and got[5] == 81


def test_do_loop_semantics():
    got = run("""
        n = 3
        k = 0
        DO i = 1, n
          n = 10                  ! the bound was evaluated once: still three trips
          k = k + 1
        END DO
        i1 = i                    ! 4: one past the last value
        m = 0
        DO j = 5, 1, -2           ! 5, 3, 1
          m = m + j
        END DO
        i2 = j                    ! -1
        i3 = 0
        DO j = 3, 1               ! zero trips
          i3 = 99
        END DO
        i4 = 0
        outer: DO i = 1, 5
          DO j = 1, 5
            i4 = i4 + 1
            IF (i * j >= 6) EXIT outer
          END DO
        END DO outer
        i5 = i
        i6 = j
    """, "k i1 m i2 i3 i4 i5 i6")
    assert got == [3, 4, 9, -1, 0, 8, 2, 3]


def test_arrays_sections_and_reductions():
    got = run("""
        v (:) = 1.5               ! whole array with lower bound 0
        v (0) = -2.0
        v (4) = 7.0
        w (:) = v (1:3) * two     ! section assignment, element by element
        r1 = SUM (v (:))
        r2 = MINVAL (v (1:4))
        r3 = MAX (v (0), w (1), 0.25)
        r4 = MIN (3, 2) + MAX (1, 2, 5)
        ALLOCATE (q (2, 0:1))
        q (:,:) = 0.0
        q (2, 1) = 4.0
        q (:, 0) = q (:, 1) + 1.0
        r5 = q (1, 0) + 10.0 * q (2, 0)
        r6 = SUM (q (2, :))
    """, "r1 r2 r3 r4 r5 r6")
    assert got == [-2.0 + 1.5 * 3 + 7.0, 1.5, 3.0, 7.0, 51.0, 9.0]


def test_logical_if_and_stop():
    got = run("""
        a = 2.0
        flag = (a > 1.0) .AND. .NOT. (a >= 3.0) .OR. (a /= a)
        r1 = 0.0
        IF (flag) r1 = 1.0
        IF (a .LT. 1.0) THEN
          r2 = 1.0
        ELSE IF (a <= 2.0) THEN
          r2 = 2.0
        ELSE
          r2 = 3.0
        END IF
        IF (r2 == 2.0) THEN
          WRITE (*,*) 'dropped', r2
          STOP
        END IF
        r1 = 5.0
    """, "r1 r2")
    assert got == "STOP 1"


def test_bounds_checked_runtime_catches_an_out_of_range_subscript():
    with pytest.raises(subprocess.CalledProcessError):
        run("""
            i = 5
            v (i) = 1.0
        """, "a")


@pytest.mark.parametrize("stmt,why", [
    ("r1 = undeclared_thing + 1.0", "undeclared"),
    ("r1 = v", "whole-array reference"),
    ("CALL something_else (a)", "outside the translated set"),
    ("r1 = 1.0d0", "double-precision"),
    ("i1 = i2 .AND. i3", "non-logical"),
    ("GO TO 10", "unsupported statement"),
    ("two = 3.0", "PARAMETER"),
    ("v (1, 2) = 0.0", "rank mismatch"),
])
def test_outside_the_subset_is_refused(stmt, why):
    with pytest.raises(f2cpp.TranslateError) as e:
        translate(stmt)
    assert why.split()[0].lower() in str(e.value).lower(), str(e.value)


# ---- randomised expressions: the translator against an independent evaluator of the standard's rules ----

class _Gen:
    """Random expression trees over REAL a,b,c and INTEGER i,j,k.  `text` prints a tree with the
    FEWEST parentheses Fortran's precedence allows, so the translator's parser has to rebuild the
    tree; `value` evaluates the tree itself with the standard's typing (int/int truncates, mixed
    operands promote to REAL, every REAL operation rounds to float32)."""
    PREC = {"**": 5, "*": 4, "/": 4, "neg": 3, "+": 2, "-": 2}

    def __init__(self, rng, env):
        self.rng, self.env = rng, env

    def tree(self, depth):
        r = self.rng
        if depth == 0 or r.random() < 0.25:
            kind = r.integers(0, 4)
            if kind == 0:
                return ("var", str(r.choice(["a", "b", "c"])))
            if kind == 1:
                return ("var", str(r.choice(["i", "j", "k"])))
            if kind == 2:
                return ("rlit", str(r.choice(["0.5", "2.0", "1.25", "3.0E-1", "7.5"])))
            return ("ilit", str(int(r.integers(1, 6))))
        op = str(r.choice(["+", "-", "*", "/", "**", "neg", "max", "min", "abs"]))
        if op == "neg":
            return ("neg", self.tree(depth - 1))
        if op == "abs":
            return ("abs", self.tree(depth - 1))
        if op in ("max", "min"):
            return (op, self.tree(depth - 1), self.tree(depth - 1))
        if op == "**":
            return ("**", self.tree(depth - 1), ("ilit", str(int(r.integers(2, 4)))))
        return (op, self.tree(depth - 1), self.tree(depth - 1))

    def text(self, t, parent=0, right=False, first_term=True):
        k = t[0]
        if k == "var" or k in ("rlit", "ilit"):
            return t[1]
        if k in ("max", "min"):
            return f"{k} ({self.text(t[1])}, {self.text(t[2])})"
        if k == "abs":
            return f"abs ({self.text(t[1])})"
        if k == "neg":
            # a sign is only legal at the start of a level-2 expression: parenthesise elsewhere
            inner = "-" + self.text(t[1], self.PREC["*"], False)
            return inner if (parent <= self.PREC["+"] and first_term and not right) else f"({inner})"
        p = self.PREC[k]
        if k == "**":
            s = f"{self.text(t[1], p + 1)} ** {self.text(t[2], p, True)}"
        else:
            s = f"{self.text(t[1], p, False, first_term)} {k} {self.text(t[2], p + (0 if k in '+*' and False else 1), True, False)}"
        need = p < parent or (p == parent and right)
        return f"({s})" if need else s

    def value(self, t):
        import numpy as np
        k = t[0]
        if k == "var":
            return self.env[t[1]]
        if k == "rlit":
            return np.float32(float(t[1]))
        if k == "ilit":
            return int(t[1])
        if k == "neg":
            v = self.value(t[1])
            return -v
        if k == "abs":
            v = self.value(t[1])
            return abs(v) if isinstance(v, int) else np.float32(abs(v))
        a, b = self.value(t[1]), self.value(t[2])
        if k == "**":
            n = b
            if isinstance(a, int):
                return a ** n
            y = a if n % 2 else np.float32(1.0)
            x = a
            n >>= 1
            while n:
                x = np.float32(x * x)
                if n % 2:
                    y = np.float32(y * x)
                n >>= 1
            return y
        both_int = isinstance(a, int) and isinstance(b, int)
        if not both_int:
            a, b = np.float32(a), np.float32(b)
        if k in ("max", "min"):
            return (max if k == "max" else min)(a, b)
        if k == "+":
            return a + b if both_int else np.float32(a + b)
        if k == "-":
            return a - b if both_int else np.float32(a - b)
        if k == "*":
            return a * b if both_int else np.float32(a * b)
        if both_int:
            if b == 0:
                raise ZeroDivisionError
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q
        if b == 0:
            raise ZeroDivisionError
        return np.float32(a / b)


def test_random_expressions_follow_the_standard():
    import numpy as np
    rng = np.random.default_rng(2024)
    env = {"a": np.float32(1.75), "b": np.float32(-0.625), "c": np.float32(3.5), "i": 7, "j": -3, "k": 2}
    gen = _Gen(rng, env)
    cases = []
    while len(cases) < 250:
        t = gen.tree(4)
        try:
            with np.errstate(all="raise"):
                v = gen.value(t)
        except (ZeroDivisionError, FloatingPointError, OverflowError):
            continue
        if isinstance(v, int):
            if abs(v) > 10 ** 8:
                continue
        elif not np.isfinite(v) or abs(v) > 1e30:
            continue
        cases.append((gen.text(t), v))
    # one program: r(n) = expression n (REAL target: integer results are converted on assignment)
    decl = "real :: a, b, c\ninteger :: i, j, k\nreal :: r (%d)\n" % len(cases)
    body = "a = 1.75\nb = -0.625\nc = 3.5\ni = 7\nj = -3\nk = 2\n" + \
        "\n".join(f"r ({n + 1}) = {txt}" for n, (txt, _) in enumerate(cases))
    scope = f2cpp.Scope("t")
    f2cpp.collect_declarations([(n + 1, f2cpp.fold_case(s)) for n, s in enumerate(decl.strip().split("\n"))],
                               "decl", scope, "")
    d, ctor = f2cpp.member_declarations(scope)
    unit = f2cpp.Unit("run", "snippet", scope, [(100 + n, f2cpp.fold_case(s)) for n, s in enumerate(body.split("\n"))])
    src = ("#include \"%s\"\n#include <cstdio>\nstruct T {\n%s\n  T() {\n%s\n  }\n  [[noreturn]] void f2c_stop(int, int) "
           "{ std::exit(1); }\n  void run() {\n%s\n  }\n};\nint main() { T t; t.run(); for (int n = 1; n <= %d; ++n) "
           "std::printf(\"%%.9g\\n\", (double)t.r_(n)); }\n"
           % (os.path.join(ROOT, "oracle", "f2c_rt.h"), "\n".join(d), "\n".join(ctor), "\n".join(unit.translate()),
              len(cases)))
    with tempfile.TemporaryDirectory() as tmp:
        cpp, exe = os.path.join(tmp, "t.cpp"), os.path.join(tmp, "t")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DF2C_BOUNDS", "-o", exe, cpp], check=True)
        got = [np.float32(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert len(got) == len(cases)
    bad = [(txt, float(np.float32(v)), float(g)) for (txt, v), g in zip(cases, got) if np.float32(v) != g]
    assert not bad, bad[:5]
