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
