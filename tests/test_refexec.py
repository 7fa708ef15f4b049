"""The Fortran front end that produces the reference-source golden vectors (oracle/refexec.py), tested on its own:
small fixed-form sources written here, executed, and compared with what the Fortran standard says they compute.
(The sources below are this repository's own test programs, not reference code.)"""
import math

import numpy as np
import pytest

from oracle import refexec as rx


def _gen(tmp_path, src, externals=None):
    f = tmp_path / "t.f"
    f.write_text(src)
    lib = rx.Library()
    lib.add_file(str(f))
    return rx.CodeGen(lib, externals=externals)


def test_fixed_form_reader_continuations_comments_cpp(tmp_path):
    src = (
        "C     a comment line\n"
        "      SUBROUTINE S(a, b)\n"
        "      REAL(KIND=8), INTENT(IN) :: a\n"
        "      REAL(KIND=8), INTENT(OUT) :: b   ! trailing comment with a ' quote\n"
        "#ifdef NEVER\n"
        "      b = -1D0\n"
        "#else\n"
        + "      b = a + 1._8".ljust(72) + "&ignored past column 72\n" +
        "     2  + 2.5E0_8 +\n"
        "     &  0.5d0\n"
        "#endif\n"
        "      RETURN\n"
        "      END SUBROUTINE S\n")
    g = _gen(tmp_path, src)
    assert g.get("s")(1.0, 0.0) == (5.0,)


def test_integer_division_power_and_precedence(tmp_path):
    src = (
        "      SUBROUTINE S(i, j, r)\n"
        "      INTEGER, INTENT(OUT) :: i, j\n"
        "      REAL(KIND=8), INTENT(OUT) :: r\n"
        "      i = 7/2 + (-7)/2\n"
        "      j = 2**3**2\n"
        "      r = -2.D0**2 + 10/4*2.D0\n"
        "      END SUBROUTINE S\n")
    i, j, r = _gen(tmp_path, src).get("s")(0, 0, 0.0)
    assert (i, j) == (0, 512)            # truncation toward zero; ** is right associative
    assert r == -4.0 + 2 * 2.0           # unary minus binds weaker than **; 10/4 is integer division


def test_do_variable_after_the_loop_and_exit(tmp_path):
    src = (
        "      SUBROUTINE S(n, a, b, c)\n"
        "      INTEGER, INTENT(IN) :: n\n"
        "      INTEGER, INTENT(OUT) :: a, b, c\n"
        "      INTEGER i\n"
        "      DO i=1, n\n"
        "      END DO\n"
        "      a = i\n"
        "      DO i=1, n\n"
        "         IF (i .EQ. 3) EXIT\n"
        "      END DO\n"
        "      b = i\n"
        "      DO i=10, 1, -3\n"
        "      END DO\n"
        "      c = i\n"
        "      END SUBROUTINE S\n")
    assert _gen(tmp_path, src).get("s")(5, 0, 0, 0) == (6, 3, -2)


def test_sections_are_passed_by_reference_and_explicit_shape_dummies_reshape(tmp_path):
    src = (
        "      SUBROUTINE FILL(n, v, s)\n"
        "      INTEGER, INTENT(IN) :: n\n"
        "      REAL(KIND=8), INTENT(INOUT) :: v(n)\n"
        "      REAL(KIND=8), INTENT(IN) :: s\n"
        "      INTEGER i\n"
        "      DO i=1, n\n"
        "         v(i) = s*i\n"
        "      END DO\n"
        "      END SUBROUTINE FILL\n"
        "      SUBROUTINE S(u)\n"
        "      REAL(KIND=8), INTENT(INOUT) :: u(2,3,2)\n"
        "      u = 0D0\n"
        "      CALL FILL(6, u(:,:,2), 1D0)\n"
        "      CALL FILL(2, u(:,2,1), 10D0)\n"
        "      END SUBROUTINE S\n")
    u = np.full((2, 3, 2), np.nan, order="F")
    _gen(tmp_path, src).get("s")(u)
    assert np.array_equal(u[:, :, 1].reshape(-1, order="F"), np.arange(1.0, 7.0))     # column-major sequence association
    assert np.array_equal(u[:, 1, 0], [10.0, 20.0]) and u[0, 0, 0] == 0.0


def test_functions_scalar_out_arguments_and_optional(tmp_path):
    src = (
        "      FUNCTION F(x, y)\n"
        "      REAL(KIND=8), INTENT(IN) :: x\n"
        "      REAL(KIND=8), INTENT(IN), OPTIONAL :: y\n"
        "      REAL(KIND=8) F\n"
        "      F = x\n"
        "      IF (PRESENT(y)) F = F + y\n"
        "      RETURN\n"
        "      END FUNCTION F\n"
        "      SUBROUTINE SWAPADD(a, b, t)\n"
        "      REAL(KIND=8), INTENT(INOUT) :: a, b\n"
        "      REAL(KIND=8), INTENT(OUT) :: t\n"
        "      t = a\n"
        "      a = b\n"
        "      b = t\n"
        "      t = a + b\n"
        "      END SUBROUTINE SWAPADD\n"
        "      SUBROUTINE S(v, r)\n"
        "      REAL(KIND=8), INTENT(INOUT) :: v(3)\n"
        "      REAL(KIND=8), INTENT(OUT) :: r\n"
        "      REAL(KIND=8) F, t\n"
        "      CALL SWAPADD(v(1), v(3), t)\n"
        "      r = F(t) + F(1D0, 2D0)\n"
        "      END SUBROUTINE S\n")
    v = np.array([1.0, 5.0, 9.0])
    (r,) = _gen(tmp_path, src).get("s")(v, 0.0)
    assert np.array_equal(v, [9.0, 5.0, 1.0]) and r == 10.0 + 3.0


def test_select_case_internal_procedures_and_allocation(tmp_path):
    src = (
        "      SUBROUTINE S(k, out, n)\n"
        "      INTEGER, INTENT(IN) :: k\n"
        "      INTEGER, INTENT(OUT) :: out, n\n"
        "      INTEGER cnt\n"
        "      REAL(KIND=8), ALLOCATABLE :: w(:), y(:)\n"
        "      cnt = 0\n"
        "      CALL BUMP\n"
        "      CALL BUMP\n"
        "      SELECT CASE (k)\n"
        "      CASE (1)\n"
        "         out = 10\n"
        "      CASE (2:4, 7)\n"
        "         out = 20\n"
        "      CASE DEFAULT\n"
        "         out = 30\n"
        "      END SELECT\n"
        "      out = out + cnt\n"
        "      ALLOCATE(y(5))\n"
        "      y = 1D0\n"
        "      y = w(1:2)\n"
        "      n = SIZE(y) + SIZE(w)\n"
        "      RETURN\n"
        "      CONTAINS\n"
        "      SUBROUTINE BUMP\n"
        "      cnt = cnt + 1\n"
        "      IF (.NOT.ALLOCATED(w)) ALLOCATE(w(3))\n"
        "      w = cnt\n"
        "      END SUBROUTINE BUMP\n"
        "      END SUBROUTINE S\n")
    s = _gen(tmp_path, src).get("s")
    assert s(1, 0, 0) == (12, 5)            # y is re-allocated to the shape of w(1:2) on assignment
    assert s(3, 0, 0)[0] == 22 and s(7, 0, 0)[0] == 22 and s(9, 0, 0)[0] == 32


def test_derived_types_defaults_arrays_of_types_and_value_assignment(tmp_path):
    src = (
        "      MODULE M\n"
        "      INTEGER, PARAMETER :: three = 3, six = 2*three\n"
        "      TYPE inner\n"
        "         LOGICAL :: flag = .FALSE.\n"
        "         INTEGER :: n = three\n"
        "         REAL(KIND=8), ALLOCATABLE :: v(:)\n"
        "      END TYPE inner\n"
        "      TYPE outer\n"
        "         REAL(KIND=8) :: p(six)\n"
        "         TYPE(inner) one\n"
        "         TYPE(inner), ALLOCATABLE :: many(:)\n"
        "      END TYPE outer\n"
        "      END MODULE M\n"
        "      SUBROUTINE S(o, c, any1, any2)\n"
        "      TYPE(outer), INTENT(INOUT) :: o\n"
        "      TYPE(inner), INTENT(OUT) :: c\n"
        "      LOGICAL, INTENT(OUT) :: any1, any2\n"
        "      ALLOCATE(o%many(2), o%one%v(o%one%n))\n"
        "      o%one%v = 7D0\n"
        "      any1 = ANY(o%many%flag)\n"
        "      o%many(2)%flag = .TRUE.\n"
        "      any2 = ANY(o%many%flag)\n"
        "      o%many%n = 5\n"
        "      c = o%one\n"
        "      c%v(1) = -1D0\n"
        "      END SUBROUTINE S\n")
    g = _gen(tmp_path, src)
    o, c = g.rt.new("outer"), g.rt.new("inner")
    assert o.p.shape == (6,) and o.one.n == 3 and o.many is None
    any1, any2 = g.get("s")(o, c, False, False)
    assert (any1, any2) == (False, True) and [m.n for m in o.many] == [5, 5]
    assert o.one.v.tolist() == [7.0, 7.0, 7.0]          # c = o%one copied the value: c%v(1) did not alias


def test_uninitialised_reals_are_poisoned_and_lower_bounds(tmp_path):
    src = (
        "      SUBROUTINE S(r, q)\n"
        "      REAL(KIND=8), INTENT(OUT) :: r, q\n"
        "      REAL(KIND=8) a(0:2), never\n"
        "      a(0) = 1D0\n"
        "      a(2) = 3D0\n"
        "      r = a(0) + a(2)\n"
        "      q = never\n"
        "      END SUBROUTINE S\n")
    r, q = _gen(tmp_path, src).get("s")(0.0, 0.0)
    assert r == 4.0 and math.isnan(q)


def test_left_to_right_evaluation_is_kept(tmp_path):
    """a + b + c is (a + b) + c: no re-association (what makes the oracle comparison bit-exact)"""
    src = (
        "      SUBROUTINE S(a, b, c, r1, r2)\n"
        "      REAL(KIND=8), INTENT(IN) :: a, b, c\n"
        "      REAL(KIND=8), INTENT(OUT) :: r1, r2\n"
        "      r1 = a + b + c\n"
        "      r2 = a + (b + c)\n"
        "      END SUBROUTINE S\n")
    a, b, c = 1.0, 1e-16, 1e-16
    r1, r2 = _gen(tmp_path, src).get("s")(a, b, c, 0.0, 0.0)
    assert r1 == (a + b) + c and r2 == a + (b + c) and r1 != r2


def test_unknown_procedures_fail_only_when_reached(tmp_path):
    src = (
        "      SUBROUTINE S(k, r)\n"
        "      INTEGER, INTENT(IN) :: k\n"
        "      INTEGER, INTENT(OUT) :: r\n"
        "      r = 1\n"
        "      IF (k .GT. 0) CALL NOT_IN_THE_TREE(r)\n"
        "      END SUBROUTINE S\n")
    s = _gen(tmp_path, src).get("s")
    assert s(0, 0) == (1,)
    with pytest.raises(NotImplementedError):
        s(1, 0)


REF = "/root/reference/Code/Source"


@pytest.mark.skipif(not __import__("os").path.isdir(REF), reason="the reference tree is not on this box")
def test_reference_routines_with_known_answers():
    """routines of the reference whose results are known independently, executed from their source: Gaussian
    elimination (svFSILS/GE.f) against LAPACK, 3x3 inverse / determinant / trace (svFSI/MATFUN.f) against NumPy,
    the scalar product and the cross product of svFSI/UTIL.f"""
    import os
    lib = rx.Library()
    for h in ("FSILS_TYPEDEF.h", "FSILS_STRUCT.h"):
        lib.add_include(os.path.join(REF, "svFSILS", h))
    lib.add_file(os.path.join(REF, "svFSILS", "GE.f"))
    for f in ("CONSTS.f", "TYPEMOD.f", "UTIL.f", "MATFUN.f"):
        lib.add_file(os.path.join(REF, "svFSI", f))
    g = rx.CodeGen(lib)
    rng = np.random.default_rng(0)
    n = 7
    A = rng.standard_normal((n, n)) + n * np.eye(n)
    b = rng.standard_normal(n)
    B = b.copy()
    ok = g.get("ge")(n, n, np.asfortranarray(A), B)
    assert ok and np.allclose(B, np.linalg.solve(A, b), rtol=1e-12, atol=1e-14)
    M3 = rng.standard_normal((3, 3)) + 2 * np.eye(3)
    inv = g.get("mat_inv")(np.asfortranarray(M3), 3)
    assert np.allclose(inv, np.linalg.inv(M3), rtol=1e-12, atol=1e-14)
    assert abs(g.get("mat_det")(np.asfortranarray(M3), 3) - np.linalg.det(M3)) <= 1e-12 * abs(np.linalg.det(M3))
    assert abs(g.get("mat_trace")(np.asfortranarray(M3), 3) - np.trace(M3)) <= 1e-14
    u, v = rng.standard_normal(5), rng.standard_normal(5)
    assert abs(g.get("norms")(u, v) - float(u @ v)) <= 1e-14
    V = np.asfortranarray(rng.standard_normal((3, 2)))
    assert np.allclose(g.get("cross")(V), np.cross(V[:, 0], V[:, 1]), rtol=1e-14, atol=1e-15)


def test_do_while_cycle_array_intrinsics_and_odd_spellings(tmp_path):
    src = (
        "      SUBROUTINE S(a, n, tot, mx, dp, cnt, bits)\n"
        "      INTEGER, INTENT(IN) :: n\n"
        "      REAL(KIND=8), INTENT(IN) :: a(3,n)\n"
        "      REAL(KIND=8), INTENT(OUT) :: tot, mx, dp\n"
        "      INTEGER, INTENT(OUT) :: cnt, bits\n"
        "      REAL(KIND=8) Do(3)\n"
        "      INTEGER i\n"
        "      Do = 0D0\n"
        "      i = 0\n"
        "      cnt = 0\n"
        "      DO WHILE (i .LT. n)\n"
        "         i = i + 1\n"
        "         IF (MOD(i,2).EQ.0 .OR . a(1,i).LT.0D0) CYCLE\n"
        "         Do = Do + a(:,i)\n"
        "         cnt = cnt + 1\n"
        "      END DO\n"
        "      tot = SUM(Do)\n"
        "      mx = MAXVAL(ABS(a))\n"
        "      dp = DOT_PRODUCT(a(:,1), a(:,n))\n"
        "      bits = 0\n"
        "      bits = IBSET(bits, 3)\n"
        "      IF (BTEST(bits,3) .AND. .NOT.BTEST(bits,2)) bits = bits + 100\n"
        "      END SUBROUTINE S\n")
    a = np.asfortranarray(np.array([[1.0, 2.0, 3.0], [9.0, 9.0, 9.0], [-1.0, 5.0, 5.0], [4.0, 4.0, 4.0], [0.5, 0.25, -8.0]]).T)
    tot, mx, dp, cnt, bits = _gen(tmp_path, src).get("s")(a, 5, 0.0, 0.0, 0.0, 0, 0)
    assert cnt == 2 and tot == (1 + 2 + 3) + (0.5 + 0.25 - 8.0)        # columns 1 and 5 (3 is negative, 2 and 4 are even)
    assert mx == 9.0 and dp == 1 * 0.5 + 2 * 0.25 + 3 * (-8.0) and bits == 108


def test_generic_interfaces_resolve_by_rank_and_type(tmp_path):
    src = (
        "      MODULE M\n"
        "      INTERFACE TWICE\n"
        "         MODULE PROCEDURE TWICES, TWICEV, TWICEI\n"
        "      END INTERFACE TWICE\n"
        "      CONTAINS\n"
        "      FUNCTION TWICES(x)\n"
        "      REAL(KIND=8), INTENT(IN) :: x\n"
        "      REAL(KIND=8) TWICES\n"
        "      TWICES = 2D0*x\n"
        "      END FUNCTION TWICES\n"
        "      FUNCTION TWICEV(x)\n"
        "      REAL(KIND=8), INTENT(IN) :: x(:)\n"
        "      REAL(KIND=8) TWICEV\n"
        "      TWICEV = 2D0*SUM(x)\n"
        "      END FUNCTION TWICEV\n"
        "      FUNCTION TWICEI(i)\n"
        "      INTEGER, INTENT(IN) :: i\n"
        "      REAL(KIND=8) TWICEI\n"
        "      TWICEI = -2D0*i\n"
        "      END FUNCTION TWICEI\n"
        "      END MODULE M\n"
        "      SUBROUTINE S(v, r1, r2, r3)\n"
        "      REAL(KIND=8), INTENT(IN) :: v(3)\n"
        "      REAL(KIND=8), INTENT(OUT) :: r1, r2, r3\n"
        "      r1 = TWICE(v(2))\n"
        "      r2 = TWICE(v)\n"
        "      r3 = TWICE(7)\n"
        "      END SUBROUTINE S\n")
    r1, r2, r3 = _gen(tmp_path, src).get("s")(np.array([1.0, 2.0, 4.0]), 0.0, 0.0, 0.0)
    assert (r1, r2, r3) == (4.0, 14.0, -14.0)


def test_direct_access_write_is_captured_as_bytes(tmp_path):
    src = (
        "      SUBROUTINE S(a, n)\n"
        "      INTEGER, INTENT(IN) :: n\n"
        "      REAL(KIND=8), INTENT(IN) :: a(2,n)\n"
        "      LOGICAL flag\n"
        "      flag = .TRUE.\n"
        "      OPEN(27, FILE='x.bin', ACCESS='DIRECT', RECL=64)\n"
        "      WRITE(27, REC=n+1) n, flag, 2.5D0, a\n"
        "      CLOSE(27)\n"
        "      END SUBROUTINE S\n")
    g = _gen(tmp_path, src)
    a = np.asfortranarray(np.array([[1.0, 3.0], [2.0, 4.0]]))
    g.get("s")(a, 2)
    want = np.array([2, 1], dtype="<i4").tobytes() + np.array([2.5, 1.0, 2.0, 3.0, 4.0], dtype="<f8").tobytes()
    assert g.rt.files[27][3] == want          # list order, column-major array elements, 4-byte INTEGER / LOGICAL
