"""The Fortran-to-Python translator (oracle/f90py.py) on small hand-written units: each case isolates one rule of Fortran semantics
that the parity pin of tests/test_reference_source.py relies on.  No reference source needed."""
import math

import numpy as np
import pytest

from oracle import f90py as F


def run(src, name, *args, consts=None):
    P = F.Program(dict({"dp": 8}, **(consts or {})))
    P.add(src, [name])
    return getattr(P, name)(*args)


def test_integer_division_truncates_and_mixed_division_promotes():
    src = """
    SUBROUTINE f( i, j, x, q, r, s)
      INTEGER,  INTENT(IN)  :: i, j
      REAL(dp), INTENT(IN)  :: x
      INTEGER,  INTENT(OUT) :: q
      REAL(dp), INTENT(OUT) :: r, s
      q = i / j
      r = x / j
      s = REAL(i,dp) / REAL(j,dp)
    END SUBROUTINE f
    """
    assert run(src, "f", 7, 2, np.float64(7.0), 0, 0.0, 0.0) == (3, 3.5, 3.5)
    assert run(src, "f", -7, 2, np.float64(7.0), 0, 0.0, 0.0)[0] == -3          # towards zero, not floor


def test_kindless_real_literals_and_REAL_are_single_precision():
    src = """
    FUNCTION f( x) RESULT( y)
      REAL(dp), INTENT(IN) :: x
      REAL(dp)             :: y
      y = x - 1E-09 + 0.1 - 0.1_dp + REAL(16777217)
    END FUNCTION f
    """
    want = np.float64(0.0) - np.float64(np.float32(1e-9)) + np.float64(np.float32(0.1)) - np.float64(0.1) + np.float64(np.float32(16777217))
    assert run(src, "f", np.float64(0.0)) == want
    assert np.float64(np.float32(16777217)) == 16777216.0                          # the partition_list defect in one line


def test_power_rules():
    src = """
    SUBROUTINE f( x, a, b, c)
      REAL(dp), INTENT(IN)  :: x
      REAL(dp), INTENT(OUT) :: a, b, c
      a = x**3
      b = x**3.0_dp
      c = x**(-1._dp / 3._dp) * 2**3
    END SUBROUTINE f
    """
    x = np.float64(1.1)
    a, b, c = run(src, "f", x, 0.0, 0.0, 0.0)
    assert a == (x * x) * x                                                       # integer exponent: multiplication chain (GCC powi)
    assert b == math.pow(1.1, 3.0)                                                # real exponent: libm pow
    assert c == np.float64(math.pow(1.1, -1.0 / 3.0)) * 8


def test_sum_runs_in_array_element_order_and_norm2_is_libgfortrans():
    src = """
    SUBROUTINE f( v, s, t, n)
      REAL(dp), DIMENSION(:,:), INTENT(IN)  :: v
      REAL(dp),                 INTENT(OUT) :: s, t, n
      s = SUM( v)
      t = SUM( v(2,:))
      n = NORM2( v(1,1:2))
    END SUBROUTINE f
    """
    v = np.asfortranarray(np.array([[1e16, 3.0, -1e16], [1.0, 1.0, 1.0]]))
    s, t, n = run(src, "f", v, 0.0, 0.0, 0.0)
    seq = np.float64(0.0)
    for x in v.ravel(order="F"):                                                  # column-major: (1,1), (2,1), (1,2), ...
        seq = seq + x
    assert s == seq
    assert t == 3.0
    # libgfortran's scaled algorithm: for (1e16, 3) -> scale 1e16, result 1 + (3/1e16)^2
    val = np.float64(3.0) / np.float64(1e16)
    assert n == np.float64(1e16) * np.sqrt(np.float64(1.0) + val * val)
    assert F._norm2(np.array([3.0, 4.0])) == 5.0 and F._norm2(np.array([0.0, 0.0])) == 0.0


def test_sections_whole_array_assignment_lower_bounds_and_dummies_start_at_one():
    src = """
    SUBROUTINE f( n, a, tot)
      INTEGER,                  INTENT(IN)    :: n
      REAL(dp), DIMENSION(:),   INTENT(INOUT) :: a
      REAL(dp),                 INTENT(OUT)   :: tot
      REAL(dp), DIMENSION(2:n)  :: w
      REAL(dp), DIMENSION(n)    :: z
      INTEGER :: k
      z = 1._dp
      DO k = 2, n
        w( k) = REAL( k, dp)
      END DO
      a( 2:n) = a( 2:n) + w * 2._dp
      a( 1)   = g( w)
      tot     = SUM( z) + SUM( a)
    END SUBROUTINE f
    FUNCTION g( d) RESULT( y)
      REAL(dp), DIMENSION(:), INTENT(IN) :: d
      REAL(dp) :: y
      y = d( 1)
    END FUNCTION g
    """
    P = F.Program({"dp": 8})
    P.add(src, ["g", "f"])
    a = np.zeros(4)
    (tot,) = P.f(4, a, 0.0)
    assert a.tolist() == [2.0, 4.0, 6.0, 8.0]          # w(2:4) = 2,3,4; inside g the dummy starts at 1, so d(1) is w(2)
    assert tot == 4.0 + 20.0
    with pytest.raises(IndexError):                    # the reference is built with -fbounds-check
        run("SUBROUTINE h( a)\n REAL(dp), DIMENSION(:), INTENT(INOUT) :: a\n a( 5) = 1._dp\nEND SUBROUTINE h", "h", np.zeros(4))


def test_control_flow_named_loops_cycle_exit_one_line_if_and_scalar_out_arguments():
    src = """
    SUBROUTINE f( n, cnt, last)
      INTEGER, INTENT(IN)  :: n
      INTEGER, INTENT(OUT) :: cnt, last
      INTEGER :: i
      LOGICAL :: done
      cnt  = 0
      done = .FALSE.
      i    = 0
      outer: DO WHILE ((.NOT. done) .AND. (i < 100))
        i = i + 1
        IF (MOD( i, 2) == 0) CYCLE
        IF (i > n) THEN
          done = .TRUE.
        ELSEIF (i == 7) THEN
          EXIT outer
        ELSE
          CALL bump( cnt, i)
        END IF
      END DO outer
      last = i
    END SUBROUTINE f
    SUBROUTINE bump( c, by)
      INTEGER, INTENT(INOUT) :: c
      INTEGER, INTENT(IN)    :: by
      c = c + by
    END SUBROUTINE bump
    """
    P = F.Program({"dp": 8})
    P.add(src, ["bump", "f"])
    assert P.f(20, 0, 0) == (1 + 3 + 5, 7)             # EXIT at i = 7
    assert P.f(4, 0, 0) == (1 + 3, 5)                  # stops through the flag at i = 5


def test_division_by_zero_gives_inf_and_nan_like_fortran_and_elemental_max():
    src = """
    SUBROUTINE f( a, b, x, y, m)
      REAL(dp), INTENT(IN)  :: a, b
      REAL(dp), INTENT(OUT) :: x, y
      REAL(dp), DIMENSION(3), INTENT(OUT) :: m
      x = a / b
      y = (a - a) / b
      m = MAX( 0._dp, [a, -a, b])
    END SUBROUTINE f
    """
    np.seterr(all="ignore")
    m = np.zeros(3)
    x, y = run(src, "f", np.float64(1.0), np.float64(0.0), 0.0, 0.0, m)
    assert np.isinf(x) and np.isnan(y) and m.tolist() == [1.0, 0.0, 0.0]


def test_unsupported_constructs_are_refused_not_guessed():
    for body in ("GOTO 10", "x = foo( 3)", "WHERE (a > 0) a = 1"):
        with pytest.raises(F.Unsupported):
            run(f"SUBROUTINE f( x)\n REAL(dp), INTENT(INOUT) :: x\n {body}\nEND SUBROUTINE f", "f", np.float64(1.0))
