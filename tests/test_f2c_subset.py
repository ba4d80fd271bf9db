"""The Fortran-subset translator (oracle/f2c_subset.py) on small routines whose answers are computed by hand or by numpy:
what it does to NINT ties, integer division, real -> integer assignment, mixed-mode arithmetic, do loops (bounds evaluated
once, strides, labelled loops, go to / cycle / exit), array sections, vector subscripts, where / elsewhere, whole-array
assignments, allocatable work arrays, assumed-shape dummies, optional arguments, any / sum / size, functions, column-major
1-based indexing.  These are the semantics the pin of the oracle (tests/test_oracle_vs_ref.py) rests on; none of this
touches the reference tree."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import f2c_subset as F  # noqa: E402

SRC = """
      subroutine arith ( x, n, iout, rout )
      implicit none
      integer, intent(in) :: n
      real,    intent(in) :: x(n)
      integer, intent(inout) :: iout(n,4)
      real,    intent(inout) :: rout(n,3)
      integer i, m
      do i = 1, n
         iout(i,1) = nint ( x(i) )
         m = x(i)
         iout(i,2) = m
         iout(i,3) = (7 * i) / 2 - i / 3
         iout(i,4) = mod ( 7*i, 4 ) + abs ( 3 - i ) + min ( i, 3, 5 ) + max ( 1, i - 2 )
         rout(i,1) = m * x(i) / (m - 10)
         rout(i,2) = 1. / 3 + i / 2 + 2.5e-1 * i
         rout(i,3) = x(i)**2 - abs ( x(i) ) + sqrt ( abs ( x(i) ) )
      end do
      end subroutine arith

      subroutine loops ( n, a, trace, ntrace )
      implicit none
      integer, intent(in) :: n
      integer, intent(inout) :: a(n)
      integer, intent(inout) :: trace(100)
      integer, intent(out) :: ntrace
      integer i, j, k, hi
      ntrace = 0
      hi = 3
      do i = 1, hi
         hi = 10
         ntrace = ntrace + 1
         trace(ntrace) = i
      end do
      do 10 j = 2, n, 3
         if ( a(j) .lt. 0 ) go to 10
         do k = 1, n
            if ( k .eq. j ) cycle
            if ( k .gt. j + 1 ) exit
            a(k) = a(k) + j
         enddo
         ntrace = ntrace + 1
         trace(ntrace) = 100 + j
 10   continue
      ntrace = ntrace + 1
      trace(ntrace) = j
      end subroutine loops

      subroutine sections ( n, nr, a, b, idx, c, w, tot, anyneg )
      implicit none
      integer, intent(in) :: n, nr
      real, intent(inout) :: a(n,nr)
      real, intent(in)    :: b(n)
      integer, intent(in) :: idx(n)
      real, intent(inout) :: c(:,:)
      real, intent(inout) :: w(n)
      integer, intent(out) :: tot
      logical, intent(out) :: anyneg
      real, allocatable :: tmp(:,:)
      integer, allocatable, dimension(:) :: cnt
      logical, allocatable :: neg(:)
      allocate ( tmp(n,nr), cnt(n), neg(n) )
      tmp = 0.5
      cnt = 2; neg = .false.
      a(:,1) = b(1:n) * 2.
      a(1:n,2) = a(:,1) - b(idx(:))
      tmp(:,:) = a(idx(:),:) + tmp(:,:)
      c(2,:) = tmp(n,1:nr)
      c(:,1) = c(:,1) + 1.
      where ( b(1:n) > 0.0 )
         w(:) = 1.0 / b(1:n)
      elsewhere
         w(:) = -1.0
      endwhere
      neg(2) = b(2) < 0.0
      cnt(3) = size(c,1) + size(c,2) * 10 + size(tmp)
      tot = sum(cnt)
      anyneg = any(neg)
      a = tmp
      deallocate ( tmp, cnt, neg )
      end subroutine sections

      subroutine opt ( x, y, scale )
      implicit none
      real, intent(in) :: x
      real, intent(out) :: y
      integer, OPTIONAL, intent(in) :: scale
      integer s_
      s_ = 1
      if ( present(scale) ) then
           s_ = scale
      end if
      if ( x .gt. 0. .AND. x .ge. s_ ) then
         y = x / s_
      else if ( x == 0.0 ) then
         y = -1.0
      else
         y = -2.0
      end if
      end subroutine opt

      real function trimean ( x, lo, hi )
      real, intent(in) :: x(:)
      integer, intent(in) :: lo, hi
      integer i
      real acc
      acc = 0.0
      do i = lo, hi
         acc = acc + x(i)
      end do
      if ( hi .lt. lo ) then
         trimean = -999.
         return
      endif
      trimean = acc / (hi - lo + 1)
      end function trimean
"""


@pytest.fixture(scope="module")
def lib():
    tr = F.Translator(SRC, {})
    for nm in ("arith", "loops", "sections", "opt", "trimean"):
        tr.routine(nm, nm + "_")
    d = tempfile.mkdtemp(prefix="f2c_")
    csrc = os.path.join(d, "t.c")
    open(csrc, "w").write(F.PRELUDE + "\n".join(tr.out) + "\n")
    so = os.path.join(d, "t.so")
    subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unused-variable", "-Wno-unused-label",
                           "-o", so, csrc, "-lm"])
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_arithmetic_follows_fortran(lib):
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 2.4999999, 3.7, -3.7, 12.0])
    n = len(x)
    iout = np.zeros((n, 4), np.int32, order="F")
    rout = np.zeros((n, 3), order="F")
    lib.arith_(P(x), n, P(iout), P(rout))
    i = np.arange(1, n + 1)
    assert list(iout[:, 0]) == [1, 2, 3, -1, -2, -3, 2, 4, -4, 12]               # NINT: half away from zero
    m = np.trunc(x).astype(int)
    assert np.array_equal(iout[:, 1], m)                                          # real -> integer: toward zero
    assert np.array_equal(iout[:, 2], (7 * i) // 2 - i // 3)                      # integer division
    assert np.array_equal(iout[:, 3], (7 * i) % 4 + np.abs(3 - i) + np.minimum(i, 3) + np.maximum(1, i - 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(rout[:, 0], (m * x) / (m - 10))                     # left to right, integer promoted
    assert np.array_equal(rout[:, 1], 1.0 / 3 + i // 2 + 2.5e-1 * i)              # i / 2 stays an integer division
    assert np.array_equal(rout[:, 2], x * x - np.abs(x) + np.sqrt(np.abs(x)))


def test_loops(lib):
    n = 9
    a = np.array([0, 1, 0, 0, -5, 0, 0, 3, 0], np.int32)
    trace = np.zeros(100, np.int32)
    nt = C.c_int(0)
    lib.loops_(n, P(a), P(trace), C.byref(nt))
    t = list(trace[:nt.value])
    # the bound of the first loop is evaluated once (3 trips although hi becomes 10); j = 2, 5, 8: j = 5 is skipped by the
    # go to; for j the inner loop adds j to a(1..j+1) except a(j); a do variable ends one stride past the last trip
    assert t == [1, 2, 3, 102, 108, 11]
    exp = np.array([0, 1, 0, 0, -5, 0, 0, 3, 0])
    for j in (2, 8):
        for k in range(1, n + 1):
            if k == j:
                continue
            if k > j + 1:
                break
            exp[k - 1] += j
    assert np.array_equal(a, exp)


def test_sections_vector_subscripts_where_allocatables(lib):
    rng = np.random.default_rng(0)
    n, nr = 7, 2
    a = np.asfortranarray(rng.normal(size=(n, nr)))
    b = rng.normal(size=n)
    b[2] = 0.0
    idx = rng.permutation(n).astype(np.int32) + 1
    c = np.asfortranarray(rng.normal(size=(3, nr)))
    c0 = c.copy()
    w = np.zeros(n)
    tot, anyneg = C.c_int(0), C.c_int(0)
    lib.sections_(n, nr, P(a), P(b), P(idx), P(c), c.shape[0], c.shape[1], P(w), C.byref(tot), C.byref(anyneg))
    a1 = b * 2.0
    a2 = a1 - b[idx - 1]
    tmp = np.stack([a1, a2], axis=1)[idx - 1, :] + 0.5
    assert np.array_equal(a, tmp)                                                 # a = tmp at the end
    cexp = c0.copy()
    cexp[1, :] = tmp[n - 1, :]
    cexp[:, 0] += 1.0
    assert np.array_equal(c, cexp)
    with np.errstate(divide="ignore"):
        assert np.array_equal(w, np.where(b > 0.0, 1.0 / b, -1.0))
    assert tot.value == 2 * (n - 1) + (3 + nr * 10 + n * nr)
    assert anyneg.value == int(b[1] < 0.0)


def test_optional_argument_and_else_if(lib):
    y = C.c_double(0.0)
    lib.opt_.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.opt_(6.0, C.byref(y), None)
    assert y.value == 6.0
    lib.opt_(6.0, C.byref(y), C.byref(C.c_int(4)))
    assert y.value == 1.5
    lib.opt_(3.0, C.byref(y), C.byref(C.c_int(4)))
    assert y.value == -2.0
    lib.opt_(0.0, C.byref(y), None)
    assert y.value == -1.0


def test_function_result_and_early_return(lib):
    x = np.array([1.0, 2.0, 4.0, 8.0])
    lib.trimean_.restype = C.c_double
    lib.trimean_.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    assert lib.trimean_(P(x), 4, 2, 4) == (2.0 + 4.0 + 8.0) / 3
    assert lib.trimean_(P(x), 4, 3, 2) == -999.0


def test_fixed_form_lines():
    text = "\n".join([
        "c     a comment",
        "      x = 1 ! trailing comment",
        "      if ( a .GT. b .and.",
        "     &     c .Le .d ) y = 'it''s ! not a comment'",
        " 10   continue",
        "      i = 1; j = 2",
    ])
    ll = F.logical_lines(text, 1, 6)
    assert ll[0] == (None, "x = 1", 2)
    assert ll[1][1] == "if ( a .gt. b .and. c .le.d ) y = 'it''s ! not a comment'"
    assert ll[2] == ("10", "continue", 5)
    assert [s for _, s, _ in ll[3:]] == ["i = 1", "j = 2"]
