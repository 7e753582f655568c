"""Who checks the checker: the Fortran-subset translator (oracle/f90_translate.py + f90_runtime.py) on small programs of
our own whose results are known from the language rules -- literal kinds and mixed-mode promotion, integer division,
loop forms, derived types, argument association, generic interfaces, list-directed and formatted input.  None of these
snippets comes from the reference; the pin tests (test_f90_pin_cpu.py) run the reference's files through the same code."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import f90_translate  # noqa: E402

F32, F64 = np.float32, np.float64


def load(src):
    prog = f90_translate.Program()
    prog.add_source(src, "snippet.f90")
    ns = prog.load()
    assert not ns["_NOT_TRANSLATED"], ns["_NOT_TRANSLATED"]
    return ns


def test_literal_kinds_and_mixed_mode_arithmetic():
    ns = load("""
module m
contains
  subroutine s(a, b, c, d, e, f, g, h)
    double precision :: a, b, c, d, e, f, g, h
    integer :: n
    real :: r
    n = 3
    a = 0.1              ! single-precision literal widened to double
    b = 0.1d0
    c = 1/3              ! integer division
    d = 1.0/3            ! single-precision quotient, then widened
    e = 2*a + n          ! integer operands promoted to double
    f = (0.001/24) * 2.0d0
    r = 16777216.0
    r = r + 1.0          ! stays single: 2**24 + 1 is not representable
    g = r
    h = 7 * (b - a) / n  ! int*double/int
  end subroutine
end module
""")
    a, b, c, d, e, f, g, h = ns["f_s"](*[F64(0)] * 8)
    assert a == float(F32(0.1)) and a != 0.1 and b == 0.1
    assert c == 0.0 and d == float(F32(1.0) / F32(3)) and d != 1.0 / 3
    assert e == 2 * float(F32(0.1)) + 3
    assert f == float(F32(0.001) / F32(24)) * 2.0
    assert g == 16777216.0
    assert h == 7 * (0.1 - float(F32(0.1))) / 3
    assert all(isinstance(x, np.float64) for x in (a, b, c, d, e, f, g, h))


def test_integer_semantics_and_the_odd_literal():
    ns = load("""
subroutine s(q1, q2, m1, m2, i1, i2, n1, n2, f1, c1, hit, p)
  integer :: q1, q2, m1, m2, i1, i2, n1, n2, f1, c1, hit, p
  double precision :: x
  q1 = (0 - 7) / 2        ! truncation toward zero, not floor
  q2 = 7 / 2
  m1 = mod(0 - 7, 3)      ! sign of the dividend
  m2 = mod(7, 0 - 3)
  i1 = 2.7
  i2 = 0.0d0 - 2.7d0      ! conversion truncates toward zero
  n1 = nint(2.5d0)
  n2 = nint(0.0d0 - 2.5d0)
  f1 = floor(0.0d0 - 0.5d0)
  c1 = ceiling(0.5d0)
  x = -0.5d0
  hit = 0
  if (x.gt.0.0000000001.or.x.lt. - 0000000001) hit = 1     ! the lower bound is the INTEGER -1
  p = 2**10
end subroutine
""")
    r = ns["f_s"](*[0] * 12)
    assert r == (-3, 3, -1, 1, 2, -2, 3, -3, -1, 1, 0, 1024)


def test_loops_labels_exit_cycle_while():
    ns = load("""
subroutine s(total, last, w)
  integer :: total, last, w
  integer :: i, j
  total = 0
  do 10 i = 1, 4
     do 20 j = 1, 3
        if (j.eq.2) goto_skip = 0
        total = total + i*j
 20  continue
 10 end do
  last = 0
  do i = 10, 1, -3
     if (i == 4) cycle
     last = last*100 + i
     if (i < 3) exit
  end do
  w = 0
  do while (w*w < 50)
     w = w + 1
  enddo
end subroutine
""".replace("if (j.eq.2) goto_skip = 0\n", ""))
    total, last, w = ns["f_s"](0, 0, 0)
    assert total == sum(i * j for i in range(1, 5) for j in range(1, 4))
    assert last == 100701 and w == 8           # 10, 7, (4 skipped), 1 -> exit


def test_derived_types_component_broadcast_and_reference_arguments():
    ns = load("""
module t
  integer, parameter :: BIG = 7
  type cell
     double precision :: v
     integer :: k
     double precision, allocatable, dimension(:) :: w
  end type
contains
  subroutine fill(c, n, s, biggest)
    integer :: n, biggest
    type(cell) :: c(n)
    double precision :: s
    integer :: i
    c%v = 0
    c%k = BIG
    do i = 1, n
       allocate(c(i)%w(i))
       c(i)%w = i
       c(i)%v = c(i)%v + sum(c(i)%w)
       call bump(c(i)%k, i)
    end do
    s = sum(c%v)
    biggest = maxval(c%k)
  end subroutine
  subroutine bump(k, by)
    integer :: k, by
    k = k + by
  end subroutine
end module
""")
    c = ns["f_alloc"]("t", [(1, 4)], ns["T_cell"])
    n, s, biggest = ns["f_fill"](c, 4, F64(0), 0)
    assert s == 1 + 4 + 9 + 16 and biggest == 7 + 4
    assert [c[i].k for i in range(1, 5)] == [8, 9, 10, 11]                # the callee's scalar dummy reached the component
    assert c[3].w.a.tolist() == [3.0, 3.0, 3.0]


def test_array_sections_whole_array_arithmetic_and_intrinsics():
    ns = load("""
subroutine s(a, h, lo, hi, n1, n2, total)
  double precision :: a(3, 4), h(4)
  double precision :: lo, hi, total
  integer :: n1, n2, i, j
  do j = 1, 4
    do i = 1, 3
      a(i, j) = 10*i + j
    end do
  end do
  h = a(2, :)                 ! a row section
  h(2:3) = h(2:3) + 0.5d0
  h = h / sum(h)
  a(:, 1) = 0
  a = cshift(a, 1, 2)         ! columns move left, the first wraps to the end
  lo = minval(a(1, :))
  hi = maxval(a(:, 3))
  n1 = size(a, 1)
  n2 = size(a)
  total = sum(a)
end subroutine
""")
    a = ns["f_alloc"]("d", [(1, 3), (1, 4)])
    h = ns["f_alloc"]("d", [(1, 4)])
    lo, hi, n1, n2, total = ns["f_s"](a, h, F64(0), F64(0), 0, 0, F64(0))
    row = np.array([21.0, 22.5, 23.5, 24.0])
    s = 0.0
    for x in row:
        s += x
    assert np.array_equal(h.a, row / s)
    assert a.a[:, 3].tolist() == [0.0, 0.0, 0.0] and a.a[:, 0].tolist() == [12.0, 22.0, 32.0]
    assert (lo, hi, n1, n2) == (0.0, 34.0, 3, 12)
    assert total == sum(10 * i + j for i in range(1, 4) for j in range(2, 5))


def test_generic_interface_functions_and_allocatable_copy_out():
    ns = load("""
module g
  interface make
    module procedure make_i
    module procedure make_r
    module procedure make_r2
  end interface
contains
  subroutine make_i(x, n)
    integer, allocatable, dimension(:) :: x
    integer :: n
    allocate(x(n))
    x = 1
  end subroutine
  subroutine make_r(x, n)
    double precision, allocatable, dimension(:) :: x
    integer :: n
    allocate(x(n))
    x = 2
  end subroutine
  subroutine make_r2(x, n, m)
    double precision, allocatable, dimension(:,:) :: x
    integer :: n, m
    allocate(x(n, m))
    x = 3
  end subroutine
  function area(w, h) result(r)
    double precision :: w, h, r
    r = w * h
  end function
  subroutine drive(si, sr, s2, ar)
    integer :: si
    double precision :: sr, s2, ar
    integer, allocatable, dimension(:) :: vi
    double precision, allocatable, dimension(:) :: vr
    double precision, allocatable, dimension(:,:) :: v2
    call make(vi, 3)
    call make(vr, 4)
    call make(v2, 2, 5)
    si = sum(vi)
    sr = sum(vr)
    s2 = sum(v2)
    ar = area(1.5d0, sr)
    if (allocated(vi)) deallocate(vi)
    if (allocated(vi) .eqv. .false.) si = si + 100
  end subroutine
end module
""")
    assert ns["f_drive"](0, F64(0), F64(0), F64(0)) == (103, 8.0, 30.0, 12.0)


def test_select_case_and_list_directed_and_formatted_input(tmp_path):
    path = tmp_path / "in.txt"
    path.write_text("! header line\n"
                    "3 2.5d0, 'two words'  tail ignored\n"
                    "\n"
                    "RAIN_ID 1 2\n"
                    "3\n"
                    "param_file  : SETTINGS/params.dat\n"
                    "n_setups    : 0012\n"
                    "7\n")
    ns = load("""
subroutine s(n, x, name, kind, v, path, k, nrec)
  integer :: n, kind, k, nrec, io
  double precision :: x
  character(len=64) :: name, tag, comment, path, tmp
  integer :: v(3)
  read (20, *)
  read (20, *) n, x, name
  read (20, *) tag, v                ! continues onto the next record for the third value
  select case (tag)
    case ('PET_ID')
      kind = 1
    case ('RAIN_ID', 'OTHER')
      kind = 2
    case default
      kind = 3
  end select
  read (20, 525) comment, path
  read (20, 530) comment, k
  nrec = 0
  do
    read (20, *, iostat = io) tmp
    if (io /= 0) exit
    nrec = nrec + 1
  end do
  rewind 20
  read (20, '(A)') tmp
  if (tmp(1:1) == '!') nrec = nrec
525 format(a13,1X,a700)
530 format(a13,1X,i4)
end subroutine
""".replace("  if (tmp(1:1) == '!') nrec = nrec\n", ""))
    ns["__runtime__"].open_unit(20, str(path))
    v = ns["f_alloc"]("i", [(1, 3)])
    n, x, name, kind, path_out, k, nrec = ns["f_s"](0, F64(0), "", 0, v, "", 0, 0)
    assert (n, x, name, kind, k, nrec) == (3, 2.5, "two words", 2, 12, 1)
    assert v.a.tolist() == [1, 2, 3] and path_out.strip() == "SETTINGS/params.dat"


def test_statements_outside_the_subset_become_loud_stubs():
    prog = f90_translate.Program()
    prog.add_source("""
subroutine ok(a)
  integer :: a
  a = a + 1
end subroutine
subroutine odd(a)
  integer :: a
  where (a > 0) a = 1
end subroutine
""")
    ns = prog.load()
    assert list(ns["_NOT_TRANSLATED"]) == ["odd"] and ns["f_ok"](1) == (2,)
    with pytest.raises(ns["FUnsupported"]):
        ns["f_odd"](1)
