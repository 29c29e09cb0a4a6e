"""The Fortran-subset interpreter (tests/golden/f90run.py) on hand-written snippets whose results follow from the Fortran
standard alone.  The interpreter is what pins the oracle to the reference's source, so its own semantics
are checked here independently of any PIC code: integer division, fp32 rounding of every operation, DO-variable exit
values, column-major storage and the out-of-range first index, sections, the two goto idioms, functions, select case,
where, MPI_SendRecv as an element-count transfer (also between threads) and MPI subarray datatypes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import f90run as R  # noqa: E402

F = np.float32


def run(text, name, g=None, args=(), arrays=(), ints=(), defines=()):
    g = g or R.Globals()
    f = R.Sub(text, name, defines=set(defines), global_arrays=set(arrays), global_ints=set(ints)).compile()
    return f(g, *args), g


def test_integer_division_truncates_and_mixed_arithmetic_promotes():
    src = """
    subroutine t()
      integer :: i, j
      real :: a
      i = 7/2
      j = -7/2
      a = 7/2 + 7/2.
      r1 = i
      r2 = j
      r3 = a
    end subroutine t
    """
    _, g = run(src, "t")
    assert (g.r1, g.r2) == (3, -3)
    assert g.r3 == F(6.5) and isinstance(g.r3, np.float32)


def test_every_operation_rounds_to_fp32():
    src = """
    subroutine t()
      real :: a, b, c
      a = 16777216.
      b = 1.
      c = (a + b) - a
      r1 = c
      c = 0.1*3.
      r2 = c
    end subroutine t
    """
    _, g = run(src, "t")
    assert g.r1 == F(0.0)                                   # 2^24 + 1 is not representable: the sum rounds back to 2^24
    assert g.r2 == F(F(0.1) * F(3.0)) and g.r2 != np.float64(0.1) * 3


def test_do_variable_after_the_loop_and_zero_trip_loops():
    src = """
    subroutine t()
      integer :: i, n
      n = 0
      do i = 2, 9, 3
        n = n + i
      enddo
      r1 = i
      r2 = n
      do i = 5, 4
        n = -1
      enddo
      r3 = i
      r4 = n
    end subroutine t
    """
    _, g = run(src, "t")
    assert (g.r1, g.r2, g.r3, g.r4) == (11, 2 + 5 + 8, 5, 15)


def test_column_major_storage_sections_and_the_flat_first_index():
    src = """
    subroutine t()
      integer :: i, j
      do j = 1, 3
        do i = 1, 4
          a(i, j) = 10*j + i
        enddo
      enddo
      r1 = a(6, 1)
      b(1:2, 1) = a(3:4, 2)
      b(:, 2) = a(2, :)
      r2 = sum(a(2:3, 3))
    end subroutine t
    """
    g = R.Globals(a=R.FArr((4, 3)), b=R.FArr((3, 2)))
    run(src, "t", g, arrays=("a", "b"))
    assert g.r1 == F(22)                                    # a(6,1) is the 6th element in storage order = a(2,2)
    assert list(g.a.flat[:5]) == [11, 12, 13, 14, 21]
    assert g.b.nd().tolist() == [[23, 12], [24, 22], [0, 32]]
    assert g.r2 == F(32 + 33)


def test_forward_goto_skips_and_goto_to_loop_end_cycles():
    src = """
    subroutine t()
      integer :: i, n
      n = 0
      do i = 1, 6
        if (modulo(i, 2) .eq. 0) go to 58
        n = n + i
 58     continue
      enddo
      r1 = n
      goto 70
      n = -100
 70   continue
      r2 = n
    end subroutine t
    """
    _, g = run(src, "t")
    assert (g.r1, g.r2) == (1 + 3 + 5, 9)


def test_functions_sign_modulo_and_aint():
    src = """
    integer function wrapidx(i)
      integer :: i
      wrapidx = modulo(i - 1, n0) + 1
    end function wrapidx

    subroutine t()
      r1 = wrapidx(0)
      r2 = wrapidx(9)
      r3 = sign(2.5, -0.1) + sign(2.5, 3.)
      r4 = mod(-7, 3)
      r5 = modulo(-7, 3)
      r6 = aint(-2.7)
      r7 = int(2.999)
      r8 = nint(2.5) + 10*nint(-2.5) + 100*nint(3.49)
    end subroutine t
    """
    g = R.Globals(n0=4)
    f = R.Sub(src, "wrapidx", global_ints={"n0"}).compile()
    g.wrapidx = lambda i: f(g, i)
    run(src, "t", g, ints=("n0",))
    assert (g.r1, g.r2, g.r3, g.r4, g.r5, g.r6, g.r7) == (4, 1, F(0.0), -1, 2, F(-2.0), 2)
    assert g.r8 == 3 - 30 + 300                     # nint rounds halves away from zero


def test_cpp_conditionals_select_the_build():
    src = """
    subroutine t()
#ifdef twoD
      r1 = 2
#else
      r1 = 3
#endif
#ifndef MPI
      r1 = -1
#endif
    end subroutine t
    """
    assert run(src, "t", defines=("MPI", "twoD"))[1].r1 == 2
    assert run(src, "t", defines=("MPI",))[1].r1 == 3


def test_select_case_where_cycle_and_exit():
    src = """
    subroutine t(name)
      character (len=5) name
      integer :: i, n
      n = 0
      select case(name)
      case('tdens')
        n = 1
      case('idens')
        n = 2
      end select
      r1 = n
      do i = 1, 10
        if (i .eq. 2) cycle
        if (i .gt. 4) exit
        n = n + i
      enddo
      r2 = n
      where (w .ne. 0.)
        a = a/w
      elsewhere
        a = 0.
      endwhere
    end subroutine t
    """
    g = R.Globals(a=R.FArr((4,)), w=R.FArr((4,)))
    g.a.flat[:] = [2, 3, 4, 5]
    g.w.flat[:] = [2, 0, 8, 0]
    run(src, "t", g, args=("idens",), arrays=("a", "w"))
    assert (g.r1, g.r2) == (2, 2 + 1 + 3 + 4)
    assert list(g.a.flat) == [1, 0, 0.5, 0]


SENDRECV = """
subroutine swap()
  integer :: plusrank, minusrank, count
  plusrank = modulo(rank + 1, size0)
  minusrank = modulo(rank - 1, size0)
  count = 3
  call MPI_SendRecv(a(2, :), count, mpi_read, plusrank, 100, &
                    b(1, :), count, mpi_read, minusrank, 100, comm, status, ierr)
end subroutine swap
"""


def test_mpi_sendrecv_to_oneself_moves_count_elements_in_storage_order():
    g = R.Globals(a=R.FArr((2, 4)), b=R.FArr((2, 4)), rank=0, size0=1, comm=None, mpi_read=0, status=0, ierr=0)
    g.a.flat[:] = np.arange(8)
    g.b.flat[:] = -1
    f = R.Sub(SENDRECV, "swap", global_arrays={"a", "b"}, global_ints={"rank", "size0"}).compile()
    f(g)
    assert g.b.nd()[0].tolist() == [1, 3, 5, -1]            # 3 elements of a(2,:) land in the first 3 of b(1,:)
    assert g.b.nd()[1].tolist() == [-1, -1, -1, -1]


def test_mpi_sendrecv_between_three_ranks_is_a_ring_shift():
    comm = R.Comm(3)
    gs = []
    f = R.Sub(SENDRECV, "swap", global_arrays={"a", "b"}, global_ints={"rank", "size0"}).compile()
    for rk in range(3):
        g = R.Globals(a=R.FArr((2, 4)), b=R.FArr((2, 4)), rank=rk, size0=3, mpi_read=0, status=0, ierr=0)
        g.comm = comm
        g.a.flat[:] = 100 * rk + np.arange(8)
        gs.append(g)
    R.run_ranks([(lambda g=g: f(g)) for g in gs])
    for rk in range(3):
        src_rank = (rk - 1) % 3
        assert gs[rk].b.nd()[0].tolist() == [100 * src_rank + 1, 100 * src_rank + 3, 100 * src_rank + 5, 0]


def test_mpi_subarray_datatypes_describe_boxes():
    src = """
    subroutine mk()
      integer, allocatable, dimension(:) :: sizes, subsizes, starts
      allocate(sizes(2), subsizes(2), starts(2))
      sizes(1) = 4
      sizes(2) = 3
      subsizes(1) = 2
      subsizes(2) = 2
      starts(1) = 1
      starts(2) = 0
      call MPI_Type_create_subarray(2, sizes, subsizes, starts, MPI_ORDER_FORTRAN, mpi_read, lowcap, ierr)
      starts(1) = 2
      starts(2) = 1
      call MPI_Type_create_subarray(2, sizes, subsizes, starts, MPI_ORDER_FORTRAN, mpi_read, highcap, ierr)
    end subroutine mk

    subroutine xchg()
      call MPI_SendRecv(a, 1, lowcap, 0, 7, b, 1, highcap, 0, 7, comm, status, ierr)
    end subroutine xchg
    """
    g = R.Globals(a=R.FArr((4, 3)), b=R.FArr((4, 3)), rank=0, mpi_read=0, mpi_order_fortran=0, status=0, ierr=0)
    g.a.nd()[...] = np.arange(12).reshape(3, 4).T              # a(i,j) = (i-1) + 4*(j-1)
    R.Sub(src, "mk", global_arrays={"a", "b"}).compile()(g)
    R.Sub(src, "xchg", global_arrays={"a", "b"}).compile()(g)
    want = np.zeros((4, 3), F)
    want[2:4, 1:3] = g.a.nd()[1:3, 0:2]
    assert np.array_equal(g.b.nd(), want)


@pytest.mark.skipif(not os.path.isdir(os.environ.get("TRISTAN_REFERENCE", "/root/reference")), reason="reference checkout not present")
def test_committed_goldens_are_what_the_generator_makes(tmp_path, monkeypatch):
    """in the build container: re-run two fast generator groups and compare with the committed files"""
    import importlib
    monkeypatch.setenv("TRISTAN_GOLDEN_OUT", str(tmp_path))
    gen = importlib.import_module("make_ref_golden")
    gen = importlib.reload(gen)
    gen.gen_deposit()
    gen.gen_shock()
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("ref_deposit.npz", "ref_shock.npz"):
        a, b = np.load(os.path.join(here, name)), np.load(os.path.join(str(tmp_path), name))
        assert set(a.files) == set(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (name, k)


def test_double_precision_entities_data_and_the_minstd_step():
    src = """
    real function lcg(dseed)
      real(dprec) :: dseed
      real(dprec) :: s2p31, s2p31m, seed
      data s2p31m/2147483647.d0/, s2p31/2147483648.d0/
      seed = dseed
      seed = dmod(16807.d0*seed, s2p31m)
      lcg = seed/s2p31
      dseed = seed
    end function lcg
    """
    g = R.Globals(dseed=np.float64(123457.0))
    f = R.Sub(src, "lcg", alias_globals={"dseed"}).compile()
    v = f(g, None)
    assert g.dseed == np.float64((16807 * 123457) % 2147483647) and isinstance(g.dseed, np.float64)
    assert v == F(g.dseed / 2147483648.0) and isinstance(v, np.float32)
    v2 = f(g, None)
    assert g.dseed == np.float64((16807 * ((16807 * 123457) % 2147483647)) % 2147483647) and v2 != v


def test_scalar_arguments_are_passed_by_reference_and_optional_arguments():
    src = """
    subroutine polar(r, phi, x, y, scale)
      real :: r, phi, x, y
      real, optional :: scale
      x = r*cos(phi)
      y = r*sin(phi)
      if (present(scale)) then
        x = x*scale
      endif
    end subroutine polar

    subroutine t()
      real :: a, b
      call polar(2., 0.5, a, b)
      r1 = a
      r2 = b
      call polar(2., 0.5, q(2), b, 3.)
    end subroutine t
    """
    g = R.Globals(q=R.FArr((3,)))
    f = R.Sub(src, "polar").compile()
    g.polar = lambda *a: f(g, *a)
    run(src, "t", g, arrays=("q",))
    import ctypes
    import ctypes.util
    m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    m.cosf.restype = m.sinf.restype = ctypes.c_float
    m.cosf.argtypes = m.sinf.argtypes = [ctypes.c_float]
    cx, sy = F(F(2) * F(m.cosf(0.5))), F(F(2) * F(m.sinf(0.5)))
    assert (g.r1, g.r2) == (cx, sy)
    assert g.q.flat[1] == F(cx * F(3)) and g.q.flat[0] == 0


def test_int_of_a_non_finite_value_is_the_integer_indefinite():
    src = """
    subroutine t()
      integer :: k
      real :: a
      a = 0.
      k = int(alog10(a))
      r1 = k
      r2 = 10.**2.5
    end subroutine t
    """
    _, g = run(src, "t")
    assert g.r1 == -2 ** 31
    assert abs(float(g.r2) - 10 ** 2.5) < 1e-4 * 10 ** 2.5 and isinstance(g.r2, np.float32)


def test_mpi_allreduce_between_ranks_and_its_log():
    src = """
    subroutine t()
      real :: lo, lo1
      lo = 10. + rank
      call mpi_allreduce(lo, lo1, 1, mpi_read, mpi_min, mpi_comm_world, ierr)
      r1 = lo1
      call mpi_allreduce(h, hin, 4, mpi_read, mpi_sum, mpi_comm_world, ierr)
    end subroutine t
    """
    comm = R.Comm(3)
    f = R.Sub(src, "t", global_arrays={"h", "hin"}, global_ints={"rank"}).compile()
    gs = []
    for rk in range(3):
        g = R.Globals(rank=rk, h=R.FArr((2, 2)), hin=R.FArr((2, 2)), mpi_read=0, mpi_min="min", mpi_sum="sum", mpi_comm_world=0, ierr=0)
        g.comm = comm
        g.h.flat[:] = rk + 1
        gs.append(g)
    R.run_ranks([(lambda g=g: f(g)) for g in gs])
    for rk, g in enumerate(gs):
        assert g.r1 == F(10.0)
        assert list(g.hin.flat) == [6, 6, 6, 6]
        assert float(g.allreduce_log[0]) == 10.0 + rk and list(g.allreduce_log[1]) == [rk + 1] * 4


@pytest.mark.skipif(not os.path.isdir(os.environ.get("TRISTAN_REFERENCE", "/root/reference")), reason="reference checkout not present")
def test_the_references_restart_reader_accepts_files_written_by_the_package(tmp_path):
    """in the build container: restart() of code/restart.F90:209-253, run from its text, reads restflds / restprtl files that
    tristan_mp_pu_master_densdecomp_b200.restart wrote -- all items, no end-of-record, nothing left over"""
    from tristan_mp_pu_master_densdecomp_b200 import restart
    ref = os.environ.get("TRISTAN_REFERENCE", "/root/reference")
    heads = {f: R.preprocess(open(os.path.join(ref, "code", f)).read(), {"MPI"}) for f in ("fields.F90", "particles.F90", "aux.F90")}

    def kind_of(name):
        for t in heads.values():
            try:
                return R.module_kind(t, name)
            except KeyError:
                pass
        return "real"
    names = ("mxrest", "myrest", "mzrest", "dseed", "lapst", "xinject", "xinject2", "xinject3", "leftwall", "walloc", "totalpartnum")
    kinds = {n: kind_of(n) for n in names}
    assert kinds["dseed"] == kinds["xinject"] == kinds["walloc"] == "real8" and kinds["leftwall"] == "real"
    rng = np.random.default_rng(5)
    mx, my, mz, maxptl, ions, lecs = 7, 6, 5, 40, 9, 6
    fields = [rng.standard_normal((mz, my, mx)).astype(F) for _ in range(6)]
    p = np.zeros(maxptl, restart.PARTICLE_DTYPE)
    for k in p.dtype.names:
        p[k] = (rng.standard_normal(maxptl) * 3).astype(p.dtype[k]) if p.dtype[k].kind == "f" else rng.integers(-50, 50, maxptl)
    fld, prt = str(tmp_path / "restflds.d"), str(tmp_path / "restprtl.d")
    restart.write_fields(fld, fields, dseed=987654321.0, lap=77, xinject=3.5, xinject2=101.25, xinject3=0.5, leftwall=15.0, walloc=20.125)
    restart.write_particles(prt, p, ions, lecs, maxptl, totalpartnum=4242)
    text = open(os.path.join(ref, "code", "restart.F90")).read()
    ints = {"ions", "lecs", "maxptl", "maxhlf", "maxptl0", "size0", "rank", "debug", "mpi_status_size"} | {n for n, k in kinds.items() if k == "int"}
    sub = R.Sub(text, "restart", defines={"MPI"}, global_arrays={"ex", "ey", "ez", "bx", "by", "bz", "p"}, global_ints=ints,
                global_kinds=kinds).compile()
    g = R.Globals(rank=0, size0=1, debug=False, maxptl0=maxptl, maxptl=0, maxhlf=0, ions=0, lecs=0, mpi_status_size=5,
                  frestartfldlap="", frestartprtlap="", mpi_comm_world=0, mpi_integer=0,
                  **{nm: R.FArr((mx, my, mz)) for nm in ("ex", "ey", "ez", "bx", "by", "bz")})
    q = np.zeros(maxptl, p.dtype)
    g.p = R.RecArr(q)
    g.units = {7: [open(fld, "rb").read()], 8: [open(prt, "rb").read()]}
    g.reorder_particles = lambda: None
    for nm in ("mpi_irecv", "mpi_wait", "mpi_isend"):
        setattr(g, nm, lambda *a: None)
    sub(g)
    assert (g.mxrest, g.myrest, g.mzrest) == (mx, my, mz)
    assert g.readers[7].left == 0 and g.readers[8].left == 0          # every byte of both records was asked for
    for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
        assert np.array_equal(getattr(g, nm).nd().transpose(2, 1, 0), fields[a]), nm
    assert (g.dseed, g.lapst, g.xinject, g.xinject2, g.xinject3, g.leftwall, g.walloc) == (987654321.0, 77 + 1, 3.5, 101.25, 0.5, F(15.0), 20.125)
    assert (g.ions, g.lecs, g.totalpartnum, g.maxhlf) == (ions, lecs, 4242, maxptl // 2)
    assert np.array_equal(q[:ions], p[:ions]) and np.array_equal(q[maxptl // 2:maxptl // 2 + lecs], p[maxptl // 2:maxptl // 2 + lecs])


def test_operator_precedence_and_associativity_follow_the_standard():
    src = """
    subroutine t()
      real :: a, b, c
      logical :: p, q, s
      a = 2.
      b = 3.
      c = 4.
      r1 = -a**2
      r2 = a**b**2
      r3 = a - b + c
      r4 = a/b*c
      r5 = -a*b
      r6 = a + b*c**2
      r7 = 7/2*2
      r8 = 2*7/2
      p = .true.
      q = .false.
      s = .false.
      l1 = .not. q .and. s
      l2 = p .or. q .and. s
      l3 = a .lt. b .and. b .lt. c
      l4 = .not. a .gt. b
    end subroutine t
    """
    _, g = run(src, "t")
    assert g.r1 == F(-4.0)                      # ** binds tighter than unary minus
    assert g.r2 == F(512.0)                     # ** associates to the right: 2**(3**2)
    assert g.r3 == F(3.0) and g.r4 == F(F(F(2) / F(3)) * F(4))
    assert g.r5 == F(-6.0) and g.r6 == F(50.0)
    assert g.r7 == 6 and g.r8 == 7              # integer division, left to right
    assert bool(g.l1) is False                  # (.not. q) .and. s
    assert bool(g.l2) is True                   # .and. before .or.
    assert bool(g.l3) is True and bool(g.l4) is True     # relational before .not.


def test_cshift_sum_and_whole_array_expressions():
    src = """
    subroutine t()
      b = cshift(a, 1, 1)
      c = cshift(a, -1, 2)
      d = 0.5*(a + b)
      r1 = sum(a(2, :))
    end subroutine t
    """
    g = R.Globals(**{n: R.FArr((3, 2)) for n in "abcd"})
    g.a.nd()[...] = np.array([[1, 4], [2, 5], [3, 6]], F)
    run(src, "t", g, arrays="abcd")
    assert g.b.nd().tolist() == [[2, 5], [3, 6], [1, 4]]           # result(i) = a(i + shift), circular, along dim 1
    assert g.c.nd().tolist() == [[4, 1], [5, 2], [6, 3]]           # shift -1 along dim 2
    assert g.d.nd().tolist() == [[1.5, 4.5], [2.5, 5.5], [2.0, 5.0]]
    assert g.r1 == F(7.0)


def test_integer_powers_multiply_by_repeated_squaring():
    src = """
    subroutine t()
      real :: x
      x = 1.1
      r2 = x**2
      r3 = x**3
      r4 = x**4
      r5 = x**(-2)
    end subroutine t
    """
    _, g = run(src, "t")
    x = F(1.1)
    x2 = F(x * x)
    assert g.r2 == x2 and g.r3 == F(x * x2) and g.r4 == F(x2 * x2) and g.r5 == F(F(1) / x2)
