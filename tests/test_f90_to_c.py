"""CPU tests of oracle/f90_to_c.py, the Fortran-subset translator that turns the reference's own source
into the parity authority (oracle/_ref).  The Fortran below is written for these tests (it is not
reference code); each construct the EVP path uses is translated, compiled with gcc and executed, and the
result is compared with the value Fortran semantics prescribe.  Unsupported constructs must raise
TranslateError -- the translator never guesses.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import f90_to_c as T

MODULE = """
      module demo_mod
      implicit none
      integer (kind=int_kind), parameter :: nlev = 3
      real (kind=dbl_kind), parameter :: &
         c0 = 0.0_dbl_kind, c1 = 1.0_dbl_kind, c2 = 2.0_dbl_kind, &
         p5 = 0.5_dbl_kind , & ! a comment with a ' quote
         third = c1/3.0_dbl_kind, &
         tiny = 1.0e-11_dbl_kind, big = 1.5d3
      real (kind=dbl_kind) :: scale
      logical (kind=log_kind) :: flag
      contains

      subroutine kernel (nx, ny, a, b, mask, out, cube, prof, total)
      integer (kind=int_kind), intent(in) :: nx, ny
      real (kind=dbl_kind), dimension (nx,ny), intent(in) :: a, b
      logical (kind=log_kind), dimension (nx,ny), intent(in) :: &
         mask      ! continuation with a trailing comment
      real (kind=dbl_kind), dimension (nx,ny), intent(out) :: out
      real (kind=dbl_kind), dimension (nx,ny,nlev), intent(out) :: cube
      real (kind=dbl_kind), dimension (-1:nlev), intent(inout) :: prof
      real (kind=dbl_kind), intent(out) :: total
      integer (kind=int_kind) :: i, j, k
      real (kind=dbl_kind) :: x, y
      real (kind=dbl_kind), dimension (nx,ny) :: work

      work(:,:) = c0
      cube = c2
      total = c0
      do j = 1, ny
      do i = 1, nx
         x = a(i,j)
         y = b(i,j)
         if (mask(i,j) .and. .not. x > y) then
            out(i,j) = -x**2 + y**3 / (c1 + x*x)      ! ** binds tighter than unary minus
         else if (.not. mask(i,j)) then
            out(i,j) = sign(c1, y) * max(x, y, p5) - min(x, y)
         else
            out(i,j) = sqrt(abs(x - y)) * scale + real(i+j,kind=dbl_kind) * third
         endif
#ifdef WITH_EXTRA
         out(i,j) = out(i,j) + big
#else
         out(i,j) = out(i,j) + tiny
#endif
#if defined(WITH_EXTRA) || defined(OTHER)
         work(i,j) = c1
#endif
         if (flag) total = total + out(i,j)
         do k = 1, nlev
            cube(i,j,k) = cube(i,j,k) * out(i,j) + &
                          real(k,kind=dbl_kind) &
                        + work(i,j)
         enddo
      enddo
      enddo
      do k = -1, nlev
         prof(k) = prof(k) + real(k,kind=dbl_kind) * p5 + x**0.5_dbl_kind
      enddo
      call helper (nx, ny, cube(:,:,2), total)
      end subroutine kernel

      subroutine helper (nx, ny, plane, acc)
      integer (kind=int_kind), intent(in) :: nx, ny
      real (kind=dbl_kind), dimension (nx,ny), intent(inout) :: plane
      real (kind=dbl_kind), intent(inout) :: acc
      integer (kind=int_kind) :: i, j
      do j = 1, ny
      do i = 1, nx
         plane(i,j) = plane(i,j) + c1
         if (mod(i+j,2) == 0) acc = acc + c1
      enddo
      enddo
      end subroutine helper
      end module demo_mod
"""


def _build(tmp_path, defines):
    src = tmp_path / "demo.F90"
    src.write_text(MODULE)
    tr = T.Translator(defines)
    tr.module_decls(str(src))
    for name in ("helper", "kernel"):
        missing = tr.subroutine(str(src), name)
        assert not [m for m in missing if m not in tr.subs], missing
    cfile = tmp_path / "demo.c"
    cfile.write_text(tr.emit_file())
    lib = tmp_path / ("libdemo_%s.so" % "_".join(defines or ["plain"]))
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    subprocess.check_call([cc, "-std=gnu11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unused",
                           "-Wno-parentheses", "-o", str(lib), str(cfile), "-lm"])
    return C.CDLL(str(lib))


def _expected(a, b, mask, scale, flag, extra, prof0):
    nx, ny = a.shape
    out = np.zeros((nx, ny))
    cube = np.full((nx, ny, 3), 2.0)
    work = np.zeros((nx, ny))
    total = 0.0
    third = 1.0 / 3.0
    x = y = 0.0
    for j in range(ny):
        for i in range(nx):
            x, y = a[i, j], b[i, j]
            if mask[i, j] and not (x > y):
                o = -(x * x) + (y * y * y) / (1.0 + x * x)
            elif not mask[i, j]:
                o = np.copysign(1.0, y) * max(max(x, y), 0.5) - min(x, y)
            else:
                o = np.sqrt(abs(x - y)) * scale + float(i + 1 + j + 1) * third
            o = o + (1.5e3 if extra else 1.0e-11)
            if extra:
                work[i, j] = 1.0
            out[i, j] = o
            if flag:
                total += o
            for k in range(3):
                cube[i, j, k] = cube[i, j, k] * o + float(k + 1) + work[i, j]
    prof = prof0.copy()
    for k in range(-1, 4):
        prof[k + 1] = prof[k + 1] + float(k) * 0.5 + x ** 0.5
    cube[:, :, 1] += 1.0
    for j in range(ny):
        for i in range(nx):
            if (i + 1 + j + 1) % 2 == 0:
                total += 1.0
    return out, cube, prof, total


@pytest.mark.parametrize("defines", [(), ("WITH_EXTRA",), ("OTHER",)], ids=["plain", "WITH_EXTRA", "OTHER"])
def test_translated_subset_executes_with_fortran_semantics(tmp_path, defines):
    L = _build(tmp_path, list(defines))
    L.ref_init_parameters()
    rng = np.random.default_rng(3)
    nx, ny = 7, 5
    a = np.asfortranarray(rng.uniform(0.1, 2.0, (nx, ny)))
    b = np.asfortranarray(rng.uniform(-2.0, 2.0, (nx, ny)))
    mask = np.asfortranarray((rng.random((nx, ny)) > 0.4).astype(np.int32))
    out = np.zeros((nx, ny), order="F")
    cube = np.zeros((nx, ny, 3), order="F")
    prof0 = rng.random(5)
    prof = prof0.copy()
    total = C.c_double(-1.0)
    C.c_double.in_dll(L, "v_scale").value = 1.75
    C.c_int32.in_dll(L, "v_flag").value = 1
    dp = C.POINTER(C.c_double)
    L.v_kernel(C.byref(C.c_int32(nx)), C.byref(C.c_int32(ny)), a.ctypes.data_as(dp), b.ctypes.data_as(dp),
               mask.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(dp), cube.ctypes.data_as(dp),
               prof.ctypes.data_as(dp), C.byref(total))
    e_out, e_cube, e_prof, e_total = _expected(a, b, mask, 1.75, True, "WITH_EXTRA" in defines, prof0)
    if "OTHER" in defines:   # '#if defined(A) || defined(B)': work = 1 without the WITH_EXTRA offset
        e_cube = e_cube + 1.0
    np.testing.assert_array_equal(out, e_out)
    np.testing.assert_array_equal(cube, e_cube)
    np.testing.assert_array_equal(prof, e_prof)
    assert total.value == e_total


@pytest.mark.parametrize("expr,want", [
    ("-a**2", "-f_powi(v_a, 2)"),
    ("(a+b)**2*c", "f_powi((v_a+v_b), 2)*v_c"),
    ("w(i,j)**2", "f_powi(v_w(v_i,v_j), 2)"),
    (".not. a > b .and. c", "!(v_a>v_b) && v_c"),
    ("a /= b .or. a == 1.0_dbl_kind", "v_a != v_b || v_a==1.0"),
    ("1.5d-3 + 2.e0", "1.5e-3+2.e0"),
    ("x%ilo + 1", "v_x.v_ilo+1"),
    ("a .ge. b", "v_a >= v_b"),
])
def test_expression_rewrites(expr, want):
    scope = T.Scope({"w": T.Decl("w", "real", ["n", "m"])}, {})
    assert T.conv_expr(expr, scope).replace(" ", "") == want.replace(" ", "")


@pytest.mark.parametrize("body,why", [
    ("where (a > c0) a = c0", "where"),
    ("do i = 1, n, 2\n a(i) = c0\n enddo", "stride"),
    ("b = a(1:2)", "section"),
    ("a = b + a(:)*q(:,1)", "rank"),
    ("call helper(a + b)", "expression argument"),
    ("x = 2**i", "variable exponent: integer ** integer is not pow()"),
    ("x = c0**x**2", "chained **"),
    ("x = unknown_fn(c0) ! resolved only if declared", None),
])
def test_unsupported_constructs_are_refused(tmp_path, body, why):
    src = tmp_path / "bad.F90"
    src.write_text("""
      subroutine bad (n, a, b, q)
      integer (kind=int_kind), intent(in) :: n
      real (kind=dbl_kind), dimension (n) :: a, b
      real (kind=dbl_kind), dimension (n,2) :: q
      real (kind=dbl_kind) :: x, c0
      integer (kind=int_kind) :: i
      %s
      end subroutine bad
""" % body)
    tr = T.Translator([])
    if why is None:   # unknown names are reported to the caller, who must resolve them or fail
        assert "unknown_fn" in tr.subroutine(str(src), "bad")
    else:
        with pytest.raises(T.TranslateError):
            tr.subroutine(str(src), "bad")


def test_cpp_nested_conditionals():
    text = "\n".join(["a", "#ifdef X", "b", "#if !defined(Y) && defined(X)", "c", "#else", "d", "#endif", "#else", "e",
                      "#endif", "f"])
    keep = lambda defs: [ln for ln in T.cpp(text, set(defs)) if ln]
    assert keep([]) == ["a", "e", "f"]
    assert keep(["X"]) == ["a", "b", "c", "f"]
    assert keep(["X", "Y"]) == ["a", "b", "d", "f"]


# ------------------------------------------------------------------------------------------------
# round 2: the constructs of the reference's block / halo machinery (serial/ice_boundary.F90, ice_blocks.F90)
# ------------------------------------------------------------------------------------------------
HALO_LIKE = """
      module halo_mod
      implicit none
      integer (int_kind), parameter :: dir_north = 1, dir_south = 2, dir_east = 3
      type, public :: table
         integer (int_kind) :: count, rows
         logical (log_kind) :: flag
         integer (int_kind), dimension(:,:), pointer :: addr
      end type
      real (dbl_kind), dimension(:,:), allocatable :: buf
      contains

      function neighbour(id, direction, bndy) result (nbr)
      integer (int_kind), intent(in) :: id, direction
      character (*), intent(in) :: bndy
      integer (int_kind) :: nbr
      integer (int_kind) :: step
      call get_param(id, step=step)
      select case (direction)
      case (dir_north)
         nbr = id + step
         if (nbr > 10) then
            select case (bndy)
            case ('open', 'closed')
               nbr = 0
            case ('cyclic')
               nbr = nbr - 10
            case ('tripole':'tripoleT')
               nbr = -id
            case default
               call stop_it('unknown boundary')
            end select
         endif
      case (dir_south:dir_east)
         nbr = id - step
      case default
         nbr = -999
      end select
      end function neighbour

      subroutine fill (t, kind, scale, total, opt)
      type (table), intent(inout) :: t
      character (*), intent(in) :: kind
      real (dbl_kind), intent(in) :: scale
      real (dbl_kind), intent(out) :: total
      real (dbl_kind), intent(in), optional :: opt
      integer (int_kind) :: i, j, n, nx
      integer (int_kind), dimension(:), pointer :: glob
      real (dbl_kind) :: f
      if (present(opt)) then
         f = opt
      else
         f = 0.5_dbl_kind
      endif
      nx = 0
      if (allocated(buf)) nx = size(buf,dim=1)
      call get_param(abs(-3), list=glob)
      n = t%count
      outer: do j = 1, t%rows
         do i = 1, 3
            n = n + 1
            t%addr(1,n) = glob(i) + neighbour(i, dir_north, kind)
            t%addr(2,n) = j
            if (kind == 'cyclic' .and. .not. t%flag) t%addr(2,n) = -j
         end do
      end do outer
      t%count = n
      buf = scale
      total = f * nx
      call helper2(t, -n, 'tag')
      end subroutine fill

      subroutine helper2 (t, m, label)
      type (table), intent(inout) :: t
      integer (int_kind), intent(in) :: m
      character (*), intent(in) :: label
      if (label /= 'tag') call stop_it('bad label')
      t%rows = t%rows + m
      end subroutine helper2
      end module halo_mod
"""

HALO_HOST = r"""
static int g_stopped;
static void v_stop_it(const char *m) { (void)m; g_stopped = 1; }
struct f_kw_get_param { int32_t *v_step; int32_t **v_list; };
static int32_t g_list[3] = {100, 200, 300};
static void v_get_param(int32_t *id, struct f_kw_get_param kw) {
    if (kw.v_step) *kw.v_step = *id == 9 ? 5 : 1;
    if (kw.v_list) *kw.v_list = g_list;
}
int stopped(void) { return g_stopped; }
"""


def test_halo_machinery_constructs(tmp_path):
    """select case on integers (value, range, default) and strings (list, range, default), named do, character
    dummies and string literals, == / /= on strings, derived type with a pointer-array component, optional dummy
    with present(), allocated() / size(a,dim=n) of a module allocatable, pointer array attached by the callee through
    a keyword argument, expression and string actual arguments, a function with a result clause called inside an
    expression, whole-array assignment to an allocatable -- translated, compiled and executed."""
    src = tmp_path / "halo_mod.F90"
    src.write_text(HALO_LIKE)
    tr = T.Translator(())
    for n in ("buf_n1", "buf_n2"):
        tr.add_global(T.Decl(n, "integer"))
    tr.module_decls(str(src), only={"dir_north", "dir_south", "dir_east"})
    tr.type_def(str(src), "table", {"addr": ["2", None]})
    tr.add_global(T.Decl("buf", "real", ["buf_n1", "buf_n2"], pointer=True))
    missing = set(tr.function(str(src), "neighbour"))
    missing |= set(tr.subroutine(str(src), "helper2"))
    missing |= set(tr.subroutine(str(src), "fill"))
    assert missing - set(tr.subs) == {"get_param", "stop_it"}
    code = tr.emit_file()
    # host functions must precede the generated routines: put them right after the prelude + types
    k = code.index("void ref_init_parameters")
    cfile = tmp_path / "halo.c"
    cfile.write_text(code[:k] + HALO_HOST + code[k:])
    lib = tmp_path / "libhalo.so"
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    subprocess.check_call([cc, "-std=gnu11", "-O2", "-fPIC", "-shared", "-Wno-unused", "-Wno-parentheses", "-o",
                           str(lib), str(cfile), "-lm"])
    L = C.CDLL(str(lib))
    L.ref_init_parameters()
    L.v_neighbour.restype = C.c_int32
    L.v_neighbour.argtypes = [C.c_int32, C.c_int32, C.c_char_p]
    assert L.v_neighbour(4, 1, b"open") == 5
    assert L.v_neighbour(10, 1, b"open") == 0 and L.v_neighbour(10, 1, b"closed") == 0
    assert L.v_neighbour(10, 1, b"cyclic") == 1
    assert L.v_neighbour(9, 1, b"cyclic") == 4            # step 5 from the keyword call
    assert L.v_neighbour(10, 1, b"tripole") == -10 and L.v_neighbour(10, 1, b"tripolet") == -10
    assert L.v_neighbour(7, 2, b"x") == 6 and L.v_neighbour(7, 3, b"x") == 6 and L.v_neighbour(7, 4, b"x") == -999
    assert L.stopped() == 0
    L.v_neighbour(10, 1, b"weird")
    assert L.stopped() == 1

    class Table(C.Structure):
        _fields_ = [("count", C.c_int32), ("rows", C.c_int32), ("flag", C.c_int32), ("addr", C.POINTER(C.c_int32))]

    addr = np.zeros((2, 40), dtype=np.int32, order="F")
    t = Table(2, 2, 0, addr.ctypes.data_as(C.POINTER(C.c_int32)))
    buf = np.zeros((4, 3), order="F")
    C.c_int32.in_dll(L, "v_buf_n1").value = 4
    C.c_int32.in_dll(L, "v_buf_n2").value = 3
    C.c_void_p.in_dll(L, "v_buf_").value = buf.ctypes.data
    total = C.c_double(0.0)
    L.v_fill(C.byref(t), b"cyclic", C.byref(C.c_double(2.5)), C.byref(total), None)
    assert total.value == 0.5 * 4 and np.all(buf == 2.5)
    assert t.count == 8 and t.rows == 2 - 8
    want1 = [100 + 2, 200 + 3, 300 + 4] * 2               # glob(i) + neighbour(i, north, 'cyclic')
    assert addr[0, 2:8].tolist() == want1
    assert addr[1, 2:8].tolist() == [-1, -1, -1, -2, -2, -2]   # flag false and kind == 'cyclic'
    L.v_fill(C.byref(t), b"open", C.byref(C.c_double(1.0)), C.byref(total), C.byref(C.c_double(3.0)))
    assert total.value == 3.0 * 4


def test_translate_range_of_a_routine(tmp_path):
    """A statement range of a routine as a function of its own: the construct that starts at a given statement
    (to its matching end) and a range between two labelled statements; untouched locals are not declared."""
    src = tmp_path / "rng.F90"
    src.write_text("""
      module rng_mod
      contains
      subroutine big (n, acc, other)
      integer (int_kind), intent(in) :: n
      integer (int_kind), intent(inout) :: acc
      real (dbl_kind), dimension(:,:), intent(in) :: other
      integer (int_kind) :: i, j, unused
      unused = 7
      acc = 1000
      do j = 1, n
         do i = 1, j
            if (i == j) then
               acc = acc + i
            endif
         end do
      end do
      acc = -1
      tail: do i = 1, n
         acc = acc + 2
      end do tail
      end subroutine big
      end module rng_mod
""")
    tr = T.Translator(())
    tr.translate_range(str(src), "big", "subroutine", r"^do\s+j\s*=", None, "big_loop", ["n", "acc"])
    tr.translate_range(str(src), "big", "subroutine", r"^tail\s*:", r"^end\s*do\s+tail$", "big_tail", ["n", "acc"])
    code = tr.emit_file()
    assert "v_unused" not in code and "v_other" not in code
    cfile = tmp_path / "rng.c"
    cfile.write_text(code)
    lib = tmp_path / "librng.so"
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    subprocess.check_call([cc, "-std=gnu11", "-O2", "-fPIC", "-shared", "-Wno-unused", "-o", str(lib), str(cfile)])
    L = C.CDLL(str(lib))
    acc = C.c_int32(5)
    L.v_big_loop(C.byref(C.c_int32(4)), C.byref(acc))
    assert acc.value == 5 + 1 + 2 + 3 + 4
    L.v_big_tail(C.byref(C.c_int32(3)), C.byref(acc))
    assert acc.value == 15 + 6
