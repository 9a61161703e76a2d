!
! eigen_libs_mod.f90 -- iso_c_binding replacement of the reference's module eigen_libs_mod
! (src/eigen_libs.F:70-216) over libeigenexa_b200.so.  A Fortran application keeps
!     use eigen_libs_mod
!     call eigen_init()                      ! or eigen_init(comm, order)
!     call eigen_get_matdims(n, nx, ny)
!     call eigen_s(n, nvec, a, lda, w, z, ldz, m_forward=48, m_backward=128, mode='A')
!     call eigen_free()
! unchanged and links with  eigen_libs_mod.o eigen_init_mpi.o -leigenexa_b200  instead of -lEigenExa.
! NOT built in this repository (no Fortran compiler / MPI in the image); INTEGRATION.md section 2.
!
! Argument meaning, defaults and the silent-return error convention are the reference's: nvec defaults to n,
! m_forward to 48, m_backward to 128, mode to 'A' (src/eigen_libs.F:178-192); a(lda,*), z(ldz,*) hold the local part
! of the 2D cyclic distribution (src/eigen_libs0.F:1986-2166); w(1:n) is returned on every rank, ascending.
!
module eigen_libs_mod
  use iso_c_binding
  implicit none
  private
  public :: eigen_init, eigen_free, eigen_get_matdims, eigen_s, eigen_sx
  public :: eigen_get_procs, eigen_get_id, eigen_get_errinfo
  public :: eigen_loop_start, eigen_loop_end, eigen_translate_l2g, eigen_translate_g2l
  public :: eigen_owner_node, eigen_owner_index

  interface
    subroutine ee_init_mpi_f(comm_f, order) bind(C, name="eigen_init_mpi_f")     ! bindings/eigen_init_mpi.c
      import :: c_int, c_char
      integer(c_int), value :: comm_f
      character(kind=c_char), intent(in) :: order(*)
    end subroutine
    subroutine eigen_free() bind(C, name="eigen_free")                           ! src/eigen_libs.F:204-216
    end subroutine
    subroutine ee_s(n, nvec, a, lda, w, z, ldz, mf, mb, mode) bind(C, name="eigen_s")
      import :: c_int, c_double, c_char
      integer(c_int), value :: n, nvec, lda, ldz, mf, mb
      real(c_double) :: a(lda, *), w(*), z(ldz, *)
      character(kind=c_char), intent(in) :: mode(*)
    end subroutine
    subroutine ee_sx(n, nvec, a, lda, w, z, ldz, mf, mb, mode) bind(C, name="eigen_sx")
      import :: c_int, c_double, c_char
      integer(c_int), value :: n, nvec, lda, ldz, mf, mb
      real(c_double) :: a(lda, *), w(*), z(ldz, *)
      character(kind=c_char), intent(in) :: mode(*)
    end subroutine
    subroutine ee_get_matdims(n, nx, ny, mf, mb, mode) bind(C, name="eigen_get_matdims")
      import :: c_int, c_char
      integer(c_int), value :: n, mf, mb
      integer(c_int) :: nx, ny
      character(kind=c_char), intent(in) :: mode(*)
    end subroutine
    subroutine eigen_get_procs(nnod, x_nnod, y_nnod) bind(C, name="eigen_get_procs")
      import :: c_int
      integer(c_int) :: nnod, x_nnod, y_nnod
    end subroutine
    subroutine eigen_get_id(inod, x_inod, y_inod) bind(C, name="eigen_get_id")    ! 1-based ids
      import :: c_int
      integer(c_int) :: inod, x_inod, y_inod
    end subroutine
    subroutine eigen_get_errinfo(info) bind(C, name="eigen_get_errinfo")
      import :: c_int
      integer(c_int) :: info
    end subroutine
    ! index helpers, explicit (nnod, inod) forms (src/eigen_libs0.F:1816-2258)
    function eigen_loop_start(istart, nnod, inod) bind(C, name="eigen_loop_start") result(r)
      import :: c_int
      integer(c_int), value :: istart, nnod, inod
      integer(c_int) :: r
    end function
    function eigen_loop_end(iend, nnod, inod) bind(C, name="eigen_loop_end") result(r)
      import :: c_int
      integer(c_int), value :: iend, nnod, inod
      integer(c_int) :: r
    end function
    function eigen_translate_l2g(ictr, nnod, inod) bind(C, name="eigen_translate_l2g") result(r)
      import :: c_int
      integer(c_int), value :: ictr, nnod, inod
      integer(c_int) :: r
    end function
    function eigen_translate_g2l(ictr, nnod, inod) bind(C, name="eigen_translate_g2l") result(r)
      import :: c_int
      integer(c_int), value :: ictr, nnod, inod
      integer(c_int) :: r
    end function
    function eigen_owner_node(ictr, nnod, inod) bind(C, name="eigen_owner_node") result(r)
      import :: c_int
      integer(c_int), value :: ictr, nnod, inod
      integer(c_int) :: r
    end function
    function eigen_owner_index(ictr, nnod, inod) bind(C, name="eigen_owner_index") result(r)
      import :: c_int
      integer(c_int), value :: ictr, nnod, inod
      integer(c_int) :: r
    end function
  end interface

contains

  subroutine eigen_init(comm, order)                     ! src/eigen_libs.F:70-104
    use mpi, only : MPI_COMM_WORLD
    integer, intent(in), optional :: comm
    character(*), intent(in), optional :: order
    integer :: comm0
    character(kind=c_char) :: order0(2)
    comm0 = MPI_COMM_WORLD
    if (present(comm)) comm0 = comm
    order0(1) = 'C'
    if (present(order)) then
      if (len(order) >= 1) order0(1) = order(1:1)
    end if
    order0(2) = c_null_char
    call ee_init_mpi_f(int(comm0, c_int), order0)
  end subroutine eigen_init

  subroutine eigen_get_matdims(n, nx, ny, m_forward, m_backward, mode)   ! src/eigen_libs.F:106-148
    integer, intent(in) :: n
    integer, intent(out) :: nx, ny
    integer, intent(in), optional :: m_forward, m_backward
    character(*), intent(in), optional :: mode
    integer(c_int) :: mf, mb, nx0, ny0
    character(kind=c_char) :: mode0(2)
    mf = 48; mb = 128
    if (present(m_forward)) mf = m_forward
    if (present(m_backward)) mb = m_backward
    mode0(1) = 'O'
    if (present(mode)) then
      if (len(mode) >= 1) mode0(1) = mode(1:1)
    end if
    mode0(2) = c_null_char
    call ee_get_matdims(int(n, c_int), nx0, ny0, mf, mb, mode0)
    nx = nx0; ny = ny0
  end subroutine eigen_get_matdims

  subroutine eigen_s(n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode)   ! src/eigen_libs.F:150-202
    integer, intent(in) :: n
    integer, intent(in), optional :: nvec
    integer, intent(in) :: lda, ldz
    real(8), intent(inout) :: a(lda, *)
    real(8), intent(out) :: w(*), z(ldz, *)
    integer, intent(in), optional :: m_forward, m_backward
    character(*), intent(in), optional :: mode
    call solve(.false., n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode)
  end subroutine eigen_s

  subroutine eigen_sx(n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode)  ! src/eigen_sx.F:30-59
    integer, intent(in) :: n
    integer, intent(in), optional :: nvec
    integer, intent(in) :: lda, ldz
    real(8), intent(inout) :: a(lda, *)
    real(8), intent(out) :: w(*), z(ldz, *)
    integer, intent(in), optional :: m_forward, m_backward
    character(*), intent(in), optional :: mode
    call solve(.true., n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode)
  end subroutine eigen_sx

  subroutine solve(penta, n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode)
    logical, intent(in) :: penta
    integer, intent(in) :: n, lda, ldz
    integer, intent(in), optional :: nvec, m_forward, m_backward
    real(8), intent(inout) :: a(lda, *)
    real(8), intent(out) :: w(*), z(ldz, *)
    character(*), intent(in), optional :: mode
    integer(c_int) :: nv, mf, mb
    character(kind=c_char) :: mode0(2)
    nv = n; mf = 48; mb = 128                       ! defaults of the reference (src/eigen_libs.F:178-192)
    if (present(nvec)) nv = nvec
    if (present(m_forward)) mf = m_forward
    if (present(m_backward)) mb = m_backward
    mode0(1) = 'A'
    if (present(mode)) then
      if (len(mode) >= 1) mode0(1) = mode(1:1)
    end if
    mode0(2) = c_null_char
    if (penta) then
      call ee_sx(int(n, c_int), nv, a, int(lda, c_int), w, z, int(ldz, c_int), mf, mb, mode0)
    else
      call ee_s(int(n, c_int), nv, a, int(lda, c_int), w, z, int(ldz, c_int), mf, mb, mode0)
    end if
  end subroutine solve

end module eigen_libs_mod
