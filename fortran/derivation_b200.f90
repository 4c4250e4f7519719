!> derivation_b200.f90 -- drop-in replacement of module `derivation` (reference
!> src/derivation.f90): the same 20 public routines with the same (df, f, d) signature
!> (der_type, src/initialization.f90:86-91), each forwarding to the sm_100a CUDA stencil kernel
!> through the C ABI (include/o3d_b200.h, o3d_der).  Results are bit-identical to the reference
!> routines (tests/test_gpu_operators.py::test_all_18_derivative_routines_bit_exact).
!> dery1D (src/derivation.f90:950, init-time 1-D helper) stays host code, restated below.
!>
!> These host-pointer forms copy f to the device and df back on every call: they exist so that
!> callers outside the time loop (initial_conditions.f90:89-90,439-442, utils.f90:283-347)
!> keep working unchanged.  The time loop itself should go through integration_b200.f90.
module derivation
  use iso_c_binding
  use o3d_b200_c
  implicit none

contains

  subroutine derx_00(df, f, d)   ! replaces src/derivation.f90:6
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 1_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derx_00")
  end subroutine derx_00

  subroutine derxp_11(df, f, d)   ! replaces src/derivation.f90:62
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 1_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derxp_11")
  end subroutine derxp_11

  subroutine derxi_11(df, f, d)   ! replaces src/derivation.f90:111
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 1_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derxi_11")
  end subroutine derxi_11

  subroutine dery_00(df, f, d)   ! replaces src/derivation.f90:165
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 1_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "dery_00")
  end subroutine dery_00

  subroutine deryp_11(df, f, d)   ! replaces src/derivation.f90:219
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 1_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "deryp_11")
  end subroutine deryp_11

  subroutine deryi_11(df, f, d)   ! replaces src/derivation.f90:269
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 1_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "deryi_11")
  end subroutine deryi_11

  subroutine derz_00(df, f, d)   ! replaces src/derivation.f90:323
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 1_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derz_00")
  end subroutine derz_00

  subroutine derzp_11(df, f, d)   ! replaces src/derivation.f90:377
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 1_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzp_11")
  end subroutine derzp_11

  subroutine derzi_11(df, f, d)   ! replaces src/derivation.f90:427
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 1_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzi_11")
  end subroutine derzi_11

  subroutine derz_2dsim(df, f, d)   ! replaces src/derivation.f90:481
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 1_c_int, 3_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derz_2dsim")
  end subroutine derz_2dsim

  subroutine derxx_00(df, f, d)   ! replaces src/derivation.f90:497
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 2_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derxx_00")
  end subroutine derxx_00

  subroutine derxxi_11(df, f, d)   ! replaces src/derivation.f90:544
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 2_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derxxi_11")
  end subroutine derxxi_11

  subroutine derxxp_11(df, f, d)   ! replaces src/derivation.f90:593
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(0_c_int, 2_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derxxp_11")
  end subroutine derxxp_11

  subroutine deryy_00(df, f, d)   ! replaces src/derivation.f90:642
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 2_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "deryy_00")
  end subroutine deryy_00

  subroutine deryyp_11(df, f, d)   ! replaces src/derivation.f90:691
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 2_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "deryyp_11")
  end subroutine deryyp_11

  subroutine deryyi_11(df, f, d)   ! replaces src/derivation.f90:740
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(1_c_int, 2_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "deryyi_11")
  end subroutine deryyi_11

  subroutine derzz_00(df, f, d)   ! replaces src/derivation.f90:789
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 2_c_int, 0_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzz_00")
  end subroutine derzz_00

  subroutine derzzi_11(df, f, d)   ! replaces src/derivation.f90:836
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 2_c_int, 2_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzzi_11")
  end subroutine derzzi_11

  subroutine derzzp_11(df, f, d)   ! replaces src/derivation.f90:885
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 2_c_int, 1_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzzp_11")
  end subroutine derzzp_11

  subroutine derzz_2dsim(df, f, d)   ! replaces src/derivation.f90:934
    real(kind=8), intent(in) :: f(:,:,:), d
    real(kind=8), intent(out) :: df(:,:,:)
    call o3d_check(o3d_der(2_c_int, 2_c_int, 3_c_int, df, f, d, &
         int(size(f,1),c_int), int(size(f,2),c_int), int(size(f,3),c_int)), "derzz_2dsim")
  end subroutine derzz_2dsim

  !> dery1D stays on the host: a 1-D, init-time helper (reference src/derivation.f90:950-992).
  !> 6th-order centred interior; one-sided 2nd-order at the end points, centred 2nd-order next
  !> to them, centred 4th-order on the third point from each end.
  subroutine dery1D(df, f, dy)
    real(kind=8), intent(in) :: f(:)
    real(kind=8), intent(in) :: dy
    real(kind=8), intent(out) :: df(:)
    real(kind=8) :: s, a, b, c
    integer :: n, j
    n = size(f)
    s = 60.d0 * dy
    a = 1.d0 / s
    b = 9.d0 / s
    c = 45.d0 / s
    do j = 4, n - 3
       df(j) = a * (f(j+3) - f(j-3)) - b * (f(j+2) - f(j-2)) + c * (f(j+1) - f(j-1))
    end do
    df(1)   = (-f(3) + 4.d0 * f(2) - 3.d0 * f(1)) / (2.d0 * dy)
    df(n)   = (3.d0 * f(n) - 4.d0 * f(n-1) + f(n-2)) / (2.d0 * dy)
    df(2)   = (f(3) - f(1)) / (2.d0 * dy)
    df(n-1) = (f(n) - f(n-2)) / (2.d0 * dy)
    df(3)   = (-f(5) + 8.d0 * f(4) - 8.d0 * f(2) + f(1)) / (12.d0 * dy)
    df(n-2) = (-f(n) + 8.d0 * f(n-1) - 8.d0 * f(n-3) + f(n-4)) / (12.d0 * dy)
  end subroutine dery1D

end module derivation
