!> poisson_b200.f90 -- drop-in replacement of module `poisson` (reference src/poisson.f90:6,132,257):
!> same three public routines, same poi_type argument list (src/initialization.f90:93-102).
!> The SOR runs on the device (red-black fast path, or the bit-exact lexicographic wavefront
!> ordering when o3d_set_sor_order(1) was called); omega is intent(inout) and persists exactly
!> as in the reference (src/integration.f90:222,247).
module poisson
  use iso_c_binding
  use o3d_b200_c
  implicit none

contains

  subroutine poisson_solver_0000(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn)
    real(kind=8), intent(inout) :: pp(:,:,:), omega
    real(kind=8), intent(in) :: rhs(:,:,:)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(in) :: eps
    integer, intent(in) :: nx, ny, nz, kmax, idyn
    integer(c_int) :: iters
    real(c_double) :: dmax
    print *, "* Poisson solver 0000 start"
    call o3d_check(o3d_poisson_solver_0000(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, &
         idyn, iters, dmax), "poisson_solver_0000")
    print *, "* Poisson solver 0000 end: iter, dmax, omega = ", iters, dmax, omega
  end subroutine poisson_solver_0000

  subroutine poisson_solver_0011(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn)
    real(kind=8), intent(inout) :: pp(:,:,:), omega
    real(kind=8), intent(in) :: rhs(:,:,:)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(in) :: eps
    integer, intent(in) :: nx, ny, nz, kmax, idyn
    integer(c_int) :: iters
    real(c_double) :: dmax
    print *, "* Poisson solver 0011 start"
    call o3d_check(o3d_poisson_solver_0011(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, &
         idyn, iters, dmax), "poisson_solver_0011")
    print *, "* Poisson solver 0011 end: iter, dmax, omega = ", iters, dmax, omega
  end subroutine poisson_solver_0011

  subroutine poisson_solver_111111(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn)
    real(kind=8), intent(inout) :: pp(:,:,:), omega
    real(kind=8), intent(in) :: rhs(:,:,:)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(in) :: eps
    integer, intent(in) :: nx, ny, nz, kmax, idyn
    integer(c_int) :: iters
    real(c_double) :: dmax
    print *, "* Poisson solver 111111 start"
    call o3d_check(o3d_poisson_solver_111111(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, &
         idyn, iters, dmax), "poisson_solver_111111")
    print *, "* Poisson solver 111111 end: iter, dmax, omega = ", iters, dmax, omega
  end subroutine poisson_solver_111111

end module poisson
