!> poisson_multigrid_b200.f90 -- drop-in replacement of module `poisson_multigrid` (reference
!> src/poisson_multigrid.f90:10).  The reference routine declares phi(0:nx+1,0:ny+1,0:nz+1) while
!> its only caller passes pp(nx,ny,nz) (src/integration.f90:227-244): undefined behaviour.  This
!> replacement takes the arrays as the caller really passes them -- (nx,ny,nz) -- and solves the
!> SAME 7-point operator and boundary rule as poisson_solver with geometric V(npre,npost) cycles
!> on the device until max|rhs - L phi| / |A| < tol (DESIGN.md "Multigrid").
module poisson_multigrid
  use iso_c_binding
  ! The boundary rule of the operator is the Poisson variant of the C-side schemes().  This module
  ! cannot bind it itself: src/Makefile compiles poisson_multigrid BEFORE initialization, whose
  ! flags it would need.  Its only caller, correct_pression (src/integration.f90:227-244), takes
  ! the divergence first -- through the diffoper shim, which binds the closures -- and the
  ! integration shim does not come through here at all (o3d_correct_pression runs the V-cycles).
  use o3d_b200_c
  implicit none
  private
  public :: solve_poisson_multigrid

contains

  subroutine solve_poisson_multigrid(phi, rhs, dx, dy, dz, nx, ny, nz, nlevels, npre, npost, tol)
    integer, intent(in) :: nx, ny, nz, nlevels, npre, npost
    real(kind=8), intent(inout) :: phi(nx, ny, nz)
    real(kind=8), intent(in) :: rhs(nx, ny, nz)
    real(kind=8), intent(in) :: dx, dy, dz, tol
    integer(c_int) :: cycles
    real(c_double) :: dmax
    call o3d_check(o3d_solve_poisson_multigrid(phi, rhs, dx, dy, dz, nx, ny, nz, nlevels, npre, &
         npost, tol, cycles, dmax), "solve_poisson_multigrid")
    print *, "* Multigrid: V-cycles, max|r|/|A| = ", cycles, dmax
  end subroutine solve_poisson_multigrid

end module poisson_multigrid
