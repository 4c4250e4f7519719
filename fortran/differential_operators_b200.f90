!> differential_operators_b200.f90 -- drop-in replacement of module `diffoper` (reference
!> src/differential_operators.f90:7,40,79): divergence, rotational, calculate_Q_criterion.
!> One fused CUDA launch each instead of 3 / 6 / 9 derivative sweeps plus temporaries.
module diffoper
  use iso_c_binding
  use initialization
  use o3d_b200_c
  implicit none

contains

  subroutine divergence(divf, fx, fy, fz, dx, dy, dz, nx, ny, nz, odd)
    integer, intent(in) :: nx, ny, nz, odd
    real(kind=8), intent(in) :: fx(nx,ny,nz), fy(nx,ny,nz), fz(nx,ny,nz)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(out) :: divf(nx,ny,nz)
    call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
    call o3d_check(o3d_divergence(divf, fx, fy, fz, dx, dy, dz, nx, ny, nz, odd), "divergence")
  end subroutine divergence

  subroutine rotational(rotx, roty, rotz, ux, uy, uz, dx, dy, dz, nx, ny, nz)
    integer, intent(in) :: nx, ny, nz
    real(kind=8), intent(in) :: ux(nx,ny,nz), uy(nx,ny,nz), uz(nx,ny,nz)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(out) :: rotx(nx,ny,nz), roty(nx,ny,nz), rotz(nx,ny,nz)
    call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
    call o3d_check(o3d_rotational(rotx, roty, rotz, ux, uy, uz, dx, dy, dz, nx, ny, nz), &
         "rotational")
  end subroutine rotational

  subroutine calculate_Q_criterion(Q, ux, uy, uz, dx, dy, dz, nx, ny, nz)
    integer, intent(in) :: nx, ny, nz
    real(kind=8), intent(in) :: ux(nx,ny,nz), uy(nx,ny,nz), uz(nx,ny,nz)
    real(kind=8), intent(in) :: dx, dy, dz
    real(kind=8), intent(out) :: Q(nx,ny,nz)
    call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
    call o3d_check(o3d_calculate_q_criterion(Q, ux, uy, uz, dx, dy, dz, nx, ny, nz), &
         "calculate_Q_criterion")
  end subroutine calculate_Q_criterion

end module diffoper
