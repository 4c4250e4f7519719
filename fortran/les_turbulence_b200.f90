!> les_turbulence_b200.f90 -- drop-in replacement of module `les_turbulence` (reference
!> src/les_turbulence.f90:10).  calculate_tau_ij / calculate_dtau_ij_dxj (:99,:179) have no
!> callers in the reference and are not provided.
module les_turbulence
  use iso_c_binding
  use IOfunctions
  use initialization
  use o3d_b200_c
  implicit none

contains

  subroutine calculate_nu_t(nu_t, ux, uy, uz, dx, dy, dz, cs, delta)
    real(kind=8), intent(out) :: nu_t(:,:,:)
    real(kind=8), intent(in) :: ux(:,:,:), uy(:,:,:), uz(:,:,:)
    real(kind=8), intent(in) :: dx, dy, dz, cs, delta
    real(c_double) :: st(6)
    call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
    call o3d_check(o3d_calculate_nu_t(nu_t, ux, uy, uz, dx, dy, dz, cs, delta, &
         int(size(ux,1),c_int), int(size(ux,2),c_int), int(size(ux,3),c_int), st), &
         "calculate_nu_t")
    ! function_stats(nu_t) + print, src/les_turbulence.f90:89-90: computed on the device
    call print_nu_t_statistics(st)
  end subroutine calculate_nu_t

end module les_turbulence
