!> output_b200.f90 -- device-side replacements of the reference's field writers, same argument
!> lists (so the call sites in src/osinco3d_main.f90:68,87,148,183 compile unchanged):
!>   save_fields     src/IOfunctions.f90:360-402    restart file fields_NNNNNN.bin
!>   write_all_data  src/visualization.f90:243-276  outputs/<name>_<num>.bin
!> In resident mode the device holds the authoritative state, so the host arrays in the argument
!> lists are NOT read: the library snapshots the fields on the GPU, copies them out on a separate
!> stream and a writer thread does the file I/O while the time loop continues (vort and qcrit are
!> computed on the device; the rotational / calculate_Q_criterion calls before write_all_data at
!> src/osinco3d_main.f90:130-133 can then be dropped).  write_xdmf (text metadata) is unchanged.
!> Call o3d_output_wait() before reading the files back or at the end of the run.
module output_b200
  use iso_c_binding
  use o3d_b200_c
  use integration, only : o3d_session_handle
  implicit none

contains

  subroutine save_fields(x, y, z, ux, uy, uz, pp, phi, nx, ny, nz, time, itime)
    integer, intent(in) :: nx, ny, nz, itime
    real(kind=8), intent(in) :: x(nx), y(ny), z(nz)
    real(kind=8), intent(in) :: ux(nx,ny,nz), uy(nx,ny,nz), uz(nx,ny,nz), pp(nx,ny,nz), phi(nx,ny,nz)
    real(kind=8), intent(in) :: time
    character(len=30) :: filename
    write(filename, '(A,I0.6,A)') "fields_", itime, ".bin"
    print *, "* Save flow state in: ", filename
    call o3d_check(o3d_s_save_fields(o3d_session_handle(), trim(filename)//c_null_char, time, &
         x, y, z), "save_fields")
  end subroutine save_fields

  subroutine write_all_data(ux, uy, uz, rotx, roty, rotz, qcriterion, pp, phi, nu_t, num, nscr, iles)
    real(kind=8), intent(in) :: ux(:,:,:), uy(:,:,:), uz(:,:,:), pp(:,:,:)
    real(kind=8), intent(in) :: rotx(:,:,:), roty(:,:,:), rotz(:,:,:)
    real(kind=8), intent(in) :: qcriterion(:,:,:), nu_t(:,:,:), phi(:,:,:)
    integer, intent(in) :: num, nscr, iles
    ! nscr / iles were given to the session at creation (o3d_config)
    call o3d_check(o3d_s_write_all_data(o3d_session_handle(), "outputs"//c_null_char, num), &
         "write_all_data")
  end subroutine write_all_data

  subroutine o3d_output_wait()
    call o3d_check(o3d_s_io_wait(o3d_session_handle()), "output wait")
  end subroutine o3d_output_wait

end module output_b200
