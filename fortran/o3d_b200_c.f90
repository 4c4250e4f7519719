!> o3d_b200_c.f90 -- ISO_C_BINDING interface blocks for libo3d_b200.so (include/o3d_b200.h).
!>
!> This module is the ONLY thing the Fortran side needs to know about the CUDA library.  It is
!> used by the drop-in replacement modules in this directory (derivation_b200.f90,
!> differential_operators_b200.f90, les_turbulence_b200.f90, poisson_b200.f90,
!> poisson_multigrid_b200.f90, integration_b200.f90), which keep the reference's module names,
!> procedure names and argument lists, so osinco3d_main.f90, initialization.f90,
!> initial_conditions.f90, utils.f90, IOfunctions.f90 and visualization.f90 compile unchanged.
!>
!> NOTE: the build container has no Fortran compiler, so these sources are shipped untested by a
!> compiler; the same C entry points are exercised from ctypes by tests/ (INTEGRATION.md).
module o3d_b200_c
  use iso_c_binding
  implicit none

  integer(c_int), parameter :: O3D_OK = 0, O3D_ERR_DIVERGED = 6

  !> mirror of `struct o3d_config` (include/o3d_b200.h); o3d_config_size() lets the shim verify
  !> that the two layouts agree before the first call
  type, bind(C) :: o3d_config
     integer(c_int) :: nx, ny, nz
     real(c_double) :: dx, dy, dz
     integer(c_int) :: nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d
     real(c_double) :: re, sc, cs, delta
     real(c_double) :: dt
     real(c_double) :: adt(3), bdt(3), cdt(3)
     integer(c_int) :: itscheme, iles, nscr
     real(c_double) :: omega, eps
     integer(c_int) :: kmax, idyn, multigrid
     integer(c_int) :: sor_order
     integer(c_int) :: sor_check_every
     integer(c_int) :: rank, nranks
     integer(c_signed_char) :: nccl_id(128)
     integer(c_int) :: reserved(8)
  end type o3d_config

  !> field ids (enum in include/o3d_b200.h)
  integer(c_int), parameter :: O3D_F_UX = 0, O3D_F_UY = 1, O3D_F_UZ = 2, O3D_F_PP = 3, &
       O3D_F_PHI = 4, O3D_F_UX_PRED = 5, O3D_F_UY_PRED = 6, O3D_F_UZ_PRED = 7, O3D_F_NU_T = 8, &
       O3D_F_RHS = 9, O3D_F_FUX1 = 10, O3D_F_FUY1 = 13, O3D_F_FUZ1 = 16, O3D_F_FPHI1 = 19, &
       O3D_F_DIVU = 22
  integer(c_int), parameter :: O3D_RED_MIN = 0, O3D_RED_MAX = 1, O3D_RED_SUM = 2, &
       O3D_RED_ABSMAX = 3

  interface
     !--- error text ---------------------------------------------------------------------
     function o3d_last_error() bind(C, name="o3d_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function o3d_last_error

     !--- schemes(), src/initialization.f90:226-304 ---------------------------------------
     function o3d_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d) &
          bind(C, name="o3d_schemes") result(rc)
       import :: c_int
       integer(c_int), value :: nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d
       integer(c_int) :: rc
     end function o3d_schemes

     !--- der_type, src/initialization.f90:86-91 (generic form) ---------------------------
     function o3d_der(axis, order, closure, df, f, d, nx, ny, nz) &
          bind(C, name="o3d_der") result(rc)
       import :: c_int, c_double
       integer(c_int), value :: axis, order, closure, nx, ny, nz
       real(c_double), value :: d
       real(c_double), intent(out) :: df(*)
       real(c_double), intent(in) :: f(*)
       integer(c_int) :: rc
     end function o3d_der

     !--- divergence, src/differential_operators.f90:7 -----------------------------------
     function o3d_divergence(divf, fx, fy, fz, dx, dy, dz, nx, ny, nz, odd) &
          bind(C, name="o3d_divergence") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: divf(*)
       real(c_double), intent(in) :: fx(*), fy(*), fz(*)
       real(c_double), value :: dx, dy, dz
       integer(c_int), value :: nx, ny, nz, odd
       integer(c_int) :: rc
     end function o3d_divergence

     function o3d_rotational(rotx, roty, rotz, ux, uy, uz, dx, dy, dz, nx, ny, nz) &
          bind(C, name="o3d_rotational") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: rotx(*), roty(*), rotz(*)
       real(c_double), intent(in) :: ux(*), uy(*), uz(*)
       real(c_double), value :: dx, dy, dz
       integer(c_int), value :: nx, ny, nz
       integer(c_int) :: rc
     end function o3d_rotational

     function o3d_calculate_q_criterion(q, ux, uy, uz, dx, dy, dz, nx, ny, nz) &
          bind(C, name="o3d_calculate_q_criterion") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: q(*)
       real(c_double), intent(in) :: ux(*), uy(*), uz(*)
       real(c_double), value :: dx, dy, dz
       integer(c_int), value :: nx, ny, nz
       integer(c_int) :: rc
     end function o3d_calculate_q_criterion

     !--- calculate_nu_t, src/les_turbulence.f90:10 ---------------------------------------
     function o3d_calculate_nu_t(nu_t, ux, uy, uz, dx, dy, dz, cs, delta, nx, ny, nz, stats6) &
          bind(C, name="o3d_calculate_nu_t") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: nu_t(*)
       real(c_double), intent(in) :: ux(*), uy(*), uz(*)
       real(c_double), value :: dx, dy, dz, cs, delta
       integer(c_int), value :: nx, ny, nz
       real(c_double), intent(out) :: stats6(6)
       integer(c_int) :: rc
     end function o3d_calculate_nu_t

     !--- predict_velocity, src/integration.f90:14-16 -------------------------------------
     function o3d_predict_velocity(ux_pred, uy_pred, uz_pred, ux, uy, uz, fux, fuy, fuz, re, &
          adt, bdt, cdt, itime, itscheme, dx, dy, dz, nx, ny, nz, iles, cs, delta, nu_t) &
          bind(C, name="o3d_predict_velocity") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: ux_pred(*), uy_pred(*), uz_pred(*)
       real(c_double), intent(in) :: ux(*), uy(*), uz(*)
       real(c_double), intent(inout) :: fux(*), fuy(*), fuz(*)
       real(c_double), value :: re, dx, dy, dz, cs, delta
       real(c_double), intent(in) :: adt(3), bdt(3), cdt(3)
       integer(c_int), value :: itime, itscheme, nx, ny, nz, iles
       real(c_double), intent(inout) :: nu_t(*)
       integer(c_int) :: rc
     end function o3d_predict_velocity

     !--- poi_type, src/initialization.f90:93-102; variant = the bound pointer ------------
     function o3d_poisson_solver(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, &
          iters, dmax) bind(C, name="o3d_poisson_solver") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: pp(*), omega
       real(c_double), intent(in) :: rhs(*)
       real(c_double), value :: dx, dy, dz, eps
       integer(c_int), value :: nx, ny, nz, kmax, idyn
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_poisson_solver
     function o3d_poisson_solver_0000(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, &
          iters, dmax) bind(C, name="o3d_poisson_solver_0000") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: pp(*), omega
       real(c_double), intent(in) :: rhs(*)
       real(c_double), value :: dx, dy, dz, eps
       integer(c_int), value :: nx, ny, nz, kmax, idyn
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_poisson_solver_0000
     function o3d_poisson_solver_0011(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, &
          iters, dmax) bind(C, name="o3d_poisson_solver_0011") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: pp(*), omega
       real(c_double), intent(in) :: rhs(*)
       real(c_double), value :: dx, dy, dz, eps
       integer(c_int), value :: nx, ny, nz, kmax, idyn
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_poisson_solver_0011
     function o3d_poisson_solver_111111(pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, &
          idyn, iters, dmax) bind(C, name="o3d_poisson_solver_111111") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: pp(*), omega
       real(c_double), intent(in) :: rhs(*)
       real(c_double), value :: dx, dy, dz, eps
       integer(c_int), value :: nx, ny, nz, kmax, idyn
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_poisson_solver_111111

     !--- solve_poisson_multigrid, src/poisson_multigrid.f90:10 ---------------------------
     function o3d_solve_poisson_multigrid(phi, rhs, dx, dy, dz, nx, ny, nz, nlevels, npre, &
          npost, tol, cycles, dmax) bind(C, name="o3d_solve_poisson_multigrid") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: phi(*)
       real(c_double), intent(in) :: rhs(*)
       real(c_double), value :: dx, dy, dz, tol
       integer(c_int), value :: nx, ny, nz, nlevels, npre, npost
       integer(c_int), intent(out) :: cycles
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_solve_poisson_multigrid

     !--- correct_pression / correct_velocity / transeq, src/integration.f90:199,257,332 --
     function o3d_correct_pression(pp, ux_pred, uy_pred, uz_pred, dx, dy, dz, nx, ny, nz, dt, &
          omega, eps, kmax, idyn, multigrid, iters, dmax) &
          bind(C, name="o3d_correct_pression") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: pp(*), omega
       real(c_double), intent(in) :: ux_pred(*), uy_pred(*), uz_pred(*)
       real(c_double), value :: dx, dy, dz, dt, eps
       integer(c_int), value :: nx, ny, nz, kmax, idyn, multigrid
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_correct_pression

     function o3d_correct_velocity(ux, uy, uz, ux_pred, uy_pred, uz_pred, pp, dt, dx, dy, dz, &
          nx, ny, nz) bind(C, name="o3d_correct_velocity") result(rc)
       import :: c_int, c_double
       real(c_double), intent(out) :: ux(*), uy(*), uz(*)
       real(c_double), intent(in) :: ux_pred(*), uy_pred(*), uz_pred(*), pp(*)
       real(c_double), value :: dt, dx, dy, dz
       integer(c_int), value :: nx, ny, nz
       integer(c_int) :: rc
     end function o3d_correct_velocity

     function o3d_transeq(phi, ux, uy, uz, src, fphi, re, sc, adt, bdt, cdt, itime, itscheme, &
          dx, dy, dz, nx, ny, nz, iles, nu_t) bind(C, name="o3d_transeq") result(rc)
       import :: c_int, c_double
       real(c_double), intent(inout) :: phi(*), fphi(*)
       real(c_double), intent(in) :: ux(*), uy(*), uz(*), src(*), nu_t(*)
       real(c_double), value :: re, sc, dx, dy, dz
       real(c_double), intent(in) :: adt(3), bdt(3), cdt(3)
       integer(c_int), value :: itime, itscheme, nx, ny, nz, iles
       integer(c_int) :: rc
     end function o3d_transeq

     !--- statistics_calc (src/utils.f90:243) and function_stats (src/functions.f90:27) ---
     function o3d_statistics_calc(ux, uy, uz, nx, ny, nz, dx, dy, dz, re, t, out17) &
          bind(C, name="o3d_statistics_calc") result(rc)
       import :: c_int, c_double
       real(c_double), intent(in) :: ux(*), uy(*), uz(*)
       integer(c_int), value :: nx, ny, nz
       real(c_double), value :: dx, dy, dz, re, t
       real(c_double), intent(out) :: out17(17)
       integer(c_int) :: rc
     end function o3d_statistics_calc
     function o3d_function_stats(f, nx, ny, nz, stats6) &
          bind(C, name="o3d_function_stats") result(rc)
       import :: c_int, c_double
       real(c_double), intent(in) :: f(*)
       integer(c_int), value :: nx, ny, nz
       real(c_double), intent(out) :: stats6(6)
       integer(c_int) :: rc
     end function o3d_function_stats

     !> page-lock a host array so that the copies of the stateless procedures run at full PCIe
     !> speed and overlap (copy pipeline, DESIGN.md 4.7): call once after `allocate`, e.g.
     !>   rc = o3d_host_register(c_loc(ux), int(8, c_long_long) * size(ux, kind=c_long_long))
     function o3d_host_register(ptr, bytes) bind(C, name="o3d_host_register") result(rc)
       import :: c_int, c_ptr, c_long_long
       type(c_ptr), value :: ptr
       integer(c_long_long), value :: bytes
       integer(c_int) :: rc
     end function o3d_host_register
     function o3d_host_unregister(ptr) bind(C, name="o3d_host_unregister") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ptr
       integer(c_int) :: rc
     end function o3d_host_unregister
     !> z chunks of the pipelined predict_velocity / correct_velocity (0 = off, default 16)
     function o3d_set_pipeline(chunks) bind(C, name="o3d_set_pipeline") result(rc)
       import :: c_int
       integer(c_int), value :: chunks
       integer(c_int) :: rc
     end function o3d_set_pipeline
     !> 1: the history-level copies at the end of predict_velocity (fu?(:,:,:,3) = fu?(:,:,:,2),
     !> fu?(:,:,:,2) = fu?(:,:,:,1)) are made in the host arrays instead of being downloaded
     function o3d_set_hostshift(on) bind(C, name="o3d_set_hostshift") result(rc)
       import :: c_int
       integer(c_int), value :: on
       integer(c_int) :: rc
     end function o3d_set_hostshift

     !=== section B of include/o3d_b200.h: device-resident session ==========================
     function o3d_session_create(cfg, ses) bind(C, name="o3d_session_create") result(rc)
       import :: c_int, c_ptr, o3d_config
       type(o3d_config), intent(in) :: cfg
       type(c_ptr), intent(out) :: ses
       integer(c_int) :: rc
     end function o3d_session_create
     function o3d_session_destroy(ses) bind(C, name="o3d_session_destroy") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int) :: rc
     end function o3d_session_destroy
     function o3d_config_size() bind(C, name="o3d_config_size") result(n)
       import :: c_int
       integer(c_int) :: n
     end function o3d_config_size
     function o3d_upload(ses, field, host) bind(C, name="o3d_upload") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field
       real(c_double), intent(in) :: host(*)
       integer(c_int) :: rc
     end function o3d_upload
     function o3d_download(ses, field, host) bind(C, name="o3d_download") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field
       real(c_double), intent(out) :: host(*)
       integer(c_int) :: rc
     end function o3d_download
     ! local planes [k0, k0+nk) only (0-based k0): a (nx,ny,nk) block, e.g. ux(:,:,k0+1:k0+nk)
     function o3d_upload_planes(ses, field, host, k0, nk) bind(C, name="o3d_upload_planes") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field, k0, nk
       real(c_double), intent(in) :: host(*)
       integer(c_int) :: rc
     end function o3d_upload_planes
     function o3d_download_planes(ses, field, host, k0, nk) bind(C, name="o3d_download_planes") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field, k0, nk
       real(c_double), intent(out) :: host(*)
       integer(c_int) :: rc
     end function o3d_download_planes
     ! how the last red-black solve ran: persistent kernel / peer-memory halos (1 / 0 each)
     function o3d_s_sor_path(ses, persistent, peer) bind(C, name="o3d_s_sor_path") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int), intent(out) :: persistent, peer
       integer(c_int) :: rc
     end function o3d_s_sor_path
     function o3d_s_predict_velocity(ses, itime) bind(C, name="o3d_s_predict_velocity") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int), value :: itime
       integer(c_int) :: rc
     end function o3d_s_predict_velocity
     function o3d_s_correct_pression(ses, iters, dmax) &
          bind(C, name="o3d_s_correct_pression") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: dmax
       integer(c_int) :: rc
     end function o3d_s_correct_pression
     function o3d_s_correct_velocity(ses) bind(C, name="o3d_s_correct_velocity") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int) :: rc
     end function o3d_s_correct_velocity
     function o3d_s_transeq(ses, itime) bind(C, name="o3d_s_transeq") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int), value :: itime
       integer(c_int) :: rc
     end function o3d_s_transeq
     function o3d_s_divergence(ses, fx, fy, fz, dst, odd) &
          bind(C, name="o3d_s_divergence") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int), value :: fx, fy, fz, dst, odd
       integer(c_int) :: rc
     end function o3d_s_divergence
     function o3d_s_function_stats(ses, field, stats6) &
          bind(C, name="o3d_s_function_stats") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field
       real(c_double), intent(out) :: stats6(6)
       integer(c_int) :: rc
     end function o3d_s_function_stats
     function o3d_s_reduce(ses, field, op, res) bind(C, name="o3d_s_reduce") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       integer(c_int), value :: field, op
       real(c_double), intent(out) :: res
       integer(c_int) :: rc
     end function o3d_s_reduce
     function o3d_s_statistics(ses, t, out17) bind(C, name="o3d_s_statistics") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), value :: t
       real(c_double), intent(out) :: out17(17)
       integer(c_int) :: rc
     end function o3d_s_statistics
     function o3d_get_omega(ses, omega) bind(C, name="o3d_get_omega") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), intent(out) :: omega
       integer(c_int) :: rc
     end function o3d_get_omega
     function o3d_set_omega(ses, omega) bind(C, name="o3d_set_omega") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), value :: omega
       integer(c_int) :: rc
     end function o3d_set_omega
     !> the Poisson controls are per-call arguments of correct_pression in the reference
     !> (src/integration.f90:199-200); the resident shim pushes them into the session with this
     function o3d_session_set_poisson(ses, eps, kmax, idyn, multigrid) &
          bind(C, name="o3d_session_set_poisson") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), value :: eps
       integer(c_int), value :: kmax, idyn, multigrid
       integer(c_int) :: rc
     end function o3d_session_set_poisson
     function o3d_s_step_diagnostics(ses, out23) bind(C, name="o3d_s_step_diagnostics") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), intent(out) :: out23(23)
       integer(c_int) :: rc
     end function o3d_s_step_diagnostics
     function o3d_s_old_values(ses) bind(C, name="o3d_s_old_values") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int) :: rc
     end function o3d_s_old_values
     function o3d_s_calculate_residuals(ses, dt, t_ref, u_ref, out15) &
          bind(C, name="o3d_s_calculate_residuals") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ses
       real(c_double), value :: dt, t_ref, u_ref
       real(c_double), intent(out) :: out15(15)
       integer(c_int) :: rc
     end function o3d_s_calculate_residuals
     function o3d_s_vorticity_magnitude(ses, dst) &
          bind(C, name="o3d_s_vorticity_magnitude") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int), value :: dst
       integer(c_int) :: rc
     end function o3d_s_vorticity_magnitude
     !--- field output in the reference's binary formats (asynchronous; o3d_s_io_wait drains) ---
     function o3d_s_save_fields(ses, filename, time, x, y, z) &
          bind(C, name="o3d_s_save_fields") result(rc)
       import :: c_int, c_ptr, c_double, c_char
       type(c_ptr), value :: ses
       character(kind=c_char), intent(in) :: filename(*)
       real(c_double), value :: time
       real(c_double), intent(in) :: x(*), y(*), z(*)
       integer(c_int) :: rc
     end function o3d_s_save_fields
     function o3d_s_read_fields(ses, filename, time, x, y, z) &
          bind(C, name="o3d_s_read_fields") result(rc)
       import :: c_int, c_ptr, c_double, c_char
       type(c_ptr), value :: ses
       character(kind=c_char), intent(in) :: filename(*)
       real(c_double), intent(inout) :: time
       real(c_double), intent(inout) :: x(*), y(*), z(*)
       integer(c_int) :: rc
     end function o3d_s_read_fields
     function o3d_s_write_binary(ses, filename, field) &
          bind(C, name="o3d_s_write_binary") result(rc)
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ses
       character(kind=c_char), intent(in) :: filename(*)
       integer(c_int), value :: field
       integer(c_int) :: rc
     end function o3d_s_write_binary
     function o3d_s_write_all_data(ses, dir, num) &
          bind(C, name="o3d_s_write_all_data") result(rc)
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ses
       character(kind=c_char), intent(in) :: dir(*)
       integer(c_int), value :: num
       integer(c_int) :: rc
     end function o3d_s_write_all_data
     function o3d_s_io_wait(ses) bind(C, name="o3d_s_io_wait") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ses
       integer(c_int) :: rc
     end function o3d_s_io_wait
  end interface

contains

  !> The reference reports errors with print + stop; every shim funnels its status through here.
  subroutine o3d_check(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: cmsg(:)
    character(len=1024) :: msg
    integer :: i
    if (rc == O3D_OK) return
    call c_f_pointer(o3d_last_error(), cmsg, [1024])
    msg = ""
    do i = 1, 1024
       if (cmsg(i) == c_null_char) exit
       msg(i:i) = cmsg(i)
    end do
    print *, "libo3d_b200 error in ", where, ": status ", rc, " ", trim(msg)
    stop
  end subroutine o3d_check

  !> C mirror of schemes() (src/initialization.f90:226-304): binds the closures of the composite
  !> operators and the Poisson variant on the C side to the boundary flags of module
  !> `initialization`.  The stateless shims call it before every composite operator (a few host
  !> instructions), so the unchanged driver needs no extra line after its own `call schemes()`.
  subroutine o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
    integer, intent(in) :: nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d
    call o3d_check(o3d_schemes(int(nbcx1, c_int), int(nbcxn, c_int), int(nbcy1, c_int), &
         int(nbcyn, c_int), int(nbcz1, c_int), int(nbczn, c_int), int(sim2d, c_int)), "schemes")
  end subroutine o3d_bind_schemes

end module o3d_b200_c
