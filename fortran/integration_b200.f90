!> integration_b200.f90 -- drop-in replacement of module `integration` (reference
!> src/integration.f90:14,199,257,332): predict_velocity, correct_pression, correct_velocity,
!> transeq with the reference's argument lists.
!>
!> Two modes, chosen by `o3d_resident`:
!>  .false. (pure drop-in): every call forwards its HOST arrays to the stateless C entry points
!>          (o3d_predict_velocity ...); inputs are copied to the device and outputs back on each
!>          call.  Bit-for-bit the reference's data flow; PCIe-bound (27 - 34 N doubles per step, INTEGRATION.md section 3).
!>  .true.  (resident, the intended production mode): the fields live in a device session created
!>          on the first call from the module `initialization` globals.  ux, uy, uz, pp, phi and
!>          the AB histories are uploaded ONCE.  What comes back per step is `o3d_mirror`:
!>            1 (default, works with the UNCHANGED driver): every array the reference's main loop
!>              reads after the step is refreshed on the host -- ux_pred, uy_pred, uz_pred
!>              (divergence(divu_pred, ...), osinco3d_main.f90:116), pp (visualize_2d,
!>              write_all_data, :133-150), nu_t, ux, uy, uz, phi: 8-9 N doubles D2H per step;
!>            0: nothing is mirrored; the driver must then take its prints from
!>              o3d_s_step_diagnostics and its files from fortran/output_b200.f90
!>              (INTEGRATION.md section 3 lists the ~10 lines of osinco3d_main.f90:116-183 to
!>              replace), and ux_pred / pp / nu_t / ux / uy / uz / phi on the host are STALE.
!>          The driver must not modify ux, uy, uz, pp, phi between steps
!>          (osinco3d_main.f90:97-188 does not).
module integration
  use iso_c_binding
  use initialization, only : nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d, sc0 => sc, &
       nscr0 => nscr
  use IOfunctions
  use o3d_b200_c
  implicit none

  logical :: o3d_resident = .true.
  integer :: o3d_mirror = 1
  type(c_ptr), private :: ses = c_null_ptr
  logical, private :: primed = .false.

contains

  !> the resident session, for the output shims of fortran/output_b200.f90
  function o3d_session_handle() result(h)
    type(c_ptr) :: h
    h = ses
  end function o3d_session_handle

  subroutine o3d_open_session(ux, uy, uz, fux, fuy, fuz, re, adt, bdt, cdt, itscheme, &
       dx, dy, dz, nx, ny, nz, iles, cs, delta)
    real(kind=8), intent(in) :: ux(:,:,:), uy(:,:,:), uz(:,:,:)
    real(kind=8), intent(in) :: fux(:,:,:,:), fuy(:,:,:,:), fuz(:,:,:,:)
    real(kind=8), intent(in) :: re, adt(3), bdt(3), cdt(3), dx, dy, dz, cs, delta
    integer, intent(in) :: itscheme, nx, ny, nz, iles
    type(o3d_config) :: c
    integer :: l
    if (o3d_config_size() /= int(c_sizeof(c), c_int)) then
       print *, "o3d_config layout mismatch between Fortran mirror and libo3d_b200.so"
       stop
    end if
    c%nx = nx; c%ny = ny; c%nz = nz
    c%dx = dx; c%dy = dy; c%dz = dz
    c%nbcx1 = nbcx1; c%nbcxn = nbcxn; c%nbcy1 = nbcy1; c%nbcyn = nbcyn
    c%nbcz1 = nbcz1; c%nbczn = nbczn; c%sim2d = sim2d
    c%re = re; c%sc = sc0; c%cs = cs; c%delta = delta
    c%dt = adt(1)                       ! adt(1) = dt, src/initialization.f90:194
    c%adt = adt; c%bdt = bdt; c%cdt = cdt
    c%itscheme = itscheme; c%iles = iles
    c%nscr = nscr0                      ! Scalar namelist (src/initialization.f90:62,122): the
                                        ! session's o3d_step / write_all_data need the real value
    c%omega = 1.d0; c%eps = 1.d-6; c%kmax = 1; c%idyn = 0; c%multigrid = 0   ! set per call below
    c%sor_order = 0; c%sor_check_every = 0
    c%rank = 0; c%nranks = 1
    c%nccl_id = 0_c_signed_char
    c%reserved = 0
    call o3d_check(o3d_session_create(c, ses), "o3d_session_create")
    call o3d_check(o3d_upload(ses, O3D_F_UX, ux), "upload ux")
    call o3d_check(o3d_upload(ses, O3D_F_UY, uy), "upload uy")
    call o3d_check(o3d_upload(ses, O3D_F_UZ, uz), "upload uz")
    do l = 2, 3   ! history levels 2,3 are inputs; level 1 is overwritten (src/integration.f90:129)
       call o3d_check(o3d_upload(ses, O3D_F_FUX1 + l - 1, fux(:,:,:,l)), "upload fux")
       call o3d_check(o3d_upload(ses, O3D_F_FUY1 + l - 1, fuy(:,:,:,l)), "upload fuy")
       call o3d_check(o3d_upload(ses, O3D_F_FUZ1 + l - 1, fuz(:,:,:,l)), "upload fuz")
    end do
  end subroutine o3d_open_session

  subroutine predict_velocity(ux_pred, uy_pred, uz_pred, ux, uy, uz, &
       fux, fuy, fuz, re, adt, bdt, cdt, itime, itscheme, &
       dx, dy, dz, nx, ny, nz, iles, cs, delta, nu_t)
    ! (intents as in the reference, src/integration.f90:53-61: ux, uy, uz are intent(inout)
    ! there although predict_velocity never assigns them)
    real(kind=8), intent(inout) :: ux(:,:,:), uy(:,:,:), uz(:,:,:)
    real(kind=8), intent(inout) :: fux(:,:,:,:), fuy(:,:,:,:), fuz(:,:,:,:)
    real(kind=8), intent(in) :: re, adt(3), bdt(3), cdt(3)
    real(kind=8), intent(in) :: dx, dy, dz, cs, delta
    integer, intent(in) :: itime, itscheme, nx, ny, nz, iles
    real(kind=8), intent(out) :: ux_pred(:,:,:), uy_pred(:,:,:), uz_pred(:,:,:)
    real(kind=8), intent(inout) :: nu_t(:,:,:)
    print *, "* Predict velocity"
    if (.not. o3d_resident) then
       call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
       call o3d_check(o3d_predict_velocity(ux_pred, uy_pred, uz_pred, ux, uy, uz, fux, fuy, fuz, &
            re, adt, bdt, cdt, itime, itscheme, dx, dy, dz, nx, ny, nz, iles, cs, delta, nu_t), &
            "predict_velocity")
       return
    end if
    if (.not. c_associated(ses)) then
       call o3d_open_session(ux, uy, uz, fux, fuy, fuz, re, adt, bdt, cdt, itscheme, &
            dx, dy, dz, nx, ny, nz, iles, cs, delta)
    end if
    call o3d_check(o3d_s_predict_velocity(ses, itime), "predict_velocity")
    if (o3d_mirror /= 0) then
       ! intent(out) arrays the unchanged driver reads at osinco3d_main.f90:116-119
       call o3d_check(o3d_download(ses, O3D_F_UX_PRED, ux_pred), "download ux_pred")
       call o3d_check(o3d_download(ses, O3D_F_UY_PRED, uy_pred), "download uy_pred")
       call o3d_check(o3d_download(ses, O3D_F_UZ_PRED, uz_pred), "download uz_pred")
       if (iles == 1) call o3d_check(o3d_download(ses, O3D_F_NU_T, nu_t), "download nu_t")
    end if
  end subroutine predict_velocity

  subroutine correct_pression(pp, ux_pred, uy_pred, uz_pred, dx, dy, dz, &
       nx, ny, nz, dt, omega, eps, kmax, idyn, multigrid)
    real(kind=8), intent(inout) :: pp(:,:,:), omega
    real(kind=8), intent(in) :: ux_pred(:,:,:), uy_pred(:,:,:), uz_pred(:,:,:)
    real(kind=8), intent(in) :: dx, dy, dz, dt, eps
    integer, intent(in) :: nx, ny, nz, kmax, idyn, multigrid
    integer(c_int) :: iters
    real(c_double) :: dmax
    print *, "* Correction pression"
    if (.not. o3d_resident) then
       call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
       call o3d_check(o3d_correct_pression(pp, ux_pred, uy_pred, uz_pred, dx, dy, dz, nx, ny, nz, &
            dt, omega, eps, kmax, idyn, multigrid, iters, dmax), "correct_pression")
       return
    end if
    if (.not. primed) then
       ! first step: pp is the initial guess (src/integration.f90:247); Poisson controls are
       ! per-call arguments in the reference, so they are pushed into the session here
       call o3d_check(o3d_upload(ses, O3D_F_PP, pp), "upload pp")
       call o3d_check(o3d_set_omega(ses, omega), "set omega")
       call o3d_check(o3d_session_set_poisson(ses, eps, kmax, idyn, multigrid), "set poisson")
       primed = .true.
    end if
    call o3d_check(o3d_s_correct_pression(ses, iters, dmax), "correct_pression")
    call o3d_check(o3d_get_omega(ses, omega), "get omega")   ! omega is intent(inout), :222
    ! pp is intent(inout): visualize_2d / write_all_data read it (osinco3d_main.f90:133-150)
    if (o3d_mirror /= 0) call o3d_check(o3d_download(ses, O3D_F_PP, pp), "download pp")
  end subroutine correct_pression

  subroutine correct_velocity(ux, uy, uz, ux_pred, uy_pred, uz_pred, pp, dt, dx, dy, dz, nx, ny, nz)
    real(kind=8), intent(in) :: ux_pred(:,:,:), uy_pred(:,:,:), uz_pred(:,:,:), pp(:,:,:)
    real(kind=8), intent(in) :: dt, dx, dy, dz
    integer, intent(in) :: nx, ny, nz
    ! intent(inout) as in the reference (src/integration.f90:287); with o3d_mirror = 0 the host
    ! copies are deliberately left as they are
    real(kind=8), intent(inout) :: ux(:,:,:), uy(:,:,:), uz(:,:,:)
    integer(c_int) :: rc
    print *, "* Correct velocity"
    if (.not. o3d_resident) then
       call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
       rc = o3d_correct_velocity(ux, uy, uz, ux_pred, uy_pred, uz_pred, pp, dt, dx, dy, dz, &
            nx, ny, nz)
    else
       rc = o3d_s_correct_velocity(ses)
       if (o3d_mirror /= 0) then
          ! the new velocity for the driver's prints / output (3 N doubles D2H per step)
          call o3d_check(o3d_download(ses, O3D_F_UX, ux), "download ux")
          call o3d_check(o3d_download(ses, O3D_F_UY, uy), "download uy")
          call o3d_check(o3d_download(ses, O3D_F_UZ, uz), "download uz")
       end if
    end if
    if (rc == O3D_ERR_DIVERGED) then
       ! src/integration.f90:309-325: NaN or max(u) > 1000 -> report and stop
       call write_velocity_diverged()
       stop
    end if
    call o3d_check(rc, "correct_velocity")
  end subroutine correct_velocity

  subroutine transeq(phi, ux, uy, uz, src, fphi, re, sc, adt, bdt, cdt, &
       itime, itscheme, dx, dy, dz, nx, ny, nz, iles, nu_t)
    real(kind=8), intent(inout) :: phi(:,:,:), fphi(:,:,:,:)
    real(kind=8), intent(in) :: ux(:,:,:), uy(:,:,:), uz(:,:,:), src(:,:,:), nu_t(:,:,:)
    real(kind=8), intent(in) :: re, sc, adt(3), bdt(3), cdt(3), dx, dy, dz
    integer, intent(in) :: itime, itscheme, nx, ny, nz, iles
    logical, save :: phi_up = .false.
    if (.not. o3d_resident) then
       call o3d_bind_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d)
       call o3d_check(o3d_transeq(phi, ux, uy, uz, src, fphi, re, sc, adt, bdt, cdt, itime, &
            itscheme, dx, dy, dz, nx, ny, nz, iles, nu_t), "transeq")
       return
    end if
    if (.not. phi_up) then
       call o3d_check(o3d_upload(ses, O3D_F_PHI, phi), "upload phi")
       call o3d_check(o3d_upload(ses, O3D_F_FPHI1 + 1, fphi(:,:,:,2)), "upload fphi2")
       call o3d_check(o3d_upload(ses, O3D_F_FPHI1 + 2, fphi(:,:,:,3)), "upload fphi3")
       phi_up = .true.
    end if
    call o3d_check(o3d_s_transeq(ses, itime), "transeq")
    if (o3d_mirror /= 0) call o3d_check(o3d_download(ses, O3D_F_PHI, phi), "download phi")
  end subroutine transeq

end module integration
