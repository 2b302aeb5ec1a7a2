!=======================================================================
! ice_dyn_evp_b200.F90 -- drop-in replacement for source/ice_dyn_evp.F90 of
! COSIMA/cice4 that runs the EVP dynamics on an NVIDIA B200 through the C ABI
! of libevp_b200.so (include/evp_b200.h).
!
! The module keeps the reference module's name and everything other units
! import from it: kdyn, ndte, evp_damping, yield_curve (source/ice_init.F90:91),
! dragio, cosw, sinw (:97, AusCOM builds), evp, init_evp, principal_stress,
! set_evp_parameters.  ice_step_mod, ice_history, ice_restart and the
! cice4 / access-om / access-cm drivers are compiled unchanged.
!
! What stays in Fortran: the module state (ice_state, ice_flux, ice_grid
! arrays), ice_strength (source/ice_mechred.F90:1869) between the two device
! phases, the timers.  Everything else of evp() runs on the device.
!
! NOTE: no Fortran compiler exists in the image this repository is built in;
! this file is the integration recipe of INTEGRATION.md and has not been
! compiled there.  The Python mirror cice4_b200/evp.py binds the same entry
! points and is what the parity tests drive.
!
! Build (sketch):  <fc> -c ice_dyn_evp_b200.F90 ...  ; link with -levp_b200
! CPP: AusCOM / ACCICE / coupled / ACCESS select the same variants as in the
! reference (bld/Macros.nci:56-82); they become run-time flags of the library.
!=======================================================================
      module ice_dyn_evp

      use, intrinsic :: iso_c_binding
      use ice_kinds_mod
      use ice_fileunits
      use ice_communicate, only: my_task, master_task, get_num_procs
      use ice_domain_size
      use ice_constants
#ifdef AusCOM
      use cpl_parameters
      use cpl_arrays_setup, only : sicemass
#endif

      implicit none
      save

      ! namelist parameters (same names as the reference module)
      integer (kind=int_kind) :: kdyn, ndte
      logical (kind=log_kind) :: evp_damping
      character (len=char_len) :: yield_curve
#if defined(AusCOM) || defined(ACCICE)
      real (kind=dbl_kind) :: dragio, cosw, sinw
#else
      real (kind=dbl_kind), parameter :: cosw = c1, sinw = c0
#endif

      ! ---- mirrors of the structs in include/evp_b200.h ----------------
      type, bind(C) :: evp_b200_dims
         integer(c_int32_t) :: nx_block, ny_block, max_blocks, nblocks
         integer(c_int32_t) :: nx_global, ny_global, ew_boundary, ns_boundary
         type(c_ptr) :: ilo, ihi, jlo, jhi, iglob_lo, jglob_lo
         integer(c_int32_t) :: slab_jlo, slab_jhi, rank, nranks, device
      end type

      type, bind(C) :: evp_b200_params
         real(c_double) :: dt
         integer(c_int32_t) :: ndte, evp_damping
         real(c_double) :: dragio, cosw, sinw, rhoi, rhos, rhow, gravit, puny
         integer(c_int32_t) :: coupled_tilt, use_ocnslope, hemisphere_turning, wind_from_strax
         integer(c_int32_t) :: kstrength, krdg_partic, krdg_redist, ncat
         real(c_double) :: mu_rdg
         integer(c_int32_t) :: math_mode, pin_host, use_graph, tile_threads, tile_rows, kernel_variant
         integer(c_int32_t) :: state_residency, exchange_mode
      end type

      type, bind(C) :: evp_b200_static_fields
         type(c_ptr) :: dxt, dyt, dxhy, dyhx, cxp, cyp, cxm, cym
         type(c_ptr) :: tarea, tarear, tinyarea, uarea, uarear, fcor
         type(c_ptr) :: tmask, umask
         type(c_ptr) :: HTE, HTN
      end type

      type, bind(C) :: evp_b200_inputs
         type(c_ptr) :: aice, vice, vsno, strairxT, strairyT, uocn, vocn, ss_tltx, ss_tlty
         type(c_ptr) :: aice0, aicen, vicen
      end type

      type, bind(C) :: evp_b200_state
         type(c_ptr) :: uvel, vvel
         type(c_ptr) :: stressp_1, stressp_2, stressp_3, stressp_4
         type(c_ptr) :: stressm_1, stressm_2, stressm_3, stressm_4
         type(c_ptr) :: stress12_1, stress12_2, stress12_3, stress12_4
         type(c_ptr) :: iceumask
      end type

      type, bind(C) :: evp_b200_outputs
         type(c_ptr) :: strairx, strairy, strtltx, strtlty, strintx, strinty
         type(c_ptr) :: strocnx, strocny, strocnxT, strocnyT, fm, prs_sig
         type(c_ptr) :: divu, shear, rdg_conv, rdg_shear, strength, sicemass, sig1, sig2
      end type

      interface
         integer(c_int) function evp_b200_init(dims, params, grid, handle) bind(C, name='evp_b200_init')
            import :: c_int, c_ptr, evp_b200_dims, evp_b200_params, evp_b200_static_fields
            type(evp_b200_dims), intent(in) :: dims
            type(evp_b200_params), intent(in) :: params
            type(evp_b200_static_fields), intent(in) :: grid
            type(c_ptr), intent(out) :: handle
         end function
         integer(c_int) function evp_b200_prep(handle, inp, st, icetmask_out) bind(C, name='evp_b200_prep')
            import :: c_int, c_ptr, evp_b200_inputs, evp_b200_state
            type(c_ptr), value :: handle
            type(evp_b200_inputs), intent(in) :: inp
            type(evp_b200_state), intent(in) :: st
            type(c_ptr), value :: icetmask_out
         end function
         integer(c_int) function evp_b200_run(handle, strength, st, outp) bind(C, name='evp_b200_run')
            import :: c_int, c_ptr, evp_b200_state, evp_b200_outputs
            type(c_ptr), value :: handle
            type(c_ptr), value :: strength
            type(evp_b200_state), intent(in) :: st
            type(evp_b200_outputs), intent(in) :: outp
         end function
         integer(c_int) function evp_b200_principal_stress_n(handle, n, sp1, sm1, s12, prs, s1, s2) &
               bind(C, name='evp_b200_principal_stress_n')
            import :: c_int, c_int64_t, c_ptr
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n
            type(c_ptr), value :: sp1, sm1, s12, prs, s1, s2
         end function
         integer(c_int) function evp_b200_download_state(handle, st) bind(C, name='evp_b200_download_state')
            import :: c_int, c_ptr, evp_b200_state
            type(c_ptr), value :: handle
            type(evp_b200_state), intent(in) :: st
         end function
         integer(c_int) function evp_b200_invalidate_device_state(handle) &
               bind(C, name='evp_b200_invalidate_device_state')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
         end function
         integer(c_int) function evp_b200_download_velocity(handle, uvel, vvel) &
               bind(C, name='evp_b200_download_velocity')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, uvel, vvel
         end function
         integer(c_int) function evp_b200_diagnostics(handle, out4) bind(C, name='evp_b200_diagnostics')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: out4(4)
         end function
         integer(c_int) function evp_b200_diagnostics_energy(handle, out8) bind(C, name='evp_b200_diagnostics_energy')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: out8(8)
         end function
         integer(c_int) function evp_b200_comm_unique_id(id) bind(C, name='evp_b200_comm_unique_id')
            import :: c_int, c_int8_t
            integer(c_int8_t), intent(out) :: id(128)
         end function
         integer(c_int) function evp_b200_comm_init(handle, id) bind(C, name='evp_b200_comm_init')
            import :: c_int, c_ptr, c_int8_t
            type(c_ptr), value :: handle
            integer(c_int8_t), intent(in) :: id(128)
         end function
         type(c_ptr) function evp_b200_last_error() bind(C, name='evp_b200_last_error')
            import :: c_ptr
         end function
         integer(c_int) function evp_b200_finalize(handle) bind(C, name='evp_b200_finalize')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
         end function
      end interface

      type(c_ptr), private :: b200_handle = c_null_ptr

      ! 0: the whole state travels every call (host arrays always current); 1: the stresses stay on the
      ! device; 2: uvel, vvel, iceumask too.  With 1 or 2 dumpfile / ice_write_hist must be preceded by
      ! `call evp_state_to_host` and restartfile followed by `call evp_state_from_host` (INTEGRATION.md).
      integer (kind=int_kind) :: evp_b200_residency = 0

      ! int32 0/1 copies of the Fortran logical masks (compiler-independent .true. pattern)
      integer(c_int32_t), allocatable, target, private :: &
         tmask_i4(:,:,:), umask_i4(:,:,:), iceumask_i4(:,:,:), icetmask_i4(:,:,:)
      integer(c_int32_t), allocatable, target, private :: &
         b_ilo(:), b_ihi(:), b_jlo(:), b_jhi(:), b_iglo(:), b_jglo(:)

      real (kind=dbl_kind), allocatable, target :: fcor_blk(:,:,:)   ! Coriolis parameter (1/s)

      contains

!=======================================================================
      subroutine b200_check(rc, what)
      use ice_exit, only: abort_ice
      integer(c_int), intent(in) :: rc
      character(*), intent(in) :: what
      if (rc /= 0) call abort_ice('ice_dyn_evp(b200): '//what//' failed, see libevp_b200 last_error')
      end subroutine b200_check

!=======================================================================
! Address of a module array of ice_state / ice_flux / ice_grid.  C_LOC needs an argument with the TARGET (or
! POINTER) attribute, and those modules -- which stay unchanged -- declare their arrays without it.  An
! assumed-size TARGET dummy is associated with a contiguous whole-array actual by reference (sequence association,
! no copy-in / copy-out for an explicit-shape module array), so C_LOC of the dummy is the address of the array
! itself; the arrays are SAVEd module variables, their storage does not move.  (Formally a pointer taken from a
! TARGET dummy whose actual has no TARGET attribute is not guaranteed beyond the call -- the alternative is to add
! `target` to the declarations in ice_state.F90, ice_flux.F90 and ice_grid.F90, see INTEGRATION.md.)
      type(c_ptr) function b200_addr(x)
      real (kind=dbl_kind), target, intent(in) :: x(*)
      b200_addr = c_loc(x)
      end function b200_addr

!=======================================================================
! init_evp: same duties as source/ice_dyn_evp.F90:441-526, plus the device setup.
      subroutine init_evp (dt)
      use ice_blocks
      use ice_domain
      use ice_state
      use ice_flux
      use ice_grid
      use ice_mechred, only: kstrength, krdg_partic, krdg_redist, mu_rdg
      real (kind=dbl_kind), intent(in) :: dt
      integer (kind=int_kind) :: i, j, iblk, nranks_y
      type (block) :: this_block
      type (evp_b200_dims) :: d
      type (evp_b200_params) :: p
      type (evp_b200_static_fields) :: g
      integer(c_int8_t) :: uid(128)

      if (my_task == master_task) then
         write(nu_diag,*) 'dt  = ',dt
         write(nu_diag,*) 'dte = ',dt/real(ndte,kind=dbl_kind)
         write(nu_diag,*) 'tdamp =', 0.36_dbl_kind*dt
      endif

      allocate(fcor_blk(nx_block,ny_block,max_blocks))
      allocate(tmask_i4(nx_block,ny_block,max_blocks), umask_i4(nx_block,ny_block,max_blocks), &
               iceumask_i4(nx_block,ny_block,max_blocks), icetmask_i4(nx_block,ny_block,max_blocks))
      tmask_i4 = 0; umask_i4 = 0; iceumask_i4 = 0; icetmask_i4 = 0

      do iblk = 1, nblocks
      do j = 1, ny_block
      do i = 1, nx_block
         uvel(i,j,iblk) = c0
         vvel(i,j,iblk) = c0
         divu (i,j,iblk) = c0
         shear(i,j,iblk) = c0
         rdg_conv (i,j,iblk) = c0
         rdg_shear(i,j,iblk) = c0
         fcor_blk(i,j,iblk) = c2*omega*sin(ULAT(i,j,iblk))
         stressp_1 (i,j,iblk) = c0; stressp_2 (i,j,iblk) = c0
         stressp_3 (i,j,iblk) = c0; stressp_4 (i,j,iblk) = c0
         stressm_1 (i,j,iblk) = c0; stressm_2 (i,j,iblk) = c0
         stressm_3 (i,j,iblk) = c0; stressm_4 (i,j,iblk) = c0
         stress12_1(i,j,iblk) = c0; stress12_2(i,j,iblk) = c0
         stress12_3(i,j,iblk) = c0; stress12_4(i,j,iblk) = c0
         iceumask(i,j,iblk) = .false.
         if (tmask(i,j,iblk)) tmask_i4(i,j,iblk) = 1
         if (umask(i,j,iblk)) umask_i4(i,j,iblk) = 1
      enddo
      enddo
      enddo

      ! block table: this_block%ilo.. and the global index of the first physical cell
      allocate(b_ilo(nblocks), b_ihi(nblocks), b_jlo(nblocks), b_jhi(nblocks), b_iglo(nblocks), b_jglo(nblocks))
      do iblk = 1, nblocks
         this_block = get_block(blocks_ice(iblk),iblk)
         b_ilo(iblk) = this_block%ilo;  b_ihi(iblk) = this_block%ihi
         b_jlo(iblk) = this_block%jlo;  b_jhi(iblk) = this_block%jhi
         b_iglo(iblk) = this_block%i_glob(this_block%ilo)
         b_jglo(iblk) = this_block%j_glob(this_block%jlo)
      enddo

      d%nx_block = nx_block; d%ny_block = ny_block; d%max_blocks = max_blocks; d%nblocks = nblocks
      d%nx_global = nx_global; d%ny_global = ny_global
      d%ew_boundary = b200_bnd(ew_boundary_type); d%ns_boundary = b200_bnd(ns_boundary_type)
      d%ilo = c_loc(b_ilo); d%ihi = c_loc(b_ihi); d%jlo = c_loc(b_jlo); d%jhi = c_loc(b_jhi)
      d%iglob_lo = c_loc(b_iglo); d%jglob_lo = c_loc(b_jglo)
      ! y-slab of this task: requires a 1 x N task layout (distribution_type='cartesian',
      ! processor_shape='slenderX1' with nprocs_x = 1), see INTEGRATION.md
      d%slab_jlo = minval(b_jglo)
      d%slab_jhi = maxval(b_jglo + (b_jhi - b_jlo))
      d%rank = my_task; d%nranks = get_num_procs(); d%device = -1

      p%dt = dt; p%ndte = ndte; p%evp_damping = merge(1, 0, evp_damping)
      p%dragio = dragio; p%cosw = cosw; p%sinw = sinw
      p%rhoi = rhoi; p%rhos = rhos; p%rhow = rhow; p%gravit = gravit; p%puny = puny
      p%coupled_tilt = 0; p%use_ocnslope = 0; p%hemisphere_turning = 0; p%wind_from_strax = 0
#ifdef coupled
      p%coupled_tilt = 1
#endif
#if defined(AusCOM) || defined(ACCICE)
      p%hemisphere_turning = 1
      p%use_ocnslope = merge(1, 0, use_ocnslope)
#endif
#ifdef ACCESS
      p%wind_from_strax = 1
#endif
      p%kstrength = kstrength; p%krdg_partic = krdg_partic; p%krdg_redist = krdg_redist
      p%ncat = ncat; p%mu_rdg = mu_rdg
      p%math_mode = 0; p%pin_host = 1; p%use_graph = 1
      p%tile_threads = 0; p%tile_rows = 0; p%kernel_variant = 0
      p%state_residency = evp_b200_residency; p%exchange_mode = 0

      g%dxt = b200_addr(dxt); g%dyt = b200_addr(dyt); g%dxhy = b200_addr(dxhy); g%dyhx = b200_addr(dyhx)
      g%cxp = b200_addr(cxp); g%cyp = b200_addr(cyp); g%cxm = b200_addr(cxm); g%cym = b200_addr(cym)
      g%tarea = b200_addr(tarea); g%tarear = b200_addr(tarear); g%tinyarea = b200_addr(tinyarea)
      g%uarea = b200_addr(uarea); g%uarear = b200_addr(uarear); g%fcor = c_loc(fcor_blk)
      g%tmask = c_loc(tmask_i4); g%umask = c_loc(umask_i4)
      g%HTE = b200_addr(HTE); g%HTN = b200_addr(HTN)      ! optional: 2-plane metric path (checked bit for bit at init)

      call b200_check(evp_b200_init(d, p, g, b200_handle), 'evp_b200_init')

      if (d%nranks > 1) then
         ! ncclUniqueId from task 0, broadcast with the model's own broadcast layer
         if (my_task == master_task) call b200_check(evp_b200_comm_unique_id(uid), 'comm_unique_id')
         call b200_bcast_bytes(uid)
         call b200_check(evp_b200_comm_init(b200_handle, uid), 'evp_b200_comm_init')
      endif
      end subroutine init_evp

!=======================================================================
      integer(c_int32_t) function b200_bnd(name)
      character(*), intent(in) :: name
      select case (trim(name))
      case ('open');    b200_bnd = 0
      case ('closed');  b200_bnd = 1
      case ('cyclic');  b200_bnd = 2
      case ('tripole'); b200_bnd = 3
      case ('tripoleT'); b200_bnd = 4
      case default;     b200_bnd = -1      ! rejected by evp_b200_init
      end select
      end function b200_bnd

      subroutine b200_bcast_bytes(buf)
      use ice_broadcast, only: broadcast_array
      integer(c_int8_t), intent(inout) :: buf(128)
      integer (kind=int_kind) :: tmp(128)
      tmp = int(buf, int_kind)
      call broadcast_array(tmp, master_task)
      buf = int(tmp, c_int8_t)
      end subroutine b200_bcast_bytes

!=======================================================================
! set_evp_parameters: kept for callers; the library derives the same scalars
! from (dt, ndte) at init (source/ice_dyn_evp.F90:535-577).
      subroutine set_evp_parameters (dt)
      real (kind=dbl_kind), intent(in) :: dt
      end subroutine set_evp_parameters

!=======================================================================
! evp: same interface and side effects as source/ice_dyn_evp.F90:119-432.
      subroutine evp (dt)
      use ice_blocks
      use ice_domain
      use ice_state
      use ice_flux
      use ice_grid
      use ice_timers
      use ice_mechred, only: ice_strength
      real (kind=dbl_kind), intent(in) :: dt
      integer (kind=int_kind) :: iblk, i, j, ilo, ihi, jlo, jhi, icellt
      integer (kind=int_kind), dimension (nx_block*ny_block) :: indxti, indxtj
      type (block) :: this_block
      type (evp_b200_inputs) :: inp
      type (evp_b200_state) :: st
      type (evp_b200_outputs) :: outp

      call ice_timer_start(timer_dynamics)

      where (iceumask) ; iceumask_i4 = 1 ; elsewhere ; iceumask_i4 = 0 ; end where

      inp%aice = b200_addr(aice); inp%vice = b200_addr(vice); inp%vsno = b200_addr(vsno)
#ifdef ACCESS
      inp%strairxT = b200_addr(strax); inp%strairyT = b200_addr(stray)
#else
      inp%strairxT = b200_addr(strairxT); inp%strairyT = b200_addr(strairyT)
#endif
      inp%uocn = b200_addr(uocn); inp%vocn = b200_addr(vocn)
      inp%ss_tltx = b200_addr(ss_tltx); inp%ss_tlty = b200_addr(ss_tlty)
      inp%aice0 = c_null_ptr; inp%aicen = c_null_ptr; inp%vicen = c_null_ptr

      call b200_state_ptrs(st)

      ! phase 1 on the device: evp_prep1, HALO icetmask, to_ugrid, t2ugrid_vector, evp_prep2
      call b200_check(evp_b200_prep(b200_handle, inp, st, c_loc(icetmask_i4)), 'evp_b200_prep')

      ! ice_strength stays on the host: same T-cell list, same order as :850-859
      do iblk = 1, nblocks
         this_block = get_block(blocks_ice(iblk),iblk)
         ilo = this_block%ilo; ihi = this_block%ihi
         jlo = this_block%jlo; jhi = this_block%jhi
         icellt = 0
         do j = jlo, jhi+1
         do i = ilo, ihi+1
            if (icetmask_i4(i,j,iblk) == 1) then
               icellt = icellt + 1
               indxti(icellt) = i
               indxtj(icellt) = j
            endif
         enddo
         enddo
         call ice_strength (nx_block, ny_block, ilo, ihi, jlo, jhi, icellt, indxti, indxtj, &
                            aice(:,:,iblk), vice(:,:,iblk), aice0(:,:,iblk), &
                            aicen(:,:,:,iblk), vicen(:,:,:,iblk), strength(:,:,iblk))
      enddo

      outp%strairx = b200_addr(strairx); outp%strairy = b200_addr(strairy)
      outp%strtltx = b200_addr(strtltx); outp%strtlty = b200_addr(strtlty)
      outp%strintx = b200_addr(strintx); outp%strinty = b200_addr(strinty)
      outp%strocnx = b200_addr(strocnx); outp%strocny = b200_addr(strocny)
      outp%strocnxT = b200_addr(strocnxT); outp%strocnyT = b200_addr(strocnyT)
      outp%fm = b200_addr(fm); outp%prs_sig = b200_addr(prs_sig)
      outp%divu = b200_addr(divu); outp%shear = b200_addr(shear)
      outp%rdg_conv = b200_addr(rdg_conv); outp%rdg_shear = b200_addr(rdg_shear)
      outp%strength = b200_addr(strength)      ! comes back halo-updated, as after :337-338
      outp%sig1 = b200_addr(sig1); outp%sig2 = b200_addr(sig2)   ! fused principal-stress epilogue (ice_history :1939)
#ifdef AusCOM
      outp%sicemass = b200_addr(sicemass)
#else
      outp%sicemass = c_null_ptr
#endif

      ! phase 2 on the device: HALO strength,u,v ; ndte x (stress, stepu, HALO) ; evp_finish ; u2tgrid_vector
      call b200_check(evp_b200_run(b200_handle, b200_addr(strength), st, outp), 'evp_b200_run')

      if (evp_b200_residency /= 2) iceumask = (iceumask_i4 == 1)

      call ice_timer_stop(timer_dynamics)
      end subroutine evp

!=======================================================================
! state_residency = 1 / 2: bring the device-resident state into the module arrays (call before dumpfile,
! source/ice_restart.F90:197-246, and before ice_write_hist) ...
      subroutine evp_state_to_host
      use ice_state
      use ice_flux
      type (evp_b200_state) :: st
      if (evp_b200_residency == 0) return
      call b200_state_ptrs(st)
      call b200_check(evp_b200_download_state(b200_handle, st), 'evp_b200_download_state')
      iceumask = (iceumask_i4 == 1)
      end subroutine evp_state_to_host

! ... and tell the library that the module arrays were changed by the host (after restartfile,
! source/ice_restart.F90:427-487): the next evp call uploads them again
      subroutine evp_state_from_host
      if (evp_b200_residency == 0) return
      call b200_check(evp_b200_invalidate_device_state(b200_handle), 'evp_b200_invalidate_device_state')
      end subroutine evp_state_from_host

! state_residency = 2: only the velocities, for the transport scheme right after evp
! (source/ice_step_mod.F90:575-585)
      subroutine evp_velocity_to_host
      use ice_state, only: uvel, vvel
      if (evp_b200_residency /= 2) return
      call b200_check(evp_b200_download_velocity(b200_handle, b200_addr(uvel), b200_addr(vvel)), 'evp_b200_download_velocity')
      end subroutine evp_velocity_to_host

      subroutine b200_state_ptrs(st)
      use ice_state
      use ice_flux
      type (evp_b200_state), intent(out) :: st
      st%uvel = b200_addr(uvel); st%vvel = b200_addr(vvel)
      st%stressp_1 = b200_addr(stressp_1); st%stressp_2 = b200_addr(stressp_2)
      st%stressp_3 = b200_addr(stressp_3); st%stressp_4 = b200_addr(stressp_4)
      st%stressm_1 = b200_addr(stressm_1); st%stressm_2 = b200_addr(stressm_2)
      st%stressm_3 = b200_addr(stressm_3); st%stressm_4 = b200_addr(stressm_4)
      st%stress12_1 = b200_addr(stress12_1); st%stress12_2 = b200_addr(stress12_2)
      st%stress12_3 = b200_addr(stress12_3); st%stress12_4 = b200_addr(stress12_4)
      st%iceumask = c_loc(iceumask_i4)
      end subroutine b200_state_ptrs

!=======================================================================
! principal_stress: same argument list as source/ice_dyn_evp.F90:1558-1561; computed by the library
! (evp_b200_principal_stress_n on one (nx_block, ny_block) slice, as ice_history calls it block by block)
      subroutine principal_stress(nx_block, ny_block, stressp_1, stressm_1, stress12_1, prs_sig, sig1, sig2)
      integer (kind=int_kind), intent(in) :: nx_block, ny_block
      real (kind=dbl_kind), dimension (nx_block,ny_block), intent(in), target :: &
         stressp_1, stressm_1, stress12_1, prs_sig
      real (kind=dbl_kind), dimension (nx_block,ny_block), intent(out), target :: sig1, sig2
      call b200_check(evp_b200_principal_stress_n(b200_handle, int(nx_block*ny_block, c_int64_t), &
                         c_loc(stressp_1), c_loc(stressm_1), c_loc(stress12_1), c_loc(prs_sig), &
                         c_loc(sig1), c_loc(sig2)), 'evp_b200_principal_stress_n')
      end subroutine principal_stress

      end module ice_dyn_evp
