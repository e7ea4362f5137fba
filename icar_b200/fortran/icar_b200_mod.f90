!> ----------------------------------------------------------------------------
!!  iso_c_binding shim over libicar_b200 (include/icar_b200.h).
!!
!!  Drop-in for ICAR's advection and microphysics plugins: it exports the same
!!  three procedures per family the reference uses
!!      adv_init / adv_var_request / advect      (src/physics/advection_driver.f90)
!!      mp_init  / mp_var_request  / mp          (src/physics/mp_driver.f90)
!!  so that src/main/init.f90 and src/main/time_step.f90 only have to `use` this
!!  module instead of `advection` / `microphysics`.
!!
!!  NOTE: no Fortran compiler exists in the build image of this repository, so this
!!  file is delivered as source that has NOT been compiled here; it is kept
!!  deliberately thin (interfaces + pointer plumbing).  The C ABI it binds is
!!  exercised from Python/ctypes with the same column-major (i,k,j) buffers.
!!
!!  Resident mode: fields are uploaded once (static metric terms in adv_init, state
!!  at the first step), stay on the GPU across steps and are downloaded by
!!  b200_sync_to_host before output / restart (driver.f90:169-195).
!! ----------------------------------------------------------------------------
module icar_b200_mod
    use, intrinsic :: iso_c_binding
    use data_structures
    use icar_constants
    use options_interface, only : options_t
    use domain_interface,  only : domain_t
    ! variable requests are unchanged: reuse the reference's own
    use adv_upwind,        only : upwind_var_request
    use module_mp_simple,  only : mp_simple_var_request
    use microphysics,      only : reference_mp_var_request => mp_var_request   ! Thompson's request list (mp_driver.f90:118-145)
    implicit none
    private
    public :: adv_init, adv_var_request, advect
    public :: mp_init, mp_var_request, mp
    public :: b200_halo_send, b200_halo_retrieve, b200_sync_to_host, b200_finalize, b200_physics_steps, b200_step
    public :: b200_diagnostic_update, b200_diagnostic_output, b200_update_forcing, b200_apply_forcing, b200_enforce_limits, b200_upload_winds
    public :: b200_linear_perturb, b200_write_restart, b200_read_restart

    ! enum icar_var
    integer(c_int), parameter :: VAR_U=0, VAR_V=1, VAR_W=2, VAR_PRESSURE=3, VAR_EXNER=4, VAR_DENSITY=5, &
        VAR_TH=6, VAR_QV=7, VAR_QC=8, VAR_QR=9, VAR_QS=10, VAR_JACOBIAN=11, VAR_JACOBIAN_U=12,       &
        VAR_JACOBIAN_V=13, VAR_JACOBIAN_W=14, VAR_ADVECTION_DZ=15, VAR_DZ_MASS=16, VAR_PRECIP=17,     &
        VAR_SNOWFALL=18, VAR_QI=19, VAR_NI=20, VAR_NR=21, VAR_QG=22, VAR_GRAUPEL=23

    type, bind(c) :: icar_grid
        integer(c_int) :: ims, ime, kms, kme, jms, jme
        integer(c_int) :: its, ite, kts, kte, jts, jte
        integer(c_int) :: ids, ide, kds, kde, jds, jde
        integer(c_int) :: halo_size
    end type

    type, bind(c) :: icar_options
        integer(c_int) :: advection, microphysics, mpdata_order, flux_corrected_transport
        integer(c_int) :: advect_density, update_interval, top_mp_level
        real(c_float)  :: dx
        integer(c_int) :: exact_arithmetic
    end type

    !> icar_mp_options == mp_options_type (opt_types.f90:31-45) up to the two logicals, which are C ints
    type, bind(c) :: icar_mp_options
        real(c_float)  :: Nt_c, TNO, am_s, rho_g
        real(c_float)  :: av_s, bv_s, fv_s, av_i
        real(c_float)  :: av_g, bv_g
        real(c_float)  :: Ef_si, Ef_rs, Ef_rg, Ef_ri
        real(c_float)  :: C_cubes, C_sqrd
        real(c_float)  :: mu_r
        real(c_float)  :: t_adjust
        integer(c_int) :: Ef_rw_l, Ef_sw_l
    end type

    interface
        integer(c_int) function icar_thompson_init(ctx, opt) bind(c)
            import :: c_int, c_ptr, icar_mp_options
            type(c_ptr), value :: ctx
            type(icar_mp_options), intent(in) :: opt
        end function
        subroutine icar_default_options(opt) bind(c)
            import :: icar_options
            type(icar_options), intent(out) :: opt
        end subroutine
        integer(c_int) function icar_create(device, grid, opt, ctx) bind(c)
            import :: c_int, c_ptr, icar_grid, icar_options
            integer(c_int), value :: device
            type(icar_grid), intent(in) :: grid
            type(icar_options), intent(in) :: opt
            type(c_ptr), intent(out) :: ctx
        end function
        integer(c_int) function icar_destroy(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        type(c_ptr) function icar_last_error(ctx) bind(c)
            import :: c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_upload(ctx, var, host) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: var
            type(c_ptr), value :: host
        end function
        integer(c_int) function icar_download(ctx, var, host) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: var
            type(c_ptr), value :: host
        end function
        integer(c_int) function icar_set_model_time(ctx, seconds) bind(c)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: seconds
        end function
        integer(c_int) function icar_advect(ctx, dt) bind(c)
            import :: c_int, c_ptr, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: dt
        end function
        integer(c_int) function icar_mp(ctx, dt_in, halo, subset) bind(c)
            import :: c_int, c_ptr, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: dt_in
            integer(c_int), value :: halo, subset
        end function
        integer(c_int) function icar_physics_steps(ctx, dt, nsteps, use_graph) bind(c)
            import :: c_int, c_ptr, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: dt
            integer(c_int), value :: nsteps, use_graph
        end function
        integer(c_int) function icar_linear_winds_init(ctx, global_terrain, nx_global, ny_global, buffer) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            type(c_ptr), value :: global_terrain
            integer(c_int), value :: nx_global, ny_global, buffer
        end function
        integer(c_int) function icar_linear_perturb(ctx, U, V, Nsq, global_z_interface, global_dz_interface,   &
                                                    layer_height, dz_levels, minimum_layer_size,              &
                                                    linear_contribution, linear_update_fraction) bind(c)
            import :: c_int, c_ptr, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: U, V, Nsq
            type(c_ptr), value :: global_z_interface, global_dz_interface, layer_height, dz_levels
            real(c_float), value :: minimum_layer_size, linear_contribution, linear_update_fraction
        end function
        integer(c_int) function icar_balance_uvw(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_write_restart(ctx, path) bind(c)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: path(*)
        end function
        integer(c_int) function icar_read_restart(ctx, path) bind(c)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: path(*)
        end function
        integer(c_int) function icar_device_count(n) bind(c)
            import :: c_int
            integer(c_int), intent(out) :: n
        end function
        integer(c_int) function icar_diagnostic_update(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_diagnostic_output(ctx, dzdx, dzdy, temperature, pressure_i, temperature_i, psfc, &
                                                       u_mass, v_mass, w_real) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, dzdx, dzdy, temperature, pressure_i, temperature_i, psfc, u_mass, v_mass, w_real
        end function
        integer(c_int) function icar_enforce_limits(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_step(ctx, end_time, dz_levels, cfl_reduction_factor, cfl_strictness, steps_taken) bind(c)
            import :: c_int, c_ptr, c_double, c_float
            type(c_ptr), value :: ctx
            real(c_double), value :: end_time
            type(c_ptr), value :: dz_levels
            real(c_float), value :: cfl_reduction_factor
            integer(c_int), value :: cfl_strictness
            type(c_ptr), value :: steps_taken
        end function
        integer(c_int) function icar_upload_tendency(ctx, var, host) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: var
            type(c_ptr), value :: host
        end function
        integer(c_int) function icar_apply_forcing(ctx, dt_seconds) bind(c)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: dt_seconds
        end function
        integer(c_int) function icar_halo_send(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_halo_retrieve(ctx) bind(c)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function icar_comm_unique_id(id) bind(c)
            import :: c_int, c_char
            character(kind=c_char) :: id(128)
        end function
        integer(c_int) function icar_comm_init(ctx, image, nimages, id) bind(c)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: image, nimages
            character(kind=c_char), intent(in) :: id(128)
        end function
    end interface

    type(c_ptr), save :: ctx = c_null_ptr
    logical, save :: state_on_device = .false.
    type(icar_mp_options), save :: mpo
    logical, save :: have_mpo = .false., thompson = .false.

contains

    subroutine check(rc, where)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        if (rc /= 0) then
            write(*,*) "icar_b200 error ", rc, " in ", where
            error stop "icar_b200"          ! the reference's own error convention (time_step.f90:322-328)
        endif
    end subroutine

    subroutine up3(var, a)
        integer(c_int), intent(in) :: var
        real, intent(in), target, contiguous :: a(:,:,:)
        call check(icar_upload(ctx, var, c_loc(a)), "icar_upload")
    end subroutine

    subroutine down3(var, a)
        integer(c_int), intent(in) :: var
        real, intent(inout), target, contiguous :: a(:,:,:)
        call check(icar_download(ctx, var, c_loc(a)), "icar_download")
    end subroutine

    !> advection_driver.f90:39-49
    subroutine adv_var_request(options)
        type(options_t), intent(inout) :: options
        call upwind_var_request(options)
    end subroutine

    !> advection_driver.f90:24-37: creates the GPU context for this image's tile and uploads the
    !! static metric terms (jacobians, advection_dz, dz_mass) once.
    subroutine adv_init(domain, options)
        type(domain_t),  intent(in) :: domain
        type(options_t), intent(in) :: options
        type(icar_grid)    :: g
        type(icar_options) :: o
        character(kind=c_char) :: id(128)
        integer(c_int) :: ndev

        g%ims = domain%grid%ims; g%ime = domain%grid%ime; g%kms = domain%grid%kms; g%kme = domain%grid%kme
        g%jms = domain%grid%jms; g%jme = domain%grid%jme
        g%its = domain%grid%its; g%ite = domain%grid%ite; g%kts = domain%grid%kts; g%kte = domain%grid%kte
        g%jts = domain%grid%jts; g%jte = domain%grid%jte
        g%ids = domain%grid%ids; g%ide = domain%grid%ide; g%kds = domain%grid%kds; g%kde = domain%grid%kde
        g%jds = domain%grid%jds; g%jde = domain%grid%jde
        g%halo_size = domain%grid%halo_size

        call icar_default_options(o)
        o%advection    = options%physics%advection
        o%microphysics = options%physics%microphysics
        o%mpdata_order = options%adv_options%mpdata_order
        o%flux_corrected_transport = merge(1, 0, options%adv_options%flux_corrected_transport)
        o%advect_density  = merge(1, 0, options%parameters%advect_density)
        o%update_interval = options%mp_options%update_interval
        o%top_mp_level    = options%mp_options%top_mp_level
        o%dx = domain%dx

        ! one image = one GPU of the node (images are placed on the node's devices round robin)
        call check(icar_device_count(ndev), "icar_device_count")
        call check(icar_create(mod(this_image()-1, max(ndev, 1_c_int)), g, o, ctx), "icar_create")

        if (num_images() > 1) then          ! x-slab decompositions only (yimages == 1)
            if (this_image() == 1) call check(icar_comm_unique_id(id), "icar_comm_unique_id")
            call co_broadcast(id, 1)
            call check(icar_comm_init(ctx, this_image(), num_images(), id), "icar_comm_init")
        endif

        call up3(VAR_JACOBIAN,     domain%jacobian)
        call up3(VAR_JACOBIAN_U,   domain%jacobian_u)
        call up3(VAR_JACOBIAN_V,   domain%jacobian_v)
        call up3(VAR_JACOBIAN_W,   domain%jacobian_w)
        call up3(VAR_ADVECTION_DZ, domain%advection_dz)
        call up3(VAR_DZ_MASS,      domain%dz_mass%data_3d)
        thompson = (options%physics%microphysics == kMP_THOMPSON)
        if (thompson .and. have_mpo) call check(icar_thompson_init(ctx, mpo), "icar_thompson_init")
    end subroutine

    !> mp_driver.f90:200-237 (only the schemes this library builds)
    subroutine mp_var_request(options)
        type(options_t), intent(inout) :: options
        if (options%physics%microphysics == kMP_SB04) call mp_simple_var_request(options)
        if (options%physics%microphysics == kMP_THOMPSON) call reference_mp_var_request(options)
    end subroutine

    !> mp_driver.f90:74-198.  mp_simple needs nothing; for Thompson this is thompson_init(options%mp_options)
    !! (mp_thompson.f90:342): the lookup tables are built on the GPU (no *_mpt.dat cache files).
    !! icar_create already ran it with the namelist defaults; this call passes the user's mp_parameters.
    !! init_physics calls mp_init BEFORE adv_init (init.f90:136,146), so when the context does not exist
    !! yet the options are kept and applied at the end of adv_init.
    subroutine mp_init(options)
        type(options_t), intent(in) :: options
        if (options%physics%microphysics /= kMP_THOMPSON) return
        mpo%Nt_c = options%mp_options%Nt_c;   mpo%TNO = options%mp_options%TNO
        mpo%am_s = options%mp_options%am_s;   mpo%rho_g = options%mp_options%rho_g
        mpo%av_s = options%mp_options%av_s;   mpo%bv_s = options%mp_options%bv_s
        mpo%fv_s = options%mp_options%fv_s;   mpo%av_i = options%mp_options%av_i
        mpo%av_g = options%mp_options%av_g;   mpo%bv_g = options%mp_options%bv_g
        mpo%Ef_si = options%mp_options%Ef_si; mpo%Ef_rs = options%mp_options%Ef_rs
        mpo%Ef_rg = options%mp_options%Ef_rg; mpo%Ef_ri = options%mp_options%Ef_ri
        mpo%C_cubes = options%mp_options%C_cubes; mpo%C_sqrd = options%mp_options%C_sqrd
        mpo%mu_r = options%mp_options%mu_r;   mpo%t_adjust = options%mp_options%t_adjust
        mpo%Ef_rw_l = merge(1, 0, options%mp_options%Ef_rw_l)
        mpo%Ef_sw_l = merge(1, 0, options%mp_options%Ef_sw_l)
        have_mpo = .true.
        if (c_associated(ctx)) call check(icar_thompson_init(ctx, mpo), "icar_thompson_init")
    end subroutine

    !> Resident mode.  The state (winds, pressure, scalars) is uploaded once, at the first mp() call; after that the
    !! whole loop body of step() that touches it runs on the device, and the call sites of the reference that would
    !! otherwise change the HOST copies are rerouted (INTEGRATION.md section 1):
    !!     diagnostic_update(domain, options)  ->  b200_diagnostic_update()      (time_step.f90:474)
    !!     domain%apply_forcing(dt)            ->  b200_apply_forcing(dt)        (time_step.f90:537)
    !!     domain%enforce_limits()             ->  b200_enforce_limits()         (time_step.f90:541)
    !!     after every forcing update          ->  b200_update_forcing(domain)   (driver.f90:134-151: the new dqdt_3d)
    !!     after update_winds on the host      ->  b200_upload_winds(domain)
    !! Without these the lateral boundary tendencies of th / qv / u / v / p never reach the device copies.
    subroutine sync_to_device(domain)
        type(domain_t), intent(inout) :: domain
        if (.not. state_on_device) then
            call b200_upload_winds(domain)
            call up3(VAR_PRESSURE, domain%pressure%data_3d)
            call up3(VAR_TH, domain%potential_temperature%data_3d)
            call up3(VAR_QV, domain%water_vapor%data_3d)
            call up3(VAR_QC, domain%cloud_water_mass%data_3d)
            call up3(VAR_QR, domain%rain_mass%data_3d)
            call up3(VAR_QS, domain%snow_mass%data_3d)
            if (thompson) then
                call up3(VAR_QI, domain%cloud_ice_mass%data_3d);  call up3(VAR_NI, domain%cloud_ice_number%data_3d)
                call up3(VAR_NR, domain%rain_number%data_3d);     call up3(VAR_QG, domain%graupel_mass%data_3d)
            endif
            call check(icar_diagnostic_update(ctx), "icar_diagnostic_update")     ! exner, density from p and th
            state_on_device = .true.
        endif
        call check(icar_set_model_time(ctx, domain%model_time%seconds()), "icar_set_model_time")
    end subroutine

    !> u, v, w as the host has them (after update_winds / balance_uvw on the host)
    subroutine b200_upload_winds(domain)
        type(domain_t), intent(inout) :: domain
        call up3(VAR_U, domain%u%data_3d);  call up3(VAR_V, domain%v%data_3d);  call up3(VAR_W, domain%w%data_3d)
    end subroutine

    !> diagnostic_update(domain, options), time_step.f90:49-198, for what mp / advect read: exner and density
    subroutine b200_diagnostic_update()
        if (state_on_device) call check(icar_diagnostic_update(ctx), "icar_diagnostic_update")
    end subroutine

    !> the output-only fields of diagnostic_update (time_step.f90:88-108, 165-194) from the device state, into the host
    !! arrays of `domain` that are associated: call before output instead of syncing u, v, w, p, th to the host
    subroutine b200_diagnostic_output(domain)
        type(domain_t), intent(inout), target :: domain
        call check(icar_diagnostic_output(ctx, c_loc(domain%dzdx), c_loc(domain%dzdy),                              &
                   loc3(domain%temperature%data_3d), loc3(domain%pressure_interface%data_3d),                        &
                   loc3(domain%temperature_interface%data_3d), loc2(domain%surface_pressure%data_2d),               &
                   loc3(domain%u_mass%data_3d), loc3(domain%v_mass%data_3d), loc3(domain%w_real%data_3d)),           &
                   "icar_diagnostic_output")
    contains
        type(c_ptr) function loc3(a)
            real, intent(in), pointer :: a(:,:,:)
            loc3 = c_null_ptr
            if (associated(a)) loc3 = c_loc(a)
        end function
        type(c_ptr) function loc2(a)
            real, intent(in), pointer :: a(:,:)
            loc2 = c_null_ptr
            if (associated(a)) loc2 = c_loc(a)
        end function
    end subroutine

    !> after the boundary object has new tendencies (update_delta_fields, domain_obj.f90:2318-2381): var%dqdt_3d of the
    !! forced variables goes to the device
    subroutine b200_update_forcing(domain)
        type(domain_t), intent(inout), target :: domain
        call tend(VAR_U, domain%u%dqdt_3d);   call tend(VAR_V, domain%v%dqdt_3d)
        call tend(VAR_PRESSURE, domain%pressure%dqdt_3d)
        call tend(VAR_TH, domain%potential_temperature%dqdt_3d)
        call tend(VAR_QV, domain%water_vapor%dqdt_3d)
    contains
        subroutine tend(var, a)
            integer(c_int), intent(in) :: var
            real, intent(in), pointer :: a(:,:,:)
            if (associated(a)) call check(icar_upload_tendency(ctx, var, c_loc(a)), "icar_upload_tendency")
        end subroutine
    end subroutine

    !> domain%apply_forcing(dt), domain_obj.f90:2383-2448, on the device copies
    subroutine b200_apply_forcing(dt_seconds)
        double precision, intent(in) :: dt_seconds
        call check(icar_apply_forcing(ctx, real(dt_seconds, c_double)), "icar_apply_forcing")
    end subroutine

    !> domain%enforce_limits(), domain_obj.f90:2228-2243, on the device copies
    subroutine b200_enforce_limits()
        call check(icar_enforce_limits(ctx), "icar_enforce_limits")
    end subroutine

    !> mp(domain, options, dt_in, halo, subset) -- mp_driver.f90:673
    subroutine mp(domain, options, dt_in, halo, subset)
        type(domain_t),  intent(inout) :: domain
        type(options_t), intent(in)    :: options
        real,            intent(in)    :: dt_in
        integer,         intent(in), optional :: halo, subset
        integer(c_int) :: h, s
        h = -1; s = -1
        if (present(halo))   h = halo
        if (present(subset)) s = subset
        if (present(halo) .or. (.not.present(subset))) call sync_to_device(domain)   ! first mp call of the step
        call check(icar_mp(ctx, dt_in, h, s), "icar_mp")
    end subroutine

    !> domain%halo_send / halo_retrieve (domain_obj.f90:104-144) for the device-resident scalars
    subroutine b200_halo_send()
        call check(icar_halo_send(ctx), "icar_halo_send")
    end subroutine
    subroutine b200_halo_retrieve()
        call check(icar_halo_retrieve(ctx), "icar_halo_retrieve")
    end subroutine

    !> n passes of mp(halo) -> halo_send -> mp(subset) -> halo_retrieve -> advect at a fixed dt, replayed from
    !> a CUDA graph: what the `do while` loop of step() (time_step.f90:464-549) does between two dt changes
    !> when microphysics and advection are the only physics switched on
    subroutine b200_physics_steps(dt, n)
        real,    intent(in) :: dt
        integer, intent(in) :: n
        call check(icar_physics_steps(ctx, dt, int(n, c_int), 1_c_int), "icar_physics_steps")
    end subroutine

    !> step(domain, end_time, options) -- time_step.f90:438-549 -- with the whole loop on the device: update_dt,
    !> diagnostic_update, mp / halo exchange / advect, apply_forcing and enforce_limits of every internal time step.
    !> end_time_seconds is on the clock the context was given (icar_set_model_time); the host copies are stale
    !> afterwards until b200_sync_to_host.
    subroutine b200_step(end_time_seconds, options)
        double precision, intent(in) :: end_time_seconds
        type(options_t),  intent(in) :: options
        real(c_float), allocatable, target :: dz(:)
        dz = options%parameters%dz_levels(1:options%parameters%nz)
        call check(icar_step(ctx, real(end_time_seconds, c_double), c_loc(dz),                        &
                             real(options%parameters%cfl_reduction_factor, c_float),                  &
                             int(options%parameters%cfl_strictness, c_int), c_null_ptr), "icar_step")
    end subroutine

    !> advect(domain, options, dt) -- advection_driver.f90:51
    subroutine advect(domain, options, dt)
        type(domain_t),  intent(inout) :: domain
        type(options_t), intent(in)    :: options
        real,            intent(in)    :: dt
        call check(icar_advect(ctx, dt), "icar_advect")
    end subroutine

    !> before output / restart / any host physics that reads the scalars
    subroutine b200_sync_to_host(domain)
        type(domain_t), intent(inout) :: domain
        call down3(VAR_TH, domain%potential_temperature%data_3d)
        call down3(VAR_QV, domain%water_vapor%data_3d)
        call down3(VAR_QC, domain%cloud_water_mass%data_3d)
        call down3(VAR_QR, domain%rain_mass%data_3d)
        call down3(VAR_QS, domain%snow_mass%data_3d)
        call check(icar_download(ctx, VAR_PRECIP,   c_loc(domain%accumulated_precipitation%data_2dd)), "icar_download")
        call check(icar_download(ctx, VAR_SNOWFALL, c_loc(domain%accumulated_snowfall%data_2dd)), "icar_download")
        if (thompson) then
            call down3(VAR_QI, domain%cloud_ice_mass%data_3d);  call down3(VAR_NI, domain%cloud_ice_number%data_3d)
            call down3(VAR_NR, domain%rain_number%data_3d);     call down3(VAR_QG, domain%graupel_mass%data_3d)
            call check(icar_download(ctx, VAR_GRAUPEL, c_loc(domain%graupel%data_2dd)), "icar_download")
        endif
    end subroutine

    !> linear_perturb(domain, options, vsmooth, reverse, useDensity, update) -- linear_winds.f90:1317-1366, as
    !> update_winds calls it (wind.f90:310-332), for the space_varying_dz case and a uniform background state
    !> (U, V, Nsq): the perturbation is computed on the GPU for that state instead of being interpolated from the
    !> look-up table, added to the resident u, v, and w is re-balanced.  The caller uploads domain%u, domain%v first.
    subroutine b200_linear_perturb(domain, options, U, V, Nsq)
        type(domain_t),  intent(inout), target :: domain      ! global_terrain etc. are allocatable components
        type(options_t), intent(in)    :: options
        real,            intent(in)    :: U, V, Nsq
        logical, save :: initialized = .false.
        if (.not. initialized) then                 ! setup_linwinds, linear_winds.f90:1180-1222
            call check(icar_linear_winds_init(ctx, c_loc(domain%global_terrain),                       &
                                              int(size(domain%global_terrain, 1), c_int),              &
                                              int(size(domain%global_terrain, 2), c_int),              &
                                              int(options%lt_options%buffer, c_int)), "icar_linear_winds_init")
            initialized = .true.
        endif
        call check(icar_linear_perturb(ctx, U, V, Nsq, c_loc(domain%global_z_interface),               &
                                       c_loc(domain%global_dz_interface), c_null_ptr, c_null_ptr,      &
                                       options%lt_options%minimum_layer_size,                          &
                                       options%lt_options%linear_contribution,                         &
                                       options%lt_options%linear_update_fraction), "icar_linear_perturb")
        call check(icar_balance_uvw(ctx), "icar_balance_uvw")     ! wind.f90:330-332
    end subroutine

    !> output_t%save_file / restart_model (output_obj.f90:41-80, restart.f90:11-90): one binary file per image
    subroutine b200_write_restart(filename)
        character(len=*), intent(in) :: filename
        call check(icar_write_restart(ctx, trim(filename)//c_null_char), "icar_write_restart")
    end subroutine

    subroutine b200_read_restart(filename)
        character(len=*), intent(in) :: filename
        call check(icar_read_restart(ctx, trim(filename)//c_null_char), "icar_read_restart")
        state_on_device = .true.
    end subroutine

    subroutine b200_finalize()
        if (c_associated(ctx)) call check(icar_destroy(ctx), "icar_destroy")
        ctx = c_null_ptr
    end subroutine
end module icar_b200_mod
