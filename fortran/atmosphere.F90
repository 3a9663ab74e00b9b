!> atmosphere_mod on libisca_b200.so -- drop-in replacement of src/atmos_spectral/driver/solo/atmosphere.F90 of the reference.
!!
!! Same public interface (atmosphere_init, atmosphere, atmosphere_end, atmosphere_domain), same namelists read under the same names
!! from input.nml (atmosphere_nml, spectral_dynamics_nml, hs_forcing_nml, spectral_init_cond_nml; constants through constants_mod),
!! same restart files (INPUT/atmosphere.res.nc, INPUT/spectral_dynamics.res.nc -> RESTART/...), same diag_manager fields of module
!! 'dynamics'.  src/atmos_solo/atmos_model.F90, FMS (mpp, fms_io, diag_manager, time_manager), field_table / diag_table and the isca
!! Python experiment scripts stay unchanged.  The model state lives on the GPU; this module keeps host mirrors only for restarts
!! and for the diagnostics that are due.
!!
!! Put this file and isca_b200_c.F90 in the executable's path_names instead of atmosphere.F90, drop atmos_spectral/model,
!! atmos_spectral/tools (except spec_mpp.F90, which owns the FMS domains) and shared/fft from the GPU executable, and add
!! `-L<repo>/isca_b200/lib -lisca_b200 -lcudart` to LDFLAGS of the mkmf template.
!!
!! This build image has no Fortran compiler, so this file has not been compiled here; tests/host/shim_driver.c performs the same call
!! sequence through the C ABI with arrays in Fortran order (tests/test_fortran_shim.py), and tools/gen_fortran_interface.py keeps the
!! bind(C) types in step with the headers.  Namelist values the library does not implement are rejected by isca_b200_create with the
!! reference's wording and reach error_mesg(..., FATAL) here.
module atmosphere_mod

#ifdef INTERNAL_FILE_NML
  use mpp_mod, only: input_nml_file
#else
  use fms_mod, only: open_namelist_file, close_file
#endif
  use iso_c_binding
  use isca_b200_c
  use fms_mod,            only: write_version_number, file_exist, field_size, stdlog, mpp_pe, mpp_root_pe, mpp_npes, error_mesg, FATAL, &
                                read_data, write_data, set_domain, nullify_domain
  use mpp_mod,            only: mpp_broadcast, mpp_gather
  use constants_mod,      only: grav, rdgas, kappa, radius, omega, pi
  use time_manager_mod,   only: time_type, get_time, operator(+)
  use interpolator_mod,   only: interpolate_type, interpolator_init, interpolator, CONSTANT
  use spec_mpp_mod,       only: spec_mpp_init, grid_domain, spectral_domain, get_grid_domain, get_spec_domain, atmosphere_domain
  use diag_manager_mod,   only: diag_axis_init, register_diag_field, register_static_field, send_data, need_data
  use field_manager_mod,  only: MODEL_ATMOS
  use tracer_manager_mod, only: get_number_tracers, get_tracer_index, NO_TRACER

  implicit none
  private
  public :: atmosphere_init, atmosphere, atmosphere_end, atmosphere_domain

  character(len=128), parameter :: version = 'isca_b200 atmosphere_mod shim'
  character(len=128), parameter :: tagname = 'libisca_b200'
  character(len=8),   parameter :: mod_name = 'dynamics'

  !------------------------------------------------------------------ atmosphere_nml (atmosphere.F90:84-86)
  logical :: idealized_moist_model = .false.
  namelist /atmosphere_nml/ idealized_moist_model

  !------------------------------------------------------------------ mixed_layer_nml do_sc_sst / sst_file (mixed_layer.F90:116,146): the
  ! prescribed SST is read on the host (interpolator_init in atmosphere_init, as mixed_layer_init :318-320 does) and handed to the library
  logical :: do_sc_sst = .false.
  character(len=256) :: sst_file = ''
  type(interpolate_type), save :: sst_interp
  real, allocatable :: sst_new(:,:)

  !------------------------------------------------------------------ spectral_dynamics_nml (spectral_dynamics.F90:152-224), same defaults
  logical :: do_mass_correction = .true., do_water_correction = .true., do_energy_correction = .true., &
             use_virtual_temperature = .false., use_implicit = .true., triang_trunc = .true., graceful_shutdown = .false., &
             make_symmetric = .false., json_logging = .false.
  integer :: damping_order = 2, damping_order_vor = -1, damping_order_div = -1, cutoff_wn = 15, lon_max = 128, lat_max = 64, &
             num_fourier = 42, num_spherical = 43, fourier_inc = 1, num_levels = 18, num_steps = 1
  integer, dimension(2) :: print_interval = (/1, 0/)
  character(len=64) :: vert_coord_option = 'even_sigma', damping_option = 'resolution_dependent', vert_advect_uv = 'second_centered', &
                       vert_advect_t = 'second_centered', vert_difference_option = 'simmons_and_burridge', &
                       initial_state_option = 'quiescent'
  real :: damping_coeff = 1.15740741e-4, damping_coeff_vor = -1., damping_coeff_div = -1., eddy_sponge_coeff = 0., &
          zmu_sponge_coeff = 0., zmv_sponge_coeff = 0., robert_coeff = .04, alpha_implicit = .5, longitude_origin = 0., &
          scale_heights = 4., surf_res = .1, p_press = .1, p_sigma = .3, exponent = 2.5, ocean_topog_smoothing = .93, &
          initial_sphum = 0.0, reference_sea_level_press = 101325., water_correction_limit = 0.0, raw_filter_coeff = 1.0
  real, dimension(2) :: valid_range_t = (/100., 500./)
  namelist /spectral_dynamics_nml/ use_virtual_temperature, damping_option, cutoff_wn, damping_order, damping_coeff, damping_order_vor, &
                                   damping_coeff_vor, damping_order_div, damping_coeff_div, do_mass_correction, do_water_correction, &
                                   do_energy_correction, vert_advect_uv, vert_advect_t, use_implicit, longitude_origin, robert_coeff, &
                                   alpha_implicit, vert_difference_option, reference_sea_level_press, lon_max, lat_max, num_levels, &
                                   num_fourier, num_spherical, fourier_inc, triang_trunc, vert_coord_option, scale_heights, surf_res, &
                                   p_press, p_sigma, exponent, ocean_topog_smoothing, initial_sphum, valid_range_t, eddy_sponge_coeff, &
                                   zmu_sponge_coeff, zmv_sponge_coeff, print_interval, num_steps, initial_state_option, &
                                   water_correction_limit, raw_filter_coeff, graceful_shutdown, json_logging, make_symmetric

  !------------------------------------------------------------------ hs_forcing_nml (hs_forcing.F90:74-122).  The group lists every variable
  ! of the reference's group so that any input.nml reads; this shim forwards the Held_Suarez option (the parameters in the first two
  ! lines), the EXOPLANET / top_down / local-heating options are served by the model of include/isca_b200_hs.h and end in FATAL here.
  logical :: no_forcing = .false., do_conserve_energy = .true., relax_to_specified_wind = .false.
  real :: t_zero = 315., t_strat = 200., delh = 60., delv = 10., eps = 0., sigma_b = 0.7, P00 = 1.e5, ka = -40., ks = -4., kf = -1., &
          trflux = 1.e-5, trsink = -4.
  real :: local_heating_srfamp = 0.0, local_heating_xwidth = 10., local_heating_ywidth = 10., local_heating_xcenter = 180., &
          local_heating_ycenter = 45., local_heating_vert_decay = 1.e4, p_trop = 1.e4, alpha = 2./7., peri_time = 0.25, smaxis = 1.5e6, &
          albedo = 0.3, lapse = 6.5, h_a = 2, tau_s = 5, orbital_period = 31557600., heat_capacity = 4.2e6, ml_depth = 1, &
          spinup_time = 10800.
  character(len=256) :: local_heating_option = '', equilibrium_t_option = 'Held_Suarez', local_heating_file = '', u_wind_file = 'u', &
                        v_wind_file = 'v', equilibrium_t_file = 'temp', stratosphere_t_option = 'extend_tp'
  namelist /hs_forcing_nml/ no_forcing, t_zero, t_strat, delh, delv, eps, sigma_b, ka, ks, kf, do_conserve_energy, trflux, trsink, &
                            local_heating_srfamp, local_heating_xwidth, local_heating_ywidth, local_heating_xcenter, &
                            local_heating_ycenter, local_heating_vert_decay, local_heating_option, local_heating_file, &
                            relax_to_specified_wind, u_wind_file, v_wind_file, equilibrium_t_option, equilibrium_t_file, p_trop, alpha, &
                            peri_time, smaxis, albedo, lapse, h_a, tau_s, orbital_period, heat_capacity, ml_depth, spinup_time, &
                            stratosphere_t_option, P00

  !------------------------------------------------------------------ spectral_init_cond_nml (spectral_init_cond.F90:68-74)
  real :: initial_temperature = 264.
  character(len=64) :: topography_option = 'flat'
  namelist /spectral_init_cond_nml/ initial_temperature, topography_option

  !------------------------------------------------------------------ module state
  type(c_ptr) :: h = c_null_ptr              ! dynamical-core handle (owned by hm when the moist model runs)
  type(c_ptr) :: hm = c_null_ptr             ! idealized moist model handle
  integer, parameter :: num_time_levels = 2
  integer :: is, ie, js, je, ms, me, ns, ne, num_tracers, nsphum
  integer :: previous, current
  logical :: module_is_initialized = .false.
  integer :: dt_integer
  type(time_type) :: Time_step
  ! host mirrors, filled on demand (diagnostics that are due, restart)
  real(c_double), allocatable, target :: buf3(:,:,:), buf2(:,:)
  complex(c_double_complex), allocatable :: sbuf3(:,:,:), sbuf2(:,:)
  ! diag_manager ids of module 'dynamics' (spectral_dynamics.F90:1541-1700)
  integer :: id_ps, id_u, id_v, id_t, id_vor, id_div, id_omega, id_sphum, id_pres_full, id_pres_half, id_zfull, id_zhalf, id_pk, id_bk
  ! derived fields of spectral_diagnostics (spectral_dynamics.F90:1613-1700, 1747-1835), formed on the device
  integer, parameter :: n_derived = 18
  character(len=16), parameter :: derived_name(n_derived) = (/ character(len=16) :: 'ucomp_sq', 'vcomp_sq', 'ucomp_vcomp', 'ucomp_omega', &
      'vcomp_omega', 'ucomp_temp', 'vcomp_temp', 'vcomp_vor', 'omega_temp', 'wspd', 'temp_sq', 'omega_sq', 'ucomp_height', &
      'vcomp_height', 'omega_height', 'sphum_u', 'sphum_v', 'sphum_w' /)
  integer(c_int), parameter :: derived_field(n_derived) = (/ ISCA_F_UU, ISCA_F_VV, ISCA_F_UV, ISCA_F_UW, ISCA_F_VW, ISCA_F_UT, ISCA_F_VT, &
      ISCA_F_V_VOR, ISCA_F_OMEGA_T, ISCA_F_WSPD, ISCA_F_TT, ISCA_F_OMEGA_OMEGA, ISCA_F_UZ, ISCA_F_VZ, ISCA_F_OMEGA_Z, ISCA_F_UTR0, &
      ISCA_F_VTR0, ISCA_F_WTR0 /)
  integer :: id_derived(n_derived) = -1, id_slp = -1
  !------------------------------------------------------------------ isca_b200_nml: options of this shim
  ! device_time_average = .true.: every registered field is accumulated on the device each step (isca_b200_diag_accumulate) and the
  ! finished mean is handed to diag_manager when output is due -- diag_table then lists the fields WITHOUT time averaging at the
  ! output interval.  .false. (default): instantaneous values when need_data() says output is due, as the reference's send_data calls
  ! would deliver for snapshot fields; fields the diag_table averages every step go through isca_b200_moist_step_io instead.
  logical :: device_time_average = .false.
  namelist /isca_b200_nml/ device_time_average

contains

  !=================================================================================================================================
  subroutine atmosphere_init(Time_init, Time, Time_step_in)
    type(time_type), intent(in) :: Time_init, Time, Time_step_in
    type(isca_config) :: cfg
    type(isca_physics_config) :: pcfg
    type(isca_moist_config) :: mcfg
    character(kind=c_char), target :: nccl_id(128)
    character(kind=c_char), allocatable :: ipc_mine(:), ipc_all(:)
    integer :: seconds, days, nml_unit, io, rc, nt, slot
    real, dimension(2) :: time_pointers
    integer, dimension(4) :: siz
    character(len=64) :: file
    character(len=256) :: message

    if (module_is_initialized) return
    call write_version_number(version, tagname)

#ifdef INTERNAL_FILE_NML
    read (input_nml_file, nml=atmosphere_nml, iostat=io)
    read (input_nml_file, nml=spectral_dynamics_nml, iostat=io)
    read (input_nml_file, nml=hs_forcing_nml, iostat=io)
    read (input_nml_file, nml=spectral_init_cond_nml, iostat=io)
    read (input_nml_file, nml=isca_b200_nml, iostat=io)
#else
    if (file_exist('input.nml')) then
      nml_unit = open_namelist_file()
      read (nml_unit, atmosphere_nml, iostat=io);          rewind(nml_unit)
      read (nml_unit, spectral_dynamics_nml, iostat=io);   rewind(nml_unit)
      read (nml_unit, hs_forcing_nml, iostat=io);          rewind(nml_unit)
      read (nml_unit, spectral_init_cond_nml, iostat=io);  rewind(nml_unit)
      read (nml_unit, isca_b200_nml, iostat=io)
      call close_file(nml_unit)
    end if
#endif
    write (stdlog(), atmosphere_nml)
    write (stdlog(), spectral_dynamics_nml)

    Time_step = Time_step_in
    call get_time(Time_step, seconds, days)
    dt_integer = 86400*days + seconds

    call get_number_tracers(MODEL_ATMOS, num_prog=num_tracers)
    nsphum = get_tracer_index(MODEL_ATMOS, 'sphum')

    ! options of the reference the library does not carry are FATAL here, with the library's own checks behind them
    if (.not. triang_trunc)   call error_mesg('atmosphere_init', 'isca_b200: rhomboidal truncation (triang_trunc = .false.) is not built', FATAL)
    if (fourier_inc /= 1)     call error_mesg('atmosphere_init', 'isca_b200: fourier_inc must be 1', FATAL)
    if (num_steps /= 1)       call error_mesg('atmosphere_init', 'isca_b200: num_steps must be 1', FATAL)
    if (trim(damping_option) /= 'resolution_dependent') &
      call error_mesg('atmosphere_init', 'isca_b200: damping_option must be resolution_dependent', FATAL)
    if (trim(vert_difference_option) /= 'simmons_and_burridge') &
      call error_mesg('atmosphere_init', 'isca_b200: vert_difference_option must be simmons_and_burridge', FATAL)
    if (trim(initial_state_option) /= 'quiescent') &
      call error_mesg('atmosphere_init', 'isca_b200: initial_state_option must be quiescent (or a restart)', FATAL)
    if (trim(topography_option) /= 'flat') &
      call error_mesg('atmosphere_init', 'isca_b200: topography is handed over with isca_b200_set_surf_geopotential by a site-specific reader', FATAL)
    if (num_tracers > 1 .or. (num_tracers == 1 .and. nsphum == NO_TRACER)) &
      call error_mesg('atmosphere_init', 'isca_b200: the field_table may hold the sphum grid tracer only', FATAL)
    if (.not. idealized_moist_model) then
      if (trim(equilibrium_t_option) /= 'Held_Suarez' .or. len_trim(local_heating_option) > 0 .or. relax_to_specified_wind) &
        call error_mesg('atmosphere_init', 'isca_b200: this shim forwards the Held_Suarez option of hs_forcing_nml; the other options '// &
                        'are reached through include/isca_b200_hs.h', FATAL)
    end if

    ! FMS domains: latitudes in contiguous blocks, zonal wavenumbers as spec_mpp_mod distributes them for I/O (spec_mpp.F90:61-110)
    call spec_mpp_init(num_fourier, num_spherical, lon_max, lat_max)
    call get_grid_domain(is, ie, js, je)
    call get_spec_domain(ms, me, ns, ne)

    !---------------------------------------------------------------- IscaConfig <- namelists (every field)
    call isca_b200_default_config(cfg)
    cfg%lon_max = lon_max;  cfg%lat_max = lat_max;  cfg%num_fourier = num_fourier;  cfg%num_spherical = num_spherical
    cfg%num_levels = num_levels
    cfg%dt_atmos = real(dt_integer, c_double)
    cfg%damping_order = damping_order;  cfg%damping_order_vor = damping_order_vor;  cfg%damping_order_div = damping_order_div
    cfg%damping_coeff = damping_coeff;  cfg%damping_coeff_vor = damping_coeff_vor;  cfg%damping_coeff_div = damping_coeff_div
    cfg%eddy_sponge_coeff = eddy_sponge_coeff;  cfg%zmu_sponge_coeff = zmu_sponge_coeff;  cfg%zmv_sponge_coeff = zmv_sponge_coeff
    cfg%do_mass_correction = l2i(do_mass_correction);  cfg%do_energy_correction = l2i(do_energy_correction)
    cfg%do_water_correction = l2i(do_water_correction)
    cfg%use_virtual_temperature = l2i(use_virtual_temperature);  cfg%use_implicit = l2i(use_implicit)
    cfg%make_symmetric = l2i(make_symmetric)
    cfg%robert_coeff = robert_coeff;  cfg%raw_filter_coeff = raw_filter_coeff;  cfg%alpha_implicit = alpha_implicit
    select case (trim(vert_coord_option))
      case ('even_sigma');   cfg%vert_coord_option = 0
      case ('uneven_sigma'); cfg%vert_coord_option = 1
      case ('hybrid');       cfg%vert_coord_option = 3
      case default
        call error_mesg('atmosphere_init', '"'//trim(vert_coord_option)//'" is not a valid value for vert_coord_option '// &
                        '(input: pass pk / bk through cfg%pk, cfg%bk)', FATAL)
    end select
    cfg%scale_heights = scale_heights;  cfg%surf_res = surf_res;  cfg%exponent = exponent;  cfg%p_press = p_press;  cfg%p_sigma = p_sigma
    cfg%vert_advect_uv = advect_id(vert_advect_uv);  cfg%vert_advect_t = advect_id(vert_advect_t)
    cfg%reference_sea_level_press = reference_sea_level_press;  cfg%initial_sphum = initial_sphum
    cfg%water_correction_limit = water_correction_limit
    cfg%valid_range_t(1) = valid_range_t(1);  cfg%valid_range_t(2) = valid_range_t(2)
    cfg%initial_temperature = initial_temperature
    cfg%num_tracers = num_tracers
    cfg%tracer_robert_coeff = -1.0_c_double              ! field_table robert_filter default: the model's robert_coeff
    cfg%no_forcing = l2i(no_forcing);  cfg%do_conserve_energy = l2i(do_conserve_energy)
    cfg%t_zero = t_zero;  cfg%t_strat = t_strat;  cfg%delh = delh;  cfg%delv = delv;  cfg%eps = eps;  cfg%sigma_b = sigma_b;  cfg%P00 = P00
    cfg%ka = ka;  cfg%ks = ks;  cfg%kf = kf;  cfg%trflux = trflux;  cfg%trsink = trsink
    cfg%radius = radius;  cfg%omega = omega;  cfg%grav = grav;  cfg%rdgas = rdgas;  cfg%kappa = kappa
    cfg%pk = c_null_ptr;  cfg%bk = c_null_ptr

    !---------------------------------------------------------------- one process per GPU: the NCCL id comes from the root PE
    nccl_id = c_null_char
    if (mpp_npes() > 1) then
      if (mpp_pe() == mpp_root_pe()) then
        rc = isca_b200_nccl_unique_id(nccl_id)
        if (rc /= 0) call fatal('atmosphere_init')
      end if
      call mpp_broadcast(nccl_id, 128, mpp_root_pe())
    end if

    if (idealized_moist_model) then
      rc = isca_b200_physics_default_config(pcfg)        ! scheme namelists: read and forwarded by idealized_moist_nml_to_config
      rc = isca_b200_moist_default_config(mcfg)
      call idealized_moist_nml_to_config(pcfg, mcfg)
      if (mpp_npes() > 1) then
        rc = isca_b200_moist_create_ranked(cfg, pcfg, mcfg, int(mpp_pe(), c_int), int(mpp_npes(), c_int), c_loc(nccl_id), hm)
      else
        rc = isca_b200_moist_create_ranked(cfg, pcfg, mcfg, 0_c_int, 1_c_int, c_null_ptr, hm)
      end if
      if (rc /= 0) call fatal_moist('atmosphere_init')
      h = isca_b200_moist_dycore(hm)
    else
      if (mpp_npes() > 1) then
        rc = isca_b200_create(cfg, int(mpp_pe(), c_int), int(mpp_npes(), c_int), c_loc(nccl_id), h)
      else
        rc = isca_b200_create(cfg, 0_c_int, 1_c_int, c_null_ptr, h)
      end if
      if (rc /= 0) call fatal('atmosphere_init')
    end if

    ! peer-memory transposes inside one node: exchange the CUDA IPC handles of the Fourier buffers (128 bytes per PE)
    if (mpp_npes() > 1) then
      allocate (ipc_mine(128), ipc_all(128*mpp_npes()))
      rc = isca_b200_ipc_handles(h, ipc_mine)
      if (rc /= 0) call fatal('atmosphere_init')
      call mpp_gather(ipc_mine, ipc_all)                 ! every PE needs all handles: gather on the root, then broadcast
      call mpp_broadcast(ipc_all, 128*mpp_npes(), mpp_root_pe())
      rc = isca_b200_set_peer_handles(h, ipc_all)
      if (rc /= 0) call fatal('atmosphere_init')
      deallocate (ipc_mine, ipc_all)
    end if

    allocate (buf3(is:ie, js:je, num_levels + 1), buf2(is:ie, js:je))
    allocate (sbuf3(ms:me, ns:ne, num_levels), sbuf2(ms:me, ns:ne))

    !---------------------------------------------------------------- restart (atmosphere.F90:197-223, spectral_dynamics.F90:509-575) or cold start
    file = 'INPUT/atmosphere.res.nc'
    if (file_exist(trim(file)) .and. file_exist('INPUT/spectral_dynamics.res.nc')) then
      call field_size(trim(file), 'ug', siz)
      if (lon_max /= siz(1) .or. lat_max /= siz(2)) then
        write (message, *) 'Resolution of restart data does not match resolution specified on namelist. Restart data: lon_max=', &
                           siz(1), ', lat_max=', siz(2), '  Namelist: lon_max=', lon_max, ', lat_max=', lat_max
        call error_mesg('atmosphere_init', message, FATAL)
      end if
      ! the library shards zonal wavenumbers in snake order (balanced triangle rows), spec_mpp_mod in contiguous blocks: on several PEs the
      ! spectral restart arrays have to be redistributed (mpp_global_field + a local pick of the library's m list, isca_b200_decomposition)
      if (mpp_npes() > 1) call error_mesg('atmosphere_init', 'isca_b200 shim: restart of the spectral state on several PEs needs the '// &
                                          'redistribution step described in INTEGRATION.md section 3', FATAL)
      call nullify_domain()
      call read_data(trim(file), 'time_pointers', time_pointers)
      previous = int(time_pointers(1));  current = int(time_pointers(2))
      do nt = 1, num_time_levels
        slot = nt - 1
        call upload_grid_level(trim(file), nt, slot)
        call upload_spectral_level('INPUT/spectral_dynamics.res.nc', nt, slot)
      end do
      call read_data('INPUT/spectral_dynamics.res.nc', 'vorg', buf3(:,:,1:num_levels), grid_domain)
      block
        real(c_double), allocatable :: divg(:,:,:)
        allocate (divg(is:ie, js:je, num_levels))
        call read_data('INPUT/spectral_dynamics.res.nc', 'divg', divg, grid_domain)
        rc = isca_b200_set_vor_div_grid(h, buf3(:,:,1:num_levels), divg)
        deallocate (divg)
      end block
      if (rc /= 0) call fatal('atmosphere_init')
      rc = isca_b200_set_time_pointers(h, int(previous - 1, c_int), int(current - 1, c_int))
      if (rc /= 0) call fatal('atmosphere_init')
    else
      previous = 1;  current = 1
      rc = isca_b200_cold_start(h)                       ! spectral_initialize_fields.F90:45-135 on the device
      if (rc /= 0) call fatal('atmosphere_init')
    end if

    if (idealized_moist_model) then
      rc = isca_b200_moist_init(hm)                      ! idealized_moist_phys_init
      if (rc /= 0) call fatal_moist('atmosphere_init')
      if (do_sc_sst) call init_prescribed_sst()          ! mixed_layer_init :318-320 (needs the handle: Gaussian weights from its tables)
      call get_time(Time, seconds, days)
      rc = isca_b200_moist_set_time(hm, int(days, c_long_long), int(seconds, c_int))
    end if

    call register_dynamics_diagnostics(Time)
    module_is_initialized = .true.
  end subroutine atmosphere_init

  !=================================================================================================================================
  subroutine atmosphere(Time)
    type(time_type), intent(in) :: Time
    type(time_type) :: Time_next
    integer :: rc, future

    if (.not. module_is_initialized) call error_mesg('atmosphere', 'atmosphere module is not initialized', FATAL)
    Time_next = Time + Time_step

    ! physics (hs_forcing or idealized_moist_phys) + spectral_dynamics + compute_pressures_and_heights(future) + the time-level swap,
    ! all on the device (atmosphere.F90:276-352)
    if (idealized_moist_model) then
      ! mixed_layer_nml do_sc_sst (mixed_layer.F90:681-691): the SST of the time stepped to is read by interpolator_mod on the host
      ! and handed over; the library's mixed_layer then moves t_surf to it instead of stepping the slab
      if (do_sc_sst) then
        call interpolator(sst_interp, Time_next, sst_new, trim(sst_file))
        rc = isca_b200_moist_set_sst(hm, sst_new)
        if (rc /= 0) call fatal_moist('atmosphere')
      end if
      rc = isca_b200_moist_step(hm, 1_c_int)
      if (rc /= 0) call fatal_moist('atmosphere')
    else
      rc = isca_b200_step(h, 1_c_int)
      if (rc /= 0) call fatal('atmosphere')              ! e.g. 'temperatures out of valid range'
    end if

    if (previous == current) then
      future = num_time_levels + 1 - current
    else
      future = previous
    end if
    previous = current
    current = future

    call send_dynamics_diagnostics(Time_next)            ! spectral_diagnostics (spectral_dynamics.F90:1709-1867): fields that are due
  end subroutine atmosphere

  !=================================================================================================================================
  subroutine atmosphere_end
    integer :: nt, rc
    character(len=64) :: file

    if (.not. module_is_initialized) return
    ! restart files with the reference's variable names (atmosphere.F90:362-375, spectral_dynamics.F90:1502-1531)
    file = 'RESTART/atmosphere.res'
    call nullify_domain()
    call write_data(trim(file), 'time_pointers', (/real(previous), real(current)/))
    do nt = 1, num_time_levels
      rc = isca_b200_get_field(h, ISCA_F_U, int(nt - 1, c_int), buf3);  call write_data(trim(file), 'ug', buf3(:,:,1:num_levels), grid_domain)
      rc = isca_b200_get_field(h, ISCA_F_V, int(nt - 1, c_int), buf3);  call write_data(trim(file), 'vg', buf3(:,:,1:num_levels), grid_domain)
      rc = isca_b200_get_field(h, ISCA_F_T, int(nt - 1, c_int), buf3);  call write_data(trim(file), 'tg', buf3(:,:,1:num_levels), grid_domain)
      rc = isca_b200_get_field(h, ISCA_F_PS, int(nt - 1, c_int), buf2); call write_data(trim(file), 'psg', buf2, grid_domain)
      if (num_tracers == 1) then
        rc = isca_b200_get_field(h, ISCA_F_TRACER0, int(nt - 1, c_int), buf3)
        call write_data(trim(file), 'sphum', buf3(:,:,1:num_levels), grid_domain)
      end if
    end do
    rc = isca_b200_get_field(h, ISCA_F_WG_FULL, ISCA_LEVEL_CURRENT, buf3);  call write_data(trim(file), 'wg_full', buf3(:,:,1:num_levels), grid_domain)
    file = 'RESTART/spectral_dynamics.res'
    if (mpp_npes() > 1) call error_mesg('atmosphere_end', 'isca_b200 shim: the spectral restart on several PEs needs the redistribution '// &
                                        'step described in INTEGRATION.md section 3', FATAL)
    do nt = 1, num_time_levels
      rc = isca_b200_get_spectral(h, ISCA_S_VOR, int(nt - 1, c_int), sbuf3)
      call write_data(trim(file), 'vors_real', real(sbuf3), spectral_domain);  call write_data(trim(file), 'vors_imag', aimag(sbuf3), spectral_domain)
      rc = isca_b200_get_spectral(h, ISCA_S_DIV, int(nt - 1, c_int), sbuf3)
      call write_data(trim(file), 'divs_real', real(sbuf3), spectral_domain);  call write_data(trim(file), 'divs_imag', aimag(sbuf3), spectral_domain)
      rc = isca_b200_get_spectral(h, ISCA_S_T, int(nt - 1, c_int), sbuf3)
      call write_data(trim(file), 'ts_real', real(sbuf3), spectral_domain);    call write_data(trim(file), 'ts_imag', aimag(sbuf3), spectral_domain)
      rc = isca_b200_get_spectral(h, ISCA_S_LNPS, int(nt - 1, c_int), sbuf2)
      call write_data(trim(file), 'ln_ps_real', real(sbuf2), spectral_domain); call write_data(trim(file), 'ln_ps_imag', aimag(sbuf2), spectral_domain)
    end do
    rc = isca_b200_get_field(h, ISCA_F_VOR, ISCA_LEVEL_CURRENT, buf3);  call write_data(trim(file), 'vorg', buf3(:,:,1:num_levels), grid_domain)
    rc = isca_b200_get_field(h, ISCA_F_DIV, ISCA_LEVEL_CURRENT, buf3);  call write_data(trim(file), 'divg', buf3(:,:,1:num_levels), grid_domain)

    call set_domain(grid_domain)
    if (idealized_moist_model) then
      rc = isca_b200_moist_destroy(hm)                   ! also destroys the dynamical core it owns
    else
      rc = isca_b200_destroy(h)
    end if
    h = c_null_ptr;  hm = c_null_ptr
    deallocate (buf3, buf2, sbuf3, sbuf2)
    module_is_initialized = .false.
  end subroutine atmosphere_end

  !=================================================================================================================================
  ! helpers
  !=================================================================================================================================
  integer(c_int32_t) function l2i(flag)
    logical, intent(in) :: flag
    l2i = merge(1_c_int32_t, 0_c_int32_t, flag)
  end function l2i

  integer(c_int32_t) function advect_id(name)
    character(len=*), intent(in) :: name
    if (trim(name) /= 'second_centered') &
      call error_mesg('atmosphere_init', '"'//trim(name)//'" is not a supported vertical advection scheme of u, v, T (second_centered)', FATAL)
    advect_id = 0
  end function advect_id

  !> one time level of INPUT/atmosphere.res.nc -> device (set_grid_state takes the reference's (lon, lat, lev) arrays unchanged)
  subroutine upload_grid_level(file, nt, slot)
    character(len=*), intent(in) :: file
    integer, intent(in) :: nt, slot
    real(c_double), allocatable, target :: u(:,:,:), v(:,:,:), t(:,:,:), q(:,:,:), ps(:,:)
    integer :: rc
    allocate (u(is:ie, js:je, num_levels), v(is:ie, js:je, num_levels), t(is:ie, js:je, num_levels), ps(is:ie, js:je))
    call read_data(file, 'ug', u, grid_domain, timelevel=nt)
    call read_data(file, 'vg', v, grid_domain, timelevel=nt)
    call read_data(file, 'tg', t, grid_domain, timelevel=nt)
    call read_data(file, 'psg', ps, grid_domain, timelevel=nt)
    if (num_tracers == 1) then
      allocate (q(is:ie, js:je, num_levels))
      call read_data(file, 'sphum', q, grid_domain, timelevel=nt)
      rc = isca_b200_set_grid_state(h, int(slot, c_int), u, v, t, ps, c_loc(q))
    else
      rc = isca_b200_set_grid_state(h, int(slot, c_int), u, v, t, ps, c_null_ptr)
    end if
    if (rc /= 0) call fatal('atmosphere_init')
  end subroutine upload_grid_level

  !> one time level of INPUT/spectral_dynamics.res.nc -> device (real / imaginary parts as the reference stores them)
  subroutine upload_spectral_level(file, nt, slot)
    character(len=*), intent(in) :: file
    integer, intent(in) :: nt, slot
    real, allocatable :: re3(:,:,:), im3(:,:,:), re2(:,:), im2(:,:)
    complex(c_double_complex), allocatable :: vo(:,:,:), di(:,:,:), ts(:,:,:), lp(:,:)
    integer :: rc
    allocate (re3(ms:me, ns:ne, num_levels), im3(ms:me, ns:ne, num_levels), re2(ms:me, ns:ne), im2(ms:me, ns:ne))
    allocate (vo(ms:me, ns:ne, num_levels), di(ms:me, ns:ne, num_levels), ts(ms:me, ns:ne, num_levels), lp(ms:me, ns:ne))
    call read_data(file, 'vors_real', re3, spectral_domain, timelevel=nt);  call read_data(file, 'vors_imag', im3, spectral_domain, timelevel=nt)
    vo = cmplx(re3, im3, kind=c_double_complex)
    call read_data(file, 'divs_real', re3, spectral_domain, timelevel=nt);  call read_data(file, 'divs_imag', im3, spectral_domain, timelevel=nt)
    di = cmplx(re3, im3, kind=c_double_complex)
    call read_data(file, 'ts_real', re3, spectral_domain, timelevel=nt);    call read_data(file, 'ts_imag', im3, spectral_domain, timelevel=nt)
    ts = cmplx(re3, im3, kind=c_double_complex)
    call read_data(file, 'ln_ps_real', re2, spectral_domain, timelevel=nt); call read_data(file, 'ln_ps_imag', im2, spectral_domain, timelevel=nt)
    lp = cmplx(re2, im2, kind=c_double_complex)
    rc = isca_b200_set_spectral_state(h, int(slot, c_int), vo, di, ts, lp)
    if (rc /= 0) call fatal('atmosphere_init')
  end subroutine upload_spectral_level

  !> The scheme namelists of the moist model -> the two configuration structs.  Every group is read by its own contained routine (the
  !! groups share variable names: do_simple appears in five of them), declared under the reference's variable names and defaults.  A
  !! group lists the variables the library consumes AND every variable the shipped test cases (exp/test_cases) set in that group, so that
  !! their input.nml files read unchanged; options the library does not carry end in FATAL.  A read error (a variable of the site's
  !! input.nml that the group below does not list) is FATAL as well -- never a silent fall-back to defaults.
  subroutine idealized_moist_nml_to_config(pcfg, mcfg)
    type(isca_physics_config), intent(inout) :: pcfg
    type(isca_moist_config), intent(inout) :: mcfg
    logical :: do_rrtm_radiation, two_stream_gray
    call read_idealized_moist_phys_nml()
    call read_mixed_layer_nml()
    call read_vert_turb_driver_nml()
    call read_diffusivity_nml()
    call read_surface_flux_nml()
    call read_lscale_cond_nml()
    call read_qe_moist_convection_nml()
    call read_two_stream_gray_rad_nml()
    call read_damping_driver_nml()
    call read_sat_vapor_pres_nml()
    pcfg%grav = grav;  pcfg%rdgas = rdgas;  pcfg%cp_air = rdgas/kappa
    if (do_rrtm_radiation .and. two_stream_gray) &
      call error_mesg('atmosphere_init', 'do_rrtm_radiation and two_stream_gray cannot both be .true.', FATAL)
    ! do_rrtm_radiation: rrtm_radiation_nml is forwarded with isca_b200_moist_use_rrtm (include/isca_b200_rrtm.h) by the site's copy of
    ! this routine; the ozone field read by interpolator_mod goes through isca_b200_moist_set_ozone whenever the alarm is due
  contains
    !> one namelist group from input.nml; `reader` does the READ so that the group stays local to its routine
    subroutine check(io, group)
      integer, intent(in) :: io
      character(len=*), intent(in) :: group
      ! io > 0: the group is present but holds a variable this shim does not declare (or a malformed value); io < 0: group absent (defaults)
      if (io > 0) call error_mesg('atmosphere_init', 'isca_b200 shim: error reading '//group//' (a variable of input.nml that '// &
                                  'fortran/atmosphere.F90 does not declare in this group?)', FATAL)
    end subroutine check
    integer function open_nml()
#ifdef INTERNAL_FILE_NML
      open_nml = -1
#else
      open_nml = open_namelist_file()
#endif
    end function open_nml

    subroutine read_idealized_moist_phys_nml()      ! idealized_moist_phys.F90:109-183
      logical :: turb = .false., do_virtual = .false., do_damping = .false., mixed_layer_bc = .false., do_simple = .false., bucket = .false., &
                 do_lcl_diffusivity_depth = .false.
      character(len=256) :: convection_scheme = 'UNSET', land_option = 'none', land_file_name = 'INPUT/land.nc'
      real :: roughness_heat = 0.05, roughness_moist = 0.05, roughness_mom = 0.05, init_bucket_depth_land = 20., max_bucket_depth_land = 0.15
      integer :: u, io
      namelist /idealized_moist_phys_nml/ turb, do_virtual, two_stream_gray, do_rrtm_radiation, do_damping, mixed_layer_bc, do_simple, &
                                          convection_scheme, roughness_heat, roughness_moist, roughness_mom, bucket, &
                                          do_lcl_diffusivity_depth, land_option, land_file_name, init_bucket_depth_land, &
                                          max_bucket_depth_land
      two_stream_gray = .true.;  do_rrtm_radiation = .false.
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=idealized_moist_phys_nml, iostat=io)
#else
      u = open_nml();  read (u, idealized_moist_phys_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'idealized_moist_phys_nml')
      if (.not. (turb .and. mixed_layer_bc)) &
        call error_mesg('atmosphere_init', 'isca_b200: idealized_moist_phys needs turb = .true. and mixed_layer_bc = .true.', FATAL)
      if (bucket) call error_mesg('atmosphere_init', 'isca_b200: bucket hydrology is not built', FATAL)
      if (do_lcl_diffusivity_depth) call error_mesg('atmosphere_init', 'isca_b200: do_lcl_diffusivity_depth is not built', FATAL)
      select case (trim(convection_scheme))
        case ('NONE', 'none');              mcfg%convection_scheme = 0
        case ('SIMPLE_BETTS_MILLER');       mcfg%convection_scheme = 1
        case ('DRY', 'dry');                mcfg%convection_scheme = 2
        case ('FULL_BETTS_MILLER');         mcfg%convection_scheme = 3
        case default
          call error_mesg('atmosphere_init', '"'//trim(convection_scheme)//'" is not a convection scheme of this library', FATAL)
      end select
      mcfg%do_damping = l2i(do_damping)
      mcfg%roughness_mom = roughness_mom;  mcfg%roughness_heat = roughness_heat;  mcfg%roughness_moist = roughness_moist
      pcfg%use_virtual_temp_vert_diff = l2i(do_virtual)
      ! land_option / land_file_name: the maps derived from them are handed over with isca_b200_moist_set_surface by the site's land set-up
    end subroutine read_idealized_moist_phys_nml

    subroutine read_mixed_layer_nml()               ! mixed_layer.F90:84-156
      real :: depth = 40.0, albedo_value = 0.06, tconst = 305.0, delta_T = 40.0, land_albedo_prefactor = 1.0, land_h_capacity_prefactor = 1.0
      logical :: evaporation = .true., prescribe_initial_dist = .false., do_qflux = .false.
      character(len=256) :: land_option = 'none'
      integer :: u, io
      namelist /mixed_layer_nml/ depth, albedo_value, evaporation, tconst, delta_T, prescribe_initial_dist, do_qflux, do_sc_sst, sst_file, &
                                 land_option, land_albedo_prefactor, land_h_capacity_prefactor      ! do_sc_sst, sst_file: module variables
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=mixed_layer_nml, iostat=io)
#else
      u = open_nml();  read (u, mixed_layer_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'mixed_layer_nml')
      mcfg%mixed_layer_depth = depth;  mcfg%albedo_value = albedo_value
      pcfg%evaporation = l2i(evaporation)
      ! prescribe_initial_dist / tconst / delta_T: the initial t_surf (mixed_layer.F90:347) is set with isca_b200_moist_set_t_surf after
      ! isca_b200_moist_init; do_qflux: isca_b200_moist_set_ocean_qflux with the field of qflux_mod; land_*: isca_b200_moist_set_surface
    end subroutine read_mixed_layer_nml

    subroutine read_vert_turb_driver_nml()          ! vert_turb_driver.F90:100-118
      logical :: use_tau = .true., do_mellor_yamada = .true., do_diffusivity = .false., do_simple = .false., do_shallow_conv = .false., &
                 do_edt = .false., do_entrain = .false.
      real :: constant_gust = 1.0, gust_factor = 1.0
      character(len=16) :: gust_scheme = 'constant'
      integer :: u, io
      namelist /vert_turb_driver_nml/ use_tau, constant_gust, do_mellor_yamada, do_diffusivity, do_simple, do_shallow_conv, do_edt, &
                                      do_entrain, gust_scheme, gust_factor
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=vert_turb_driver_nml, iostat=io)
#else
      u = open_nml();  read (u, vert_turb_driver_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'vert_turb_driver_nml')
      if (do_mellor_yamada .or. .not. do_diffusivity .or. do_shallow_conv .or. do_edt .or. do_entrain .or. trim(gust_scheme) /= 'constant') &
        call error_mesg('vert_turb_driver', 'isca_b200: only do_diffusivity = .true. with gust_scheme = constant is built', FATAL)
      mcfg%use_tau = l2i(use_tau);  mcfg%constant_gust = constant_gust
    end subroutine read_vert_turb_driver_nml

    subroutine read_diffusivity_nml()               ! diffusivity.F90:124-153
      logical :: fixed_depth = .false., free_atm_diff = .false., free_atm_skyhi_diff = .false., pbl_mcm = .false., ampns = .false., &
                 do_entrain = .true., do_simple = .false., use_pog_bug_fix = .true.
      real :: depth_0 = 5000.0, frac_inner = 0.1, rich_crit_pbl = 1.0, entr_ratio = 0.2, parcel_buoy = 2.0, znom = 1000.0, &
              rich_crit_diff = 0.25, mix_len = 30., rich_prandtl = 1.0, background_m = 0.0, background_t = 0.0, ampns_max = 1.0E20
      integer :: u, io
      namelist /diffusivity_nml/ fixed_depth, depth_0, frac_inner, rich_crit_pbl, entr_ratio, parcel_buoy, znom, free_atm_diff, &
                                 free_atm_skyhi_diff, pbl_mcm, rich_crit_diff, mix_len, rich_prandtl, background_m, background_t, &
                                 ampns, ampns_max, do_entrain, do_simple, use_pog_bug_fix
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=diffusivity_nml, iostat=io)
#else
      u = open_nml();  read (u, diffusivity_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'diffusivity_nml')
      pcfg%fixed_depth = l2i(fixed_depth);  pcfg%diffusivity_do_entrain = l2i(do_entrain);  pcfg%diffusivity_do_simple = l2i(do_simple)
      pcfg%free_atm_diff = l2i(free_atm_diff);  pcfg%free_atm_skyhi_diff = l2i(free_atm_skyhi_diff);  pcfg%pbl_mcm = l2i(pbl_mcm)
      pcfg%use_pog_bug_fix = l2i(use_pog_bug_fix);  pcfg%ampns = l2i(ampns);  pcfg%ampns_max = ampns_max
      pcfg%depth_0 = depth_0;  pcfg%frac_inner = frac_inner;  pcfg%rich_crit_pbl = rich_crit_pbl;  pcfg%entr_ratio = entr_ratio
      pcfg%parcel_buoy = parcel_buoy;  pcfg%znom = znom;  pcfg%rich_crit_diff = rich_crit_diff;  pcfg%mix_len = mix_len
      pcfg%rich_prandtl = rich_prandtl;  pcfg%background_m = background_m;  pcfg%background_t = background_t
    end subroutine read_diffusivity_nml

    subroutine read_surface_flux_nml()              ! surface_flux.F90:225-253
      logical :: no_neg_q = .false., use_virtual_temp = .true., alt_gustiness = .false., old_dtaudv = .false., use_mixing_ratio = .false., &
                 do_simple = .false., ncar_ocean_flux = .false., ncar_ocean_flux_orig = .false., raoult_sat_vap = .false.
      real :: gust_const = 1.0, gust_min = 0.0, land_humidity_prefactor = 1.0, land_evap_prefactor = 1.0
      integer :: u, io
      namelist /surface_flux_nml/ no_neg_q, use_virtual_temp, alt_gustiness, gust_const, gust_min, old_dtaudv, use_mixing_ratio, &
                                  ncar_ocean_flux, ncar_ocean_flux_orig, raoult_sat_vap, do_simple, land_humidity_prefactor, &
                                  land_evap_prefactor
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=surface_flux_nml, iostat=io)
#else
      u = open_nml();  read (u, surface_flux_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'surface_flux_nml')
      if (ncar_ocean_flux .or. ncar_ocean_flux_orig .or. raoult_sat_vap) &
        call error_mesg('surface_flux', 'isca_b200: ncar_ocean_flux and raoult_sat_vap are not built', FATAL)
      pcfg%no_neg_q = l2i(no_neg_q);  pcfg%use_virtual_temp = l2i(use_virtual_temp);  pcfg%alt_gustiness = l2i(alt_gustiness)
      pcfg%old_dtaudv = l2i(old_dtaudv);  pcfg%use_mixing_ratio = l2i(use_mixing_ratio);  pcfg%surface_flux_do_simple = l2i(do_simple)
      pcfg%gust_const = gust_const;  pcfg%gust_min = gust_min
      pcfg%land_humidity_prefactor = land_humidity_prefactor;  pcfg%land_evap_prefactor = land_evap_prefactor
    end subroutine read_surface_flux_nml

    subroutine read_lscale_cond_nml()               ! lscale_cond.F90:48-52
      real :: hc = 1.0
      logical :: do_evap = .false., do_simple = .false.
      integer :: u, io
      namelist /lscale_cond_nml/ hc, do_evap, do_simple
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=lscale_cond_nml, iostat=io)
#else
      u = open_nml();  read (u, lscale_cond_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'lscale_cond_nml')
      if (.not. do_simple) call error_mesg('lscale_cond', 'isca_b200: lscale_cond_nml do_simple = .true. only (no ice phase; every '// &
                                           'shipped test case sets it)', FATAL)
      pcfg%hc = hc;  pcfg%do_evap = l2i(do_evap)
    end subroutine read_lscale_cond_nml

    subroutine read_qe_moist_convection_nml()       ! qe_moist_convection.F90:61-75
      real :: tau_bm = 7200., rhbm = 0.8, Tmin = 173., Tmax = 335., val_inc = 0.01
      integer :: u, io
      namelist /qe_moist_convection_nml/ tau_bm, rhbm, Tmin, Tmax, val_inc
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=qe_moist_convection_nml, iostat=io)
#else
      u = open_nml();  read (u, qe_moist_convection_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'qe_moist_convection_nml')
      pcfg%tau_bm = tau_bm;  pcfg%rhbm = rhbm;  pcfg%Tmin = Tmin;  pcfg%Tmax = Tmax;  pcfg%val_inc = val_inc
    end subroutine read_qe_moist_convection_nml

    subroutine read_two_stream_gray_rad_nml()       ! two_stream_gray_rad.F90:72-118
      real :: solar_constant = 1360.0, del_sol = 1.4, del_sw = 0.0, ir_tau_eq = 6.0, ir_tau_pole = 1.5, atm_abs = 0.0, sw_diff = 0.0, &
              linear_tau = 0.1, wv_exponent = 4.0, solar_exponent = 4.0, odp = 1.0, equinox_day = 0.75, carbon_conc = 360.0
      logical :: do_seasonal = .false., do_read_co2 = .false., use_time_average_coszen = .false.
      integer :: solday = -10
      character(len=32) :: rad_scheme = 'frierson'
      character(len=256) :: co2_file = 'co2', co2_variable_name = 'co2'
      integer :: u, io
      namelist /two_stream_gray_rad_nml/ solar_constant, del_sol, del_sw, ir_tau_eq, ir_tau_pole, atm_abs, sw_diff, linear_tau, wv_exponent, &
                                         solar_exponent, odp, rad_scheme, do_seasonal, equinox_day, solday, use_time_average_coszen, &
                                         do_read_co2, co2_file, co2_variable_name, carbon_conc
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=two_stream_gray_rad_nml, iostat=io)
#else
      u = open_nml();  read (u, two_stream_gray_rad_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'two_stream_gray_rad_nml')
      pcfg%solar_constant = solar_constant;  pcfg%del_sol = del_sol;  pcfg%del_sw = del_sw;  pcfg%ir_tau_eq = ir_tau_eq
      pcfg%ir_tau_pole = ir_tau_pole;  pcfg%atm_abs = atm_abs;  pcfg%sw_diff = sw_diff;  pcfg%linear_tau = linear_tau
      pcfg%wv_exponent = wv_exponent;  pcfg%solar_exponent = solar_exponent;  pcfg%odp = odp;  pcfg%carbon_conc = carbon_conc
      select case (trim(rad_scheme))
        case ('frierson', 'FRIERSON');   pcfg%rad_scheme = 0
        case ('byrne', 'BYRNE');         pcfg%rad_scheme = 1
        case ('geen', 'GEEN');           pcfg%rad_scheme = 2
        case ('schneider', 'SCHNEIDER'); pcfg%rad_scheme = 3
        case default
          call error_mesg('two_stream_gray_rad', '"'//trim(rad_scheme)//'" is not a valid radiation scheme.', FATAL)
      end select
      ! do_seasonal: isca_b200_moist_set_seasonal (solday, equinox_day, use_time_average_coszen) after create; do_read_co2: the value
      ! interpolator_mod reads from co2_file goes through isca_b200_moist_set_co2 before each step (include/isca_b200_physics.h)
      if (do_seasonal .or. do_read_co2) call error_mesg('two_stream_gray_rad', 'isca_b200 shim: forward do_seasonal / do_read_co2 with '// &
          'isca_b200_moist_set_seasonal / isca_b200_moist_set_co2 in the site''s copy of this routine', FATAL)
    end subroutine read_two_stream_gray_rad_nml

    subroutine read_damping_driver_nml()            ! damping_driver.f90:60-78
      real :: trayfric = 0., sponge_pbottom = 50.
      logical :: do_rayleigh = .false., do_conserve_energy = .false., do_mg_drag = .false., do_cg_drag = .false., do_topo_drag = .false.
      integer :: nlev_rayfric = 1
      integer :: u, io
      namelist /damping_driver_nml/ trayfric, sponge_pbottom, do_rayleigh, do_conserve_energy, nlev_rayfric, do_mg_drag, do_cg_drag, do_topo_drag
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=damping_driver_nml, iostat=io)
#else
      u = open_nml();  read (u, damping_driver_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'damping_driver_nml')
      if (do_mg_drag .or. do_cg_drag .or. do_topo_drag) call error_mesg('damping_driver', 'isca_b200: only the Rayleigh sponge is built', FATAL)
      pcfg%trayfric = merge(trayfric, 0.0, do_rayleigh);  pcfg%sponge_pbottom = sponge_pbottom
      pcfg%do_conserve_energy = l2i(do_conserve_energy)
    end subroutine read_damping_driver_nml

    subroutine read_sat_vapor_pres_nml()            ! sat_vapor_pres.F90 (do_simple = .true. in every Frierson / MiMA test case)
      logical :: do_simple = .false., show_bad_value_count_by_slice = .true., show_all_bad_values = .false., use_exact_qs = .false., &
                 do_not_calculate = .false.
      integer :: u, io
      namelist /sat_vapor_pres_nml/ do_simple, show_bad_value_count_by_slice, show_all_bad_values, use_exact_qs, do_not_calculate
#ifdef INTERNAL_FILE_NML
      read (input_nml_file, nml=sat_vapor_pres_nml, iostat=io)
#else
      u = open_nml();  read (u, sat_vapor_pres_nml, iostat=io);  call close_file(u)
#endif
      call check(io, 'sat_vapor_pres_nml')
      pcfg%sat_vapor_pres_do_simple = l2i(do_simple)
    end subroutine read_sat_vapor_pres_nml
  end subroutine idealized_moist_nml_to_config

  !> mixed_layer_init :318-320 for do_sc_sst: interpolator_init on the cell boundaries of this PE's grid block.  The boundaries are those
  !! of get_grid_boundaries (transforms.F90: longitudes half a cell either side of the points, sin(latitude) boundaries accumulated from
  !! the Gaussian weights), built from the library's host tables.
  subroutine init_prescribed_sst()
    real, allocatable :: lonb2(:,:), latb2(:,:), wts(:), slb(:)
    real(c_double), allocatable :: tab(:)
    integer :: i, j, rc
    interface
      integer(c_int) function isca_b200_get_table(hh, tid, host, count) bind(C, name="isca_b200_get_table")
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: hh
        integer(c_int), value :: tid, count
        real(c_double), intent(out) :: host(*)
      end function isca_b200_get_table
    end interface
    allocate (sst_new(is:ie, js:je), lonb2(is:ie + 1, js:je + 1), latb2(is:ie + 1, js:je + 1), wts(lat_max), slb(lat_max + 1), tab(lat_max))
    rc = isca_b200_get_table(h, 1_c_int, tab, int(lat_max, c_int))                   ! ISCA_TB_WTS_LAT
    if (rc /= 0) call fatal('atmosphere_init')
    wts = tab
    slb(1) = -1.0
    do j = 1, lat_max
      slb(j + 1) = slb(j) + wts(j)
    end do
    slb(lat_max + 1) = 1.0
    do i = is, ie + 1
      lonb2(i, :) = (real(i) - 1.5)*2.0*pi/real(lon_max) + longitude_origin
    end do
    do j = js, je + 1
      latb2(:, j) = asin(slb(j))
    end do
    call interpolator_init(sst_interp, trim(sst_file)//'.nc', lonb2, latb2, data_out_of_bounds=(/CONSTANT/))
    deallocate (lonb2, latb2, wts, slb, tab)
  end subroutine init_prescribed_sst

  !> axes and fields of diag_manager module 'dynamics' (spectral_dynamics.F90:1541-1700), registered under the reference's names
  subroutine register_dynamics_diagnostics(Time)
    type(time_type), intent(in) :: Time
    integer :: id_lon, id_lat, id_pfull, id_phalf, k, j, i
    integer, dimension(3) :: axes_3d_full, axes_3d_half
    real, allocatable :: lon(:), lat(:), p_full_ref(:), p_half_ref(:), tab(:)
    integer :: rc
    allocate (lon(lon_max), lat(lat_max), p_full_ref(num_levels), p_half_ref(num_levels + 1), tab(max(lon_max, lat_max, num_levels + 1)))
    rc = isca_b200_get_table_f(3, lon, lon_max)                       ! ISCA_TB_DEG_LON
    rc = isca_b200_get_table_f(2, lat, lat_max)                       ! ISCA_TB_DEG_LAT
    rc = isca_b200_get_table_f(4, p_half_ref, num_levels + 1)         ! pk
    rc = isca_b200_get_table_f(5, tab, num_levels + 1)                ! bk
    p_half_ref = (p_half_ref + tab(1:num_levels + 1)*reference_sea_level_press)*0.01
    do k = 1, num_levels
      p_full_ref(k) = 0.5*(p_half_ref(k) + p_half_ref(k + 1))
    end do
    id_lon = diag_axis_init('lon', lon, 'degrees_E', 'x', 'longitude', set_name=mod_name, Domain2=grid_domain)
    id_lat = diag_axis_init('lat', lat, 'degrees_N', 'y', 'latitude', set_name=mod_name, Domain2=grid_domain)
    id_phalf = diag_axis_init('phalf', p_half_ref, 'hPa', 'z', 'approx half pressure level', direction=-1, set_name=mod_name)
    id_pfull = diag_axis_init('pfull', p_full_ref, 'hPa', 'z', 'approx full pressure level', direction=-1, set_name=mod_name, edges=id_phalf)
    axes_3d_full = (/id_lon, id_lat, id_pfull/);  axes_3d_half = (/id_lon, id_lat, id_phalf/)
    id_ps    = register_diag_field(mod_name, 'ps',    (/id_lon, id_lat/), Time, 'surface pressure', 'pascals')
    id_u     = register_diag_field(mod_name, 'ucomp', axes_3d_full, Time, 'zonal wind component', 'm/sec')
    id_v     = register_diag_field(mod_name, 'vcomp', axes_3d_full, Time, 'meridional wind component', 'm/sec')
    id_t     = register_diag_field(mod_name, 'temp',  axes_3d_full, Time, 'temperature', 'deg_k')
    id_vor   = register_diag_field(mod_name, 'vor',   axes_3d_full, Time, 'Vorticity', 'sec**-1')
    id_div   = register_diag_field(mod_name, 'div',   axes_3d_full, Time, 'Divergence', 'sec**-1')
    id_omega = register_diag_field(mod_name, 'omega', axes_3d_full, Time, 'dp/dt vertical velocity', 'Pa/sec')
    id_sphum = register_diag_field(mod_name, 'sphum', axes_3d_full, Time, 'specific humidity', 'kg/kg')
    id_pres_full = register_diag_field(mod_name, 'pres_full', axes_3d_full, Time, 'pressure at full model levels', 'pascals')
    id_pres_half = register_diag_field(mod_name, 'pres_half', axes_3d_half, Time, 'pressure at half model levels', 'pascals')
    id_zfull = register_diag_field(mod_name, 'height',      axes_3d_full, Time, 'geopotential height at full model levels', 'm')
    id_zhalf = register_diag_field(mod_name, 'height_half', axes_3d_half, Time, 'geopotential height at half model levels', 'm')
    do k = 1, n_derived
      if (k > 15 .and. num_tracers /= 1) cycle
      id_derived(k) = register_diag_field(mod_name, trim(derived_name(k)), axes_3d_full, Time, trim(derived_name(k)), 'mks')
    end do
    id_slp = register_diag_field(mod_name, 'slp', (/id_lon, id_lat/), Time, 'sea level pressure', 'pascals')
    id_pk = register_static_field(mod_name, 'pk', (/id_phalf/), 'vertical coordinate pressure values', 'pascals')
    id_bk = register_static_field(mod_name, 'bk', (/id_phalf/), 'vertical coordinate sigma values', 'none')
    deallocate (lon, lat, p_full_ref, p_half_ref, tab)
  contains
    integer function isca_b200_get_table_f(table_id, dst, n)
      integer, intent(in) :: table_id, n
      real, intent(out) :: dst(:)
      real(c_double), allocatable :: tmp(:)
      interface
        integer(c_int) function isca_b200_get_table(hh, tid, host, count) bind(C, name="isca_b200_get_table")
          import :: c_int, c_ptr, c_double
          type(c_ptr), value :: hh
          integer(c_int), value :: tid, count
          real(c_double), intent(out) :: host(*)
        end function isca_b200_get_table
      end interface
      allocate (tmp(n))
      isca_b200_get_table_f = isca_b200_get_table(h, int(table_id, c_int), tmp, int(n, c_int))
      dst(1:n) = tmp
      deallocate (tmp)
    end function isca_b200_get_table_f
  end subroutine register_dynamics_diagnostics

  !> send_data for the registered fields whose output is due at Time_next (only those are copied off the device)
  subroutine send_dynamics_diagnostics(Time_next)
    type(time_type), intent(in) :: Time_next
    logical :: used
    integer :: rc, k
    integer(c_int) :: cnt
    call send2(id_ps, ISCA_F_PS);                    call send2(id_slp, ISCA_F_SLP)
    call send3(id_u, ISCA_F_U, num_levels);          call send3(id_v, ISCA_F_V, num_levels)
    call send3(id_t, ISCA_F_T, num_levels);          call send3(id_vor, ISCA_F_VOR, num_levels)
    call send3(id_div, ISCA_F_DIV, num_levels);      call send3(id_omega, ISCA_F_WG_FULL, num_levels)
    call send3(id_pres_full, ISCA_F_P_FULL, num_levels);  call send3(id_pres_half, ISCA_F_P_HALF, num_levels + 1)
    call send3(id_zfull, ISCA_F_Z_FULL, num_levels);      call send3(id_zhalf, ISCA_F_Z_HALF, num_levels + 1)
    if (num_tracers == 1) call send3(id_sphum, ISCA_F_TRACER0, num_levels)
    do k = 1, n_derived
      call send3(id_derived(k), derived_field(k), num_levels)
    end do
  contains
    !> the field (instantaneous, or its device-side mean since the last output) into buf3 / buf2 when output is due
    logical function due(id, field, host)
      integer, intent(in) :: id
      integer(c_int), intent(in) :: field
      real(c_double), intent(out) :: host(*)
      due = .false.
      if (id <= 0) return
      if (device_time_average) then
        rc = isca_b200_diag_accumulate(h, field)
        if (rc /= 0) call fatal('spectral_diagnostics')
        if (.not. need_data(id, Time_next)) return
        rc = isca_b200_diag_fetch(h, field, host, 1_c_int, cnt)
      else
        if (.not. need_data(id, Time_next)) return
        rc = isca_b200_get_field(h, field, ISCA_LEVEL_CURRENT, host)
      end if
      if (rc /= 0) call fatal('spectral_diagnostics')
      due = .true.
    end function due
    subroutine send3(id, field, nlev)
      integer, intent(in) :: id, nlev
      integer(c_int), intent(in) :: field
      if (due(id, field, buf3)) used = send_data(id, buf3(:,:,1:nlev), Time_next)
    end subroutine send3
    subroutine send2(id, field)
      integer, intent(in) :: id
      integer(c_int), intent(in) :: field
      if (due(id, field, buf2)) used = send_data(id, buf2, Time_next)
    end subroutine send2
  end subroutine send_dynamics_diagnostics

  subroutine fatal(routine)
    character(len=*), intent(in) :: routine
    call error_mesg(routine, isca_c_string(isca_b200_last_error(h)), FATAL)
  end subroutine fatal

  subroutine fatal_moist(routine)
    character(len=*), intent(in) :: routine
    call error_mesg(routine, isca_c_string(isca_b200_moist_last_error(hm)), FATAL)
  end subroutine fatal_moist

end module atmosphere_mod
