!> isca_b200_c -- ISO_C_BINDING interface of libisca_b200.so (include/isca_b200.h, include/isca_b200_physics.h) for the Fortran host
!! side of the reference: the replacement atmosphere_mod of fortran/atmosphere.F90 uses it, atmos_model.F90 and FMS stay unchanged.
!!
!! The derived types mirror the C structs field by field (bind(C) gives the C layout); tools/gen_fortran_interface.py regenerates the
!! type definitions from the headers and tests/test_fortran_shim.py checks that this file is in step with them.  No Fortran compiler
!! exists in the build image of this repository: the file is exercised through the C driver tests/host/shim_driver.c, which makes the
!! same calls with arrays in Fortran order.
module isca_b200_c
  use iso_c_binding
  implicit none
  public

  integer(c_int32_t), parameter :: ISCA_B200_ABI_VERSION = 2
  ! field ids of isca_b200_get_field / isca_b200_get_spectral, time-level selectors, scalar ids (include/isca_b200.h)
  integer(c_int), parameter :: ISCA_F_PS = 0, ISCA_F_U = 1, ISCA_F_V = 2, ISCA_F_T = 3, ISCA_F_VOR = 4, ISCA_F_DIV = 5, &
                               ISCA_F_WG_FULL = 6, ISCA_F_P_FULL = 7, ISCA_F_P_HALF = 8, ISCA_F_Z_FULL = 9, &
                               ISCA_F_Z_HALF = 10, ISCA_F_TRACER0 = 16, ISCA_F_WSPD = 32, ISCA_F_UU = 33, ISCA_F_VV = 34, &
                               ISCA_F_UV = 35, ISCA_F_V_VOR = 36, ISCA_F_TT = 37, ISCA_F_OMEGA_OMEGA = 38, &
                               ISCA_F_OMEGA_T = 39, ISCA_F_UW = 40, ISCA_F_VW = 41, ISCA_F_UT = 42, ISCA_F_VT = 43, &
                               ISCA_F_UZ = 44, ISCA_F_VZ = 45, ISCA_F_OMEGA_Z = 46, ISCA_F_UTR0 = 48, ISCA_F_VTR0 = 49, &
                               ISCA_F_WTR0 = 50, ISCA_F_SLP = 56
  integer(c_int), parameter :: ISCA_S_VOR = 0, ISCA_S_DIV = 1, ISCA_S_T = 2, ISCA_S_LNPS = 3
  integer(c_int), parameter :: ISCA_LEVEL_CURRENT = -1, ISCA_LEVEL_PREVIOUS = -2
  integer(c_int), parameter :: ISCA_SC_MEAN_PS = 0, ISCA_SC_MEAN_ENERGY = 1, ISCA_SC_T_MIN = 2, ISCA_SC_T_MAX = 3
  ! field ids of isca_b200_moist_get (isca_b200/moist.py FIELDS_2D)
  integer(c_int), parameter :: ISCA_M_T_SURF = 0, ISCA_M_PRECIP = 1, ISCA_M_FLUX_T = 2, ISCA_M_FLUX_Q = 3, ISCA_M_Z_PBL = 4, &
                               ISCA_M_NET_SW = 5, ISCA_M_LW_DOWN = 6, ISCA_M_CONV_RAIN = 7, ISCA_M_CAPE = 8, ISCA_M_OLR = 17, ISCA_M_TOA_SW = 18

  ! struct IscaConfig: spectral_dynamics_nml, hs_forcing_nml, spectral_init_cond_nml, constants_nml, main_nml dt_atmos
  type, bind(C) :: isca_config
    integer(c_int32_t)   :: abi_version
    integer(c_int32_t)   :: lon_max
    integer(c_int32_t)   :: lat_max
    integer(c_int32_t)   :: num_fourier
    integer(c_int32_t)   :: num_spherical
    integer(c_int32_t)   :: num_levels
    real(c_double)       :: dt_atmos
    integer(c_int32_t)   :: damping_order
    integer(c_int32_t)   :: damping_order_vor
    integer(c_int32_t)   :: damping_order_div
    real(c_double)       :: damping_coeff
    real(c_double)       :: damping_coeff_vor
    real(c_double)       :: damping_coeff_div
    real(c_double)       :: eddy_sponge_coeff
    real(c_double)       :: zmu_sponge_coeff
    real(c_double)       :: zmv_sponge_coeff
    integer(c_int32_t)   :: do_mass_correction
    integer(c_int32_t)   :: do_energy_correction
    integer(c_int32_t)   :: do_water_correction
    integer(c_int32_t)   :: use_virtual_temperature
    integer(c_int32_t)   :: use_implicit
    integer(c_int32_t)   :: make_symmetric
    real(c_double)       :: robert_coeff
    real(c_double)       :: raw_filter_coeff
    real(c_double)       :: alpha_implicit
    integer(c_int32_t)   :: vert_coord_option
    real(c_double)       :: scale_heights
    real(c_double)       :: surf_res
    real(c_double)       :: exponent
    real(c_double)       :: p_press
    real(c_double)       :: p_sigma
    integer(c_int32_t)   :: vert_advect_uv
    integer(c_int32_t)   :: vert_advect_t
    real(c_double)       :: reference_sea_level_press
    real(c_double)       :: initial_sphum
    real(c_double)       :: water_correction_limit
    real(c_double)       :: valid_range_t(2)
    real(c_double)       :: initial_temperature
    integer(c_int32_t)   :: num_tracers
    real(c_double)       :: tracer_robert_coeff
    integer(c_int32_t)   :: no_forcing
    integer(c_int32_t)   :: do_conserve_energy
    real(c_double)       :: t_zero
    real(c_double)       :: t_strat
    real(c_double)       :: delh
    real(c_double)       :: delv
    real(c_double)       :: eps
    real(c_double)       :: sigma_b
    real(c_double)       :: P00
    real(c_double)       :: ka
    real(c_double)       :: ks
    real(c_double)       :: kf
    real(c_double)       :: trflux
    real(c_double)       :: trsink
    real(c_double)       :: radius
    real(c_double)       :: omega
    real(c_double)       :: grav
    real(c_double)       :: rdgas
    real(c_double)       :: kappa
    type(c_ptr)          :: pk
    type(c_ptr)          :: bk
  end type isca_config

  ! struct IscaPhysicsConfig: the scheme namelists of idealized_moist_phys
  type, bind(C) :: isca_physics_config
    integer(c_int)       :: abi_version
    integer(c_int)       :: num_lon
    integer(c_int)       :: num_lat
    integer(c_int)       :: num_levels
    real(c_double)       :: grav
    real(c_double)       :: rdgas
    real(c_double)       :: rvgas
    real(c_double)       :: cp_air
    real(c_double)       :: hlv
    real(c_double)       :: tfreeze
    real(c_double)       :: stefan
    real(c_double)       :: pstd_mks
    real(c_double)       :: es0
    real(c_double)       :: hc
    integer(c_int)       :: do_evap
    real(c_double)       :: solar_constant
    real(c_double)       :: del_sol
    real(c_double)       :: del_sw
    real(c_double)       :: ir_tau_eq
    real(c_double)       :: ir_tau_pole
    real(c_double)       :: atm_abs
    real(c_double)       :: sw_diff
    real(c_double)       :: linear_tau
    real(c_double)       :: wv_exponent
    real(c_double)       :: solar_exponent
    real(c_double)       :: odp
    real(c_double)       :: diabatic_acce
    real(c_double)       :: trayfric
    real(c_double)       :: sponge_pbottom
    integer(c_int)       :: do_conserve_energy
    integer(c_int)       :: vert_diff_do_conserve_energy
    integer(c_int)       :: use_virtual_temp_vert_diff
    integer(c_int)       :: evaporation
    real(c_double)       :: rich_crit
    real(c_double)       :: drag_min
    real(c_double)       :: zeta_trans
    real(c_double)       :: vonkarm
    integer(c_int)       :: neutral
    integer(c_int)       :: stable_option
    integer(c_int)       :: no_neg_q
    integer(c_int)       :: use_virtual_temp
    integer(c_int)       :: alt_gustiness
    integer(c_int)       :: old_dtaudv
    integer(c_int)       :: use_mixing_ratio
    integer(c_int)       :: surface_flux_do_simple
    real(c_double)       :: gust_const
    real(c_double)       :: gust_min
    real(c_double)       :: land_humidity_prefactor
    real(c_double)       :: land_evap_prefactor
    integer(c_int)       :: fixed_depth
    integer(c_int)       :: diffusivity_do_entrain
    integer(c_int)       :: diffusivity_do_simple
    integer(c_int)       :: free_atm_diff
    integer(c_int)       :: pbl_mcm
    integer(c_int)       :: use_pog_bug_fix
    real(c_double)       :: depth_0
    real(c_double)       :: frac_inner
    real(c_double)       :: rich_crit_pbl
    real(c_double)       :: entr_ratio
    real(c_double)       :: parcel_buoy
    real(c_double)       :: znom
    real(c_double)       :: background_m
    real(c_double)       :: background_t
    real(c_double)       :: tau_bm
    real(c_double)       :: rhbm
    real(c_double)       :: Tmin
    real(c_double)       :: Tmax
    real(c_double)       :: val_inc
    integer(c_int)       :: rad_scheme
    real(c_double)       :: ir_tau_co2_win
    real(c_double)       :: ir_tau_wv_win1
    real(c_double)       :: ir_tau_wv_win2
    real(c_double)       :: ir_tau_co2
    real(c_double)       :: ir_tau_wv1
    real(c_double)       :: ir_tau_wv2
    real(c_double)       :: window
    real(c_double)       :: carbon_conc
    real(c_double)       :: single_albedo
    real(c_double)       :: back_scatter
    real(c_double)       :: lw_tau_0_gp
    real(c_double)       :: sw_tau_0_gp
    real(c_double)       :: lw_tau_exponent_gp
    real(c_double)       :: sw_tau_exponent_gp
    real(c_double)       :: bog_a
    real(c_double)       :: bog_b
    real(c_double)       :: bog_mu
    integer(c_int)       :: sat_vapor_pres_do_simple
    integer(c_int)       :: free_atm_skyhi_diff
    integer(c_int)       :: ampns
    real(c_double)       :: rich_crit_diff
    real(c_double)       :: mix_len
    real(c_double)       :: rich_prandtl
    real(c_double)       :: ampns_max
  end type isca_physics_config

  ! struct IscaMoistConfig: idealized_moist_phys_nml, mixed_layer_nml, vert_turb_driver_nml
  type, bind(C) :: isca_moist_config
    integer(c_int)       :: abi_version
    integer(c_int)       :: convection_scheme
    integer(c_int)       :: do_damping
    real(c_double)       :: roughness_mom
    real(c_double)       :: roughness_heat
    real(c_double)       :: roughness_moist
    real(c_double)       :: mixed_layer_depth
    real(c_double)       :: albedo_value
    real(c_double)       :: rho_cp
    real(c_double)       :: constant_gust
    integer(c_int)       :: use_tau
  end type isca_moist_config

  interface
    ! ---------------------------------------------------------------- dynamical core (atmosphere_mod with hs_forcing)
    subroutine isca_b200_default_config(cfg) bind(C, name="isca_b200_default_config")
      import :: isca_config
      type(isca_config), intent(out) :: cfg
    end subroutine isca_b200_default_config
    integer(c_int) function isca_b200_create(cfg, rank, nranks, nccl_unique_id, h) bind(C, name="isca_b200_create")
      import :: isca_config, c_int, c_ptr
      type(isca_config), intent(in) :: cfg
      integer(c_int), value :: rank, nranks
      type(c_ptr), value :: nccl_unique_id          ! 128 bytes from isca_b200_nccl_unique_id on the root PE, broadcast; c_null_ptr for 1 PE
      type(c_ptr), intent(out) :: h
    end function isca_b200_create
    integer(c_int) function isca_b200_destroy(h) bind(C, name="isca_b200_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function isca_b200_destroy
    type(c_ptr) function isca_b200_last_error(h) bind(C, name="isca_b200_last_error")
      import :: c_ptr
      type(c_ptr), value :: h
    end function isca_b200_last_error
    integer(c_int) function isca_b200_nccl_unique_id(out128) bind(C, name="isca_b200_nccl_unique_id")
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: out128(128)
    end function isca_b200_nccl_unique_id
    integer(c_int) function isca_b200_ipc_handles(h, out128) bind(C, name="isca_b200_ipc_handles")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: h
      character(kind=c_char), intent(out) :: out128(128)
    end function isca_b200_ipc_handles
    integer(c_int) function isca_b200_set_peer_handles(h, all_handles) bind(C, name="isca_b200_set_peer_handles")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: all_handles(*)          ! 128 bytes per PE, in PE order
    end function isca_b200_set_peer_handles
    integer(c_int) function isca_b200_cold_start(h) bind(C, name="isca_b200_cold_start")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function isca_b200_cold_start
    integer(c_int) function isca_b200_set_grid_state(h, slot, ug, vg, tg, psg, tracers) bind(C, name="isca_b200_set_grid_state")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: slot
      real(c_double), intent(in) :: ug(*), vg(*), tg(*), psg(*)
      type(c_ptr), value :: tracers                 ! c_loc of grid_tracers(:,:,:,slot,nsphum) or c_null_ptr
    end function isca_b200_set_grid_state
    integer(c_int) function isca_b200_set_spectral_state(h, slot, vors, divs, ts, ln_ps) bind(C, name="isca_b200_set_spectral_state")
      import :: c_int, c_ptr, c_double_complex
      type(c_ptr), value :: h
      integer(c_int), value :: slot
      complex(c_double_complex), intent(in) :: vors(*), divs(*), ts(*), ln_ps(*)
    end function isca_b200_set_spectral_state
    integer(c_int) function isca_b200_set_vor_div_grid(h, vorg, divg) bind(C, name="isca_b200_set_vor_div_grid")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: vorg(*), divg(*)
    end function isca_b200_set_vor_div_grid
    integer(c_int) function isca_b200_set_surf_geopotential(h, sg) bind(C, name="isca_b200_set_surf_geopotential")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: sg(*)
    end function isca_b200_set_surf_geopotential
    integer(c_int) function isca_b200_set_time_pointers(h, previous_slot, current_slot) bind(C, name="isca_b200_set_time_pointers")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: previous_slot, current_slot
    end function isca_b200_set_time_pointers
    integer(c_int) function isca_b200_get_time_pointers(h, previous_slot, current_slot) bind(C, name="isca_b200_get_time_pointers")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), intent(out) :: previous_slot, current_slot
    end function isca_b200_get_time_pointers
    integer(c_int) function isca_b200_step(h, n_steps) bind(C, name="isca_b200_step")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: n_steps
    end function isca_b200_step
    integer(c_int) function isca_b200_get_field(h, field_id, level, host) bind(C, name="isca_b200_get_field")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field_id, level
      real(c_double), intent(out) :: host(*)
    end function isca_b200_get_field
    integer(c_int) function isca_b200_get_spectral(h, field_id, level, host) bind(C, name="isca_b200_get_spectral")
      import :: c_int, c_ptr, c_double_complex
      type(c_ptr), value :: h
      integer(c_int), value :: field_id, level
      complex(c_double_complex), intent(out) :: host(*)
    end function isca_b200_get_spectral
    integer(c_int) function isca_b200_get_scalar(h, scalar_id, value) bind(C, name="isca_b200_get_scalar")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: scalar_id
      real(c_double), intent(out) :: value
    end function isca_b200_get_scalar
    integer(c_int) function isca_b200_spectral_dynamics_tracers(h, dt_psg, dt_ug, dt_vg, dt_tg, dt_tracers, psg, ug, vg, tg, &
                                                                grid_tracers, wg_full, p_full) bind(C, name="isca_b200_spectral_dynamics_tracers")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, dt_psg, dt_ug, dt_vg, dt_tg, dt_tracers, psg, ug, vg, tg, grid_tracers, wg_full, p_full
    end function isca_b200_spectral_dynamics_tracers
    integer(c_int) function isca_b200_diag_accumulate(h, field_id) bind(C, name="isca_b200_diag_accumulate")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: field_id
    end function isca_b200_diag_accumulate
    integer(c_int) function isca_b200_diag_fetch(h, field_id, host, reset, count_out) bind(C, name="isca_b200_diag_fetch")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field_id, reset
      real(c_double), intent(out) :: host(*)
      integer(c_int), intent(out) :: count_out
    end function isca_b200_diag_fetch
    integer(c_int) function isca_b200_spherical_to_grid(h, spec, grid, nlev) bind(C, name="isca_b200_spherical_to_grid")
      import :: c_int, c_ptr, c_double, c_double_complex
      type(c_ptr), value :: h
      complex(c_double_complex), intent(in) :: spec(*)
      real(c_double), intent(out) :: grid(*)
      integer(c_int), value :: nlev
    end function isca_b200_spherical_to_grid
    integer(c_int) function isca_b200_grid_to_spherical(h, grid, spec, nlev, do_truncation) bind(C, name="isca_b200_grid_to_spherical")
      import :: c_int, c_ptr, c_double, c_double_complex
      type(c_ptr), value :: h
      real(c_double), intent(in) :: grid(*)
      complex(c_double_complex), intent(out) :: spec(*)
      integer(c_int), value :: nlev, do_truncation
    end function isca_b200_grid_to_spherical
    ! ---------------------------------------------------------------- idealized moist model (atmosphere_mod with idealized_moist_phys)
    integer(c_int) function isca_b200_physics_default_config(cfg) bind(C, name="isca_b200_physics_default_config")
      import :: c_int, isca_physics_config
      type(isca_physics_config), intent(out) :: cfg
    end function isca_b200_physics_default_config
    integer(c_int) function isca_b200_moist_default_config(cfg) bind(C, name="isca_b200_moist_default_config")
      import :: c_int, isca_moist_config
      type(isca_moist_config), intent(out) :: cfg
    end function isca_b200_moist_default_config
    integer(c_int) function isca_b200_moist_create_ranked(dyn, phys, mc, rank, nranks, nccl_unique_id, m) &
        bind(C, name="isca_b200_moist_create_ranked")
      import :: c_int, c_ptr, isca_config, isca_physics_config, isca_moist_config
      type(isca_config), intent(in) :: dyn
      type(isca_physics_config), intent(in) :: phys
      type(isca_moist_config), intent(in) :: mc
      integer(c_int), value :: rank, nranks
      type(c_ptr), value :: nccl_unique_id
      type(c_ptr), intent(out) :: m
    end function isca_b200_moist_create_ranked
    integer(c_int) function isca_b200_moist_destroy(m) bind(C, name="isca_b200_moist_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: m
    end function isca_b200_moist_destroy
    type(c_ptr) function isca_b200_moist_last_error(m) bind(C, name="isca_b200_moist_last_error")
      import :: c_ptr
      type(c_ptr), value :: m
    end function isca_b200_moist_last_error
    type(c_ptr) function isca_b200_moist_dycore(m) bind(C, name="isca_b200_moist_dycore")
      import :: c_ptr
      type(c_ptr), value :: m
    end function isca_b200_moist_dycore
    integer(c_int) function isca_b200_moist_init(m) bind(C, name="isca_b200_moist_init")
      import :: c_int, c_ptr
      type(c_ptr), value :: m
    end function isca_b200_moist_init
    integer(c_int) function isca_b200_moist_step(m, n_steps) bind(C, name="isca_b200_moist_step")
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: n_steps
    end function isca_b200_moist_step
    integer(c_int) function isca_b200_moist_get(m, id, host) bind(C, name="isca_b200_moist_get")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: m
      integer(c_int), value :: id
      real(c_double), intent(out) :: host(*)
    end function isca_b200_moist_get
    integer(c_int) function isca_b200_moist_set_t_surf(m, host) bind(C, name="isca_b200_moist_set_t_surf")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: m
      real(c_double), intent(in) :: host(*)
    end function isca_b200_moist_set_t_surf
    integer(c_int) function isca_b200_moist_set_ocean_qflux(m, host) bind(C, name="isca_b200_moist_set_ocean_qflux")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: m
      real(c_double), intent(in) :: host(*)
    end function isca_b200_moist_set_ocean_qflux
    integer(c_int) function isca_b200_moist_set_ozone(m, o3) bind(C, name="isca_b200_moist_set_ozone")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: m
      real(c_double), intent(in) :: o3(*)
    end function isca_b200_moist_set_ozone
    integer(c_int) function isca_b200_moist_set_time(m, days, seconds) bind(C, name="isca_b200_moist_set_time")
      import :: c_int, c_ptr, c_long_long
      type(c_ptr), value :: m
      integer(c_long_long), value :: days
      integer(c_int), value :: seconds
    end function isca_b200_moist_set_time
    integer(c_int) function isca_b200_moist_use_rrtm(m, rc, dc, table_path) bind(C, name="isca_b200_moist_use_rrtm")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: m, rc, dc              ! c_loc of IscaRrtmConfig / IscaRrtmDriverConfig (include/isca_b200_rrtm.h)
      character(kind=c_char), intent(in) :: table_path(*)
    end function isca_b200_moist_use_rrtm
  end interface

contains

  !> the C string of isca_b200_last_error / isca_b200_moist_last_error as a Fortran character value
  function isca_c_string(p) result(s)
    type(c_ptr), intent(in) :: p
    character(len=:), allocatable :: s
    character(kind=c_char), pointer :: c(:)
    integer :: n
    if (.not. c_associated(p)) then
      s = ''
      return
    end if
    call c_f_pointer(p, c, [1024])
    n = 0
    do while (n < 1024)
      if (c(n + 1) == c_null_char) exit
      n = n + 1
    end do
    allocate(character(len=n) :: s)
    s = transfer(c(1:n), s)
  end function isca_c_string

end module isca_b200_c
