/* isca_b200_physics.h -- C ABI of the per-column physics kernels.
 *
 * Each entry point replaces one Fortran subroutine called from idealized_moist_phys
 * (atmos_spectral/driver/solo/idealized_moist_phys.F90:733-1127).  Arrays are host pointers in the reference's
 * Fortran memory order (lon fastest, then lat, then level; level 1 = model top), double precision.
 * All functions return 0 on success; on failure isca_b200_physics_last_error() describes the error.
 */
#ifndef ISCA_B200_PHYSICS_H
#define ISCA_B200_PHYSICS_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct IscaPhysics_t* IscaPhysics;

/* physical constants (shared/constants/constants.F90) and the scheme namelists */
typedef struct IscaPhysicsConfig {
  int abi_version;                /* 4 */
  int num_lon, num_lat, num_levels;
  double grav, rdgas, rvgas, cp_air, hlv, tfreeze, stefan, pstd_mks;
  /* sat_vapor_pres_nml: do_simple tables only (sat_vapor_pres_k.F90:161-266) */
  double es0;
  /* lscale_cond_nml (lscale_cond.F90:48-52) */
  double hc; int do_evap;
  /* two_stream_gray_rad_nml, rad_scheme = 'frierson' (two_stream_gray_rad.F90:72-113) */
  double solar_constant, del_sol, del_sw, ir_tau_eq, ir_tau_pole, atm_abs, sw_diff, linear_tau,
         wv_exponent, solar_exponent, odp, diabatic_acce;
  /* damping_driver_nml: trayfric, sponge_pbottom (damping_driver.f90:60-78) */
  double trayfric, sponge_pbottom; int do_conserve_energy;
  /* vert_diff_nml (vert_diff.F90:70-77; do_mcm_plev = .false. only) */
  int vert_diff_do_conserve_energy, use_virtual_temp_vert_diff;
  /* mixed_layer_nml: evaporation (mixed_layer.F90:87); heat capacity, q-flux and albedo maps are set by
   * isca_b200_mixed_layer_init */
  int evaporation;
  /* monin_obukhov_nml (monin_obukhov.F90:76-83) and VONKARM (constants.F90:239) */
  double rich_crit, drag_min, zeta_trans, vonkarm; int neutral, stable_option;
  /* surface_flux_nml (surface_flux.F90:225-253); bucket hydrology, ncar_ocean_flux and raoult_sat_vap are not built */
  int no_neg_q, use_virtual_temp, alt_gustiness, old_dtaudv, use_mixing_ratio, surface_flux_do_simple;
  double gust_const, gust_min, land_humidity_prefactor, land_evap_prefactor;
  /* diffusivity_nml (diffusivity.F90:124-153).  pbl_mcm and use_pog_bug_fix = .false. are rejected at create; the parameters of
   * free_atm_diff are at the end of the struct */
  int fixed_depth, diffusivity_do_entrain, diffusivity_do_simple, free_atm_diff, pbl_mcm, use_pog_bug_fix;
  double depth_0, frac_inner, rich_crit_pbl, entr_ratio, parcel_buoy, znom, background_m, background_t;
  /* qe_moist_convection_nml (qe_moist_convection.F90:61-75) */
  double tau_bm, rhbm, Tmin, Tmax, val_inc;
  /* two_stream_gray_rad_nml, the other values of rad_scheme (two_stream_gray_rad.F90:89-118, 214-238):
   * 0 'frierson', 1 'byrne' (Byrne & O'Gorman 2013), 2 'geen' (Geen et al. 2016, window band + water-vapour shortwave),
   * 3 'schneider' (Schneider & Liu 2009 giant planet).  do_seasonal: isca_b200_two_stream_gray_rad_set_insolation /
   * isca_b200_moist_set_seasonal; do_read_co2 is not built. */
  int rad_scheme;
  double ir_tau_co2_win, ir_tau_wv_win1, ir_tau_wv_win2, ir_tau_co2, ir_tau_wv1, ir_tau_wv2, window, carbon_conc;
  double single_albedo, back_scatter, lw_tau_0_gp, sw_tau_0_gp, lw_tau_exponent_gp, sw_tau_exponent_gp;
  double bog_a, bog_b, bog_mu;
  /* sat_vapor_pres_nml do_simple (sat_vapor_pres.F90; 1 in every Frierson / MiMA test case): 1 = the Clausius-Clapeyron form
   * es0*610.78*exp(-hlv/rvgas*(1/T - 1/tfreeze)), 0 = compute_es_k (Goff-Gratch / Smithsonian tables, ice below freezing,
   * sat_vapor_pres_k.F90:331-381) with finite-difference derivative tables */
  int sat_vapor_pres_do_simple;
  /* diffusivity_nml free_atm_diff = .true. (diffusivity_free, diffusivity.F90:604-697; the axisymmetric test case): Richardson-number
   * mixing-length diffusivities above the boundary layer */
  int free_atm_skyhi_diff, ampns;
  double rich_crit_diff, mix_len, rich_prandtl, ampns_max;
} IscaPhysicsConfig;

int isca_b200_physics_default_config(IscaPhysicsConfig* cfg);
int isca_b200_physics_create(const IscaPhysicsConfig* cfg, IscaPhysics* out);
int isca_b200_physics_destroy(IscaPhysics p);
const char* isca_b200_physics_last_error(IscaPhysics p);   /* p may be NULL */

/* sat_vapor_pres_init_k (sat_vapor_pres_k.F90:161-266): the handle's TABLE | DTABLE | D2TABLE (n = 5231 values each: -173 C ... 350 C in
 * 0.1 K steps) for cfg's tfreeze / hlv / rvgas / es0 / sat_vapor_pres_do_simple.  Host computation: needs no GPU and no handle. */
int isca_b200_sat_vapor_pres_tables(const IscaPhysicsConfig* cfg, int n, double* tables);

/* lookup_es_des (sat_vapor_pres_k.F90:1132-1158): n temperatures -> es, des.  Out-of-table temperatures fail. */
int isca_b200_lookup_es_des(IscaPhysics p, int n, const double* temp, double* es, double* des);

/* compute_qs (sat_vapor_pres_k.F90:457-540; q absent): qs and dqs/dT at (temp, press). */
int isca_b200_compute_qs(IscaPhysics p, int n, const double* temp, const double* press, double* qs, double* dqsdT);

/* lscale_cond (lscale_cond.F90:79-208): tin, qin, pfull [K][J][I]; phalf [K+1][J][I];
 * out rain [J][I], tdel, qdel [K][J][I]. */
int isca_b200_lscale_cond(IscaPhysics p, const double* tin, const double* qin, const double* pfull,
                          const double* phalf, double* rain, double* tdel, double* qdel);

/* two_stream_gray_rad_down(is, js, Time, lat, lon, p_half, t, net_surf_sw_down, surf_lw_down, albedo, q)
 * (two_stream_gray_rad.F90:386-655, no seasonal cycle):
 * lat [J][I] radians, p_half [K+1][J][I], t [K][J][I], albedo [J][I], q [K][J][I] specific humidity (read by the byrne and
 * geen schemes only; may be NULL otherwise); out net_surf_sw_down, surf_lw_down [J][I]. */
int isca_b200_two_stream_gray_rad_down(IscaPhysics p, const double* lat, const double* p_half, const double* t,
                                       const double* albedo, const double* q, double* net_surf_sw_down, double* surf_lw_down);

/* two_stream_gray_rad_up (two_stream_gray_rad.F90:659-776): same lat/p_half/t/q as the down call (the down sweep is
 * recomputed in registers rather than stored in module arrays), t_surf, albedo [J][I]; tdt [K][J][I] is incremented; olr [J][I]
 * (may be NULL). */
int isca_b200_two_stream_gray_rad_up(IscaPhysics p, const double* lat, const double* p_half, const double* t,
                                     const double* t_surf, const double* albedo, const double* q, double* tdt, double* olr);

/* do_read_co2 (two_stream_gray_rad.F90:209-212, 519-521): the CO2 concentration (ppmv) the host read from `co2_file` for Time_diag
 * (`carbon_conc = maxval(co2f)`), used by the longwave of the following two_stream_gray_rad_down / _up calls; as in the reference
 * the shortwave of the geen scheme (:466, evaluated before the file is read) still sees the value of the previous down call. */
int isca_b200_two_stream_gray_rad_set_co2(IscaPhysics p, double carbon_conc);

/* do_seasonal (two_stream_gray_rad.F90:417-447): the insolation [J][I] (= solar_constant * coszen from astronomy_mod
 * diurnal_solar, see isca_b200_diurnal_solar) that the following two_stream_gray_rad_down / _up calls use instead of the analytic
 * annual-mean profile; NULL switches back.  It takes precedence over the scheme's own profile, as in the reference. */
int isca_b200_two_stream_gray_rad_set_insolation(IscaPhysics p, const double* insolation);

/* damping_driver, rayleigh sponge (damping_driver.f90:404-420, 594-636): p_full, u, v [K][J][I],
 * pref [K+1] reference pressures; udt, vdt, tdt [K][J][I] are the damping tendencies (overwritten). */
int isca_b200_rayleigh_damping(IscaPhysics p, double delt, const double* p_full, const double* u, const double* v,
                               const double* pref, double* udt, double* vdt, double* tdt);

/* gcm_vert_diff_down (vert_diff.F90:270-402), sphum the only diffused tracer, no kbot.  All 3-D arrays [K][J][I]
 * (p_half [K+1][J][I]); tau_u, tau_v [J][I] are updated in place; dt_u, dt_v, dt_t are updated in place; dt_q is
 * read only; dissipative_heat is written.  The tridiagonal factors (vert_diff_mod e_global, f_t_global, f_q_global) and
 * the surf_diff_type fields stay on the device inside the handle until gcm_vert_diff_up. */
int isca_b200_gcm_vert_diff_down(IscaPhysics p, double delt, const double* u, const double* v, const double* t,
                                 const double* q, const double* diff_m, const double* diff_t, const double* p_half,
                                 const double* p_full, const double* z_full, double* tau_u, double* tau_v,
                                 const double* dtau_du, const double* dtau_dv, double* dt_u, double* dt_v,
                                 double* dt_t, const double* dt_q, double* dissipative_heat);

/* surf_diff_type / module-state inspection (tests): id 0 delta_t, 1 dflux_t, 2 delta_tr(sphum), 3 dflux_tr(sphum),
 * 4 dtmass, 5 delta_u, 6 delta_v ([J][I]); 16 e_global, 17 f_t_global, 18 f_q_global ([K][J][I]). */
int isca_b200_get_tri_surf(IscaPhysics p, int id, double* host);

/* mixed_layer_init subset: per-column heat capacity (land_sea_heat_capacity, J/m2/K) and ocean_qflux [J][I]. */
int isca_b200_mixed_layer_init(IscaPhysics p, const double* heat_capacity, const double* ocean_qflux);

/* mixed_layer (mixed_layer.F90:568-745; do_calc_eff_heat_cap path, no prescribed SST / ice / flux anomalies):
 * implicit slab update of t_surf [J][I] (in place) and of the handle's Tri_surf delta_t, delta_tr(sphum).
 * delta_t_surf (may be NULL) receives the increment. */
/* mixed_layer_nml do_sc_sst = .true. (mixed_layer.F90:495-502, 681-691): sst [J][I] = the SST the host's interpolator_mod read from
 * sst_file for the time stepped to (Time_next).  While set, mixed_layer moves t_surf to it (delta_t_surf = sst - t_surf) instead of
 * stepping the slab; NULL switches back to the slab ocean.  specify_sst_over_ocean_only and do_ape_sst are not built. */
int isca_b200_mixed_layer_set_sst(IscaPhysics p, const double* sst);
int isca_b200_mixed_layer(IscaPhysics p, double dt, double* t_surf, const double* flux_t, const double* flux_q,
                          const double* flux_r, const double* net_surf_sw_down, const double* surf_lw_down,
                          const double* dhdt_surf, const double* dedt_surf, const double* dedq_surf,
                          const double* drdt_surf, const double* dhdt_atm, const double* dedq_atm,
                          double* delta_t_surf);

/* gcm_vert_diff_up (vert_diff.F90:406-467): back-substitution -> dt_t, dt_q [K][J][I] (overwritten). */
int isca_b200_gcm_vert_diff_up(IscaPhysics p, double delt, double* dt_t, double* dt_q);

/* Monin-Obukhov similarity kernels on n independent points (monin_obukhov_kernel.F90): mo_drag = monin_obukhov_drag_1d
 * (:122-241, with the Newton iteration monin_obukhov_solve_zeta :245-411), mo_profile = monin_obukhov_profile_1d (:498-640),
 * stable_mix = monin_obukhov_stable_mix (:810-868), mo_diff = monin_obukhov_diff (:35-118; z [nk][n], k_m, k_h [nk][n]). */
int isca_b200_mo_drag(IscaPhysics p, int n, const double* pt, const double* pt0, const double* z, const double* z0,
                      const double* zt, const double* zq, const double* speed, double* drag_m, double* drag_t,
                      double* drag_q, double* u_star, double* b_star);
int isca_b200_mo_profile(IscaPhysics p, int n, double zref, double zref_t, const double* z, const double* z0,
                         const double* zt, const double* zq, const double* u_star, const double* b_star,
                         double* del_m, double* del_t, double* del_q);
int isca_b200_stable_mix(IscaPhysics p, int n, const double* rich, double* mix);
int isca_b200_mo_diff(IscaPhysics p, int n, int nk, const double* z, const double* u_star, const double* b_star,
                      double* k_m, double* k_h);

/* surface_flux (surface_flux.F90:338-700; bucket = .false., every point available).  All arrays [J][I]. */
typedef struct IscaSurfaceFluxArgs {
  /* in */
  const double *t_atm, *q_atm, *u_atm, *v_atm, *p_atm, *z_atm, *p_surf, *t_surf, *t_ca, *u_surf, *v_surf,
               *rough_mom, *rough_heat, *rough_moist, *rough_scale, *gust;
  const int* land;                 /* 1 = land */
  /* inout */
  double* q_surf;
  /* out */
  double *flux_t, *flux_q, *flux_r, *flux_u, *flux_v, *cd_m, *cd_t, *cd_q, *w_atm, *u_star, *b_star, *q_star,
         *dhdt_surf, *dedt_surf, *dedq_surf, *drdt_surf, *dhdt_atm, *dedq_atm, *dtaudu_atm, *dtaudv_atm,
         *ex_del_m, *ex_del_h, *ex_del_q, *temp_2m, *u_10m, *v_10m, *q_2m, *rh_2m;
} IscaSurfaceFluxArgs;
int isca_b200_surface_flux(IscaPhysics p, const IscaSurfaceFluxArgs* args);

/* diffusivity (diffusivity.F90:263-354: pbl_depth :358-456, diffusivity_pbl :458-526, diffusivity_entr :732-750), no kbot,
 * no ind_lcl.  t, q, u, v, p_full, z_full [K][J][I]; p_half, z_half [K+1][J][I]; u_star, b_star [J][I];
 * out h [J][I]; k_m, k_t [K][J][I] in/out (the incoming values are added, as the reference does). */
int isca_b200_diffusivity(IscaPhysics p, const double* t, const double* q, const double* u, const double* v,
                          const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                          const double* u_star, const double* b_star, double* h, double* k_m, double* k_t);

/* qe_moist_convection (qe_moist_convection.F90:157-186 -> SBM_convection_scheme :189-369), argument order of the reference
 * without coldT.  Tin, qin, p_full [K][J][I], p_half [K+1][J][I]; out rain, snow, CAPE, CIN, invtau_* [J][I] (snow = 0),
 * deltaT, deltaq, qref, Tref [K][J][I] (increments over dt, not rates), int convflag, kLZBs, kLCLs [J][I] (levels 1-based,
 * 0 = none).  As in the reference only the last column of invtau_q_relaxation / invtau_t_relaxation is non-zero. */
int isca_b200_qe_moist_convection(IscaPhysics p, double dt, const double* Tin, const double* qin, const double* p_full,
                                  const double* p_half, double* rain, double* snow, double* deltaT, double* deltaq,
                                  double* qref, int* convflag, int* kLZBs, double* cape, double* cin,
                                  double* invtau_q_relaxation, double* invtau_t_relaxation, double* Tref, int* kLCLs);

/* betts_miller_nml (atmos_param/betts_miller/betts_miller.f90:56-70), the full Betts-Miller scheme of convection_scheme =
 * 'FULL_BETTS_MILLER'.  do_taucape is rejected (the reference rescales the module's tau_bm inside the grid loop, :237-240, so its
 * result depends on the order of the columns); capetaubm / tau_min belong to it. */
typedef struct IscaBettsMillerConfig {
  int abi_version;                 /* 1 */
  int do_simp, do_shallower, do_changeqref, do_envsat, do_taucape;
  double tau_bm, rhbm, capetaubm, tau_min, buoyancy_kick;
} IscaBettsMillerConfig;
int isca_b200_betts_miller_default_config(IscaBettsMillerConfig* cfg);
/* betts_miller_init: the namelist of the handle's betts_miller calls (defaults until called) */
int isca_b200_betts_miller_init(IscaPhysics p, const IscaBettsMillerConfig* cfg);
/* betts_miller(dt, tin, qin, pfull, phalf, coldT, rain, snow, tdel, qdel, q_ref, bmflag, klzbs, cape, cin, t_ref, invtau_bm_t,
 * invtau_bm_q, capeflag, klcls) (betts_miller.f90:86-438 with capecalcnew :444-776 and lcltabl :779-845), without coldT / mask /
 * conv.  tin, qin, pfull [K][J][I], phalf [K+1][J][I]; out rain, snow, cape, cin, invtau_bm_t, invtau_bm_q, capeflag [J][I] (snow = 0;
 * capeflag = 0: the reference never assigns it; may be NULL), tdel, qdel, q_ref, t_ref [K][J][I] (increments over dt), int bmflag
 * (0 no CAPE, 1 shallow, 2 deep), klzbs, klcls [J][I] (levels 1-based, 0 = none). */
int isca_b200_betts_miller(IscaPhysics p, double dt, const double* tin, const double* qin, const double* pfull, const double* phalf,
                           double* rain, double* snow, double* tdel, double* qdel, double* q_ref, int* bmflag, int* klzbs, double* cape,
                           double* cin, double* t_ref, double* invtau_bm_t, double* invtau_bm_q, double* capeflag, int* klcls);

/* dry_convection(Time, tg, p_full, p_half, dt_tg, cape, cin, lzb, lcl) (atmos_param/dry_convection/dry_convection.f90:105-186 with
 * capecalc :190-299), the Schneider & Walker dry convective adjustment of convection_scheme = 'dry'.  tau, gamma: dry_convection_nml
 * (no defaults in the reference).  tg, p_full [K][J][I], p_half [K+1][J][I]; out dt_tg [K][J][I] (K/s), cape, cin [J][I], int lzb, lcl
 * [J][I] (1-based levels).  The reference's FATALs ("LCL defined, LZB not defined", "LCL above LZB") are returned as errors. */
int isca_b200_dry_convection(IscaPhysics p, double tau, double gamma, const double* tg, const double* p_full, const double* p_half,
                             double* dt_tg, double* cape, double* cin, int* lzb, int* lcl);

/* device-resident timing of one kernel (which: 0 lscale_cond, 1 gray_rad_down, 2 gray_rad_up, 3 rayleigh,
 * 4 gcm_vert_diff_down, 5 gcm_vert_diff_up, 9 betts_miller; any other value times gcm_vert_diff_up) on
 * synthetic resident columns; returns average ms per launch (CUDA events) and the algorithmic bytes per launch. */
int isca_b200_physics_time(IscaPhysics p, int which, int reps, double* ms, double* bytes);

/* ---------------------------------------------------------------------------------------------------------------
 * idealized_moist_model: atmosphere_mod boundary with idealized_moist_phys as the physics (atmosphere.F90:263-266, 300-302;
 * idealized_moist_phys.F90:322-731, 819-1395) for the grey-radiation slab-ocean aquaplanet (Frierson test case):
 * qe_moist_convection | none -> lscale_cond -> two_stream_gray_rad_down -> surface_flux -> two_stream_gray_rad_up ->
 * [damping_driver rayleigh] -> vert_turb_driver (do_diffusivity) -> gcm_vert_diff_down -> mixed_layer -> gcm_vert_diff_up,
 * then spectral_dynamics with the sphum grid tracer.  State, surface fields and all tendencies stay on the device.
 * Not built: bucket hydrology, land, clouds, RRTMG/Socrates, the other convection schemes (create fails loudly). */
#include "isca_b200.h"
typedef struct IscaMoist_t* IscaMoist;
typedef struct IscaMoistConfig {
  int abi_version;                 /* 2 */
  int convection_scheme;           /* 0 'NONE', 1 'SIMPLE_BETTS_MILLER', 2 'DRY', 3 'FULL_BETTS_MILLER' (idealized_moist_phys.F90:391-426;
                                    * betts_miller_nml: isca_b200_moist_set_betts_miller; 'DRY' needs
                                    * isca_b200_moist_set_dry_convection; large-scale condensation is then skipped, :977) */
  int do_damping;                  /* damping_driver rayleigh sponge */
  double roughness_mom, roughness_heat, roughness_moist;      /* idealized_moist_phys_nml :136-138 */
  double mixed_layer_depth, albedo_value, rho_cp;             /* mixed_layer_nml depth, albedo_value; constants RHO_CP */
  double constant_gust;            /* vert_turb_driver_nml, gust_scheme = 'constant' */
  int use_tau;                     /* vert_turb_driver_nml (vert_turb_driver.F90:109, 202-214): 1 (default) = diffusivity from the
                                    * `current` fields; 0 = from previous + delta_t * tendencies, as every shipped test case sets */
} IscaMoistConfig;

int isca_b200_moist_default_config(IscaMoistConfig* cfg);
/* dyn: the dynamical-core namelist (num_tracers must be 1 = sphum); phys: scheme namelists (its grid sizes are overwritten) */
int isca_b200_moist_create(const IscaConfig* dyn, const IscaPhysicsConfig* phys, const IscaMoistConfig* mc, IscaMoist* out);
/* One process per GPU: this rank's latitude block of the same model (columns are independent; the dynamical core and the
 * sphum tracer exchange over NCCL / peer memory as isca_b200_create describes).  Host arrays of the get/set calls are
 * (lon, lat_local). */
int isca_b200_moist_create_ranked(const IscaConfig* dyn, const IscaPhysicsConfig* phys, const IscaMoistConfig* mc, int rank, int nranks,
                                  const void* nccl_unique_id, IscaMoist* out);
int isca_b200_moist_destroy(IscaMoist m);
const char* isca_b200_moist_last_error(IscaMoist m);            /* m may be NULL */
/* the dynamical core inside (isca_b200_cold_start / set_grid_state / get_field ... operate on it) */
IscaHandle isca_b200_moist_dycore(IscaMoist m);
/* idealized_moist_phys_init after the atmospheric state is set: t_surf = tg(lowest level, current) + 1 (:643), q_surf = 0,
 * gust = 1, heat capacity = depth * RHO_CP, albedo = albedo_value */
int isca_b200_moist_init(IscaMoist m);
/* n calls of atmosphere(Time) */
int isca_b200_moist_step(IscaMoist m, int n_steps);
/* 2-D fields [J][I] of the last step: 0 t_surf, 1 precip (kg/m2/s), 2 flux_t, 3 flux_q, 4 z_pbl, 5 net_surf_sw_down,
 * 6 surf_lw_down, 7 convective rain, 8 cape, 9 convflag, 10 q_surf, 11 u_star, 12 b_star, 13 flux_u, 14 flux_v, 15 delta_t_surf;
 * 3-D [K][J][I]: 32 dt_ug, 33 dt_vg, 34 dt_tg, 35 dt_tracers(sphum) (physics tendencies), 36 diff_m, 37 diff_t */
int isca_b200_moist_get(IscaMoist m, int id, double* host);
/* atmosphere(Time) together with the host traffic of the step, software-pipelined: o3_host (may be NULL) = the ozone field of
 * isca_b200_moist_set_ozone, copied to the device before the step; then n_out fields are copied to host_out[i] (page-locked memory for
 * the copies to overlap): kinds[i] = 0 a field of isca_b200_get_field (ids[i], levels[i]), 1 a 2-D / 3-D field of isca_b200_moist_get
 * (ids[i] in 0-8, 17, 18, 34, 35; levels[i] ignored).  The call RETURNS BEFORE the downloads have finished: they overlap the next step
 * (what send_data of the reference's diag_manager consumes one step later); host_out is complete after isca_b200_moist_io_sync, which
 * also reports the FATAL conditions of the steps since the last sync.  isca_b200_moist_io_wait(m, age) waits only for the downloads of the
 * last call (age 0) or of the call before it (age 1) and leaves the pipeline running: with two alternating sets of host arrays, the set
 * of call n-1 is consumed after call n was issued. */
int isca_b200_moist_step_io(IscaMoist m, const double* o3_host, int n_out, const int* kinds, const int* ids, const int* levels,
                            double* const* host_out);
int isca_b200_moist_io_sync(IscaMoist m);
int isca_b200_moist_io_wait(IscaMoist m, int age);
/* Measurement aid: n_steps eager steps with a CUDA event after every kernel group; writes the average milliseconds per group to
   ms_out and the ';'-separated group names to `names` ("phys_*" = the column-physics kernels in call order, "phys_rrtmg_call" = a
   whole radiation call on the steps where the alarm fires, the rest = the dynamical core's groups of isca_b200_profile_step).
   Returns the number of groups, -1 on error. */
int isca_b200_moist_profile_step(IscaMoist m, int n_steps, double* ms_out, int max_groups, char* names, int capacity);
int isca_b200_moist_set_t_surf(IscaMoist m, const double* host);
/* do_sc_sst for the moist model: the prescribed SST [J][I] (this rank's latitude block) of the next isca_b200_moist_step calls
 * (see isca_b200_mixed_layer_set_sst); NULL = slab ocean again.  The shim calls it whenever interpolator_mod delivers a new field. */
int isca_b200_moist_set_sst(IscaMoist m, const double* host);
/* Surface properties [J][I] that idealized_moist_phys_init / mixed_layer_init derive from the land options (land mask file,
 * land_h_capacity_prefactor, land_albedo_prefactor, land_roughness_prefactor; idealized_moist_phys.F90:565-616, mixed_layer.F90:380-470):
 * id 20 albedo, 21 rough_mom, 22 rough_heat, 23 rough_moist, 24 surface heat capacity (J/m2/K), 25 land mask (0. / 1.; used by
 * surface_flux for land_humidity_prefactor / land_evap_prefactor).  Call after isca_b200_moist_init (which fills the aquaplanet values). */
int isca_b200_moist_set_surface(IscaMoist m, int id, const double* host);
/* do_read_co2 for the resident moist model (see isca_b200_two_stream_gray_rad_set_co2): call before the isca_b200_moist_step whose
 * Time the value belongs to */
int isca_b200_moist_set_co2(IscaMoist m, double carbon_conc);
/* betts_miller_nml of convection_scheme = 'FULL_BETTS_MILLER' (defaults until called) */
int isca_b200_moist_set_betts_miller(IscaMoist m, const IscaBettsMillerConfig* cfg);
/* dry_convection_nml: relaxation time scale tau [s] and lapse-rate factor gamma of convection_scheme = 'DRY' */
int isca_b200_moist_set_dry_convection(IscaMoist m, double tau, double gamma);
/* mixed_layer_init: ocean_qflux [J][I] (W/m2; `do_qflux` / `do_warmpool` of mixed_layer_nml, atmos_param/qflux/qflux.f90, or a
 * q-flux file), added to the slab's heat budget every step.  Call after isca_b200_moist_init (which zeroes it). */
int isca_b200_moist_set_ocean_qflux(IscaMoist m, const double* host);
/* average ms per step of the last isca_b200_moist_step call (CUDA events), and of its physics part */
int isca_b200_moist_timing(IscaMoist m, double* ms_step, double* ms_physics);

#ifdef __cplusplus
}
#endif
#endif
