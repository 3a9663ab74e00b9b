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
  int abi_version;                /* 1 */
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
} IscaPhysicsConfig;

int isca_b200_physics_default_config(IscaPhysicsConfig* cfg);
int isca_b200_physics_create(const IscaPhysicsConfig* cfg, IscaPhysics* out);
int isca_b200_physics_destroy(IscaPhysics p);
const char* isca_b200_physics_last_error(IscaPhysics p);   /* p may be NULL */

/* lookup_es_des (sat_vapor_pres_k.F90:1132-1158): n temperatures -> es, des.  Out-of-table temperatures fail. */
int isca_b200_lookup_es_des(IscaPhysics p, int n, const double* temp, double* es, double* des);

/* compute_qs (sat_vapor_pres_k.F90:457-540; q absent): qs and dqs/dT at (temp, press). */
int isca_b200_compute_qs(IscaPhysics p, int n, const double* temp, const double* press, double* qs, double* dqsdT);

/* lscale_cond (lscale_cond.F90:79-208): tin, qin, pfull [K][J][I]; phalf [K+1][J][I];
 * out rain [J][I], tdel, qdel [K][J][I]. */
int isca_b200_lscale_cond(IscaPhysics p, const double* tin, const double* qin, const double* pfull,
                          const double* phalf, double* rain, double* tdel, double* qdel);

/* two_stream_gray_rad_down (two_stream_gray_rad.F90:386-655, frierson, no seasonal cycle):
 * lat [J][I] radians, p_half [K+1][J][I], t [K][J][I], albedo [J][I];
 * out net_surf_sw_down, surf_lw_down [J][I]. */
int isca_b200_two_stream_gray_rad_down(IscaPhysics p, const double* lat, const double* p_half, const double* t,
                                       const double* albedo, double* net_surf_sw_down, double* surf_lw_down);

/* two_stream_gray_rad_up (two_stream_gray_rad.F90:659-776): same lat/p_half/t as the down call (the down sweep is
 * recomputed in registers rather than stored), t_surf, albedo [J][I]; tdt [K][J][I] is incremented; olr [J][I]
 * (may be NULL). */
int isca_b200_two_stream_gray_rad_up(IscaPhysics p, const double* lat, const double* p_half, const double* t,
                                     const double* t_surf, const double* albedo, double* tdt, double* olr);

/* damping_driver, rayleigh sponge (damping_driver.f90:404-420, 594-636): p_full, u, v [K][J][I],
 * pref [K+1] reference pressures; udt, vdt, tdt [K][J][I] are the damping tendencies (overwritten). */
int isca_b200_rayleigh_damping(IscaPhysics p, double delt, const double* p_full, const double* u, const double* v,
                               const double* pref, double* udt, double* vdt, double* tdt);

/* device-resident timing of one kernel (which: 0 lscale_cond, 1 gray_rad_down, 2 gray_rad_up, 3 rayleigh) on
 * synthetic resident columns; returns average ms per launch (CUDA events) and the algorithmic bytes per launch. */
int isca_b200_physics_time(IscaPhysics p, int which, int reps, double* ms, double* bytes);

#ifdef __cplusplus
}
#endif
#endif
