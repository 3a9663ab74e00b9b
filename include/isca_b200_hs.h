/* isca_b200_hs.h -- C ABI of hs_forcing_mod with the namelist options beyond the Held-Suarez default
 * (atmos_param/hs_forcing/hs_forcing.F90).
 *
 * The default Held-Suarez forcing is fused into the dynamical core's grid kernel (isca_b200_step, include/isca_b200.h).  This header
 * adds the general module: (i) isca_b200_hs_forcing = subroutine hs_forcing (:148-272) on host arrays, one call per physics step,
 * with equilibrium_t_option = 'Held_Suarez' | 'EXOPLANET' | 'EXOPLANET2' | 'top_down', stratosphere_t_option, local_heating_option =
 * 'Isidoro' and the tracer source/sink; (ii) isca_b200_hs_model_* = atmosphere_mod of the dry model (atmosphere.F90:276-352) with that
 * forcing: state, forcing and dynamics resident on the device.  The options that read netCDF files through interpolator_mod
 * (equilibrium_t_option / local_heating_option = 'from_file', relax_to_specified_wind) are rejected at create time.
 *
 * Arrays are host pointers in the reference's Fortran memory order (lon fastest, then lat, then level; level 1 = model top),
 * double precision.  All functions return 0 on success; isca_b200_hs_last_error() describes a failure.
 */
#ifndef ISCA_B200_HS_H
#define ISCA_B200_HS_H
#include "isca_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct IscaHsForcing_t* IscaHsForcing;
typedef struct IscaHsModel_t* IscaHsModel;

enum { ISCA_HS_HELD_SUAREZ = 0, ISCA_HS_EXOPLANET = 1, ISCA_HS_EXOPLANET2 = 2, ISCA_HS_TOP_DOWN = 3 };       /* equilibrium_t_option */
enum { ISCA_HS_EXTEND_TP = 0, ISCA_HS_C_ABOVE_TP = 1, ISCA_HS_HS_LIKE = 2, ISCA_HS_STRAT_OTHER = 3 };        /* stratosphere_t_option */

/* hs_forcing_nml (hs_forcing.F90:74-122) and the constants_mod / astronomy_nml values the module reads */
typedef struct IscaHsForcingConfig {
  int abi_version;                 /* 1 */
  int num_lon, num_lat, num_levels;
  int no_forcing, do_conserve_energy;
  int equilibrium_t_option, stratosphere_t_option;
  int local_heating_option;        /* 0 '' (none), 1 'Isidoro' */
  int num_angles;                  /* astronomy_nml (orbit table of the EXOPLANET options) */
  double t_zero, t_strat, delh, delv, eps, sigma_b, P00, p_trop, alpha, ka, ks, kf, trflux, trsink;
  double local_heating_srfamp, local_heating_xwidth, local_heating_ywidth, local_heating_xcenter, local_heating_ycenter,
         local_heating_vert_decay;
  double peri_time, smaxis, albedo, lapse, h_a, tau_s, heat_capacity, ml_depth, spinup_time;
  /* constants_mod: KAPPA, RDGAS (CP_AIR = RDGAS/KAPPA), GRAV, STEFAN, solar_const, OMEGA, orbital_period (the value update_orbit
   * multiplies by 86400, :822-828), orbital_rate (2 pi / orbital_period at constants_init; <= 0: computed so) */
  double kappa, rdgas, grav, stefan, solar_const, omega, orbital_period, orbital_rate;
  double ecc, obliq, per;          /* astronomy_nml */
} IscaHsForcingConfig;

int isca_b200_hs_forcing_default_config(IscaHsForcingConfig* cfg);
const char* isca_b200_hs_last_error(void);

/* hs_forcing_init (hs_forcing.F90:276-470).  lat [J][I] radians and Time = (days, seconds) are used by the top_down spin-up of the
 * surface heat capacity (:331-366; no INPUT/hs_forcing.res.nc -- hand a restart over with isca_b200_hs_forcing_set_tg_prev);
 * lat may be NULL for the other options. */
int isca_b200_hs_forcing_create(const IscaHsForcingConfig* cfg, const double* lat, long long days, int seconds, IscaHsForcing* out);
int isca_b200_hs_forcing_destroy(IscaHsForcing h);

/* hs_forcing(is, ie, js, je, dt, Time, lon, lat, p_half, p_full, u, v, t, r, um, vm, tm, rm, udt, vdt, tdt, rdt, zfull)
 * (hs_forcing.F90:148-272; no mask / kbot).  lon, lat [J][I] radians; p_half [K+1][J][I]; p_full, u, v, t, um, vm, tm, zfull, udt, vdt,
 * tdt [K][J][I]; r, rm, rdt [num_tracers][K][J][I] (r is not read, as in the reference; all three may be NULL when num_tracers = 0).
 * udt, vdt, tdt, rdt are incremented.  zfull is read by top_down only (NULL otherwise).  teq [K][J][I] and h_trop [J][I] (the
 * module's diagnostics) may be NULL.  With top_down the handle's tg_prev is advanced by the call. */
int isca_b200_hs_forcing(IscaHsForcing h, double dt, long long days, int seconds, const double* lon, const double* lat,
                         const double* p_half, const double* p_full, const double* u, const double* v, const double* t,
                         const double* r, const double* um, const double* vm, const double* tm, const double* rm,
                         double* udt, double* vdt, double* tdt, double* rdt, const double* zfull, int num_tracers,
                         double* teq, double* h_trop);

/* tg_prev [J][I] of the top_down option (RESTART/hs_forcing.res, hs_forcing.F90:337, 498) */
int isca_b200_hs_forcing_get_tg_prev(IscaHsForcing h, double* tg_prev);
int isca_b200_hs_forcing_set_tg_prev(IscaHsForcing h, const double* tg_prev);

/* ---- the dry model with the general forcing: atmosphere_init / atmosphere / atmosphere_end (atmosphere.F90:120-352) ---------
 * dyn: the dynamical-core configuration (its hs_forcing_nml members are ignored; num_tracers 0 or 1); hs: num_lon / num_lat /
 * num_levels, kappa, rdgas, grav are overwritten with the core's.  Single rank. */
int isca_b200_hs_model_create(const IscaConfig* dyn, const IscaHsForcingConfig* hs, IscaHsModel* out);
int isca_b200_hs_model_destroy(IscaHsModel m);
/* the dynamical core of the model (state set / get, tables): owned by the model */
IscaHandle isca_b200_hs_model_dycore(IscaHsModel m);
/* model time of the next step (Time of atmosphere(Time)); call before isca_b200_hs_model_init */
int isca_b200_hs_model_set_time(IscaHsModel m, long long days, int seconds);
/* restart of the top_down option: tg_prev [J][I] read from INPUT/hs_forcing.res.nc (hs_forcing.F90:337); call before
 * isca_b200_hs_model_init, which then skips the spin-up.  isca_b200_hs_model_get id 2 returns the value to write at the end. */
int isca_b200_hs_model_set_tg_prev(IscaHsModel m, const double* tg_prev);
/* hs_forcing_init on the model grid; call after the initial state is in place */
int isca_b200_hs_model_init(IscaHsModel m);
/* n_steps calls of atmosphere(Time): hs_forcing(Time + Time_step) on (previous) fields and p / z of (current), spectral_dynamics */
int isca_b200_hs_model_step(IscaHsModel m, int n_steps);
/* diagnostics of the last step: id 0 teq [K][J][I], 1 h_trop [J][I], 2 tg_prev [J][I], 3 tdt (forcing temperature tendency) [K][J][I] */
int isca_b200_hs_model_get(IscaHsModel m, int id, double* host);

#ifdef __cplusplus
}
#endif
#endif
