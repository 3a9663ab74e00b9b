/* isca_b200_rrtm.h -- C ABI of the RRTMG clear-sky radiation.
 *
 * Replaces, for the configuration Isca runs (`icld = 0`, `iaer = 0`, `idrv = 0`: clear sky, no aerosol;
 * atmos_param/rrtm_radiation/rrtm_radiation.F90:113-116):
 *   rrtmg_lw   rrtmg_lw/gcm_model/src/rrtmg_lw_rad.nomcica.f90:81   (inatm, setcoef, taumol, rtrnmr)
 *   rrtmg_sw   rrtmg_sw/gcm_model/src/rrtmg_sw_rad.nomcica.f90:73   (inatm_sw, setcoef_sw, spcvrt_sw: taumol_sw, reftra_sw, vrtqdr_sw)
 *   interp_temp + the column part of run_rrtmg   rrtm_radiation.F90:502-544, 816-1000
 * The g-point reduction of rrtmg_lw_ini / rrtmg_sw_ini is done when the coefficient file is built
 * (tools/make_rrtmg_tables.py -> isca_b200/data/rrtmg_tables.bin); create() loads that file.
 *
 * Host arrays, double precision.  The rrtmg_lw / rrtmg_sw entry points take the reference's own layout: (ncol, nlay)
 * Fortran order (column index fastest), layer 1 = lowest layer, pressures in hPa, gases as volume mixing ratios.
 * All functions return 0 on success; isca_b200_rrtm_last_error() describes a failure.  There is no CPU fallback.
 */
#ifndef ISCA_B200_RRTM_H
#define ISCA_B200_RRTM_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct IscaRrtm_t* IscaRrtm;

typedef struct IscaRrtmConfig {
  int abi_version;                  /* 1 */
  int num_lon, num_lat, num_levels; /* grid of run_rrtmg (this rank's latitude block); rrtmg_lw/sw take ncol per call */
  /* constants_mod: CP_AIR (rrtmg_*_ini(cp_air), idealized_moist_phys.F90:781-782), RDGAS, GAS_CONSTANT, WTMH2O, WTMOZONE */
  double cp_air, rdgas, gas_constant, wtmh2o, wtmozone;
  /* rrtm_radiation_nml (rrtm_radiation.F90:117-226) */
  double co2ppmv, h2o_lower_limit, temp_lower_limit, temp_upper_limit, solrad, solr_cnst;
  int include_secondary_gases;
  double ch4_val, n2o_val, o2_val, cfc11_val, cfc12_val, cfc22_val, ccl4_val;
  int convert_sphum_to_vmr, input_o3_file_is_mmr;
  int lonstep;                      /* run_rrtmg: radiation on every lonstep-th longitude, linear interpolation in between */
} IscaRrtmConfig;

int isca_b200_rrtm_default_config(IscaRrtmConfig* cfg);
int isca_b200_rrtm_create(const IscaRrtmConfig* cfg, const char* table_path, IscaRrtm* out);
int isca_b200_rrtm_destroy(IscaRrtm r);
const char* isca_b200_rrtm_last_error(IscaRrtm r);          /* r may be NULL */

/* rrtmg_lw(ncol, nlay, icld=0, idrv=0, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
 *          cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, ... -> uflx, dflx, hr):
 * play, tlay and the gases (ncol, nlay); plev, tlev (ncol, nlay+1); tsfc (ncol); emis (ncol, 16);
 * out uflx, dflx (ncol, nlay+1) W/m2; hr (ncol, nlay) K/day.  Clear-sky and total-sky outputs coincide. */
int isca_b200_rrtmg_lw(IscaRrtm r, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                       const double* tlev, const double* tsfc, const double* h2ovmr, const double* o3vmr,
                       const double* co2vmr, const double* ch4vmr, const double* n2ovmr, const double* o2vmr,
                       const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr, const double* ccl4vmr,
                       const double* emis, double* uflx, double* dflx, double* hr);

/* rrtmg_sw(ncol, nlay, icld=0, iaer=0, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
 *          asdir=asdif=aldir=aldif=albedo, coszen, adjes, dyofyr=0, scon, ... -> swuflx, swdflx, swhr):
 * albedo, coszen (ncol); columns with coszen < 1e-10 return zeros as in the reference. */
int isca_b200_rrtmg_sw(IscaRrtm r, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                       const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                       const double* n2ovmr, const double* o2vmr, const double* albedo, const double* coszen,
                       double adjes, double scon, double* swuflx, double* swdflx, double* swhr);

/* interp_temp + run_rrtmg on a radiation step, model layout: 3-D arrays [K][J][I] (level 1 = model top), p_half /
 * z_half [K+1][J][I], 2-D [J][I]; pressures Pa.  q = specific humidity, o3 = ozone as read from the ozone file (NULL =
 * no ozone; mass mixing ratio if input_o3_file_is_mmr), coszen = the zenith angle run_rrtmg computed (astronomy stays
 * on the host side).  tdt [K][J][I] is incremented by the radiative heating (K/s); tdt_rad (may be NULL) receives it
 * (the `store_intermediate_rad` copy); flux_sw = net surface SW down, flux_lw = surface LW down; olr, toa_sw may be NULL. */
int isca_b200_run_rrtmg(IscaRrtm r, const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                        const double* t, const double* q, const double* o3, const double* t_surf, const double* albedo,
                        const double* coszen, double* tdt, double* tdt_rad, double* flux_sw, double* flux_lw, double* olr,
                        double* toa_sw);

/* average ms per launch (CUDA events) of the LW (which = 0) / SW (1) kernel on the columns of the last call */
int isca_b200_rrtm_time(IscaRrtm r, int which, int reps, double* ms);

/* ---------------------------------------------------------------------------------------------------------------
 * Radiation time stepping and zenith angle of run_rrtmg (rrtm_radiation.F90:640-745) + astronomy_nml
 * (shared/astronomy/astronomy.f90:141-164), and RRTMG as the radiation of the idealized moist model
 * (idealized_moist_phys.F90:1167-1177: `do_rrtm_radiation` replaces `two_stream_gray`). */
typedef struct IscaRrtmDriverConfig {
  int abi_version;                 /* 1 */
  int dt_rad, dt_rad_avg;          /* seconds; dt_rad <= 0: every step; dt_rad_avg <= 0: dt_rad */
  int do_rad_time_avg, store_intermediate_rad, solday, frierson_solar_rad;
  double equinox_day, del_sol, del_sw;
  double ecc, obliq, per;          /* astronomy_nml */
  int num_angles;
  double day_in_s, year_in_s;      /* length_of_day(), length_of_year() of the calendar (360-day year in the test cases) */
} IscaRrtmDriverConfig;

int isca_b200_rrtm_driver_default_config(IscaRrtmDriverConfig* dc);

/* diurnal_solar(lat, lon, gmt, time_since_ae, cosz, fracday, rrsun [, dt]) (astronomy.f90:1123-1410), n points, host arrays;
 * dt <= 0: no time averaging.  rrsun may be NULL. */
int isca_b200_diurnal_solar(IscaRrtm r, const IscaRrtmDriverConfig* dc, int n, const double* lat, const double* lon, double gmt,
                            double time_since_ae, double dt, double* cosz, double* fracday, double* rrsun);

#include "isca_b200_physics.h"
/* Switch the moist model's radiation from two_stream_gray_rad to RRTMG (rc->num_lon/num_lat/num_levels are overwritten with the
 * model's).  Must be called before isca_b200_moist_init. */
int isca_b200_moist_use_rrtm(IscaMoist m, const IscaRrtmConfig* rc, const IscaRrtmDriverConfig* dc, const char* table_path);
/* two_stream_gray_rad_nml do_seasonal = .true. (two_stream_gray_rad.F90:417-447) for the moist model's grey radiation: every step
 * insolation = solar_constant * coszen(Time) from astronomy_mod diurnal_solar instead of the analytic annual-mean profile.
 * dc: solday (>= 0: perpetual day of the year; < 0, the namelist default -10: follow the model clock -- NOTE that
 * isca_b200_rrtm_driver_default_config sets the rrtm_radiation_nml default solday = 0, which on THIS path is a perpetual day 0:
 * a caller who wants the seasonal cycle must set solday = -10 after taking the defaults), equinox_day,
 * do_rad_time_avg (= use_time_average_coszen), dt_rad_avg (seconds; <= 0: dt_atmos), ecc / obliq / per / num_angles,
 * day_in_s / year_in_s; the other fields are ignored.  Must be called before isca_b200_moist_init; excludes isca_b200_moist_use_rrtm. */
int isca_b200_moist_set_seasonal(IscaMoist m, const IscaRrtmDriverConfig* dc);
/* ozone as read from the ozone file, [K][J][I] on this rank's latitude block (do_read_ozone; NULL = no ozone) */
int isca_b200_moist_set_ozone(IscaMoist m, const double* o3);
/* model time of the next isca_b200_moist_step (Time of atmosphere(Time)); advanced by dt_atmos every step */
int isca_b200_moist_set_time(IscaMoist m, long long days, int seconds);

#ifdef __cplusplus
}
#endif
#endif
