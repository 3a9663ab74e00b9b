// rrtm_internal.h -- what moist_model.cu needs from rrtm.cu (device-pointer entry points on a caller-owned stream)
#pragma once
#include "../../include/isca_b200_rrtm.h"
#include <cuda_runtime.h>
#include <vector>

// run_rrtmg on device arrays in the model layout ([K][J][I], Pa); see isca_b200_run_rrtmg
int isca_rrtm_run_device(IscaRrtm r, cudaStream_t st, const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                         const double* t, const double* q, const double* o3, const double* t_surf, const double* albedo, const double* coszen,
                         double* tdt, double* tdt_rad, double* flux_sw, double* flux_lw, double* olr, double* toa_sw);
// the zenith-angle block of run_rrtmg (rrtm_radiation.F90:700-745) for model time `total_seconds`, on device lat / lon [n]
int isca_rrtm_coszen_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb_angle, cudaStream_t st, double total_seconds,
                            int n, const double* lat, const double* lon, double* coszen, double* fracday);
// the do_seasonal block of two_stream_gray_rad_down (two_stream_gray_rad.F90:417-447) for Time = (days, seconds): coszen on device
// lat / lon [n].  dc supplies solday (>= 0: perpetual day), equinox_day, do_rad_time_avg (= use_time_average_coszen), dt_rad_avg
// (seconds, > 0), the astronomy_nml values and the calendar lengths
int isca_gray_coszen_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb_angle, cudaStream_t st, double days,
                            double seconds, int n, const double* lat, const double* lon, double* coszen);
// diurnal_solar_2d (astronomy.f90:1123-1410) on device lat / lon [n]; dt <= 0: instantaneous.  dc: the astronomy_nml values
int isca_diurnal_solar_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb_angle, cudaStream_t st, double gmt,
                              double time_since_ae, double dt, int n, const double* lat, const double* lon, double* coszen);
// astronomy_mod orbit table (astronomy.f90:orbit)
std::vector<double> isca_rrtm_orbit(const IscaRrtmDriverConfig& dc);
