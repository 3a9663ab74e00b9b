// hs_forcing_column.h -- the per-column arithmetic of hs_forcing_mod (atmos_param/hs_forcing/hs_forcing.F90) as
// `__host__ __device__` functions: hs_forcing.cu calls them from one thread per column, tests/host/hs_host.cpp (test
// infrastructure) calls the same functions in a serial loop so that the formulas can be checked on a machine without a GPU.
#pragma once
#include <cmath>
#include <cstddef>

#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

namespace isca_hs {

struct HsParams {
  int K; size_t plane;
  int do_conserve_energy, eq_opt, strat_opt, local_heating;
  double tka, tks, vkf, sigma_b, t_zero, t_strat, delh, delv, eps, P00, p_trop, alpha, kappa, cp_air;
  double xwidth, ywidth, xcenter, ycenter, srfamp, vert_decay;                 // radians / K s-1 (hs_forcing_init :370-381)
  double lapse, h_a, tau_s, stefan, solar_const, albedo, ml_heat;              // ml_heat = ml_depth * heat_capacity
  double trflux, rdamp;                                                        // tracer_source_sink: rdamp = 1/sink time (or 0)
};

// s, t_radbal, t_trop, h_trop, t_surf of the top-down forcing (:347-357, :946-958)
HS_HD void hs_radiative_surface(const HsParams& p, double lat, double dec, double& t_trop, double& h_trop, double& t_surf) {
  const double PI = 3.14159265358979323846;
  double inv = -tan(lat) * tan(dec);                                           // calc_hour_angle (:832-849)
  if (inv > 1) inv = 1;
  if (inv < -1) inv = -1;
  const double ha = acos(inv);
  const double s = p.solar_const / PI * (ha * sin(lat) * sin(dec) + cos(lat) * cos(dec) * sin(ha));
  const double t_radbal = pow((1 - p.albedo) * s / p.stefan, 0.25);
  t_trop = t_radbal / pow(2.0, 0.25);
  const double a = 1.3863 * t_trop;
  h_trop = 1.0 / (16 * p.lapse) * (a + sqrt(a * a + 32 * p.lapse * p.tau_s * p.h_a * t_trop));
  t_surf = t_trop + h_trop * p.lapse;
}

// one update of the slab temperature: tg = stefan*dt/(ml_depth*heat_capacity)*(t_surf**4 - tg_prev**4) + tg_prev (:359, :959)
HS_HD double hs_slab_update(const HsParams& p, double dt, double t_surf, double tg_prev) {
  const double ts2 = t_surf * t_surf, tp2 = tg_prev * tg_prev;
  return p.stefan * dt / p.ml_heat * (ts2 * ts2 - tp2 * tp2) + tg_prev;
}

// hs_forcing (:148-272) for one column: rayleigh_damping, the dissipative heating, newtonian_damping / top_down_newtonian_damping,
// local_heating.  Arrays are [k * plane + col]; coszen: diurnal_exoplanet (EXOPLANET); dec: update_orbit (top_down);
// tg_prev (top_down) is read and updated; teq_out / h_trop_out may be null.
HS_HD void hs_column(const HsParams& p, size_t col, double dt, double lat, double lon, double coszen, double dec,
                     const double* p_half, const double* p_full, const double* u, const double* v, const double* t,
                     const double* um, const double* vm, const double* zfull, double* udt, double* vdt, double* tdt,
                     double* teq_out, double* tg_prev, double* h_trop_out) {
  const int K = p.K; const size_t pl = p.plane;
  const double twopi = 2 * 3.14159265358979323846;
  const double ps = p_half[(size_t)K * pl + col];
  const double rps = 1. / ps;
  const double sin_lat = sin(lat), cos_lat = cos(lat);
  const double sin_lat_2 = sin_lat * sin_lat, cos_lat_2 = 1.0 - sin_lat_2, cos_lat_4 = cos_lat_2 * cos_lat_2;
  const double vcoeff = -p.vkf / (1.0 - p.sigma_b), tcoeff = (p.tks - p.tka) / (1.0 - p.sigma_b);
  double t_star = p.t_zero - p.delh * sin_lat_2 - p.eps * sin_lat;
  const double tstr = p.t_strat - p.eps * sin_lat;
  if (p.eq_opt == 1) t_star = p.t_zero - p.delh * (1 - coszen) - p.eps * sin_lat;
  double t_trop = 0.0, h_trop = 0.0;
  if (p.eq_opt == 3) {
    double t_surf;
    hs_radiative_surface(p, lat, dec, t_trop, h_trop, t_surf);
    const double tg = hs_slab_update(p, dt, t_surf, tg_prev[col]);
    tg_prev[col] = tg;
    t_trop = tg - h_trop * p.lapse;
    if (h_trop_out) h_trop_out[col] = h_trop;
  }
  double lon_lat_factor = 0.0;
  if (p.local_heating) {
    const double lon_temp = lon - twopi * floor(lon / twopi);
    const double a = (lon_temp - p.xcenter) / p.xwidth, b = (lat - p.ycenter) / p.ywidth;
    lon_lat_factor = p.srfamp * exp(-.5 * (a * a)) * exp(-.5 * (b * b));
  }
  for (int k = 0; k < K; ++k) {
    const size_t e = (size_t)k * pl + col;
    const double pf = p_full[e];
    const double sigma = pf * rps;
    const bool in_bl = (sigma <= 1.0 && sigma > p.sigma_b);
    double utnd = 0.0, vtnd = 0.0;
    if (in_bl) { const double vfactr = vcoeff * (sigma - p.sigma_b); utnd = vfactr * u[e]; vtnd = vfactr * v[e]; }
    double dT = tdt[e];
    if (p.do_conserve_energy) dT = dT + (-((um[e] + .5 * utnd * dt) * utnd + (vm[e] + .5 * vtnd * dt) * vtnd) / p.cp_air);
    udt[e] = udt[e] + utnd; vdt[e] = vdt[e] + vtnd;
    double teq;
    if (p.eq_opt == 0 || p.eq_opt == 1) {
      const double p_norm = pf / p.P00;
      const double the = t_star - p.delv * (p.eq_opt == 0 ? cos_lat_2 : coszen) * log(p_norm);
      teq = fmax(the * pow(p_norm, p.kappa), tstr);
    } else if (p.eq_opt == 2) {
      teq = fmax(p.t_strat * cos_lat * pow(pf / p.p_trop, p.alpha), p.t_strat);
    } else {
      const double zkm = zfull[e] / 1000;
      teq = t_trop + p.lapse * (h_trop - zkm);
      if (p.strat_opt == 1) { if (zkm >= h_trop) teq = tstr; }
      else if (p.strat_opt == 2) teq = fmax(teq, tstr);
      else if (p.strat_opt == 0) { if (zkm >= h_trop) teq = t_trop; }
      else teq = fmax(teq, 0.);
    }
    double tdamp = p.tka;
    if (in_bl) tdamp = p.tka + cos_lat_4 * (tcoeff * (sigma - p.sigma_b));
    dT = dT + (-tdamp * (t[e] - teq));
    if (p.local_heating) dT = dT + lon_lat_factor * exp((pf - ps) / p.vert_decay);
    tdt[e] = dT;
    if (teq_out) teq_out[e] = teq;
  }
}

// the tracer part of hs_forcing (:248-265) for one column of one tracer: rst = rm + dt*rdt; rdt += source - rdamp*rst
HS_HD void hs_tracer_column(const HsParams& p, size_t col, double dt, const double* p_half, const double* rm, double* rdt) {
  const int K = p.K; const size_t pl = p.plane;
  for (int k = 0; k < K; ++k) {
    const size_t e = (size_t)k * pl + col;
    const double rst = rm[e] + dt * rdt[e];
    double source = 0.0;
    if (k == K - 1) source = p.trflux / (p_half[e + pl] - p_half[e]);
    rdt[e] = rdt[e] + (source - p.rdamp * rst);
  }
}

}  // namespace isca_hs
