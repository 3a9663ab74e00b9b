// grid.cu -- grid-space column kernels (one thread per (lon, lat) column, k-recurrences in registers).
//
//   grid_step_kernel   hs_forcing (atmos_param/hs_forcing/hs_forcing.F90:148-272, 508-679) +
//                      initialize_corrections (model/spectral_dynamics.F90:1306-1338) +
//                      pressure_variables (model/press_and_geopot.F90:152-221) +
//                      compute_pressure_gradient (:1192-1209), four_in_one (:1038-1112),
//                      compute_geopotential (press_and_geopot.F90:314-359),
//                      vert_advection second_centered (atmos_shared/vert_advection/vert_advection.F90:163-173,440-476),
//                      horizontal_advection tail (tools/transforms.F90:826), Coriolis/KE block (:893-902)
//   colsum / reduce    area_weighted_global_mean (tools/transforms.F90:1059-1077),
//                      mass_weighted_global_integral (model/global_integral.F90:49-81)
//   apply_* kernels    compute_corrections (model/spectral_dynamics.F90:1213-1302)
//   press_heights      compute_pressures_and_heights (press_and_geopot.F90:363-387)
#include "device.h"
#include "grid.h"

namespace isca {

// pressure_variables for one column, one level at a time, bottom-up.
struct PressLevel { double p_half_k, p_half_k1, ln_half_k, ln_half_k1, ln_full, p_full; };

// ln(p_half(k)): for a pure sigma coordinate (pk == 0 everywhere) p_half = bk*ps, so the logarithm is
// ln(bk) (host table) + ln(ps) (one log per column) instead of one log per level; the two forms differ by
// at most ~2 ulp of ln p (~4e-15 absolute), far inside the parity tolerance.
__device__ __forceinline__ double ln_p_half(const DevTables& t, const Params& pr, int k, double p_half_k, double ln_ps) {
  return pr.pure_sigma ? (t.ln_bk[k] + ln_ps) : log(p_half_k);
}

__device__ __forceinline__ void press_level(const DevTables& t, const Params& pr, int k, double ps, double ln_ps,
                                            double ln_half_k1, PressLevel& o) {
  o.p_half_k = t.pk[k] + t.bk[k] * ps;
  o.p_half_k1 = t.pk[k + 1] + t.bk[k + 1] * ps;
  o.ln_half_k1 = ln_half_k1;
  if (k == 0 && pr.pkbk0_zero) {
    o.ln_half_k = 0.0;
    o.ln_full = ln_half_k1 + (-1.0);                 // ln_top_level_factor (press_and_geopot.F90:103,186)
  } else {
    o.ln_half_k = ln_p_half(t, pr, k, o.p_half_k, ln_ps);
    const double alpha = 1.0 - o.p_half_k * (o.ln_half_k1 - o.ln_half_k) / (o.p_half_k1 - o.p_half_k);
    o.ln_full = o.ln_half_k1 - alpha;
  }
  o.p_full = exp(o.ln_full);
}

// Column kernel, k-split: one warp = 32 consecutive longitudes x one chunk of CH levels; KW = 4 warps
// (chunks) per column group.  The vertical recurrences (cumulative mass divergence top-down,
// hydrostatic integral bottom-up, energy integral) are done per chunk in registers and stitched
// through shared memory, which quadruples the available parallelism of a 131 072-column grid and
// shortens every dependent chain by 4.
constexpr int GS_KW = 4;

template <int CH>
__global__ void __launch_bounds__(32 * GS_KW, (CH <= 10) ? 6 : 4)
grid_step_kernel(DevTables t, Params pr, GridStepArgs a) {
  const GeomDev& g = t.g;
  const int I = g.I, K = g.K;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int jl = blockIdx.y;
  const int j = g.j0 + jl;
  const bool live = (i < I);
  const size_t col = (size_t)jl * I + (live ? i : 0);
  const size_t plane = (size_t)g.Jloc * I;
  __shared__ double s_tot[GS_KW][32], s_gh[GS_KW][32], s_en[GS_KW][32];
  // per-thread chunk arrays live in shared memory ([c][thread], conflict-free) to keep registers for occupancy
  __shared__ double s_cum[CH + 1][32 * GS_KW], s_phi[CH][32 * GS_KW];
  const int tx = threadIdx.x;

  const int k_lo = w * CH;
  const int k_hi = (k_lo + CH < K) ? (k_lo + CH) : K;          // chunk = [k_lo, k_hi), may be empty

  const double ps_c = a.ps_cur[col], ps_p = a.ps_prev[col];
  const double cosm = t.cosm_lat[j];
  // the energy fixer's temperature increment is applied on read (see apply_fixers_kernel)
  const double sh_c = a.scal[SC_TSHIFT0 + a.slot_cur], sh_p = a.scal[SC_TSHIFT0 + a.slot_prev];
  // compute_pressure_gradient: dx_psg = psg * S2G(dx ln ps), then divide_by_cos
  const double dx_psg = (ps_c * a.dx_lnps[col]) * cosm;
  const double dy_psg = (ps_c * a.dy_lnps[col]) * cosm;

  // ---- pass 1 (top-down within the chunk): cumulative mass divergence (four_in_one :1073-1083)
  {
    double run = 0.0;
    s_cum[0][tx] = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = k_lo + c;
      double dmean = 0.0;
      if (k < k_hi) {
        const size_t e = (size_t)k * plane + col;
        const double dp = t.dpk[k] + t.dbk[k] * ps_c;
        dmean = a.div_cur[e] * dp + t.dbk[k] * (a.u_cur[e] * dx_psg + a.v_cur[e] * dy_psg);
      }
      run = run + dmean;
      s_cum[c + 1][tx] = run;
    }
    s_tot[w][lane] = run;
  }
  __syncthreads();
  double cum_off = 0.0, dmean_total = 0.0;
#pragma unroll
  for (int q = 0; q < GS_KW; ++q) {
    if (q == w) cum_off = dmean_total;
    dmean_total = dmean_total + s_tot[q][lane];
  }
  if (w == 0 && live) a.dt_lnps[col] = (0.0 - dmean_total) / ps_c;   // dt_psg = dt_psg - dmean_tot ; dt_ln_psg = dt_psg/psg

  // Held-Suarez latitude constants (newtonian_damping :527-545)
  const double lat = t.rad_lat[j];
  const double sin_lat = sin(lat);
  const double sin_lat_2 = sin_lat * sin_lat;
  const double cos_lat_2 = 1.0 - sin_lat_2;
  const double cos_lat_4 = cos_lat_2 * cos_lat_2;
  const double t_star = pr.t_zero - pr.delh * sin_lat_2 - pr.eps * sin_lat;
  const double tstr = pr.t_strat - pr.eps * sin_lat;
  const double tcoeff = (pr.tks - pr.tka) / (1.0 - pr.sigma_b);
  const double vcoeff = -pr.vkf / (1.0 - pr.sigma_b);
  const double ps_hs = t.pk[K] + t.bk[K] * ps_c;     // ps = p_half(:,:,size(p_half,3))
  const double rps = 1. / ps_hs;
  const double fcor = t.coriolis[j];
  const double delta_t = pr.delta_t;
  const double ln_p00 = log(pr.P00);

  // ---- pass 2 (bottom-up within the chunk)
  double gh_local = 0.0;                              // geopot_half relative to the chunk's bottom interface
  double energy_int = 0.0;
  double ln_half_below = 0.0;
  const double ln_ps = pr.pure_sigma ? log(ps_c) : 0.0;
  const double inv_cp = 1.0 / pr.cp_air;
  double u_dn = 0.0, v_dn = 0.0, T_dn = 0.0, u_k = 0.0, v_k = 0.0, T_k = 0.0;
  if (k_hi > k_lo) {
    ln_half_below = ln_p_half(t, pr, k_hi, t.pk[k_hi] + t.bk[k_hi] * ps_c, ln_ps);
    const size_t e = (size_t)(k_hi - 1) * plane + col;
    u_k = a.u_cur[e]; v_k = a.v_cur[e]; T_k = a.t_cur[e] + sh_c;
    if (k_hi < K) { u_dn = a.u_cur[e + plane]; v_dn = a.v_cur[e + plane]; T_dn = a.t_cur[e + plane] + sh_c; }
  }
#pragma unroll
  for (int c = CH - 1; c >= 0; --c) {
    const int k = k_lo + c;
    if (k >= k_hi) continue;
    const size_t e = (size_t)k * plane + col;
    double u_up = 0.0, v_up = 0.0, T_up = 0.0;       // level k-1
    if (k > 0) { u_up = a.u_cur[e - plane]; v_up = a.v_cur[e - plane]; T_up = a.t_cur[e - plane] + sh_c; }
    PressLevel pl;
    press_level(t, pr, k, ps_c, ln_ps, ln_half_below, pl);

    // ---------------- physics: hs_forcing on (u,v,T)(previous), pressures of `current`
    double dt_u = 0.0, dt_v = 0.0, dt_T = 0.0;
    const double u_p = a.u_prev[e], v_p = a.v_prev[e], T_p = a.t_prev[e] + sh_p;
    if (pr.physics_on && !pr.no_forcing) {
      const double sigma = pl.p_full * rps;
      double utnd = 0.0, vtnd = 0.0;
      const bool in_bl = (sigma <= 1.0 && sigma > pr.sigma_b);
      if (in_bl) { const double vfactr = vcoeff * (sigma - pr.sigma_b); utnd = vfactr * u_p; vtnd = vfactr * v_p; }
      if (pr.do_conserve_energy) {
        const double ttnd = -((u_p + .5 * utnd * delta_t) * utnd + (v_p + .5 * vtnd * delta_t) * vtnd) * inv_cp;
        dt_T = dt_T + ttnd;
      }
      dt_u = dt_u + utnd; dt_v = dt_v + vtnd;
      // p_norm = p_full/P00; log(p_norm) = ln_p_full - log(P00); p_norm**kappa = exp(kappa*log(p_norm))
      const double ln_pn = pl.ln_full - ln_p00;
      const double the = t_star - pr.delv * cos_lat_2 * ln_pn;
      double teq = the * exp(pr.kappa * ln_pn);
      teq = fmax(teq, tstr);
      double tdamp = pr.tka;
      if (in_bl) { const double tfactr = tcoeff * (sigma - pr.sigma_b); tdamp = pr.tka + cos_lat_4 * tfactr; }
      dt_T = dt_T + (-tdamp * (T_p - teq));
    } else if (a.dt_u_in) {                          // tendencies supplied by the caller (spectral_dynamics API)
      dt_u = a.dt_u_in[e]; dt_v = a.dt_v_in[e]; dt_T = a.dt_t_in[e];
    }
    // initialize_corrections: energy of (previous + dt*delta_t), mass-weighted with psg(previous)
    {
      const double up = u_p + dt_u * delta_t, vp = v_p + dt_v * delta_t;
      const double en = 0.5 * (up * up + vp * vp) + pr.cp_air * (T_p + dt_T * delta_t);
      const double dpp = (t.pk[k + 1] + t.bk[k + 1] * ps_p) - (t.pk[k] + t.bk[k] * ps_p);
      energy_int = energy_int + en * dpp;
    }

    // ---------------- four_in_one, level k
    const double Tv = T_k;                            // use_virtual_temperature = .false. (dry) -> virtual_t = tg
    const double dp = t.dpk[k] + t.dbk[k] * ps_c;
    const double dp_inv = 1 / dp;
    const double dlog_1 = pl.ln_half_k1 - pl.ln_full;
    const double dlog_2 = pl.ln_full - pl.ln_half_k;
    const double dlog_3 = pl.ln_half_k1 - pl.ln_half_k;
    const double x1 = (t.bk[k + 1] * dlog_1 + t.bk[k] * dlog_2) * dp_inv;
    const double x2 = x1 * dx_psg;
    const double x3 = x1 * dy_psg;
    dt_u = dt_u - pr.rdgas * Tv * x2;
    dt_v = dt_v - pr.rdgas * Tv * x3;
    const double cum_k = cum_off + s_cum[c][tx], cum_k1 = cum_off + s_cum[c + 1][tx];
    const double dmean = a.div_cur[e] * dp + t.dbk[k] * (u_k * dx_psg + v_k * dy_psg);
    const double x4 = (cum_k * dlog_3 + dmean * dlog_1) * dp_inv;
    const double x5 = x4 - u_k * x2 - v_k * x3;
    dt_T = dt_T - pr.kappa * Tv * x5;
    if (a.wg_full && live) a.wg_full[e] = -x5 * pl.p_full;
    // wg at the two interfaces of level k (:1102-1108)
    const double w_top = (k == 0) ? 0.0 : (-cum_k + dmean_total * t.bk[k]);
    const double w_bot = (k == K - 1) ? 0.0 : (-cum_k1 + dmean_total * t.bk[k + 1]);
    if (a.wg && live) { a.wg[e] = w_top; if (k == K - 1) a.wg[e + plane] = w_bot; }

    // ---------------- compute_geopotential (relative to the chunk's bottom interface)
    const double gfull = gh_local + pr.rdgas * T_k * (pl.ln_half_k1 - pl.ln_full);
    if (!(k == 0 && pr.pk0_zero)) gh_local = gh_local + pr.rdgas * T_k * (pl.ln_half_k1 - pl.ln_half_k);

    // ---------------- vert_advection, second_centered, advective form, time_level = current
    {
      const double rdz = 1.0 / (pl.p_half_k1 - pl.p_half_k);  // dp = p_half(k+1) - p_half(k); one reciprocal for the 3 fields
      const double fu_t = (k == 0) ? w_top * u_k : w_top * (0.5 * (u_k + u_up));
      const double fu_b = (k == K - 1) ? w_bot * u_k : w_bot * (0.5 * (u_dn + u_k));
      dt_u = dt_u + (-(fu_b - fu_t - u_k * (w_bot - w_top)) * rdz);
      const double fv_t = (k == 0) ? w_top * v_k : w_top * (0.5 * (v_k + v_up));
      const double fv_b = (k == K - 1) ? w_bot * v_k : w_bot * (0.5 * (v_dn + v_k));
      dt_v = dt_v + (-(fv_b - fv_t - v_k * (w_bot - w_top)) * rdz);
      const double ft_t = (k == 0) ? w_top * T_k : w_top * (0.5 * (T_k + T_up));
      const double ft_b = (k == K - 1) ? w_bot * T_k : w_bot * (0.5 * (T_dn + T_k));
      dt_T = dt_T + (-(ft_b - ft_t - T_k * (w_bot - w_top)) * rdz);
    }
    // ---------------- horizontal_advection of T (dx, dy already divided by cos in the FFT epilogue)
    dt_T = dt_T - u_k * a.dx_t[e] - v_k * a.dy_t[e];
    // ---------------- Coriolis / vorticity
    const double absv = a.vor_cur[e] + fcor;
    dt_u = dt_u + absv * v_k;
    dt_v = dt_v - absv * u_k;

    // ---------------- outputs for the forward transforms
    if (live) {
      a.out_A[e] = dt_u * cosm;                        // vor_div_from_uv_grid: divide_by_cos(u_grid)
      a.out_B[e] = dt_v * cosm;
      a.out_T[e] = dt_T;
    }
    s_phi[c][tx] = gfull + .5 * (u_k * u_k + v_k * v_k);

    ln_half_below = pl.ln_half_k;
    u_dn = u_k; v_dn = v_k; T_dn = T_k;
    u_k = u_up; v_k = v_up; T_k = T_up;
  }
  s_gh[w][lane] = gh_local;
  s_en[w][lane] = energy_int;
  __syncthreads();
  // geopotential at the chunk's bottom interface: surf_geopotential + contributions of the chunks below
  double gh_off = a.phis[col];
#pragma unroll
  for (int q = GS_KW - 1; q >= 0; --q) if (q > w) gh_off = gh_off + s_gh[q][lane];
  if (live) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = k_lo + c;
      if (k < k_hi) a.out_phi[(size_t)k * plane + col] = gh_off + s_phi[c][tx];
    }
    if (w == 0) {
      double en = 0.0;
#pragma unroll
      for (int q = GS_KW - 1; q >= 0; --q) en = en + s_en[q][lane];
      const double wt = t.wts_lat[j];
      a.part[0 * plane + col] = wt * ps_p;             // area_weighted_global_mean(psg(previous))
      a.part[1 * plane + col] = wt * en;               // mass_weighted_global_integral(energy, psg(previous))
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pure sigma coordinate (pk == 0 at every interface, bk(K+1) == 1): p_half = bk*ps, so every logarithm,
// exponential and quotient of pressure_variables / four_in_one / hs_forcing factors into a host-computed
// per-level constant times a per-column scalar (ps, 1/ps, ln ps, (ps/P00)^kappa):
//   ln p_full(k) = LF(k) + ln ps,  p_full(k) = PF(k)*ps,  dlog_1 = AL(k),  dlog_2 = D2(k),  dlog_3 = D3(k),
//   dp = DB(k)*ps,  sigma = PF(k),  p_norm^kappa = PFK(k)*(ps/P00)^kappa.
// The column loop then contains no transcendental and no division.  Mathematically identical to the
// generic kernel; the arithmetic is re-associated, so results agree to a few ulp, not bit for bit
// (the parity tests hold both kernels to the same tolerance).  Mass fluxes are carried divided by ps.
// ---------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(32 * GS_KW, (CH <= 10) ? 6 : 4)
grid_step_sigma_kernel(DevTables t, Params pr, GridStepArgs a) {
  const GeomDev& g = t.g;
  const int I = g.I, K = g.K;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int jl = blockIdx.y;
  const int j = g.j0 + jl;
  const bool live = (i < I);
  const size_t col = (size_t)jl * I + (live ? i : 0);
  const size_t plane = (size_t)g.Jloc * I;
  __shared__ double s_tot[GS_KW][32], s_gh[GS_KW][32], s_en[GS_KW][32];
  __shared__ double s_cum[CH + 1][32 * GS_KW], s_phi[CH][32 * GS_KW];
  const int tx = threadIdx.x;
  const SigmaTables& sg = t.sig;

  const int k_lo = w * CH;
  const int k_hi = (k_lo + CH < K) ? (k_lo + CH) : K;

  const double ps_c = a.ps_cur[col], ps_p = a.ps_prev[col];
  const double cosm = t.cosm_lat[j];
  const double sh_c = a.scal[SC_TSHIFT0 + a.slot_cur], sh_p = a.scal[SC_TSHIFT0 + a.slot_prev];
  const double gx = a.dx_lnps[col] * cosm, gy = a.dy_lnps[col] * cosm;      // (1/ps) * grad(ps) / cos
  const double ln_ps = log(ps_c);

  // ---- pass 1: cumulative (mass divergence)/ps, top-down within the chunk
  {
    double run = 0.0;
    s_cum[0][tx] = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = k_lo + c;
      double dm = 0.0;
      if (k < k_hi) {
        const size_t e = (size_t)k * plane + col;
        dm = sg.db[k] * (a.div_cur[e] + (a.u_cur[e] * gx + a.v_cur[e] * gy));
      }
      run = run + dm;
      s_cum[c + 1][tx] = run;
    }
    s_tot[w][lane] = run;
  }
  __syncthreads();
  double cum_off = 0.0, tot = 0.0;
#pragma unroll
  for (int q = 0; q < GS_KW; ++q) {
    if (q == w) cum_off = tot;
    tot = tot + s_tot[q][lane];
  }
  if (w == 0 && live) a.dt_lnps[col] = 0.0 - tot;                           // dt_ln_psg = -dmean_tot/psg

  // Held-Suarez latitude / column factors
  const double lat = t.rad_lat[j];
  const double sin_lat = sin(lat);
  const double sin_lat_2 = sin_lat * sin_lat;
  const double cos_lat_2 = 1.0 - sin_lat_2;
  const double cos_lat_4 = cos_lat_2 * cos_lat_2;
  const double t_star = pr.t_zero - pr.delh * sin_lat_2 - pr.eps * sin_lat;
  const double tstr = pr.t_strat - pr.eps * sin_lat;
  const double tcoeff = (pr.tks - pr.tka) / (1.0 - pr.sigma_b);
  const double vcoeff = -pr.vkf / (1.0 - pr.sigma_b);
  const double fcor = t.coriolis[j];
  const double delta_t = pr.delta_t;
  const double inv_cp = 1.0 / pr.cp_air;
  const double ln_pn0 = ln_ps - log(pr.P00);                                 // ln(ps/P00)
  const double pk_col = exp(pr.kappa * ln_pn0);                              // (ps/P00)^kappa
  const bool hs_on = (pr.physics_on && !pr.no_forcing);

  // ---- pass 2: bottom-up within the chunk
  double gh_local = 0.0, energy_int = 0.0;
  double u_dn = 0.0, v_dn = 0.0, T_dn = 0.0, u_k = 0.0, v_k = 0.0, T_k = 0.0;
  if (k_hi > k_lo) {
    const size_t e = (size_t)(k_hi - 1) * plane + col;
    u_k = a.u_cur[e]; v_k = a.v_cur[e]; T_k = a.t_cur[e] + sh_c;
    if (k_hi < K) { u_dn = a.u_cur[e + plane]; v_dn = a.v_cur[e + plane]; T_dn = a.t_cur[e + plane] + sh_c; }
  }
#pragma unroll
  for (int c = CH - 1; c >= 0; --c) {
    const int k = k_lo + c;
    if (k >= k_hi) continue;
    const size_t e = (size_t)k * plane + col;
    double u_up = 0.0, v_up = 0.0, T_up = 0.0;
    if (k > 0) { u_up = a.u_cur[e - plane]; v_up = a.v_cur[e - plane]; T_up = a.t_cur[e - plane] + sh_c; }
    const double dbk = sg.db[k], rdb = sg.rdb[k], al = sg.al[k], d3 = sg.d3[k], pf = sg.pf[k];

    // ---------------- hs_forcing
    double dt_u = 0.0, dt_v = 0.0, dt_T = 0.0;
    const double u_p = a.u_prev[e], v_p = a.v_prev[e], T_p = a.t_prev[e] + sh_p;
    if (hs_on) {
      const double sigma = pf;                                                // p_full/ps
      double utnd = 0.0, vtnd = 0.0;
      const bool in_bl = (sigma <= 1.0 && sigma > pr.sigma_b);
      if (in_bl) { const double vfactr = vcoeff * (sigma - pr.sigma_b); utnd = vfactr * u_p; vtnd = vfactr * v_p; }
      if (pr.do_conserve_energy) dt_T = dt_T + (-((u_p + .5 * utnd * delta_t) * utnd + (v_p + .5 * vtnd * delta_t) * vtnd) * inv_cp);
      dt_u = dt_u + utnd; dt_v = dt_v + vtnd;
      const double ln_pn = sg.lf[k] + ln_pn0;                                 // ln(p_full/P00)
      double teq = (t_star - pr.delv * cos_lat_2 * ln_pn) * (sg.pfk[k] * pk_col);
      teq = fmax(teq, tstr);
      double tdamp = pr.tka;
      if (in_bl) tdamp = pr.tka + cos_lat_4 * (tcoeff * (sigma - pr.sigma_b));
      dt_T = dt_T + (-tdamp * (T_p - teq));
    } else if (a.dt_u_in) {
      dt_u = a.dt_u_in[e]; dt_v = a.dt_v_in[e]; dt_T = a.dt_t_in[e];
    }
    {   // initialize_corrections energy integral, dp(previous) = db*ps(previous)
      const double up = u_p + dt_u * delta_t, vp = v_p + dt_v * delta_t;
      energy_int = energy_int + (0.5 * (up * up + vp * vp) + pr.cp_air * (T_p + dt_T * delta_t)) * (dbk * ps_p);
    }

    // ---------------- four_in_one
    const double x2 = sg.x1c[k] * gx, x3 = sg.x1c[k] * gy;                    // x1*dx_psg, x1*dy_psg
    dt_u = dt_u - pr.rdgas * T_k * x2;
    dt_v = dt_v - pr.rdgas * T_k * x3;
    const double cum_k = cum_off + s_cum[c][tx], cum_k1 = cum_off + s_cum[c + 1][tx];
    const double dm = dbk * (a.div_cur[e] + (u_k * gx + v_k * gy));
    const double x4 = (cum_k * d3 + dm * al) * rdb;
    const double x5 = x4 - u_k * x2 - v_k * x3;
    dt_T = dt_T - pr.kappa * T_k * x5;
    if (a.wg_full && live) a.wg_full[e] = -x5 * (pf * ps_c);
    const double w_top = (k == 0) ? 0.0 : (-cum_k + tot * sg.b[k]);           // wg/ps at the interfaces
    const double w_bot = (k == K - 1) ? 0.0 : (-cum_k1 + tot * sg.b[k + 1]);
    if (a.wg && live) { a.wg[e] = w_top * ps_c; if (k == K - 1) a.wg[e + plane] = w_bot * ps_c; }

    // ---------------- compute_geopotential
    const double gfull = gh_local + pr.rdgas * T_k * al;
    if (k != 0) gh_local = gh_local + pr.rdgas * T_k * d3;

    // ---------------- vert_advection (second_centered, advective form): (w/ps) / db
    {
      const double fu_t = (k == 0) ? w_top * u_k : w_top * (0.5 * (u_k + u_up));
      const double fu_b = (k == K - 1) ? w_bot * u_k : w_bot * (0.5 * (u_dn + u_k));
      dt_u = dt_u + (-(fu_b - fu_t - u_k * (w_bot - w_top)) * rdb);
      const double fv_t = (k == 0) ? w_top * v_k : w_top * (0.5 * (v_k + v_up));
      const double fv_b = (k == K - 1) ? w_bot * v_k : w_bot * (0.5 * (v_dn + v_k));
      dt_v = dt_v + (-(fv_b - fv_t - v_k * (w_bot - w_top)) * rdb);
      const double ft_t = (k == 0) ? w_top * T_k : w_top * (0.5 * (T_k + T_up));
      const double ft_b = (k == K - 1) ? w_bot * T_k : w_bot * (0.5 * (T_dn + T_k));
      dt_T = dt_T + (-(ft_b - ft_t - T_k * (w_bot - w_top)) * rdb);
    }
    dt_T = dt_T - u_k * a.dx_t[e] - v_k * a.dy_t[e];
    const double absv = a.vor_cur[e] + fcor;
    dt_u = dt_u + absv * v_k;
    dt_v = dt_v - absv * u_k;
    if (live) {
      a.out_A[e] = dt_u * cosm;
      a.out_B[e] = dt_v * cosm;
      a.out_T[e] = dt_T;
    }
    s_phi[c][tx] = gfull + .5 * (u_k * u_k + v_k * v_k);
    u_dn = u_k; v_dn = v_k; T_dn = T_k;
    u_k = u_up; v_k = v_up; T_k = T_up;
  }
  s_gh[w][lane] = gh_local;
  s_en[w][lane] = energy_int;
  __syncthreads();
  double gh_off = a.phis[col];
#pragma unroll
  for (int q = GS_KW - 1; q >= 0; --q) if (q > w) gh_off = gh_off + s_gh[q][lane];
  if (live) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = k_lo + c;
      if (k < k_hi) a.out_phi[(size_t)k * plane + col] = gh_off + s_phi[c][tx];
    }
    if (w == 0) {
      double en = 0.0;
#pragma unroll
      for (int q = GS_KW - 1; q >= 0; --q) en = en + s_en[q][lane];
      const double wt = t.wts_lat[j];
      a.part[0 * plane + col] = wt * ps_p;
      a.part[1 * plane + col] = wt * en;
    }
  }
}

void launch_grid_step(const DevTables& t, const Params& pr, const GridStepArgs& a, cudaStream_t st) {
  dim3 block(32 * GS_KW), grid((t.g.I + 31) / 32, t.g.Jloc);
  const int ch = (t.g.K + GS_KW - 1) / GS_KW;
  if (pr.sigma_fast) {
    if (ch <= 4) grid_step_sigma_kernel<4><<<grid, block, 0, st>>>(t, pr, a);
    else if (ch <= 7) grid_step_sigma_kernel<7><<<grid, block, 0, st>>>(t, pr, a);
    else if (ch <= 10) grid_step_sigma_kernel<10><<<grid, block, 0, st>>>(t, pr, a);
    else if (ch <= 15) grid_step_sigma_kernel<15><<<grid, block, 0, st>>>(t, pr, a);
    else grid_step_sigma_kernel<20><<<grid, block, 0, st>>>(t, pr, a);
    return;
  }
  if (ch <= 4) grid_step_kernel<4><<<grid, block, 0, st>>>(t, pr, a);
  else if (ch <= 7) grid_step_kernel<7><<<grid, block, 0, st>>>(t, pr, a);
  else if (ch <= 10) grid_step_kernel<10><<<grid, block, 0, st>>>(t, pr, a);
  else if (ch <= 15) grid_step_kernel<15><<<grid, block, 0, st>>>(t, pr, a);
  else grid_step_kernel<20><<<grid, block, 0, st>>>(t, pr, a);
}

// ---------------------------------------------------------------------------------------------
// deterministic reduction of nq per-column arrays: out[q] = sum / min / max of part[q][0..n)
// One CTA, fixed summation tree -> bitwise reproducible run to run.
// ---------------------------------------------------------------------------------------------
// Stage 1: RED_NB CTAs, each reduces a fixed contiguous slice in a fixed order; stage 2: one CTA
// combines the RED_NB partials.  Fixed slices + fixed trees -> bitwise reproducible run to run.
constexpr int RED_NB = 128;
__device__ __forceinline__ double red_op(int op, double x, double y) {
  return (op == 0) ? x + y : ((op == 1) ? fmin(x, y) : fmax(x, y));
}
__device__ __forceinline__ double red_init(int op) { return (op == 0) ? 0.0 : ((op == 1) ? 1.0e300 : -1.0e300); }

__global__ void __launch_bounds__(256)
reduce_stage1_kernel(const double* __restrict__ part, size_t n, const int* __restrict__ ops, double* __restrict__ tmp) {
  __shared__ double sh[256];
  const int q = blockIdx.y, op = ops[q];
  const size_t per = (n + RED_NB - 1) / RED_NB;
  const size_t lo = (size_t)blockIdx.x * per, hi = (lo + per < n) ? lo + per : n;
  double acc = red_init(op);
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) acc = red_op(op, acc, part[(size_t)q * n + i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = red_op(op, sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) tmp[q * RED_NB + blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(RED_NB)
reduce_stage2_kernel(const double* __restrict__ tmp, const int* __restrict__ ops, double* __restrict__ out) {
  __shared__ double sh[RED_NB];
  const int q = blockIdx.x, op = ops[q];
  sh[threadIdx.x] = tmp[q * RED_NB + threadIdx.x];
  __syncthreads();
  for (int s = RED_NB / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = red_op(op, sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[q] = sh[0];
}
void launch_reduce(const double* part, size_t n, int nq, const int* ops, double* out, double* tmp, cudaStream_t st) {
  reduce_stage1_kernel<<<dim3(RED_NB, nq), 256, 0, st>>>(part, n, ops, tmp);
  reduce_stage2_kernel<<<nq, RED_NB, 0, st>>>(tmp, ops, out);
}

// ---------------------------------------------------------------------------------------------
// post-transform: compute_corrections (spectral_dynamics.F90:1213-1302), mass + energy fixers in one pass.
// The energy integral is taken with the mass-corrected surface pressure ps' = f*ps; as dp = dpk + dbk*ps' it is
// A + f*B with A = sum e*dpk, B = sum e*dbk*ps, so both partial sums are formed before f is known and all global sums of the
// step travel in one reduction (and one all-reduce).  The water fixer's integrals of the future level (spectral_dynamics.F90:1245-1278)
// are taken with the same corrected ps: their column sums come from the tracer PPM sweep as (sum q*dpk, sum q*dbk) below / above
// water_correction_limit and enter the same way.  part[0] = w*ps, part[1] = w*A, part[2] = w*B, part[3..6] = w*(water A, B*ps below the
// limit; A, B*ps above), part[7] = -min T, part[8] = max T
// ---------------------------------------------------------------------------------------------
__global__ void colsum_fixers_kernel(DevTables t, Params pr, const double* __restrict__ u, const double* __restrict__ v,
                                     const double* __restrict__ T, const double* __restrict__ ps, const double* __restrict__ wpart,
                                     double* __restrict__ part) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double p_s = ps[col];
  double va = 0.0, vb = 0.0, tmin = 1.0e300, tmax = -1.0e300;
  for (int k = 0; k < g.K; ++k) {
    const size_t e = (size_t)k * plane + col;
    const double uu = u[e], vv = v[e], tt = T[e];
    const double en = 0.5 * (uu * uu + vv * vv) + pr.cp_air * tt;
    va = va + en * (t.pk[k + 1] - t.pk[k]);
    vb = vb + en * ((t.bk[k + 1] - t.bk[k]) * p_s);
    tmin = fmin(tmin, tt); tmax = fmax(tmax, tt);
  }
  const double w = t.wts_lat[g.j0 + jl];
  part[col] = w * p_s;
  part[plane + col] = w * va;
  part[2 * plane + col] = w * vb;
  double wa_c = 0.0, wb_c = 0.0, wa_n = 0.0, wb_n = 0.0;
  if (wpart) { wa_c = wpart[col]; wb_c = wpart[plane + col] * p_s; wa_n = wpart[2 * plane + col]; wb_n = wpart[3 * plane + col] * p_s; }
  part[3 * plane + col] = w * wa_c;
  part[4 * plane + col] = w * wb_c;
  part[5 * plane + col] = w * wa_n;
  part[6 * plane + col] = w * wb_n;
  part[7 * plane + col] = -tmin;
  part[8 * plane + col] = tmax;
}
void launch_colsum_fixers(const DevTables& t, const Params& pr, const double* u, const double* v, const double* T,
                          const double* ps, const double* wpart, double* part, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  colsum_fixers_kernel<<<grid, 128, 0, st>>>(t, pr, u, v, T, ps, wpart, part);
}

// mass_correction_factor = mean_ps_prev / mean_ps_tmp (:1228-1231); temperature_correction (:1236-1241).
// The temperature increment of the energy fixer (tg(future) = tg(future) + temperature_correction, :1239) is NOT applied with a
// pass over the 3-D field: it is stored as the pending shift of the future slot and added by every reader (grid_step;
// materialize_t for host mirrors), which yields the same rounded values fl(T + tc).  Only ps and the spectral (0,0)
// coefficients are updated here.
__global__ void apply_fixers_kernel(DevTables t, Params pr, int slot_fut, double* __restrict__ ps, double2* __restrict__ lnps_fut,
                                    double2* __restrict__ lnps_cur, double2* __restrict__ ts_fut, double2* __restrict__ ts_cur,
                                    double rc_raw, double* __restrict__ scal, double denom, int owns_m0, int do_mass, int do_energy) {
  const GeomDev& g = t.g;
  const size_t n = (size_t)g.Jloc * g.I;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const double mean_prev = scal[SC_SUM_PS_PREV] / denom;
  const double mean_tmp = scal[SC_SUM_PS_FUT] / denom;
  const double f = do_mass ? (mean_prev / mean_tmp) : 1.0;
  const double mean_e_prev = scal[SC_SUM_EN_PREV] / denom / pr.grav;
  const double mean_e_tmp = (scal[SC_SUM_EN_FUT] + f * scal[SC_SUM_EN_FUTB]) / denom / pr.grav;
  const double tc = do_energy ? pr.grav * (mean_e_prev - mean_e_tmp) / (pr.cp_air * mean_prev) : 0.0;
  const double ntmin = scal[SC_NTMIN], tmax = scal[SC_TMAX];
  if (idx < n && do_mass) ps[idx] = f * ps[idx];
  if (blockIdx.x == 0) {
    if (owns_m0 && do_energy) {
      for (int k = threadIdx.x; k < g.K; k += blockDim.x) {
        ts_fut[k].x = ts_fut[k].x + sqrt(2.) * tc;                               // ts(0,0,:,future) (:1241)
        if (ts_cur) ts_cur[k].x = ts_cur[k].x + rc_raw * (sqrt(2.) * tc);         // fused leapfrog_2level_B
      }
    }
    if (threadIdx.x == 0) {
      if (owns_m0 && do_mass) {
        const double inc = sqrt(2.) * log(f);
        lnps_fut[0].x = lnps_fut[0].x + inc;                                     // ln_ps(0,0,future) (:1231)
        if (lnps_cur) lnps_cur[0].x = lnps_cur[0].x + rc_raw * inc;               // fused leapfrog_2level_B sees the fixed value
      }
      scal[SC_MEAN_PS_PREV] = mean_prev;
      scal[SC_MASS_FACTOR] = f;
      scal[SC_TSHIFT0 + slot_fut] = tc;
      scal[SC_MEAN_EN_PREV] = mean_e_prev;
      scal[SC_T_CORR] = tc;
      scal[SC_TMIN] = -ntmin;
      scal[SC_W_CORR] = scal[SC_WA_CORR] + f * scal[SC_WB_CORR];                  // water integrals with the corrected ps
      scal[SC_W_NOT] = scal[SC_WA_NOT] + f * scal[SC_WB_NOT];
      scal[SC_W_ALL] = (scal[SC_WA_CORR] + f * scal[SC_WB_CORR]) + (scal[SC_WA_NOT] + f * scal[SC_WB_NOT]);
      if (-ntmin < pr.vr_tmin || tmax > pr.vr_tmax) scal[SC_T_FLAG] = 1.0;       // valid_range_t (:940)
    }
  }
}
void launch_apply_fixers(const DevTables& t, const Params& pr, int slot_fut, double* ps, double2* lnps_fut, double2* lnps_cur,
                         double2* ts_fut, double2* ts_cur, double rc_raw, double* scal, double denom, int owns_m0, int do_mass,
                         int do_energy, cudaStream_t st) {
  size_t n = (size_t)t.g.Jloc * t.g.I;
  apply_fixers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t, pr, slot_fut, ps, lnps_fut, lnps_cur, ts_fut, ts_cur, rc_raw, scal,
                                                                  denom, owns_m0, do_mass, do_energy);
}

// T(slot) += pending shift; shift = 0   (host mirrors, restart, the spectral_dynamics host API)
__global__ void materialize_t_kernel(double* __restrict__ T, size_t n, const double* __restrict__ scal, int slot) {
  const double sh = scal[SC_TSHIFT0 + slot];
  if (sh == 0.0) return;                                   // nothing pending: the launch costs a few microseconds
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) T[i] = T[i] + sh;
}
__global__ void clear_shift_kernel(double* scal, int slot) { scal[SC_TSHIFT0 + slot] = 0.0; }
void launch_materialize_t(const DevTables& t, double* T, double* scal, int slot, cudaStream_t st) {
  const size_t n = (size_t)t.g.K * t.g.Jloc * t.g.I;
  materialize_t_kernel<<<148 * 8, 256, 0, st>>>(T, n, scal, slot);
  clear_shift_kernel<<<1, 1, 0, st>>>(scal, slot);
}

// ---------------------------------------------------------------------------------------------
// compute_pressures_and_heights (dry): p_half, p_full, z_half, z_full for diagnostics / physics API
// ---------------------------------------------------------------------------------------------
__global__ void press_heights_kernel(DevTables t, Params pr, const double* __restrict__ T, const double* __restrict__ ps,
                                     const double* __restrict__ phis, double* __restrict__ p_full, double* __restrict__ p_half,
                                     double* __restrict__ z_full, double* __restrict__ z_half) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const int K = g.K;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double p_s = ps[col];
  double gh_below = phis[col];
  const double ln_ps = pr.pure_sigma ? log(p_s) : 0.0;
  double ln_half_below = ln_p_half(t, pr, K, t.pk[K] + t.bk[K] * p_s, ln_ps);
  if (p_half) p_half[(size_t)K * plane + col] = t.pk[K] + t.bk[K] * p_s;
  if (z_half) z_half[(size_t)K * plane + col] = gh_below / pr.grav;
  for (int k0 = K - 1; k0 >= 0; k0 -= 4) {            // four levels of loads in flight, consumed serially
    double t4[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) if (k0 - d >= 0) t4[d] = T[(size_t)(k0 - d) * plane + col];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int k = k0 - d;
      if (k < 0) break;
      const size_t e = (size_t)k * plane + col;
      PressLevel pl;
      press_level(t, pr, k, p_s, ln_ps, ln_half_below, pl);
      const double tt = t4[d];
      const double gfull = gh_below + pr.rdgas * tt * (pl.ln_half_k1 - pl.ln_full);
      double gh = 0.0;
      if (!(k == 0 && pr.pk0_zero)) gh = gh_below + pr.rdgas * tt * (pl.ln_half_k1 - pl.ln_half_k);
      if (p_full) p_full[e] = pl.p_full;
      if (p_half) p_half[e] = pl.p_half_k;
      if (z_full) z_full[e] = gfull / pr.grav;
      if (z_half) z_half[e] = gh / pr.grav;
      gh_below = gh; ln_half_below = pl.ln_half_k;
    }
  }
}
void launch_press_heights(const DevTables& t, const Params& pr, const double* T, const double* ps, const double* phis,
                          double* p_full, double* p_half, double* z_full, double* z_half, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  press_heights_kernel<<<grid, 128, 0, st>>>(t, pr, T, ps, phis, p_full, p_half, z_full, z_half);
}

// small helpers
__global__ void scale_rows_kernel(DevTables t, double* __restrict__ f, int nlev, int which) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const double s = (which == 0) ? t.cosm_lat[g.j0 + jl] : t.cos_lat[g.j0 + jl];
  const size_t plane = (size_t)g.Jloc * g.I;
  for (int k = 0; k < nlev; ++k) f[(size_t)k * plane + (size_t)jl * g.I + i] *= s;
}
void launch_divide_by_cos(const DevTables& t, double* f, int nlev, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  scale_rows_kernel<<<grid, 128, 0, st>>>(t, f, nlev, 0);
}

}  // namespace isca
