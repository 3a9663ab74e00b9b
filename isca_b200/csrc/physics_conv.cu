// physics_conv.cu -- simplified Betts-Miller (quasi-equilibrium) moist convection, one thread per column.
//
//   qe_moist_convection / SBM_convection_scheme / CAPE_calculation / CAPE_below_LCL / CAPE_above_LCL / set_reference_profiles /
//   Pq_calculation / Pt_calculation / do_deep_convection / do_shallow_convection / level_of_zero_precip / get_lcl_temp
//                         atmos_param/qe_moist_convection/qe_moist_convection.F90:157-1082
//   LCL temperature table (get_val_min_max, generate_lcl_table, lcl_temp) :113-154, 1086-1175 -- built on the host at create
//
// Trip counts depend on the column (LCL, LZB, level of zero precipitation): warps diverge, but every global access stays
// coalesced across the 32 columns of a warp.  Parcel temperature / mixing ratio live in thread-local arrays; the four 3-D outputs
// double as the working profiles.  Level indices kLZB, kLCL are reported 1-based (0 = none), as the reference does.
#include "physics_common.h"

using namespace isca_phys;

namespace {

struct ConvConst {
  double tau_bm, rhbm, Tmin, val_min, val_max, val_inc;
  const double* lcl_table; int ntab;
  double rdgas, rvgas, cp_air, hlv, kappa, grav;
};

__device__ __forceinline__ double mixing_ratio(const ConvConst& c, double e, double p) { return c.rdgas * e / c.rvgas / (p - e); }
__device__ __forceinline__ double virtual_temp(const ConvConst& c, double T, double r) {
  double q = r / (1.0 + r);
  return T * (1.0 + q * (c.rvgas / c.rdgas - 1.0));
}

// bytes/column: read Tin, qin, p_full (3K) + p_half (K+1); write deltaT, deltaq, qref, Tref (4K) + 7  ~ (8K + 8) * 8
__global__ void __launch_bounds__(128, ISCA_COL_MINB) sbm_convection_kernel(SvpDev s, ConvConst c, int ncol, int K, double dt,
    const double* __restrict__ Tin, const double* __restrict__ qin, const double* __restrict__ p_full, const double* __restrict__ p_half,
    double* __restrict__ rain, double* __restrict__ deltaT, double* __restrict__ deltaq, double* __restrict__ qref, double* __restrict__ Tref,
    int* __restrict__ convflag, int* __restrict__ kLZBs, int* __restrict__ kLCLs, double* __restrict__ CAPE_o, double* __restrict__ CIN_o,
    double* __restrict__ itq_o, double* __restrict__ itt_o, int* err) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const size_t nc = ncol;
  const double small = 1.0e-10, pref = 1.0e5;
  double Tp[ISCA_KMAX], rp[ISCA_KMAX];
  int bad = 0;
  auto T_in = [&](int k) { return Tin[(size_t)k * nc + col]; };
  auto q_in = [&](int k) { return qin[(size_t)k * nc + col]; };
  auto r_in = [&](int k) { double q = q_in(k); return q / (1.0 - q); };
  auto pf = [&](int k) { return p_full[(size_t)k * nc + col]; };
  auto ph = [&](int k) { return p_half[(size_t)k * nc + col]; };
  auto es_of = [&](double T) { double e, d; if (!svp_lookup(s, T, e, d)) bad |= 1; return e; };
  auto tv_in = [&](int k) { return virtual_temp(c, T_in(k), r_in(k)); };
  auto dlnph = [&](int k) { return log(ph(k + 1) / ph(k)); };
  const int ks = K - 1;

  // ---- CAPE_calculation :373-420
  for (int k = 0; k < K; ++k) { Tp[k] = T_in(k); rp[k] = r_in(k); }
  bool nocape = true, skip = false;
  double cape = 0.0, cin = 0.0;
  int kLZB = 0, kLCL = 0;
  auto nocape_reset = [&]() { kLZB = 0; cin = 0.0; for (int k = 0; k < K; ++k) { Tp[k] = T_in(k); rp[k] = r_in(k); } };
  {
    const double T0 = T_in(ks), r0 = r_in(ks);
    const double rs = mixing_ratio(c, es_of(T0), pf(ks));
    if (r0 >= rs) {                                            // saturated at the lowest level (CAPE_below_LCL :448-456)
      kLCL = K;
      Tp[ks] = T0 + (r0 - rs) / ((c.cp_air / (c.hlv + small)) + (c.hlv * rs) / c.rvgas / (T0 * T0));
      rp[ks] = mixing_ratio(c, es_of(Tp[ks]), pf(ks));
    } else {
      const double theta0 = T0 * pow(pref / pf(ks), c.kappa);
      if (r0 <= 0.0) skip = true;
      else {
        double value = log(pow(theta0, -1.0 / c.kappa) * pref * r0 / (c.rdgas / c.rvgas + r0));
        double TLCL = 0.0;
        {                                                      // get_lcl_temp :1053-1082
          int iv = (int)floor((value - c.val_min) / c.val_inc) + 1;
          if (!(value >= c.val_min) || !(value <= c.val_max) || iv + 1 > c.ntab) { bad |= 2; iv = 1; value = c.val_min; }
          double w_floor = c.val_min + (iv - 1) * c.val_inc;
          double w_ceil = (value - w_floor) / c.val_inc;
          TLCL = c.lcl_table[iv] * w_ceil - c.lcl_table[iv - 1] * (w_ceil - 1.0);
        }
        double pLCL = pref * pow(TLCL / theta0, 1.0 / c.kappa);
        if (pLCL < pf(0)) { pLCL = pf(0); TLCL = theta0 * pow(pLCL / pref, c.kappa); }
        int k = ks;
        cin = 0.0;
        while (pf(k) > pLCL) {
          Tp[k] = theta0 * pow(pf(k) / pref, c.kappa);
          rp[k] = mixing_ratio(c, es_of(Tp[k]), pf(k));
          cin = cin + c.rdgas * (tv_in(k) - virtual_temp(c, Tp[k], r0)) * dlnph(k);
          --k;
        }
        kLCL = k + 1;
        double a = c.kappa * TLCL + (c.hlv / c.cp_air) * r0;
        double b = (c.hlv * c.hlv) * r0 / (c.cp_air * c.rvgas * (TLCL * TLCL));
        double dtdlnp = a / (1.0 + b);
        Tp[k] = TLCL + dtdlnp * log(pf(k) / pLCL) / 2;
        if (Tp[k] < c.Tmin && nocape) { skip = true; nocape_reset(); }
        else {
          rp[k] = mixing_ratio(c, es_of(Tp[k]), (pf(k) + pLCL) / 2);
          a = c.kappa * Tp[k] + (c.hlv / c.cp_air) * rp[k];
          b = (c.hlv * c.hlv) * rp[k] / (c.cp_air * c.rvgas * (Tp[k] * Tp[k]));
          dtdlnp = a / (1.0 + b);
          Tp[k] = TLCL + dtdlnp * log(pf(k) / pLCL);
          if (Tp[k] < c.Tmin && nocape) { skip = true; nocape_reset(); }
          else {
            rp[k] = mixing_ratio(c, es_of(Tp[k]), pf(k));
            double tvp = virtual_temp(c, Tp[k], rp[k]), tve = tv_in(k);
            if (tvp < tve && nocape) cin = cin + c.rdgas * (tve - tvp) * dlnph(k);
            else { cape = cape + c.rdgas * (tvp - tve) * dlnph(k); nocape = false; }
          }
        }
      }
    }
    // CAPE_above_LCL :597-682
    if (skip) { if (nocape) nocape_reset(); }
    else {
      for (int k = kLCL - 2; k >= 0; --k) {
        double a = c.kappa * Tp[k + 1] + (c.hlv / c.cp_air) * rp[k + 1];
        double b = (c.hlv * c.hlv) * rp[k + 1] / (c.cp_air * c.rvgas * (Tp[k + 1] * Tp[k + 1]));
        double dtdlnp = a / (1.0 + b);
        double lp = log(pf(k) / pf(k + 1));
        Tp[k] = Tp[k + 1] + dtdlnp * lp / 2;
        if (Tp[k] < c.Tmin && nocape) { nocape_reset(); break; }
        rp[k] = mixing_ratio(c, es_of(Tp[k]), (pf(k) + pf(k + 1)) / 2);
        a = c.kappa * Tp[k] + (c.hlv / c.cp_air) * rp[k];
        b = (c.hlv * c.hlv) * rp[k] / (c.cp_air * c.rvgas * (Tp[k] * Tp[k]));
        dtdlnp = a / (1.0 + b);
        Tp[k] = Tp[k + 1] + dtdlnp * lp;
        if (Tp[k] < c.Tmin && nocape) { nocape_reset(); break; }
        rp[k] = mixing_ratio(c, es_of(Tp[k]), pf(k));
        double tvp = virtual_temp(c, Tp[k], rp[k]), tve = tv_in(k);
        if (tvp < tve && nocape) cin = cin + c.rdgas * (tve - tvp) * dlnph(k);
        else if (tvp < tve && !nocape) { kLZB = k + 2; break; }
        else { cape = cape + c.rdgas * (tvp - tve) * dlnph(k); nocape = false; }
      }
    }
  }

  // ---- SBM_convection_scheme :299-356
  auto DT = [&](int k) -> double& { return deltaT[(size_t)k * nc + col]; };
  auto DQ = [&](int k) -> double& { return deltaq[(size_t)k * nc + col]; };
  auto QR = [&](int k) -> double& { return qref[(size_t)k * nc + col]; };
  auto TR = [&](int k) -> double& { return Tref[(size_t)k * nc + col]; };
  auto full = [&](int k1, int k2) {                            // set_profiles_to_full_model_values, 1-based inclusive
    for (int k = k1 - 1; k < k2; ++k) { TR(k) = T_in(k); QR(k) = q_in(k); DT(k) = 0.0; DQ(k) = 0.0; }
  };
  int flag = 0;
  double Pq = 0.0, itq = 0.0, itt = 0.0;
  if (cape > 0.0) {
    flag = 1;
    const int lz = kLZB > 1 ? kLZB : 1;                        // kLZB = 0 (no LZB below the top) is out of bounds in the reference
    for (int k = lz - 1; k < K; ++k) {                         // set_reference_profiles :768-796
      TR(k) = Tp[k];
      double eref = c.rhbm * pf(k) * rp[k] / (rp[k] + (c.rdgas / c.rvgas));
      double r = mixing_ratio(c, eref, pf(k));
      QR(k) = r / (1.0 + r);
    }
    full(1, lz - 1 > 1 ? lz - 1 : 1);
    if (lz == 1) { /* level 1 was just reset to the model values by the reference as well (k = max(kLZB-1,1) = 1) */ }
    Pq = 0.0;                                                  // Pq_calculation :715-736
    for (int k = lz - 1; k < K; ++k) { double d = -(q_in(k) - QR(k)) * dt / c.tau_bm; DQ(k) = d; Pq = Pq + d * (ph(k) - ph(k + 1)); }
    Pq = Pq / c.grav;
    double Pt = 0.0;                                           // Pt_calculation :739-764
    for (int k = lz - 1; k < K; ++k) {
      double d = -(T_in(k) - TR(k)) * dt / c.tau_bm; DT(k) = d;
      Pt = Pt + (c.cp_air / (c.hlv + small)) * d * (ph(k + 1) - ph(k));
    }
    Pt = Pt / c.grav;
    if (Pq > 0.0 && Pt > 0.0) {
      flag = 2;
      if (Pq > Pt) {                                           // do_change_time_scale_deepconv :1001-1017
        itq = Pt / Pq / c.tau_bm;
        for (int k = lz - 1; k < K; ++k) DQ(k) = c.tau_bm * itq * DQ(k);
        Pq = Pt;
        itt = 1.0 / c.tau_bm;
      } else {                                                 // do_change_Tref_deepconv :968-999
        double deltak = 0.0;
        for (int k = lz - 1; k < K; ++k) deltak = deltak - (DT(k) + (c.hlv / c.cp_air) * DQ(k)) * (ph(k + 1) - ph(k));
        deltak = deltak / (ph(K) - ph(lz - 1));
        for (int k = lz - 1; k < K; ++k) { TR(k) = TR(k) + deltak * c.tau_bm / dt; DT(k) = DT(k) + deltak; }
      }
    } else if (Pt > 0.0) {                                     // do_shallow_convection :800-929
      int k = lz;
      while (Pq < 0.0 && k <= K) { Pq = Pq - DQ(k - 1) * (ph(k - 1) - ph(k)) / c.grav; ++k; }
      const int k_top = k - 1;
      const bool found = Pq > 0.0;
      if (k_top > lz) full(lz, k_top - 1);
      if (found) {                                             // change_Tref_LZB_shallowconv :889-920
        double cc = Pq * c.grav / (DQ(k_top - 1) * (ph(k_top) - ph(k_top - 1)));
        DQ(k_top - 1) = DQ(k_top - 1) * cc;
        DT(k_top - 1) = DT(k_top - 1) * cc;
        double deltak = 0.0;
        for (int kk = k_top - 1; kk < K; ++kk) deltak = deltak + DT(kk) * (ph(kk) - ph(kk + 1));
        deltak = deltak / (ph(K) - ph(k_top - 1));
        if (k_top != K) for (int kk = k_top - 1; kk < K; ++kk) { DT(kk) = DT(kk) + deltak; TR(kk) = TR(kk) + deltak * c.tau_bm / dt; }
      } else {
        if (k_top == lz) full(K, K); else full(lz, k_top);
      }
      Pq = 0.0;
    } else { Pq = 0.0; full(1, K); }
  } else { Pq = 0.0; full(1, K); }
  // levels above the reference profiles that the branches above did not touch hold the model values / zero increments
  // (deltaq, deltaT are zero-initialised, Tref = Tp = Tin there in the reference); with cape > 0 they were set by full(1, lz-1)
  rain[col] = Pq; convflag[col] = flag; kLZBs[col] = kLZB; kLCLs[col] = kLCL; CAPE_o[col] = cape; CIN_o[col] = cin;
  // the reference zeroes the whole relaxation-rate arrays inside its column loop (:283-284): only the last column keeps a value
  itq_o[col] = col == ncol - 1 ? itq : 0.0;
  itt_o[col] = col == ncol - 1 ? itt : 0.0;
  if (bad) atomicOr(err, bad);
}

}  // namespace

namespace isca_phys {

// LCL temperature table (qe_moist_convection.F90:113-154, 1086-1175) from the host copy of the saturation table
int build_lcl_table(IscaPhysics p) {
  const IscaPhysicsConfig& g = p->cfg;
  const std::vector<double>& tb = p->svp_host;
  const int n = p->svp.n;
  const double kappa = g.rdgas / g.cp_air;
  bool bad = false;
  auto es = [&](double T) {
    double tmp = T - p->svp.tminl, x = p->svp.dtinvl * (tmp + p->svp.tepsl);
    if (!(x > -1.0 && x < (double)n)) { bad = true; return 1.0; }
    int ind = (int)x;
    double dl = tmp - p->svp.dtres * ind;
    return tb[ind] + dl * (tb[n + ind] + dl * tb[2 * n + ind]);
  };
  p->lcl_val_min = std::log(es(g.Tmin) / std::pow(g.Tmin, 1.0 / kappa));
  p->lcl_val_max = std::log(es(g.Tmax) / std::pow(g.Tmax, 1.0 / kappa));
  if (bad || !(p->lcl_val_max > p->lcl_val_min) || !(g.val_inc > 0.0)) return fail(p, "qe_moist_convection_init: Tmin/Tmax outside the saturation vapour pressure table");
  int size = (int)std::ceil((p->lcl_val_max - p->lcl_val_min) / g.val_inc);
  std::vector<double> tab(size);
  double guess = g.Tmin;
  for (int k = 0; k < size; ++k) {
    double value = p->lcl_val_min + k * g.val_inc;
    double T = guess, dT = 1.0e-7 + 1.0;
    int iter = 0;
    while (std::fabs(dT) > 1.0e-7 && iter < 100) {
      double f = value - std::log(es(T) * std::pow(T, -1.0 / kappa));
      double df = 1.0 / kappa * std::pow(T, -1.0) - g.hlv / g.rvgas * std::pow(T, -2.0);
      dT = f / df; T = T - dT; ++iter;
    }
    if (!(dT < 1.0e-7) || bad) { fail(p, "qe_moist_convection: LCL calculation did not converge. Precision not achieved."); return 2; }
    tab[k] = T; guess = T;
  }
  if (!p->lcl_tab.ensure(size)) return fail(p, "cudaMalloc failed");
  PCK(cudaMemcpyAsync(p->lcl_tab.p, tab.data(), size * sizeof(double), cudaMemcpyHostToDevice, p->st));
  PCK(cudaStreamSynchronize(p->st));
  p->lcl_n = size;
  return 0;
}

void launch_sbm_convection(IscaPhysics p, double dt, const double* Tin, const double* qin, const double* p_full, const double* p_half,
                           double* rain, double* deltaT, double* deltaq, double* qref, double* Tref, int* convflag, int* kLZBs, int* kLCLs,
                           double* cape, double* cin, double* itq, double* itt) {
  ConvConst c;
  c.tau_bm = p->cfg.tau_bm; c.rhbm = p->cfg.rhbm; c.Tmin = p->cfg.Tmin; c.val_min = p->lcl_val_min; c.val_max = p->lcl_val_max;
  c.val_inc = p->cfg.val_inc; c.lcl_table = p->lcl_tab.p; c.ntab = p->lcl_n;
  c.rdgas = p->cfg.rdgas; c.rvgas = p->cfg.rvgas; c.cp_air = p->cfg.cp_air; c.hlv = p->cfg.hlv; c.kappa = p->cfg.rdgas / p->cfg.cp_air;
  c.grav = p->cfg.grav;
  sbm_convection_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>(p->svp, c, (int)p->ncol, p->K, dt, Tin, qin, p_full, p_half, rain, deltaT, deltaq,
                                                               qref, Tref, convflag, kLZBs, kLCLs, cape, cin, itq, itt, p->d_err);
}

}  // namespace isca_phys

extern "C" int isca_b200_qe_moist_convection(IscaPhysics p, double dt, const double* Tin, const double* qin, const double* p_full,
                                             const double* p_half, double* rain, double* snow, double* deltaT, double* deltaq,
                                             double* qref, int* convflag, int* kLZBs, double* cape, double* cin,
                                             double* invtau_q_relaxation, double* invtau_t_relaxation, double* Tref, int* kLCLs) {
  if (!p) return fail(nullptr, "null handle");
  if (!snow || !convflag || !kLZBs || !kLCLs) return fail(p, "null output array");
  if (!p->lcl_err.empty()) return fail(p, p->lcl_err);
  size_t nc = p->ncol, n3 = nc * p->K;
  Dev* b = p->buf;
  if (up(p, b[0], Tin, n3) || up(p, b[1], qin, n3) || up(p, b[2], p_full, n3) || up(p, b[3], p_half, n3 + nc)) return 1;
  for (int i = 4; i < 8; ++i) if (!b[i].ensure(n3)) return fail(p, "cudaMalloc failed");
  for (int i = 8; i < 15; ++i) if (!b[i].ensure(nc)) return fail(p, "cudaMalloc failed");
  int* iflag = reinterpret_cast<int*>(b[13].p);                // three int planes share two double-sized buffers
  int* ilzb = iflag + nc;
  int* ilcl = reinterpret_cast<int*>(b[14].p);
  launch_sbm_convection(p, dt, b[0].p, b[1].p, b[2].p, b[3].p, b[8].p, b[4].p, b[5].p, b[6].p, b[7].p, iflag, ilzb, ilcl, b[9].p, b[10].p,
                        b[11].p, b[12].p);
  if (down(p, b[8], rain, nc) || down(p, b[4], deltaT, n3) || down(p, b[5], deltaq, n3) || down(p, b[6], qref, n3) || down(p, b[7], Tref, n3) ||
      down(p, b[9], cape, nc) || down(p, b[10], cin, nc) || down(p, b[11], invtau_q_relaxation, nc) || down(p, b[12], invtau_t_relaxation, nc)) return 1;
  PCK(cudaMemcpyAsync(convflag, iflag, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaMemcpyAsync(kLZBs, ilzb, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaMemcpyAsync(kLCLs, ilcl, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  std::memset(snow, 0, nc * sizeof(double));                   // snow = 0. (:367)
  int e = 0;
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaStreamSynchronize(p->st));
  if (e) {
    PCK(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->st));
    if (e & 2) return fail(p, "qe_moist_convection: get_lcl_temp: value outside the LCL temperature table (too low / too high)");
    return fail(p, "qe_moist_convection: lookup_es: temperature outside the saturation vapour pressure table (table overflow)");
  }
  return 0;
}
