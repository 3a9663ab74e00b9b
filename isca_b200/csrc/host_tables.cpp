// host_tables.cpp -- init-time tables, computed on the host in fp64 with the reference's
// recurrences, uploaded once.  Citations are relative to /root/reference/src.
#include "common.h"
#include <algorithm>
#include <cstring>

namespace isca {

static const double PI = 3.14159265358979323846;   // shared/constants/constants.F90:240

// ---------------------------------------------------------------------------------------------
// Decomposition.  Grid: latitudes in P equal contiguous blocks (tools/spec_mpp.F90:61-75).
// Spectral: zonal wavenumbers dealt to ranks in snake order so that sum_m (M-m+2) is balanced
// (the reference's contiguous m blocks, spec_mpp.F90:77-80, are load-imbalanced under triangular
// truncation -- SURVEY F9).
// ---------------------------------------------------------------------------------------------
void build_geometry(const IscaConfig& c, int rank, int nranks, Geometry& g) {
  g.I = c.lon_max; g.J = c.lat_max; g.K = c.num_levels; g.M = c.num_fourier; g.N = c.num_spherical;
  g.Jh = g.J / 2; g.P = nranks; g.rank = rank;
  if (g.N != g.M + 1) throw std::runtime_error("num_spherical must equal num_fourier+1 (triangular truncation)");
  if (g.J % 2) throw std::runtime_error("lat_max must be even");
  if (g.J % nranks) throw std::runtime_error("lat_max must be divisible by the number of ranks (spec_mpp.F90:69-75)");
  if (g.I < 2 * g.M + 1 && g.I / 2 < g.M) throw std::runtime_error("lon_max too small for num_fourier");
  if ((g.I & (g.I - 1)) != 0 || g.I < 32 || g.I > 1024) throw std::runtime_error("lon_max must be a power of two in [32, 1024]");
  if (g.K > ISCA_KMAX) throw std::runtime_error("num_levels exceeds ISCA_KMAX");
  g.Jloc = g.J / nranks; g.j0 = rank * g.Jloc;
  g.owner.assign(g.M + 1, 0);
  std::vector<std::vector<int>> lists(nranks);
  for (int m = 0; m <= g.M; ++m) {
    int r = m % (2 * nranks);
    if (r >= nranks) r = 2 * nranks - 1 - r;
    g.owner[m] = r;
    lists[r].push_back(m);
  }
  g.nm_rank.resize(nranks); g.roff.assign(nranks + 1, 0); g.nm_max = 0;
  for (int r = 0; r < nranks; ++r) {
    g.nm_rank[r] = (int)lists[r].size();
    g.roff[r + 1] = g.roff[r] + g.nm_rank[r];
    g.nm_max = std::max(g.nm_max, g.nm_rank[r]);
  }
  g.pos.assign(g.M + 1, 0);
  for (int r = 0; r < nranks; ++r)
    for (size_t i = 0; i < lists[r].size(); ++i) g.pos[lists[r][i]] = g.roff[r] + (int)i;
  g.m_of = lists[rank];
  g.nm = (int)g.m_of.size();
  g.off.assign(g.nm + 1, 0);
  for (int mi = 0; mi < g.nm; ++mi) g.off[mi + 1] = g.off[mi] + (g.M - g.m_of[mi] + 2);
  g.T = g.off[g.nm];
}

// tools/gauss_and_legendre.F90:111-183
void compute_gaussian(int n_hem, std::vector<double>& sin_hem, std::vector<double>& wts_hem) {
  double converg = 1.0;
  for (int i = 0; i < 15; ++i) converg *= 0.1;      // .1**precision(real*8)
  converg = std::pow(0.1, 15);
  const int itermax = 10;
  const int n = 2 * n_hem;
  sin_hem.assign(n_hem, 0.0); wts_hem.assign(n_hem, 0.0);
  for (int i = 1; i <= n_hem; ++i) {
    double z = std::cos(PI * (i - 0.25) / (n + 0.5));
    double pp = 0.0;
    bool ok = false;
    for (int iter = 0; iter < itermax; ++iter) {
      double p1 = 1.0, p2 = 0.0, p3;
      for (int j = 1; j <= n; ++j) {
        p3 = p2; p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < converg) { ok = true; break; }
    }
    if (!ok) throw std::runtime_error("compute_gaussian: abscissas failed to converge in itermax iterations");
    sin_hem[i - 1] = z;
    wts_hem[i - 1] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

// matrix_invert.F90:38-130 (Gauss-Jordan with column pivoting on the row's largest element).
// a is row-major n x n: a[i*n+j] = matrix(i,j).
void invert_matrix(std::vector<double>& a, int n) {
  std::vector<double> ac(2 * n * n, 0.0);     // ac(i,j), i < 2n, j < n  -> ac[i*n + j]
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) ac[i * n + j] = a[i * n + j];
  for (int j = 0; j < n; ++j) ac[(n + j) * n + j] = 1.0;
  std::vector<double> h(2 * n), dd(2 * n);
  double det = 1.0;
  for (int k = 0; k < n; ++k) {
    int mx = 0; double rmax = std::fabs(ac[k * n + k]);
    for (int i = 0; i < n - k; ++i)
      if (std::fabs(ac[k * n + k + i]) > rmax) { rmax = std::fabs(ac[k * n + k + i]); mx = i; }
    int L = mx + k;
    if (k - L < 0) {
      for (int i = k; i < 2 * n; ++i) std::swap(ac[i * n + k], ac[i * n + L]);
      det = -det;
    }
    det *= ac[k * n + k];
    if (std::fabs(det) < 1.0e-30) throw std::runtime_error("invert: the input matrix appears to be singular");
    double piv = ac[k * n + k];
    for (int i = k; i < 2 * n; ++i) h[i] = ac[i * n + k] / piv;
    std::vector<double> rowk(n);
    for (int j = 0; j < n; ++j) rowk[j] = ac[k * n + j];
    for (int i = k; i < 2 * n; ++i)
      for (int j = 0; j < n; ++j) ac[i * n + j] = ac[i * n + j] - h[i] * rowk[j];
    for (int i = k; i < 2 * n; ++i) ac[i * n + k] = h[i];
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) a[i * n + j] = ac[(n + i) * n + j];
}

// press_and_geopot.F90:152-221 for a single surface pressure (simmons_and_burridge)
static void pressure_variables_1d(const std::vector<double>& pk, const std::vector<double>& bk, double ps,
                                  std::vector<double>& p_half, std::vector<double>& ln_p_half,
                                  std::vector<double>& p_full, std::vector<double>& ln_p_full) {
  int K = (int)pk.size() - 1;
  p_half.resize(K + 1); ln_p_half.resize(K + 1); p_full.resize(K); ln_p_full.resize(K);
  for (int k = 0; k <= K; ++k) p_half[k] = pk[k] + bk[k] * ps;
  if (pk[0] == 0.0 && bk[0] == 0.0) {
    for (int k = 1; k <= K; ++k) ln_p_half[k] = std::log(p_half[k]);
    for (int k = 1; k < K; ++k) {
      double alpha = 1.0 - p_half[k] * (ln_p_half[k + 1] - ln_p_half[k]) / (p_half[k + 1] - p_half[k]);
      ln_p_full[k] = ln_p_half[k + 1] - alpha;
    }
    ln_p_full[0] = ln_p_half[1] + (-1.0);
    ln_p_half[0] = 0.0;
  } else {
    for (int k = 0; k <= K; ++k) ln_p_half[k] = std::log(p_half[k]);
    for (int k = 0; k < K; ++k) {
      double alpha = 1.0 - p_half[k] * (ln_p_half[k + 1] - ln_p_half[k]) / (p_half[k + 1] - p_half[k]);
      ln_p_full[k] = ln_p_half[k + 1] - alpha;
    }
  }
  for (int k = 0; k < K; ++k) p_full[k] = std::exp(ln_p_full[k]);
}

namespace {
struct ImplicitRef {
  int K; double rdgas, kappa, ref_ps;
  const std::vector<double>*pk, *bk, *dpk, *dbk;
  std::vector<double> t, lh, lf;
  // implicit.F90:414-480 (real 1-D)
  void linear_tp_tendency(const std::vector<double>& div, double& dt_p, std::vector<double>& dt_t) const {
    double dmean_tot = 0.0;
    std::vector<double> vert_vel(K + 1, 0.0), temp(K + 1, 0.0);
    dt_t.assign(K, 0.0);
    for (int k = 0; k < K; ++k) {
      double dp = (*dpk)[k] + (*dbk)[k] * ref_ps, dp_inv = 1 / dp;
      double dlog_1 = lh[k + 1] - lf[k], dlog_3 = lh[k + 1] - lh[k];
      double dmean = div[k] * dp;
      dt_t[k] = -kappa * t[k] * (dmean_tot * dlog_3 + dmean * dlog_1) * dp_inv;
      dmean_tot = dmean_tot + dmean;
      vert_vel[k + 1] = -dmean_tot;
    }
    dt_p = -dmean_tot;
    for (int k = 1; k < K; ++k) vert_vel[k] = vert_vel[k] + dmean_tot * (*bk)[k];
    for (int k = 1; k < K; ++k) temp[k] = -vert_vel[k] * (t[k] - t[k - 1]);
    for (int k = 0; k < K; ++k) {
      double dp = (*dpk)[k] + (*dbk)[k] * ref_ps, dp_inv = 1 / dp;
      dt_t[k] = dt_t[k] + .5 * dp_inv * (temp[k + 1] + temp[k]);
    }
  }
  // implicit.F90:329-359
  void linear_geopotential(const std::vector<double>& del_t, const std::vector<double>& dlh,
                           const std::vector<double>& dlf, std::vector<double>& geopot) const {
    std::vector<double> gh(K + 1, 0.0);
    geopot.assign(K, 0.0);
    for (int k = K - 1; k >= 1; --k)
      gh[k] = gh[k + 1] + rdgas * (del_t[k] * (lh[k + 1] - lh[k]) + t[k] * (dlh[k + 1] - dlh[k]));
    for (int k = 0; k < K; ++k)
      geopot[k] = gh[k + 1] + rdgas * (del_t[k] * (lh[k + 1] - lf[k]) + t[k] * (dlh[k + 1] - dlf[k]));
  }
};
}  // namespace

void build_tables(const IscaConfig& c, const Geometry& g, HostTables& t) {
  const int I = g.I, J = g.J, K = g.K, M = g.M, N = g.N, Jh = g.Jh;
  const double a = c.radius;
  // ---- Gaussian grid: define_gaussian (tools/spherical_fourier.F90:397-431)
  compute_gaussian(Jh, t.sin_hem, t.wts_hem);
  t.sin_lat.assign(J, 0); t.wts_lat.assign(J, 0); t.cos_lat.assign(J, 0); t.cosm_lat.assign(J, 0);
  t.cosm2_lat.assign(J, 0); t.deg_lat.assign(J, 0); t.rad_lat.assign(J, 0); t.coriolis.assign(J, 0);
  for (int j = 0; j < Jh; ++j) t.sin_lat[j] = -t.sin_hem[j];
  for (int j = 0; j < Jh; ++j) {
    t.sin_lat[J - 1 - j] = -t.sin_lat[j];
    t.wts_lat[j] = t.wts_hem[j];
    t.wts_lat[J - 1 - j] = t.wts_hem[j];
  }
  t.global_sum_of_wts = 0.0;
  for (int j = 0; j < J; ++j) {
    t.cos_lat[j] = std::sqrt(1 - t.sin_lat[j] * t.sin_lat[j]);
    t.cosm_lat[j] = 1. / t.cos_lat[j];
    t.cosm2_lat[j] = 1. / (t.cos_lat[j] * t.cos_lat[j]);
    t.deg_lat[j] = std::asin(t.sin_lat[j]) * 180.0 / PI;
    t.rad_lat[j] = t.deg_lat[j] * PI / 180.;             // atmosphere.F90:250-253
    t.coriolis[j] = 2 * c.omega * t.sin_lat[j];          // spectral_dynamics.F90:432
    t.global_sum_of_wts += t.wts_lat[j];                 // transforms.F90:329
  }
  t.deg_lon.resize(I);
  for (int i = 0; i < I; ++i) t.deg_lon[i] = i * 360.0 / I;   // grid_fourier.F90:105-118 (origin 0)

  // ---- vertical coordinate (init/vert_coordinate.F90:89-310)
  t.pk.assign(K + 1, 0.0); t.bk.assign(K + 1, 0.0);
  if (c.vert_coord_option == 0) {
    for (int k = 1; k <= K; ++k) t.bk[k - 1] = (double)(k - 1) / (double)K;
    t.bk[K] = 1.0;
  } else if (c.vert_coord_option == 1) {
    double s2 = 1.0 - c.surf_res;
    if (c.scale_heights == 0. || c.exponent == 0. || c.surf_res <= 0. || c.surf_res > 1.0)
      throw std::runtime_error("compute_vert_coord: invalid scale_heights/exponent/surf_res");
    for (int k = 1; k <= K; ++k) {
      double zeta = (1. - ((double)(k - 1) / (double)K));
      double z = c.surf_res * zeta + s2 * std::pow(zeta, c.exponent);
      t.bk[k - 1] = std::exp(-z * c.scale_heights);
    }
    t.bk[K] = 1.0; t.bk[0] = 0.0;
  } else if (c.vert_coord_option == 2) {
    if (!c.pk || !c.bk) throw std::runtime_error("vert_coord_option=input needs pk and bk");
    for (int k = 0; k <= K; ++k) { t.pk[k] = c.pk[k]; t.bk[k] = c.bk[k]; }
  } else if (c.vert_coord_option == 3) {
    // 'hybrid' (vert_coordinate.F90:141-147): sigma near the surface, pressure aloft, blended by transition() (:161-183);
    // both profiles are compute_uneven_sigma with zero_top = .false. (:248-272)
    if (c.scale_heights == 0. || c.exponent == 0. || c.surf_res <= 0. || c.surf_res > 1.0)
      throw std::runtime_error("compute_vert_coord: invalid scale_heights/exponent/surf_res");
    if (c.p_sigma < c.p_press) throw std::runtime_error("compute_vert_coord: p_sigma must be greater than p_press");
    double s2 = 1.0 - c.surf_res;
    std::vector<double> prof(K + 1);
    for (int k = 1; k <= K; ++k) {
      double zeta = (1. - ((double)(k - 1) / (double)K));
      double z = c.surf_res * zeta + s2 * std::pow(zeta, c.exponent);
      prof[k - 1] = std::exp(-z * c.scale_heights);
    }
    prof[K] = 1.0;
    for (int k = 0; k <= K; ++k) {
      double f;
      if (prof[k] <= c.p_press) f = 0.0;
      else if (prof[k] >= c.p_sigma) f = 1.0;
      else { double x = prof[k] - c.p_press, xx = c.p_sigma - c.p_press; double sn = std::sin(0.5 * PI * x / xx); f = sn * sn; }
      double a = 0.0 * f + prof[k] * (1.0 - f);            // a_sigma = 0, a_press = prof
      t.bk[k] = prof[k] * f + 0.0 * (1.0 - f);             // b_sigma = prof, b_press = 0
      t.pk[k] = c.reference_sea_level_press * a;
    }
  } else throw std::runtime_error("unsupported vert_coord_option");
  t.dpk.resize(K); t.dbk.resize(K);
  for (int k = 0; k < K; ++k) { t.dpk[k] = t.pk[k + 1] - t.pk[k]; t.dbk[k] = t.bk[k + 1] - t.bk[k]; }

  // ---- Legendre tables (tools/gauss_and_legendre.F90:47-108), packed rows of owned m only
  t.leg.assign((size_t)g.T * Jh, 0.0); t.legw.assign((size_t)g.T * Jh, 0.0);
  {
    std::vector<double> eps((size_t)(N + 1) * (M + 1)), b(M + 1, 0.0), poly((size_t)(N + 1) * (M + 1));
    for (int n = 0; n <= N; ++n)
      for (int m = 0; m <= M; ++m) {
        double m2 = (double)m * m, l2 = (double)(m + n) * (m + n);
        eps[(size_t)n * (M + 1) + m] = std::sqrt((l2 - m2) / (4.0 * l2 - 1.0));
      }
    for (int m = 1; m <= M; ++m) b[m] = std::sqrt(0.5 * (2.0 * (double)m + 1.0) / (double)m);
    for (int j = 0; j < Jh; ++j) {
      double s = t.sin_hem[j];
      double cl = std::sqrt(1 - s * s);
      auto P = [&](int m, int n) -> double& { return poly[(size_t)n * (M + 1) + m]; };
      auto E = [&](int m, int n) -> double { return eps[(size_t)n * (M + 1) + m]; };
      P(0, 0) = std::sqrt(0.5);
      for (int m = 1; m <= M; ++m) P(m, 0) = b[m] * cl * P(m - 1, 0);
      for (int m = 0; m <= M; ++m) P(m, 1) = s * P(m, 0) / E(m, 1);
      for (int n = 2; n <= N; ++n)
        for (int m = 0; m <= M; ++m) P(m, n) = (s * P(m, n - 1) - E(m, n - 1) * P(m, n - 2)) / E(m, n);
      for (int mi = 0; mi < g.nm; ++mi) {
        int m = g.m_of[mi], Nm = M - m + 2;
        for (int n = 0; n < Nm; ++n) {
          size_t p = (size_t)(g.off[mi] + n);
          t.leg[p * Jh + j] = P(m, n);
          t.legw[p * Jh + j] = P(m, n) * t.wts_hem[j];     // spherical_fourier.F90:389-391
        }
      }
    }
  }

  // ---- spherical_init coefficient tables (tools/spherical.F90:137-216), per packed row
  auto epsf = [](double m, double L) { return std::sqrt((L * L - m * m) / (4.0 * L * L - 1.0)); };
  const int T = g.T;
  t.eigen.assign(T, 0); t.coef_uvm.assign(T, 0); t.coef_uvc.assign(T, 0); t.coef_uvp.assign(T, 0);
  t.coef_alpm.assign(T, 0); t.coef_alpp.assign(T, 0); t.coef_dym.assign(T, 0); t.coef_dx.assign(T, 0);
  t.coef_dyp.assign(T, 0); t.trunc_mask.assign(T, 0); t.row_m.assign(T, 0); t.row_n.assign(T, 0);
  t.damping.assign(T, 0); t.damping_vor.assign(T, 0); t.damping_div.assign(T, 0);
  t.eddy_sponge.assign(T, 0); t.zmu_sponge.assign(T, 0); t.zmv_sponge.assign(T, 0);
  // spectral_damping_init (model/spectral_damping.F90:56-168), resolution_dependent
  int o_v = c.damping_order_vor == -1 ? c.damping_order : c.damping_order_vor;
  int o_d = c.damping_order_div == -1 ? c.damping_order : c.damping_order_div;
  double c_v = c.damping_coeff_vor == -1. ? c.damping_coeff : c.damping_coeff_vor;
  double c_d = c.damping_coeff_div == -1. ? c.damping_coeff : c.damping_coeff_div;
  double eig_ref = (double)(N - 1) * ((double)(N - 1) + 1.0) / (a * a);   // eigen(0, num_spherical-1)
  for (int mi = 0; mi < g.nm; ++mi) {
    int m = g.m_of[mi], Nm = M - m + 2;
    for (int n = 0; n < Nm; ++n) {
      int p = g.off[mi] + n;
      double fm = m, L = m + n;
      double e = epsf(fm, L);
      if (L == 0) e = 0.0;   // (0-0)/(0-1) = -0 -> sqrt = 0
      t.row_m[p] = mi; t.row_n[p] = n;
      t.trunc_mask[p] = (m + n > N - 1 || (c.make_symmetric && m > 0)) ? 0.0 : 1.0;     // spherical.F90:183-185
      t.eigen[p] = L * (L + 1.0) / (a * a);
      if (L > 0) { t.coef_uvm[p] = -a * e / L; t.coef_uvc[p] = -a * fm / (L * (L + 1.0)); }
      double e1 = epsf(fm, L + 1.0);              // epsilon(m, n+1)
      if (n <= N - 1) {
        t.coef_uvp[p] = -a * e1 / (L + 1.0);
        t.coef_alpp[p] = L * e1 / a;
        t.coef_dyp[p] = (L + 2.0) * e1 / a;
      }
      t.coef_alpm[p] = (L + 1.0) * e / a;
      t.coef_dym[p] = (L - 1.0) * e / a;
      t.coef_dx[p] = fm / a;
      double ratio = t.eigen[p] / eig_ref;
      t.damping[p] = c.damping_coeff * std::pow(ratio, (double)c.damping_order);
      t.damping_vor[p] = c_v * std::pow(ratio, (double)o_v);
      t.damping_div[p] = c_d * std::pow(ratio, (double)o_d);
      t.eddy_sponge[p] = c.eddy_sponge_coeff * t.eigen[p];
      double eig0n = (double)n * (n + 1.0) / (a * a);       // eigen(0, n)
      t.zmu_sponge[p] = c.zmu_sponge_coeff * eig0n;
      t.zmv_sponge[p] = c.zmv_sponge_coeff * eig0n;
    }
  }

  // ---- semi-implicit reference operators (model/implicit.F90:79-214)
  {
    ImplicitRef r;
    r.K = K; r.rdgas = c.rdgas; r.kappa = c.rdgas / (c.rdgas / c.kappa); r.ref_ps = c.reference_sea_level_press;
    r.pk = &t.pk; r.bk = &t.bk; r.dpk = &t.dpk; r.dbk = &t.dbk;
    r.t.assign(K, 300.);                              // spectral_dynamics.F90:473
    std::vector<double> ph, pf;
    pressure_variables_1d(t.pk, t.bk, r.ref_ps, ph, r.lh, pf, r.lf);
    std::vector<double> dlh(K + 1), dlf(K), l1h, l1f, l2h, l2f;
    for (int k = 1; k <= K; ++k) dlh[k] = t.bk[k] / (t.pk[k] + t.bk[k] * r.ref_ps);
    if (t.pk[0] == 0.0) dlh[0] = 1.0 / r.ref_ps; else dlh[0] = t.bk[0] / (t.pk[0] + t.bk[0] * r.ref_ps);
    const double eps = 1.e-5;
    pressure_variables_1d(t.pk, t.bk, r.ref_ps * (1.0 - 0.5 * eps), ph, l1h, pf, l1f);
    pressure_variables_1d(t.pk, t.bk, r.ref_ps * (1.0 + 0.5 * eps), ph, l2h, pf, l2f);
    for (int k = 0; k < K; ++k) dlf[k] = (l2f[k] - l1f[k]) / (eps * r.ref_ps);
    // build_matrix
    std::vector<double> tau(K * K, 0.0), gamma(K * K, 0.0), nu(K, 0.0), zero(K, 0.0), zero1(K + 1, 0.0);
    for (int k = 0; k < K; ++k) {
      std::vector<double> in(K, 0.0), dt_t, gp; in[k] = 1.0; double dt_p;
      r.linear_tp_tendency(in, dt_p, dt_t);
      nu[k] = -dt_p;
      for (int kk = 0; kk < K; ++kk) tau[kk * K + k] = -dt_t[kk];
      r.linear_geopotential(in, zero1, zero, gp);
      for (int kk = 0; kk < K; ++kk) gamma[kk * K + k] = gp[kk];
    }
    std::vector<double> h1(K), h2;
    for (int k = 0; k < K; ++k) {                      // pres_grad_funct :389-410
      double dlog_1 = r.lh[k + 1] - r.lf[k], dlog_2 = r.lf[k] - r.lh[k];
      h1[k] = c.rdgas * r.t[k] * (t.bk[k + 1] * dlog_1 + t.bk[k] * dlog_2) / (t.dpk[k] + t.dbk[k] * r.ref_ps);
    }
    r.linear_geopotential(zero, dlh, dlf, h2);
    t.h.resize(K);
    for (int k = 0; k < K; ++k) t.h[k] = h1[k] + h2[k];
    t.div_mat.assign(K * K, 0.0);
    for (int k = 0; k < K; ++k)
      for (int kk = 0; kk < K; ++kk) {
        double s = t.h[k] * nu[kk];
        for (int kkk = 0; kkk < K; ++kkk) s = s + gamma[k * K + kkk] * tau[kkk * K + kk];
        t.div_mat[k * K + kk] = s;
      }
    t.ref_ln_p_half = r.lh; t.ref_ln_p_full = r.lf; t.ref_t = r.t; t.ref_ps = r.ref_ps;
  }

  // ---- finite-volume tracer grid: fv_advection_init (model/fv_advection.F90:59-121) with the latitude
  //      boundaries of transforms_init (tools/transforms.F90:314-321)
  {
    std::vector<double> yy(J + 1), y(J);
    yy[0] = -.5 * PI;
    double sum_wts = 0.;
    for (int j = 1; j <= J - 1; ++j) { sum_wts = sum_wts + t.wts_lat[j - 1]; yy[j] = std::asin(sum_wts - 1.); }
    yy[J] = .5 * PI;
    t.fv_c.resize(J); t.fv_cc.resize(J + 1); t.fv_dy.assign(J + 4, 0.0); t.fv_dyy.assign(J + 1, 0.0);
    t.fv_dy_plus.resize(J + 2); t.fv_dy_minus.resize(J + 2);
    for (int j = 0; j < J; ++j) { y[j] = 0.5 * (yy[j + 1] + yy[j]); t.fv_c[j] = std::cos(y[j]); }
    for (int j = 0; j <= J; ++j) t.fv_cc[j] = std::cos(yy[j]);
    std::vector<double>& dy = t.fv_dy;                      // dy[j+1] = dy(j), j = -1..J+2
    for (int j = 1; j <= J; ++j) dy[j + 1] = yy[j] - yy[j - 1];
    dy[0] = dy[3]; dy[1] = dy[2]; dy[J + 2] = dy[J + 1]; dy[J + 3] = dy[J];
    for (int j = 2; j <= J; ++j) t.fv_dyy[j - 1] = y[j - 1] - y[j - 2];
    t.fv_dyy[0] = 2 * (y[0] - yy[0]);
    t.fv_dyy[J] = 2 * (yy[J] - y[J - 1]);
    for (int j = 0; j <= J + 1; ++j) {
      t.fv_dy_plus[j] = dy[j + 1] / (dy[j + 1] + dy[j + 2]);
      t.fv_dy_minus[j] = dy[j + 1] / (dy[j] + dy[j + 1]);
    }
    for (auto& v : t.fv_dy) v = v * c.radius;
    for (auto& v : t.fv_dyy) v = v * c.radius;
    t.fv_dx = 2.0 * PI * c.radius / (double)I;
  }

  // ---- FFT twiddles exp(-2 pi i k / I), k < I, from long double for full fp64 accuracy
  t.twiddle.resize((size_t)2 * I);
  for (int k = 0; k < I; ++k) {
    long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)I;
    t.twiddle[2 * k] = (double)cosl(ang);
    t.twiddle[2 * k + 1] = (double)sinl(ang);
  }
}

void build_wave_matrices(const IscaConfig& c, const Geometry& g, const HostTables& t, double xi,
                         std::vector<double>& wm) {
  const int K = g.K, nw = g.N - 1;
  wm.assign((size_t)(nw + 1) * K * K, 0.0);
  std::vector<double> mat(K * K);
  for (int L = 0; L <= nw; ++L) {
    double factor = xi * xi * L * (L + 1) / (c.radius * c.radius);
    for (int k = 0; k < K; ++k)
      for (int kk = 0; kk < K; ++kk) mat[k * K + kk] = (k == kk ? 1.0 : 0.0) + factor * t.div_mat[k * K + kk];
    invert_matrix(mat, K);
    std::copy(mat.begin(), mat.end(), wm.begin() + (size_t)L * K * K);
  }
}

}  // namespace isca
