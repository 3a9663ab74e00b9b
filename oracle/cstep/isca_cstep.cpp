// isca_cstep.cpp -- CPU restatement (C++17 + OpenMP) of one Isca time step: atmosphere(Time) of the dry spectral dynamical core with
// Held-Suarez forcing and the sphum grid tracer.  TEST INFRASTRUCTURE / CPU BASELINE ONLY: it is the second, independent checker beside
// the NumPy oracle (tests/test_cstep.py holds the two against each other) and the multi-core CPU arm of bench.py (`--impl reference`,
// `cpu_baseline`).  Nothing under isca_b200/ includes, links or calls it.
//
// Follows, routine by routine (paths relative to /root/reference/src):
//   atmos_spectral/tools/spherical_fourier.F90:177-339   Legendre transforms with the reference's RECTANGULAR n range (0..num_spherical
//                                                        for every m; the GPU path computes the triangle only)
//   shared/fft/fft.F90:483-718, fft99.F90:195-209        real FFT along longitude (forward scaled 1/N, inverse unscaled)
//   atmos_spectral/tools/spherical.F90:270-600           spectral operators, triangular truncation
//   atmos_spectral/tools/transforms.F90:379-533,599-831  compositions, divide_by_cos, area_weighted_global_mean (:1059-1077)
//   atmos_spectral/model/press_and_geopot.F90:116-387    pressure variables (Simmons-Burridge), geopotential
//   atmos_spectral/model/spectral_dynamics.F90:780-1338  spectral_dynamics, four_in_one, update_tracers, corrections
//   atmos_spectral/model/implicit.F90:241-480            implicit_correction (matrices are built by the caller, see oracle/cstep.py)
//   atmos_spectral/model/spectral_damping.F90:172-291, leapfrog.F90:58-105
//   atmos_shared/vert_advection/vert_advection.F90:70-478  second_centered and finite_volume_parabolic
//   atmos_spectral/model/fv_advection.F90:126-560        Lin-Rood A-grid van-Leer advection, polar mirror rows
//   atmos_param/hs_forcing/hs_forcing.F90:148-272,508-724  Held-Suarez forcing
// Init-time tables (Gaussian grid, Legendre functions, coefficient tables, vertical coordinate, semi-implicit matrices, damping
// coefficients, finite-volume grid metrics) are handed in by the caller; the timed step is entirely in this file.
// Parallelism: OpenMP over levels / latitudes / spectral rows (the reference uses one MPI rank per core over latitudes and zonal
// wavenumbers); arithmetic order per output element is that of the reference loops, so results do not depend on the thread count.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
typedef std::complex<double> cplx;

struct Params {                      // plain C struct, mirrored by oracle/cstep.py
  int I, J, K, M, N;
  int num_tracers, do_mass, do_energy, do_water, use_implicit, no_forcing, do_conserve_energy;
  double dt_atmos, robert_coeff, raw_filter_coeff, tracer_robert_coeff, radius, grav, rdgas, kappa, cp_air;
  double water_correction_limit, valid_t_lo, valid_t_hi, ref_ps;
  double t_zero, t_strat, delh, delv, eps, sigma_b, P00, tka, tks, vkf, trflux, trsink;
  double fv_dx, alpha_implicit;
};
struct Tables {                      // pointers into caller-owned arrays (copied at create)
  const double *legendre, *legendre_wts;                       // [J/2][N+1][M+1]
  const double *cosm_lat, *wts_lat, *rad_lat, *coriolis;       // [J]
  const double *triangle_mask, *eigen, *coef_uvm, *coef_uvc, *coef_uvp, *coef_alpm, *coef_alpp, *coef_dym, *coef_dx, *coef_dyp;   // [N+1][M+1]
  const double *pk, *bk;                                       // [K+1]
  const double *damping, *damping_vor, *damping_div;           // [N+1][M+1]
  const double *ref_t, *ref_ln_p_half, *ref_ln_p_full, *h;     // [K], [K+1], [K], [K]
  const double *wave_dt, *wave_2dt;                            // [N][K][K]  (total wavenumber 0..N-1)
  const double *fv_c, *fv_cc, *fv_dy, *fv_dyy, *fv_dy_plus, *fv_dy_minus;   // [J], [J+1], [J+4], [J+1], [J+2], [J+2]
};

struct Core {
  Params p;
  std::vector<double> leg, legw, cosm, wts, radlat, cor, tri, eigen, uvm, uvc, uvp, alpm, alpp, dym, cdx, dyp, pk, bk, dpk, dbk, damp, dampv, dampd,
      ref_t, ref_lh, ref_lf, hvec, wave_dt, wave_2dt, fvc, fvcc, fvdy, fvdyy, fvdyp, fvdym;
  // state
  std::vector<cplx> vors, divs, ts, ln_ps;                    // [2][K][NS], [2][NS]
  std::vector<double> ug, vg, tg, psg, vorg, divg, q, wg_full, surf_geo, p_half, p_full, z_half, z_full;
  int previous = 0, current = 0;
  double mean_ps_prev = 0, mean_e_prev = 0, mean_w_prev = 0;
  // FFT
  int H = 0, logH = 0;
  std::vector<cplx> tw_fft, tw_real;                          // e^{-2 pi i k / H}, e^{-2 pi i k / I}
  std::vector<int> bitrev;
  std::string err;

  size_t NS() const { return (size_t)(p.N + 1) * (p.M + 1); }
  size_t JI() const { return (size_t)p.J * p.I; }

  // ------------------------------------------------------------------ FFT (power-of-two lon_max)
  void fft_init() {
    H = p.I / 2; logH = 0;
    while ((1 << logH) < H) ++logH;
    tw_fft.resize(H); tw_real.resize(H + 1); bitrev.resize(H);
    const double PI = 3.14159265358979323846;
    for (int k = 0; k < H; ++k) tw_fft[k] = cplx(cos(2 * PI * k / H), -sin(2 * PI * k / H));
    for (int k = 0; k <= H; ++k) tw_real[k] = cplx(cos(2 * PI * k / p.I), -sin(2 * PI * k / p.I));
    for (int i = 0; i < H; ++i) { int r = 0; for (int b = 0; b < logH; ++b) if (i & (1 << b)) r |= 1 << (logH - 1 - b); bitrev[i] = r; }
  }
  // in-place complex FFT of length H; sign -1 forward, +1 inverse (unscaled)
  void cfft(cplx* z, int sign) const {
    for (int i = 0; i < H; ++i) if (bitrev[i] > i) std::swap(z[i], z[bitrev[i]]);
    for (int len = 2; len <= H; len <<= 1) {
      const int half = len >> 1, step = H / len;
      for (int s = 0; s < H; s += len)
        for (int k = 0; k < half; ++k) {
          cplx w = tw_fft[k * step];
          if (sign > 0) w = std::conj(w);
          cplx a = z[s + k], b = z[s + k + half] * w;
          z[s + k] = a + b; z[s + k + half] = a - b;
        }
    }
  }
  // one longitude line: real x[I] -> c[0..M] = (1/I) sum_j x_j e^{-2 pi i j k / I}
  void fft_fwd_line(const double* x, cplx* c, cplx* work) const {
    for (int n = 0; n < H; ++n) work[n] = cplx(x[2 * n], x[2 * n + 1]);
    cfft(work, -1);
    const double s = 1.0 / p.I;
    for (int k = 0; k <= p.M; ++k) {
      cplx zk = work[k % H], zc = std::conj(work[(H - k) % H]);
      cplx e = 0.5 * (zk + zc), o = cplx(0.0, -0.5) * (zk - zc);
      c[k] = s * (e + tw_real[k] * o);
    }
  }
  // c[0..M] (zero beyond) -> real x[I] = sum_{k=-I/2..I/2} c_k e^{2 pi i j k / I}
  void fft_inv_line(const cplx* c, double* x, cplx* work) const {
    const int M = p.M;
    for (int k = 0; k < H; ++k) {
      cplx ck = k <= M ? c[k] : cplx(0, 0);
      const int kk = H - k;
      cplx cm = std::conj(kk <= M ? c[kk] : cplx(0, 0));
      if (k == 0) cm = std::conj(H <= M ? c[H] : cplx(0, 0));
      cplx e = ck + cm, o = (ck - cm) * std::conj(tw_real[k]);
      work[k] = e + cplx(0.0, 1.0) * o;
    }
    cfft(work, +1);
    for (int n = 0; n < H; ++n) { x[2 * n] = work[n].real(); x[2 * n + 1] = work[n].imag(); }
  }

  // ------------------------------------------------------------------ Legendre (spherical_fourier.F90:177-339), rectangular n range
  // spec [nlev][N+1][M+1] -> four [nlev][J][M+1]
  void spherical_to_fourier(const cplx* spec, cplx* four, int nlev) const {
    const int M1 = p.M + 1, N1 = p.N + 1, J = p.J, JH = J / 2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nlev; ++k)
      for (int j = 0; j < JH; ++j) {
        cplx xe[1024], xo[1024];
        for (int m = 0; m < M1; ++m) { xe[m] = 0; xo[m] = 0; }
        const double* P = &leg[(size_t)j * N1 * M1];
        const cplx* S = spec + (size_t)k * N1 * M1;
        for (int n = 0; n < N1; n += 2) for (int m = 0; m < M1; ++m) xe[m] += S[n * M1 + m] * P[n * M1 + m];
        for (int n = 1; n < N1; n += 2) for (int m = 0; m < M1; ++m) xo[m] += S[n * M1 + m] * P[n * M1 + m];
        cplx* south = four + ((size_t)k * J + j) * M1;
        cplx* north = four + ((size_t)k * J + (J - 1 - j)) * M1;
        for (int m = 0; m < M1; ++m) { south[m] = xe[m] - xo[m]; north[m] = xe[m] + xo[m]; }
      }
  }
  void fourier_to_spherical(const cplx* four, cplx* spec, int nlev) const {
    const int M1 = p.M + 1, N1 = p.N + 1, J = p.J, JH = J / 2;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nlev; ++k) {
      cplx* S = spec + (size_t)k * N1 * M1;
      for (size_t i = 0; i < (size_t)N1 * M1; ++i) S[i] = 0;
      cplx xe[1024], xo[1024];
      for (int j = 0; j < JH; ++j) {
        const cplx* south = four + ((size_t)k * J + j) * M1;
        const cplx* north = four + ((size_t)k * J + (J - 1 - j)) * M1;
        for (int m = 0; m < M1; ++m) { xe[m] = north[m] + south[m]; xo[m] = north[m] - south[m]; }
        const double* W = &legw[(size_t)j * N1 * M1];
        for (int n = 0; n < N1; n += 2) for (int m = 0; m < M1; ++m) S[n * M1 + m] += xe[m] * W[n * M1 + m];
        for (int n = 1; n < N1; n += 2) for (int m = 0; m < M1; ++m) S[n * M1 + m] += xo[m] * W[n * M1 + m];
      }
    }
  }
  void spherical_to_grid(const cplx* spec, double* grid, int nlev) const {
    const int M1 = p.M + 1, J = p.J, I = p.I;
    std::vector<cplx> four((size_t)nlev * J * M1);
    spherical_to_fourier(spec, four.data(), nlev);
#pragma omp parallel
    {
      std::vector<cplx> work(H);
#pragma omp for schedule(static)
      for (long l = 0; l < (long)nlev * J; ++l) fft_inv_line(&four[(size_t)l * M1], grid + (size_t)l * I, work.data());
    }
  }
  void grid_to_spherical(const double* grid, cplx* spec, int nlev, bool trunc) const {
    const int M1 = p.M + 1, J = p.J, I = p.I;
    std::vector<cplx> four((size_t)nlev * J * M1);
#pragma omp parallel
    {
      std::vector<cplx> work(H);
#pragma omp for schedule(static)
      for (long l = 0; l < (long)nlev * J; ++l) fft_fwd_line(grid + (size_t)l * I, &four[(size_t)l * M1], work.data());
    }
    fourier_to_spherical(four.data(), spec, nlev);
    if (trunc) {
      const size_t ns = NS();
#pragma omp parallel for schedule(static)
      for (int k = 0; k < nlev; ++k) for (size_t i = 0; i < ns; ++i) spec[k * ns + i] *= tri[i];
    }
  }
  void divide_by_cos(double* g, int nlev) const {
    const int J = p.J, I = p.I;
#pragma omp parallel for schedule(static)
    for (long l = 0; l < (long)nlev * J; ++l) { const double c = cosm[l % J]; double* r = g + (size_t)l * I; for (int i = 0; i < I; ++i) r[i] *= c; }
  }
  // ------------------------------------------------------------------ spectral operators (spherical.F90)
  static cplx times_i(cplx s) { return cplx(-s.imag(), s.real()); }
  void lon_deriv_cos(const cplx* s, cplx* d, int nlev) const {
    const size_t ns = NS();
    for (int k = 0; k < nlev; ++k) for (size_t i = 0; i < ns; ++i) d[k * ns + i] = cdx[i] * times_i(s[k * ns + i]);
  }
  void lat_deriv_cos(const cplx* s, cplx* d, int nlev) const {
    const int M1 = p.M + 1, N = p.N; const size_t ns = NS();
    for (int k = 0; k < nlev; ++k) {
      const cplx* a = s + k * ns; cplx* o = d + k * ns;
      for (size_t i = 0; i < ns; ++i) o[i] = 0;
      for (int n = 1; n <= N; ++n) for (int m = 0; m < M1; ++m) o[n * M1 + m] = -a[(n - 1) * M1 + m] * dym[n * M1 + m];
      for (int n = 0; n < N; ++n) for (int m = 0; m < M1; ++m) o[n * M1 + m] += a[(n + 1) * M1 + m] * dyp[n * M1 + m];
    }
  }
  void ucos_vcos(const cplx* vor, const cplx* div, cplx* u, cplx* v, int nlev) const {
    const int M1 = p.M + 1, N = p.N; const size_t ns = NS();
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nlev; ++k) {
      const cplx* vo = vor + k * ns; const cplx* di = div + k * ns; cplx* uu = u + k * ns; cplx* vv = v + k * ns;
      for (size_t i = 0; i < ns; ++i) { uu[i] = uvc[i] * times_i(di[i]); vv[i] = uvc[i] * times_i(vo[i]); }
      for (int n = 1; n <= N; ++n) for (int m = 0; m < M1; ++m) {
        uu[n * M1 + m] += uvm[n * M1 + m] * vo[(n - 1) * M1 + m];
        vv[n * M1 + m] -= uvm[n * M1 + m] * di[(n - 1) * M1 + m];
      }
      for (int n = 0; n < N; ++n) for (int m = 0; m < M1; ++m) {
        uu[n * M1 + m] -= uvp[n * M1 + m] * vo[(n + 1) * M1 + m];
        vv[n * M1 + m] += uvp[n * M1 + m] * di[(n + 1) * M1 + m];
      }
    }
  }
  void alpha_operator(const cplx* a, const cplx* b, double isign, cplx* al, int nlev) const {
    const int M1 = p.M + 1, N = p.N; const size_t ns = NS();
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nlev; ++k) {
      const cplx* aa = a + k * ns; const cplx* bb = b + k * ns; cplx* o = al + k * ns;
      for (size_t i = 0; i < ns; ++i) o[i] = cdx[i] * times_i(aa[i]);
      for (int n = 1; n <= N; ++n) for (int m = 0; m < M1; ++m) o[n * M1 + m] -= isign * alpm[n * M1 + m] * bb[(n - 1) * M1 + m];
      for (int n = 0; n < N; ++n) for (int m = 0; m < M1; ++m) o[n * M1 + m] += isign * alpp[n * M1 + m] * bb[(n + 1) * M1 + m];
    }
  }
  void uv_grid_from_vor_div(const cplx* vor, const cplx* div, double* u, double* v, int nlev) const {
    std::vector<cplx> us(nlev * NS()), vs(nlev * NS());
    ucos_vcos(vor, div, us.data(), vs.data(), nlev);
    spherical_to_grid(us.data(), u, nlev); divide_by_cos(u, nlev);
    spherical_to_grid(vs.data(), v, nlev); divide_by_cos(v, nlev);
  }
  void vor_div_from_uv_grid(const double* u, const double* v, cplx* vor, cplx* div, int nlev) const {
    const size_t n3 = (size_t)nlev * JI(), ns = NS();
    std::vector<double> tmp(n3);
    std::vector<cplx> dx(nlev * ns), dy(nlev * ns);
    std::copy(u, u + n3, tmp.begin()); divide_by_cos(tmp.data(), nlev); grid_to_spherical(tmp.data(), dx.data(), nlev, false);
    std::copy(v, v + n3, tmp.begin()); divide_by_cos(tmp.data(), nlev); grid_to_spherical(tmp.data(), dy.data(), nlev, false);
    alpha_operator(dy.data(), dx.data(), -1.0, vor, nlev);
    alpha_operator(dx.data(), dy.data(), +1.0, div, nlev);
    for (int k = 0; k < nlev; ++k) for (size_t i = 0; i < ns; ++i) { vor[k * ns + i] *= tri[i]; div[k * ns + i] *= tri[i]; }
  }
  double area_weighted_global_mean(const double* f) const {      // transforms.F90:1059-1077
    const int J = p.J, I = p.I;
    double sum = 0.0, sw = 0.0;
    for (int j = 0; j < J; ++j) { double r = 0.0; for (int i = 0; i < I; ++i) r += wts[j] * f[(size_t)j * I + i]; sum += r; sw += wts[j]; }
    return sum / (sw * I);
  }
  // ------------------------------------------------------------------ pressure / geopotential (press_and_geopot.F90)
  // arrays [k][JI]; ps [JI]
  void pressure_variables(const double* ps, double* ph, double* lnh, double* pf, double* lnf) const {
    const int K = p.K; const size_t n = JI();
    const bool top0 = pk[0] == 0.0 && bk[0] == 0.0;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)n; ++c) {
      for (int k = 0; k <= K; ++k) ph[k * n + c] = pk[k] + bk[k] * ps[c];
      if (top0) {
        for (int k = 1; k <= K; ++k) lnh[k * n + c] = log(ph[k * n + c]);
        for (int k = 1; k < K; ++k) {
          double alpha = 1.0 - ph[k * n + c] * (lnh[(k + 1) * n + c] - lnh[k * n + c]) / (ph[(k + 1) * n + c] - ph[k * n + c]);
          lnf[k * n + c] = lnh[(k + 1) * n + c] - alpha;
        }
        lnf[c] = lnh[n + c] + (-1.0);
        lnh[c] = 0.0;
      } else {
        for (int k = 0; k <= K; ++k) lnh[k * n + c] = log(ph[k * n + c]);
        for (int k = 0; k < K; ++k) {
          double alpha = 1.0 - ph[k * n + c] * (lnh[(k + 1) * n + c] - lnh[k * n + c]) / (ph[(k + 1) * n + c] - ph[k * n + c]);
          lnf[k * n + c] = lnh[(k + 1) * n + c] - alpha;
        }
      }
      for (int k = 0; k < K; ++k) pf[k * n + c] = exp(lnf[k * n + c]);
    }
  }
  void compute_geopotential(const double* t, const double* lnh, const double* lnf, double* gf, double* gh) const {
    const int K = p.K; const size_t n = JI();
    const int ktop = pk[0] == 0.0 ? 1 : 0;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)n; ++c) {
      for (int k = 0; k <= K; ++k) gh[k * n + c] = 0.0;
      gh[K * n + c] = surf_geo[c];
      for (int k = K - 1; k >= ktop; --k) gh[k * n + c] = gh[(k + 1) * n + c] + p.rdgas * t[k * n + c] * (lnh[(k + 1) * n + c] - lnh[k * n + c]);
      for (int k = 0; k < K; ++k) gf[k * n + c] = gh[(k + 1) * n + c] + p.rdgas * t[k * n + c] * (lnh[(k + 1) * n + c] - lnf[k * n + c]);
    }
  }
  void pressures_and_heights(int lev) {
    const int K = p.K; const size_t n = JI();
    std::vector<double> lnh((K + 1) * n), lnf(K * n);
    double* ph = &p_half[(size_t)lev * (K + 1) * n]; double* pf = &p_full[(size_t)lev * K * n];
    double* zh = &z_half[(size_t)lev * (K + 1) * n]; double* zf = &z_full[(size_t)lev * K * n];
    pressure_variables(&psg[lev * n], ph, lnh.data(), pf, lnf.data());
    compute_geopotential(&tg[(size_t)lev * K * n], lnh.data(), lnf.data(), zf, zh);
    const double gi = p.grav;
    for (size_t i = 0; i < (size_t)K * n; ++i) zf[i] /= gi;
    for (size_t i = 0; i < (size_t)(K + 1) * n; ++i) zh[i] /= gi;
  }
  double mass_weighted_global_integral(const double* field, const double* ps) const {   // global_integral.F90:49-81
    const int K = p.K; const size_t n = JI();
    std::vector<double> vi(n, 0.0);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)n; ++c) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) { double dp = (pk[k + 1] + bk[k + 1] * ps[c]) - (pk[k] + bk[k] * ps[c]); s = s + field[k * n + c] * dp; }
      vi[c] = s;
    }
    return area_weighted_global_mean(vi.data()) / p.grav;
  }
  // ------------------------------------------------------------------ vertical advection (vert_advection.F90), ADVECTIVE_FORM
  // second_centered on arrays [k][JI]; out += tendency
  void vert_adv_centered(const double* w, const double* dz, const double* r, double* out) const {
    const int K = p.K; const size_t n = JI();
#pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)n; ++c) {
      double flux_up = w[c] * r[c];                         // flux(ks) = w(ks) * r(ks)
      for (int k = 0; k < K; ++k) {
        double flux_dn = k == K - 1 ? w[(size_t)K * n + c] * r[(size_t)(K - 1) * n + c]
                                    : w[(size_t)(k + 1) * n + c] * (0.5 * (r[(size_t)(k + 1) * n + c] + r[(size_t)k * n + c]));
        out[k * n + c] += -(flux_dn - flux_up - r[k * n + c] * (w[(size_t)(k + 1) * n + c] - w[k * n + c])) / dz[k * n + c];
        flux_up = flux_dn;
      }
    }
  }
  // finite_volume_parabolic of one column (arrays of length K / K+1); returns the tendency in rdt
  static void ppm_column(int K, double dt, const double* w, const double* dz, const double* r, double* rdt) {
    double slp[256], rl[256], rr[256], zwt1[256], zwt2[256], zwt3[256], flux[257];
    // slope_z (linear = .false., limit = .true.)
    for (int k = 0; k < K; ++k) slp[k] = 0.0;
    {
      double grad[256];
      grad[0] = 0.0;
      for (int k = 1; k < K; ++k) grad[k] = (r[k] - r[k - 1]) / (dz[k] + dz[k - 1]);
      for (int k = 1; k < K - 1; ++k)
        slp[k] = (grad[k + 1] * (2.0 * dz[k - 1] + dz[k]) + grad[k] * (2.0 * dz[k + 1] + dz[k])) * dz[k] / (dz[k - 1] + dz[k] + dz[k + 1]);
      for (int k = 1; k < K - 1; ++k) {
        double rmin = std::min(std::min(r[k - 1], r[k]), r[k + 1]), rmax = std::max(std::max(r[k - 1], r[k]), r[k + 1]);
        double sg = slp[k] >= 0.0 ? 1.0 : -1.0;
        slp[k] = sg * std::min(std::min(fabs(slp[k]), 2.0 * (r[k] - rmin)), 2.0 * (rmax - r[k]));
      }
      slp[0] = 0.0; slp[K - 1] = 0.0;
    }
    for (int k = 2; k < K - 1; ++k) {                       // compute_weights
      double denom1 = 1.0 / (dz[k - 1] + dz[k]), denom2 = 1.0 / (dz[k - 2] + dz[k - 1] + dz[k] + dz[k + 1]);
      double denom3 = 1.0 / (2 * dz[k - 1] + dz[k]), denom4 = 1.0 / (dz[k - 1] + 2 * dz[k]);
      double num3 = dz[k - 2] + dz[k - 1], num4 = dz[k] + dz[k + 1];
      double x = num3 * denom3 - num4 * denom4, y = 2.0 * dz[k - 1] * dz[k];
      double z0 = dz[k - 1] * denom1;
      zwt1[k] = z0 + x * y * denom1 * denom2;
      zwt2[k] = dz[k - 1] * num3 * denom3 * denom2;
      zwt3[k] = dz[k] * num4 * denom4 * denom2;
    }
    for (int k = 0; k < K; ++k) { rl[k] = 0.0; rr[k] = 0.0; }
    for (int k = 2; k < K - 1; ++k) {
      rl[k] = r[k - 1] + zwt1[k] * (r[k] - r[k - 1]) - zwt2[k] * slp[k] + zwt3[k] * slp[k - 1];
      rr[k - 1] = rl[k];
    }
    rl[1] = r[1] - 0.5 * slp[1];
    rr[K - 2] = r[K - 2] + 0.5 * slp[K - 2];
    rl[0] = r[0] - 0.5 * slp[0]; rr[0] = r[0] + 0.5 * slp[0];
    rl[K - 1] = r[K - 1] - 0.5 * slp[K - 1]; rr[K - 1] = r[K - 1] + 0.5 * slp[K - 1];
    for (int k = 0; k < K; ++k) {                           // Colella-Woodward limiter (:340-356)
      if ((rr[k] - r[k]) * (r[k] - rl[k]) <= 0.0) { rl[k] = r[k]; rr[k] = r[k]; }
      if (k == 0 || k == K - 1) continue;
      double rm = rr[k] - rl[k], a = rm * (r[k] - 0.5 * (rr[k] + rl[k])), b = rm * rm / 6.0;
      if (a > b) rl[k] = 3.0 * r[k] - 2.0 * rr[k];
      if (a < -b) rr[k] = 3.0 * r[k] - 2.0 * rl[k];
    }
    const double tt = 2.0 / 3.0;
    flux[0] = w[0] * r[0]; flux[K] = w[K] * r[K - 1];
    for (int k = 1; k < K; ++k) {
      const double wk = w[k];
      const bool pos = wk >= 0.0;
      double cn = pos ? dt * wk / dz[k - 1] : -dt * wk / dz[k];
      int kk = pos ? k - 1 : k;
      double rsum = 0.0, dzsum = 0.0;
      const double dtw = pos ? dt * wk : -dt * wk;
      const bool big = cn > 1.0;
      if (big) {
        const int step = pos ? -1 : 1;
        while (dzsum + dz[kk] < dtw) {
          if (kk == 0) break;
          dzsum += dz[kk]; rsum += r[kk]; kk += step;
          if (kk >= K) { kk = K - 1; break; }
        }
      }
      double xx = big ? (dtw - dzsum) / dz[kk] : cn;
      double rm = rr[kk] - rl[kk], r6 = 6.0 * (r[kk] - 0.5 * (rr[kk] + rl[kk]));
      if (pos && kk == 0) r6 = 0.0;
      if (!pos && kk == K - 1) r6 = 0.0;
      double rst = pos ? rr[kk] - 0.5 * xx * (rm - (1.0 - tt * xx) * r6) : rl[kk] + 0.5 * xx * (rm + (1.0 - tt * xx) * r6);
      if (big) rst = (xx * rst + rsum) / cn;
      flux[k] = wk * rst;
    }
    for (int k = 0; k < K; ++k) rdt[k] = -(flux[k + 1] - flux[k] - r[k] * (w[k + 1] - w[k])) / dz[k];
  }
  // ------------------------------------------------------------------ fv_advection.F90 for one level: q, ua, va [J][I] -> dq_dt [J][I]
  void fv_level(const double* ua, const double* va, const double* q, double dt, double* dq) const {
    const int nx = p.I, ny = p.J, hx = nx / 2;
    auto DY = [&](int j) { return fvdy[j + 1]; };           // Fortran dy(j), j = -1..ny+2
    // rows with halo: index jj = j + 1 for Fortran j = -1..ny+2  (2-row halo), polar rows mirrored at the antipodal longitude
    std::vector<double> qx((size_t)(ny + 4) * nx), q1x((size_t)(ny + 4) * nx), vx((size_t)(ny + 2) * nx), q2((size_t)ny * nx),
        uc((size_t)ny * nx), vc((size_t)(ny + 1) * nx), sx((size_t)ny * nx), fl((size_t)ny * nx), sl((size_t)(ny + 2) * nx), fy((size_t)(ny + 1) * nx);
    auto halo2 = [&](const double* src, double* dst, double sign) {
      for (int j = 1; j <= ny; ++j) std::copy(src + (size_t)(j - 1) * nx, src + (size_t)j * nx, dst + (size_t)(j + 1) * nx);
      for (int r = 0; r < 2; ++r)
        for (int i = 0; i < nx; ++i) {
          dst[(size_t)(1 - r) * nx + i] = sign * src[(size_t)r * nx + (i + hx) % nx];                     // j = 0, -1  <- rows 1, 2
          dst[(size_t)(ny + 2 + r) * nx + i] = sign * src[(size_t)(ny - 1 - r) * nx + (i + hx) % nx];      // j = ny+1, ny+2 <- rows ny, ny-1
        }
    };
    halo2(q, qx.data(), 1.0);
    for (int j = 1; j <= ny; ++j) std::copy(va + (size_t)(j - 1) * nx, va + (size_t)j * nx, vx.begin() + (size_t)j * nx);
    for (int i = 0; i < nx; ++i) { vx[i] = -va[(i + hx) % nx]; vx[(size_t)(ny + 1) * nx + i] = -va[(size_t)(ny - 1) * nx + (i + hx) % nx]; }
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) uc[(size_t)j * nx + i] = 0.5 * (ua[(size_t)j * nx + (i + nx - 1) % nx] + ua[(size_t)j * nx + i]);
    for (int j = 1; j <= ny + 1; ++j) for (int i = 0; i < nx; ++i) vc[(size_t)(j - 1) * nx + i] = 0.5 * (vx[(size_t)(j - 1) * nx + i] + vx[(size_t)j * nx + i]);
    // dq_dt = q * div
    for (int j = 1; j <= ny; ++j)
      for (int i = 0; i < nx; ++i) {
        double d = (vc[(size_t)j * nx + i] * fvcc[j] - vc[(size_t)(j - 1) * nx + i] * fvcc[j - 1]) / (fvc[j - 1] * DY(j));
        d = d + (uc[(size_t)(j - 1) * nx + (i + 1) % nx] - uc[(size_t)(j - 1) * nx + i]) / (fvc[j - 1] * p.fv_dx);
        dq[(size_t)(j - 1) * nx + i] = q[(size_t)(j - 1) * nx + i] * d;
      }
    auto find_cell = [&](int i1, double b) {                // 1-based source cell ii = i-1 - floor(b), wrapped
      long ii = (long)(i1 - 1) - (long)floor(b);
      if (ii > nx) ii -= nx;
      if (ii < 1) ii += nx;
      return (int)ii;
    };
    // q1 = q + semi_x(q, dt/2) (with halo), q2 = q + semi_y(q, dt/2)
    {
      std::vector<double> q1((size_t)ny * nx);
      const double hdt = 0.5 * dt;
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          double b = ua[(size_t)j * nx + i] * hdt / (p.fv_dx * fvc[j]);
          int il = find_cell(i + 1, b), ir = il + 1 > nx ? 1 : il + 1;
          double bb = b - floor(b);
          q1[(size_t)j * nx + i] = q[(size_t)j * nx + i] + (bb * q[(size_t)j * nx + il - 1] + (1.0 - bb) * q[(size_t)j * nx + ir - 1] - q[(size_t)j * nx + i]);
        }
      halo2(q1.data(), q1x.data(), 1.0);
      for (int j = 1; j <= ny; ++j)
        for (int i = 0; i < nx; ++i) {
          double v = va[(size_t)(j - 1) * nx + i];
          double up = v * hdt * (qx[(size_t)j * nx + i] - qx[(size_t)(j + 1) * nx + i]) / fvdyy[j - 1];
          double dn = v * hdt * (qx[(size_t)(j + 1) * nx + i] - qx[(size_t)(j + 2) * nx + i]) / fvdyy[j];
          q2[(size_t)(j - 1) * nx + i] = q[(size_t)(j - 1) * nx + i] + (v >= 0.0 ? up : dn);
        }
    }
    // vanleer_x on q2
    for (int j = 0; j < ny; ++j) {
      const double* qq = &q2[(size_t)j * nx];
      double bmax = 0.0;
      for (int i = 0; i < nx; ++i) {
        double qm = qq[(i + nx - 1) % nx], qp = qq[(i + 1) % nx], q0 = qq[i];
        double slope = ((qp - q0) + (q0 - qm)) / 2;
        double qmin = std::min(std::min(qm, q0), qp), qmax = std::max(std::max(qm, q0), qp);
        double sg = slope >= 0.0 ? 1.0 : -1.0;
        sx[(size_t)j * nx + i] = sg * std::min(std::min(fabs(slope), 2.0 * (q0 - qmin)), 2.0 * (qmax - q0));
        bmax = std::max(bmax, fabs(uc[(size_t)j * nx + i] * dt / (p.fv_dx * fvc[j])));
      }
      for (int i = 0; i < nx; ++i) {
        double b = uc[(size_t)j * nx + i] * dt / (p.fv_dx * fvc[j]);
        double f = 0.0;
        if (bmax > 1.0) {                                   // integer_flux_x (:483-521)
          long n = (long)trunc(b); const int i1 = i + 1;
          if (n >= 1) {
            if (i1 - n >= 1) for (long t = i1 - n; t <= i1 - 1; ++t) f += qq[t - 1];
            else { for (long t = 1; t <= i1 - 1; ++t) f += qq[t - 1]; for (long t = i1 - n + nx; t <= nx; ++t) f += qq[t - 1]; }
          } else if (n <= -1) {
            double s = 0.0;
            if (i1 - 1 - n <= nx) for (long t = i1; t <= i1 - 1 - n; ++t) s += qq[t - 1];
            else { for (long t = i1; t <= nx; ++t) s += qq[t - 1]; for (long t = 1; t <= i1 - 1 - n - nx; ++t) s += qq[t - 1]; }
            f = -s;
          }
        }
        double bb = b - trunc(b);
        int ii = find_cell(i + 1, b) - 1;
        double sgb = bb >= 0.0 ? 1.0 : -1.0;
        fl[(size_t)j * nx + i] = f + bb * (qq[ii] + 0.5 * sx[(size_t)j * nx + ii] * (sgb - bb));
      }
      for (int i = 0; i < nx; ++i) dq[(size_t)j * nx + i] = dq[(size_t)j * nx + i] - (fl[(size_t)j * nx + (i + 1) % nx] - fl[(size_t)j * nx + i]) / dt;
    }
    // vanleer_sphere on q1x
    for (int j = 0; j <= ny + 1; ++j)
      for (int i = 0; i < nx; ++i) {
        double qj = q1x[(size_t)(j + 1) * nx + i], qp = q1x[(size_t)(j + 2) * nx + i], qm = q1x[(size_t)j * nx + i];
        double slope = (qp - qj) * fvdyp[j] + (qj - qm) * fvdym[j];
        double qmin = std::min(std::min(qm, qj), qp), qmax = std::max(std::max(qm, qj), qp);
        double sg = slope >= 0.0 ? 1.0 : -1.0;
        sl[(size_t)j * nx + i] = sg * std::min(std::min(fabs(slope), 2.0 * (qj - qmin)), 2.0 * (qmax - qj));
      }
    for (int j = 1; j <= ny + 1; ++j)
      for (int i = 0; i < nx; ++i) {
        double v = vc[(size_t)(j - 1) * nx + i];
        double dtdy_m = dt / DY(j - 1), dtdy = dt / DY(j);
        double fp = v * fvcc[j - 1] * (q1x[(size_t)j * nx + i] + 0.5 * sl[(size_t)(j - 1) * nx + i] * (1.0 - dtdy_m * v));
        double fm = v * fvcc[j - 1] * (q1x[(size_t)(j + 1) * nx + i] - 0.5 * sl[(size_t)j * nx + i] * (1.0 + dtdy * v));
        fy[(size_t)(j - 1) * nx + i] = (j == 1 || j == ny + 1) ? 0.0 : (v >= 0.0 ? fp : fm);
      }
    for (int j = 1; j <= ny; ++j) {
      const double dyc = 1.0 / (DY(j) * fvc[j - 1]);
      for (int i = 0; i < nx; ++i) dq[(size_t)(j - 1) * nx + i] = dq[(size_t)(j - 1) * nx + i] - dyc * (fy[(size_t)j * nx + i] - fy[(size_t)(j - 1) * nx + i]);
    }
  }
  // ------------------------------------------------------------------ implicit.F90: linear operators on one coefficient column
  void linear_tp_tendency(const cplx* div, cplx& dt_p, cplx* dt_t) const {
    const int K = p.K;
    cplx dmean_tot = 0, vert_vel[257], temp[257];
    vert_vel[0] = 0;
    for (int k = 0; k < K; ++k) {
      double dp = dpk[k] + dbk[k] * p.ref_ps, dp_inv = 1 / dp;
      double dlog_1 = ref_lh[k + 1] - ref_lf[k], dlog_3 = ref_lh[k + 1] - ref_lh[k];
      cplx dmean = div[k] * dp;
      dt_t[k] = -p.kappa * ref_t[k] * (dmean_tot * dlog_3 + dmean * dlog_1) * dp_inv;
      dmean_tot = dmean_tot + dmean;
      vert_vel[k + 1] = -dmean_tot;
    }
    dt_p = -dmean_tot;
    for (int k = 1; k < K; ++k) vert_vel[k] = vert_vel[k] + dmean_tot * bk[k];
    for (int k = 0; k <= K; ++k) temp[k] = 0;
    for (int k = 1; k < K; ++k) temp[k] = -vert_vel[k] * (ref_t[k] - ref_t[k - 1]);
    for (int k = 0; k < K; ++k) {
      double dp = dpk[k] + dbk[k] * p.ref_ps, dp_inv = 1 / dp;
      dt_t[k] = dt_t[k] + 0.5 * dp_inv * (temp[k + 1] + temp[k]);
    }
  }
  void linear_geopotential_t(const cplx* del_t, cplx* g) const {   // del_ln_p_half = del_ln_p_full = 0
    const int K = p.K;
    cplx gh[257];
    for (int k = 0; k <= K; ++k) gh[k] = 0;
    for (int k = K - 1; k >= 1; --k) gh[k] = gh[k + 1] + p.rdgas * (del_t[k] * (ref_lh[k + 1] - ref_lh[k]));
    for (int k = 0; k < K; ++k) g[k] = gh[k + 1] + p.rdgas * (del_t[k] * (ref_lh[k + 1] - ref_lf[k]));
  }
  void implicit_correction(cplx* dt_divs, cplx* dt_ts, cplx* dt_ln_ps, double dt_in, int prev, int cur) const {
    const int K = p.K, M1 = p.M + 1, N1 = p.N + 1; const size_t ns = NS();
    const double xi = dt_in * p.alpha_implicit;
    const double* wave = dt_in == p.dt_atmos ? wave_dt.data() : wave_2dt.data();
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N1; ++n)
      for (int m = 0; m < M1; ++m) {
        const size_t i = (size_t)n * M1 + m;
        cplx dv[256], tt[256], tmp[256], geo[256], ddiv[256], work[256];
        cplx dps;
        for (int k = 0; k < K; ++k) dv[k] = divs[((size_t)prev * K + k) * ns + i] - divs[((size_t)cur * K + k) * ns + i];
        linear_tp_tendency(dv, dps, tmp);
        for (int k = 0; k < K; ++k) tt[k] = dt_ts[k * ns + i] + tmp[k];
        cplx dlp = dt_ln_ps[i] + dps / p.ref_ps;
        for (int k = 0; k < K; ++k) tmp[k] = ts[((size_t)prev * K + k) * ns + i] - ts[((size_t)cur * K + k) * ns + i] + xi * tt[k];
        cplx ps_temp = ln_ps[prev * ns + i] - ln_ps[cur * ns + i] + xi * dlp;
        linear_geopotential_t(tmp, geo);
        for (int k = 0; k < K; ++k) ddiv[k] = dt_divs[k * ns + i] + eigen[i] * (geo[k] + hvec[k] * ps_temp * p.ref_ps);
        const int L = n + m;
        if (L <= p.N - 1) {
          const double* W = wave + (size_t)L * K * K;
          for (int k = 0; k < K; ++k) { cplx s = 0; for (int kk = 0; kk < K; ++kk) s += W[k * K + kk] * ddiv[kk]; work[k] = s; }
          for (int k = 0; k < K; ++k) ddiv[k] = work[k];
        }
        linear_tp_tendency(ddiv, dps, tmp);
        for (int k = 0; k < K; ++k) { dt_divs[k * ns + i] = ddiv[k]; dt_ts[k * ns + i] = tt[k] + xi * tmp[k]; }
        dt_ln_ps[i] = dlp + xi * dps / p.ref_ps;
      }
  }

  // ------------------------------------------------------------------ Held-Suarez forcing (hs_forcing.F90)
  void hs_forcing(double dt, int prev, int cur, double* udt, double* vdt, double* tdt, double* rdt) const {
    const int K = p.K, J = p.J, I = p.I; const size_t n = JI();
    if (p.no_forcing) return;
    const double* ph = &p_half[(size_t)cur * (K + 1) * n]; const double* pf = &p_full[(size_t)cur * K * n];
    const double* u = &ug[(size_t)prev * K * n]; const double* v = &vg[(size_t)prev * K * n]; const double* t = &tg[(size_t)prev * K * n];
    const double vcoeff = -p.vkf / (1.0 - p.sigma_b), tcoeff = (p.tks - p.tka) / (1.0 - p.sigma_b);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < J; ++j) {
        const double sl = sin(radlat[j]), sl2 = sl * sl, cl2 = 1.0 - sl2, cl4 = cl2 * cl2;
        const double t_star = p.t_zero - p.delh * sl2 - p.eps * sl, tstr = p.t_strat - p.eps * sl;
        for (int i = 0; i < I; ++i) {
          const size_t c = (size_t)j * I + i, x = (size_t)k * n + c;
          const double ps = ph[(size_t)K * n + c], sigma = pf[x] * (1.0 / ps);
          const bool act = sigma <= 1.0 && sigma > p.sigma_b;
          const double vf = vcoeff * (sigma - p.sigma_b);
          const double ut = act ? vf * u[x] : 0.0, vt = act ? vf * v[x] : 0.0;
          if (p.do_conserve_energy) tdt[x] += -((u[x] + 0.5 * ut * dt) * ut + (v[x] + 0.5 * vt * dt) * vt) / p.cp_air;
          udt[x] += ut; vdt[x] += vt;
          const double p_norm = pf[x] / p.P00;
          double teq = (t_star - p.delv * cl2 * log(p_norm)) * pow(p_norm, p.kappa);
          teq = std::max(teq, tstr);
          const double tdamp = act ? p.tka + cl4 * (tcoeff * (sigma - p.sigma_b)) : p.tka;
          tdt[x] += -tdamp * (t[x] - teq);
        }
      }
    if (p.num_tracers) {
      double rdamp = p.trsink;
      if (rdamp > 0.0) rdamp = 1.0 / rdamp;
      const double* r = &q[(size_t)prev * K * n];
#pragma omp parallel for schedule(static)
      for (long c = 0; c < (long)n; ++c)
        for (int k = 0; k < K; ++k) {
          const size_t x = (size_t)k * n + c;
          double rst = r[x] + dt * rdt[x];
          double source = k == K - 1 ? p.trflux / (ph[(size_t)K * n + c] - ph[(size_t)(K - 1) * n + c]) : 0.0;
          rdt[x] += source - rdamp * rst;
        }
    }
  }

  // ------------------------------------------------------------------ one call of atmosphere(Time)
  int step() {
    const int K = p.K, J = p.J, I = p.I; const size_t n = JI(), n3 = (size_t)K * n, ns = NS(), s3 = (size_t)K * ns;
    const int prev = previous, cur = current, fut = 1 - cur;
    const double delta_t = prev == cur ? p.dt_atmos : 2 * p.dt_atmos;
    std::vector<double> dt_ug(n3, 0.0), dt_vg(n3, 0.0), dt_tg(n3, 0.0), dt_psg(n, 0.0), dt_tr(p.num_tracers ? n3 : 0, 0.0);
    hs_forcing(delta_t, prev, cur, dt_ug.data(), dt_vg.data(), dt_tg.data(), dt_tr.data());
    // ---- spectral_dynamics
    const double* ugc = &ug[(size_t)cur * n3]; const double* vgc = &vg[(size_t)cur * n3]; const double* tgc = &tg[(size_t)cur * n3];
    const double* psc = &psg[(size_t)cur * n];
    if (p.do_mass) mean_ps_prev = area_weighted_global_mean(&psg[(size_t)prev * n]);
    if (p.do_energy) {
      std::vector<double> e(n3);
      const double* up = &ug[(size_t)prev * n3]; const double* vp = &vg[(size_t)prev * n3]; const double* tp = &tg[(size_t)prev * n3];
#pragma omp parallel for schedule(static)
      for (long x = 0; x < (long)n3; ++x) {
        double a = up[x] + dt_ug[x] * delta_t, b = vp[x] + dt_vg[x] * delta_t;
        e[x] = 0.5 * (a * a + b * b) + p.cp_air * (tp[x] + dt_tg[x] * delta_t);
      }
      mean_e_prev = mass_weighted_global_integral(e.data(), &psg[(size_t)prev * n]);
    }
    if (p.do_water) {
      std::vector<double> w(n3);
      const double* qp = &q[(size_t)prev * n3];
      for (size_t x = 0; x < n3; ++x) w[x] = qp[x] + delta_t * dt_tr[x];
      mean_w_prev = mass_weighted_global_integral(w.data(), &psg[(size_t)prev * n]);
    }
    std::vector<double> ph((K + 1) * n), lnh((K + 1) * n), pf(n3), lnf(n3);
    pressure_variables(psc, ph.data(), lnh.data(), pf.data(), lnf.data());
    std::copy(ph.begin(), ph.end(), p_half.begin() + (size_t)cur * (K + 1) * n);
    std::copy(pf.begin(), pf.end(), p_full.begin() + (size_t)cur * n3);
    // compute_pressure_gradient
    std::vector<double> dx_ps(n), dy_ps(n);
    {
      std::vector<cplx> dxs(ns), dys(ns);
      lon_deriv_cos(&ln_ps[cur * ns], dxs.data(), 1); lat_deriv_cos(&ln_ps[cur * ns], dys.data(), 1);
      spherical_to_grid(dxs.data(), dx_ps.data(), 1); spherical_to_grid(dys.data(), dy_ps.data(), 1);
      for (size_t c = 0; c < n; ++c) { dx_ps[c] = psc[c] * dx_ps[c]; dy_ps[c] = psc[c] * dy_ps[c]; }
      divide_by_cos(dx_ps.data(), 1); divide_by_cos(dy_ps.data(), 1);
    }
    // four_in_one
    std::vector<double> wg((K + 1) * n, 0.0);
    {
      const double kappa = p.rdgas / p.cp_air;
#pragma omp parallel for schedule(static)
      for (long c = 0; c < (long)n; ++c) {
        double dmean_tot = 0.0;
        const double ps = psc[c];
        for (int k = 0; k < K; ++k) {
          const size_t x = (size_t)k * n + c;
          double dp = dpk[k] + dbk[k] * ps, dp_inv = 1 / dp;
          double dlog_1 = lnh[(size_t)(k + 1) * n + c] - lnf[x], dlog_2 = lnf[x] - lnh[x], dlog_3 = lnh[(size_t)(k + 1) * n + c] - lnh[x];
          double x1 = (bk[k + 1] * dlog_1 + bk[k] * dlog_2) * dp_inv, x2 = x1 * dx_ps[c], x3 = x1 * dy_ps[c];
          dt_ug[x] = dt_ug[x] - p.rdgas * tgc[x] * x2;
          dt_vg[x] = dt_vg[x] - p.rdgas * tgc[x] * x3;
          double dmean = divg[x] * dp + dbk[k] * (ugc[x] * dx_ps[c] + vgc[x] * dy_ps[c]);
          double x4 = (dmean_tot * dlog_3 + dmean * dlog_1) * dp_inv, x5 = x4 - ugc[x] * x2 - vgc[x] * x3;
          dt_tg[x] = dt_tg[x] - kappa * tgc[x] * x5;
          wg_full[x] = -x5 * pf[x];
          dmean_tot = dmean_tot + dmean;
          wg[(size_t)(k + 1) * n + c] = -dmean_tot;
        }
        dt_psg[c] = dt_psg[c] - dmean_tot;
        for (int k = 1; k < K; ++k) wg[(size_t)k * n + c] = wg[(size_t)k * n + c] + dmean_tot * bk[k];
        wg[c] = 0.0; wg[(size_t)K * n + c] = 0.0;
      }
    }
    std::vector<double> phi(n3), gh((K + 1) * n);
    compute_geopotential(tgc, lnh.data(), lnf.data(), phi.data(), gh.data());
    std::vector<cplx> dt_ln_ps(ns), dt_ts(s3), dt_vors(s3), dt_divs(s3);
    {
      std::vector<double> t2(n);
      for (size_t c = 0; c < n; ++c) t2[c] = dt_psg[c] / psc[c];
      grid_to_spherical(t2.data(), dt_ln_ps.data(), 1, true);
    }
    std::vector<double> dp(n3);
    for (int k = 0; k < K; ++k) for (size_t c = 0; c < n; ++c) dp[k * n + c] = ph[(size_t)(k + 1) * n + c] - ph[k * n + c];
    vert_adv_centered(wg.data(), dp.data(), ugc, dt_ug.data());
    vert_adv_centered(wg.data(), dp.data(), vgc, dt_vg.data());
    vert_adv_centered(wg.data(), dp.data(), tgc, dt_tg.data());
    {   // horizontal_advection of T
      std::vector<cplx> dxs(s3), dys(s3);
      lon_deriv_cos(&ts[(size_t)cur * s3], dxs.data(), K); lat_deriv_cos(&ts[(size_t)cur * s3], dys.data(), K);
      std::vector<double> dxg(n3), dyg(n3);
      spherical_to_grid(dxs.data(), dxg.data(), K); divide_by_cos(dxg.data(), K);
      spherical_to_grid(dys.data(), dyg.data(), K); divide_by_cos(dyg.data(), K);
#pragma omp parallel for schedule(static)
      for (long x = 0; x < (long)n3; ++x) dt_tg[x] = dt_tg[x] - ugc[x] * dxg[x] - vgc[x] * dyg[x];
    }
    grid_to_spherical(dt_tg.data(), dt_ts.data(), K, true);
#pragma omp parallel for schedule(static)
    for (long x = 0; x < (long)n3; ++x) {
      const int j = (int)((x % n) / I);
      double absv = vorg[x] + cor[j];
      dt_ug[x] = dt_ug[x] + absv * vgc[x];
      dt_vg[x] = dt_vg[x] - absv * ugc[x];
    }
    vor_div_from_uv_grid(dt_ug.data(), dt_vg.data(), dt_vors.data(), dt_divs.data(), K);
    {
      std::vector<double> pke(n3);
#pragma omp parallel for schedule(static)
      for (long x = 0; x < (long)n3; ++x) pke[x] = phi[x] + 0.5 * (ugc[x] * ugc[x] + vgc[x] * vgc[x]);
      std::vector<cplx> sp(s3);
      grid_to_spherical(pke.data(), sp.data(), K, true);
      for (int k = 0; k < K; ++k) for (size_t i = 0; i < ns; ++i) dt_divs[k * ns + i] = dt_divs[k * ns + i] - sp[k * ns + i] * (-eigen[i]);
    }
    if (p.use_implicit) implicit_correction(dt_divs.data(), dt_ts.data(), dt_ln_ps.data(), delta_t, prev, cur);
    // spectral damping (no sponges: checked at create)
    for (int k = 0; k < K; ++k)
      for (size_t i = 0; i < ns; ++i) {
        const size_t x = k * ns + i;
        dt_vors[x] = (1.0 / (1.0 + dampv[i] * delta_t)) * (dt_vors[x] - dampv[i] * vors[(size_t)prev * s3 + x]);
        dt_divs[x] = (1.0 / (1.0 + dampd[i] * delta_t)) * (dt_divs[x] - dampd[i] * divs[(size_t)prev * s3 + x]);
        dt_ts[x] = (1.0 / (1.0 + damp[i] * delta_t)) * (dt_ts[x] - damp[i] * ts[(size_t)prev * s3 + x]);
      }
    // leapfrog_2level_A
    const double rc = p.robert_coeff, raw = p.raw_filter_coeff;
    std::vector<cplx> part_v(s3), part_d(s3), part_t(s3), part_p(ns);
    auto leap = [&](std::vector<cplx>& a, const std::vector<cplx>& dta, std::vector<cplx>& part, size_t sz) {
      for (size_t x = 0; x < sz; ++x) {
        cplx pfv = a[prev * sz + x] - 2.0 * a[cur * sz + x];
        part[x] = pfv;
        if (prev == cur) { a[fut * sz + x] = a[prev * sz + x] + delta_t * dta[x]; a[cur * sz + x] = a[cur * sz + x] + rc * pfv * raw; }
        else { a[cur * sz + x] = a[cur * sz + x] + rc * pfv * raw; a[fut * sz + x] = a[prev * sz + x] + delta_t * dta[x]; }
      }
    };
    leap(ln_ps, dt_ln_ps, part_p, ns); leap(vors, dt_vors, part_v, s3); leap(divs, dt_divs, part_d, s3); leap(ts, dt_ts, part_t, s3);
    spherical_to_grid(&divs[(size_t)fut * s3], divg.data(), K);
    spherical_to_grid(&vors[(size_t)fut * s3], vorg.data(), K);
    uv_grid_from_vor_div(&vors[(size_t)fut * s3], &divs[(size_t)fut * s3], &ug[(size_t)fut * n3], &vg[(size_t)fut * n3], K);
    spherical_to_grid(&ts[(size_t)fut * s3], &tg[(size_t)fut * n3], K);
    spherical_to_grid(&ln_ps[(size_t)fut * ns], &psg[(size_t)fut * n], 1);
    for (size_t c = 0; c < n; ++c) psg[fut * n + c] = exp(psg[fut * n + c]);
    {
      double tmin = 1e300, tmax = -1e300;
      for (size_t x = 0; x < n3; ++x) { tmin = std::min(tmin, tg[fut * n3 + x]); tmax = std::max(tmax, tg[fut * n3 + x]); }
      if (tmin < p.valid_t_lo || tmax > p.valid_t_hi) { err = "temperatures out of valid range"; return 1; }
    }
    // update_tracers (grid tracer sphum)
    std::vector<double> part_q;
    const double rct = p.tracer_robert_coeff < 0 ? p.robert_coeff : p.tracer_robert_coeff;
    if (p.num_tracers) {
      std::vector<double> trf(n3), dq(n3);
      const double* qp = &q[(size_t)prev * n3];
      for (size_t x = 0; x < n3; ++x) trf[x] = qp[x] + delta_t * dt_tr[x];
#pragma omp parallel for schedule(dynamic, 1)
      for (int k = 0; k < K; ++k) fv_level(ugc + (size_t)k * n, vgc + (size_t)k * n, &trf[(size_t)k * n], delta_t, &dq[(size_t)k * n]);
      for (size_t x = 0; x < n3; ++x) trf[x] = trf[x] + delta_t * dq[x];
#pragma omp parallel for schedule(static)
      for (long c = 0; c < (long)n; ++c) {
        double w[257], dz[256], r[256], rdt[256];
        for (int k = 0; k <= K; ++k) w[k] = wg[(size_t)k * n + c];
        for (int k = 0; k < K; ++k) { dz[k] = dp[(size_t)k * n + c]; r[k] = trf[(size_t)k * n + c]; }
        ppm_column(K, delta_t, w, dz, r, rdt);
        for (int k = 0; k < K; ++k) trf[(size_t)k * n + c] = r[k] + delta_t * rdt[k];
      }
      part_q.resize(n3);
      for (size_t x = 0; x < n3; ++x) {
        double pfv = q[prev * n3 + x] - 2.0 * q[cur * n3 + x];
        part_q[x] = pfv;
        q[cur * n3 + x] = q[cur * n3 + x] + rct * pfv * raw;
        q[fut * n3 + x] = trf[x];
      }
    }
    // compute_corrections
    if (p.do_mass) {
      double mean_ps = area_weighted_global_mean(&psg[(size_t)fut * n]);
      double f = mean_ps_prev / mean_ps;
      for (size_t c = 0; c < n; ++c) psg[fut * n + c] = f * psg[fut * n + c];
      ln_ps[fut * ns] = ln_ps[fut * ns] + sqrt(2.0) * log(f);
    }
    if (p.do_energy) {
      std::vector<double> e(n3);
      const double* uf = &ug[(size_t)fut * n3]; const double* vf = &vg[(size_t)fut * n3]; const double* tf = &tg[(size_t)fut * n3];
#pragma omp parallel for schedule(static)
      for (long x = 0; x < (long)n3; ++x) e[x] = 0.5 * (uf[x] * uf[x] + vf[x] * vf[x]) + p.cp_air * tf[x];
      double mean_e = mass_weighted_global_integral(e.data(), &psg[(size_t)fut * n]);
      double tc = p.grav * (mean_e_prev - mean_e) / (p.cp_air * mean_ps_prev);
      for (size_t x = 0; x < n3; ++x) tg[fut * n3 + x] = tg[fut * n3 + x] + tc;
      for (int k = 0; k < K; ++k) ts[(size_t)fut * s3 + k * ns] = ts[(size_t)fut * s3 + k * ns] + sqrt(2.0) * tc;
    }
    if (p.do_water) {
      double* qf = &q[(size_t)fut * n3];
      const double* psf = &psg[(size_t)fut * n];
      std::vector<double> a(n3), b(n3);
      for (size_t x = 0; x < n3; ++x) { bool m = pf[x] >= p.water_correction_limit; a[x] = m ? qf[x] : 0.0; b[x] = m ? 0.0 : qf[x]; }
      double mean_w = mass_weighted_global_integral(qf, psf), corr = mass_weighted_global_integral(a.data(), psf),
             ncorr = mass_weighted_global_integral(b.data(), psf);
      if (mean_w > 0.0) {
        double wf = mean_w_prev / mean_w;
        wf = wf * (1.0 + ncorr / corr) - ncorr / corr;
        for (size_t x = 0; x < n3; ++x) if (pf[x] >= p.water_correction_limit) qf[x] = wf * qf[x];
      }
    }
    previous = cur; current = fut;
    // complete_robert_filter -> leapfrog_2level_B
    {
      const int pp = previous, cc = current;
      auto lb = [&](std::vector<cplx>& a, const std::vector<cplx>& part, size_t sz) {
        for (size_t x = 0; x < sz; ++x) {
          a[pp * sz + x] = a[pp * sz + x] + rc * a[cc * sz + x] * raw;
          a[cc * sz + x] = a[cc * sz + x] + rc * (part[x] + a[cc * sz + x]) * (raw - 1.0);
        }
      };
      lb(ln_ps, part_p, ns); lb(vors, part_v, s3); lb(divs, part_d, s3); lb(ts, part_t, s3);
      if (p.num_tracers)
        for (size_t x = 0; x < n3; ++x) {
          q[pp * n3 + x] = q[pp * n3 + x] + rct * q[cc * n3 + x] * raw;
          q[cc * n3 + x] = q[cc * n3 + x] + rct * (part_q[x] + q[cc * n3 + x]) * (raw - 1.0);
        }
    }
    pressures_and_heights(current);
    return 0;
  }
};

template <class T>
void cp(std::vector<T>& dst, const T* src, size_t n) { dst.assign(src, src + n); }
}  // namespace

extern "C" {

void* cstep_create(const Params* pp, const Tables* t) {
  Core* c = new Core();
  c->p = *pp;
  const Params& p = c->p;
  if (p.I & (p.I - 1) || p.I < 8 || p.M + 1 > 1024 || p.K > 255 || p.I / 2 < p.M) { delete c; return nullptr; }
  const size_t ns = c->NS(), J = p.J, K = p.K, n = c->JI();
  cp(c->leg, t->legendre, (J / 2) * ns); cp(c->legw, t->legendre_wts, (J / 2) * ns);
  cp(c->cosm, t->cosm_lat, J); cp(c->wts, t->wts_lat, J); cp(c->radlat, t->rad_lat, J); cp(c->cor, t->coriolis, J);
  cp(c->tri, t->triangle_mask, ns); cp(c->eigen, t->eigen, ns); cp(c->uvm, t->coef_uvm, ns); cp(c->uvc, t->coef_uvc, ns); cp(c->uvp, t->coef_uvp, ns);
  cp(c->alpm, t->coef_alpm, ns); cp(c->alpp, t->coef_alpp, ns); cp(c->dym, t->coef_dym, ns); cp(c->cdx, t->coef_dx, ns); cp(c->dyp, t->coef_dyp, ns);
  cp(c->pk, t->pk, K + 1); cp(c->bk, t->bk, K + 1);
  c->dpk.resize(K); c->dbk.resize(K);
  for (size_t k = 0; k < K; ++k) { c->dpk[k] = c->pk[k + 1] - c->pk[k]; c->dbk[k] = c->bk[k + 1] - c->bk[k]; }
  cp(c->damp, t->damping, ns); cp(c->dampv, t->damping_vor, ns); cp(c->dampd, t->damping_div, ns);
  cp(c->ref_t, t->ref_t, K); cp(c->ref_lh, t->ref_ln_p_half, K + 1); cp(c->ref_lf, t->ref_ln_p_full, K); cp(c->hvec, t->h, K);
  cp(c->wave_dt, t->wave_dt, (size_t)p.N * K * K); cp(c->wave_2dt, t->wave_2dt, (size_t)p.N * K * K);
  cp(c->fvc, t->fv_c, J); cp(c->fvcc, t->fv_cc, J + 1); cp(c->fvdy, t->fv_dy, J + 4); cp(c->fvdyy, t->fv_dyy, J + 1);
  cp(c->fvdyp, t->fv_dy_plus, J + 2); cp(c->fvdym, t->fv_dy_minus, J + 2);
  c->vors.assign(2 * K * ns, 0); c->divs.assign(2 * K * ns, 0); c->ts.assign(2 * K * ns, 0); c->ln_ps.assign(2 * ns, 0);
  c->ug.assign(2 * K * n, 0); c->vg.assign(2 * K * n, 0); c->tg.assign(2 * K * n, 0); c->psg.assign(2 * n, 0);
  c->vorg.assign(K * n, 0); c->divg.assign(K * n, 0); c->q.assign(p.num_tracers ? 2 * K * n : 0, 0); c->wg_full.assign(K * n, 0);
  c->surf_geo.assign(n, 0); c->p_half.assign(2 * (K + 1) * n, 0); c->p_full.assign(2 * K * n, 0);
  c->z_half.assign(2 * (K + 1) * n, 0); c->z_full.assign(2 * K * n, 0);
  c->fft_init();
  return c;
}
void cstep_destroy(void* h) { delete (Core*)h; }
const char* cstep_error(void* h) { return ((Core*)h)->err.c_str(); }
int cstep_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// both time levels: spectral [2][K][N+1][M+1] complex, grid [2][K][J][I], psg [2][J][I], vorg/divg [K][J][I], q [2][K][J][I] (or NULL)
void cstep_set_state(void* h, const cplx* vors, const cplx* divs, const cplx* ts, const cplx* ln_ps, const double* ug, const double* vg,
                     const double* tg, const double* psg, const double* vorg, const double* divg, const double* q, int previous, int current) {
  Core* c = (Core*)h;
  std::copy(vors, vors + c->vors.size(), c->vors.begin()); std::copy(divs, divs + c->divs.size(), c->divs.begin());
  std::copy(ts, ts + c->ts.size(), c->ts.begin()); std::copy(ln_ps, ln_ps + c->ln_ps.size(), c->ln_ps.begin());
  std::copy(ug, ug + c->ug.size(), c->ug.begin()); std::copy(vg, vg + c->vg.size(), c->vg.begin());
  std::copy(tg, tg + c->tg.size(), c->tg.begin()); std::copy(psg, psg + c->psg.size(), c->psg.begin());
  std::copy(vorg, vorg + c->vorg.size(), c->vorg.begin()); std::copy(divg, divg + c->divg.size(), c->divg.begin());
  if (q && c->p.num_tracers) std::copy(q, q + c->q.size(), c->q.begin());
  c->previous = previous; c->current = current;
  for (int lev : {previous, current}) c->pressures_and_heights(lev);
}
void cstep_get_state(void* h, cplx* vors, cplx* divs, cplx* ts, cplx* ln_ps, double* ug, double* vg, double* tg, double* psg, double* vorg,
                     double* divg, double* q, double* wg_full, int* previous, int* current) {
  Core* c = (Core*)h;
  std::copy(c->vors.begin(), c->vors.end(), vors); std::copy(c->divs.begin(), c->divs.end(), divs);
  std::copy(c->ts.begin(), c->ts.end(), ts); std::copy(c->ln_ps.begin(), c->ln_ps.end(), ln_ps);
  std::copy(c->ug.begin(), c->ug.end(), ug); std::copy(c->vg.begin(), c->vg.end(), vg);
  std::copy(c->tg.begin(), c->tg.end(), tg); std::copy(c->psg.begin(), c->psg.end(), psg);
  std::copy(c->vorg.begin(), c->vorg.end(), vorg); std::copy(c->divg.begin(), c->divg.end(), divg);
  if (q && c->p.num_tracers) std::copy(c->q.begin(), c->q.end(), q);
  if (wg_full) std::copy(c->wg_full.begin(), c->wg_full.end(), wg_full);
  *previous = c->previous; *current = c->current;
}
int cstep_step(void* h, int nsteps) {
  Core* c = (Core*)h;
  for (int i = 0; i < nsteps; ++i) if (c->step()) return 1;
  return 0;
}
// transform-level entry points (for the tests): spec [nlev][N+1][M+1] <-> grid [nlev][J][I]
void cstep_spherical_to_grid(void* h, const cplx* spec, double* grid, int nlev) { ((Core*)h)->spherical_to_grid(spec, grid, nlev); }
void cstep_grid_to_spherical(void* h, const double* grid, cplx* spec, int nlev, int trunc) { ((Core*)h)->grid_to_spherical(grid, spec, nlev, trunc != 0); }

}  // extern "C"
