// spectral.cu -- spectral-space kernels on the packed triangular layout.
//
//   spec_gradient     compute_gradient_cos           tools/spherical.F90:270-351
//   spec_ucos_vcos    compute_ucos_vcos              tools/spherical.F90:409-469
//   spec_tend_adjust  compute_vor_div + truncation + laplacian + adjust_dt_divs
//                     (tools/spherical.F90:472-561, transforms.F90:742-783,
//                      model/spectral_dynamics.F90:900-904, model/implicit.F90:289-325)
//   spec_wave_matvec  dt_divs(m,n,:) = wave_matrix(:,:,L) x dt_divs(m,n,:)     implicit.F90:270-278
//   spec_update       rest of implicit_correction (:280-284), compute_spectral_damping{,_vor,_div}
//                     (model/spectral_damping.F90:172-291), leapfrog_2level_A (model/leapfrog.F90:58-83)
//   spec_robert_b     leapfrog_2level_B  (leapfrog.F90:87-105) via complete_robert_filter
//
// Column kernels stage a tile of RP packed rows x K levels in shared memory as [k][col]
// (col = row*2 + re/im): the k-recurrences of linear_tp_tendency / linear_geopotential are then run
// by one thread per real column in exactly the reference's top-down / bottom-up order.
#include "device.h"
#include "spectral.h"

namespace isca {

__device__ __forceinline__ double2 times_i(double2 a) { return make_double2(-a.y, a.x); }
__device__ __forceinline__ double2 axpy2(double s, double2 a, double2 acc) {
  return make_double2(acc.x + s * a.x, acc.y + s * a.y);
}

// ---------------------------------------------------------------------------------------------
// gradient: dst[.., dx_off+k] = i m/a S ; dst[.., dy_off+k] = -dym(n) S(n-1) + dyp(n) S(n+1)
// ---------------------------------------------------------------------------------------------
__global__ void spec_gradient_kernel(DevTables t, const double2* __restrict__ src, int Ls, int nlev,
                                     double2* __restrict__ dst, int Lp, int dx_off, int dy_off) {
  const GeomDev& g = t.g;
  const int p = blockIdx.x;
  const int n = t.row_n[p];
  const int m = g.m_of[g.row_m[p]];
  const int Nm = g.M - m + 2;
  const double cdx = t.coef_dx[p], cdym = t.coef_dym[p], cdyp = t.coef_dyp[p];
  for (int k = threadIdx.x; k < nlev; k += blockDim.x) {
    double2 s = src[(size_t)p * Ls + k];
    double2 dx = times_i(s); dx.x *= cdx; dx.y *= cdx;
    double2 dy = make_double2(0.0, 0.0);
    if (n >= 1) { double2 sm = src[(size_t)(p - 1) * Ls + k]; dy.x = -sm.x * cdym; dy.y = -sm.y * cdym; }
    if (n + 1 < Nm) { double2 sp = src[(size_t)(p + 1) * Ls + k]; dy.x = dy.x + sp.x * cdyp; dy.y = dy.y + sp.y * cdyp; }
    dst[(size_t)p * Lp + dx_off + k] = dx;
    dst[(size_t)p * Lp + dy_off + k] = dy;
  }
}
void launch_spec_gradient(const DevTables& t, const double2* src, int Ls, int nlev, double2* dst, int Lp,
                          int dx_off, int dy_off, cudaStream_t st) {
  spec_gradient_kernel<<<t.g.T, 64, 0, st>>>(t, src, Ls, nlev, dst, Lp, dx_off, dy_off);
}

// ---------------------------------------------------------------------------------------------
// (vor, div) -> (u cos, v cos), all inside one batch buffer
// ---------------------------------------------------------------------------------------------
__global__ void spec_ucos_vcos_kernel(DevTables t, double2* __restrict__ buf, int Lp, int nlev,
                                      int vor_off, int div_off, int u_off, int v_off) {
  const GeomDev& g = t.g;
  const int p = blockIdx.x;
  const int n = t.row_n[p];
  const int m = g.m_of[g.row_m[p]];
  const int Nm = g.M - m + 2;
  const double uvc = t.coef_uvc[p], uvm = t.coef_uvm[p], uvp = t.coef_uvp[p];
  for (int k = threadIdx.x; k < nlev; k += blockDim.x) {
    const double2 vor = buf[(size_t)p * Lp + vor_off + k];
    const double2 div = buf[(size_t)p * Lp + div_off + k];
    double2 u = times_i(div); u.x *= uvc; u.y *= uvc;
    double2 v = times_i(vor); v.x *= uvc; v.y *= uvc;
    if (n >= 1) {
      double2 vm = buf[(size_t)(p - 1) * Lp + vor_off + k], dm = buf[(size_t)(p - 1) * Lp + div_off + k];
      u.x = u.x + uvm * vm.x; u.y = u.y + uvm * vm.y;
      v.x = v.x - uvm * dm.x; v.y = v.y - uvm * dm.y;
    }
    if (n + 1 < Nm) {
      double2 vp = buf[(size_t)(p + 1) * Lp + vor_off + k], dp = buf[(size_t)(p + 1) * Lp + div_off + k];
      u.x = u.x - uvp * vp.x; u.y = u.y - uvp * vp.y;
      v.x = v.x + uvp * dp.x; v.y = v.y + uvp * dp.y;
    }
    buf[(size_t)p * Lp + u_off + k] = u;
    buf[(size_t)p * Lp + v_off + k] = v;
  }
}
void launch_spec_ucos_vcos(const DevTables& t, double2* buf, int Lp, int nlev, int vor_off, int div_off,
                           int u_off, int v_off, cudaStream_t st) {
  spec_ucos_vcos_kernel<<<t.g.T, 64, 0, st>>>(t, buf, Lp, nlev, vor_off, div_off, u_off, v_off);
}

// Fused tail of the spectral step, all inside the next inverse batch buffer:
//   (vor, div) -> (u cos, v cos)                                   compute_ucos_vcos   spherical.F90:409-469
//   T, ln ps   -> (dx, dy) gradients used by the NEXT step        compute_gradient_cos spherical.F90:270-351
// The gradients are unaffected by the mass/energy fixers (they only change the (0,0) coefficient, whose
// gradient is identically zero) and, with raw_filter_coeff = 1, by leapfrog_2level_B.
__global__ void spec_post_kernel(DevTables t, double2* __restrict__ buf, int Lp, int K, int vor_off, int div_off, int u_off,
                                 int v_off, int t_off, int lnps_off, int dxt_off, int dyt_off, int dxl_off, int dyl_off) {
  const GeomDev& g = t.g;
  const int p = blockIdx.x;
  const int n = t.row_n[p];
  const int m = g.m_of[g.row_m[p]];
  const int Nm = g.M - m + 2;
  const double uvc = t.coef_uvc[p], uvm = t.coef_uvm[p], uvp = t.coef_uvp[p];
  const double cdx = t.coef_dx[p], cdym = t.coef_dym[p], cdyp = t.coef_dyp[p];
  const bool has_m = (n >= 1), has_p = (n + 1 < Nm);
  double2* row = buf + (size_t)p * Lp;
  const double2* rm = row - Lp;
  const double2* rp = row + Lp;
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    if (k < K) {
      const double2 vor = row[vor_off + k], div = row[div_off + k];
      double2 u = times_i(div); u.x *= uvc; u.y *= uvc;
      double2 v = times_i(vor); v.x *= uvc; v.y *= uvc;
      if (has_m) {
        const double2 vm = rm[vor_off + k], dm = rm[div_off + k];
        u.x = u.x + uvm * vm.x; u.y = u.y + uvm * vm.y;
        v.x = v.x - uvm * dm.x; v.y = v.y - uvm * dm.y;
      }
      if (has_p) {
        const double2 vp = rp[vor_off + k], dp = rp[div_off + k];
        u.x = u.x - uvp * vp.x; u.y = u.y - uvp * vp.y;
        v.x = v.x + uvp * dp.x; v.y = v.y + uvp * dp.y;
      }
      row[u_off + k] = u; row[v_off + k] = v;
    }
    // gradient of T (k < K) or of ln ps (k == K)
    const int so = (k < K) ? t_off + k : lnps_off;
    const int dxo = (k < K) ? dxt_off + k : dxl_off;
    const int dyo = (k < K) ? dyt_off + k : dyl_off;
    const double2 s0 = row[so];
    double2 dx = times_i(s0); dx.x *= cdx; dx.y *= cdx;
    double2 dy = make_double2(0.0, 0.0);
    if (has_m) { const double2 sm = rm[so]; dy.x = -sm.x * cdym; dy.y = -sm.y * cdym; }
    if (has_p) { const double2 sp = rp[so]; dy.x = dy.x + sp.x * cdyp; dy.y = dy.y + sp.y * cdyp; }
    row[dxo] = dx; row[dyo] = dy;
  }
}

// (u cos / cos^2 .. ) -> (vor, div): vor = alpha(vcos_s, ucos_s, -1), div = alpha(ucos_s, vcos_s, +1), truncated
__global__ void spec_vor_div_kernel(DevTables t, const double2* __restrict__ buf, int Lp, int nlev, int a_off, int b_off,
                                    double2* __restrict__ out, int Lo, int vor_off, int div_off) {
  const GeomDev& g = t.g;
  const int p = blockIdx.x;
  const int n = t.row_n[p];
  const int m = g.m_of[g.row_m[p]];
  const int Nm = g.M - m + 2;
  const double cdx = t.coef_dx[p], alpm = t.coef_alpm[p], alpp = t.coef_alpp[p], mask = t.trunc_mask[p];
  for (int k = threadIdx.x; k < nlev; k += blockDim.x) {
    const double2 A = buf[(size_t)p * Lp + a_off + k], B = buf[(size_t)p * Lp + b_off + k];
    double2 vor = times_i(B); vor.x *= cdx; vor.y *= cdx;
    double2 div = times_i(A); div.x *= cdx; div.y *= cdx;
    if (n >= 1) {
      double2 Am = buf[(size_t)(p - 1) * Lp + a_off + k], Bm = buf[(size_t)(p - 1) * Lp + b_off + k];
      vor.x = vor.x + alpm * Am.x; vor.y = vor.y + alpm * Am.y;
      div.x = div.x - alpm * Bm.x; div.y = div.y - alpm * Bm.y;
    }
    if (n + 1 < Nm) {
      double2 Ap = buf[(size_t)(p + 1) * Lp + a_off + k], Bp = buf[(size_t)(p + 1) * Lp + b_off + k];
      vor.x = vor.x - alpp * Ap.x; vor.y = vor.y - alpp * Ap.y;
      div.x = div.x + alpp * Bp.x; div.y = div.y + alpp * Bp.y;
    }
    out[(size_t)p * Lo + vor_off + k] = make_double2(vor.x * mask, vor.y * mask);
    out[(size_t)p * Lo + div_off + k] = make_double2(div.x * mask, div.y * mask);
  }
}
void launch_spec_vor_div(const DevTables& t, const double2* buf, int Lp, int nlev, int a_off, int b_off,
                         double2* out, int Lo, int vor_off, int div_off, cudaStream_t st) {
  spec_vor_div_kernel<<<t.g.T, 64, 0, st>>>(t, buf, Lp, nlev, a_off, b_off, out, Lo, vor_off, div_off);
}

// ---------------------------------------------------------------------------------------------
// S1: explicit spectral tendencies + adjust_dt_divs
// ---------------------------------------------------------------------------------------------
constexpr int RP = 16;            // packed rows per CTA
constexpr int NC = 2 * RP;        // real columns per CTA
constexpr int NCP = NC + 1;       // padded row stride of the [k][column] shared tiles (odd: conflict-free both ways)

// linear_tp_tendency (implicit.F90:414-480) for one real column held in shared memory [k][NC].
// ref_temperature_implicit is 300 K at every level (spectral_dynamics.F90:473), so the
// vert_vel*(t_ref(k)-t_ref(k-1)) term of :464-476 is identically zero and is not evaluated.
// Column-independent factors of the reference-state recurrences, hoisted out of the per-column loops
// (same operands, same operation order as implicit.F90:441-448 / :345-356, so the results are unchanged):
//   lc[0][k] = dp = dpk(k)+dbk(k)*p_surf_ref, lc[1][k] = 1/dp, lc[2][k] = dlog_1, lc[3][k] = dlog_3,
//   lc[4][k] = ln_p_half(k+1)-ln_p_full(k) (= dlog_1), lc[5][k] = ln_p_half(k+1)-ln_p_half(k) (= dlog_3), lc[6][k] = t_ref(k)
constexpr int NLC = 5;
__device__ __forceinline__ void load_level_consts(const DevTables& t, const Params& pr, double* lc, int K) {
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double dp = t.dpk[k] + t.dbk[k] * pr.ref_ps;
    lc[0 * K + k] = dp;
    lc[1 * K + k] = 1 / dp;
    lc[2 * K + k] = t.ref_ln_p_half[k + 1] - t.ref_ln_p_full[k];
    lc[3 * K + k] = t.ref_ln_p_half[k + 1] - t.ref_ln_p_half[k];
    lc[4 * K + k] = t.ref_t[k];
  }
}
__device__ __forceinline__ double tp_tendency_column(const Params& pr, const double* lc, const double* div, double* dt_t,
                                                     int K, int col) {
  double dmean_tot = 0.0;
  for (int k = 0; k < K; ++k) {
    const double dmean = div[k * NCP + col] * lc[k];
    dt_t[k * NCP + col] = -pr.kappa * lc[4 * K + k] * (dmean_tot * lc[3 * K + k] + dmean * lc[2 * K + k]) * lc[K + k];
    dmean_tot = dmean_tot + dmean;
  }
  return -dmean_tot;      // dt_p_surf
}

__global__ void __launch_bounds__(128)
spec_tend_adjust_kernel(DevTables t, Params pr, SpecStepArgs a) {
  extern __shared__ __align__(16) double sm[];
  const GeomDev& g = t.g;
  const int K = g.K;
  double* s_ddiv = sm;                    // [K][NC] dt_divs
  double* s_dT = s_ddiv + K * NCP;         // [K][NC] dt_ts
  double* s_dD = s_dT + K * NCP;           // [K][NC] divs(prev) - divs(cur); later scratch
  double* s_dTs = s_dD + K * NCP;          // [K][NC] ts(prev) - ts(cur)
  double* s_ps = s_dTs + K * NCP;          // [2][NC]  dt_ln_ps ; ln_ps(prev)-ln_ps(cur)
  double* s_lc = s_ps + 2 * NC;            // [NLC][K] level constants
  const int p0 = blockIdx.x * RP;
  const int tid = threadIdx.x;
  const int LpB = a.LpB;
  load_level_consts(t, pr, s_lc, K);

  // phase 1: elementwise (row, level)
  for (int idx = tid; idx < RP * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const int p = p0 + r;
    double2 dvor = make_double2(0, 0), ddiv = make_double2(0, 0), dT = make_double2(0, 0);
    double2 dD = make_double2(0, 0), dTs = make_double2(0, 0);
    if (p < g.T) {
      const int n = t.row_n[p];
      const int m = g.m_of[g.row_m[p]];
      const int Nm = g.M - m + 2;
      const double cdx = t.coef_dx[p], alpm = t.coef_alpm[p], alpp = t.coef_alpp[p], mask = t.trunc_mask[p];
      const double2 A = a.specB[(size_t)p * LpB + a.oA + k], B = a.specB[(size_t)p * LpB + a.oB + k];
      dvor = times_i(B); dvor.x *= cdx; dvor.y *= cdx;      // alpha(vcos, ucos, -1)
      ddiv = times_i(A); ddiv.x *= cdx; ddiv.y *= cdx;      // alpha(ucos, vcos, +1)
      if (n >= 1) {
        const double2 Am = a.specB[(size_t)(p - 1) * LpB + a.oA + k], Bm = a.specB[(size_t)(p - 1) * LpB + a.oB + k];
        dvor.x = dvor.x + alpm * Am.x; dvor.y = dvor.y + alpm * Am.y;
        ddiv.x = ddiv.x - alpm * Bm.x; ddiv.y = ddiv.y - alpm * Bm.y;
      }
      if (n + 1 < Nm) {
        const double2 Ap = a.specB[(size_t)(p + 1) * LpB + a.oA + k], Bp = a.specB[(size_t)(p + 1) * LpB + a.oB + k];
        dvor.x = dvor.x - alpp * Ap.x; dvor.y = dvor.y - alpp * Ap.y;
        ddiv.x = ddiv.x + alpp * Bp.x; ddiv.y = ddiv.y + alpp * Bp.y;
      }
      dvor.x *= mask; dvor.y *= mask; ddiv.x *= mask; ddiv.y *= mask;      // triangular_truncation
      const double2 phi = a.specB[(size_t)p * LpB + a.oPhi + k];
      const double eig = t.eigen[p];
      ddiv.x = ddiv.x - (phi.x * (-eig)); ddiv.y = ddiv.y - (phi.y * (-eig));   // dt_divs - compute_laplacian(phi)
      dT = a.specB[(size_t)p * LpB + a.oT + k];
      a.dt_vors[(size_t)p * K + k] = dvor;
      const double2 dpv = a.divs_prev[(size_t)p * K + k], dcu = a.divs_cur[(size_t)p * K + k];
      const double2 tpv = a.ts_prev[(size_t)p * K + k], tcu = a.ts_cur[(size_t)p * K + k];
      dD = make_double2(dpv.x - dcu.x, dpv.y - dcu.y);
      dTs = make_double2(tpv.x - tcu.x, tpv.y - tcu.y);
    }
    const int c = 2 * r;
    s_ddiv[k * NCP + c] = ddiv.x; s_ddiv[k * NCP + c + 1] = ddiv.y;
    s_dT[k * NCP + c] = dT.x;     s_dT[k * NCP + c + 1] = dT.y;
    s_dD[k * NCP + c] = dD.x;     s_dD[k * NCP + c + 1] = dD.y;
    s_dTs[k * NCP + c] = dTs.x;   s_dTs[k * NCP + c + 1] = dTs.y;
  }
  if (tid < RP) {
    const int p = p0 + tid;
    double2 dl = make_double2(0, 0), dd = make_double2(0, 0);
    if (p < g.T) {
      dl = a.specB[(size_t)p * LpB + a.oLnps];
      const double2 lp = a.lnps_prev[p], lc = a.lnps_cur[p];
      dd = make_double2(lp.x - lc.x, lp.y - lc.y);
    }
    s_ps[2 * tid] = dl.x; s_ps[2 * tid + 1] = dl.y;
    s_ps[NC + 2 * tid] = dd.x; s_ps[NC + 2 * tid + 1] = dd.y;
  }
  __syncthreads();

  // phase 2: adjust_dt_divs, one thread per real column (implicit.F90:289-325)
  if (a.use_implicit && tid < NC) {
    const int col = tid;
    const int p = p0 + (col >> 1);
    if (p < g.T) {
      const double xi = pr.xi;
      // dt_ts += tp(dD).dt_t ; dt_ln_ps += tp(dD).dt_p / ref_ps   -- reuse s_dD as output of tp (reads then writes same k)
      double dmean_tot = 0.0;
      for (int k = 0; k < K; ++k) {
        const double dmean = s_dD[k * NCP + col] * s_lc[k];
        const double dtt = -pr.kappa * s_lc[4 * K + k] * (dmean_tot * s_lc[3 * K + k] + dmean * s_lc[2 * K + k]) * s_lc[K + k];
        dmean_tot = dmean_tot + dmean;
        s_dT[k * NCP + col] = s_dT[k * NCP + col] + dtt;
      }
      const double dt_ps_temp = -dmean_tot;
      const double dlnps = s_ps[col] + dt_ps_temp / pr.ref_ps;
      s_ps[col] = dlnps;
      const double ps_temp = s_ps[NC + col] + xi * dlnps;
      // linear_geopotential(ts_temp, 0, 0) bottom-up (implicit.F90:329-359)
      const double eig = t.eigen[p];
      double gh = 0.0;                                  // geopot_half(K+1)
      for (int k = K - 1; k >= 0; --k) {
        const double ts_temp = s_dTs[k * NCP + col] + xi * s_dT[k * NCP + col];
        const double geopot = gh + pr.rdgas * (ts_temp * s_lc[2 * K + k] + s_lc[4 * K + k] * (0.0 - 0.0));
        s_ddiv[k * NCP + col] = s_ddiv[k * NCP + col] + eig * (geopot + t.h_impl[k] * ps_temp * pr.ref_ps);
        if (k >= 1) gh = gh + pr.rdgas * (ts_temp * s_lc[3 * K + k] + s_lc[4 * K + k] * (0.0 - 0.0));
      }
    }
  }
  __syncthreads();
  // phase 3: store work arrays
  for (int idx = tid; idx < RP * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const int p = p0 + r;
    if (p < g.T) {
      const int c = 2 * r;
      a.w_div[(size_t)p * K + k] = make_double2(s_ddiv[k * NCP + c], s_ddiv[k * NCP + c + 1]);
      a.w_T[(size_t)p * K + k] = make_double2(s_dT[k * NCP + c], s_dT[k * NCP + c + 1]);
    }
  }
  if (tid < RP && p0 + tid < g.T) a.w_lnps[p0 + tid] = make_double2(s_ps[2 * tid], s_ps[2 * tid + 1]);
}

// ---------------------------------------------------------------------------------------------
// S2: per total wavenumber L, y = wave_matrix(L) x  for every column with m + n = L
// ---------------------------------------------------------------------------------------------
constexpr int WM_COLS = 64;      // real columns per CTA
constexpr int WM_KG = 4;         // output-level groups per column (threads = WM_COLS * WM_KG)
__global__ void __launch_bounds__(WM_COLS * WM_KG)
spec_wave_matvec_kernel(DevTables t, double2* __restrict__ w_div) {
  extern __shared__ __align__(16) double sm[];
  const GeomDev& g = t.g;
  const int K = g.K;
  const int L = blockIdx.x;
  double* W = sm;                 // [K][K]
  double* X = W + K * K;          // [K][WM_COLS]
  const int tid = threadIdx.x;
  const int cl = tid % WM_COLS, kg = tid / WM_COLS;
  const int col = blockIdx.y * WM_COLS + cl;      // real column index: (mi_idx, reim)
  int cnt = 0;                                    // number of local m <= L (m_of is ascending)
  {
    int lo = 0, hi = g.nm;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (g.m_of[mid] <= L) lo = mid + 1; else hi = mid; }
    cnt = lo;
  }
  if (blockIdx.y * WM_COLS >= 2 * cnt) return;
  for (int i = tid; i < K * K; i += WM_COLS * WM_KG) W[i] = t.wave_matrix[(size_t)L * K * K + i];
  const bool active = col < 2 * cnt;
  size_t base = 0;
  if (active) {
    const int mi = col >> 1;
    const int p = g.off[mi] + (L - g.m_of[mi]);
    base = ((size_t)p * K) * 2 + (col & 1);
    const double* src = reinterpret_cast<const double*>(w_div);
    for (int k = kg; k < K; k += WM_KG) X[k * WM_COLS + cl] = src[base + 2 * k];
  }
  __syncthreads();
  if (active) {
    double* dst = reinterpret_cast<double*>(w_div);
    for (int k = kg; k < K; k += WM_KG) {
      double s = 0.0;
      for (int q = 0; q < K; ++q) s += W[k * K + q] * X[q * WM_COLS + cl];   // matmul(wave_matrix(:,:,L), dt_divs(m,n,:))
      dst[base + 2 * k] = s;
    }
  }
}

// The dynamic shared-memory limit of the kernel only ever grows (one process-wide high-water mark shared by every caller:
// lowering it for a handle with few levels would make later launches of a larger handle fail).
static void launch_wave_matvec(const DevTables& t, double2* w_div, cudaStream_t st) {
  const GeomDev& g = t.g;
  const int K = g.K;
  const size_t smem = sizeof(double) * ((size_t)K * K + (size_t)K * WM_COLS);
  static size_t attr = 0;
  if (smem > attr) { cudaFuncSetAttribute(spec_wave_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
  dim3 grid(g.M + 1, (2 * g.nm + WM_COLS - 1) / WM_COLS);
  spec_wave_matvec_kernel<<<grid, WM_COLS * WM_KG, smem, st>>>(t, w_div);
}

// ---------------------------------------------------------------------------------------------
// S3: finish implicit correction, damping, leapfrog part A, and emit the next inverse batch
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
spec_update_kernel(DevTables t, Params pr, SpecStepArgs a) {
  extern __shared__ __align__(16) double sm[];
  const GeomDev& g = t.g;
  const int K = g.K;
  double* s_ddiv = sm;                    // [K][NC]
  double* s_tmp = s_ddiv + K * NCP;        // [K][NC] dt_ts_temp
  double* s_ps = s_tmp + K * NCP;          // [NC] dt_ps_temp
  double* s_lc = s_ps + NC;                // [NLC][K] level constants
  const int p0 = blockIdx.x * RP;
  const int tid = threadIdx.x;
  load_level_consts(t, pr, s_lc, K);

  for (int idx = tid; idx < RP * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const int p = p0 + r;
    double2 v = make_double2(0, 0);
    if (p < g.T) v = a.w_div[(size_t)p * K + k];
    s_ddiv[k * NCP + 2 * r] = v.x; s_ddiv[k * NCP + 2 * r + 1] = v.y;
  }
  __syncthreads();
  if (tid < NC) {
    double dps = 0.0;
    if (a.use_implicit) dps = tp_tendency_column(pr, s_lc, s_ddiv, s_tmp, K, tid);
    else for (int k = 0; k < K; ++k) s_tmp[k * NCP + tid] = 0.0;
    s_ps[tid] = dps;
  }
  __syncthreads();

  const double dt = pr.delta_t, rc = pr.robert_coeff, raw = pr.raw_filter_coeff, xi = pr.xi;
  const int LpC = a.LpC;
  for (int idx = tid; idx < RP * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const int p = p0 + r;
    if (p >= g.T) continue;
    const size_t e = (size_t)p * K + k;
    const int mi = g.row_m[p];
    const int m = g.m_of[mi];
    const int n = t.row_n[p];
    // implicit_correction tail: dt_ts = dt_ts + xi*dt_ts_temp
    double2 dT = a.w_T[e];
    if (a.use_implicit) { dT.x = dT.x + xi * s_tmp[k * NCP + 2 * r]; dT.y = dT.y + xi * s_tmp[k * NCP + 2 * r + 1]; }
    double2 dD = make_double2(s_ddiv[k * NCP + 2 * r], s_ddiv[k * NCP + 2 * r + 1]);
    double2 dV = a.dt_vors[e];
    // damping (spectral_damping.F90:172-291) against the previous time level
    const double2 vp = a.vors_prev[e], dp_ = a.divs_prev[e], tp = a.ts_prev[e];
    {
      const double dv = t.damping_vor[p], cv = 1.0 / (1.0 + dv * dt);
      dV.x = cv * (dV.x - dv * vp.x); dV.y = cv * (dV.y - dv * vp.y);
      const double dd = t.damping_div[p], cd = 1.0 / (1.0 + dd * dt);
      dD.x = cd * (dD.x - dd * dp_.x); dD.y = cd * (dD.y - dd * dp_.y);
      const double dtm = t.damping[p], ct = 1.0 / (1.0 + dtm * dt);
      dT.x = ct * (dT.x - dtm * tp.x); dT.y = ct * (dT.y - dtm * tp.y);
      if (k == 0) {     // sponges act on the top level only (:220-232, :273-285)
        if (m != 0) {
          const double es = t.eddy_sponge[p];
          dV.x = (dV.x - es * vp.x) / (1. + es * dt); dV.y = (dV.y - es * vp.y) / (1. + es * dt);
          dD.x = (dD.x - es * dp_.x) / (1. + es * dt); dD.y = (dD.y - es * dp_.y) / (1. + es * dt);
        } else {
          const double zu = t.zmu_sponge[p], zv = t.zmv_sponge[p];
          dV.x = (dV.x - zu * vp.x) / (1. + zu * dt); dV.y = (dV.y - zu * vp.y) / (1. + zu * dt);
          dD.x = (dD.x - zv * dp_.x) / (1. + zv * dt); dD.y = (dD.y - zv * dp_.y) / (1. + zv * dt);
        }
      }
    }
    (void)n;
    // leapfrog_2level_A: part = a(prev) - 2 a(cur); future = a(prev) + dt*dt_a; cur += rc*part*raw
    const double2 vc = a.vors_cur[e], dc = a.divs_cur[e], tc = a.ts_cur[e];
    double2 vf, df, tf, vcn, dcn, tcn;
    vf = make_double2(vp.x + dt * dV.x, vp.y + dt * dV.y);
    df = make_double2(dp_.x + dt * dD.x, dp_.y + dt * dD.y);
    tf = make_double2(tp.x + dt * dT.x, tp.y + dt * dT.y);
    vcn = make_double2(vc.x + rc * (vp.x - 2.0 * vc.x) * raw, vc.y + rc * (vp.y - 2.0 * vc.y) * raw);
    dcn = make_double2(dc.x + rc * (dp_.x - 2.0 * dc.x) * raw, dc.y + rc * (dp_.y - 2.0 * dc.y) * raw);
    tcn = make_double2(tc.x + rc * (tp.x - 2.0 * tc.x) * raw, tc.y + rc * (tp.y - 2.0 * tc.y) * raw);
    if (a.fuse_robert_b) {   // leapfrog_2level_B: a(cur) += rc*a(fut)*raw; the fixers' (0,0) increments are patched in later
      vcn.x = vcn.x + rc * vf.x * raw; vcn.y = vcn.y + rc * vf.y * raw;
      dcn.x = dcn.x + rc * df.x * raw; dcn.y = dcn.y + rc * df.y * raw;
      tcn.x = tcn.x + rc * tf.x * raw; tcn.y = tcn.y + rc * tf.y * raw;
    }
    a.vors_cur_w[e] = vcn; a.divs_cur_w[e] = dcn; a.ts_cur_w[e] = tcn;
    a.vors_fut[e] = vf; a.divs_fut[e] = df; a.ts_fut[e] = tf;
    a.specC[(size_t)p * LpC + a.cVor + k] = vf;
    a.specC[(size_t)p * LpC + a.cDiv + k] = df;
    a.specC[(size_t)p * LpC + a.cT + k] = tf;
    if (a.keep_tend) { a.k_dt_vors[e] = dV; a.k_dt_divs[e] = dD; a.k_dt_ts[e] = dT; }
  }
  if (tid < RP && p0 + tid < g.T) {
    const int p = p0 + tid;
    double2 dl = a.w_lnps[p];
    if (a.use_implicit) { dl.x = dl.x + xi * s_ps[2 * tid] / pr.ref_ps; dl.y = dl.y + xi * s_ps[2 * tid + 1] / pr.ref_ps; }
    const double2 lp = a.lnps_prev[p], lc = a.lnps_cur[p];
    const double2 lf = make_double2(lp.x + dt * dl.x, lp.y + dt * dl.y);
    double2 lcn = make_double2(lc.x + rc * (lp.x - 2.0 * lc.x) * raw, lc.y + rc * (lp.y - 2.0 * lc.y) * raw);
    if (a.fuse_robert_b) { lcn.x = lcn.x + rc * lf.x * raw; lcn.y = lcn.y + rc * lf.y * raw; }
    a.lnps_cur_w[p] = lcn;
    a.lnps_fut[p] = lf;
    a.specC[(size_t)p * LpC + a.cLnps] = lf;
    if (a.keep_tend) a.k_dt_lnps[p] = dl;
  }
}

void launch_spec_step(const DevTables& t, const Params& pr, const SpecStepArgs& a, cudaStream_t st) {
  const GeomDev& g = t.g;
  const int K = g.K;
  const int nb = (g.T + RP - 1) / RP;
  {
    size_t smem = sizeof(double) * ((size_t)4 * K * NCP + 2 * NC + (size_t)NLC * K);
    static size_t attr = 0;
    if (smem > attr) { cudaFuncSetAttribute(spec_tend_adjust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
    spec_tend_adjust_kernel<<<nb, 128, smem, st>>>(t, pr, a);
  }
  if (a.use_implicit) launch_wave_matvec(t, a.w_div, st);
  {
    size_t smem = sizeof(double) * ((size_t)2 * K * NCP + NC + (size_t)NLC * K);
    static size_t attr = 0;
    if (smem > attr) { cudaFuncSetAttribute(spec_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
    spec_update_kernel<<<nb, 128, smem, st>>>(t, pr, a);
  }
  spec_post_kernel<<<g.T, 64, 0, st>>>(t, a.specC, a.LpC, K, a.cVor, a.cDiv, a.cU, a.cV, a.cT, a.cLnps,
                                       a.cDxT, a.cDyT, a.cDxL, a.cDyL);
}

// ---------------------------------------------------------------------------------------------
// leapfrog_2level_B (raw_filter_coeff == 1): a(prev_new) += rc * a(cur_new)
// ---------------------------------------------------------------------------------------------
__global__ void spec_robert_b_kernel(double2* __restrict__ aprev, const double2* __restrict__ acur, size_t n, double rc, double raw) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    double2 p = aprev[i], c = acur[i];
    p.x = p.x + rc * c.x * raw; p.y = p.y + rc * c.y * raw;
    aprev[i] = p;
  }
}
void launch_spec_robert_b(double2* aprev, const double2* acur, size_t n, double rc, double raw, cudaStream_t st) {
  spec_robert_b_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(aprev, acur, n, rc, raw);
}

// ---------------------------------------------------------------------------------------------
// Stand-alone implicit_correction (model/implicit.F90:241-325) on packed spectra [T][K], for the transforms_mod-level ABI
// (isca_b200_implicit_correction).  The time step runs the fused kernels above; these two kernels restate the routine one
// packed row per thread (real and imaginary parts are independent real-linear problems) and share spec_wave_matvec_kernel.
// ---------------------------------------------------------------------------------------------
struct ImplArgs {
  double2 *dt_divs, *dt_ts, *dt_lnps;
  const double2 *divs_prev, *divs_cur, *ts_prev, *ts_cur, *lnps_prev, *lnps_cur;
};

// dt_t(k) of linear_tp_tendency (implicit.F90:414-480) for one level: `before` = sum of dmean above level k, `total` = column sum
__device__ __forceinline__ double impl_tp_level(const DevTables& t, const Params& pr, int k, int K, double before, double dmean, double total) {
  const double dp = t.dpk[k] + t.dbk[k] * pr.ref_ps, dp_inv = 1 / dp;
  const double dlog_1 = t.ref_ln_p_half[k + 1] - t.ref_ln_p_full[k], dlog_3 = t.ref_ln_p_half[k + 1] - t.ref_ln_p_half[k];
  double dtt = -pr.kappa * t.ref_t[k] * (before * dlog_3 + dmean * dlog_1) * dp_inv;
  // vert_vel(k) = -before + total*bk(k) (1 <= k <= K-1), vert_vel(k+1) = -(before + dmean) + total*bk(k+1); temp = -vert_vel*(t(k)-t(k-1))
  double temp_k = 0.0, temp_k1 = 0.0;
  if (k >= 1) temp_k = -((-before) + total * t.bk[k]) * (t.ref_t[k] - t.ref_t[k - 1]);
  if (k + 1 <= K - 1) temp_k1 = -((-(before + dmean)) + total * t.bk[k + 1]) * (t.ref_t[k + 1] - t.ref_t[k]);
  return dtt + 0.5 * dp_inv * (temp_k1 + temp_k);
}

__global__ void impl_adjust_kernel(DevTables t, Params pr, ImplArgs a) {
  const GeomDev& g = t.g;
  const int K = g.K;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * g.T) return;
  const int row = idx >> 1, ri = idx & 1;
  auto C = [&](const double2* p, int k) { return reinterpret_cast<const double*>(p)[((size_t)row * K + k) * 2 + ri]; };
  auto W = [&](double2* p, int k) -> double& { return reinterpret_cast<double*>(p)[((size_t)row * K + k) * 2 + ri]; };
  auto C2 = [&](const double2* p) { return reinterpret_cast<const double*>(p)[(size_t)row * 2 + ri]; };
  // adjust_dt_divs: linear_tp_tendency(divs(previous) - divs(current))
  double total = 0.0;
  for (int k = 0; k < K; ++k) total = total + (C(a.divs_prev, k) - C(a.divs_cur, k)) * (t.dpk[k] + t.dbk[k] * pr.ref_ps);
  double before = 0.0;
  for (int k = 0; k < K; ++k) {
    const double dmean = (C(a.divs_prev, k) - C(a.divs_cur, k)) * (t.dpk[k] + t.dbk[k] * pr.ref_ps);
    W(a.dt_ts, k) = W(a.dt_ts, k) + impl_tp_level(t, pr, k, K, before, dmean, total);
    before = before + dmean;
  }
  double& dlp = reinterpret_cast<double*>(a.dt_lnps)[(size_t)row * 2 + ri];
  dlp = dlp + (-total) / pr.ref_ps;
  const double ps_temp = C2(a.lnps_prev) - C2(a.lnps_cur) + pr.xi * dlp;
  // linear_geopotential(ts_temp, 0, 0) bottom-up (:329-359) and the divergence-tendency adjustment
  const double eig = t.eigen[row];
  double gh = 0.0;                                             // geopot_half(k+1)
  for (int k = K - 1; k >= 0; --k) {
    const double ts_temp = C(a.ts_prev, k) - C(a.ts_cur, k) + pr.xi * W(a.dt_ts, k);
    const double lh1 = t.ref_ln_p_half[k + 1], lh0 = t.ref_ln_p_half[k], lf = t.ref_ln_p_full[k];
    const double geo = gh + pr.rdgas * (ts_temp * (lh1 - lf));
    W(a.dt_divs, k) = W(a.dt_divs, k) + eig * (geo + t.h_impl[k] * ps_temp * pr.ref_ps);
    if (k >= 1) gh = gh + pr.rdgas * (ts_temp * (lh1 - lh0));
  }
}

__global__ void impl_back_kernel(DevTables t, Params pr, ImplArgs a) {
  const GeomDev& g = t.g;
  const int K = g.K;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * g.T) return;
  const int row = idx >> 1, ri = idx & 1;
  auto W = [&](double2* p, int k) -> double& { return reinterpret_cast<double*>(p)[((size_t)row * K + k) * 2 + ri]; };
  double total = 0.0;
  for (int k = 0; k < K; ++k) total = total + W(a.dt_divs, k) * (t.dpk[k] + t.dbk[k] * pr.ref_ps);
  double before = 0.0;
  for (int k = 0; k < K; ++k) {
    const double dmean = W(a.dt_divs, k) * (t.dpk[k] + t.dbk[k] * pr.ref_ps);
    W(a.dt_ts, k) = W(a.dt_ts, k) + pr.xi * impl_tp_level(t, pr, k, K, before, dmean, total);
    before = before + dmean;
  }
  double& dlp = reinterpret_cast<double*>(a.dt_lnps)[(size_t)row * 2 + ri];
  dlp = dlp + pr.xi * (-total) / pr.ref_ps;
}

void launch_implicit_correction(const DevTables& t, const Params& pr, double2* dt_divs, double2* dt_ts, double2* dt_lnps,
                                const double2* divs_prev, const double2* divs_cur, const double2* ts_prev, const double2* ts_cur,
                                const double2* lnps_prev, const double2* lnps_cur, cudaStream_t st) {
  const GeomDev& g = t.g;
  const int K = g.K;
  ImplArgs a{dt_divs, dt_ts, dt_lnps, divs_prev, divs_cur, ts_prev, ts_cur, lnps_prev, lnps_cur};
  const int nb = (2 * g.T + 127) / 128;
  impl_adjust_kernel<<<nb, 128, 0, st>>>(t, pr, a);
  launch_wave_matvec(t, dt_divs, st);
  impl_back_kernel<<<nb, 128, 0, st>>>(t, pr, a);
}

}  // namespace isca
