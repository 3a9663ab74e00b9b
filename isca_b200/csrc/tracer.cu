// tracer.cu -- grid tracer path (parts of a14/a20/a22).
//
//   tracer_source_kernel   hs_forcing tracer_source_sink (atmos_param/hs_forcing/hs_forcing.F90:248-265,683-724),
//                          tr_future = q(prev) + dt*dt_tr (model/spectral_dynamics.F90:1155), and the water part of
//                          initialize_corrections (:1327-1335)
//   tracer_semi_kernel     semi_x_3d / semi_y_3d half steps of advection_sphere_3d (model/fv_advection.F90:241-284,379-433)
//   tracer_flux_kernel     a_grid_horiz_advection_3d divergence term (:126-200), vanleer_x_3d (:330-375, incl.
//                          integer_flux_x :483-521, slope_x :446-479, find_cell_x :427-442), vanleer_sphere_3d (:288-326,
//                          slope_sphere :525-550)
//   tracer_ppm_kernel      vert_advection FINITE_VOLUME_PARABOLIC, advective form
//                          (atmos_shared/vert_advection/vert_advection.F90:297-478, slope_z :504-563, compute_weights :567-629)
//                          + leapfrog part A for the grid tracer (spectral_dynamics.F90:1165-1169)
//   tracer_water_*         water fixer of compute_corrections (spectral_dynamics.F90:1245-1278) + leapfrog_2level_B
// nranks > 1: one exchange of the 2 edge rows of (tr0, u, v) with each latitude neighbour (tracer_halo_pack + grouped
// ncclSend/ncclRecv in core.cu) replaces the reference's three mpp_update_domains calls (fv_advection.F90:161-162,189,196).
#include "tracer.h"
#include "grid.h"

namespace isca {

__device__ __forceinline__ double sign1(double x) { return x >= 0.0 ? 1.0 : -1.0; }   // sign(1.0, x)
__device__ __forceinline__ double min3(double a, double b, double c) { return fmin(fmin(a, b), c); }
__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

// ---------------------------------------------------------------------------------------------
__global__ void tracer_source_kernel(DevTables t, Params pr, TracerArgs a) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const int K = g.K;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double ps_c = a.ps_cur[col], ps_p = a.ps_prev[col];
  double vi = 0.0;
  for (int k = 0; k < K; ++k) {
    const size_t e = (size_t)k * plane + col;
    const double r = a.q_prev[e];
    double rdt = a.dt_q_in ? a.dt_q_in[e] : 0.0;
    if (a.physics_on && !pr.no_forcing) {
      // rst = rm + dt*rdt (rdt = 0 on entry); source = flux/pmass in the lowest layer; sink = rdamp*rst
      double source = 0.0;
      if (k == K - 1) source = a.trflux / ((t.pk[K] + t.bk[K] * ps_c) - (t.pk[K - 1] + t.bk[K - 1] * ps_c));
      rdt = rdt + (source - a.trdamp * r);
    }
    const double trf = r + a.delta_t * rdt;
    a.tr0[e] = trf;
    const double dpp = (t.pk[k + 1] + t.bk[k + 1] * ps_p) - (t.pk[k] + t.bk[k] * ps_p);
    vi = vi + trf * dpp;                               // mass_weighted_global_integral(q_prev + dt*dt_tr, psg(previous))
  }
  a.part[col] = t.wts_lat[g.j0 + jl] * vi;
}
void launch_tracer_source(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  tracer_source_kernel<<<grid, 128, 0, st>>>(t, pr, a);
}

// A [K][Jloc][I] field together with its 2-row latitude halos [K][2][I] (null on a single rank)
struct HaloField { const double* main; const double* s; const double* n; };

// row jl (local, -2 <= jl < Jloc+2) of level k
__device__ __forceinline__ const double* halo_row(const HaloField& F, int k, int jl, int I, int Jloc) {
  if (jl >= 0 && jl < Jloc) return F.main + ((size_t)k * Jloc + jl) * I;
  if (jl < 0) return F.s + ((size_t)k * 2 + (jl + 2)) * I;
  return F.n + ((size_t)k * 2 + (jl - Jloc)) * I;
}

// value at Fortran latitude index jf in [-1, J+2] (1-based, global), longitude i (0-based), with the polar mirror
// rows (fv_advection.F90:164-178): row 0 <- row 1 at i+nx/2, row -1 <- row 2, row J+1 <- row J, ...; rows owned by a
// neighbouring rank come from the halo (mpp_update_domains of fv_advection.F90:161-162)
__device__ __forceinline__ double at_lat(const HaloField& F, int k, int jf, int i, const GeomDev& g, double pole_sign) {
  const int I = g.I, J = g.J;
  if (jf >= 1 && jf <= J) return halo_row(F, k, jf - 1 - g.j0, I, g.Jloc)[i];
  int ii = i + I / 2; if (ii >= I) ii -= I;
  const int jm = (jf < 1) ? (1 - jf) : (2 * J + 1 - jf);       // 0 -> 1, -1 -> 2, J+1 -> J, J+2 -> J-1
  return pole_sign * halo_row(F, k, jm - 1 - g.j0, I, g.Jloc)[ii];
}

// ---------------------------------------------------------------------------------------------
// edge rows of tr0, u, v packed for the two latitude neighbours: send_s = local rows 0,1 ; send_n = rows Jloc-2, Jloc-1
// ---------------------------------------------------------------------------------------------
__global__ void tracer_halo_pack_kernel(DevTables t, TracerArgs a) {
  const GeomDev& g = t.g;
  const int I = g.I, Jloc = g.Jloc, K = g.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, fk = blockIdx.z;     // r: 0,1 south rows; 2,3 north rows
  if (i >= I) return;
  const int f = fk / K, k = fk - f * K;
  const double* src = (f == 0) ? a.tr0 : (f == 1 ? a.u_cur : a.v_cur);
  const int jl = (r < 2) ? r : (Jloc - 4 + r);
  const double v = src[((size_t)k * Jloc + jl) * I + i];
  double* dst = (r < 2) ? a.send_s : a.send_n;
  dst[((size_t)fk * 2 + (r & 1)) * I + i] = v;
}
void launch_tracer_halo_pack(const DevTables& t, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, 4, 3 * t.g.K);
  tracer_halo_pack_kernel<<<grid, 128, 0, st>>>(t, a);
}

// ---------------------------------------------------------------------------------------------
// q1 = q + semi_x(q, dt/2), q2 = q + semi_y(q, dt/2).  With latitude halos q1 is also evaluated on the two halo rows on
// either side (semi_x only needs the row itself), which replaces the reference's second halo exchange of q1.
// ---------------------------------------------------------------------------------------------
__global__ void tracer_semi_kernel(DevTables t, FvTables f, TracerArgs a, int ext) {
  const GeomDev& g = t.g;
  const int I = g.I, Jloc = g.Jloc, K = g.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.z;
  const int jl = (int)blockIdx.y - ext;                    // local row, -2 .. Jloc+1 when ext == 2
  const int j = g.j0 + jl;                                 // global row, 0-based
  if (i >= I || j < 0 || j >= g.J) return;                 // beyond a pole: the mirror rows are read from the interior
  const HaloField Q{a.tr0, a.halo_s, a.halo_n};
  const HaloField U{a.u_cur, a.halo_s ? a.halo_s + (size_t)K * 2 * I : nullptr, a.halo_n ? a.halo_n + (size_t)K * 2 * I : nullptr};
  const double* qrow = halo_row(Q, k, jl, I, Jloc);
  const double dt = 0.5 * a.delta_t;
  const double qc = qrow[i];
  const bool interior = (jl >= 0 && jl < Jloc);
  // semi_x: b = ua*dt/(dx*c); cell ii = i-1-floor(b) (1-based, wrapped)
  {
    const double b = halo_row(U, k, jl, I, Jloc)[i] * dt / (f.dx * f.c[j]);
    const double fl = floor(b);
    int il = (i + 1) - 1 - (int)fl;                    // 1-based i_left
    while (il > I) il -= I;
    while (il < 1) il += I;
    int ir = il + 1; if (ir > I) ir = 1;
    const double bb = b - fl;
    const double q1v = qc + (bb * qrow[il - 1] + (1.0 - bb) * qrow[ir - 1] - qc);
    if (interior) a.q1[((size_t)k * Jloc + jl) * I + i] = q1v;
    else if (jl < 0) a.q1_halo_s[((size_t)k * 2 + (jl + 2)) * I + i] = q1v;
    else a.q1_halo_n[((size_t)k * 2 + (jl - Jloc)) * I + i] = q1v;
  }
  if (!interior) return;
  // semi_y
  {
    const size_t e = ((size_t)k * Jloc + jl) * I + i;
    const double v = a.v_cur[e];
    const int jf = j + 1;
    double dq;
    if (v >= 0.0) dq = v * dt * (at_lat(Q, k, jf - 1, i, g, 1.0) - qc) / f.dyy[jf - 1];
    else dq = v * dt * (qc - at_lat(Q, k, jf + 1, i, g, 1.0)) / f.dyy[jf];
    a.q2[e] = qc + dq;
  }
}
void launch_tracer_semi(const DevTables& t, const FvTables& f, const TracerArgs& a, cudaStream_t st) {
  const int ext = (t.g.P > 1) ? 2 : 0;
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc + 2 * ext, t.g.K);
  tracer_semi_kernel<<<grid, 128, 0, st>>>(t, f, a, ext);
}

// ---------------------------------------------------------------------------------------------
// one CTA per (latitude row, level): the row of q2 and the x-fluxes live in shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double slope_x_at(const double* __restrict__ row, int c, int I) {   // c 0-based cell
  const int cm = (c == 0) ? I - 1 : c - 1, cp = (c == I - 1) ? 0 : c + 1;
  const double q = row[c], qm = row[cm], qp = row[cp];
  const double slope = ((qp - q) + (q - qm)) / 2;
  return sign1(slope) * min3(fabs(slope), 2.0 * (q - min3(qm, q, qp)), 2.0 * (max3(qm, q, qp) - q));
}

__device__ __forceinline__ double slope_sphere_at(const HaloField& Q1, int k, const FvTables& f, int jf, int i, const GeomDev& g) {
  const double qm = at_lat(Q1, k, jf - 1, i, g, 1.0), q = at_lat(Q1, k, jf, i, g, 1.0), qp = at_lat(Q1, k, jf + 1, i, g, 1.0);
  const double slope = (qp - q) * f.dy_plus[jf] + (q - qm) * f.dy_minus[jf];
  return sign1(slope) * min3(fabs(slope), 2.0 * (q - min3(qm, q, qp)), 2.0 * (max3(qm, q, qp) - q));
}

__global__ void tracer_flux_kernel(DevTables t, FvTables f, TracerArgs a) {
  extern __shared__ double sm[];
  const GeomDev& g = t.g;
  const int I = g.I, J = g.J, Jloc = g.Jloc, K = g.K;
  const int jl = blockIdx.x, k = blockIdx.y;          // local row, 0-based
  const int j = g.j0 + jl;                            // global row, 0-based
  const int jf = j + 1;
  double* row = sm;                                   // q2 row [I]
  double* flx = sm + I;                               // x flux at the west face of cell i [I]
  const size_t base = ((size_t)k * Jloc + jl) * I;
  const double dt = a.delta_t;
  const double cj = f.c[j];
  for (int i = threadIdx.x; i < I; i += blockDim.x) row[i] = a.q2[base + i];
  __syncthreads();
  const double* ua = a.u_cur + base;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    // vanleer_x: uc(i) = 0.5 (ua(i-1) + ua(i)); b = uc*dt/(dx*c)
    const int im = (i == 0) ? I - 1 : i - 1;
    const double uc = 0.5 * (ua[im] + ua[i]);
    const double b = uc * dt / (f.dx * cj);
    const double bt = trunc(b);
    const double bb = b - bt;
    double flux = 0.0;
    const int n = (int)bt;                            // integer_flux_x: whole cells crossed
    if (n >= 1) {
      // sum(q(i-n : i-1)) in the reference's order (ascending index, wrapped part after the unwrapped one)
      const int i1 = (i + 1);                         // 1-based i
      if (i1 - n >= 1) { for (int c = i1 - n; c <= i1 - 1; ++c) flux += row[c - 1]; }
      else { double s1 = 0.0, s2 = 0.0; for (int c = 1; c <= i1 - 1; ++c) s1 += row[c - 1]; for (int c = i1 - n + I; c <= I; ++c) s2 += row[c - 1]; flux = s1 + s2; }
    } else if (n <= -1) {
      const int i1 = (i + 1);
      if (i1 - 1 - n <= I) { double s1 = 0.0; for (int c = i1; c <= i1 - 1 - n; ++c) s1 += row[c - 1]; flux = -s1; }
      else { double s1 = 0.0, s2 = 0.0; for (int c = i1; c <= I; ++c) s1 += row[c - 1]; for (int c = 1; c <= i1 - 1 - n - I; ++c) s2 += row[c - 1]; flux = -s1 - s2; }
    }
    // find_cell_x: ii = i-1-floor(b) (1-based), wrapped
    int ii = (i + 1) - 1 - (int)floor(b);
    while (ii > I) ii -= I;
    while (ii < 1) ii += I;
    const double qq = row[ii - 1];
    const double ss = slope_x_at(row, ii - 1, I);
    flx[i] = flux + bb * (qq + 0.5 * ss * (sign1(bb) - bb));
  }
  __syncthreads();
  const double* q = a.tr0 + base;
  const HaloField Q1{a.q1, a.q1_halo_s, a.q1_halo_n};
  const HaloField V{a.v_cur, a.halo_s ? a.halo_s + (size_t)2 * K * 2 * I : nullptr, a.halo_n ? a.halo_n + (size_t)2 * K * 2 * I : nullptr};
  const double* va = a.v_cur + base;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const int ip = (i == I - 1) ? 0 : i + 1, im = (i == 0) ? I - 1 : i - 1;
    // divergence term: dq_dt = q*div
    const double vc_s = 0.5 * (at_lat(V, k, jf - 1, i, g, -1.0) + va[i]);         // vc(j)
    const double vc_n = 0.5 * (va[i] + at_lat(V, k, jf + 1, i, g, -1.0));         // vc(j+1)
    const double uc_w = 0.5 * (ua[im] + ua[i]), uc_e = 0.5 * (ua[i] + ua[ip]);
    double div = (vc_n * f.cc[jf] - vc_s * f.cc[jf - 1]) / (cj * f.dy[jf + 1]);
    div = div + (uc_e - uc_w) / (cj * f.dx);
    const double qc = q[i];
    double dq = 0.0 + qc * div;
    // vanleer_x
    dq = dq - (flx[ip] - flx[i]) / dt;
    // vanleer_sphere on q1: fluxes at the south (j) and north (j+1) faces
    double fl_s = 0.0, fl_n = 0.0;
    if (jf > 1) {
      const double v = vc_s;
      if (v >= 0.0) fl_s = v * f.cc[jf - 1] * (at_lat(Q1, k, jf - 1, i, g, 1.0) + 0.5 * slope_sphere_at(Q1, k, f, jf - 1, i, g) * (1.0 - (dt / f.dy[jf]) * v));
      else fl_s = v * f.cc[jf - 1] * (at_lat(Q1, k, jf, i, g, 1.0) - 0.5 * slope_sphere_at(Q1, k, f, jf, i, g) * (1.0 + (dt / f.dy[jf + 1]) * v));
    }
    if (jf < J) {
      const double v = vc_n;
      if (v >= 0.0) fl_n = v * f.cc[jf] * (at_lat(Q1, k, jf, i, g, 1.0) + 0.5 * slope_sphere_at(Q1, k, f, jf, i, g) * (1.0 - (dt / f.dy[jf + 1]) * v));
      else fl_n = v * f.cc[jf] * (at_lat(Q1, k, jf + 1, i, g, 1.0) - 0.5 * slope_sphere_at(Q1, k, f, jf + 1, i, g) * (1.0 + (dt / f.dy[jf + 2]) * v));
    }
    dq = dq - (1.0 / (f.dy[jf + 1] * cj)) * (fl_n - fl_s);
    a.tr1[base + i] = qc + dt * dq;                   // tr_future = tr_future + delta_t*dt_tr
  }
}
void launch_tracer_flux(const DevTables& t, const FvTables& f, const TracerArgs& a, cudaStream_t st) {
  dim3 grid(t.g.Jloc, t.g.K);
  const int threads = t.g.I < 256 ? t.g.I : 256;
  tracer_flux_kernel<<<grid, threads, sizeof(double) * 2 * t.g.I, st>>>(t, f, a);
}

// ---------------------------------------------------------------------------------------------
// PPM vertical advection, one thread per column.  Nothing is kept in thread-local arrays: the limited parabola edge values
// go through two scratch planes (q1, q2 of the horizontal step, free by now), slopes and interface values are produced by a
// sliding window, layer thicknesses are recomputed from pk/bk, and the profile itself is re-read from tr1 (L1/L2 hits).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tracer_ppm_kernel(DevTables t, Params pr, TracerArgs a) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const int K = g.K;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double ps_c = a.ps_cur[col];
  const double dt = a.delta_t;
  const double* __restrict__ rp = a.tr1 + col;
  double* __restrict__ rlp = a.q1 + col;
  double* __restrict__ rrp = a.q2 + col;
  auto R = [&](int k) { return __ldg(rp + (size_t)k * plane); };
  auto DZ = [&](int k) { return (t.pk[k + 1] + t.bk[k + 1] * ps_c) - (t.pk[k] + t.bk[k] * ps_c); };     // dp = p_half(k+1) - p_half(k)
  // slope_z(linear=.false., limit=.true.) :504-563 for an interior level
  auto slope = [&](int k, double rm1, double r0, double rp1) {
    const double dm = DZ(k - 1), d0 = DZ(k), dp = DZ(k + 1);
    const double gk = (r0 - rm1) / (d0 + dm), gk1 = (rp1 - r0) / (dp + d0);
    const double sl = (gk1 * (2. * dm + d0) + gk * (2. * dp + d0)) * d0 / (dm + d0 + dp);
    const double rmin = min3(rm1, r0, rp1), rmax = max3(rm1, r0, rp1);
    return sign1(sl) * min3(fabs(sl), 2. * (r0 - rmin), 2. * (rmax - r0));
  };
  // interface value between levels k-1 and k (:300-325 with compute_weights :567-629), 2 <= k <= K-2
  auto iface = [&](int k, double rm1, double r0, double slp_k, double slp_km1) {
    const double d2 = DZ(k - 2), d1 = DZ(k - 1), d0 = DZ(k), dp = DZ(k + 1);
    const double denom1 = 1.0 / (d1 + d0);
    const double denom2 = 1.0 / (d2 + d1 + d0 + dp);
    const double denom3 = 1.0 / (2 * d1 + d0);
    const double denom4 = 1.0 / (d1 + 2 * d0);
    const double num3 = d2 + d1, num4 = d0 + dp;
    const double x = num3 * denom3 - num4 * denom4;
    const double y = 2.0 * d1 * d0;
    const double z0 = d1 * denom1;
    const double z1 = z0 + x * y * denom1 * denom2;
    const double z2 = d1 * num3 * denom3 * denom2;
    const double z3 = d0 * num4 * denom4 * denom2;
    return rm1 + z1 * (r0 - rm1) - z2 * slp_k + z3 * slp_km1;
  };
  // ---- pass 1: limited edge values of every level
  {
    double r_m1 = 0.0, r_0 = R(0), r_p1 = R(1);
    double slp_k = 0.0;                                  // slp(0) = 0
    double if_k = 0.0;                                   // interface value at the top of level k (valid for 2 <= k <= K-2)
    for (int k = 0; k < K; ++k) {
      const double r_p2 = (k + 2 < K) ? R(k + 2) : 0.0;
      const double slp_k1 = (k + 1 >= 1 && k + 1 <= K - 2) ? slope(k + 1, r_0, r_p1, r_p2) : 0.0;
      const double if_k1 = (k + 1 >= 2 && k + 1 <= K - 2) ? iface(k + 1, r_0, r_p1, slp_k1, slp_k) : 0.0;
      double rl, rr;
      if (k == 0 || k == K - 1) { rl = r_0 - 0.5 * slp_k; rr = r_0 + 0.5 * slp_k; }
      else {
        rl = (k == 1) ? r_0 - 0.5 * slp_k : if_k;
        rr = (k == K - 2) ? r_0 + 0.5 * slp_k : if_k1;
      }
      // Colella-Woodward limiter :340-356
      if ((rr - r_0) * (r_0 - rl) <= 0.0) { rl = r_0; rr = r_0; }
      if (k != 0 && k != K - 1) {
        const double rm = rr - rl;
        const double aa = rm * (r_0 - 0.5 * (rr + rl));
        const double bq = rm * rm / 6.;
        if (aa > bq) rl = 3.0 * r_0 - 2.0 * rr;
        if (aa < -bq) rr = 3.0 * r_0 - 2.0 * rl;
      }
      rlp[(size_t)k * plane] = rl; rrp[(size_t)k * plane] = rr;
      r_m1 = r_0; r_0 = r_p1; r_p1 = r_p2; slp_k = slp_k1; if_k = if_k1;
    }
    (void)r_m1;
  }
  auto RL = [&](int k) { return rlp[(size_t)k * plane]; };
  auto RR = [&](int k) { return rrp[(size_t)k * plane]; };
  auto W = [&](int k) { return __ldg(a.wg + (size_t)k * plane + col); };
  // ---- pass 2: fluxes at interfaces :360-425 and advective-form tendency :466-476
  const double tt = 2. / 3.;
  double w_k = W(0), r_k = R(0);
  double flux_above = w_k * r_k;                        // flux(ks) = w(ks)*r(ks)
  const double rc = a.robert_coeff, raw = a.raw_filter_coeff;
  for (int k = 0; k < K; ++k) {
    const double w_k1 = W(k + 1);
    const double r_k1 = (k + 1 < K) ? R(k + 1) : 0.0;
    double flux_below;
    if (k == K - 1) flux_below = w_k1 * r_k;            // flux(ke+1) = w(ke+1)*r(ke)
    else {
      const int kf = k + 1;                             // interface index
      const double wk = w_k1;
      double cn, xx, rst, rsum = 0.0;
      int kk;
      if (wk >= 0.) {
        cn = dt * wk / DZ(kf - 1);
        kk = kf - 1;
        if (cn > 1.) {
          double dzsum = 0.0; const double dtw = dt * wk;
          while (dzsum + DZ(kk) < dtw) { if (kk == 0) break; dzsum += DZ(kk); rsum += R(kk); kk = kk - 1; }
          xx = (dtw - dzsum) / DZ(kk);
        } else xx = cn;
        const double rkk = (kk == k) ? r_k : R(kk), rrk = RR(kk), rlk = RL(kk);
        const double rm = rrk - rlk;
        double r6 = 6.0 * (rkk - 0.5 * (rrk + rlk));
        if (kk == 0) r6 = 0.;
        rst = rrk - 0.5 * xx * (rm - (1.0 - tt * xx) * r6);
        if (cn > 1.) rst = (xx * rst + rsum) / cn;
      } else {
        cn = -dt * wk / DZ(kf);
        kk = kf;
        if (cn > 1.) {
          double dzsum = 0.0; const double dtw = -dt * wk;
          while (dzsum + DZ(kk) < dtw) { if (kk == 0) break; dzsum += DZ(kk); rsum += R(kk); kk = kk + 1; if (kk >= K) { kk = K - 1; break; } }
          xx = (dtw - dzsum) / DZ(kk);
        } else xx = cn;
        const double rkk = (kk == k + 1) ? r_k1 : R(kk), rrk = RR(kk), rlk = RL(kk);
        const double rm = rrk - rlk;
        double r6 = 6.0 * (rkk - 0.5 * (rrk + rlk));
        if (kk == K - 1) r6 = 0.;
        rst = rlk + 0.5 * xx * (rm + (1.0 - tt * xx) * r6);
        if (cn > 1.) rst = (xx * rst + rsum) / cn;
      }
      flux_below = wk * rst;
    }
    const double rdt = -(flux_below - flux_above - r_k * (w_k1 - w_k)) / DZ(k);
    const size_t e = (size_t)k * plane + col;
    // leapfrog part A for the grid tracer (:1165-1169): current += rc*(previous - 2 current)*raw.
    // `future` shares its storage slot with `previous` (two time levels): read before the future value is written.
    const double qp = a.q_prev[e], qc = a.q_cur[e];
    a.q_fut[e] = r_k + dt * rdt;                        // tr_future + delta_t*dt_tmp
    a.q_cur_w[e] = qc + rc * (qp - 2.0 * qc) * raw;
    flux_above = flux_below; w_k = w_k1; r_k = r_k1;
  }
}
void launch_tracer_ppm(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  tracer_ppm_kernel<<<grid, 128, 0, st>>>(t, pr, a);
}

// ---------------------------------------------------------------------------------------------
// water fixer: three mass-weighted integrals with psg(future) (all, p_full >= limit, p_full < limit)
// p_full is the `current` level's (the array passed to compute_corrections, spectral_dynamics.F90:1011)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double p_full_level(const DevTables& t, const Params& pr, int k, double ps, double& ln_half_below_io) {
  // bottom-up helper is awkward here; recompute the two logs of level k directly (press_and_geopot.F90:170-192)
  const double ph = t.pk[k] + t.bk[k] * ps, ph1 = t.pk[k + 1] + t.bk[k + 1] * ps;
  const double l1 = log(ph1);
  double lf;
  if (k == 0 && pr.pkbk0_zero) lf = l1 + (-1.0);
  else { const double l0 = log(ph); lf = l1 - (1.0 - ph * (l1 - l0) / (ph1 - ph)); }
  (void)ln_half_below_io;
  return exp(lf);
}

__global__ void tracer_water_colsum_kernel(DevTables t, Params pr, TracerArgs a) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double ps_f = a.ps_fut[col], ps_c = a.ps_cur[col];
  double v_all = 0.0, v_corr = 0.0, v_not = 0.0, dummy = 0.0;
  for (int k = 0; k < g.K; ++k) {
    const double q = a.q_fut[(size_t)k * plane + col];
    const double dp = (t.pk[k + 1] + t.bk[k + 1] * ps_f) - (t.pk[k] + t.bk[k] * ps_f);
    const double pf = p_full_level(t, pr, k, ps_c, dummy);
    const double m1 = (pf >= a.water_limit) ? 1.0 : 0.0, m0 = (pf < a.water_limit) ? 1.0 : 0.0;
    v_all = v_all + q * dp;
    v_corr = v_corr + (q * m1) * dp;
    v_not = v_not + (q * m0) * dp;
  }
  const double w = t.wts_lat[g.j0 + jl];
  a.part[col] = w * v_all; a.part[plane + col] = w * v_corr; a.part[2 * plane + col] = w * v_not;
}
void launch_tracer_water_colsum(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  tracer_water_colsum_kernel<<<grid, 128, 0, st>>>(t, pr, a);
}

// scal: [0] sum water prev, [1..3] sums all / corrected / not corrected (future)
__global__ void tracer_water_apply_kernel(DevTables t, Params pr, TracerArgs a, const double* __restrict__ scal, double denom,
                                          int do_water) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double mean_prev = scal[0] / denom / pr.grav, mean_tmp = scal[1] / denom / pr.grav;
  const double corr = scal[2] / denom / pr.grav, ncorr = scal[3] / denom / pr.grav;
  double wf = 1.0;
  const bool apply = do_water && (mean_tmp > 0.);
  if (apply) { wf = mean_prev / mean_tmp; wf = wf * (1. + ncorr / corr) - ncorr / corr; }
  const double ps_c = a.ps_cur[col];
  const double rc = a.robert_coeff, raw = a.raw_filter_coeff;
  double dummy = 0.0;
  for (int k = 0; k < g.K; ++k) {
    const size_t e = (size_t)k * plane + col;
    double q = a.q_fut[e];
    if (apply && p_full_level(t, pr, k, ps_c, dummy) >= a.water_limit) { q = wf * q; a.q_fut[e] = q; }
    a.q_cur_w[e] = a.q_cur_w[e] + rc * q * raw;      // leapfrog_2level_B for the grid tracer
  }
}
void launch_tracer_water_apply(const DevTables& t, const Params& pr, const TracerArgs& a, const double* scal, double denom,
                               int do_water, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  tracer_water_apply_kernel<<<grid, 128, 0, st>>>(t, pr, a, scal, denom, do_water);
}

}  // namespace isca
