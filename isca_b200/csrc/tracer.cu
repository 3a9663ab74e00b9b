// tracer.cu -- grid tracer path (parts of a14/a20/a22).
//
//   tracer_horiz_kernel    ONE kernel for the horizontal step of update_tracers (model/spectral_dynamics.F90:1150-1160):
//                          hs_forcing tracer_source_sink (atmos_param/hs_forcing/hs_forcing.F90:248-265,683-724),
//                          tr_future = q(prev) + dt*dt_tr (:1155), then a_grid_horiz_advection_3d (model/fv_advection.F90:126-200):
//                          divergence term, semi_x_3d / semi_y_3d half steps of advection_sphere_3d (:241-284,379-433),
//                          vanleer_x_3d (:330-375, incl. integer_flux_x :483-521, slope_x :446-479, find_cell_x :427-442),
//                          vanleer_sphere_3d (:288-326, slope_sphere :525-550).  A CTA owns a band of R latitude rows of one
//                          level; the R+4 rows it needs of tr0 and q1 = q + semi_x(q) and the R rows of q2 and of the x-fluxes
//                          live in shared memory, so none of the reference's intermediate 3-D arrays touches HBM.
//   tracer_ppm_kernel      vert_advection FINITE_VOLUME_PARABOLIC, advective form
//                          (atmos_shared/vert_advection/vert_advection.F90:297-478, slope_z :504-563, compute_weights :567-629)
//                          in ONE sweep with a sliding window (no edge-value arrays), + leapfrog part A for the grid tracer
//                          (spectral_dynamics.F90:1165-1169), + the column integrals of the water fixer
//   tracer_water_*         water fixer of compute_corrections (spectral_dynamics.F90:1245-1278) + leapfrog_2level_B
// nranks > 1: one exchange of the 2 edge rows of (tr0, u, v) with each latitude neighbour (tracer_halo_pack + grouped
// ncclSend/ncclRecv in core.cu) replaces the reference's three mpp_update_domains calls (fv_advection.F90:161-162,189,196).
#include "tracer.h"
#include "grid.h"

namespace isca {

__device__ __forceinline__ double sign1(double x) { return x >= 0.0 ? 1.0 : -1.0; }   // sign(1.0, x)
__device__ __forceinline__ double min3(double a, double b, double c) { return fmin(fmin(a, b), c); }
__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

// tr_future before advection: q(prev) + dt*(dt_tr + tracer_source_sink) at element e of level k, column col
__device__ __forceinline__ double tracer_tr0(const DevTables& t, const Params& pr, const TracerArgs& a, int k, size_t e, size_t col) {
  const double r = a.q_prev[e];
  double rdt = a.dt_q_in ? a.dt_q_in[e] : 0.0;
  if (a.physics_on && !pr.no_forcing) {
    // rst = rm + dt*rdt (rdt = 0 on entry); source = flux/pmass in the lowest layer; sink = rdamp*rst
    double source = 0.0;
    const int K = t.g.K;
    if (k == K - 1) { const double ps_c = a.ps_cur[col]; source = a.trflux / ((t.pk[K] + t.bk[K] * ps_c) - (t.pk[K - 1] + t.bk[K - 1] * ps_c)); }
    rdt = rdt + (source - a.trdamp * r);
  }
  return r + a.delta_t * rdt;
}

// ---------------------------------------------------------------------------------------------
// edge rows of tr0, u, v packed for the two latitude neighbours: send_s = local rows 0,1 ; send_n = rows Jloc-2, Jloc-1
// layout [3][K][2][I]
// ---------------------------------------------------------------------------------------------
__global__ void tracer_halo_pack_kernel(DevTables t, Params pr, TracerArgs a) {
  const GeomDev& g = t.g;
  const int I = g.I, Jloc = g.Jloc, K = g.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, fk = blockIdx.z;     // r: 0,1 south rows; 2,3 north rows
  if (i >= I) return;
  const int f = fk / K, k = fk - f * K;
  const int jl = (r < 2) ? r : (Jloc - 4 + r);
  const size_t col = (size_t)jl * I + i, e = (size_t)k * Jloc * I + col;
  const double v = (f == 0) ? tracer_tr0(t, pr, a, k, e, col) : (f == 1 ? a.u_cur[e] : a.v_cur[e]);
  double* dst = (r < 2) ? a.send_s : a.send_n;
  dst[((size_t)fk * 2 + (r & 1)) * I + i] = v;
}
void launch_tracer_halo_pack(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, 4, 3 * t.g.K);
  tracer_halo_pack_kernel<<<grid, 128, 0, st>>>(t, pr, a);
}

// ---------------------------------------------------------------------------------------------
// horizontal step.  Band row r (0 <= r < R+4) is the local latitude row jb - 2 + r.  Rows beyond a pole are the polar mirror
// rows of fv_advection.F90:164-178,266-280 (row 0 <- row 1 at i + nx/2, row -1 <- row 2, ...): they are loaded from the
// mirrored interior row with the longitude shift (semi_x commutes with the shift, so q1 on a mirror row is the mirror of q1);
// rows owned by a neighbouring rank come from the halo buffers.
// ---------------------------------------------------------------------------------------------
struct BandRow {
  int kind;        // 0 interior (or mirrored interior), 1 southern halo, 2 northern halo
  int row;         // local row (kind 0) or halo row 0/1
  int jg;          // global 0-based row whose metric (cos) applies
  int shift;       // longitude shift (I/2 on mirror rows)
  double sign;     // -1 for the meridional wind on mirror rows
};

__device__ __forceinline__ BandRow band_row(const GeomDev& g, int jl) {
  BandRow b; b.shift = 0; b.sign = 1.0;
  int jg = g.j0 + jl;
  if (jg < 0) { jg = -1 - jg; b.shift = g.I / 2; b.sign = -1.0; }
  else if (jg >= g.J) { jg = 2 * g.J - 1 - jg; b.shift = g.I / 2; b.sign = -1.0; }
  b.jg = jg;
  const int l = jg - g.j0;
  if (l >= 0 && l < g.Jloc) { b.kind = 0; b.row = l; }
  else if (l < 0) { b.kind = 1; b.row = l + 2; }
  else { b.kind = 2; b.row = l - g.Jloc; }
  return b;
}

// field f (0 tr0 [only halo kinds], 1 u, 2 v) of level k at band row b, longitude i (already shifted)
__device__ __forceinline__ double band_load(const GeomDev& g, const TracerArgs& a, const double* __restrict__ main, int f, int k,
                                            const BandRow& b, int i) {
  if (b.kind == 0) return main[((size_t)k * g.Jloc + b.row) * g.I + i];
  const double* h = (b.kind == 1) ? a.halo_s : a.halo_n;
  return h[(((size_t)f * g.K + k) * 2 + b.row) * g.I + i];
}

// R rows per band, NPT longitudes per thread (blockDim.x * NPT == I, I a power of two).  The kernel is bound by instruction
// issue and load latency, not by HBM bytes, so: all global loads of a phase are issued before their first use (fully unrolled
// row / element loops into register arrays); every per-row reciprocal comes from a host table; interior bands (EDGE = false:
// no pole, no rank boundary inside the R+4 rows) skip the mirror / halo bookkeeping; each meridional face flux is evaluated
// once per column (it is the north flux of one row and the south flux of the next) with the upwind stencil selected
// arithmetically, so that a warp never executes both branches of the limiter.
template <int R, int NPT, bool EDGE>
__device__ __forceinline__ void tracer_horiz_body(const DevTables& t, const FvTables& f, const Params& pr, const TracerArgs& a, double* sm) {
  const GeomDev& g = t.g;
  const int I = g.I, J = g.J, Jloc = g.Jloc, K = g.K;
  const int k = blockIdx.y, jb = blockIdx.x * R;
  const int nt = blockDim.x, tid = threadIdx.x, imask = I - 1;
  double* s_q = sm;                         // [R+4][I]  tr0
  double* s_q1 = s_q + (R + 4) * I;         // [R+4][I]  q + semi_x(q, dt/2)
  double* s_q2 = s_q1 + (R + 4) * I;        // [R][I]    q + semi_y(q, dt/2)
  double* s_fx = s_q2 + R * I;              // [R][I]    x flux at the west face
  const double dt = a.delta_t, dth = 0.5 * a.delta_t;
  const size_t lev = (size_t)k * Jloc * I;
  const double* __restrict__ qprev = a.q_prev + lev;
  const double* __restrict__ ucur = a.u_cur + lev;
  const double* __restrict__ vcur = a.v_cur + lev;
  const double* __restrict__ dtq = a.dt_q_in ? a.dt_q_in + lev : nullptr;

  // band row r -> where it lives.  Interior bands: local row jb-2+r.
  BandRow br[R + 4];
#pragma unroll
  for (int r = 0; r < R + 4; ++r) {
    if constexpr (EDGE) {
      int jl = jb - 2 + r;
      if (jl > Jloc + 1) jl = Jloc + 1;     // rows past the end of a ragged last band: harmless duplicates
      br[r] = band_row(g, jl);
    } else { br[r].kind = 0; br[r].row = jb - 2 + r; br[r].jg = g.j0 + jb - 2 + r; br[r].shift = 0; br[r].sign = 1.0; }
  }
  // element offset inside a level of band row r at longitude i (EDGE: shifted / halo rows handled by ld())
  auto ld = [&](const double* __restrict__ main, int fidx, int r, int i) -> double {
    if constexpr (!EDGE) return main[br[r].row * I + i];
    else {
      const int is = (i + br[r].shift) & imask;
      if (br[r].kind == 0) return main[br[r].row * I + is];
      const double* h = (br[r].kind == 1) ? a.halo_s : a.halo_n;
      return h[((fidx * K + k) * 2 + br[r].row) * I + is];
    }
  };

  // ---- phase 1: raw loads of q_prev (+ optional tendency) and u on all band rows, then tr0 -> shared memory
  double uq[R + 4][NPT];
  {
    double qv[R + 4][NPT], dq[R + 4][NPT];
    const bool forcing = a.physics_on && !pr.no_forcing;
#pragma unroll
    for (int r = 0; r < R + 4; ++r)
#pragma unroll
      for (int e = 0; e < NPT; ++e) {
        const int i = tid + e * nt;
        qv[r][e] = ld(qprev, 0, r, i);                  // halo rows: tr0 itself (packed by tracer_halo_pack_kernel)
        uq[r][e] = ld(ucur, 1, r, i);
        dq[r][e] = 0.0;
        if (dtq && br[r].kind == 0) dq[r][e] = ld(dtq, 0, r, i);
      }
    const bool bottom = forcing && (k == K - 1);
#pragma unroll
    for (int r = 0; r < R + 4; ++r)
#pragma unroll
      for (int e = 0; e < NPT; ++e) {
        double v = qv[r][e];
        if (br[r].kind == 0) {
          // tracer_tr0: rst = rm + dt*rdt; source = flux/pmass in the lowest layer; sink = rdamp*rst
          double rdt = dq[r][e];
          if (forcing) {
            double source = 0.0;
            if (bottom) {
              const double ps_c = a.ps_cur[br[r].row * I + ((tid + e * nt + br[r].shift) & imask)];
              source = a.trflux / ((t.pk[K] + t.bk[K] * ps_c) - (t.pk[K - 1] + t.bk[K - 1] * ps_c));
            }
            rdt = rdt + (source - a.trdamp * v);
          }
          v = v + a.delta_t * rdt;
        }
        s_q[r * I + tid + e * nt] = v;
      }
  }
  // meridional wind on the band rows and their two neighbours (held for phases 2 and 4); mirror rows carry the sign flip
  double vq[R + 2][NPT];
#pragma unroll
  for (int r = 0; r < R + 2; ++r)
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const double v = ld(vcur, 2, r + 1, tid + e * nt);
      if constexpr (EDGE) vq[r][e] = br[r + 1].sign * v; else vq[r][e] = v;
    }
  __syncthreads();
  // ---- phase 2: q1 on all band rows, q2 on the R rows of the band
#pragma unroll
  for (int r = 0; r < R + 4; ++r) {
    const double dthx = dth * f.rdxc[br[r].jg];
    const double* q = s_q + r * I;
    const bool band = (r >= 2 && r < R + 2);
    const int jf = br[r].jg + 1;
    const double dty_m = band ? dth * f.rdyy[jf - 1] : 0.0, dty_p = band ? dth * f.rdyy[jf] : 0.0;
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * nt;
      const double qc = q[i];
      // semi_x: b = ua*dt/(dx*c); source cells i_left = i-1-floor(b), i_right = i_left+1 (1-based, wrapped)
      const double bq = uq[r][e] * dthx;
      const double fl = floor(bq);
      const int il = (i - 1 - (int)fl) & imask;          // 0-based i_left
      const int ir = (il + 1) & imask;
      const double bb = bq - fl;
      s_q1[r * I + i] = qc + (bb * q[il] + (1.0 - bb) * q[ir] - qc);
      if (band) {
        // semi_y (interior rows: no shift, sign +1)
        const double v = vq[r - 1][e];
        const double dq = (v >= 0.0) ? v * dty_m * (q[i - I] - qc) : v * dty_p * (qc - q[i + I]);
        s_q2[(r - 2) * I + i] = qc + dq;
      }
    }
  }
  // u(i-1), u(i+1) of the band rows (interior rows: never shifted, never halo)
  double uw[R][NPT], ue[R][NPT];
#pragma unroll
  for (int r2 = 0; r2 < R; ++r2)
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * nt;
      const int rowoff = (EDGE ? min(jb + r2, Jloc - 1) : jb + r2) * I;
      uw[r2][e] = ucur[rowoff + ((i - 1) & imask)];
      ue[r2][e] = ucur[rowoff + ((i + 1) & imask)];
    }
  __syncthreads();
  // ---- phase 3: vanleer_x fluxes at the west face of every cell of the band rows
#pragma unroll
  for (int r2 = 0; r2 < R; ++r2) {
    const double* row = s_q2 + r2 * I;
    const double dtx = dt * f.rdxc[br[r2 + 2].jg];
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * nt;
      // uc(i) = 0.5 (ua(i-1) + ua(i)); b = uc*dt/(dx*c)
      const double uc = 0.5 * (uw[r2][e] + uq[r2 + 2][e]);
      const double bq = uc * dtx;
      const double bt = trunc(bq);
      const double bb = bq - bt;
      double flux = 0.0;
      const int n = (int)bt;                            // integer_flux_x: whole cells crossed
      if (n != 0) {
        if (n >= 1) {
          // sum(q(i-n : i-1)) in the reference's order (ascending index, wrapped part after the unwrapped one)
          const int i1 = (i + 1);                         // 1-based i
          if (i1 - n >= 1) { for (int c = i1 - n; c <= i1 - 1; ++c) flux += row[c - 1]; }
          else { double s1 = 0.0, s2 = 0.0; for (int c = 1; c <= i1 - 1; ++c) s1 += row[c - 1]; for (int c = i1 - n + I; c <= I; ++c) s2 += row[c - 1]; flux = s1 + s2; }
        } else {
          const int i1 = (i + 1);
          if (i1 - 1 - n <= I) { double s1 = 0.0; for (int c = i1; c <= i1 - 1 - n; ++c) s1 += row[c - 1]; flux = -s1; }
          else { double s1 = 0.0, s2 = 0.0; for (int c = i1; c <= I; ++c) s1 += row[c - 1]; for (int c = 1; c <= i1 - 1 - n - I; ++c) s2 += row[c - 1]; flux = -s1 - s2; }
        }
      }
      // find_cell_x: ii = i-1-floor(b) (1-based), wrapped
      const int c0 = (i - 1 - (int)floor(bq)) & imask, cm = (c0 - 1) & imask, cp = (c0 + 1) & imask;
      const double qq = row[c0], qm = row[cm], qp = row[cp];
      const double slope = ((qp - qq) + (qq - qm)) * 0.5;
      const double ss = sign1(slope) * min3(fabs(slope), 2.0 * (qq - min3(qm, qq, qp)), 2.0 * (max3(qm, qq, qp) - qq));
      s_fx[r2 * I + i] = flux + bb * (qq + 0.5 * ss * (sign1(bb) - bb));
    }
  }
  __syncthreads();
  // ---- phase 4: divergence term, x-flux difference, vanleer_sphere fluxes on q1 (one evaluation per face and column)
  const double rdt = 1.0 / dt;
#pragma unroll
  for (int e = 0; e < NPT; ++e) {
    const int i = tid + e * nt, ip = (i + 1) & imask;
    double q1c[R + 4];
#pragma unroll
    for (int r = 0; r < R + 4; ++r) q1c[r] = s_q1[r * I + i];
    // face fi lies south of band row fi+2 (global face index = that row's 0-based latitude index)
    double ff[R + 1], vcc[R + 1];
#pragma unroll
    for (int fi = 0; fi <= R; ++fi) {
      const int jg = br[fi + 2].jg;                     // for the face north of the last band row: br[R+2] (halo / mirror row)
      const int jface = EDGE ? ((fi == R) ? br[R + 1].jg + 1 : jg) : jg;
      const double v = 0.5 * (vq[fi][e] + vq[fi + 1][e]);                      // vc(j)
      const double cc = f.cc[jface];
      vcc[fi] = v * cc;
      const bool up = (v >= 0.0);
      // upwind row: 1-based index jface (v >= 0) or jface+1 (v < 0)
      const double a0 = up ? q1c[fi] : q1c[fi + 1], b0 = up ? q1c[fi + 1] : q1c[fi + 2], c0 = up ? q1c[fi + 2] : q1c[fi + 3];
      const int jr = up ? jface : jface + 1;
      const double sl = (c0 - b0) * f.dy_plus[jr] + (b0 - a0) * f.dy_minus[jr];
      const double lim = sign1(sl) * min3(fabs(sl), 2.0 * (b0 - min3(a0, b0, c0)), 2.0 * (max3(a0, b0, c0) - b0));
      const double dtdy = dt * f.rdy[jr + 1];           // dt / dy(jr)
      const double sg = up ? 1.0 : -1.0;
      double fl = vcc[fi] * (b0 + sg * 0.5 * lim * (1.0 - sg * dtdy * v));
      if (jface == 0 || jface == J) fl = 0.0;
      ff[fi] = fl;
    }
#pragma unroll
    for (int r2 = 0; r2 < R; ++r2) {
      const int jl = jb + r2;
      if (EDGE && jl >= Jloc) break;
      const int jg = g.j0 + jl;
      const double rcdy = f.rcdy[jg], rcdx = f.rdxc[jg];
      const double u0 = uq[r2 + 2][e];
      const double uc_w = 0.5 * (uw[r2][e] + u0), uc_e = 0.5 * (u0 + ue[r2][e]);
      double div = (vcc[r2 + 1] - vcc[r2]) * rcdy;
      div = div + (uc_e - uc_w) * rcdx;
      const double qc = s_q[(r2 + 2) * I + i];
      double dq = qc * div;
      dq = dq - (s_fx[r2 * I + ip] - s_fx[r2 * I + i]) * rdt;
      dq = dq - rcdy * (ff[r2 + 1] - ff[r2]);
      a.tr1[lev + (size_t)jl * I + i] = qc + dt * dq;   // tr_future = tr_future + delta_t*dt_tr
    }
  }
}

template <int R, int NPT>
__global__ void __launch_bounds__(256, 2)
tracer_horiz_kernel(DevTables t, FvTables f, Params pr, TracerArgs a) {
  extern __shared__ double sm[];
  const int jb = blockIdx.x * R;
  const bool edge = (jb < 2) || (jb + R + 2 > t.g.Jloc);
  if (edge) tracer_horiz_body<R, NPT, true>(t, f, pr, a, sm);
  else tracer_horiz_body<R, NPT, false>(t, f, pr, a, sm);
}

template <int R, int NPT>
static void launch_horiz_t(const DevTables& t, const FvTables& f, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  const int I = t.g.I;
  const size_t smem = sizeof(double) * (size_t)(4 * R + 8) * I;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(tracer_horiz_kernel<R, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  dim3 grid((t.g.Jloc + R - 1) / R, t.g.K);
  tracer_horiz_kernel<R, NPT><<<grid, I / NPT, smem, st>>>(t, f, pr, a);
}
void launch_tracer_horiz(const DevTables& t, const FvTables& f, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  const int I = t.g.I;                                   // a power of two >= 32 (checked at create time)
  if (I <= 256) launch_horiz_t<4, 1>(t, f, pr, a, st);
  else if (I == 512) launch_horiz_t<4, 2>(t, f, pr, a, st);
  else if (I == 1024) launch_horiz_t<2, 4>(t, f, pr, a, st);
  else launch_horiz_t<1, 8>(t, f, pr, a, st);            // I == 2048: 12 rows x 16 KB
}

// ---------------------------------------------------------------------------------------------
// PPM vertical advection, one thread per column, one sweep.  The limited parabola edge values of levels k and k+1 are carried
// in a sliding window (they are all the flux at interface k+1 needs when the vertical Courant number is <= 1); the rare
// Courant > 1 extension recomputes the edges of the level it lands in.  Layer thicknesses come from pk/bk.
// The same sweep forms the column integrals the water fixer needs:
//   part[0] = w * sum_k (q_prev + dt*dt_tr) dp(ps_prev)                     (initialize_corrections, :1327-1335)
//   part[1..4] = sum_k q_fut*dpk, sum_k q_fut*dbk over the levels with p_full(current) >= water_correction_limit, and the same
//   over the other levels: mass_weighted_global_integral(q_fut [*mask], ps_fut) = part_A + ps_fut*part_B once ps_fut is known
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double p_full_level(const DevTables& t, const Params& pr, int k, double ps) {
  if (pr.sigma_fast) return t.sig.pf[k] * ps;
  // two logs of level k (press_and_geopot.F90:170-192)
  const double ph = t.pk[k] + t.bk[k] * ps, ph1 = t.pk[k + 1] + t.bk[k + 1] * ps;
  const double l1 = log(ph1);
  double lf;
  if (k == 0 && pr.pkbk0_zero) lf = l1 + (-1.0);
  else { const double l0 = log(ph); lf = l1 - (1.0 - ph * (l1 - l0) / (ph1 - ph)); }
  return exp(lf);
}

// SIG: pure-sigma fast path (pk == 0): every layer thickness is dbk*ps, so the slope / interface weights of slope_z and
// compute_weights are ratios of dbk only (host tables, SigmaTables::ppm_*) and 1/dp = rdb/ps -- no division per level.
template <bool SIG>
struct PpmCol {
  const DevTables& t; const double* __restrict__ rp; size_t plane; double ps, rps; int K;
  __device__ __forceinline__ double R(int k) const { return __ldg(rp + (size_t)k * plane); }
  __device__ __forceinline__ double DZ(int k) const {                                       // dp = p_half(k+1) - p_half(k)
    if constexpr (SIG) return t.sig.db[k] * ps;
    else return (t.pk[k + 1] + t.bk[k + 1] * ps) - (t.pk[k] + t.bk[k] * ps);
  }
  __device__ __forceinline__ double RDZ(int k) const { if constexpr (SIG) return t.sig.rdb[k] * rps; else return 1.0 / DZ(k); }
  // slope_z(linear=.false., limit=.true.) :504-563 for an interior level
  __device__ __forceinline__ double slope(int k, double rm1, double r0, double rp1) const {
    double sl;
    if constexpr (SIG) sl = (rp1 - r0) * t.sig.ppm_c1[k] + (r0 - rm1) * t.sig.ppm_c2[k];
    else {
      const double dm = DZ(k - 1), d0 = DZ(k), dp = DZ(k + 1);
      const double gk = (r0 - rm1) / (d0 + dm), gk1 = (rp1 - r0) / (dp + d0);
      sl = (gk1 * (2. * dm + d0) + gk * (2. * dp + d0)) * d0 / (dm + d0 + dp);
    }
    const double rmin = min3(rm1, r0, rp1), rmax = max3(rm1, r0, rp1);
    return sign1(sl) * min3(fabs(sl), 2. * (r0 - rmin), 2. * (rmax - r0));
  }
  // interface value between levels k-1 and k (:300-325 with compute_weights :567-629), 2 <= k <= K-2
  __device__ __forceinline__ double iface(int k, double rm1, double r0, double slp_k, double slp_km1) const {
    if constexpr (SIG) return rm1 + t.sig.ppm_z1[k] * (r0 - rm1) - t.sig.ppm_z2[k] * slp_k + t.sig.ppm_z3[k] * slp_km1;
    else {
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
    }
  }
  __device__ __forceinline__ double slope_at(int k) const { return (k >= 1 && k <= K - 2) ? slope(k, R(k - 1), R(k), R(k + 1)) : 0.0; }
  __device__ __forceinline__ double iface_at(int k) const { return iface(k, R(k - 1), R(k), slope_at(k), slope_at(k - 1)); }
  // limited edge values of level k from the unlimited ones (Colella-Woodward limiter :340-356)
  __device__ __forceinline__ void limit(int k, double r0, double& rl, double& rr) const {
    if ((rr - r0) * (r0 - rl) <= 0.0) { rl = r0; rr = r0; }
    if (k != 0 && k != K - 1) {
      const double rm = rr - rl;
      const double aa = rm * (r0 - 0.5 * (rr + rl));
      const double bq = rm * rm * (1.0 / 6.);
      if (aa > bq) rl = 3.0 * r0 - 2.0 * rr;
      if (aa < -bq) rr = 3.0 * r0 - 2.0 * rl;
    }
  }
  // slow path: edges of an arbitrary level (Courant number > 1 extension)
  __device__ __noinline__ void edges_slow(int k, double& rl, double& rr) const {
    const double r0 = R(k), s = slope_at(k);
    if (k == 0 || k == K - 1) { rl = r0 - 0.5 * s; rr = r0 + 0.5 * s; }
    else {
      rl = (k == 1) ? r0 - 0.5 * s : iface_at(k);
      rr = (k == K - 2) ? r0 + 0.5 * s : iface_at(k + 1);
    }
    limit(k, r0, rl, rr);
  }
};

template <bool SIG>
__global__ void __launch_bounds__(128)
tracer_ppm_kernel(DevTables t, Params pr, TracerArgs a) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y;
  if (i >= g.I) return;
  const int K = g.K;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  const double ps_c = a.ps_cur[col], ps_p = a.ps_prev[col];
  const double dt = a.delta_t;
  const PpmCol<SIG> c{t, a.tr1 + col, plane, ps_c, 1.0 / ps_c, K};
  auto W = [&](int k) { return __ldg(a.wg + (size_t)k * plane + col); };
  const double tt = 2. / 3.;
  const double rc = a.robert_coeff, raw = a.raw_filter_coeff;
  const bool forcing = a.physics_on && !pr.no_forcing;
  const double src_bottom = forcing ? a.trflux / ((t.pk[K] + t.bk[K] * ps_c) - (t.pk[K - 1] + t.bk[K - 1] * ps_c)) : 0.0;

  // sliding window: r(k..k+3), slopes and interface values one level ahead, edges of levels k and k+1
  double r_0 = c.R(0), r_p1 = c.R(1), r_p2 = c.R(2);
  double slp_k1, if_k1;
  double rl_k, rr_k, rl_k1 = 0.0, rr_k1 = 0.0;
  // edges of level 0 (slope 0)
  slp_k1 = c.slope(1, r_0, r_p1, r_p2);
  if_k1 = 0.0;
  rl_k = r_0; rr_k = r_0;

  double w_k = W(0);
  double flux_above = w_k * r_0;                        // flux(ks) = w(ks)*r(ks)
  double vi = 0.0, a_c = 0.0, b_c = 0.0, a_n = 0.0, b_n = 0.0;
  // loads run one level ahead of the arithmetic (software pipeline: two levels of loads in flight per thread)
  double n_r3 = (3 < K) ? c.R(3) : 0.0, n_w1 = W(1), n_qp = a.q_prev[col], n_qc = a.q_cur[col];
  double n_dtq = a.dt_q_in ? a.dt_q_in[col] : 0.0;
  for (int k = 0; k < K; ++k) {
    const size_t e = (size_t)k * plane + col;
    const double r_p3 = n_r3, w_k1 = n_w1, qp = n_qp, qc = n_qc, dtq = n_dtq;
    if (k + 1 < K) {
      n_r3 = (k + 4 < K) ? c.R(k + 4) : 0.0;
      n_w1 = W(k + 2);
      n_qp = a.q_prev[e + plane]; n_qc = a.q_cur[e + plane];
      if (a.dt_q_in) n_dtq = a.dt_q_in[e + plane];
    }
    // ---- stage A: edges of level k+1 (needs r(k+3))
    if (k + 1 < K) {
      const int ka = k + 1;
      const double slp_k2 = (ka + 1 <= K - 2) ? c.slope(ka + 1, r_p1, r_p2, r_p3) : 0.0;
      const double if_k2 = (ka + 1 >= 2 && ka + 1 <= K - 2) ? c.iface(ka + 1, r_p1, r_p2, slp_k2, slp_k1) : 0.0;
      if (ka == K - 1) { rl_k1 = r_p1 - 0.5 * slp_k1; rr_k1 = r_p1 + 0.5 * slp_k1; }
      else {
        rl_k1 = (ka == 1) ? r_p1 - 0.5 * slp_k1 : if_k1;
        rr_k1 = (ka == K - 2) ? r_p1 + 0.5 * slp_k1 : if_k2;
      }
      c.limit(ka, r_p1, rl_k1, rr_k1);
      slp_k1 = slp_k2; if_k1 = if_k2;
    }
    // ---- stage B: flux at interface k+1 (:360-425) and advective-form tendency of level k (:466-476)
    const double rdz_k = c.RDZ(k);
    double flux_below;
    if (k == K - 1) flux_below = w_k1 * r_0;            // flux(ke+1) = w(ke+1)*r(ke)
    else {
      const int kf = k + 1;                             // interface index
      const double wk = w_k1;
      double cn, xx, rst, rsum = 0.0;
      if (wk >= 0.) {
        cn = dt * wk * rdz_k;
        int kk = kf - 1;
        double rkk = r_0, rrk = rr_k, rlk = rl_k;
        if (cn > 1.) {
          double dzsum = 0.0; const double dtw = dt * wk;
          while (dzsum + c.DZ(kk) < dtw) { if (kk == 0) break; dzsum += c.DZ(kk); rsum += c.R(kk); kk = kk - 1; }
          xx = (dtw - dzsum) / c.DZ(kk);
          if (kk != k) { rkk = c.R(kk); c.edges_slow(kk, rlk, rrk); }
        } else xx = cn;
        const double rm = rrk - rlk;
        double r6 = 6.0 * (rkk - 0.5 * (rrk + rlk));
        if (kk == 0) r6 = 0.;
        rst = rrk - 0.5 * xx * (rm - (1.0 - tt * xx) * r6);
        if (cn > 1.) rst = (xx * rst + rsum) / cn;
      } else {
        cn = -dt * wk * c.RDZ(kf);
        int kk = kf;
        double rkk = r_p1, rrk = rr_k1, rlk = rl_k1;
        if (cn > 1.) {
          double dzsum = 0.0; const double dtw = -dt * wk;
          while (dzsum + c.DZ(kk) < dtw) { if (kk == 0) break; dzsum += c.DZ(kk); rsum += c.R(kk); kk = kk + 1; if (kk >= K) { kk = K - 1; break; } }
          xx = (dtw - dzsum) / c.DZ(kk);
          if (kk != k + 1) { rkk = c.R(kk); c.edges_slow(kk, rlk, rrk); }
        } else xx = cn;
        const double rm = rrk - rlk;
        double r6 = 6.0 * (rkk - 0.5 * (rrk + rlk));
        if (kk == K - 1) r6 = 0.;
        rst = rlk + 0.5 * xx * (rm + (1.0 - tt * xx) * r6);
        if (cn > 1.) rst = (xx * rst + rsum) / cn;
      }
      flux_below = wk * rst;
    }
    const double rdt = -(flux_below - flux_above - r_0 * (w_k1 - w_k)) * rdz_k;
    // water integral of initialize_corrections: (q_prev + dt*dt_tr) with psg(previous); dt_tr as in tracer_tr0
    {
      double rd = dtq;
      if (forcing) rd = rd + (((k == K - 1) ? src_bottom : 0.0) - a.trdamp * qp);
      vi = vi + (qp + dt * rd) * ((t.pk[k + 1] + t.bk[k + 1] * ps_p) - (t.pk[k] + t.bk[k] * ps_p));
    }
    // leapfrog part A for the grid tracer (:1165-1169): current += rc*(previous - 2 current)*raw.
    // `future` shares its storage slot with `previous` (two time levels): read before the future value is written.
    const double qf = r_0 + dt * rdt;                   // tr_future + delta_t*dt_tmp
    a.q_fut[e] = qf;
    a.q_cur_w[e] = qc + rc * (qp - 2.0 * qc) * raw;
    const double da = t.pk[k + 1] - t.pk[k], db = t.bk[k + 1] - t.bk[k];
    if (p_full_level(t, pr, k, ps_c) >= a.water_limit) { a_c = a_c + qf * da; b_c = b_c + qf * db; }
    else { a_n = a_n + qf * da; b_n = b_n + qf * db; }
    flux_above = flux_below; w_k = w_k1;
    r_0 = r_p1; r_p1 = r_p2; r_p2 = r_p3;
    rl_k = rl_k1; rr_k = rr_k1;
  }
  a.part[col] = t.wts_lat[g.j0 + jl] * vi;
  a.wpart[col] = a_c; a.wpart[plane + col] = b_c; a.wpart[2 * plane + col] = a_n; a.wpart[3 * plane + col] = b_n;
}
void launch_tracer_ppm(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st) {
  dim3 grid((t.g.I + 127) / 128, t.g.Jloc);
  if (pr.sigma_fast) tracer_ppm_kernel<true><<<grid, 128, 0, st>>>(t, pr, a);
  else tracer_ppm_kernel<false><<<grid, 128, 0, st>>>(t, pr, a);
}

// ---------------------------------------------------------------------------------------------
// water fixer: three mass-weighted integrals with psg(future) (all, p_full >= limit, p_full < limit) from the column sums of
// the PPM sweep; p_full is the `current` level's (the array passed to compute_corrections, spectral_dynamics.F90:1011)
// ---------------------------------------------------------------------------------------------
// scal: the device scalars (SC_W_PREV; SC_W_ALL / SC_W_CORR / SC_W_NOT formed by apply_fixers).  8 levels per thread: the correction factor
// (a dozen divisions) is formed once per thread and the loads of the 8 levels are in flight together.
constexpr int WA_LEV = 8;
__global__ void __launch_bounds__(256)
tracer_water_apply_kernel(DevTables t, Params pr, TracerArgs a, const double* __restrict__ scal, double denom, int do_water) {
  const GeomDev& g = t.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, jl = blockIdx.y, k0 = blockIdx.z * WA_LEV;
  if (i >= g.I) return;
  const size_t col = (size_t)jl * g.I + i, plane = (size_t)g.Jloc * g.I;
  double qf[WA_LEV], qc[WA_LEV];
#pragma unroll
  for (int d = 0; d < WA_LEV; ++d)
    if (k0 + d < g.K) { const size_t e = (size_t)(k0 + d) * plane + col; qf[d] = a.q_fut[e]; qc[d] = a.q_cur_w[e]; }
  const double ps_c = a.ps_cur[col];
  const double mean_prev = scal[SC_W_PREV] / denom / pr.grav, mean_tmp = scal[SC_W_ALL] / denom / pr.grav;
  const double corr = scal[SC_W_CORR] / denom / pr.grav, ncorr = scal[SC_W_NOT] / denom / pr.grav;
  double wf = 1.0;
  const bool apply = do_water && (mean_tmp > 0.);
  if (apply) { wf = mean_prev / mean_tmp; wf = wf * (1. + ncorr / corr) - ncorr / corr; }
  const double rc = a.robert_coeff, raw = a.raw_filter_coeff;
#pragma unroll
  for (int d = 0; d < WA_LEV; ++d) {
    const int k = k0 + d;
    if (k >= g.K) break;
    const size_t e = (size_t)k * plane + col;
    double q = qf[d];
    if (apply && p_full_level(t, pr, k, ps_c) >= a.water_limit) { q = wf * q; a.q_fut[e] = q; }
    a.q_cur_w[e] = qc[d] + rc * q * raw;              // leapfrog_2level_B for the grid tracer
  }
}
void launch_tracer_water_apply(const DevTables& t, const Params& pr, const TracerArgs& a, const double* scal, double denom,
                               int do_water, cudaStream_t st) {
  dim3 grid((t.g.I + 255) / 256, t.g.Jloc, (t.g.K + WA_LEV - 1) / WA_LEV);
  tracer_water_apply_kernel<<<grid, t.g.I < 256 ? t.g.I : 256, 0, st>>>(t, pr, a, scal, denom, do_water);
}

}  // namespace isca
