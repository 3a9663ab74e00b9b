// legendre.cu -- per-zonal-wavenumber Legendre transforms as batched fp64 tensor-core GEMMs.
//
// Replaces trans_spherical_to_fourier_3d / trans_fourier_to_spherical_3d
// (atmos_spectral/tools/spherical_fourier.F90:177-261, 264-339).
//
// For one zonal wavenumber m the inverse transform is the dense contraction
//     E[jh][c] = sum_{n even} P(m,n,jh) S(m,n,c)     O[jh][c] = sum_{n odd} P(m,n,jh) S(m,n,c)
//     F(south jh) = E - O,   F(north mirror) = E + O                       (:228-236)
// and the forward transform
//     S(m,n,c) = sum_jh Pw(m,n,jh) * (F_N(jh) + F_S(jh))   n even,   (F_N - F_S) for n odd   (:311-318)
// with c running over (level, re/im) of every field in the batch.  fp64 has no tcgen05 kind
// (SURVEY F4), so the contraction runs on the fp64 tensor pipe through mma.sync m8n8k4 (SASS DMMA).
// Only the triangular range n < M-m+2 is touched (the reference loops over the full rectangle).
#include "device.h"

namespace isca {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Fourier buffer, m-owner side ("layout A"): [(s*nm + mi)*Jloc + jl][C], s = j / Jloc, jl = j % Jloc.
__device__ __forceinline__ size_t fourA_index(const GeomDev& g, int mi, int j, int C) {
  int s = j / g.Jloc, jl = j - s * g.Jloc;
  return ((size_t)(s * g.nm + mi) * g.Jloc + jl) * (size_t)C;
}

// ---------------------------------------------------------------------------------------------
// inverse: spectral -> Fourier
//   CTA tile: JT latitudes (hemisphere index) x CT=32 columns, 4 warps arranged WJ x (4/WJ).
// ---------------------------------------------------------------------------------------------
constexpr int LEG_CT = 32;      // columns (doubles) per CTA tile
constexpr int LEG_KC = 16;      // n rows per parity per k-chunk

template <int JT, int WJ>
__global__ void __launch_bounds__(128)
legendre_inv_kernel(DevTables t, const double2* __restrict__ spec, double* __restrict__ four, int Lp, int ct_begin) {
  constexpr int WC = 4 / WJ;
  constexpr int WTJ = JT / WJ, WTC = LEG_CT / WC;
  constexpr int MT = WTJ / 8, NT = WTC / 8;
  constexpr int PS = JT + 4, SS = LEG_CT + 4;          // smem strides == 4 (mod 16): conflict-free fragments
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  typedef double (*PsT)[2][LEG_KC][PS];
  typedef double (*SsT)[2][LEG_KC][SS];
  PsT Ps = reinterpret_cast<PsT>(leg_smem_raw);                                        // [stage][parity][n][jh]
  SsT Ssm = reinterpret_cast<SsT>(leg_smem_raw + sizeof(double) * 2 * 2 * LEG_KC * PS);  // [stage][parity][n][c]

  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  const int c0 = (ct_begin + blockIdx.x) * LEG_CT;
  const int mi = blockIdx.y;
  const int jt0 = blockIdx.z * JT;
  const int m = g.m_of[mi];
  const int Nm = g.M - m + 2;
  const int row0 = g.off[mi];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wj = warp % WJ, wc = warp / WJ;
  const double* specd = reinterpret_cast<const double*>(spec);

  double acc[2][MT][NT][2];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
      for (int b = 0; b < NT; ++b) { acc[p][a][b][0] = 0.0; acc[p][a][b][1] = 0.0; }

  const int nchunks = (Nm + 2 * LEG_KC - 1) / (2 * LEG_KC);

  auto load_chunk = [&](int stage, int ch) {
    const int nbase = ch * 2 * LEG_KC;
    // P tile: 2*KC rows x JT doubles, 16-byte pieces
    constexpr int PV = JT / 2;                           // double2 per row
    for (int idx = tid; idx < 2 * LEG_KC * PV; idx += 128) {
      int r = idx / PV, v = idx - r * PV;
      int par = r & 1, rr = r >> 1;
      int n = nbase + r;                                 // r = 2*rr + par  -> n parity == par
      double* dst = &Ps[stage][par][rr][2 * v];
      if (n < Nm) cp_async16(dst, t.leg + (size_t)(row0 + n) * g.Jh + jt0 + 2 * v);
      else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    constexpr int SV = LEG_CT / 2;
    for (int idx = tid; idx < 2 * LEG_KC * SV; idx += 128) {
      int r = idx / SV, v = idx - r * SV;
      int par = r & 1, rr = r >> 1;
      int n = nbase + r;
      double* dst = &Ssm[stage][par][rr][2 * v];
      if (n < Nm) cp_async16(dst, specd + (size_t)(row0 + n) * C + c0 + 2 * v);
      else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    cp_async_commit();
  };

  load_chunk(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait_all();
    __syncthreads();
    if (ch + 1 < nchunks) load_chunk((ch + 1) & 1, ch + 1);
    const int st = ch & 1;
    // rows of this chunk that exist: n < Nm  ->  per parity ceil((Nm - nbase)/2), in k-steps of 4
    const int rem = Nm - ch * 2 * LEG_KC;
    const int ksteps = (rem >= 2 * LEG_KC) ? (LEG_KC / 4) : ((rem + 7) >> 3);
#pragma unroll 1
    for (int kk = 0; kk < ksteps; ++kk) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        double a[MT], b[NT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a[mt] = Ps[st][par][kk * 4 + (lane & 3)][wj * WTJ + mt * 8 + (lane >> 2)];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = Ssm[st][par][kk * 4 + (lane & 3)][wc * WTC + nt * 8 + (lane >> 2)];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[par][mt][nt][0], acc[par][mt][nt][1], a[mt], b[nt]);
      }
    }
  }

  // epilogue: south = even - odd, north = even + odd.  On several GPUs the CTA's two JT x 32 output tiles are staged in shared memory
  // (the pipeline buffers are free now) so that every row leaves as ONE 256-byte run written by a whole warp: full NVLink packets for the
  // stores into the lat-owner buffer of the rank that owns the latitude (the fragment layout itself gives 64-byte pieces per row; measured
  // on one GPU the staging costs 0.02 ms, so it is only used for peer stores).  Row stride 40 doubles: the 16-byte fragment stores of a
  // quarter warp fall into distinct banks.
  if (!g.p2p) {         // one GPU: the fragment layout stores 64-byte pieces per row, which HBM takes at full sector efficiency
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int jh = jt0 + wj * WTJ + mt * 8 + (lane >> 2);
      double* rowS = four + fourA_index(g, mi, jh, C);
      double* rowN = four + fourA_index(g, mi, g.J - 1 - jh, C);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = c0 + wc * WTC + nt * 8 + (lane & 3) * 2;
        double e0 = acc[0][mt][nt][0], e1 = acc[0][mt][nt][1];
        double o0 = acc[1][mt][nt][0], o1 = acc[1][mt][nt][1];
        *reinterpret_cast<double2*>(rowS + c) = make_double2(e0 - o0, e1 - o1);
        *reinterpret_cast<double2*>(rowN + c) = make_double2(e0 + o0, e1 + o1);
      }
    }
    return;
  }
  constexpr int OS = LEG_CT + 8;
  __syncthreads();
  double* outS = reinterpret_cast<double*>(leg_smem_raw);
  double* outN = outS + JT * OS;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = wj * WTJ + mt * 8 + (lane >> 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = wc * WTC + nt * 8 + (lane & 3) * 2;
      double e0 = acc[0][mt][nt][0], e1 = acc[0][mt][nt][1];
      double o0 = acc[1][mt][nt][0], o1 = acc[1][mt][nt][1];
      *reinterpret_cast<double2*>(outS + r * OS + c) = make_double2(e0 - o0, e1 - o1);
      *reinterpret_cast<double2*>(outN + r * OS + c) = make_double2(e0 + o0, e1 + o1);
    }
  }
  __syncthreads();
  for (int rr = warp; rr < 2 * JT; rr += 4) {
    const int north = rr >= JT, r = north ? rr - JT : rr;
    const int jh = jt0 + r;
    const int j = north ? g.J - 1 - jh : jh;
    const int sj = j / g.Jloc;                          // peer memory over NVLink
    const size_t prow = (size_t)(g.roff[g.rank] + mi) * g.Jloc;
    double* row = g.peerB[sj] + (prow + (j - sj * g.Jloc)) * (size_t)C;
    row[c0 + lane] = (north ? outN : outS)[r * OS + lane];
  }
}

template <int JT, int WJ>
static void launch_inv_t(const DevTables& t, const double2* spec, double* four, int Lp, cudaStream_t st, int ct_begin, int ct_count) {
  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  const size_t smem = sizeof(double) * 2 * 2 * LEG_KC * ((JT + 4) + (LEG_CT + 4));
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(legendre_inv_kernel<JT, WJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(ct_count < 0 ? C / LEG_CT - ct_begin : ct_count, g.nm, g.Jh / JT);
  legendre_inv_kernel<JT, WJ><<<grid, 128, smem, st>>>(t, spec, four, Lp, ct_begin);
}
void launch_legendre_inv(const DevTables& t, const double2* spec, double* four, int Lp, cudaStream_t st, int ct_begin, int ct_count) {
  const GeomDev& g = t.g;
  if (g.Jh % 64 == 0) launch_inv_t<64, 4>(t, spec, four, Lp, st, ct_begin, ct_count);
  else if (g.Jh % 32 == 0) launch_inv_t<32, 2>(t, spec, four, Lp, st, ct_begin, ct_count);
  else launch_inv_t<16, 1>(t, spec, four, Lp, st, ct_begin, ct_count);
}

// ---------------------------------------------------------------------------------------------
// forward: Fourier -> spectral
//   CTA tile: 32 n rows (16 even + 16 odd) x CT=32 columns; warp = (parity, column half).
//   Reduction over the hemisphere latitudes jh in chunks of 16; the (F_N +- F_S) fold is the prologue.
// ---------------------------------------------------------------------------------------------
constexpr int FWD_NT = 32;      // n rows per CTA tile (both parities)

template <int FWD_KC>
__global__ void __launch_bounds__(128)
legendre_fwd_kernel(DevTables t, const double* __restrict__ four, double2* __restrict__ spec, int Lp,
                    const unsigned char* __restrict__ lev_trunc, int ct_begin) {
  constexpr int XS = LEG_CT + 4;            // 36 == 4 mod 16
  constexpr int WS = FWD_KC + 4;            // == 4 mod 16
  constexpr int NQ = FWD_KC / 8;            // double2 per thread per tile (X: KC x 16, W: 32 x KC/2; 128 threads)
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  typedef double (*XsT)[2][FWD_KC][XS];
  typedef double (*WsT)[2][FWD_NT / 2][WS];
  XsT Xs = reinterpret_cast<XsT>(leg_smem_raw);                                             // [stage][plus/minus][jh][c]
  WsT Ws = reinterpret_cast<WsT>(leg_smem_raw + sizeof(double) * 2 * 2 * FWD_KC * XS);     // [stage][parity][n][jh]

  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  // grid: x = n tile (fastest), y = column tile, z = m.  The n tiles of one (m, column tile) read the same Fourier rows X; launched
  // next to each other they share them through L2 instead of fetching them from HBM once per tile (ncu round 1: 2.4x the algorithmic bytes)
  const int c0 = (ct_begin + blockIdx.y) * LEG_CT;
  const int mi = blockIdx.z;
  const int m = g.m_of[mi];
  const int Nm = g.M - m + 2;
  const int nt0 = blockIdx.x * FWD_NT;
  if (nt0 >= Nm) return;
  const int row0 = g.off[mi];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp & 1, wc = warp >> 1;

  double acc[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

  // register-staged software pipeline: the global loads of chunk ch+1 are in flight while chunk ch is multiplied
  double2 xs_[NQ], xn_[NQ], w_[NQ];
  auto prefetch = [&](int jh0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = tid + q * 128;
      const int r = idx / (LEG_CT / 2), v = idx - r * (LEG_CT / 2);
      const int jh = jh0 + r;
      xs_[q] = *reinterpret_cast<const double2*>(four + fourA_index(g, mi, jh, C) + c0 + 2 * v);
      xn_[q] = *reinterpret_cast<const double2*>(four + fourA_index(g, mi, g.J - 1 - jh, C) + c0 + 2 * v);
      const int rw = idx / (FWD_KC / 2), vw = idx - rw * (FWD_KC / 2);
      const int n = nt0 + rw;
      w_[q] = make_double2(0.0, 0.0);
      if (n < Nm) w_[q] = *reinterpret_cast<const double2*>(t.legw + (size_t)(row0 + n) * g.Jh + jh0 + 2 * vw);
    }
  };
  auto stage = [&](int st) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = tid + q * 128;
      const int r = idx / (LEG_CT / 2), v = idx - r * (LEG_CT / 2);
      Xs[st][0][r][2 * v] = xn_[q].x + xs_[q].x; Xs[st][0][r][2 * v + 1] = xn_[q].y + xs_[q].y;   // x_even (:311)
      Xs[st][1][r][2 * v] = xn_[q].x - xs_[q].x; Xs[st][1][r][2 * v + 1] = xn_[q].y - xs_[q].y;   // x_odd  (:312)
      const int rw = idx / (FWD_KC / 2), vw = idx - rw * (FWD_KC / 2);
      Ws[st][rw & 1][rw >> 1][2 * vw] = w_[q].x; Ws[st][rw & 1][rw >> 1][2 * vw + 1] = w_[q].y;
    }
  };

  const int nch = g.Jh / FWD_KC;
  prefetch(0);
  stage(0);
  __syncthreads();
  for (int ch = 0; ch < nch; ++ch) {
    const int st = ch & 1;
    if (ch + 1 < nch) prefetch((ch + 1) * FWD_KC);
#pragma unroll
    for (int kk = 0; kk < FWD_KC / 4; ++kk) {
      double a[2], b[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[mt] = Ws[st][par][mt * 8 + (lane >> 2)][kk * 4 + (lane & 3)];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) b[nt] = Xs[st][par][kk * 4 + (lane & 3)][wc * 16 + nt * 8 + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    if (ch + 1 < nch) stage(st ^ 1);
    __syncthreads();
  }

  double* specd = reinterpret_cast<double*>(spec);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int n = nt0 + 2 * (mt * 8 + (lane >> 2)) + par;
    if (n >= Nm) continue;
    const bool beyond = (m + n > g.M) || (g.symmetric && m > 0);      // triangle_mask = 0 (spherical.F90:183-185)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int c = c0 + wc * 16 + nt * 8 + (lane & 3) * 2;
      double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      if (beyond && lev_trunc[c >> 1]) v = make_double2(0.0, 0.0);     // triangular_truncation (spherical.F90:564-600)
      *reinterpret_cast<double2*>(specd + (size_t)(row0 + n) * C + c) = v;
    }
  }
}

void launch_legendre_fwd(const DevTables& t, const double* four, double2* spec, int Lp,
                         const unsigned char* lev_trunc, cudaStream_t st, int ct_begin, int ct_count) {
  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  dim3 grid((g.M + 2 + FWD_NT - 1) / FWD_NT, ct_count < 0 ? C / LEG_CT - ct_begin : ct_count, g.nm);
  if (false) {                     // KC = 32 measured slower (fewer resident CTAs): profiles/r01_experiments.md
    constexpr int KC = 32;
    const size_t smem = sizeof(double) * (2 * 2 * KC * (LEG_CT + 4) + 2 * 2 * (FWD_NT / 2) * (KC + 4));
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(legendre_fwd_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    legendre_fwd_kernel<KC><<<grid, 128, smem, st>>>(t, four, spec, Lp, lev_trunc, ct_begin);
  } else {
    constexpr int KC = 16;
    const size_t smem = sizeof(double) * (2 * 2 * KC * (LEG_CT + 4) + 2 * 2 * (FWD_NT / 2) * (KC + 4));
    legendre_fwd_kernel<KC><<<grid, 128, smem, st>>>(t, four, spec, Lp, lev_trunc, ct_begin);
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversion: reference rectangular (m, n, lev) complex <-> packed [T][Lp]
// ---------------------------------------------------------------------------------------------
__global__ void pack_spec_kernel(DevTables t, const double2* __restrict__ rect, double2* __restrict__ packed,
                                 int nlev, int Lp, int lev0) {
  const GeomDev& g = t.g;
  int p = blockIdx.x;
  int mi = g.row_m[p], n = t.row_n[p], m = g.m_of[mi];
  for (int k = threadIdx.x; k < nlev; k += blockDim.x)
    packed[(size_t)p * Lp + lev0 + k] = rect[((size_t)k * (g.N + 1) + n) * (g.M + 1) + m];
}
__global__ void unpack_spec_kernel(DevTables t, const double2* __restrict__ packed, double2* __restrict__ rect,
                                   int nlev, int Lp, int lev0) {
  const GeomDev& g = t.g;
  int p = blockIdx.x;
  int mi = g.row_m[p], n = t.row_n[p], m = g.m_of[mi];
  for (int k = threadIdx.x; k < nlev; k += blockDim.x)
    rect[((size_t)k * (g.N + 1) + n) * (g.M + 1) + m] = packed[(size_t)p * Lp + lev0 + k];
}
void launch_pack_spec(const DevTables& t, const double2* rect, double2* packed, int nlev, int Lp, int lev0, cudaStream_t st) {
  pack_spec_kernel<<<t.g.T, 64, 0, st>>>(t, rect, packed, nlev, Lp, lev0);
}
void launch_unpack_spec(const DevTables& t, const double2* packed, double2* rect, int nlev, int Lp, int lev0, cudaStream_t st) {
  unpack_spec_kernel<<<t.g.T, 64, 0, st>>>(t, packed, rect, nlev, Lp, lev0);
}

}  // namespace isca
