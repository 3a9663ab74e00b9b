// rrtm_kernels.h -- the CUDA kernels of the RRTMG row: rrtmg_lw_kernel, rrtmg_sw_kernel, the run_rrtmg glue kernels
// (model layout <-> RRTMG layout, interp_temp, units, lonstep) and the zenith-angle kernel.  Included by rrtm.cu (nvcc, the product)
// and by the test-only thread emulator tests/host/rrtm_emu.cpp, which runs the same kernel bodies on CPU threads -- one OS thread
// per CUDA thread of a block, std::barrier for __syncthreads, a per-warp exchange buffer for __shfl_xor_sync -- so that the phase
// structure, the shared-memory staging and the shuffle reductions can be checked without a GPU.  The product never runs it.
//
// Mapping: one CTA per column.  Phase A: one thread per layer runs inatm + setcoef (+ the Planck interpolation) into
// shared memory; thread 0 forms the column sums (precipitable water -> diffusivity angle, laytrop -> solar source
// layers).  Phase B: one thread per g-point (140 LW / 112 SW) evaluates the gaseous optical depths layer by layer with
// the descriptor-driven generic band code of rrtm_column.h and runs the radiative-transfer sweeps (rtrnmr clear-sky /
// reftra + vrtqdr); the radiances of a level are summed over the g-points with warp shuffles + a fixed-order
// cross-warp sum (deterministic).  Phase C: one thread per level writes fluxes and heating rates.
#pragma once
#include "rrtm_column.h"

namespace rrtm_k {
using namespace rrtm;

struct ColIn {                       // device pointers, (ncol, nlay) column-fastest; NULL gas = the constant beside it
  int ncol, nlay;
  const double *play, *plev, *tlay, *tlev, *tsfc, *emis, *albedo, *coszen;
  const double* gas[NSP]; double gas_c[NSP];
  const double* xs[4]; double xs_c[4];
  double *uflx, *dflx, *hr;
  double heatfac, adjflux;
  double* lays;                      // longwave: setcoef planes [LAYP_N][nlay][ncol] written by rrtmg_lw_setcoef_kernel (scratch)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum eight per-lane values over the 32 lanes of a warp with 9 shuffles instead of 8 x 5: each exchange step halves the number of
// values a lane is responsible for (lane bit 4 picks the half v[0..3] / v[4..7], bit 3 the next half, bit 2 the last one), the two last
// steps are a plain butterfly.  Afterwards v[0] of lane L holds the total of value number ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 +
// ((L >> 2) & 1); the order of the additions is fixed, so the result is reproducible.
__device__ __forceinline__ void warp_sum8(double (&v)[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h3 ? v[i] : v[i + 2], keep = h3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double send = h2 ? v[0] : v[1], keep = h2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int warp_sum8_slot(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

struct DevRed {                      // per-level sum over the g-points: warp shuffle, lane 0 stores the warp's partial
  double* part; int ps, warp, lane;  // part[warp][2][ps], ps = nlay + 1
  __device__ __forceinline__ void up(int lev, double v) { v = warp_sum(v); if (lane == 0) part[(warp * 2 + 0) * ps + lev] = v; }
  __device__ __forceinline__ void down(int lev, double v) { v = warp_sum(v); if (lane == 0) part[(warp * 2 + 1) * ps + lev] = v; }
};

constexpr int LW_THREADS = 160, SW_THREADS = 128;
// tile = layers whose term lists are resident at a time (<= 8 = the batch of warp_sum8); MINB = CTAs per SM the register allocation
// is held to.  Tunable at compile time for the occupancy experiments of tools/rrtm_variants.sh; the defaults are the fastest measured
// (profiles/r02/r02_experiments.md: LW tile 4 / 6 CTAs 32.3 ms, SW tile 4 / 8 CTAs 15.3 ms per T170 call).
#ifndef ISCA_LW_TILE
#define ISCA_LW_TILE 4
#endif
#ifndef ISCA_SW_TILE
#define ISCA_SW_TILE 4
#endif
#ifndef ISCA_LW_MINB
#define ISCA_LW_MINB 6
#endif
#ifndef ISCA_SW_MINB
#define ISCA_SW_MINB 8
#endif
constexpr int LW_TILE = ISCA_LW_TILE, SW_TILE = ISCA_SW_TILE;      // layers whose (layer, band) term lists are resident in shared memory at a time
constexpr double FLUXFAC = 3.14159265358979323846 * 2.0e4;      // pi * 2.e4 with pi = 2*asin(1)

// dynamic shared memory (sized by the number of layers, so that more CTAs fit per SM); the CPU thread emulator of
// tests/host/rrtm_emu.cpp defines ISCA_RRTM_EMU and gets a static buffer instead
#ifdef ISCA_RRTM_EMU
#define ISCA_DYN_SMEM(name) alignas(16) static double name[40000]
#else
#define ISCA_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#endif
__host__ __device__ inline size_t lw_smem_doubles(int nl) {
  return (sizeof(Layer) * (size_t)nl + 15) / 16 * 2 + sizeof(LwRec) * (size_t)(LW_TILE * NB_LW) / 8 + (size_t)(LW_THREADS / 32) * 2 * (nl + 1)
       + 3 * NB_LW + 2 * (size_t)(nl + 1) + (size_t)(2 * nl + 1) + (size_t)(2 * nl + 2) / 2 + 8;
}
__host__ __device__ inline size_t sw_smem_doubles(int nl) {
  return (sizeof(Layer) * (size_t)nl + sizeof(SwRec) * (size_t)(SW_TILE * NB_SW)) / 8 + (size_t)(SW_THREADS / 32) * 2 * (nl + 1)
       + 2 * (size_t)(nl + 1) + NB_SW + 8;
}

// Longwave.  Phase A: one thread per layer, inatm + setcoef + Planck functions.  Phase B, tile by tile (LW_TILE layers, top-down):
// (i) one thread per (layer, band) builds the term list of that layer and band (everything of taugbNN that does not depend on the
// g-point: species ratios, interpolation weights, `pow` of the adjusted minor-gas amounts); (ii) one thread per g-point forms
// tau = sum_i w_i A[off_i + g] and advances rtrnmr's downward sweep.  The upward sweep needs no optical depths again
// (atrans / bbugas of the column stay in thread-local memory).  Phase C: fluxes and heating rates.
__global__ void __launch_bounds__(LW_THREADS, ISCA_LW_MINB) rrtmg_lw_kernel(const double* __restrict__ A, Tab tb, const LwBand* __restrict__ bands, ColIn in) {
  ISCA_DYN_SMEM(smem);
  const int col = blockIdx.x, tid = threadIdx.x, nl = in.nlay, nc = in.ncol;
  const int PS = nl + 1;             // row stride of the Planck arrays
  Layer* lay = reinterpret_cast<Layer*>(smem);
  LwRec* recs = reinterpret_cast<LwRec*>(smem + (sizeof(Layer) * (size_t)nl + 15) / 16 * 2);
  double* part = reinterpret_cast<double*>(recs + LW_TILE * NB_LW);      // [warp][2][PS]
  double* plankbnd = part + (LW_THREADS / 32) * 2 * PS;
  double* secdiff = plankbnd + NB_LW;
  double* semiss = secdiff + NB_LW;
  double* pz = semiss + NB_LW;
  double* fnet = pz + PS;
  // the Planck functions of the layers and levels are interpolated by the g-point threads themselves (two table reads each, L1-resident)
  // from the index / fraction pairs below instead of being staged for all 16 bands: 10 KB less shared memory = one more CTA per SM
  double* pl_frac = fnet + PS;                           // [nl] layers, then [nl + 1] levels
  int* pl_ind = reinterpret_cast<int*>(pl_frac + 2 * nl + 1);
  // ---- phase A: inatm + setcoef per layer
  for (int l = tid; l < nl; l += LW_THREADS) {
    double vmr[NSP], xs[4];
    for (int i = 0; i < NSP; ++i) vmr[i] = in.gas[i] ? in.gas[i][col + (size_t)nc * l] : in.gas_c[i];
    for (int i = 0; i < 4; ++i) xs[i] = in.xs[i] ? in.xs[i][col + (size_t)nc * l] : in.xs_c[i];
    double pb = in.plev[col + (size_t)nc * l], pa = in.plev[col + (size_t)nc * (l + 1)];
    double coldry = coldry_of(pb, pa, vmr[0]);
    double tav = in.tlay[col + (size_t)nc * l];
    lw_setcoef_layer(A, tb, in.play[col + (size_t)nc * l], tav, coldry, vmr, xs, lay[l]);
    planck_index(tav, pl_ind[l], pl_frac[l]);
    planck_index(in.tlev[col + (size_t)nc * (l + 1)], pl_ind[nl + l + 1], pl_frac[nl + l + 1]);
    pz[l + 1] = pa;
    if (l == 0) pz[0] = pb;
  }
  if (tid < NB_LW) semiss[tid] = in.emis ? in.emis[col + (size_t)nc * tid] : 1.0;
  if (tid == LW_THREADS - 1) planck_index(in.tlev[col], pl_ind[nl], pl_frac[nl]);
  __syncthreads();
  if (tid == 0) {
    double pb[NB_LW];
    planck16(A, tb, in.tsfc[col], pb, 1);
    for (int ib = 0; ib < NB_LW; ++ib) plankbnd[ib] = semiss[ib] * pb[ib];
    // inatm: precipitable water (rrtmg_lw_rad.nomcica.f90:846-856), sequential over the layers
    double amttl = 0.0, wvttl = 0.0;
    for (int l = 0; l < nl; ++l) {
      double wv = lay[l].col[SP_H2O] * 1.0e20;
      amttl += lay[l].coldry + wv;
      wvttl += wv;
    }
    double wvsh = (AMW * wvttl) / (AMD * amttl);
    double pwvcm = wvsh * (1.0e3 * pz[0]) / (1.0e2 * GRAV);
    for (int ib = 0; ib < NB_LW; ++ib) secdiff[ib] = lw_secdiff(ib, pwvcm);
  }
  // ---- phase B: one g-point per thread, rtrnmr clear sky (rrtmg_lw_rtrnmr.f90:390-480 down, :560-640 up)
  const int valid = tid < NG_LW;
  const int g = valid ? tid : NG_LW - 1;
  int ib = 0;
  while (ib < NB_LW - 1 && g >= bands[ib + 1].g0) ++ib;
  const int gb = g - bands[ib].g0;
  const double delwave[NB_LW] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
  const double w = valid ? 0.5 * delwave[ib] : 0.0;
  const int warp = tid >> 5, lane = tid & 31, slot = warp_sum8_slot(lane);
  DevRed red{part, PS, warp, lane};
  const double* exptfn = A + tb.exptfn;
  double atrans[KMAX], bbugas[KMAX];
  double radld = 0.0, frac1 = 0.0;
  red.down(nl, 0.0);
  for (int hi = nl; hi >= 1; hi -= LW_TILE) {          // layers lev = hi .. lo (1-based), top-down
    const int lo = hi - LW_TILE + 1 > 1 ? hi - LW_TILE + 1 : 1, cnt = hi - lo + 1;
    __syncthreads();                                    // phase A (first tile) / the previous tile's term lists are consumed
    for (int task = tid; task < cnt * NB_LW; task += LW_THREADS) {     // band-major: the lanes of a warp share 3-4 bands (code paths), not 16
      const int b = task / cnt, li = task - b * cnt;
      lw_terms(A, tb, bands[b], lay[lo - 1 + li], recs[li * NB_LW + b]);
    }
    __syncthreads();
    const double sd = secdiff[ib];
    double v8[8];
    double pl_up = planck_at(A, tb, ib, pl_ind[nl + hi], pl_frac[nl + hi]);       // Planck function of level hi (top of the tile)
#pragma unroll
    for (int j = 0; j < LW_TILE; ++j) {
      const int lev = hi - j;
      v8[j] = 0.0;
      if (lev >= lo) {
        double tau, plfrac;
        lw_tau_rec(A, recs[(lev - lo) * NB_LW + ib], gb, tau, plfrac);
        if (lev == 1) frac1 = plfrac;
        const double pl_dn = planck_at(A, tb, ib, pl_ind[nl + lev - 1], pl_frac[nl + lev - 1]);
        lw_layer(exptfn, sd, tau, plfrac, planck_at(A, tb, ib, pl_ind[lev - 1], pl_frac[lev - 1]), pl_up, pl_dn,
                 radld, atrans[lev - 1], bbugas[lev - 1]);
        pl_up = pl_dn;
        v8[j] = radld * w;
      }
    }
    warp_sum8(v8, lane);                                 // the tile's eight downward radiances at once
    if ((lane & 3) == 0 && slot < cnt) part[(warp * 2 + 1) * PS + hi - 1 - slot] = v8[0];
  }
  {
    double rad0 = frac1 * plankbnd[ib];
    double radlu = rad0 + (1.0 - semiss[ib]) * radld;
    red.up(0, radlu * w);
    for (int l0 = 1; l0 <= nl; l0 += 8) {
      double v8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int lev = l0 + j;
        v8[j] = 0.0;
        if (lev <= nl) {
          radlu = radlu + (bbugas[lev - 1] - radlu) * atrans[lev - 1];
          v8[j] = radlu * w;
        }
      }
      warp_sum8(v8, lane);
      if ((lane & 3) == 0 && l0 + slot <= nl) part[(warp * 2 + 0) * PS + l0 + slot] = v8[0];
    }
  }
  __syncthreads();
  // ---- phase C: fluxes and heating rates
  for (int lev = tid; lev <= nl; lev += LW_THREADS) {
    double u = 0.0, d = 0.0;
    for (int wp = 0; wp < LW_THREADS / 32; ++wp) { u += part[(wp * 2 + 0) * PS + lev]; d += part[(wp * 2 + 1) * PS + lev]; }
    u *= FLUXFAC; d *= FLUXFAC;
    in.uflx[col + (size_t)nc * lev] = u;
    in.dflx[col + (size_t)nc * lev] = d;
    fnet[lev] = u - d;
  }
  __syncthreads();
  for (int l = tid; l < nl; l += LW_THREADS)
    in.hr[col + (size_t)nc * l] = in.heatfac * (fnet[l] - fnet[l + 1]) / (pz[l] - pz[l + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// Longwave, column-per-lane mapping (round 2, the production kernel).
//
// The g-point-per-thread kernel above is bound by the L1TEX pipe (ncu, T170: 85 % of its peak; 11.6 G warp instructions per call): each
// warp instruction serves the 32 g-points of ONE column, the term lists have to travel through shared memory, and their construction
// diverges 16 ways.  Here a LANE is a COLUMN (32 consecutive longitudes of one latitude) and a WARP owns a group of bands:
//   * the term list of a (layer, band) is built by every lane for its own column with the same code path in all lanes, kept in
//     thread-local memory, and read once per term -- not once per (term, g-point);
//   * the table rows of neighbouring columns coincide (same pressure / temperature bins), so a table read is a broadcast (one L1
//     wavefront for 32 columns) instead of 256 bytes per column;
//   * the Planck functions of a layer are interpolated once per (layer, band), not per g-point;
//   * the sum over the g-points is a serial sum in registers: no shuffles, no barriers inside the sweeps.
// A CTA = 32 columns x 4 warps; the four band groups are balanced by sum(ng x terms).  The per-layer transmittance / source pairs of
// the upward sweep stay in thread-local memory as before.  setcoef is evaluated once per (column, layer) by a pre-pass kernel (below).
// ---------------------------------------------------------------------------------------------------------------
constexpr int LWC_WARPS = 4;
// longest-processing-time assignment of cost(band) = 550 (setcoef + term list per layer) + ng * (1.6 terms + 60):
// warp 0: bands 5, 8, 6, 14; warp 1: 3, 1, 10, 13; warp 2: 4, 2, 12, 15; warp 3: 7, 9, 11, 16 (1-based band numbers), ~5200 each
__device__ __constant__ int LWC_GROUP_OF_BAND[NB_LW] = {1, 2, 1, 2, 0, 0, 3, 0, 3, 1, 3, 2, 1, 0, 2, 3};   // 0-based band -> warp

#ifndef ISCA_LWC_PREPASS
#define ISCA_LWC_PREPASS 1
#endif
// 0: scalar table reads, unrolling left to the compiler (fastest measured: 24.7 ms per T170 L40 call); 2 / 4: that many terms in flight
// with explicit 128-bit reads (tab2; measured slower with the evict-last hint, which pushes the thread-local arrays out of L1)
#ifndef ISCA_LWC_UNROLL
#define ISCA_LWC_UNROLL 0
#endif
// setcoef does not depend on the band: rrtmg_lw_setcoef_kernel evaluates inatm + setcoef once per (column, layer) into planes
// [LAYP_N][layer][column] (1.1 GB at T170 L40), and the band sweeps read back only the planes their band descriptor uses.
// Plane 0 holds the seven small integers packed into one exactly representable double.
enum { LAYP_PK = 0, LAYP_FAC00, LAYP_FAC01, LAYP_FAC10, LAYP_FAC11, LAYP_SELFFAC, LAYP_SELFFRAC, LAYP_FORFAC, LAYP_FORFRAC,
       LAYP_MINORFRAC, LAYP_SCALEMINOR, LAYP_SCALEMINORN2, LAYP_COLBRD, LAYP_COLDRY, LAYP_PAVEL, LAYP_COL, LAYP_WX = LAYP_COL + NSP,
       LAYP_N = LAYP_WX + 4 };

__device__ __forceinline__ void lw_layer_inputs(const ColIn& in, int col, int l, const double* __restrict__ A, const Tab& tb, Layer& L) {
  const size_t nc = in.ncol;
  double vmr[NSP], xs[4];
#pragma unroll
  for (int i = 0; i < NSP; ++i) vmr[i] = in.gas[i] ? in.gas[i][col + nc * l] : in.gas_c[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) xs[i] = in.xs[i] ? in.xs[i][col + nc * l] : in.xs_c[i];
  const double pb = in.plev[col + nc * l], pa = in.plev[col + nc * (l + 1)];
  lw_setcoef_layer(A, tb, in.play[col + nc * l], in.tlay[col + nc * l], coldry_of(pb, pa, vmr[0]), vmr, xs, L);
}

__global__ void __launch_bounds__(128) rrtmg_lw_setcoef_kernel(const double* __restrict__ A, Tab tb, ColIn in) {
  const size_t nc = in.ncol, ps = nc * (size_t)in.nlay;
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= ps) return;
  const int l = (int)(i / nc), col = (int)(i - (size_t)l * nc);
  Layer L;
  lw_layer_inputs(in, col, l, A, tb, L);
  double* o = in.lays + i;
  const int pk = L.jp | (L.jt << 6) | (L.jt1 << 9) | (L.indself << 12) | (L.indfor << 16) | (L.indminor << 18) | (L.lower << 23);
  o[LAYP_PK * ps] = (double)pk;
  o[LAYP_FAC00 * ps] = L.fac00; o[LAYP_FAC01 * ps] = L.fac01; o[LAYP_FAC10 * ps] = L.fac10; o[LAYP_FAC11 * ps] = L.fac11;
  o[LAYP_SELFFAC * ps] = L.selffac; o[LAYP_SELFFRAC * ps] = L.selffrac; o[LAYP_FORFAC * ps] = L.forfac; o[LAYP_FORFRAC * ps] = L.forfrac;
  o[LAYP_MINORFRAC * ps] = L.minorfrac; o[LAYP_SCALEMINOR * ps] = L.scaleminor; o[LAYP_SCALEMINORN2 * ps] = L.scaleminorn2;
  o[LAYP_COLBRD * ps] = L.colbrd; o[LAYP_COLDRY * ps] = L.coldry; o[LAYP_PAVEL * ps] = L.pavel;
#pragma unroll
  for (int k = 0; k < NSP; ++k) o[(LAYP_COL + k) * ps] = L.col[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[(LAYP_WX + k) * ps] = L.wx[k];
}

// The term list of one (column, layer, band) from the setcoef planes: reads exactly the planes lw_terms uses for this band / region.
// One copy of this code serves every band (not inlined into the four g-point-count instantiations of the sweep: instruction cache).
__device__ __noinline__ void lw_record_from_planes(const double* __restrict__ A, const Tab& tb, const LwBand& B,
                                                   const double* __restrict__ p, size_t ps, LwRec& rec) {
  Layer L;
  const int pk = (int)p[LAYP_PK * ps];
  L.jp = pk & 63; L.jt = (pk >> 6) & 7; L.jt1 = (pk >> 9) & 7; L.indself = (pk >> 12) & 15; L.indfor = (pk >> 16) & 3;
  L.indminor = (pk >> 18) & 31; L.lower = (pk >> 23) & 1;
  const LwRegion& R = B.r[L.lower ? 0 : 1];
  L.fac00 = p[LAYP_FAC00 * ps]; L.fac01 = p[LAYP_FAC01 * ps]; L.fac10 = p[LAYP_FAC10 * ps]; L.fac11 = p[LAYP_FAC11 * ps];
  if (R.major >= 1) L.col[R.spA] = p[(LAYP_COL + R.spA) * ps];
  if (R.spB >= 0) { L.col[R.spA] = p[(LAYP_COL + R.spA) * ps]; L.col[R.spB] = p[(LAYP_COL + R.spB) * ps]; }
  if (R.self_off >= 0) { L.selffac = p[LAYP_SELFFAC * ps]; L.selffrac = p[LAYP_SELFFRAC * ps]; }
  if (R.for_off >= 0) { L.forfac = p[LAYP_FORFAC * ps]; L.forfrac = p[LAYP_FORFRAC * ps]; }
  if (R.nminor > 0) {
    L.minorfrac = p[LAYP_MINORFRAC * ps];
    for (int m = 0; m < R.nminor; ++m) {
      const Minor& M = R.minor[m];
      if (M.scale == SC_COL) L.col[M.sp] = p[(LAYP_COL + M.sp) * ps];
      else if (M.scale == SC_ADJ) { L.col[M.sp] = p[(LAYP_COL + M.sp) * ps]; L.coldry = p[LAYP_COLDRY * ps]; }
      else if (M.scale == SC_BRD_N2) { L.colbrd = p[LAYP_COLBRD * ps]; L.scaleminorn2 = p[LAYP_SCALEMINORN2 * ps]; }
      else if (M.scale == SC_BRD) { L.colbrd = p[LAYP_COLBRD * ps]; L.scaleminor = p[LAYP_SCALEMINOR * ps]; }
      else { L.col[SP_O2] = p[(LAYP_COL + SP_O2) * ps]; L.scaleminor = p[LAYP_SCALEMINOR * ps]; }
    }
  }
  for (int c = 0; c < R.ncfc; ++c) L.wx[R.cfc_wx[c]] = p[(LAYP_WX + R.cfc_wx[c]) * ps];
  if (R.corr != 0) L.pavel = p[LAYP_PAVEL * ps];
  lw_terms(A, tb, B, L, rec);
}

// one band of one column: rtrnmr's clear-sky down and up sweeps.  The downward radiances of the band's NG (padded) g-points stay in
// registers; optical depths are formed and consumed in groups of H = min(NG, 8) g-points to bound the live registers.
template <int NG>
__device__ __forceinline__ void lw_band_column(const double* __restrict__ A, const Tab& tb, const LwBand& B, int ib, const ColIn& in, int col,
                                               int nl, double pwvcm, double* __restrict__ fu, double* __restrict__ fd,
                                               double* __restrict__ AT, double* __restrict__ BBU) {
  constexpr int H = NG < 8 ? NG : 8;
  const size_t nc = in.ncol;
  const int ng = B.ng;
  const double* exptfn = A + tb.exptfn;
  const double delwave[NB_LW] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
  const double wb = 0.5 * delwave[ib];
  const double sd = lw_secdiff(ib, pwvcm);
  const double semiss = in.emis ? in.emis[col + nc * ib] : 1.0;
  double radld[NG];
  double* FR1 = BBU + (size_t)nl * NG;                          // Planck fractions of layer 1 (needed again at the surface)
#pragma unroll
  for (int g = 0; g < NG; ++g) radld[g] = 0.0;
  int ind; double fr;
  planck_index(in.tlev[col + nc * nl], ind, fr);
  double pl_up = planck_at(A, tb, ib, ind, fr);                // Planck function of the level above the current layer
  for (int lev = nl; lev >= 1; --lev) {
    const int l = lev - 1;
    LwRec rec;
#if ISCA_LWC_PREPASS
    lw_record_from_planes(A, tb, B, in.lays + (col + nc * l), nc * (size_t)nl, rec);
#else
    {
      Layer L;                                                   // inatm + setcoef of the layer
      lw_layer_inputs(in, col, l, A, tb, L);
      lw_terms(A, tb, B, L, rec);
    }
#endif
    planck_index(in.tlay[col + nc * l], ind, fr);
    const double blay = planck_at(A, tb, ib, ind, fr);
    planck_index(in.tlev[col + nc * l], ind, fr);
    const double pl_dn = planck_at(A, tb, ib, ind, fr);
    double sum = 0.0;
#pragma unroll
    for (int h0 = 0; h0 < NG; h0 += H) {
      double tau[H];
#pragma unroll
      for (int g = 0; g < H; ++g) tau[g] = 0.0;
#if ISCA_LWC_UNROLL == 0
      for (int i = 0; i < rec.n; ++i) {                          // one (weight, row) pair per term, H table reads each
        const double w = rec.w[i];
        const double* row = A + rec.off[i] + h0;
#pragma unroll
        for (int g = 0; g < H; ++g) if (h0 + g < ng) tau[g] += w * row[g];
      }
#pragma unroll
      for (int g = 0; g < H; ++g) {
        const int gg = h0 + g;
        if (gg < ng) {
          double t = tau[g];
          if (rec.gs_off >= 0) t *= A[rec.gs_off + gg];
          double plfrac = 0.0;
          if (rec.f0 >= 0) { const double a = A[rec.f0 + gg]; plfrac = rec.f1 >= 0 ? a + rec.ffs * (A[rec.f1 + gg] - a) : a; }
          if (lev == 1) FR1[gg] = plfrac;
          double at, bbu;
          lw_layer(exptfn, sd, t, plfrac, blay, pl_up, pl_dn, radld[gg], at, bbu);
          AT[l * NG + gg] = at; BBU[l * NG + gg] = bbu;
          sum += radld[gg];
        }
      }
    }
#else
      // one (weight, row) pair per term; the lists are padded to multiples of 8 terms, ISCA_LWC_UNROLL terms (x H / 2 128-bit table
      // reads) are in flight at a time; the additions stay in list order
      for (int i = 0; i < rec.n; i += ISCA_LWC_UNROLL) {
        double w[ISCA_LWC_UNROLL], v[ISCA_LWC_UNROLL][H];
#pragma unroll
        for (int u = 0; u < ISCA_LWC_UNROLL; ++u) {
          w[u] = rec.w[i + u];
          const double* row = A + rec.off[i + u] + h0;
#pragma unroll
          for (int g = 0; g < H; g += 2) if (h0 + g < ng) tab2(row + g, v[u][g], v[u][g + 1]);
        }
#pragma unroll
        for (int u = 0; u < ISCA_LWC_UNROLL; ++u)
#pragma unroll
          for (int g = 0; g < H; ++g) if (h0 + g < ng) tau[g] += w[u] * v[u][g];
      }
      double gsc[H], fa[H], fb[H];                               // per-g-point scaling and Planck fractions of the band's table rows
#pragma unroll
      for (int g = 0; g < H; g += 2)
        if (h0 + g < ng) {
          if (rec.gs_off >= 0) tab2(A + rec.gs_off + h0 + g, gsc[g], gsc[g + 1]);
          if (rec.f0 >= 0) tab2(A + rec.f0 + h0 + g, fa[g], fa[g + 1]);
          if (rec.f1 >= 0) tab2(A + rec.f1 + h0 + g, fb[g], fb[g + 1]);
        }
#pragma unroll
      for (int g = 0; g < H; ++g) {
        const int gg = h0 + g;
        if (gg < ng) {
          double t = tau[g];
          if (rec.gs_off >= 0) t *= gsc[g];
          double plfrac = 0.0;
          if (rec.f0 >= 0) plfrac = rec.f1 >= 0 ? fa[g] + rec.ffs * (fb[g] - fa[g]) : fa[g];
          if (lev == 1) FR1[gg] = plfrac;
          double at, bbu;
          lw_layer(exptfn, sd, t, plfrac, blay, pl_up, pl_dn, radld[gg], at, bbu);
          AT[l * NG + gg] = at; BBU[l * NG + gg] = bbu;
          sum += radld[gg];
        }
      }
    }
#endif
    fd[l] += sum * wb;
    pl_up = pl_dn;
  }
  planck_index(in.tsfc[col], ind, fr);
  const double plankbnd = semiss * planck_at(A, tb, ib, ind, fr);
  double radlu[NG];
  {
    double sum = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      radlu[g] = 0.0;
      if (g < ng) { radlu[g] = FR1[g] * plankbnd + (1.0 - semiss) * radld[g]; sum += radlu[g]; }
    }
    fu[0] += sum * wb;
  }
  for (int lev = 1; lev <= nl; ++lev) {
    const int l = lev - 1;
    double sum = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g)
      if (g < ng) { radlu[g] = radlu[g] + (BBU[l * NG + g] - radlu[g]) * AT[l * NG + g]; sum += radlu[g]; }
    fu[lev] += sum * wb;
  }
}

#ifndef ISCA_LWC_MINB
#define ISCA_LWC_MINB 5
#endif
__global__ void __launch_bounds__(32 * LWC_WARPS, ISCA_LWC_MINB) rrtmg_lw_col_kernel(const double* __restrict__ A, Tab tb, const LwBand* __restrict__ bands, ColIn in) {
  __shared__ double sfu[(KMAX + 1) * 32], sfd[(KMAX + 1) * 32];        // [level][column of the CTA]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nl = in.nlay;
  const size_t nc = in.ncol;
  const int col_raw = blockIdx.x * 32 + lane;
  const bool live = col_raw < in.ncol;
  const int col = live ? col_raw : in.ncol - 1;                   // padding lanes repeat the last column (no divergence, no stores)
  double fu[KMAX + 1], fd[KMAX + 1];
  double AT[KMAX * 16], BBU[(KMAX + 1) * 16];
  for (int lev = 0; lev <= nl; ++lev) { fu[lev] = 0.0; fd[lev] = 0.0; }
  // inatm: precipitable water of the column (rrtmg_lw_rad.nomcica.f90:846-856) -> diffusivity angle of the bands
  double pwvcm;
  {
    double amttl = 0.0, wvttl = 0.0;
    for (int l = 0; l < nl; ++l) {
      const double h2o = in.gas[SP_H2O] ? in.gas[SP_H2O][col + nc * l] : in.gas_c[SP_H2O];
      const double coldry = coldry_of(in.plev[col + nc * l], in.plev[col + nc * (l + 1)], h2o);
      const double wv = (1.0e-20 * (coldry * h2o)) * 1.0e20;   // colh2o * 1.e20 as setcoef / inatm form it
      amttl += coldry + wv;
      wvttl += wv;
    }
    const double wvsh = (AMW * wvttl) / (AMD * amttl);
    pwvcm = wvsh * (1.0e3 * in.plev[col]) / (1.0e2 * GRAV);
  }
  for (int ib = 0; ib < NB_LW; ++ib) {
    if (LWC_GROUP_OF_BAND[ib] != warp) continue;
    const LwBand& B = bands[ib];
    if (B.ng > 8) lw_band_column<16>(A, tb, B, ib, in, col, nl, pwvcm, fu, fd, AT, BBU);
    else if (B.ng > 4) lw_band_column<8>(A, tb, B, ib, in, col, nl, pwvcm, fu, fd, AT, BBU);
    else if (B.ng > 2) lw_band_column<4>(A, tb, B, ib, in, col, nl, pwvcm, fu, fd, AT, BBU);
    else lw_band_column<2>(A, tb, B, ib, in, col, nl, pwvcm, fu, fd, AT, BBU);
  }
  // sum of the four band groups in a fixed order (warp 0, 1, 2, 3): reproducible
  for (int w = 0; w < LWC_WARPS; ++w) {
    if (warp == w)
      for (int lev = 0; lev <= nl; ++lev) {
        if (w == 0) { sfu[lev * 32 + lane] = fu[lev]; sfd[lev * 32 + lane] = fd[lev]; }
        else { sfu[lev * 32 + lane] += fu[lev]; sfd[lev * 32 + lane] += fd[lev]; }
      }
    __syncthreads();
  }
  if (!live) return;
  for (int lev = warp; lev <= nl; lev += LWC_WARPS) {
    in.uflx[col + nc * lev] = sfu[lev * 32 + lane] * FLUXFAC;
    in.dflx[col + nc * lev] = sfd[lev * 32 + lane] * FLUXFAC;
  }
  for (int l = warp; l < nl; l += LWC_WARPS) {
    const double f0 = sfu[l * 32 + lane] * FLUXFAC - sfd[l * 32 + lane] * FLUXFAC;
    const double f1 = sfu[(l + 1) * 32 + lane] * FLUXFAC - sfd[(l + 1) * 32 + lane] * FLUXFAC;
    in.hr[col + nc * l] = in.heatfac * (f0 - f1) / (in.plev[col + nc * l] - in.plev[col + nc * (l + 1)]);
  }
}

// Shortwave (spcvrt_sw clear sky + reftra_sw + vrtqdr_sw).  Same structure as the longwave kernel; the sweeps are re-ordered so that
// the optical depths are needed only once: pass 1 runs bottom-up (layer optics, two-stream reflectance / transmittance of the layer
// and, in the same step, vrtqdr's upward combination zrup / zrupd, which is a bottom-up recurrence), pass 2 top-down (direct-beam
// transmittance, downward combination, fluxes).  Values and their order of evaluation are those of the reference.
__global__ void __launch_bounds__(SW_THREADS, ISCA_SW_MINB) rrtmg_sw_kernel(const double* __restrict__ A, Tab tb, const SwBand* __restrict__ bands, ColIn in) {
  ISCA_DYN_SMEM(smem);
  const int col = blockIdx.x, tid = threadIdx.x, nl = in.nlay, nc = in.ncol;
  const int PS = nl + 1;
  const double cosz = in.coszen[col];
  if (cosz < 1.0e-10) {              // `if (coszen(iplon) < zepzen) ... cycle` (rrtmg_sw_rad.nomcica.f90)
    for (int lev = tid; lev <= nl; lev += SW_THREADS) { in.uflx[col + (size_t)nc * lev] = 0.0; in.dflx[col + (size_t)nc * lev] = 0.0; }
    for (int l = tid; l < nl; l += SW_THREADS) in.hr[col + (size_t)nc * l] = 0.0;
    return;
  }
  Layer* lay = reinterpret_cast<Layer*>(smem);
  SwRec* recs = reinterpret_cast<SwRec*>(smem + (sizeof(Layer) * (size_t)nl + 7) / 8);
  double* part = reinterpret_cast<double*>(recs + SW_TILE * NB_SW);      // [warp][2][PS]
  double* pz = part + (SW_THREADS / 32) * 2 * PS;
  double* fnet = pz + PS;
  int* lsol = reinterpret_cast<int*>(fnet + PS);                           // [NB_SW] + laytrop
  int* laytrop_s = lsol + NB_SW;
  for (int l = tid; l < nl; l += SW_THREADS) {
    double vmr[NSP];
    for (int i = 0; i < NSP; ++i) vmr[i] = in.gas[i] ? in.gas[i][col + (size_t)nc * l] : in.gas_c[i];
    double pb = in.plev[col + (size_t)nc * l], pa = in.plev[col + (size_t)nc * (l + 1)];
    sw_setcoef_layer(A, tb, in.play[col + (size_t)nc * l], in.tlay[col + (size_t)nc * l], coldry_of(pb, pa, vmr[0]), vmr, lay[l]);
    pz[l + 1] = pa;
    if (l == 0) pz[0] = pb;
  }
  __syncthreads();
  if (tid == 0) { int n = 0; for (int l = 0; l < nl; ++l) n += lay[l].lower; *laytrop_s = n; }
  __syncthreads();
  if (tid < NB_SW) lsol[tid] = sw_laysolfr(bands[tid], lay, nl, *laytrop_s);
  const int valid = tid < NG_SW;
  const int g = valid ? tid : NG_SW - 1;
  int ib = 0;
  while (ib < NB_SW - 1 && g >= bands[ib + 1].g0) ++ib;
  const int gb = g - bands[ib].g0;
  const double* exp_tbl = A + tb.exp_tbl;
  const double albedo = in.albedo[col], prmu0 = cosz;
  const int warp = tid >> 5, lane = tid & 31, slot = warp_sum8_slot(lane);
  // per-thread column arrays, index jk = 0 (top layer) .. nl-1 (bottom layer), levels 0 (top) .. nl (surface)
  double zref[KMAX], zrefd[KMAX], ztra[KMAX], ztrad[KMAX], zdbt[KMAX], zrup[KMAX + 1], zrupd[KMAX + 1];
  double sflux = 0.0;
  zrup[nl] = albedo; zrupd[nl] = albedo;
  for (int lo = 1; lo <= nl; lo += SW_TILE) {          // layers (1-based, bottom-up) lo .. hi
    const int hi = lo + SW_TILE - 1 < nl ? lo + SW_TILE - 1 : nl, cnt = hi - lo + 1;
    __syncthreads();
    for (int task = tid; task < cnt * NB_SW; task += SW_THREADS) {     // band-major (see the longwave kernel)
      const int b = task / cnt, li = task - b * cnt;
      sw_terms(A, bands[b], lay[lo - 1 + li], recs[li * NB_SW + b]);
    }
    __syncthreads();
    const int ls = lsol[ib];
    for (int l1 = lo; l1 <= hi; ++l1) {
      const SwRec& rc = recs[(l1 - lo) * NB_SW + ib];
      const int jk = nl - l1;                            // ikl = l1 - 1 = nl - 1 - jk
      double taug, taur;
      sw_tau_rec(A, rc, gb, taug, taur);
      if (l1 == ls) sflux = sw_src_rec(A, bands[ib], rc, gb);
      sw_layer(exp_tbl, prmu0, taug, taur, zrup[jk + 1], zrupd[jk + 1], zref[jk], zrefd[jk], ztra[jk], ztrad[jk], zdbt[jk], zrup[jk], zrupd[jk]);
    }
  }
  {
    // top-down transmittances and the fluxes at every level
    const double zinc = in.adjflux * sflux * prmu0 * (valid ? 1.0 : 0.0);
    double ztdn = 1.0, zrdnd = 0.0, tdbt = 1.0, tdbt_prev = 1.0;     // tdbt = ztdbt(jk), tdbt_prev = ztdbt(jk-1)
    for (int j0 = 0; j0 <= nl; j0 += 4) {                // four levels = eight values (up, down) per warp_sum8
      double v8[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jk = j0 + j;
        v8[2 * j] = 0.0; v8[2 * j + 1] = 0.0;
        if (jk <= nl) {
          if (jk >= 1) { tdbt_prev = tdbt; tdbt = zdbt[jk - 1] * tdbt_prev; }
          const int jl = jk >= 1 ? jk - 1 : 0;
          double fu, fd;
          sw_level(jk, tdbt, tdbt_prev, zref[jl], zrefd[jl], ztra[jl], ztrad[jl], zrup[jk], zrupd[jk], ztdn, zrdnd, fu, fd);
          v8[2 * j] = zinc * fu; v8[2 * j + 1] = zinc * fd;
        }
      }
      warp_sum8(v8, lane);
      const int jk = j0 + (slot >> 1);
      if ((lane & 3) == 0 && jk <= nl) part[(warp * 2 + (slot & 1)) * PS + nl - jk] = v8[0];
    }
  }
  __syncthreads();
  for (int lev = tid; lev <= nl; lev += SW_THREADS) {
    double u = 0.0, d = 0.0;
    for (int wp = 0; wp < SW_THREADS / 32; ++wp) { u += part[(wp * 2 + 0) * PS + lev]; d += part[(wp * 2 + 1) * PS + lev]; }
    in.uflx[col + (size_t)nc * lev] = u;
    in.dflx[col + (size_t)nc * lev] = d;
    fnet[lev] = d - u;
  }
  __syncthreads();
  for (int l = tid; l < nl; l += SW_THREADS)    // swhr(nlayers) = 0 in the reference
    in.hr[col + (size_t)nc * l] = l == nl - 1 ? 0.0 : (fnet[l + 1] - fnet[l]) * in.heatfac / (pz[l] - pz[l + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// Shortwave, column-per-lane mapping (see the longwave kernel above): a lane = a column, a warp = a band group; bands are processed in
// chunks of at most 8 g-points whose recurrences run in registers.  Pass 1 bottom-up (layer optics, reftra, upward combination), pass 2
// top-down (direct beam, downward combination, fluxes); the seven per-(layer, g-point) values pass 2 needs stay in thread-local memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SWC_WARPS = 4, SWC_NG = 8;
__device__ __constant__ int SWC_GROUP_OF_BAND[NB_SW] = {0, 0, 2, 2, 0, 1, 2, 2, 3, 1, 3, 3, 3, 1};        // 0-based band -> warp

__device__ __forceinline__ void sw_band_chunk(const double* __restrict__ A, const Tab& tb, const SwBand& B, int g0, int ngc, const ColIn& in,
                                              int col, int nl, int lsol, double prmu0, double albedo, double* __restrict__ fu,
                                              double* __restrict__ fd, double* __restrict__ W) {
  constexpr int NG = SWC_NG;
  const size_t nc = in.ncol;
  const double* exp_tbl = A + tb.exp_tbl;
  // W: seven arrays [level or layer][NG]
  double* zref = W; double* zrefd = zref + KMAX * NG; double* ztra = zrefd + KMAX * NG; double* ztrad = ztra + KMAX * NG;
  double* zdbt = ztrad + KMAX * NG; double* zrup = zdbt + KMAX * NG; double* zrupd = zrup + (KMAX + 1) * NG;
  double rup[NG], rupd[NG], sflux[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) { rup[g] = albedo; rupd[g] = albedo; sflux[g] = 0.0; zrup[nl * NG + g] = albedo; zrupd[nl * NG + g] = albedo; }
  for (int l1 = 1; l1 <= nl; ++l1) {                   // bottom-up: layer optics + upward combination
    const int l = l1 - 1, jk = nl - l1;
    Layer L;
    {
      double vmr[NSP];
#pragma unroll
      for (int i = 0; i < NSP; ++i) vmr[i] = in.gas[i] ? in.gas[i][col + nc * l] : in.gas_c[i];
      const double pb = in.plev[col + nc * l], pa = in.plev[col + nc * (l + 1)];
      sw_setcoef_layer(A, tb, in.play[col + nc * l], in.tlay[col + nc * l], coldry_of(pb, pa, vmr[0]), vmr, L);
    }
    SwRec rec;
    sw_terms(A, B, L, rec);
    double taug[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) taug[g] = 0.0;
    for (int i = 0; i < rec.n; ++i) {
      const double w = rec.w[i];
      const double* row = A + rec.off[i] + g0;
#pragma unroll
      for (int g = 0; g < NG; ++g) if (g < ngc) taug[g] += w * row[g];
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g < ngc) {
        const int gg = g0 + g;
        const double tg = taug[g] + rec.tconst;
        const double a = A[rec.r0 + rec.rg * gg];
        const double tr = rec.rc0 * (rec.r1 >= 0 ? a + rec.rc1 * (A[rec.r1 + gg] - a) : a);
        if (l1 == lsol) sflux[g] = sw_src_rec(A, B, rec, gg);
        double rf, rfd, tt, ttd, db, ru, rud;
        sw_layer(exp_tbl, prmu0, tg, tr, rup[g], rupd[g], rf, rfd, tt, ttd, db, ru, rud);
        zref[jk * NG + g] = rf; zrefd[jk * NG + g] = rfd; ztra[jk * NG + g] = tt; ztrad[jk * NG + g] = ttd; zdbt[jk * NG + g] = db;
        zrup[jk * NG + g] = ru; zrupd[jk * NG + g] = rud;
        rup[g] = ru; rupd[g] = rud;
      }
    }
  }
  double ztdn[NG], zrdnd[NG], tdbt[NG], tdbt_prev[NG], zinc[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) { ztdn[g] = 1.0; zrdnd[g] = 0.0; tdbt[g] = 1.0; tdbt_prev[g] = 1.0; zinc[g] = in.adjflux * sflux[g] * prmu0; }
  for (int jk = 0; jk <= nl; ++jk) {                   // top-down: downward combination + fluxes
    const int j = jk >= 1 ? jk - 1 : 0;
    double su = 0.0, sdn = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g < ngc) {
        if (jk >= 1) { tdbt_prev[g] = tdbt[g]; tdbt[g] = zdbt[(jk - 1) * NG + g] * tdbt_prev[g]; }
        double u, d;
        sw_level(jk, tdbt[g], tdbt_prev[g], zref[j * NG + g], zrefd[j * NG + g], ztra[j * NG + g], ztrad[j * NG + g], zrup[jk * NG + g],
                 zrupd[jk * NG + g], ztdn[g], zrdnd[g], u, d);
        su += zinc[g] * u; sdn += zinc[g] * d;
      }
    }
    fu[nl - jk] += su; fd[nl - jk] += sdn;
  }
}

#ifndef ISCA_SWC_MINB
#define ISCA_SWC_MINB 1
#endif
__global__ void __launch_bounds__(32 * SWC_WARPS, ISCA_SWC_MINB) rrtmg_sw_col_kernel(const double* __restrict__ A, Tab tb, const SwBand* __restrict__ bands, ColIn in) {
  __shared__ double sfu[(KMAX + 1) * 32], sfd[(KMAX + 1) * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nl = in.nlay;
  const size_t nc = in.ncol;
  const int col_raw = blockIdx.x * 32 + lane;
  const bool live = col_raw < in.ncol;
  const int col = live ? col_raw : in.ncol - 1;
  const double cosz = in.coszen[col];
  const bool day = cosz >= 1.0e-10;                    // `if (coszen(iplon) < zepzen) ... cycle` (rrtmg_sw_rad.nomcica.f90)
  double fu[KMAX + 1], fd[KMAX + 1];
  double W[(5 * KMAX + 2 * (KMAX + 1)) * SWC_NG];
  int jp[KMAX];
  for (int lev = 0; lev <= nl; ++lev) { fu[lev] = 0.0; fd[lev] = 0.0; }
  if (day) {
    int laytrop = 0;
    for (int l = 0; l < nl; ++l) {                     // pressure indices of every layer: laytrop and the solar source layers
      const double plog = log(in.play[col + nc * l]);
      int j = (int)(36.0 - 5.0 * (plog + 0.04));
      jp[l] = j < 1 ? 1 : (j > 58 ? 58 : j);
      laytrop += plog > 4.56;
    }
    const double albedo = in.albedo[col];
    for (int ib = 0; ib < NB_SW; ++ib) {
      if (SWC_GROUP_OF_BAND[ib] != warp) continue;
      const SwBand& B = bands[ib];
      const int lsol = sw_laysolfr_jp(B, jp, nl, laytrop);
      const int nchunk = (B.ng + SWC_NG - 1) / SWC_NG, per = (B.ng + nchunk - 1) / nchunk;
      for (int c = 0; c < nchunk; ++c) {
        const int g0 = c * per, ngc = (g0 + per <= B.ng) ? per : B.ng - g0;
        sw_band_chunk(A, tb, B, g0, ngc, in, col, nl, lsol, cosz, albedo, fu, fd, W);
      }
    }
  }
  for (int w = 0; w < SWC_WARPS; ++w) {
    if (warp == w)
      for (int lev = 0; lev <= nl; ++lev) {
        if (w == 0) { sfu[lev * 32 + lane] = fu[lev]; sfd[lev * 32 + lane] = fd[lev]; }
        else { sfu[lev * 32 + lane] += fu[lev]; sfd[lev * 32 + lane] += fd[lev]; }
      }
    __syncthreads();
  }
  if (!live) return;
  for (int lev = warp; lev <= nl; lev += SWC_WARPS) {
    in.uflx[col + nc * lev] = sfu[lev * 32 + lane];
    in.dflx[col + nc * lev] = sfd[lev * 32 + lane];
  }
  for (int l = warp; l < nl; l += SWC_WARPS) {         // swhr(nlayers) = 0 in the reference
    const double f0 = sfd[l * 32 + lane] - sfu[l * 32 + lane], f1 = sfd[(l + 1) * 32 + lane] - sfu[(l + 1) * 32 + lane];
    in.hr[col + nc * l] = l == nl - 1 ? 0.0 : (f1 - f0) * in.heatfac / (in.plev[col + nc * l] - in.plev[col + nc * (l + 1)]);
  }
}

// ---- run_rrtmg glue: model layout [K][J][I] top-down (Pa) -> RRTMG layout (ncol, nlay) bottom-up (hPa) ----
struct PrepArgs {
  int ncol, K;                          // ncol = rrtm columns (= model columns / lonstep)
  const double *p_full, *p_half, *z_full, *z_half, *t, *q, *o3;
  double *play, *plev, *tlay, *tlev, *h2o, *o3v;
  double h2o_fac, o3_fac, h2o_lower_limit, t_lo, t_hi; int convert;
  // lonstep > 1 (`p_full(1:si:lonstep,:,:)`, rrtm_radiation.F90:831-846): every lonstep-th longitude; the 2-D inputs are gathered too
  int lonstep, I; size_t ncol_model;
  const double *t_surf, *albedo, *coszen; double *tsfc_s, *albedo_s, *coszen_s;
};

// interp_temp (rrtm_radiation.F90:502-544) and the reshape / unit block of run_rrtmg (:816-870)
__global__ void rrtm_prepare_kernel(PrepArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;          // rrtm column
  if (c >= a.ncol) return;
  const int K = a.K; const size_t nc = a.ncol, nm = a.ncol_model;
  size_t s = c;                                                 // model column it is taken from
  if (a.lonstep > 1) {
    const int Is = a.I / a.lonstep, j = c / Is, i = c - j * Is;
    s = (size_t)j * a.I + (size_t)i * a.lonstep;
    a.tsfc_s[c] = a.t_surf[s]; a.albedo_s[c] = a.albedo[s]; a.coszen_s[c] = a.coszen[s];
  }
  auto lim = [&](double x) { return fmin(fmax(x, a.t_lo), a.t_hi); };
  for (int k = 0; k < K; ++k) {                 // model level k (0 = top) -> rrtm layer K-1-k
    const int l = K - 1 - k;
    double tk = a.t[s + nm * k];
    a.play[c + nc * l] = a.p_full[s + nm * k] * 0.01;
    a.tlay[c + nc * l] = lim(tk);
    double q = a.q[s + nm * k];
    double v = a.convert ? (q / (1.0 - q)) * a.h2o_fac : q;
    a.h2o[c + nc * l] = fmax(v, a.h2o_lower_limit);
    a.o3v[c + nc * l] = a.o3 ? a.o3[s + nm * k] * a.o3_fac : 0.0;
    // half level k (interface above layer k) -> rrtm level K-k
    double th;
    if (k == 0) th = 0.5 * (3.0 * tk - a.t[s + nm * 1]);
    else {
      double zf0 = a.z_full[s + nm * (k - 1)], zf1 = a.z_full[s + nm * k], zh = a.z_half[s + nm * k];
      double dzk2 = 1.0 / (zf0 - zf1), dzk = (zh - zf1) * dzk2, dzk1 = (zf0 - zh) * dzk2;
      th = tk * dzk1 + a.t[s + nm * (k - 1)] * dzk;
    }
    a.tlev[c + nc * (K - k)] = lim(th);
    a.plev[c + nc * (K - k)] = a.p_half[s + nm * k] * 0.01;
  }
  {
    double zf0 = a.z_full[s + nm * (K - 2)], zf1 = a.z_full[s + nm * (K - 1)];
    double th = a.t[s + nm * (K - 2)] + (a.z_half[s + nm * K] - zf0) * (a.t[s + nm * (K - 1)] - a.t[s + nm * (K - 2)]) / (zf1 - zf0);
    a.tlev[c] = lim(th);
    a.plev[c] = a.p_half[s + nm * K] * 0.01;
  }
}
// `if(minval(phalf(:,sk+1)) .le. 0.) phalf(:,sk+1) = pfull(:,sk)*0.5` -- the top interface pressure pk(1) + bk(1)*ps is the
// same in every column (bk(1) = 0), so the reference's all-or-nothing replacement equals this per-column test
__global__ void rrtm_fix_top_kernel(int ncol, int K, const double* play, double* plev) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncol && plev[c + (size_t)ncol * K] <= 0.0) plev[c + (size_t)ncol * K] = play[c + (size_t)ncol * (K - 1)] * 0.5;
}

struct FinishArgs {
  int ncol, K;                          // ncol = model columns
  const double *swhr, *lwhr, *swu, *swd, *lwu, *lwd;
  double *tdt, *tdt_rad, *flux_sw, *flux_lw, *olr, *toa_sw;
  int lonstep, I;                       // lonstep > 1: linear interpolation in longitude, closed toroidally (rrtm_radiation.F90:918-935)
};
__global__ void rrtm_finish_kernel(FinishArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;          // model column
  if (c >= a.ncol) return;
  const int K = a.K; const size_t nm = a.ncol;
  const double daypersec = 1.0 / 86400.0;
  if (a.lonstep <= 1) {
    for (int k = 0; k < K; ++k) {
      const int l = K - 1 - k;
      double h = a.swhr[c + nm * l] * daypersec + a.lwhr[c + nm * l] * daypersec;
      if (a.tdt) a.tdt[c + nm * k] += h;
      if (a.tdt_rad) a.tdt_rad[c + nm * k] = h;
    }
    if (a.flux_sw) a.flux_sw[c] = a.swd[c] - a.swu[c];
    if (a.flux_lw) a.flux_lw[c] = a.lwd[c];
    if (a.olr) a.olr[c] = a.lwu[c + nm * K] - a.lwd[c + nm * K];
    if (a.toa_sw) a.toa_sw[c] = a.swd[c + nm * K] - a.swu[c + nm * K];
    return;
  }
  const int Is = a.I / a.lonstep, j = c / a.I, i = c - j * a.I;
  const int ic = i / a.lonstep, ij = i - ic * a.lonstep, i1 = ic + 1 < Is ? ic + 1 : 0;
  const double di = (double)ij * (1.0 / (double)a.lonstep);     // di = (ij-1)*dlon, dlon = 1./lonstep
  const size_t nr = nm / a.lonstep, c0 = (size_t)j * Is + ic, c1 = (size_t)j * Is + i1;
  for (int k = 0; k < K; ++k) {
    const int l = K - 1 - k;
    double s0 = a.swhr[c0 + nr * l] * daypersec, l0 = a.lwhr[c0 + nr * l] * daypersec;
    double s1 = a.swhr[c1 + nr * l] * daypersec, l1 = a.lwhr[c1 + nr * l] * daypersec;
    double h = di * (s1 + l1) + (1.0 - di) * (s0 + l0);
    if (a.tdt) a.tdt[c + nm * k] += h;
    if (a.tdt_rad) a.tdt_rad[c + nm * k] = h;
  }
  if (a.flux_sw) a.flux_sw[c] = di * (a.swd[c1] - a.swu[c1]) + (1.0 - di) * (a.swd[c0] - a.swu[c0]);
  if (a.flux_lw) a.flux_lw[c] = di * a.lwd[c1] + (1.0 - di) * a.lwd[c0];
  if (a.olr) a.olr[c] = di * (a.lwu[c1 + nr * K] - a.lwd[c1 + nr * K]) + (1.0 - di) * (a.lwu[c0 + nr * K] - a.lwd[c0 + nr * K]);
  if (a.toa_sw) a.toa_sw[c] = di * (a.swd[c1 + nr * K] - a.swu[c1 + nr * K]) + (1.0 - di) * (a.swd[c0 + nr * K] - a.swu[c0 + nr * K]);
}

// diurnal_solar_2d (astronomy.f90:1123-1410; allow_negative_cosz absent): the chain of `where` statements in order
__global__ void coszen_kernel(int n, const double* __restrict__ lat, const double* __restrict__ lon, double gmt, double dec, double dt,
                              int frierson, double del_sol, double del_sw, double* __restrict__ cosz_out, double* __restrict__ fracday_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double PI = 3.14159265358979323846, twopi = 2.0 * PI;
  double la = lat[i];
  if (frierson) {                         // rrtm_radiation.F90:686-689
    double sl = sin(la);
    double p2 = (1.0 - 3.0 * sl * sl) / 4.0;
    cosz_out[i] = 0.25 * (1.0 + del_sol * p2 + del_sw * sl);
    if (fracday_out) fracday_out[i] = 1.0;
    return;
  }
  double aa = sin(la) * sin(dec), bb = cos(la) * cos(dec);
  double t = gmt + lon[i] - PI;
  if (t >= PI) t -= twopi;
  if (t < -PI) t += twopi;
  // half_day
  double l2 = la;
  if (la == 0.5 * PI) l2 = la - 1.0e-05;
  if (la == -0.5 * PI) l2 = la + 1.0e-05;
  double chd = -tan(l2) * tan(dec);
  double h = chd <= -1.0 ? PI : (chd >= 1.0 ? 0.0 : acos(chd));
  double cosz, fracday;
  if (dt > 0.0) {
    double tt = t + dt, st = sin(t), stt = sin(tt), sh = sin(h);
    cosz = 0.0;
    if (t < -h && tt < -h) cosz = 0.0;
    if (t < -h && fabs(tt) <= h) cosz = (aa * (tt + h) / (tt - t)) + bb * (stt + sh) / (tt - t);
    if (t < -h && h != 0.0 && h < tt) cosz = aa * (2. * h) / (tt - t) + bb * (sh + sh) / (tt - t);
    if (fabs(t) <= h && fabs(tt) <= h) cosz = aa + bb * (stt - st) / (tt - t);
    if (fabs(t) <= h && h < tt) cosz = (aa * (h - t) / (tt - t)) + bb * (sh - st) / (tt - t);
    if (twopi - h < tt && t <= h) cosz = aa * ((tt + (2. * h) - t - twopi) / (tt - t)) + bb * (((sh - st) / (tt - t)) + ((stt + sh) / (tt - t)));
    if (h < t && twopi - h >= tt) cosz = 0.0;
    if (h < t && twopi - h < tt && tt < twopi + h) cosz = aa * (tt + h - twopi) / (tt - t) + bb * (stt + sh) / (tt - t);
    if (h < t && twopi - h < tt && tt > twopi + h) cosz = aa * (2. * h) / (tt - t) + bb * (sh + sh) / (tt - t);
    fracday = 0.0;
    if (t < -h && tt < -h) fracday = 0.0;
    if (t < -h && fabs(tt) <= h) fracday = (tt + h) / dt;
    if (t < -h && h < tt) fracday = (h + h) / dt;
    if (fabs(t) <= h && fabs(tt) <= h) fracday = (tt - t) / dt;
    if (fabs(t) <= h && h < tt) fracday = (h - t) / dt;
    if (h < t) fracday = 0.0;
    if (twopi - h < tt) fracday = fracday + (tt + h - twopi) / dt;
  } else {
    if (fabs(t) < h) { cosz = aa + bb * cos(t); fracday = 1.0; } else { cosz = 0.0; fracday = 0.0; }
  }
  cosz_out[i] = fmax(0.0, cosz);
  if (fracday_out) fracday_out[i] = fracday;
}


}  // namespace rrtm_k
