// fft.cu -- longitude FFT as a shared-memory Stockham pass (fp64).
//
// Replaces fft_grid_to_fourier / fft_fourier_to_grid (shared/fft/fft.F90:483-591, 600-718) and the
// Temperton FFT99 passes behind them (shared/fft/fft99.F90).  Convention (fft99.F90:195-209):
//   forward  c_k = (1/N) sum_j x_j exp(-2 pi i j k / N)      (only k <= num_fourier kept, transforms.F90:509)
//   inverse  x_j = sum_{k=0}^{N-1} c_k exp(+2 pi i j k / N)  with Hermitian completion, c_k = 0 for k > num_fourier
// N = lon_max is a power of two.  A real transform of length N is done as a complex transform of
// length H = N/2 (radix-4 Stockham autosort passes + one radix-2 pass when log2(H) is odd) plus the
// split/merge step.  One CTA handles LT lines = LT consecutive batch levels at one latitude, so the
// Fourier-side accesses ([m][lat][level] complex) are contiguous runs of LT*16 bytes.
#include "device.h"

namespace isca {

// Fourier buffer, lat-owner side ("layout B"): [(pos[m]*Jloc + jl)][C]
__device__ __forceinline__ size_t fourB_index(const GeomDev& g, int m, int jl, int C) {
  return ((size_t)g.pos[m] * g.Jloc + jl) * (size_t)C;
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// twiddle exp(sign * 2 pi i q / I): table holds exp(-2 pi i q / I), q < I
template <int SIGN>
__device__ __forceinline__ double2 tw(const double2* __restrict__ table, int q) {
  double2 w = table[q];
  if (SIGN > 0) w.y = -w.y;
  return w;
}

// Complex FFT of length H on one line held in shared memory, executed by H/4 threads (lt = thread
// index within the line).  Stockham autosort: ping-pong between a and b.  Returns pointer to result.
// All threads of the CTA must call this (it contains __syncthreads()).
template <int SIGN>
__device__ double2* cfft_line(double2* a, double2* b, int H, int I, int lt, const double2* __restrict__ table) {
  const int Q = H >> 2;                      // threads per line
  int Ns = 1;
  // radix-4 passes while 4*Ns <= H
  while (Ns * 4 <= H) {
    __syncthreads();
    {
      const int j = lt;                      // j < H/4
      const int k = j & (Ns - 1);            // j % Ns
      const int stride_tw = I / (Ns * 4);    // twiddle step: exp(-+2 pi i k t / (4 Ns))
      double2 v0 = a[j], v1 = a[j + Q], v2 = a[j + 2 * Q], v3 = a[j + 3 * Q];
      if (Ns > 1) {
        v1 = cmul(v1, tw<SIGN>(table, k * stride_tw));
        v2 = cmul(v2, tw<SIGN>(table, 2 * k * stride_tw));
        v3 = cmul(v3, tw<SIGN>(table, 3 * k * stride_tw));
      }
      // radix-4 butterfly
      double2 s0 = cadd(v0, v2), d0 = csub(v0, v2), s1 = cadd(v1, v3), d1 = csub(v1, v3);
      // multiply d1 by -i (forward) or +i (inverse)
      double2 d1r = (SIGN < 0) ? make_double2(d1.y, -d1.x) : make_double2(-d1.y, d1.x);
      const int j0 = ((j - k) << 2) + k;     // (j / Ns) * Ns * 4 + k
      b[j0] = cadd(s0, s1);
      b[j0 + Ns] = cadd(d0, d1r);
      b[j0 + 2 * Ns] = csub(s0, s1);
      b[j0 + 3 * Ns] = csub(d0, d1r);
    }
    double2* tswap = a; a = b; b = tswap;
    Ns <<= 2;
  }
  if (Ns < H) {                              // one radix-2 pass: Ns * 2 == H
    __syncthreads();
    const int half = H >> 1;
    for (int j = lt; j < half; j += Q) {
      const int k = j & (Ns - 1);
      const int stride_tw = I / (Ns * 2);
      double2 v0 = a[j], v1 = a[j + half];
      v1 = cmul(v1, tw<SIGN>(table, k * stride_tw));
      const int j0 = ((j - k) << 1) + k;
      b[j0] = cadd(v0, v1);
      b[j0 + Ns] = csub(v0, v1);
    }
    double2* tswap = a; a = b; b = tswap;
  }
  __syncthreads();
  return a;
}

constexpr int FFT_LT = 8;     // lines (levels) per CTA

// ---------------------------------------------------------------------------------------------
// inverse: Fourier (layout B) -> grid planes.   grid = (ceil(nlev/LT), Jloc), block = LT * H/4
// ---------------------------------------------------------------------------------------------
__global__ void fft_inv_kernel(DevTables t, const double* __restrict__ four, const LevDesc* __restrict__ levs,
                               int nlev, int Lp) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const GeomDev& g = t.g;
  const int I = g.I, H = I >> 1, Q = H >> 2, C = 2 * Lp;
  double2* bufA = reinterpret_cast<double2*>(fft_smem);             // [LT][H + 1]  (X[0..H])
  double2* bufB = bufA + FFT_LT * (H + 1);                          // [LT][H]
  const int lev0 = blockIdx.x * FFT_LT;
  const int jl = blockIdx.y;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;

  // 1. load X[m][lev] (m <= M) -- zero elsewhere (transforms.F90:424)
  for (int idx = tid; idx < FFT_LT * (H + 1); idx += nthreads) {
    int k = idx / FFT_LT, l = idx - k * FFT_LT;
    double2 v = make_double2(0.0, 0.0);
    if (k <= g.M && lev0 + l < nlev)
      v = *reinterpret_cast<const double2*>(four + fourB_index(g, k, jl, C) + 2 * (lev0 + l));
    bufA[l * (H + 1) + k] = v;
  }
  __syncthreads();
  // 2. merge: Z[k] = (X[k] + conj X[H-k]) + i w^{-k} (X[k] - conj X[H-k]),  k < H
  const int line = tid / Q, lt = tid - line * Q;
  {
    double2* X = bufA + line * (H + 1);
    double2* Z = bufB + line * H;
    for (int k = lt; k < H; k += Q) {
      double2 xk = X[k], xc = cconj(X[H - k]);
      if (k == 0) { xk.y = 0.0; xc = make_double2(0.0, 0.0); xc = cconj(X[H]); }
      double2 e = cadd(xk, xc), o = csub(xk, xc);
      double2 wo = cmul(o, tw<+1>(t.twiddle, k));
      Z[k] = make_double2(e.x - wo.y, e.y + wo.x);       // e + i*wo
    }
  }
  // 3. complex inverse FFT of length H
  double2* res = cfft_line<+1>(bufB + line * H, bufA + line * (H + 1), H, I, lt, t.twiddle);
  // 4. store: x[2n] = Re z[n], x[2n+1] = Im z[n]
  const int lev = lev0 + line;
  if (lev < nlev) {
    const LevDesc d = levs[lev];
    double2* out = reinterpret_cast<double2*>(d.ptr + (size_t)jl * I);
    const double sc = (d.op == 1) ? t.cosm_lat[g.j0 + jl] : 1.0;
    for (int n = lt; n < H; n += Q) {
      double2 z = res[n];
      if (d.op == 2) { z.x = exp(z.x); z.y = exp(z.y); }
      else if (d.op == 1) { z.x *= sc; z.y *= sc; }
      out[n] = z;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward: grid planes -> Fourier (layout B)
// ---------------------------------------------------------------------------------------------
__global__ void fft_fwd_kernel(DevTables t, double* __restrict__ four, const LevDesc* __restrict__ levs,
                               int nlev, int Lp) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const GeomDev& g = t.g;
  const int I = g.I, H = I >> 1, Q = H >> 2, C = 2 * Lp;
  double2* bufA = reinterpret_cast<double2*>(fft_smem);             // [LT][H + 1]
  double2* bufB = bufA + FFT_LT * (H + 1);                          // [LT][H]
  const int lev0 = blockIdx.x * FFT_LT;
  const int jl = blockIdx.y;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int line = tid / Q, lt = tid - line * Q;
  const int lev = lev0 + line;

  // 1. load z[n] = x[2n] + i x[2n+1]
  {
    double2* Z = bufB + line * H;
    if (lev < nlev) {
      const LevDesc d = levs[lev];
      const double2* in = reinterpret_cast<const double2*>(d.ptr + (size_t)jl * I);
      for (int n = lt; n < H; n += Q) Z[n] = in[n];
    } else {
      for (int n = lt; n < H; n += Q) Z[n] = make_double2(0.0, 0.0);
    }
  }
  // 2. complex forward FFT
  double2* res = cfft_line<-1>(bufB + line * H, bufA + line * (H + 1), H, I, lt, t.twiddle);
  // res is either bufA-line or bufB-line; split needs Z[k] and Z[H-k] -> write X into the other buffer
  double2* other = (res == bufB + line * H) ? (bufA + line * (H + 1)) : (bufB + line * H);
  {
    const double inv = 1.0 / (double)I;
    for (int k = lt; k <= g.M; k += Q) {
      double2 zk = res[k];
      double2 zc = cconj(res[(k == 0) ? 0 : (H - k)]);
      double2 e = cadd(zk, zc), o = csub(zk, zc);
      double2 wo = cmul(o, tw<-1>(t.twiddle, k));
      // X[k] = 0.5*(e) - 0.5*i*wo
      other[k] = make_double2(0.5 * (e.x + wo.y) * inv, 0.5 * (e.y - wo.x) * inv);
    }
  }
  __syncthreads();
  // 3. store X[m][lev], m <= M: LT complex contiguous per m
  for (int idx = tid; idx < FFT_LT * (g.M + 1); idx += nthreads) {
    int k = idx / FFT_LT, l = idx - k * FFT_LT;
    if (lev0 + l < nlev) {
      double2* src_line = ((res == bufB + line * H) ? bufA : bufB);
      // all lines use the same ping-pong parity, so `other` of line l is:
      double2* ol = (res == bufB + line * H) ? (bufA + l * (H + 1)) : (bufB + l * H);
      (void)src_line;
      *reinterpret_cast<double2*>(four + fourB_index(g, k, jl, C) + 2 * (lev0 + l)) = ol[k];
    }
  }
}

static size_t fft_smem_bytes(int I) { int H = I / 2; return sizeof(double2) * FFT_LT * (2 * H + 1); }

void launch_fft_inv(const DevTables& t, const double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st) {
  const GeomDev& g = t.g;
  const int H = g.I / 2, Q = H / 4;
  size_t smem = fft_smem_bytes(g.I);
  static size_t attr = 0;
  if (smem > attr) { cudaFuncSetAttribute(fft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
  dim3 grid((nlev + FFT_LT - 1) / FFT_LT, g.Jloc);
  fft_inv_kernel<<<grid, FFT_LT * Q, smem, st>>>(t, four, levs, nlev, Lp);
}
void launch_fft_fwd(const DevTables& t, double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st) {
  const GeomDev& g = t.g;
  const int H = g.I / 2, Q = H / 4;
  size_t smem = fft_smem_bytes(g.I);
  static size_t attr = 0;
  if (smem > attr) { cudaFuncSetAttribute(fft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
  dim3 grid((nlev + FFT_LT - 1) / FFT_LT, g.Jloc);
  fft_fwd_kernel<<<grid, FFT_LT * Q, smem, st>>>(t, four, levs, nlev, Lp);
}

}  // namespace isca
