// fft.cu -- longitude FFT as a shared-memory Stockham pass (fp64).
//
// Replaces fft_grid_to_fourier / fft_fourier_to_grid (shared/fft/fft.F90:483-591, 600-718) and the
// Temperton FFT99 passes behind them (shared/fft/fft99.F90).  Convention (fft99.F90:195-209):
//   forward  c_k = (1/N) sum_j x_j exp(-2 pi i j k / N)      (only k <= num_fourier kept, transforms.F90:509)
//   inverse  x_j = sum_{k=0}^{N-1} c_k exp(+2 pi i j k / N)  with Hermitian completion, c_k = 0 for k > num_fourier
// N = lon_max is a power of two.  A real transform of length N is a complex transform of length H = N/2
// plus a split/merge step.  The complex transform is a Stockham autosort FFT with two or three
// high-radix passes (radix 16/8/4 butterflies held entirely in registers):
//     pass 1 reads its inputs straight from global memory (forward) or from the staged Fourier tile
//     (inverse), the last pass writes straight to global memory (inverse) -- the data crosses shared
//     memory once per extra pass instead of once per radix-2/4 stage.
// One CTA handles LT = 8 lines = 8 consecutive batch levels at one latitude, so that the Fourier-side
// accesses ([m][lat][level] complex) are contiguous 128-byte runs.
#include "device.h"
#include <cstdlib>

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

// twiddle exp(SIGN * 2 pi i q / I): table holds exp(-2 pi i q / I), q < I
template <int SIGN>
__device__ __forceinline__ double2 tw(const double2* __restrict__ table, int q) {
  double2 w = __ldg(table + q);
  if (SIGN > 0) w.y = -w.y;
  return w;
}

// ---- in-register DFT of length N in {2,4,8,16}, natural order in and out ----------------------
// exp(-2 pi i k / 16), k < 8
__device__ __forceinline__ double2 w16(int k) {
  constexpr double c[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                           0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
  constexpr double s[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                           1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  return make_double2(c[k], -s[k]);
}

template <int N, int SIGN>
struct Dft {
  static __device__ __forceinline__ void run(double2 (&v)[N]) {
    double2 e[N / 2], o[N / 2];
#pragma unroll
    for (int k = 0; k < N / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    Dft<N / 2, SIGN>::run(e);
    Dft<N / 2, SIGN>::run(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      double2 w = w16(k * (16 / N));
      if (SIGN > 0) w.y = -w.y;
      double2 t;
      if (k == 0) t = o[0];
      else if (4 * k == N) t = (SIGN < 0) ? make_double2(o[k].y, -o[k].x) : make_double2(-o[k].y, o[k].x);
      else t = cmul(o[k], w);
      v[k] = cadd(e[k], t);
      v[k + N / 2] = csub(e[k], t);
    }
  }
};
template <int SIGN>
struct Dft<2, SIGN> {
  static __device__ __forceinline__ void run(double2 (&v)[2]) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  }
};
template <int SIGN>
struct Dft<1, SIGN> {
  static __device__ __forceinline__ void run(double2 (&)[1]) {}
};

constexpr int FFT_LT = 8;     // lines (levels) per CTA

template <int H, int R1>
struct FftShape {
  static constexpr int PADP = R1;                                   // one pad element every R1 (first-pass write stride)
  static constexpr int LS = ((H + H / R1 + 1) | 1);                 // odd line stride (in double2): conflict-free across lines
  __device__ static __forceinline__ int pad(int i) { return i + i / PADP; }
};

// One Stockham pass of radix R (Ns = product of the earlier radices) split in two halves so that a
// single shared-memory buffer per line suffices: every thread first pulls all its butterflies into
// registers (read half), the CTA synchronises, then the results are scattered back (write half).
template <int H, int R, int R1, int SIGN, int Q>
__device__ __forceinline__ void pass_read(const double2* __restrict__ src, double2 (&v)[(H / R + Q - 1) / Q][R], int Ns, int I,
                                          int lt, const double2* __restrict__ table) {
  typedef FftShape<H, R1> S;
  constexpr int NB = (H / R + Q - 1) / Q;
  const int stride_tw = I / (Ns * R);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = lt + b * Q;
    if (j < H / R) {
      const int k = j & (Ns - 1);
#pragma unroll
      for (int r = 0; r < R; ++r) v[b][r] = src[S::pad(j + r * (H / R))];
#pragma unroll
      for (int r = 1; r < R; ++r) v[b][r] = cmul(v[b][r], tw<SIGN>(table, k * r * stride_tw));
      Dft<R, SIGN>::run(v[b]);
    }
  }
}
template <int H, int R, int R1, int Q>
__device__ __forceinline__ void pass_write(double2* __restrict__ dst, const double2 (&v)[(H / R + Q - 1) / Q][R], int Ns, int lt) {
  typedef FftShape<H, R1> S;
  constexpr int NB = (H / R + Q - 1) / Q;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = lt + b * Q;
    if (j < H / R) {
      const int k = j & (Ns - 1);
      const int j0 = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) dst[S::pad(j0 + r * Ns)] = v[b][r];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inverse: Fourier (layout B) -> grid planes.   grid = (ceil(nlev/LT), Jloc), block = LT * Q
// ---------------------------------------------------------------------------------------------
template <int H, int R1, int R2, int R3, int Q, int MINB>
__global__ void __launch_bounds__(FFT_LT * Q, MINB)
fft_inv_kernel(DevTables t, const double* __restrict__ four, const LevDesc* __restrict__ levs, int nlev, int Lp, int lev_begin) {
  typedef FftShape<H, R1> S;
  constexpr int I = 2 * H;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  double2* buf = reinterpret_cast<double2*>(fft_smem);              // [LT][LS]
  const int lev0 = lev_begin + blockIdx.x * FFT_LT;
  const int jl = blockIdx.y;
  const int tid = threadIdx.x;

  // thread -> (line, lt) with the LINE index fastest: a warp is 8 lines (= 8 consecutive levels, one 128-byte run of the Fourier
  // buffer per wavenumber) x 4 butterfly slots, so a warp-wide load touches 4 cache lines instead of 16 (the L1 tag stage, one
  // line per cycle, was the limiter: profiles/r01h), and the scattered twiddle loads of a warp collapse to 4 addresses.  Shared
  // memory stays conflict-free: a quarter-warp is 8 lines at the same offset and the line stride LS is odd.  (The forward
  // kernel keeps lt fastest: its global loads are the grid rows, 256-byte runs per line; measured 0.083 vs 0.098 ms.)
  const int line = tid % FFT_LT, lt = tid / FFT_LT;
  double2* A = buf + line * S::LS;
  const int lev = lev0 + line;

  // 1.+2. pass 1 (radix R1, Ns = 1) straight from global memory: the merged spectrum
  //    Z[k] = (X[k] + conj X[H-k]) + i w^{-k} (X[k] - conj X[H-k]),  k < H,   X[k] = 0 for k > M (transforms.F90:424)
  //    A warp covers 8 adjacent levels x 4 consecutive m: whole 128-byte lines, and all
  //    2*R1 loads of a thread are independent (no staging pass, no barrier before the butterflies).
  {
    constexpr int NB1 = (H / R1 + Q - 1) / Q;
    double2 v[NB1][R1];
    const bool live = lev < nlev;
    const double* base = four + (size_t)jl * C + 2 * lev;
    const size_t mstride = (size_t)g.Jloc * C;
#pragma unroll
    for (int b = 0; b < NB1; ++b) {
      const int j = lt + b * Q;
      if (j < H / R1) {
        constexpr int RC = (R1 > 8) ? 8 : R1;          // loads are issued in chunks of RC butterflies legs (register budget)
#pragma unroll
        for (int r0 = 0; r0 < R1; r0 += RC) {
          double2 xk[RC], xc[RC];
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            const int k = j + (r0 + q) * (H / R1);
            const int kc = H - k;
            xk[q] = make_double2(0.0, 0.0); xc[q] = make_double2(0.0, 0.0);
            if (live && k <= g.M) xk[q] = *reinterpret_cast<const double2*>(base + (size_t)g.pos[k] * mstride);
            if (live && kc <= g.M) xc[q] = *reinterpret_cast<const double2*>(base + (size_t)g.pos[kc] * mstride);
          }
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            const int k = j + (r0 + q) * (H / R1);
            double2 a = xk[q];
            const double2 c = cconj(xc[q]);
            if (k == 0) a.y = 0.0;
            const double2 e = cadd(a, c), o = csub(a, c);
            const double2 wo = cmul(o, tw<+1>(t.twiddle, k));
            v[b][r0 + q] = make_double2(e.x - wo.y, e.y + wo.x);     // e + i*wo
          }
        }
        Dft<R1, +1>::run(v[b]);
      }
    }
    pass_write<H, R1, R1, Q>(A, v, 1, lt);
  }
  __syncthreads();
  int Ns = R1;
  if (R3 > 1) {
    constexpr int NB2 = (H / R2 + Q - 1) / Q;
    double2 v[NB2][R2];
    pass_read<H, R2, R1, +1, Q>(A, v, Ns, I, lt, t.twiddle);
    __syncthreads();
    pass_write<H, R2, R1, Q>(A, v, Ns, lt);
    __syncthreads();
    Ns *= R2;
  }
  // 3. last pass: smem -> registers -> global (x[2n] = Re z[n], x[2n+1] = Im z[n])
  constexpr int RL = (R3 > 1) ? R3 : R2;
  {
    constexpr int NBL = (H / RL + Q - 1) / Q;
    double2 v[NBL][RL];
    pass_read<H, RL, R1, +1, Q>(A, v, Ns, I, lt, t.twiddle);
    if (lev < nlev) {
      const LevDesc d = levs[lev];
      double2* out = reinterpret_cast<double2*>(d.ptr + (size_t)jl * I);
      const double sc = (d.op == 1) ? t.cosm_lat[g.j0 + jl] : 1.0;
#pragma unroll
      for (int b = 0; b < NBL; ++b) {
        const int j = lt + b * Q;                    // Ns * RL == H  ->  j0 = j, outputs at j + r * Ns
        if (j < H / RL) {
#pragma unroll
          for (int r = 0; r < RL; ++r) {
            double2 z = v[b][r];
            if (d.op == 2) { z.x = exp(z.x); z.y = exp(z.y); }
            else if (d.op == 1) { z.x *= sc; z.y *= sc; }
            out[j + r * Ns] = z;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward: grid planes -> Fourier (layout B)
// ---------------------------------------------------------------------------------------------
template <int H, int R1, int R2, int R3, int Q, int MINB>
__global__ void __launch_bounds__(FFT_LT * Q, MINB)
fft_fwd_kernel(DevTables t, double* __restrict__ four, const LevDesc* __restrict__ levs, int nlev, int Lp, int lev_begin) {
  typedef FftShape<H, R1> S;
  constexpr int I = 2 * H;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const GeomDev& g = t.g;
  const int C = 2 * Lp;
  double2* buf = reinterpret_cast<double2*>(fft_smem);              // [LT][LS]
  int* s_row = reinterpret_cast<int*>(buf + FFT_LT * S::LS);        // [M+1] destination row of wavenumber m
  int* s_own = s_row + (g.M + 1);                                   // [M+1] destination rank (peer-memory mode)
  const int lev0 = lev_begin + blockIdx.x * FFT_LT;
  const int jl = blockIdx.y;
  const int tid = threadIdx.x;
  constexpr int NT = FFT_LT * Q;
  // the m -> destination tables go to shared memory (visible after the __syncthreads of the passes below): global lookups in the
  // store loop put a second memory latency in front of every store
  for (int q = tid; q <= g.M; q += NT) {
    if (g.p2p) { const int r = g.owner[q]; s_own[q] = r; s_row[q] = g.rank * g.nm_rank[r] + g.lidx[q]; }
    else { s_own[q] = 0; s_row[q] = g.pos[q]; }
  }
  const int line = tid / Q, lt = tid - line * Q;
  const int lev = lev0 + line;
  double2* A = buf + line * S::LS;

  // 1. pass 1 (radix R1, Ns = 1) straight from global: z[n] = x[2n] + i x[2n+1]
  {
    constexpr int NB1 = (H / R1 + Q - 1) / Q;
    double2 v[NB1][R1];
    const double2* in = nullptr;
    if (lev < nlev) in = reinterpret_cast<const double2*>(levs[lev].ptr + (size_t)jl * I);
#pragma unroll
    for (int b = 0; b < NB1; ++b) {
      const int j = lt + b * Q;
      if (j < H / R1) {
#pragma unroll
        for (int r = 0; r < R1; ++r) v[b][r] = in ? in[j + r * (H / R1)] : make_double2(0.0, 0.0);
        Dft<R1, -1>::run(v[b]);
      }
    }
    pass_write<H, R1, R1, Q>(A, v, 1, lt);
  }
  __syncthreads();
  int Ns = R1;
  {
    constexpr int NB2 = (H / R2 + Q - 1) / Q;
    double2 v[NB2][R2];
    pass_read<H, R2, R1, -1, Q>(A, v, Ns, I, lt, t.twiddle);
    __syncthreads();
    pass_write<H, R2, R1, Q>(A, v, Ns, lt);
    __syncthreads();
    Ns *= R2;
  }
  if (R3 > 1) {
    constexpr int R3E = (R3 > 1) ? R3 : 2;
    constexpr int NB3 = (H / R3E + Q - 1) / Q;
    double2 v[NB3][R3E];
    pass_read<H, R3E, R1, -1, Q>(A, v, Ns, I, lt, t.twiddle);
    __syncthreads();
    pass_write<H, R3E, R1, Q>(A, v, Ns, lt);
    __syncthreads();
  }
  // 2. split + store: X[k] = 0.5 (Z[k] + conj Z[H-k]) - 0.5 i w^k (Z[k] - conj Z[H-k]), scaled by 1/I, k <= M;
  //    threads ordered (level fastest) so that every m writes LT complex = 128 contiguous bytes
  const double inv = 1.0 / (double)I;
#pragma unroll 4
  for (int idx = tid; idx < FFT_LT * (g.M + 1); idx += NT) {
    const int k = idx / FFT_LT, l = idx - k * FFT_LT;
    if (lev0 + l >= nlev) continue;
    const double2* Z = buf + l * S::LS;
    const double2 zk = Z[S::pad(k)];
    const double2 zc = cconj(Z[S::pad((k == 0) ? 0 : (H - k))]);
    const double2 e = cadd(zk, zc), o = csub(zk, zc);
    const double2 wo = cmul(o, tw<-1>(t.twiddle, k));
    const double2 x = make_double2(0.5 * (e.x + wo.y) * inv, 0.5 * (e.y - wo.x) * inv);
    // peer-memory mode: the row lives in the m-owner buffer of the rank that owns wavenumber k (store over NVLink)
    double* dst = (g.p2p ? g.peerA[s_own[k]] : four) + ((size_t)s_row[k] * g.Jloc + jl) * (size_t)C;
    *reinterpret_cast<double2*>(dst + 2 * (lev0 + l)) = x;
  }
}

template <int H, int R1, int R2, int R3, int Q, int MINB>
static void launch_inv_shape(const DevTables& t, const double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin) {
  typedef FftShape<H, R1> S;
  const size_t smem = sizeof(double2) * FFT_LT * S::LS;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(fft_inv_kernel<H, R1, R2, R3, Q, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid((nlev - lev_begin + FFT_LT - 1) / FFT_LT, t.g.Jloc);
  fft_inv_kernel<H, R1, R2, R3, Q, MINB><<<grid, FFT_LT * Q, smem, st>>>(t, four, levs, nlev, Lp, lev_begin);
}
template <int H, int R1, int R2, int R3, int Q, int MINB>
static void launch_fwd_shape(const DevTables& t, double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin) {
  typedef FftShape<H, R1> S;
  const size_t smem = sizeof(double2) * FFT_LT * S::LS + sizeof(int) * 2 * (t.g.M + 1);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(fft_fwd_kernel<H, R1, R2, R3, Q, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(double2) * FFT_LT * S::LS + sizeof(int) * 2 * (H + 1))); attr = true; }
  dim3 grid((nlev - lev_begin + FFT_LT - 1) / FFT_LT, t.g.Jloc);
  fft_fwd_kernel<H, R1, R2, R3, Q, MINB><<<grid, FFT_LT * Q, smem, st>>>(t, four, levs, nlev, Lp, lev_begin);
}

// radix plans: H = I/2 = R1*R2*R3, Q = H / max radix threads per line
static int fft_plan_override() {
  static int plan = -1;
  if (plan < 0) { const char* e = std::getenv("ISCA_B200_FFT_PLAN"); plan = e ? std::atoi(e) : 0; }
  return plan;
}
void launch_fft_inv(const DevTables& t, const double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin) {
  if (t.g.I == 512 && fft_plan_override() == 884) { launch_inv_shape<256, 8, 8, 4, 32, 3>(t, four, levs, nlev, Lp, st, lev_begin); return; }
  switch (t.g.I) {
    case 1024: launch_inv_shape<512, 8, 8, 8, 64, 1>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 512:  launch_inv_shape<256, 16, 16, 1, 16, 4>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 256:  launch_inv_shape<128, 8, 16, 1, 8, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 128:  launch_inv_shape<64, 8, 8, 1, 8, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 64:   launch_inv_shape<32, 4, 8, 1, 4, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 32:   launch_inv_shape<16, 4, 4, 1, 4, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    default: break;   // rejected in build_geometry
  }
}
void launch_fft_fwd(const DevTables& t, double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin) {
  if (t.g.I == 512 && fft_plan_override() == 884) { launch_fwd_shape<256, 8, 8, 4, 32, 4>(t, four, levs, nlev, Lp, st, lev_begin); return; }
  switch (t.g.I) {
    case 1024: launch_fwd_shape<512, 8, 8, 8, 64, 2>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 512:  launch_fwd_shape<256, 16, 16, 1, 16, 5>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 256:  launch_fwd_shape<128, 8, 16, 1, 8, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 128:  launch_fwd_shape<64, 8, 8, 1, 8, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 64:   launch_fwd_shape<32, 4, 8, 1, 4, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    case 32:   launch_fwd_shape<16, 4, 4, 1, 4, 8>(t, four, levs, nlev, Lp, st, lev_begin); break;
    default: break;
  }
}

}  // namespace isca
