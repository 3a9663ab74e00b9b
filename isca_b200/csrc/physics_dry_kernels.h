// physics_dry_kernels.h -- dry convective adjustment of Schneider & Walker (atmos_param/dry_convection/dry_convection.f90:105-299:
// dry_convection + capecalc), one thread per column.  Included by physics_dry.cu (nvcc) and by the test-only thread emulator
// tests/host/rrtm_emu.cpp.  Planes are [lev][col] (a warp reads 32 consecutive columns of a level); the parcel profile `tp` lives in
// the caller-provided scratch plane instead of a thread-local array.
#pragma once
#include <cmath>

namespace dryconv_k {

// Levels are 1-based in the comments (k = 1 top ... K bottom, btm = K) like the Fortran; arrays are 0-based.
__global__ void dry_convection_kernel(int ncol, int K, double tau, double gamma, double cons1, double rdgas, const double* __restrict__ tg,
                                      const double* __restrict__ p_full, const double* __restrict__ p_half, double* __restrict__ tp,
                                      double* __restrict__ dt_tg, double* __restrict__ cape_out, double* __restrict__ cin_out,
                                      int* __restrict__ lzb_out, int* __restrict__ lcl_out, int* __restrict__ err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const size_t nc = ncol;
  const int btm = K;
  auto T = [&](int k) { return tg[c + nc * (k - 1)]; };
  auto TP = [&](int k) -> double& { return tp[c + nc * (k - 1)]; };
  // capecalc: dry adiabat with the prescribed lapse-rate factor gamma, lifted from the lowest level
  TP(btm) = T(btm);
  for (int k = btm - 1; k >= 1; --k) {
    double zdpkpk = exp(cons1 * log(p_full[c + nc * (k - 1)] / p_full[c + nc * k]));
    double below = TP(k + 1);
    TP(k) = below + gamma * (below * zdpkpk - below);
  }
  double cape = 0.0, cin = 0.0;
  int lzb = btm, lcl = btm;
  for (int k = btm - 1; k >= 1; --k) {
    const double tgk = T(k);
    const double lg = log(p_half[c + nc * k] / p_half[c + nc * (k - 1)]);
    if (TP(k) > tgk) {                       // unstable parcel
      if (lzb == btm) {                      // not above a lower cloud
        cape = cape + rdgas * (TP(k) - tgk) * lg;
        if (TP(k + 1) < T(k + 1)) lcl = k;
        if (k == 1 || TP(k - 1) < T(k - 1)) lzb = k;
      } else TP(k) = tgk;                    // above cloud level: parcel temperature = ambient
    }
    if (TP(k) <= tgk) {                      // stable parcel
      if (lzb == btm) {
        if (lcl == btm) cin = cin - rdgas * (TP(k) - tgk) * lg;
      } else TP(k) = tgk;
    }
  }
  if (cin > cape) for (int k = 1; k <= K; ++k) TP(k) = T(k);
  if ((lcl != btm && lzb == btm) || lcl < lzb) atomicExch(err, (lcl < lzb) ? 65 : 64);   // the reference's two FATALs
  if (lcl == btm && lzb == btm) { cape = 0.0; cin = 0.0; }
  // dry_convection: energy-conserving shift of the parcel profile between LZB and the bottom, relaxation over tau
  double ener_int = 0.0, dp = 0.0;
  for (int k = 1; k <= K; ++k) {
    if (k >= lzb && k <= btm) {
      double dph = p_half[c + nc * k] - p_half[c + nc * (k - 1)];
      ener_int = ener_int + dph * (T(k) - TP(k));
      dp = dp + dph;
    } else TP(k) = T(k);
  }
  ener_int = ener_int / dp;
  for (int k = btm; k >= lzb; --k) TP(k) = TP(k) + ener_int;
  for (int k = 1; k <= K; ++k) dt_tg[c + nc * (k - 1)] = (TP(k) - T(k)) / tau;
  cape_out[c] = cape; cin_out[c] = cin; lzb_out[c] = lzb; lcl_out[c] = lcl;
}

}  // namespace dryconv_k
