// device.h -- device-side table bundle and kernel launch declarations.
#pragma once
#include "common.h"
#include <cuda_runtime.h>

namespace isca {

// per-level constants of the pure-sigma fast path (grid.cu: grid_step_sigma_kernel); device pointers, [K] (b: [K+1])
struct SigmaTables {
  const double *b;      // bk
  const double *db;     // bk(k+1) - bk(k)
  const double *rdb;    // 1 / db
  const double *al;     // alpha(k) = ln_p_half(k+1) - ln_p_full(k)            (dlog_1)
  const double *d3;     // ln bk(k+1) - ln bk(k)                                (dlog_3; 0 at the top level)
  const double *lf;     // ln(p_full/ps)
  const double *pf;     // p_full/ps
  const double *pfk;    // (p_full/ps)^kappa
  const double *x1c;    // (bk(k+1)*dlog_1 + bk(k)*dlog_2) / db                 (x1 * ps)
  // PPM vertical advection (tracer.cu): slope_z weights (levels 1..K-2) and compute_weights z1..z3 (levels 2..K-2) of
  // atmos_shared/vert_advection/vert_advection.F90:504-629 with dz = db (the surface pressure cancels)
  const double *ppm_c1, *ppm_c2, *ppm_z1, *ppm_z2, *ppm_z3;
};

// All device-resident constant tables.  Pointers are device pointers.
struct DevTables {
  GeomDev g;
  SigmaTables sig;
  // Gaussian grid [J] (global index) and vertical coordinate
  const double *sin_lat, *cos_lat, *cosm_lat, *wts_lat, *coriolis, *rad_lat;
  const double *pk, *bk, *dpk, *dbk;           // [K+1],[K+1],[K],[K]
  const double *ln_bk;                         // [K+1] log(bk) (0 where bk == 0); used when pk == 0 everywhere
  // Legendre tables, packed [T][Jh]
  const double *leg, *legw;
  // per packed row [T]
  const double *eigen, *coef_uvm, *coef_uvc, *coef_uvp, *coef_alpm, *coef_alpp, *coef_dym, *coef_dx, *coef_dyp;
  const double *trunc_mask, *damping, *damping_vor, *damping_div, *eddy_sponge, *zmu_sponge, *zmv_sponge;
  const int *row_n;                            // [T] n of packed row   (g.row_m: local mi)
  // semi-implicit
  const double *ref_ln_p_half, *ref_ln_p_full, *ref_t, *h_impl;   // [K+1],[K],[K],[K]
  const double *wave_matrix;                   // [M+1][K][K] for the current xi
  // FFT twiddles [I] complex: exp(-2 pi i k / I)
  const double2 *twiddle;
};

// One level of a batched transform: where the grid-side plane lives and what to apply.
//   inverse FFT epilogue ops: 0 store, 1 multiply by 1/cos(lat), 2 exp()
//   forward FFT prologue ops: 0 load
struct LevDesc {
  double* ptr;      // plane [Jloc][I]
  int op;
  int pad;
};

// scalar physical parameters used by the column / spectral kernels
struct Params {
  double rdgas, kappa, cp_air, grav, radius, omega;
  double ref_ps, xi, delta_t, dt_atmos;
  double robert_coeff, raw_filter_coeff;
  double virtual_factor;
  // Held-Suarez
  double tka, tks, vkf, sigma_b, t_zero, t_strat, delh, delv, eps, P00;
  int do_conserve_energy, no_forcing, physics_on;
  int first_step;                // previous == current
  int pk0_zero, pkbk0_zero;      // pk(1)==0 ; pk(1)==0 && bk(1)==0
  int pure_sigma;                // pk == 0 at every half level
  int sigma_fast;                // pure sigma, zero top, bk(K+1) == 1: grid_step_sigma_kernel is used
  double vr_tmin, vr_tmax;
};

// ---- launches (all asynchronous on `st`) ----------------------------------------------------
// Legendre: spec batch [T][Lp] complex  <->  Fourier buffer (m-owner layout), C = 2*Lp doubles
// ct_begin/ct_count: range of 32-column (16-level) tiles; lev_begin: first level (FFT processes [lev_begin, nlev))
void launch_legendre_inv(const DevTables& t, const double2* spec, double* four, int Lp, cudaStream_t st, int ct_begin = 0, int ct_count = -1);
void launch_legendre_fwd(const DevTables& t, const double* four, double2* spec, int Lp,
                         const unsigned char* lev_trunc, cudaStream_t st, int ct_begin = 0, int ct_count = -1);
// FFT: Fourier buffer (lat-owner layout) <-> grid planes described by LevDesc[nlev]
void launch_fft_inv(const DevTables& t, const double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin = 0);
void launch_fft_fwd(const DevTables& t, double* four, const LevDesc* levs, int nlev, int Lp, cudaStream_t st, int lev_begin = 0);

// layout conversion between the reference's rectangular (m,n,lev) arrays and the packed layout
void launch_pack_spec(const DevTables& t, const double2* rect, double2* packed, int nlev, int Lp, int lev0, cudaStream_t st);
void launch_unpack_spec(const DevTables& t, const double2* packed, double2* rect, int nlev, int Lp, int lev0, cudaStream_t st);

}  // namespace isca
