// spectral.h -- argument bundle of the spectral step kernels.
#pragma once
#include "device.h"

namespace isca {

struct SpecStepArgs {
  // forward-transform output batch [T][LpB]: level offsets of dt_ln_ps(1), dt_T(K), A=dt_u/cos(K), B=dt_v/cos(K), Phi+KE(K)
  const double2* specB; int LpB; int oLnps, oT, oA, oB, oPhi;
  // state, [T][K] (3-D) or [T] (2-D); *_w are the write views (same slot as *_cur)
  const double2 *vors_prev, *divs_prev, *ts_prev, *lnps_prev;
  const double2 *vors_cur, *divs_cur, *ts_cur, *lnps_cur;
  double2 *vors_cur_w, *divs_cur_w, *ts_cur_w, *lnps_cur_w;
  double2 *vors_fut, *divs_fut, *ts_fut, *lnps_fut;
  // work arrays
  double2 *dt_vors, *w_div, *w_T, *w_lnps;
  // next inverse batch [T][LpC]: offsets of vor(K), div(K), ucos(K), vcos(K), T(K), lnps(1)
  double2* specC; int LpC; int cVor, cDiv, cU, cV, cT, cLnps;
  int cDxT, cDyT, cDxL, cDyL;   // gradient levels for the next step: dx T (K), dy T (K), dx ln ps, dy ln ps
  int fuse_robert_b;
  int use_implicit;
  // optional copies of the final spectral tendencies (parity tests)
  int keep_tend; double2 *k_dt_vors, *k_dt_divs, *k_dt_ts, *k_dt_lnps;
};

void launch_spec_gradient(const DevTables& t, const double2* src, int Ls, int nlev, double2* dst, int Lp,
                          int dx_off, int dy_off, cudaStream_t st);
void launch_spec_ucos_vcos(const DevTables& t, double2* buf, int Lp, int nlev, int vor_off, int div_off,
                           int u_off, int v_off, cudaStream_t st);
void launch_spec_vor_div(const DevTables& t, const double2* buf, int Lp, int nlev, int a_off, int b_off,
                         double2* out, int Lo, int vor_off, int div_off, cudaStream_t st);
void launch_spec_step(const DevTables& t, const Params& pr, const SpecStepArgs& a, cudaStream_t st);
// stand-alone implicit_correction (implicit.F90:241-325) on packed spectra [T][K] / [T]; in place on dt_divs, dt_ts, dt_lnps
void launch_implicit_correction(const DevTables& t, const Params& pr, double2* dt_divs, double2* dt_ts, double2* dt_lnps,
                                const double2* divs_prev, const double2* divs_cur, const double2* ts_prev, const double2* ts_cur,
                                const double2* lnps_prev, const double2* lnps_cur, cudaStream_t st);
void launch_spec_robert_b(double2* aprev, const double2* acur, size_t n, double rc, double raw, cudaStream_t st);

}  // namespace isca
