// grid.h -- argument bundle of the grid-space column kernels.
#pragma once
#include "device.h"

namespace isca {

// device scalar slots
enum {
  // [0..9]: the ONE all-reduce (SUM) of a step on several ranks carries every global sum -- the previous-level sums of
  // initialize_corrections (grid_step partials), the previous-level water (tracer PPM sweep partials) and the fused fixer partials of the
  // future level; everything that depends on the mass factor f enters as A + f*B
  SC_SUM_PS_PREV = 0, SC_SUM_EN_PREV = 1, SC_W_PREV = 2,
  SC_SUM_PS_FUT = 3,                                // mass fixer: sum w*ps(future)
  SC_SUM_EN_FUT = 4, SC_SUM_EN_FUTB = 5,            // energy fixer: sum w*e*dpk and sum w*e*dbk*ps (energy = A + f*B)
  SC_WA_CORR = 6, SC_WB_CORR = 7, SC_WA_NOT = 8, SC_WB_NOT = 9,   // water fixer: sum w*q*dpk and sum w*q*dbk*ps below / above the limit
  SC_NSUM = 10,
  SC_NTMIN = 10, SC_TMAX = 11,                      // -min T, max T of this rank's latitudes (valid_range_t check: local)
  SC_MEAN_PS_PREV = 12, SC_MASS_FACTOR = 13, SC_MEAN_EN_PREV = 14, SC_T_CORR = 15, SC_T_FLAG = 16, SC_TMIN = 17,
  SC_W_ALL = 18, SC_W_CORR = 19, SC_W_NOT = 20,     // water sums of the future level, formed by apply_fixers: A + f*B
  SC_TSHIFT0 = 22, SC_TSHIFT1 = 23,   // pending energy-fixer temperature increment of storage slot 0 / 1 (applied on read)
  SC_COUNT = 24
};

struct GridStepArgs {
  // state (planes [K][Jloc][I] / [Jloc][I])
  const double *u_cur, *v_cur, *t_cur, *u_prev, *v_prev, *t_prev, *vor_cur, *div_cur;
  const double *ps_cur, *ps_prev, *phis;
  // gradients from the inverse batch
  const double *dx_t, *dy_t, *dx_lnps, *dy_lnps;
  // optional externally supplied tendencies (spectral_dynamics API), NULL otherwise
  const double *dt_u_in, *dt_v_in, *dt_t_in;
  // outputs (forward batch planes)
  double *out_A, *out_B, *out_T, *out_phi, *dt_lnps;
  double *wg_full;          // may be NULL
  double *wg;               // [K+1] interface mass fluxes (needed by the tracer PPM step); may be NULL
  double *part;             // [2][Jloc*I] per-column partials for the global means
  const double *scal;       // device scalars (pending temperature shifts)
  int slot_cur, slot_prev;
};

void launch_grid_step(const DevTables& t, const Params& pr, const GridStepArgs& a, cudaStream_t st);
void launch_reduce(const double* part, size_t n, int nq, const int* ops, double* out, double* tmp, cudaStream_t st);  // tmp: [nq*128]
// compute_corrections (spectral_dynamics.F90:1213-1302), mass, energy and water fixers fused: one column pass, one reduction,
// one apply.  part: [9][Jloc*I]; wpart: the four column sums of the tracer PPM sweep (NULL without a tracer)
void launch_colsum_fixers(const DevTables& t, const Params& pr, const double* u, const double* v, const double* T,
                          const double* ps, const double* wpart, double* part, cudaStream_t st);
void launch_apply_fixers(const DevTables& t, const Params& pr, int slot_fut, double* ps, double2* lnps_fut, double2* lnps_cur,
                         double2* ts_fut, double2* ts_cur, double rc_raw, double* scal, double denom, int owns_m0, int do_mass,
                         int do_energy, cudaStream_t st);
void launch_materialize_t(const DevTables& t, double* T, double* scal, int slot, cudaStream_t st);
void launch_press_heights(const DevTables& t, const Params& pr, const double* T, const double* ps, const double* phis,
                          double* p_full, double* p_half, double* z_full, double* z_half, cudaStream_t st);
void launch_divide_by_cos(const DevTables& t, double* f, int nlev, cudaStream_t st);

}  // namespace isca
