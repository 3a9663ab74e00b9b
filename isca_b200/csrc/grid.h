// grid.h -- argument bundle of the grid-space column kernels.
#pragma once
#include "device.h"

namespace isca {

// device scalar slots
enum {
  // [0..4]: one SUM all-reduce per step -- grid_step partials of the previous level, then the fused fixer partials
  SC_SUM_PS_PREV = 0, SC_SUM_EN_PREV = 1,
  SC_SUM_PS_FUT = 2,                                // mass fixer: sum w*ps(future)
  SC_SUM_EN_FUT = 3, SC_SUM_EN_FUTB = 4,            // energy fixer: sum w*e*dpk and sum w*e*dbk*ps (energy = A + factor*B)
  SC_NTMIN = 5, SC_TMAX = 6,                        // one MAX all-reduce: -min T, max T (range check)
  SC_MEAN_PS_PREV = 7, SC_MASS_FACTOR = 8, SC_MEAN_EN_PREV = 9, SC_T_CORR = 10, SC_T_FLAG = 11, SC_TMIN = 12,
  SC_W_PREV = 13, SC_W_ALL = 14, SC_W_CORR = 15, SC_W_NOT = 16,   // water fixer sums
  SC_TSHIFT0 = 18, SC_TSHIFT1 = 19,   // pending energy-fixer temperature increment of storage slot 0 / 1 (applied on read)
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
// compute_corrections (spectral_dynamics.F90:1213-1302), mass and energy fixers fused: one column pass, one reduction,
// one apply.  part: [5][Jloc*I]
void launch_colsum_fixers(const DevTables& t, const Params& pr, const double* u, const double* v, const double* T,
                          const double* ps, double* part, cudaStream_t st);
void launch_apply_fixers(const DevTables& t, const Params& pr, int slot_fut, double* ps, double2* lnps_fut, double2* lnps_cur,
                         double2* ts_fut, double2* ts_cur, double rc_raw, double* scal, double denom, int owns_m0, int do_mass,
                         int do_energy, cudaStream_t st);
void launch_materialize_t(const DevTables& t, double* T, double* scal, int slot, cudaStream_t st);
void launch_press_heights(const DevTables& t, const Params& pr, const double* T, const double* ps, const double* phis,
                          double* p_full, double* p_half, double* z_full, double* z_half, cudaStream_t st);
void launch_divide_by_cos(const DevTables& t, double* f, int nlev, cudaStream_t st);

}  // namespace isca
