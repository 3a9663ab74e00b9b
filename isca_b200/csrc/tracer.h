// tracer.h -- grid tracer (sphum) path: Lin-Rood A-grid horizontal advection, PPM vertical advection,
// Held-Suarez tracer source/sink, water fixer.
#pragma once
#include "device.h"

namespace isca {

// finite-volume grid metrics of fv_advection_init (atmos_spectral/model/fv_advection.F90:59-121); device pointers
struct FvTables {
  const double *c, *cc;            // cos at cell centres [J], at boundaries [J+1]
  const double *dy;                // dy(-1:J+2) * radius, stored with offset 1: dy[j+1] = dy(j)
  const double *dyy;               // dyy(1:J+1) * radius, dyy[j-1] = dyy(j)
  const double *dy_plus, *dy_minus;  // (0:J+1)
  // reciprocals: rdxc[j] = 1/(dx*c[j]) [J], rdyy = 1/dyy [J+1], rcdy[j] = 1/(c[j]*dy(j+1)) [J] (dy of the row itself), rdy = 1/dy [J+4]
  const double *rdxc, *rdyy, *rcdy, *rdy;
  double dx;
};

struct TracerArgs {
  // state
  const double *q_prev, *q_cur; double *q_cur_w, *q_fut;
  const double *u_cur, *v_cur, *ps_cur, *ps_prev, *ps_fut;
  const double *wg;                // [K+1] planes: downward mass flux at the interfaces (four_in_one)
  double *tr1;                     // work plane [K]: tr_future after the horizontal step
  double *part;                    // per-column partials [3][Jloc*I] (shared with the other global means of the step)
  double *wpart;                   // per-column water-fixer sums of the PPM sweep [4][Jloc*I]
  // latitude halos (nranks > 1; fv_advection.F90:161-162 exchanges 2 rows N and S): [3][K][2][I] = tr0, u_cur, v_cur rows
  // of the southern neighbour (global rows j0-2, j0-1) / northern neighbour (j0+Jloc, j0+Jloc+1); send_* are the packed
  // edge rows of this rank
  double *halo_s, *halo_n, *send_s, *send_n;
  double delta_t, trflux, trdamp, robert_coeff, raw_filter_coeff, water_limit;
  int physics_on;
  const double* dt_q_in;           // externally computed tendency (moist physics), added to the source; may be null
};

void launch_tracer_halo_pack(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st);
void launch_tracer_horiz(const DevTables& t, const FvTables& f, const Params& pr, const TracerArgs& a, cudaStream_t st);
void launch_tracer_ppm(const DevTables& t, const Params& pr, const TracerArgs& a, cudaStream_t st);
void launch_tracer_water_apply(const DevTables& t, const Params& pr, const TracerArgs& a, const double* scal, double denom,
                               int do_water, cudaStream_t st);

}  // namespace isca
