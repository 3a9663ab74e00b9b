// core_internal.h -- C++ interface between the dynamical core (core.cu) and the moist-model driver (moist_model.cu):
// raw device views of the resident state and a step that takes device-resident physics tendencies.
#pragma once
#include "../../include/isca_b200.h"
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>

struct IscaCoreView {
  int I, Jloc, K, j0, previous, current, num_tracers, nranks;
  double *u[2], *v[2], *T[2], *q[2], *ps[2];   // [K][Jloc][I] (ps: [Jloc][I]) per time-level slot
  double *phis, *wg_full;
  const double* rad_lat;                       // [Jloc] latitudes (radians) of this rank
  cudaStream_t st;
  double dt_atmos, grav;
};

extern "C" {
// fills the view; applies any pending energy-fixer temperature shift so that T[slot] can be read directly.  0 = ok
int isca_core_view(IscaHandle h, IscaCoreView* v);
// compute_pressures_and_heights (press_and_geopot.F90:363-387) of one time-level slot into device arrays
int isca_core_press_heights(IscaHandle h, int slot, double* p_full, double* p_half, double* z_full, double* z_half);
// one atmosphere() step with device-resident tendencies dt_ug, dt_vg, dt_tg, dt_tracers(sphum) (may be null = zero)
int isca_core_step_ext(IscaHandle h, const double* dtu, const double* dtv, const double* dtt, const double* dtq);
// valid-temperature-range check of the last steps (spectral_dynamics.F90 FATAL); 0 = ok
int isca_core_check(IscaHandle h);
// device pointer and element count of a grid field of isca_b200_get_field (materialised on the core's stream where needed); 0 = ok
int isca_core_field_device(IscaHandle h, int field_id, int level, const double** ptr, size_t* count);
// per-kernel-group CUDA-event marks shared with the moist-model driver (isca_b200_moist_profile_step): begin clears the marks and
// records "start"; mark() records an event named `name` on the core's stream (no-op unless profiling); end synchronises, adds the
// elapsed ms between consecutive marks to acc (keyed by the later mark's name, first-seen order kept in `order`) and frees the events
void isca_core_profile_begin(IscaHandle h);
void isca_core_mark(IscaHandle h, const char* name);
int isca_core_profile_end(IscaHandle h, std::vector<std::string>& order, std::map<std::string, double>& acc);
}
