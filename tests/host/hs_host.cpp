// TEST INFRASTRUCTURE ONLY -- serial host build of the hs_forcing column arithmetic of isca_b200/csrc/hs_forcing_column.h.
// hs_forcing.cu calls hs_column / hs_tracer_column / hs_radiative_surface from one CUDA thread per column; this file calls the very
// same functions in plain loops so that `pytest -m "not gpu"` can check the formulas against the NumPy oracle on a machine without a
// GPU.  Compiled by tests/test_hs_host.py with g++ into tests/host/_build/; never linked into the product library.
#include "../../isca_b200/csrc/hs_forcing_column.h"

using namespace isca_hs;

extern "C" {

// p: the HsParams fields in declaration order after (K, plane): 4 ints then 29 doubles
int hs_host_forcing(int K, long plane, const int* ip, const double* dp, double dt, double dec, const double* lat, const double* lon,
                    const double* coszen, const double* p_half, const double* p_full, const double* u, const double* v, const double* t,
                    const double* um, const double* vm, const double* zfull, double* udt, double* vdt, double* tdt, double* teq,
                    double* tg_prev, double* h_trop, int ntr, const double* rm, double* rdt) {
  HsParams p;
  p.K = K; p.plane = (size_t)plane;
  p.do_conserve_energy = ip[0]; p.eq_opt = ip[1]; p.strat_opt = ip[2]; p.local_heating = ip[3];
  double* d = &p.tka;
  for (int i = 0; i < 29; ++i) d[i] = dp[i];
  for (long c = 0; c < plane; ++c)
    hs_column(p, (size_t)c, dt, lat[c], lon[c], coszen ? coszen[c] : 0.0, dec, p_half, p_full, u, v, t, um, vm, zfull, udt, vdt, tdt, teq,
              tg_prev, h_trop);
  for (int n = 0; n < ntr; ++n)
    for (long c = 0; c < plane; ++c) hs_tracer_column(p, (size_t)c, dt, p_half, rm + (size_t)n * K * plane, rdt + (size_t)n * K * plane);
  return 0;
}

// the loop body of hs_spinup_kernel
int hs_host_spinup(long plane, const int* ip, const double* dp, int n_iter, const double* dec, const double* lat, double* tg_prev) {
  HsParams p;
  p.K = 1; p.plane = (size_t)plane;
  p.do_conserve_energy = ip[0]; p.eq_opt = ip[1]; p.strat_opt = ip[2]; p.local_heating = ip[3];
  double* d = &p.tka;
  for (int i = 0; i < 29; ++i) d[i] = dp[i];
  for (long c = 0; c < plane; ++c) {
    double tg = 250.0, prev = 250.0;
    for (int i = 0; i < n_iter; ++i) {
      prev = tg;
      double t_trop, h_trop, t_surf;
      hs_radiative_surface(p, lat[c], dec[i], t_trop, h_trop, t_surf);
      tg = hs_slab_update(p, 86400.0, t_surf, prev);
    }
    tg_prev[c] = prev;
  }
  return 0;
}

int hs_host_nparams() { return (int)((sizeof(HsParams) - offsetof(HsParams, tka)) / sizeof(double)); }

}  // extern "C"
