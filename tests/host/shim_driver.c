/* TEST INFRASTRUCTURE -- the call sequence of fortran/atmosphere.F90 (atmosphere_init / atmosphere / atmosphere_end on a cold start,
 * then a restart) made from plain C through the C ABI, with every array laid out as the Fortran shim holds it: grid fields
 * (lon, lat, lev) with the longitude fastest, spectral fields complex (m, n, lev) interleaved (re, im).  No Fortran compiler exists in
 * this image; this driver is what exercises the boundary the shim binds (tests/test_fortran_shim.py builds it with gcc against
 * include/ and isca_b200/lib/libisca_b200.so).
 *
 *   shim_driver <nsteps_a> <nsteps_b>    prints "CHECKSUM <name> <value>" lines; exit 0 = ok, 3 = no CUDA device, other = failure */
#include "../../include/isca_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double checksum(const double* a, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += a[i] * (double)(1 + i % 7);
  return s;
}
#define CK(call) do { if ((call) != 0) { fprintf(stderr, "FATAL %s: %s\n", #call, isca_b200_last_error(h)); return 2; } } while (0)

int main(int argc, char** argv) {
  const int na = argc > 1 ? atoi(argv[1]) : 6, nb = argc > 2 ? atoi(argv[2]) : 4;
  IscaConfig cfg;
  isca_b200_default_config(&cfg);
  /* held_suarez_test_case.py namelists at T21 L10, as atmosphere_init forwards them */
  cfg.lon_max = 64; cfg.lat_max = 32; cfg.num_fourier = 21; cfg.num_spherical = 22; cfg.num_levels = 10;
  cfg.dt_atmos = 1200.0;
  cfg.damping_order = 4; cfg.water_correction_limit = 200.e2; cfg.reference_sea_level_press = 1.0e5;
  cfg.valid_range_t[0] = 100.0; cfg.valid_range_t[1] = 800.0; cfg.initial_sphum = 0.0;
  cfg.vert_coord_option = 1; cfg.scale_heights = 6.0; cfg.exponent = 7.5; cfg.surf_res = 0.5;
  cfg.num_tracers = 1; cfg.do_water_correction = 1;
  IscaHandle h = NULL;
  if (isca_b200_create(&cfg, 0, 1, NULL, &h) != 0) {
    const char* e = isca_b200_last_error(NULL);
    fprintf(stderr, "create: %s\n", e);
    return strstr(e, "CUDA") || strstr(e, "device") ? 3 : 2;
  }
  const size_t I = cfg.lon_max, J = cfg.lat_max, K = cfg.num_levels, n3 = I * J * K, n2 = I * J;
  const size_t ns3 = (size_t)(cfg.num_fourier + 1) * (cfg.num_spherical + 1) * K * 2, ns2 = (size_t)(cfg.num_fourier + 1) * (cfg.num_spherical + 1) * 2;
  CK(isca_b200_cold_start(h));
  CK(isca_b200_step(h, na));                                  /* atmosphere(Time) x na */
  /* atmosphere_end: both time levels through the host mirrors, as the restart files hold them */
  double *ug[2], *vg[2], *tg[2], *ps[2], *q[2], *vors[2], *divs[2], *ts[2], *lnps[2];
  for (int nt = 0; nt < 2; ++nt) {
    ug[nt] = malloc(n3 * 8); vg[nt] = malloc(n3 * 8); tg[nt] = malloc(n3 * 8); q[nt] = malloc(n3 * 8); ps[nt] = malloc(n2 * 8);
    vors[nt] = malloc(ns3 * 8); divs[nt] = malloc(ns3 * 8); ts[nt] = malloc(ns3 * 8); lnps[nt] = malloc(ns2 * 8);
    CK(isca_b200_get_field(h, ISCA_F_U, nt, ug[nt])); CK(isca_b200_get_field(h, ISCA_F_V, nt, vg[nt]));
    CK(isca_b200_get_field(h, ISCA_F_T, nt, tg[nt])); CK(isca_b200_get_field(h, ISCA_F_PS, nt, ps[nt]));
    CK(isca_b200_get_field(h, ISCA_F_TRACER0, nt, q[nt]));
    CK(isca_b200_get_spectral(h, ISCA_S_VOR, nt, vors[nt])); CK(isca_b200_get_spectral(h, ISCA_S_DIV, nt, divs[nt]));
    CK(isca_b200_get_spectral(h, ISCA_S_T, nt, ts[nt])); CK(isca_b200_get_spectral(h, ISCA_S_LNPS, nt, lnps[nt]));
  }
  double* vorg = malloc(n3 * 8); double* divg = malloc(n3 * 8);
  CK(isca_b200_get_field(h, ISCA_F_VOR, ISCA_LEVEL_CURRENT, vorg)); CK(isca_b200_get_field(h, ISCA_F_DIV, ISCA_LEVEL_CURRENT, divg));
  int previous = 0, current = 0;
  CK(isca_b200_get_time_pointers(h, &previous, &current));
  /* uninterrupted continuation = the reference answer of the restart */
  CK(isca_b200_step(h, nb));
  double* t_ref = malloc(n3 * 8); double* ps_ref = malloc(n2 * 8);
  CK(isca_b200_get_field(h, ISCA_F_T, ISCA_LEVEL_CURRENT, t_ref)); CK(isca_b200_get_field(h, ISCA_F_PS, ISCA_LEVEL_CURRENT, ps_ref));
  printf("CHECKSUM T_uninterrupted %.17g\n", checksum(t_ref, n3));
  printf("CHECKSUM ps_uninterrupted %.17g\n", checksum(ps_ref, n2));
  CK(isca_b200_destroy(h));
  /* atmosphere_init from the "restart files" */
  h = NULL;
  if (isca_b200_create(&cfg, 0, 1, NULL, &h) != 0) { fprintf(stderr, "create(2): %s\n", isca_b200_last_error(NULL)); return 2; }
  for (int nt = 0; nt < 2; ++nt) {
    CK(isca_b200_set_grid_state(h, nt, ug[nt], vg[nt], tg[nt], ps[nt], q[nt]));
    CK(isca_b200_set_spectral_state(h, nt, vors[nt], divs[nt], ts[nt], lnps[nt]));
  }
  CK(isca_b200_set_vor_div_grid(h, vorg, divg));
  CK(isca_b200_set_time_pointers(h, previous, current));
  CK(isca_b200_step(h, nb));
  double* t_new = malloc(n3 * 8); double* ps_new = malloc(n2 * 8);
  CK(isca_b200_get_field(h, ISCA_F_T, ISCA_LEVEL_CURRENT, t_new)); CK(isca_b200_get_field(h, ISCA_F_PS, ISCA_LEVEL_CURRENT, ps_new));
  printf("CHECKSUM T_restarted %.17g\n", checksum(t_new, n3));
  printf("CHECKSUM ps_restarted %.17g\n", checksum(ps_new, n2));
  double dmax = 0.0, tmax = 0.0;
  for (size_t i = 0; i < n3; ++i) { double d = t_new[i] - t_ref[i]; if (d < 0) d = -d; if (d > dmax) dmax = d; if (t_ref[i] > tmax) tmax = t_ref[i]; }
  printf("RESTART_REL_DIFF_T %.3e\n", dmax / tmax);
  double mean_ps = 0.0;
  CK(isca_b200_get_scalar(h, ISCA_SC_MEAN_PS, &mean_ps));
  printf("MEAN_PS %.17g\n", mean_ps);
  CK(isca_b200_destroy(h));
  return dmax / tmax < 1e-12 ? 0 : 4;
}
