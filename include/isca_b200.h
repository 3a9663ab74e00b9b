/* isca_b200.h -- C ABI of the B200-native Isca spectral-dynamical-core hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer:
 * its "operator API" is the set of Fortran module procedures of atmosphere_mod,
 * spectral_dynamics_mod and transforms_mod.  Each entry point below names the reference
 * interface it replaces (paths relative to /root/reference/src).  A Fortran maintainer
 * binds these with ISO_C_BINDING (see INTEGRATION.md for the interface block and the
 * replacement atmosphere_mod shim).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is available
 *     from isca_b200_last_error().  No C++ exception crosses the boundary.  The Fortran
 *     shim turns a non-zero return into error_mesg(..., FATAL) like the reference.
 *   - arrays are in the reference's Fortran memory order:
 *       grid 3-D   (lon, lat, lev)  -> C index [lev][lat][lon], doubles
 *       grid 2-D   (lon, lat)
 *       spectral   (m, n, lev), m = 0..num_fourier, n = 0..num_spherical ("meridional
 *                  index", total wavenumber L = m + n), complex(8) stored as interleaved
 *                  (re, im) doubles -> C index [lev][n][m][2]
 *     latitudes run south to north (spectral_dynamics.F90:239), lat is the local block
 *     [lat_start, lat_start + lat_count) of this rank (grid decomposition of
 *     atmos_spectral/tools/spec_mpp.F90:61-65).
 *   - the library owns all device memory; the caller owns host buffers.  One host thread
 *     per handle, one CUDA device per rank.
 *   - there is no CPU fallback: if no CUDA device is present create() fails.
 */
#ifndef ISCA_B200_H
#define ISCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISCA_B200_ABI_VERSION 2

typedef struct IscaHandle_t* IscaHandle;

/* Every namelist value the hot path reads.  Field names and
 * defaults follow the reference namelists:
 *   spectral_dynamics_nml  atmos_spectral/model/spectral_dynamics.F90:152-224
 *   hs_forcing_nml         atmos_param/hs_forcing/hs_forcing.F90:74-122
 *   spectral_init_cond_nml atmos_spectral/init/spectral_init_cond.F90:68-74
 *   constants_nml          shared/constants/constants.F90:254-270
 *   main_nml dt_atmos      atmos_solo/atmos_model.F90:111
 * Use isca_b200_default_config() to obtain the reference defaults, then override. */
typedef struct IscaConfig {
  int32_t abi_version;            /* must be ISCA_B200_ABI_VERSION */
  /* resolution */
  int32_t lon_max, lat_max, num_fourier, num_spherical, num_levels;
  double  dt_atmos;
  /* spectral_dynamics_nml */
  int32_t damping_order, damping_order_vor, damping_order_div;
  double  damping_coeff, damping_coeff_vor, damping_coeff_div;
  double  eddy_sponge_coeff, zmu_sponge_coeff, zmv_sponge_coeff;
  int32_t do_mass_correction, do_energy_correction, do_water_correction;
  int32_t use_virtual_temperature, use_implicit;
  int32_t make_symmetric;         /* zonally symmetric model: every coefficient with m > 0 is removed wherever the triangular truncation
                                   * is applied (spherical.F90:185; the `axisymmetric` test case) */
  double  robert_coeff, raw_filter_coeff, alpha_implicit;
  int32_t vert_coord_option;      /* 0 even_sigma, 1 uneven_sigma, 2 input (pk/bk below), 3 hybrid (p_press, p_sigma) */
  double  scale_heights, surf_res, exponent, p_press, p_sigma;
  int32_t vert_advect_uv, vert_advect_t;  /* 0 second_centered (only value supported) */
  double  reference_sea_level_press, initial_sphum, water_correction_limit;
  double  valid_range_t[2];
  double  initial_temperature;
  /* tracers (field_table): 0, or 1 grid tracer "sphum" (finite_volume_parabolic) */
  int32_t num_tracers;
  double  tracer_robert_coeff;    /* < 0: use robert_coeff */
  /* hs_forcing_nml (atmosphere_nml idealized_moist_model = .false.) */
  int32_t no_forcing, do_conserve_energy;
  double  t_zero, t_strat, delh, delv, eps, sigma_b, P00, ka, ks, kf, trflux, trsink;
  /* constants_nml */
  double  radius, omega, grav, rdgas, kappa;
  /* vert_coordinate_nml (vert_coord_option == 2): num_levels+1 values each */
  const double* pk;
  const double* bk;
} IscaConfig;

/* field ids for get_field / get_spectral */
enum {
  ISCA_F_PS = 0, ISCA_F_U = 1, ISCA_F_V = 2, ISCA_F_T = 3, ISCA_F_VOR = 4, ISCA_F_DIV = 5,
  ISCA_F_WG_FULL = 6, ISCA_F_P_FULL = 7, ISCA_F_P_HALF = 8, ISCA_F_Z_FULL = 9, ISCA_F_Z_HALF = 10,
  ISCA_F_TRACER0 = 16,
  /* derived fields of spectral_diagnostics (spectral_dynamics.F90:1747-1835), formed on the device from the current level so that
   * products can be time-averaged there (isca_b200_diag_accumulate) instead of shipping their factors every step:
   * wspd = sqrt(u^2+v^2); ucomp_sq, vcomp_sq, ucomp_vcomp, vcomp_vor, temp_sq, omega_sq, omega_temp, ucomp_omega, vcomp_omega,
   * ucomp_temp, vcomp_temp, ucomp_height, vcomp_height, omega_height; sphum_u, sphum_v, sphum_w (tracer 0); slp [J][I] */
  ISCA_F_WSPD = 32, ISCA_F_UU = 33, ISCA_F_VV = 34, ISCA_F_UV = 35, ISCA_F_V_VOR = 36, ISCA_F_TT = 37, ISCA_F_OMEGA_OMEGA = 38,
  ISCA_F_OMEGA_T = 39, ISCA_F_UW = 40, ISCA_F_VW = 41, ISCA_F_UT = 42, ISCA_F_VT = 43, ISCA_F_UZ = 44, ISCA_F_VZ = 45,
  ISCA_F_OMEGA_Z = 46, ISCA_F_UTR0 = 48, ISCA_F_VTR0 = 49, ISCA_F_WTR0 = 50, ISCA_F_SLP = 56,
  ISCA_S_VOR = 0, ISCA_S_DIV = 1, ISCA_S_T = 2, ISCA_S_LNPS = 3
};
/* time-level selectors */
enum { ISCA_LEVEL_CURRENT = -1, ISCA_LEVEL_PREVIOUS = -2 };
/* scalars */
enum { ISCA_SC_MEAN_PS = 0, ISCA_SC_MEAN_ENERGY = 1, ISCA_SC_T_MIN = 2, ISCA_SC_T_MAX = 3,
       ISCA_SC_STEP_COUNT = 4, ISCA_SC_KERNEL_LAUNCHES = 5, ISCA_SC_LAST_STEP_MS = 6 };
/* 1-D tables */
enum { ISCA_TB_SIN_LAT = 0, ISCA_TB_WTS_LAT = 1, ISCA_TB_DEG_LAT = 2, ISCA_TB_DEG_LON = 3,
       ISCA_TB_PK = 4, ISCA_TB_BK = 5,
       /* isca_b200_host_table only.  Packed spectral rows: row r = (m, n) in the order ROW_M / ROW_N give (T rows) */
       ISCA_TB_ROW_M = 16, ISCA_TB_ROW_N = 17,          /* [T] as doubles */
       ISCA_TB_LEGENDRE = 18,                           /* [T][lat_max/2]: associated Legendre functions on the Gaussian latitudes of
                                                           one hemisphere (tools/spherical.F90 compute_legendre) */
       ISCA_TB_EIGEN_LAPLACIAN = 19,                    /* [T] n(n+1)/a^2 */
       ISCA_TB_DAMPING = 20,                            /* [T] spectral_damping coefficients (spectral_damping.F90) */
       ISCA_TB_REF_T = 21, ISCA_TB_IMPLICIT_H = 22,     /* [K] implicit.F90 reference temperature and h */
       ISCA_TB_DIV_MAT = 23,                            /* [K][K] implicit.F90 div_mat, row-major */
       ISCA_TB_WAVE_MATRIX = 24 };                      /* [M+1][K][K] implicit.F90 wave_matrix, see isca_b200_host_table */

/* Fill cfg with the reference's namelist defaults (T42 L18 dt=600, Held-Suarez). */
void isca_b200_default_config(IscaConfig* cfg);

/* atmosphere_init (atmos_spectral/driver/solo/atmosphere.F90:120-272) ->
 * spectral_dynamics_init (model/spectral_dynamics.F90:230-492): transforms_init,
 * press_and_geopot_init, spectral_damping_init, implicit_init, hs_forcing_init.
 * nccl_unique_id: 128-byte ncclUniqueId shared by all ranks (NULL when nranks == 1). */
int isca_b200_create(const IscaConfig* cfg, int rank, int nranks, const void* nccl_unique_id,
                     IscaHandle* out);
/* atmosphere_end (atmosphere.F90:356-399): releases all device memory. */
int isca_b200_destroy(IscaHandle h);
const char* isca_b200_last_error(IscaHandle h);   /* h may be NULL (create errors) */
/* bytes of an ncclUniqueId written to out (rank 0 calls this, the host runtime broadcasts) */
int isca_b200_nccl_unique_id(void* out128);

/* Peer-memory transpose (nranks > 1, one process per GPU of one NVLink/NVSwitch domain): every rank exports
 * its two Fourier buffers (2 x 64-byte cudaIpcMemHandle_t), the host runtime all-gathers the nranks x 128 bytes
 * (rank order) and hands them back; from then on the Legendre / FFT kernels store straight into the peers'
 * buffers and the per-transform exchange is only a one-element all-reduce used as a barrier.  Without this
 * call the transpose (tools/transforms.F90:970-1056 transpose_fourier) is a grouped ncclSend/ncclRecv. */
int isca_b200_ipc_handles(IscaHandle h, void* out128);
int isca_b200_set_peer_handles(IscaHandle h, const void* all_handles);

/* The decomposition the library uses for (rank, nranks): the rank's contiguous latitude block
 * (grid domain of tools/spec_mpp.F90:61-65) and its zonal wavenumbers (spectral domain; dealt in snake
 * order instead of spec_mpp.F90:77-80's contiguous blocks, to balance the triangle).  m_list has room for
 * num_fourier+1 entries; owner/pos have num_fourier+1 entries (pos = row of m in the lat-owner-side
 * Fourier buffer).  Needs no GPU. */
int isca_b200_decomposition(const IscaConfig* cfg, int rank, int nranks, int* lat_start, int* lat_count,
                            int* num_m, int* m_list, int* owner, int* pos);

/* The host-side tables the library builds at create time (Gaussian grid, vertical coordinate, Legendre functions, spectral
 * coefficient tables, semi-implicit matrices; host_tables.cpp), without creating a handle: needs no GPU, so the table code can be
 * checked against the oracle on any machine.  count must equal the table's size (queried with host == NULL: returns the size in
 * *count_out).  ISCA_TB_WAVE_MATRIX is built for xi = 2 * dt_atmos * alpha_implicit (the leapfrog step). */
int isca_b200_host_table(const IscaConfig* cfg, int table_id, double* host, int count, int* count_out);

/* cold start: spectral_init_cond 'quiescent' -> spectral_initialize_fields
 * (atmos_spectral/init/spectral_initialize_fields.F90:45-135), previous = current. */
int isca_b200_cold_start(IscaHandle h);

/* restart path: read_restart_or_do_coldstart (spectral_dynamics.F90:509-575) and
 * atmosphere_init (atmosphere.F90:197-223).  slot = 0 or 1 is the storage slot of the
 * reference's time-level dimension; tracers is (lon, lat, lev, ntr) or NULL. */
int isca_b200_set_grid_state(IscaHandle h, int slot, const double* ug, const double* vg,
                             const double* tg, const double* psg, const double* tracers);
int isca_b200_set_spectral_state(IscaHandle h, int slot, const double* vors, const double* divs,
                                 const double* ts, const double* ln_ps);
int isca_b200_set_vor_div_grid(IscaHandle h, const double* vorg, const double* divg);
int isca_b200_set_surf_geopotential(IscaHandle h, const double* surf_geopotential);
int isca_b200_set_time_pointers(IscaHandle h, int previous_slot, int current_slot);

/* atmosphere(Time) (atmosphere.F90:276-352) called n_steps times: physics
 * (hs_forcing.F90:148-272) + spectral_dynamics (spectral_dynamics.F90:780-1034) +
 * time-level swap.  State stays on the device; nothing is copied to the host. */
int isca_b200_step(IscaHandle h, int n_steps);
/* same as step(), but physics is skipped (dt_* = 0): spectral_dynamics with zero
 * tendencies, the "transform + semi-implicit only" configuration of BASELINE.json. */
int isca_b200_step_dynamics_only(IscaHandle h, int n_steps);

/* spectral_dynamics(...) at the reference's own argument list (spectral_dynamics.F90:780):
 * host tendencies in (dt_ug, dt_vg, dt_tg: (lon,lat,lev); dt_psg: (lon,lat); any may be NULL
 * = zero), new state out (psg_final, ug_final, vg_final, tg_final, wg_full, p_full; any may
 * be NULL).  One time step; host<->device copies happen inside the call. */
int isca_b200_spectral_dynamics(IscaHandle h, const double* dt_psg, const double* dt_ug,
                                const double* dt_vg, const double* dt_tg,
                                double* psg_final, double* ug_final, double* vg_final,
                                double* tg_final, double* wg_full, double* p_full);

/* The same with the reference's tracer arguments (spectral_dynamics.F90:780-783: dt_tracers in, grid_tracers_final out;
 * one grid tracer, sphum): dt_tracers / grid_tracers_final are (lon,lat,lev), either may be NULL. */
int isca_b200_spectral_dynamics_tracers(IscaHandle h, const double* dt_psg, const double* dt_ug,
                                        const double* dt_vg, const double* dt_tg, const double* dt_tracers,
                                        double* psg_final, double* ug_final, double* vg_final,
                                        double* tg_final, double* grid_tracers_final, double* wg_full,
                                        double* p_full);

/* lazy host mirrors for diag_manager send_data / restart writes
 * (spectral_dynamics.F90:1502-1531,1709-1867).  level: ISCA_LEVEL_CURRENT / _PREVIOUS or slot */
int isca_b200_get_field(IscaHandle h, int field_id, int level, double* host);
int isca_b200_get_spectral(IscaHandle h, int field_id, int level, double* host);
int isca_b200_get_scalar(IscaHandle h, int scalar_id, double* value);
int isca_b200_get_table(IscaHandle h, int table_id, double* host, int count);
int isca_b200_get_time_pointers(IscaHandle h, int* previous_slot, int* current_slot);

/* transforms_mod level entry points (atmos_spectral/tools/transforms.F90:379-533,700-783):
 * trans_spherical_to_grid, trans_grid_to_spherical(do_truncation), uv_grid_from_vor_div,
 * vor_div_from_uv_grid.  Host arrays, nlev levels. */
int isca_b200_spherical_to_grid(IscaHandle h, const double* spec, double* grid, int nlev);
int isca_b200_grid_to_spherical(IscaHandle h, const double* grid, double* spec, int nlev,
                                int do_truncation);
int isca_b200_uv_grid_from_vor_div(IscaHandle h, const double* vors, const double* divs,
                                   double* ug, double* vg, int nlev);
int isca_b200_vor_div_from_uv_grid(IscaHandle h, const double* ug, const double* vg,
                                   double* vors, double* divs, int nlev);

/* Stage-level entry points (unit tests / init-time callers; single rank -- with nranks > 1 the Fourier transpose sits between
 * the two stages).  Fourier arrays are the reference's (0:num_fourier, lat, lev) complex arrays:
 *   isca_b200_fft_r2c      trans_grid_to_fourier    (tools/grid_fourier.F90:129-150 -> fft_grid_to_fourier, shared/fft/fft.F90:483-591)
 *   isca_b200_fft_c2r      trans_fourier_to_grid    (tools/grid_fourier.F90:154-179 -> fft_fourier_to_grid, fft.F90:600-718)
 *   isca_b200_legendre_inv trans_spherical_to_fourier (tools/spherical_fourier.F90:177-261)
 *   isca_b200_legendre_fwd trans_fourier_to_spherical (tools/spherical_fourier.F90:264-339), + triangular_truncation if asked */
int isca_b200_fft_r2c(IscaHandle h, const double* grid, double* fourier, int nlev);
int isca_b200_fft_c2r(IscaHandle h, const double* fourier, double* grid, int nlev);
int isca_b200_legendre_inv(IscaHandle h, const double* spec, double* fourier, int nlev);
int isca_b200_legendre_fwd(IscaHandle h, const double* fourier, double* spec, int nlev,
                           int do_truncation);
/* implicit_correction(dt_divs, dt_ts, dt_ln_ps, divs, ts, ln_ps, delta_t, previous, current)
 * (atmos_spectral/model/implicit.F90:241-325): the three tendencies are updated in place; the two time levels of divs, ts,
 * ln_ps are passed as separate (m,n,lev) / (m,n) complex arrays; wave matrices are (re)built for delta_t as :260-264 does. */
int isca_b200_implicit_correction(IscaHandle h, double* dt_divs, double* dt_ts, double* dt_ln_ps,
                                  const double* divs_prev, const double* divs_cur,
                                  const double* ts_prev, const double* ts_cur,
                                  const double* ln_ps_prev, const double* ln_ps_cur, double delta_t);

/* Device-side time averaging for diag_manager (send_data with time_avg, diag_manager; spectral_diagnostics
 * spectral_dynamics.F90:1709-1867): accumulate adds the current level's field (ISCA_F_* ids) to a device accumulator; fetch
 * returns the mean over the accumulated samples (and their number), optionally resetting the accumulator. */
int isca_b200_diag_accumulate(IscaHandle h, int field_id);
int isca_b200_diag_fetch(IscaHandle h, int field_id, double* host, int reset, int* count_out);

/* kernel-level entry points used by bench.py / tests for device-resident timing:
 * run `reps` batched inverse+forward transform pairs of `nlev` levels on resident synthetic
 * data and return the average milliseconds of each stage measured with CUDA events on the
 * library's stream. stages: 0 legendre_inv, 1 fft_inv, 2 fft_fwd, 3 legendre_fwd. */
int isca_b200_time_transforms(IscaHandle h, int nlev, int reps, double ms_out[4]);
/* per-kernel-group CUDA-event timings of the last profiled step (ms); names are written as
 * a ';'-separated list into names (capacity bytes). Returns the number of groups. */
int isca_b200_profile_step(IscaHandle h, int n_steps, double* ms_out, int max_groups,
                           char* names, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* ISCA_B200_H */
