// physics_common.h -- shared definitions of the column-physics kernels (physics*.cu): the handle, device buffers,
// saturation-vapour-pressure lookup and the host<->device staging helpers of the host-array C ABI.
#pragma once
#include "../../include/isca_b200_physics.h"
#include "common.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// Resident CTAs per SM asked of the 128-thread column kernels.  At T170 (131 072 columns = 1024 CTAs = 6.9 per SM) a latency-bound
// kernel that keeps fewer than 7 CTAs resident runs a second, mostly empty wave and pays for two: (128, 7) caps the registers at 72,
// which costs these kernels a few spilled values (ptxas -v; surface_flux 276 bytes, still faster: 0.058 -> 0.046 ms).  vert_diff_down
// (125 registers) is fastest at 4: 0.29 ms against 0.31 (5) and 0.42 (7).
#ifndef ISCA_COL_MINB
#define ISCA_COL_MINB 7
#endif
#ifndef ISCA_VDD_MINB
#define ISCA_VDD_MINB 4
#endif
#ifndef ISCA_SF_MINB
#define ISCA_SF_MINB 7
#endif

namespace isca_phys {

struct SvpDev {
  const double *tab, *dtab, *d2tab;
  double tminl, dtinvl, tepsl, dtres;
  int n;
};

struct PhysConst {
  double grav, rdgas, rvgas, cp_air, hlv, stefan, pstd;
  double hc; int do_evap;
  double solar_constant, del_sol, del_sw, ir_tau_eq, ir_tau_pole, atm_abs, sw_diff, linear_tau, wv_exponent,
         solar_exponent, odp, diabatic_acce;
  // rad_scheme variants (two_stream_gray_rad.F90:89-118): 0 frierson, 1 byrne, 2 geen, 3 schneider
  int rad_scheme;
  double ir_tau_co2_win, ir_tau_wv_win1, ir_tau_wv_win2, ir_tau_co2, ir_tau_wv1, ir_tau_wv2, window, carbon_conc;
  double carbon_conc_sw;                // do_read_co2: value seen by the geen shortwave (the longwave value of the previous call)
  double lw_tau_0_gp, sw_tau_0_gp, lw_tau_exponent_gp, sw_tau_exponent_gp, gp_albedo, Ga_asym;
  double bog_a, bog_b, bog_mu, pstd_earth;
  // do_seasonal (two_stream_gray_rad.F90:417-447): per-column insolation = solar_constant * coszen computed by the driver;
  // nullptr = the analytic annual-mean profiles
  const double* insol_dev = nullptr;
};

// lookup_es_des (sat_vapor_pres_k.F90:1132-1158): table index + 2nd-order Taylor; false = outside the table
__device__ __forceinline__ bool svp_lookup(const SvpDev& s, double T, double& es, double& des) {
  double tmp = T - s.tminl;
  double x = s.dtinvl * (tmp + s.tepsl);
  if (!(x > -1.0 && x < (double)s.n)) { es = 0.0; des = 0.0; return false; }
  int ind = (int)x;                                   // truncation, like the Fortran int()
  double dl = tmp - s.dtres * (double)ind;
  double t0 = __ldg(s.tab + ind), t1 = __ldg(s.dtab + ind), t2 = __ldg(s.d2tab + ind);
  es = t0 + dl * (t1 + dl * t2);
  des = t1 + 2.0 * dl * t2;
  return true;
}

// compute_qs_k (sat_vapor_pres_k.F90:457-540, q absent)
__device__ __forceinline__ void qs_from_es(double es, double des, double press, double hc, double eps, double& qs, double& dqs) {
  des *= hc; es *= hc;
  double denom = press - (1.0 - eps) * es;
  qs = denom > 0.0 ? eps * es / denom : eps;
  dqs = eps * press * des / (denom * denom);
}

struct MoConst { double rich_crit, drag_min, zeta_trans, vonkarm, grav; int neutral, stable_option; };

struct Dev {                 // owning device array
  double* p = nullptr; size_t n = 0;
  bool ensure(size_t count) {
    if (count <= n) return true;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    if (cudaMalloc(&p, count * sizeof(double)) != cudaSuccess) return false;
    n = count; return true;
  }
  ~Dev() { if (p) cudaFree(p); }
};

// module state kept between calls, as the Fortran modules keep it (vert_diff_mod e_global/f_t_global/f_q_global and the
// surf_diff_type Tri_surf of idealized_moist_phys; mixed_layer_mod arrays)
enum StateId { ST_E_GLOBAL = 0, ST_F_T_GLOBAL, ST_F_Q_GLOBAL, ST_TRI_DELTA_T, ST_TRI_DFLUX_T, ST_TRI_DELTA_Q, ST_TRI_DFLUX_Q,
               ST_TRI_DTMASS, ST_TRI_DELTA_U, ST_TRI_DELTA_V, ST_ML_HEAT_CAP, ST_ML_QFLUX, ST_ML_SST, ST_COUNT };

}  // namespace isca_phys

struct IscaPhysics_t {
  IscaPhysicsConfig cfg;
  isca_phys::PhysConst pc;
  isca_phys::SvpDev svp;
  isca_phys::Dev tab;                   // TABLE | DTABLE | D2TABLE
  std::vector<double> svp_host;         // host copy of the same tables (LCL table construction)
  isca_phys::Dev lcl_tab;               // qe_moist_convection lcl_temp_table
  int lcl_n = 0; double lcl_val_min = 0.0, lcl_val_max = 0.0;
  std::string lcl_err;                  // qe_moist_convection_init failure, raised when the scheme is used (the reference calls that init only for SIMPLE_BETTS_MILLER)
  isca_phys::Dev buf[24];               // staging of the host-array entry points
  isca_phys::Dev state[isca_phys::ST_COUNT];
  IscaBettsMillerConfig bm{1, 1, 0, 0, 0, 0, 7200., .8, 900., 2400., 0.};      // betts_miller_nml defaults (betts_miller.f90:56-66)
  double co2_next_sw = 360.0;           // do_read_co2 bookkeeping of launch_gray_down
  isca_phys::Dev insol;                 // do_seasonal insolation [J][I] (pc.insol_dev points here while it is set)
  bool vert_diff_down_done = false;
  bool sc_sst = false;                  // mixed_layer_nml do_sc_sst: t_surf follows state[ST_ML_SST] (isca_b200_mixed_layer_set_sst)
  int* d_err = nullptr;
  cudaStream_t st = nullptr;
  bool owns_stream = true;              // false when the moist-model driver runs the kernels on the dynamical core's stream
  std::string err;
  size_t ncol = 0; int K = 0;
};

namespace isca_phys {

std::string& thread_error();
inline int fail(IscaPhysics p, const std::string& m) { if (p) p->err = m; thread_error() = m; return 1; }

#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return isca_phys::fail(p, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

inline int up(IscaPhysics p, Dev& d, const double* h, size_t n) {
  if (!h) return fail(p, "null input array");
  if (!d.ensure(n)) return fail(p, "cudaMalloc failed");
  PCK(cudaMemcpyAsync(d.p, h, n * sizeof(double), cudaMemcpyHostToDevice, p->st));
  return 0;
}
inline int down(IscaPhysics p, const Dev& d, double* h, size_t n) {
  if (!h) return fail(p, "null output array");
  PCK(cudaMemcpyAsync(h, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  return 0;
}
// waits for the stream and turns the device error flag (saturation-table overflow) into a failure
inline int finish(IscaPhysics p, const char* what) {
  int e = 0;
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaStreamSynchronize(p->st));
  if (e) {
    PCK(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->st));
    return fail(p, std::string(what) + ": lookup_es: temperature outside the saturation vapour pressure table (table overflow)");
  }
  return 0;
}
inline int col_blocks(IscaPhysics p, int threads) { return (int)((p->ncol + threads - 1) / threads); }

// device-pointer launches shared between files
void launch_lscale(IscaPhysics p, const double* t, const double* q, const double* pf, const double* ph, double* rain, double* td, double* qd);
void launch_gray_down(IscaPhysics p, const double* lat, const double* ph, const double* t, const double* q, const double* alb, double* sw, double* lw);
void launch_gray_up(IscaPhysics p, const double* lat, const double* ph, const double* t, const double* q, const double* ts, const double* alb, double* tdt, double* olr);
int rayleigh_nlev(const double* pref, int K, double pb);
void launch_rayleigh(IscaPhysics p, int nlev, double delt, const double* pf, const double* u, const double* v, double* udt, double* vdt, double* tdt,
                     int full = 1);
int prepare_vert_diff_state(IscaPhysics p);
void launch_vert_diff_down(IscaPhysics p, double delt, const double* u, const double* v, const double* t, const double* q,
                           const double* diff_m, const double* diff_t, const double* p_half, const double* z_full, double* tau_u,
                           double* tau_v, const double* dtau_du, const double* dtau_dv, double* dt_u, double* dt_v, double* dt_t,
                           const double* dt_q, double* diss);
void launch_vert_diff_up(IscaPhysics p, double delt, double* dt_t, double* dt_q);
void launch_mixed_layer(IscaPhysics p, double dt, double* t_surf, const double* flux_t, const double* flux_q, const double* flux_r,
                        const double* net_surf_sw_down, const double* surf_lw_down, const double* dhdt_surf, const double* dedt_surf,
                        const double* dedq_surf, const double* drdt_surf, const double* dhdt_atm, const double* dedq_atm, double* delta_t_surf);
void launch_surface_flux(IscaPhysics p, const IscaSurfaceFluxArgs& dev);     // physics_surface.cu; device pointers
int build_lcl_table(IscaPhysics p);                                             // physics_conv.cu
void launch_sbm_convection(IscaPhysics p, double dt, const double* Tin, const double* qin, const double* p_full, const double* p_half,
                           double* rain, double* deltaT, double* deltaq, double* qref, double* Tref, int* convflag, int* kLZBs, int* kLCLs,
                           double* cape, double* cin, double* itq, double* itt);
void launch_betts_miller(IscaPhysics p, double dt, const double* tin, const double* qin, const double* p_full, const double* p_half,
                         double* rain, double* tdel, double* qdel, double* q_ref, double* t_ref, int* bmflag, int* klzbs, int* klcls,
                         double* cape, double* cin, double* invtau_t, double* invtau_q);                         // physics_bm.cu
void launch_dry_convection(IscaPhysics p, double tau, double gamma, const double* tg, const double* p_full, const double* p_half, double* tp,
                           double* dt_tg, double* cape, double* cin, int* lzb, int* lcl);                         // physics_dry.cu
void launch_diffusivity(IscaPhysics p, const double* t, const double* q, const double* u, const double* v, const double* z_full,
                        const double* z_half, const double* u_star, const double* b_star, double* h, double* k_m, double* k_t,
                        const double* tdt = nullptr, const double* qdt = nullptr, const double* udt = nullptr, const double* vdt = nullptr,
                        double dt = 0.0, int add_input = 1);

}  // namespace isca_phys
