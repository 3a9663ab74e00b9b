// moist_model.cu -- idealized_moist_model driver: the sequence of idealized_moist_phys
// (atmos_spectral/driver/solo/idealized_moist_phys.F90:819-1395) on the device-resident state of the dynamical core, followed
// by spectral_dynamics with the physics tendencies (atmosphere.F90:276-352).  Every kernel runs on the core's stream.
#include "physics_common.h"
#include "core_internal.h"
#include "rrtm_internal.h"

using namespace isca_phys;

struct IscaMoist_t {
  IscaMoistConfig mc;
  IscaHandle dyn = nullptr;
  IscaPhysics phy = nullptr;
  std::string err;
  int I = 0, J = 0, K = 0;
  size_t nc = 0, n3 = 0;
  bool initialized = false;
  double ms_step = 0.0, ms_phys = 0.0;
  std::vector<double> pref;
  // 3-D work [K][J][I] (p_half, z_half: K+1)
  Dev p_full[2], p_half[2], z_full[2], z_half[2];            // index 0: previous, 1: current (recomputed every step)
  Dev dt_u, dt_v, dt_t, dt_q, tg_tmp, qg_tmp, c_dT, c_dq, c_qref, c_Tref, diff_m, diff_t, diss, w1, w2, w3;
  // 2-D
  Dev lat2d, z_surf, t_surf, q_surf, gust, albedo, zero2, rough_m, rough_h, rough_q, z_atm, precip, rain, conv_rain, cape, cin, itq, itt,
      net_sw, lw_down, z_pbl, dts, sf;                         // sf: 28 surface-flux output planes
  Dev iwork;                                                   // int planes: land | convflag | kLZB | kLCL
  // do_rrtm_radiation (idealized_moist_phys.F90:1167-1177): RRTMG instead of two_stream_gray_rad
  IscaRrtm rr = nullptr;
  IscaRrtmDriverConfig rdc{};
  std::vector<double> orb_angle;
  Dev tdt_rad, coszen, lon2d, o3, olr, toa_sw;
  bool have_o3 = false;
  double time_s = 0.0;                                         // Time of atmosphere(Time), seconds since Time_init
  double dt_last = 0.0;                                        // rrtm_vars dt_last (radiation alarm)
  long n_rad_calls = 0;
  // two_stream_gray_rad_nml do_seasonal (two_stream_gray_rad.F90:417-447): insolation = solar_constant * coszen(Time) every step
  bool seasonal = false;
  IscaRrtmDriverConfig sdc{};
  std::vector<double> s_orb;
  double dry_tau = 0.0, dry_gamma = 0.0;                       // dry_convection_nml (convection_scheme = 'DRY')
  // isca_b200_moist_step_io: software-pipelined host I/O of a step (copy streams, staging buffer, events)
  cudaStream_t io_h2d = nullptr, io_d2h = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_snap = nullptr, ev_d2h = nullptr, ev_d2h_prev = nullptr, ev_main = nullptr;
  Dev io_stage;
  bool io_ready = false, io_pending = false;
};

namespace {

thread_local std::string g_merr;
int mfail(IscaMoist m, const std::string& s) { if (m) m->err = s; g_merr = s; return 1; }
#define MCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return mfail(m, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

__global__ void fill_kernel(double* a, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void lat2d_kernel(double* lat2d, const double* rad_lat, int I, int J) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < I) lat2d[(size_t)j * I + i] = rad_lat[j];
}
// idealized_moist_phys.F90:873-880: tg_tmp = conv_dt_tg + tg(previous); rates = increments / delta_t; dt += rates; precip = rain/delta_t
__global__ void conv_post_kernel(size_t n3, size_t nc, double delta_t, const double* __restrict__ dT, const double* __restrict__ dq,
                                 const double* __restrict__ tg, const double* __restrict__ qg, double* __restrict__ tg_tmp,
                                 double* __restrict__ qg_tmp, double* __restrict__ dt_t, double* __restrict__ dt_q,
                                 const double* __restrict__ rain, double* __restrict__ conv_rain, double* __restrict__ precip) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) {
    double a = dT[i], b = dq[i];
    tg_tmp[i] = a + tg[i]; qg_tmp[i] = b + qg[i];
    dt_t[i] = dt_t[i] + a / delta_t; dt_q[i] = dt_q[i] + b / delta_t;
    if (i < nc) { double r = rain[i] / delta_t; conv_rain[i] = r; precip[i] = r; }
  }
}
// :981-1000: condensation increments -> rates, precip += rain/delta_t
__global__ void cond_post_kernel(size_t n3, size_t nc, double delta_t, const double* __restrict__ dT, const double* __restrict__ dq,
                                 double* __restrict__ dt_t, double* __restrict__ dt_q, const double* __restrict__ rain,
                                 double* __restrict__ precip) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) {
    dt_t[i] = dt_t[i] + dT[i] / delta_t; dt_q[i] = dt_q[i] + dq[i] / delta_t;
    if (i < nc) precip[i] = precip[i] + rain[i] / delta_t;
  }
}
__global__ void add3_kernel(size_t n, double* a, const double* da, double* b, const double* db, double* c, const double* dc) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    a[i] = a[i] + da[i]; b[i] = b[i] + db[i]; c[i] = c[i] + dc[i];
  }
}
// vert_turb_driver.F90:209-213 (use_tau = .false.): variables at time tau+1
__global__ void tau_plus1_kernel(size_t n, double dt, const double* __restrict__ um, const double* __restrict__ vm, const double* __restrict__ tm,
                                 const double* __restrict__ qm, const double* __restrict__ udt, const double* __restrict__ vdt,
                                 const double* __restrict__ tdt, const double* __restrict__ qdt, double* __restrict__ uu, double* __restrict__ vv,
                                 double* __restrict__ tt, double* __restrict__ qq) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uu[i] = um[i] + dt * udt[i]; vv[i] = vm[i] + dt * vdt[i]; tt[i] = tm[i] + dt * tdt[i]; qq[i] = qm[i] + dt * qdt[i];
  }
}
__global__ void mask_to_int_kernel(size_t n, const double* __restrict__ a, int* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] != 0.0 ? 1 : 0;
}
__global__ void add1_kernel(size_t n, double* a, const double* da) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = a[i] + da[i];
}
__global__ void lon2d_kernel(double* lon2d, int I, int J) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < I) lon2d[(size_t)j * I + i] = (i * 360.0 / I) * (3.14159265358979323846 / 180.0);      // rad_lon = deg_lon * pi/180
}
__global__ void scale_kernel(size_t n, double* out, const double* a, double c) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = c * a[i];
}
__global__ void sub_kernel(size_t n, double* out, const double* a, const double* b) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] - b[i];
}
__global__ void add_const_kernel(size_t n, double* out, const double* a, double c) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] + c;
}

inline int nblk(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : b); }

const char* SF_NAMES[28] = {"flux_t", "flux_q", "flux_r", "flux_u", "flux_v", "cd_m", "cd_t", "cd_q", "w_atm", "u_star", "b_star", "q_star",
                            "dhdt_surf", "dedt_surf", "dedq_surf", "drdt_surf", "dhdt_atm", "dedq_atm", "dtaudu_atm", "dtaudv_atm",
                            "ex_del_m", "ex_del_h", "ex_del_q", "temp_2m", "u_10m", "v_10m", "q_2m", "rh_2m"};
enum { SF_FLUX_T = 0, SF_FLUX_Q, SF_FLUX_R, SF_FLUX_U, SF_FLUX_V, SF_CD_M, SF_CD_T, SF_CD_Q, SF_W_ATM, SF_U_STAR, SF_B_STAR, SF_Q_STAR,
       SF_DHDT_SURF, SF_DEDT_SURF, SF_DEDQ_SURF, SF_DRDT_SURF, SF_DHDT_ATM, SF_DEDQ_ATM, SF_DTAUDU, SF_DTAUDV };

// one call of idealized_moist_phys + spectral_dynamics
int moist_step_once(IscaMoist m, cudaEvent_t ev_phys_end) {
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, std::string("dynamical core: ") + isca_b200_last_error(m->dyn));
  IscaPhysics p = m->phy;
  cudaStream_t st = v.st;
  const int K = m->K; const size_t nc = m->nc, n3 = m->n3;
  const int prev = v.previous, cur = v.current;
  const double delta_t = (prev == cur) ? v.dt_atmos : 2 * v.dt_atmos;            // idealized_moist_phys.F90:838-842
  // pressures and heights of both time levels (atmosphere.F90:228-247, 332-339)
  if (isca_core_press_heights(m->dyn, prev, m->p_full[0].p, m->p_half[0].p, m->z_full[0].p, m->z_half[0].p) ||
      isca_core_press_heights(m->dyn, cur, m->p_full[1].p, m->p_half[1].p, m->z_full[1].p, m->z_half[1].p))
    return mfail(m, std::string("dynamical core: ") + isca_b200_last_error(m->dyn));
  isca_core_mark(m->dyn, "phys_press_heights");
  const double *tg_p = v.T[prev], *q_p = v.q[prev], *ug_p = v.u[prev], *vg_p = v.v[prev];
  const double *pf_p = m->p_full[0].p, *ph_p = m->p_half[0].p, *pf_c = m->p_full[1].p, *ph_c = m->p_half[1].p;
  const double *zf_c = m->z_full[1].p, *zh_c = m->z_half[1].p;
  MCK(cudaMemsetAsync(m->dt_u.p, 0, n3 * sizeof(double), st));
  MCK(cudaMemsetAsync(m->dt_v.p, 0, n3 * sizeof(double), st));
  MCK(cudaMemsetAsync(m->dt_t.p, 0, n3 * sizeof(double), st));
  MCK(cudaMemsetAsync(m->dt_q.p, 0, n3 * sizeof(double), st));
  int* land = reinterpret_cast<int*>(m->iwork.p);
  int* convflag = land + nc; int* klzb = convflag + nc; int* klcl = klzb + nc;
  const double *t_in = tg_p, *q_in = q_p;
  if (m->mc.convection_scheme == 1 || m->mc.convection_scheme == 3) {
    // (fusing this post-processing into the convection kernel was measured and dropped: the latency-bound column kernel runs the extra
    // level loop at a fraction of the bandwidth the streaming kernel reaches -- +0.08 ms at T170, profiles/r02/r02_experiments.md)
    if (m->mc.convection_scheme == 1)
      launch_sbm_convection(p, delta_t, tg_p, q_p, pf_p, ph_p, m->rain.p, m->c_dT.p, m->c_dq.p, m->c_qref.p, m->c_Tref.p, convflag, klzb, klcl,
                            m->cape.p, m->cin.p, m->itq.p, m->itt.p);
    else                                                          // 'FULL_BETTS_MILLER' (:889-916): same post-processing
      launch_betts_miller(p, delta_t, tg_p, q_p, pf_p, ph_p, m->rain.p, m->c_dT.p, m->c_dq.p, m->c_qref.p, m->c_Tref.p, convflag, klzb, klcl,
                          m->cape.p, m->cin.p, m->itt.p, m->itq.p);
    isca_core_mark(m->dyn, "phys_convection");
    conv_post_kernel<<<nblk(n3), 256, 0, st>>>(n3, nc, delta_t, m->c_dT.p, m->c_dq.p, tg_p, q_p, m->tg_tmp.p, m->qg_tmp.p, m->dt_t.p, m->dt_q.p,
                                               m->rain.p, m->conv_rain.p, m->precip.p);
    isca_core_mark(m->dyn, "phys_conv_post");
    t_in = m->tg_tmp.p; q_in = m->qg_tmp.p;
  } else if (m->mc.convection_scheme == 2) {                     // 'DRY' (:918-928): dt_tg += conv_dt_tg; no precipitation
    launch_dry_convection(p, m->dry_tau, m->dry_gamma, tg_p, pf_p, ph_p, m->c_Tref.p, m->c_dT.p, m->cape.p, m->cin.p, klzb, klcl);
    add1_kernel<<<nblk(n3), 256, 0, st>>>(n3, m->dt_t.p, m->c_dT.p);
    MCK(cudaMemsetAsync(m->precip.p, 0, nc * sizeof(double), st));
    MCK(cudaMemsetAsync(m->conv_rain.p, 0, nc * sizeof(double), st));
  } else {
    MCK(cudaMemsetAsync(m->precip.p, 0, nc * sizeof(double), st));
    MCK(cudaMemsetAsync(m->conv_rain.p, 0, nc * sizeof(double), st));
  }
  if (m->mc.convection_scheme != 2) {                            // `if (r_conv_scheme .ne. DRY_CONV)` (:977): no large-scale condensation
    launch_lscale(p, t_in, q_in, pf_p, ph_p, m->rain.p, m->c_dT.p, m->c_dq.p);
    isca_core_mark(m->dyn, "phys_lscale_cond");
    cond_post_kernel<<<nblk(n3), 256, 0, st>>>(n3, nc, delta_t, m->c_dT.p, m->c_dq.p, m->dt_t.p, m->dt_q.p, m->rain.p, m->precip.p);
    isca_core_mark(m->dyn, "phys_cond_post");
  }
  if (!m->rr && m->seasonal) {                                   // Time_diag = Time (:1054); days of 86400 s as get_time returns them
    const double days = floor(m->time_s / 86400.0), seconds = m->time_s - 86400.0 * days;
    if (isca_gray_coszen_device(m->sdc, m->s_orb, st, days, seconds, (int)nc, m->lat2d.p, m->lon2d.p, m->coszen.p))
      return mfail(m, "two_stream_gray_rad: zenith-angle kernel launch failed");
    scale_kernel<<<nblk(nc), 256, 0, st>>>(nc, p->insol.p, m->coszen.p, p->pc.solar_constant);
  }
  if (!m->rr) launch_gray_down(p, m->lat2d.p, ph_c, tg_p, q_p, m->albedo.p, m->net_sw.p, m->lw_down.p);   // q = grid_tracers(previous, nsphum), :1068
  isca_core_mark(m->dyn, "phys_gray_down");
  // surface_flux on the lowest model level (:1076-1132)
  sub_kernel<<<nblk(nc), 256, 0, st>>>(nc, m->z_atm.p, zf_c + (size_t)(K - 1) * nc, m->z_surf.p);
  IscaSurfaceFluxArgs a;
  a.t_atm = tg_p + (size_t)(K - 1) * nc; a.q_atm = q_p + (size_t)(K - 1) * nc; a.u_atm = ug_p + (size_t)(K - 1) * nc;
  a.v_atm = vg_p + (size_t)(K - 1) * nc; a.p_atm = pf_c + (size_t)(K - 1) * nc; a.z_atm = m->z_atm.p; a.p_surf = ph_c + (size_t)K * nc;
  a.t_surf = m->t_surf.p; a.t_ca = m->t_surf.p; a.u_surf = m->zero2.p; a.v_surf = m->zero2.p;
  a.rough_mom = m->rough_m.p; a.rough_heat = m->rough_h.p; a.rough_moist = m->rough_q.p; a.rough_scale = m->rough_m.p; a.gust = m->gust.p;
  a.land = land; a.q_surf = m->q_surf.p;
  double** outs[28] = {&a.flux_t, &a.flux_q, &a.flux_r, &a.flux_u, &a.flux_v, &a.cd_m, &a.cd_t, &a.cd_q, &a.w_atm, &a.u_star, &a.b_star,
                       &a.q_star, &a.dhdt_surf, &a.dedt_surf, &a.dedq_surf, &a.drdt_surf, &a.dhdt_atm, &a.dedq_atm, &a.dtaudu_atm,
                       &a.dtaudv_atm, &a.ex_del_m, &a.ex_del_h, &a.ex_del_q, &a.temp_2m, &a.u_10m, &a.v_10m, &a.q_2m, &a.rh_2m};
  for (int i = 0; i < 28; ++i) *outs[i] = m->sf.p + (size_t)i * nc;
  launch_surface_flux(p, a);
  isca_core_mark(m->dyn, "phys_surface_flux");
  if (!m->rr) launch_gray_up(p, m->lat2d.p, ph_c, tg_p, q_p, m->t_surf.p, m->albedo.p, m->dt_t.p, nullptr);
  else {
    // run_rrtmg (rrtm_radiation.F90:640-660): radiation alarm; between radiation steps the stored heating and surface fluxes are reused
    const double dt_rad = m->rdc.dt_rad > 0 ? (double)m->rdc.dt_rad : (double)(long)v.dt_atmos;
    if (m->time_s - m->dt_last >= dt_rad) {
      m->dt_last = m->time_s;
      if (isca_rrtm_coszen_device(m->rdc, m->orb_angle, st, m->time_s, (int)nc, m->lat2d.p, m->lon2d.p, m->coszen.p, nullptr))
        return mfail(m, "run_rrtmg: zenith-angle kernel launch failed");
      if (isca_rrtm_run_device(m->rr, st, pf_c, ph_c, zf_c, zh_c, tg_p, q_p, m->have_o3 ? m->o3.p : nullptr, m->t_surf.p, m->albedo.p,
                               m->coszen.p, m->dt_t.p, m->tdt_rad.p, m->net_sw.p, m->lw_down.p, m->olr.p, m->toa_sw.p))
        return mfail(m, std::string("run_rrtmg: ") + isca_b200_rrtm_last_error(m->rr));
      m->n_rad_calls++;
      isca_core_mark(m->dyn, "phys_rrtmg_call");                   // the whole radiation call (only on radiation steps)
    } else if (m->rdc.store_intermediate_rad) {
      add1_kernel<<<nblk(n3), 256, 0, st>>>(n3, m->dt_t.p, m->tdt_rad.p);
    } else {
      MCK(cudaMemsetAsync(m->net_sw.p, 0, nc * sizeof(double), st));
      MCK(cudaMemsetAsync(m->lw_down.p, 0, nc * sizeof(double), st));
    }
  }
  isca_core_mark(m->dyn, "phys_radiation");
  if (m->mc.do_damping) {
    int nlev = rayleigh_nlev(m->pref.data(), K, p->cfg.sponge_pbottom);
    launch_rayleigh(p, nlev, delta_t, pf_c, ug_p, vg_p, m->w1.p, m->w2.p, m->w3.p, 0);
    const size_t ns = (size_t)nlev * nc;                         // the sponge acts on the top nlev levels only (damping_driver.f90:594-636)
    if (ns) add3_kernel<<<nblk(ns), 256, 0, st>>>(ns, m->dt_u.p, m->w1.p, m->dt_v.p, m->w2.p, m->dt_t.p, m->w3.p);
  }
  isca_core_mark(m->dyn, "phys_damping");
  // vert_turb_driver, do_diffusivity branch (vert_turb_driver.F90:277-292): on the `current` fields (use_tau = .true.) or on
  // previous + delta_t * tendencies (use_tau = .false., formed inside the kernel); the diffusivities start from zero (:248-253)
  if (m->mc.use_tau)
    launch_diffusivity(p, v.T[cur], v.q[cur], v.u[cur], v.v[cur], zf_c, zh_c, a.u_star, a.b_star, m->z_pbl.p, m->diff_m.p, m->diff_t.p,
                       nullptr, nullptr, nullptr, nullptr, 0.0, 0);
  else
    launch_diffusivity(p, tg_p, q_p, ug_p, vg_p, zf_c, zh_c, a.u_star, a.b_star, m->z_pbl.p, m->diff_m.p, m->diff_t.p,
                       m->dt_t.p, m->dt_q.p, m->dt_u.p, m->dt_v.p, delta_t, 0);
  isca_core_mark(m->dyn, "phys_diffusivity");
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->gust.p, nc, m->mc.constant_gust);
  launch_vert_diff_down(p, delta_t, ug_p, vg_p, tg_p, q_p, m->diff_m.p, m->diff_t.p, ph_c, zf_c, a.flux_u, a.flux_v, a.dtaudu_atm, a.dtaudv_atm,
                        m->dt_u.p, m->dt_v.p, m->dt_t.p, m->dt_q.p, m->diss.p);
  isca_core_mark(m->dyn, "phys_vert_diff_down");
  launch_mixed_layer(p, v.dt_atmos, m->t_surf.p, a.flux_t, a.flux_q, a.flux_r, m->net_sw.p, m->lw_down.p, a.dhdt_surf, a.dedt_surf, a.dedq_surf,
                     a.drdt_surf, a.dhdt_atm, a.dedq_atm, m->dts.p);
  launch_vert_diff_up(p, delta_t, m->dt_t.p, m->dt_q.p);
  isca_core_mark(m->dyn, "phys_mixed_layer_vert_diff_up");
  if (ev_phys_end) MCK(cudaEventRecord(ev_phys_end, st));
  if (isca_core_step_ext(m->dyn, m->dt_u.p, m->dt_v.p, m->dt_t.p, m->dt_q.p)) return mfail(m, std::string("spectral_dynamics: ") + isca_b200_last_error(m->dyn));
  m->time_s += v.dt_atmos;                                     // atmos_model.F90: Time_atmos = Time_atmos + Time_step_atmos
  return 0;
}

}  // namespace

extern "C" {

int isca_b200_moist_default_config(IscaMoistConfig* c) {
  if (!c) return 1;
  std::memset(c, 0, sizeof(*c));
  c->abi_version = 2; c->convection_scheme = 1; c->do_damping = 0; c->use_tau = 1;
  c->roughness_mom = 0.05; c->roughness_heat = 0.05; c->roughness_moist = 0.05;
  c->mixed_layer_depth = 40.0; c->albedo_value = 0.06; c->rho_cp = 1.035e3 * 3989.24495292815;
  c->constant_gust = 1.0;
  return 0;
}

const char* isca_b200_moist_last_error(IscaMoist m) { return m ? m->err.c_str() : g_merr.c_str(); }
IscaHandle isca_b200_moist_dycore(IscaMoist m) { return m ? m->dyn : nullptr; }

int isca_b200_moist_destroy(IscaMoist m) {
  if (!m) return 0;
  if (m->io_ready) {
    cudaStreamSynchronize(m->io_h2d); cudaStreamSynchronize(m->io_d2h);
    for (cudaEvent_t e : {m->ev_h2d, m->ev_snap, m->ev_d2h, m->ev_d2h_prev, m->ev_main}) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(m->io_h2d); cudaStreamDestroy(m->io_d2h);
  }
  if (m->rr) isca_b200_rrtm_destroy(m->rr);
  if (m->phy) isca_b200_physics_destroy(m->phy);
  if (m->dyn) isca_b200_destroy(m->dyn);
  delete m;
  return 0;
}

int isca_b200_moist_create(const IscaConfig* dyn, const IscaPhysicsConfig* phys, const IscaMoistConfig* mc, IscaMoist* out) {
  return isca_b200_moist_create_ranked(dyn, phys, mc, 0, 1, nullptr, out);
}

int isca_b200_moist_create_ranked(const IscaConfig* dyn, const IscaPhysicsConfig* phys, const IscaMoistConfig* mc, int rank, int nranks,
                                  const void* nccl_unique_id, IscaMoist* out) {
  IscaMoist m = nullptr;
  if (!dyn || !phys || !mc || !out) return mfail(nullptr, "null argument");
  if (mc->abi_version != 2) return mfail(nullptr, "IscaMoistConfig abi_version mismatch");
  if (mc->convection_scheme < 0 || mc->convection_scheme > 3)
    return mfail(nullptr, "idealized_moist_phys: Invalid convection scheme (NONE, SIMPLE_BETTS_MILLER, FULL_BETTS_MILLER and DRY are built)");
  if (dyn->num_tracers != 1) return mfail(nullptr, "idealized_moist_model needs the sphum grid tracer (num_tracers = 1)");
  m = new IscaMoist_t();
  m->mc = *mc;
  if (isca_b200_create(dyn, rank, nranks, nccl_unique_id, &m->dyn)) { std::string e = isca_b200_last_error(nullptr); delete m; return mfail(nullptr, "dynamical core: " + e); }
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) { std::string e = isca_b200_last_error(m->dyn); isca_b200_moist_destroy(m); return mfail(nullptr, e); }
  m->I = v.I; m->J = v.Jloc; m->K = v.K; m->nc = (size_t)v.I * v.Jloc; m->n3 = m->nc * v.K;
  IscaPhysicsConfig pc = *phys;
  pc.num_lon = v.I; pc.num_lat = v.Jloc; pc.num_levels = v.K; pc.grav = dyn->grav; pc.rdgas = dyn->rdgas; pc.cp_air = dyn->rdgas / dyn->kappa;
  if (isca_b200_physics_create(&pc, &m->phy)) { std::string e = isca_b200_physics_last_error(nullptr); isca_b200_moist_destroy(m); return mfail(nullptr, "physics: " + e); }
  if (mc->convection_scheme == 1 && !m->phy->lcl_err.empty()) {      // qe_moist_convection_init (idealized_moist_phys.F90:505)
    std::string e = m->phy->lcl_err; isca_b200_moist_destroy(m); return mfail(nullptr, "physics: " + e);
  }
  cudaStreamDestroy(m->phy->st);
  m->phy->st = v.st; m->phy->owns_stream = false;
  const size_t nc = m->nc, n3 = m->n3;
  bool ok = true;
  for (int s = 0; s < 2; ++s) ok &= m->p_full[s].ensure(n3) && m->p_half[s].ensure(n3 + nc) && m->z_full[s].ensure(n3) && m->z_half[s].ensure(n3 + nc);
  Dev* d3[] = {&m->dt_u, &m->dt_v, &m->dt_t, &m->dt_q, &m->tg_tmp, &m->qg_tmp, &m->c_dT, &m->c_dq, &m->c_qref, &m->c_Tref, &m->diff_m, &m->diff_t,
               &m->diss, &m->w1, &m->w2, &m->w3};
  for (Dev* d : d3) ok &= d->ensure(n3);
  Dev* d2[] = {&m->lat2d, &m->z_surf, &m->t_surf, &m->q_surf, &m->gust, &m->albedo, &m->zero2, &m->rough_m, &m->rough_h, &m->rough_q, &m->z_atm,
               &m->precip, &m->rain, &m->conv_rain, &m->cape, &m->cin, &m->itq, &m->itt, &m->net_sw, &m->lw_down, &m->z_pbl, &m->dts};
  for (Dev* d : d2) ok &= d->ensure(nc);
  ok &= m->sf.ensure(28 * nc) && m->iwork.ensure(2 * nc + 2);
  if (!ok || prepare_vert_diff_state(m->phy)) { isca_b200_moist_destroy(m); return mfail(nullptr, "cudaMalloc failed"); }
  // reference pressures of the sponge (idealized_moist_phys.F90:628-631): pressure_variables at PSTD_MKS
  {
    std::vector<double> pk(v.K + 1), bk(v.K + 1);
    if (isca_b200_get_table(m->dyn, ISCA_TB_PK, pk.data(), v.K + 1) || isca_b200_get_table(m->dyn, ISCA_TB_BK, bk.data(), v.K + 1)) {
      isca_b200_moist_destroy(m); return mfail(nullptr, "get_table failed");
    }
    m->pref.assign(v.K + 1, 0.0);
    std::vector<double> ph(v.K + 1), lh(v.K + 1, 0.0);
    const double ps = pc.pstd_mks;
    for (int k = 0; k <= v.K; ++k) { ph[k] = pk[k] + bk[k] * ps; if (ph[k] > 0.0) lh[k] = std::log(ph[k]); }
    for (int k = 0; k < v.K; ++k) {                             // Simmons-Burridge full levels (press_and_geopot.F90:152-221)
      double lf;
      if (k == 0 && ph[0] == 0.0) lf = lh[1] - 1.0;
      else { double alpha = 1.0 - ph[k] * (lh[k + 1] - lh[k]) / (ph[k + 1] - ph[k]); lf = lh[k + 1] - alpha; }
      m->pref[k] = std::exp(lf);
    }
    m->pref[v.K] = ps;
  }
  *out = m;
  return 0;
}

int isca_b200_moist_init(IscaMoist m) {
  if (!m) return mfail(nullptr, "null handle");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  cudaStream_t st = v.st;
  const size_t nc = m->nc;
  dim3 g2((m->I + 127) / 128, m->J);
  lat2d_kernel<<<g2, 128, 0, st>>>(m->lat2d.p, v.rad_lat, m->I, m->J);
  // z_surf = surf_geopotential / grav (:565-566)
  {
    std::vector<double> ph(nc);
    MCK(cudaMemcpyAsync(ph.data(), v.phis, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < nc; ++i) ph[i] = ph[i] / v.grav;
    MCK(cudaMemcpyAsync(m->z_surf.p, ph.data(), nc * sizeof(double), cudaMemcpyHostToDevice, st));
    MCK(cudaStreamSynchronize(st));
  }
  add_const_kernel<<<nblk(nc), 256, 0, st>>>(nc, m->t_surf.p, v.T[v.current] + (size_t)(m->K - 1) * nc, 1.0);     // :643
  MCK(cudaMemsetAsync(m->q_surf.p, 0, nc * sizeof(double), st));
  MCK(cudaMemsetAsync(m->zero2.p, 0, nc * sizeof(double), st));
  MCK(cudaMemsetAsync(m->iwork.p, 0, (2 * nc + 2) * sizeof(double), st));
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->gust.p, nc, 1.0);
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->albedo.p, nc, m->mc.albedo_value);
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->rough_m.p, nc, m->mc.roughness_mom);
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->rough_h.p, nc, m->mc.roughness_heat);
  fill_kernel<<<nblk(nc), 256, 0, st>>>(m->rough_q.p, nc, m->mc.roughness_moist);
  IscaPhysics p = m->phy;
  if (!p->state[ST_ML_HEAT_CAP].ensure(nc) || !p->state[ST_ML_QFLUX].ensure(nc)) return mfail(m, "cudaMalloc failed");
  fill_kernel<<<nblk(nc), 256, 0, st>>>(p->state[ST_ML_HEAT_CAP].p, nc, m->mc.mixed_layer_depth * m->mc.rho_cp);
  MCK(cudaMemsetAsync(p->state[ST_ML_QFLUX].p, 0, nc * sizeof(double), st));
  Dev* zero[] = {&m->precip, &m->conv_rain, &m->cape, &m->z_pbl, &m->dts, &m->net_sw, &m->lw_down};
  for (Dev* d : zero) MCK(cudaMemsetAsync(d->p, 0, nc * sizeof(double), st));
  MCK(cudaMemsetAsync(m->sf.p, 0, 28 * nc * sizeof(double), st));
  MCK(cudaStreamSynchronize(st));
  p->vert_diff_down_done = false;
  if (m->seasonal && !m->rr) {
    lon2d_kernel<<<g2, 128, 0, st>>>(m->lon2d.p, m->I, m->J);
    MCK(cudaMemsetAsync(m->coszen.p, 0, nc * sizeof(double), st));
    MCK(cudaStreamSynchronize(st));
  }
  if (m->rr) {
    lon2d_kernel<<<g2, 128, 0, st>>>(m->lon2d.p, m->I, m->J);
    Dev* z2[] = {&m->coszen, &m->olr, &m->toa_sw};
    for (Dev* d : z2) MCK(cudaMemsetAsync(d->p, 0, nc * sizeof(double), st));
    MCK(cudaMemsetAsync(m->tdt_rad.p, 0, m->n3 * sizeof(double), st));
    MCK(cudaStreamSynchronize(st));
    const double dt_rad = m->rdc.dt_rad > 0 ? (double)m->rdc.dt_rad : (double)(long)v.dt_atmos;
    m->dt_last = -dt_rad;                                      // rrtm_radiation_init: radiation at the first time step
    m->n_rad_calls = 0;
  }
  m->initialized = true;
  return 0;
}

int isca_b200_moist_step(IscaMoist m, int n_steps) {
  if (!m) return mfail(nullptr, "null handle");
  if (!m->initialized) return mfail(m, "idealized_moist_phys: module not initialized (isca_b200_moist_init has not been called)");
  if (m->mc.convection_scheme == 2 && !(m->dry_tau > 0.0))
    return mfail(m, "dry_convection: tau / gamma not set (isca_b200_moist_set_dry_convection; dry_convection_nml has no defaults)");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  struct Events {                             // destroyed on every exit path
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
  } ev;
  for (auto& x : ev.e) MCK(cudaEventCreate(&x));
  cudaEvent_t &e0 = ev.e[0], &e1 = ev.e[1], &e2 = ev.e[2], &e3 = ev.e[3];
  double phys_ms = 0.0;
  MCK(cudaEventRecord(e0, v.st));
  for (int i = 0; i < n_steps; ++i) {
    const bool last = (i == n_steps - 1);
    if (last) MCK(cudaEventRecord(e2, v.st));
    if (moist_step_once(m, last ? e1 : nullptr)) return 1;
  }
  MCK(cudaEventRecord(e3, v.st));
  // device error flags: saturation-table / LCL-table overflow, zero effective heat capacity; temperature range of the core
  int e = 0;
  MCK(cudaMemcpyAsync(&e, m->phy->d_err, sizeof(int), cudaMemcpyDeviceToHost, v.st));
  MCK(cudaStreamSynchronize(v.st));
  if (n_steps > 0) {
    float a = 0, b = 0;
    MCK(cudaEventElapsedTime(&a, e0, e3)); MCK(cudaEventElapsedTime(&b, e2, e1));
    m->ms_step = a / n_steps; phys_ms = b; m->ms_phys = phys_ms;
  }
  if (e) {
    MCK(cudaMemsetAsync(m->phy->d_err, 0, sizeof(int), v.st));
    return mfail(m, "idealized_moist_phys: lookup_es / get_lcl_temp table overflow or zero effective heat capacity (device error flag " + std::to_string(e) + ")");
  }
  if (isca_core_check(m->dyn)) return mfail(m, isca_b200_last_error(m->dyn));
  return 0;
}

// average milliseconds per kernel group (physics kernels "phys_*" + the dynamical core's groups) over n eager steps
int isca_b200_moist_profile_step(IscaMoist m, int n_steps, double* ms_out, int max_groups, char* names, int capacity) {
  if (!m || !ms_out || !names) return -1;
  if (!m->initialized) { mfail(m, "moist_profile_step: module not initialized"); return -1; }
  std::vector<std::string> order;
  std::map<std::string, double> acc;
  for (int i = 0; i < n_steps; ++i) {
    isca_core_profile_begin(m->dyn);
    int rc = moist_step_once(m, nullptr);
    if (isca_core_profile_end(m->dyn, order, acc) || rc) { if (!rc) mfail(m, "moist_profile_step: event timing failed"); return -1; }
  }
  std::string joined;
  int n = 0;
  for (auto& nm : order) {
    if (n >= max_groups) break;
    ms_out[n++] = acc[nm] / (n_steps > 0 ? n_steps : 1);
    joined += (joined.empty() ? "" : ";") + nm;
  }
  if ((int)joined.size() + 1 > capacity) { mfail(m, "moist_profile_step: names buffer too small"); return -1; }
  std::memcpy(names, joined.c_str(), joined.size() + 1);
  return n;
}

// device pointer / element count of a double-valued field of isca_b200_moist_get (id 9, convflag, is an int plane: not here)
static int moist_field_device(IscaMoist m, int id, const double** out, size_t* count) {
  const size_t nc = m->nc, n3 = m->n3;
  const double* src = nullptr; size_t n = nc;
  switch (id) {
    case 0: src = m->t_surf.p; break;
    case 1: src = m->precip.p; break;
    case 2: src = m->sf.p + SF_FLUX_T * nc; break;
    case 3: src = m->sf.p + SF_FLUX_Q * nc; break;
    case 4: src = m->z_pbl.p; break;
    case 5: src = m->net_sw.p; break;
    case 6: src = m->lw_down.p; break;
    case 7: src = m->conv_rain.p; break;
    case 8: src = m->cape.p; break;
    case 17: if (!m->rr) return mfail(m, "moist field: olr needs do_rrtm_radiation"); src = m->olr.p; break;
    case 18: if (!m->rr) return mfail(m, "moist field: toa_sw needs do_rrtm_radiation"); src = m->toa_sw.p; break;
    case 34: src = m->dt_t.p; n = n3; break;
    case 35: src = m->dt_q.p; n = n3; break;
    default: return mfail(m, "moist field: this id is not available to isca_b200_moist_step_io");
  }
  *out = src; *count = n;
  return 0;
}

// atmosphere(Time) with the host I/O of a step, software-pipelined (include/isca_b200_physics.h).  Streams: the core's (compute), one
// host->device and one device->host copy stream.  Per call: [h2d] ozone upload after the previous step has finished with the old field
// -> [compute] the step, then device-to-device snapshots of the requested fields into a staging buffer (after the previous call's
// downloads have drained it) -> [d2h] downloads of the snapshots into the caller's (pinned) arrays.  The call returns without waiting:
// the downloads of step n overlap the compute of step n+1 and the upload of its input (PCIe is full duplex).
int isca_b200_moist_step_io(IscaMoist m, const double* o3_host, int n_out, const int* kinds, const int* ids, const int* levels,
                            double* const* host_out) {
  if (!m) return mfail(nullptr, "null handle");
  if (!m->initialized) return mfail(m, "idealized_moist_phys: module not initialized (isca_b200_moist_init has not been called)");
  if (n_out < 0 || (n_out > 0 && (!kinds || !ids || !levels || !host_out))) return mfail(m, "moist_step_io: null output description");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  if (!m->io_ready) {
    MCK(cudaStreamCreateWithFlags(&m->io_h2d, cudaStreamNonBlocking)); MCK(cudaStreamCreateWithFlags(&m->io_d2h, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&m->ev_h2d, &m->ev_snap, &m->ev_d2h, &m->ev_d2h_prev, &m->ev_main}) MCK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    MCK(cudaEventRecord(m->ev_d2h, m->io_d2h)); MCK(cudaEventRecord(m->ev_d2h_prev, m->io_d2h)); MCK(cudaEventRecord(m->ev_main, v.st));
    m->io_ready = true;
  }
  if (o3_host) {
    if (!m->rr) return mfail(m, "moist_step_io: the ozone field needs do_rrtm_radiation");
    if (!m->o3.ensure(m->n3)) return mfail(m, "cudaMalloc failed");
    MCK(cudaStreamWaitEvent(m->io_h2d, m->ev_main, 0));          // the previous step no longer reads the old field
    MCK(cudaMemcpyAsync(m->o3.p, o3_host, m->n3 * sizeof(double), cudaMemcpyHostToDevice, m->io_h2d));
    MCK(cudaEventRecord(m->ev_h2d, m->io_h2d));
    MCK(cudaStreamWaitEvent(v.st, m->ev_h2d, 0));
    m->have_o3 = true;
  }
  if (moist_step_once(m, nullptr)) return 1;
  // resolve the fields (may launch materialising kernels on the compute stream), then snapshot them
  std::vector<const double*> src(n_out, nullptr);
  std::vector<size_t> cnt(n_out, 0);
  size_t total = 0;
  for (int i = 0; i < n_out; ++i) {
    if (!host_out[i]) return mfail(m, "moist_step_io: null output array");
    if (kinds[i] == 0) { if (isca_core_field_device(m->dyn, ids[i], levels[i], &src[i], &cnt[i])) return mfail(m, isca_b200_last_error(m->dyn)); }
    else if (moist_field_device(m, ids[i], &src[i], &cnt[i])) return 1;
    total += cnt[i];
  }
  if (total > m->io_stage.n) {                                  // first call (or a larger request): nothing is in flight that uses it
    MCK(cudaStreamSynchronize(m->io_d2h));
    if (!m->io_stage.ensure(total)) return mfail(m, "cudaMalloc failed");
  }
  MCK(cudaStreamWaitEvent(v.st, m->ev_d2h, 0));                 // the previous downloads have drained the staging buffer
  size_t off = 0;
  for (int i = 0; i < n_out; ++i) {
    MCK(cudaMemcpyAsync(m->io_stage.p + off, src[i], cnt[i] * sizeof(double), cudaMemcpyDeviceToDevice, v.st));
    off += cnt[i];
  }
  MCK(cudaEventRecord(m->ev_snap, v.st));
  MCK(cudaEventRecord(m->ev_main, v.st));
  MCK(cudaStreamWaitEvent(m->io_d2h, m->ev_snap, 0));
  off = 0;
  for (int i = 0; i < n_out; ++i) {
    MCK(cudaMemcpyAsync(host_out[i], m->io_stage.p + off, cnt[i] * sizeof(double), cudaMemcpyDeviceToHost, m->io_d2h));
    off += cnt[i];
  }
  std::swap(m->ev_d2h, m->ev_d2h_prev);                         // ev_d2h_prev = the call before this one (isca_b200_moist_io_wait)
  MCK(cudaEventRecord(m->ev_d2h, m->io_d2h));
  m->io_pending = true;
  return 0;
}

// wait for the downloads of the last call (age = 0) or of the call before it (age = 1), without draining the pipeline
int isca_b200_moist_io_wait(IscaMoist m, int age) {
  if (!m) return mfail(nullptr, "null handle");
  if (!m->io_ready) return 0;
  if (age != 0 && age != 1) return mfail(m, "moist_io_wait: age must be 0 or 1");
  MCK(cudaEventSynchronize(age == 0 ? m->ev_d2h : m->ev_d2h_prev));
  return 0;
}

// wait until the host arrays of the last isca_b200_moist_step_io call are complete; reports the device error flags of the steps since
int isca_b200_moist_io_sync(IscaMoist m) {
  if (!m) return mfail(nullptr, "null handle");
  if (!m->io_ready) return 0;
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  MCK(cudaEventSynchronize(m->ev_d2h));
  int e = 0;
  MCK(cudaMemcpyAsync(&e, m->phy->d_err, sizeof(int), cudaMemcpyDeviceToHost, v.st));
  MCK(cudaStreamSynchronize(v.st));
  m->io_pending = false;
  if (e) {
    MCK(cudaMemsetAsync(m->phy->d_err, 0, sizeof(int), v.st));
    return mfail(m, "idealized_moist_phys: lookup_es / get_lcl_temp table overflow or zero effective heat capacity (device error flag " + std::to_string(e) + ")");
  }
  if (isca_core_check(m->dyn)) return mfail(m, isca_b200_last_error(m->dyn));
  return 0;
}

int isca_b200_moist_get(IscaMoist m, int id, double* host) {
  if (!m) return mfail(nullptr, "null handle");
  if (!host) return mfail(m, "null output array");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  const size_t nc = m->nc, n3 = m->n3;
  const double* src = nullptr; size_t n = nc;
  switch (id) {
    case 0: src = m->t_surf.p; break;
    case 1: src = m->precip.p; break;
    case 2: src = m->sf.p + SF_FLUX_T * nc; break;
    case 3: src = m->sf.p + SF_FLUX_Q * nc; break;
    case 4: src = m->z_pbl.p; break;
    case 5: src = m->net_sw.p; break;
    case 6: src = m->lw_down.p; break;
    case 7: src = m->conv_rain.p; break;
    case 8: src = m->cape.p; break;
    case 10: src = m->q_surf.p; break;
    case 11: src = m->sf.p + SF_U_STAR * nc; break;
    case 12: src = m->sf.p + SF_B_STAR * nc; break;
    case 13: src = m->sf.p + SF_FLUX_U * nc; break;
    case 14: src = m->sf.p + SF_FLUX_V * nc; break;
    case 15: src = m->dts.p; break;
    case 16: if (!m->rr && !m->seasonal) return mfail(m, "moist_get: coszen needs do_rrtm_radiation or do_seasonal"); src = m->coszen.p; break;
    case 17: if (!m->rr) return mfail(m, "moist_get: olr needs do_rrtm_radiation"); src = m->olr.p; break;
    case 18: if (!m->rr) return mfail(m, "moist_get: toa_sw needs do_rrtm_radiation"); src = m->toa_sw.p; break;
    case 38: if (!m->rr) return mfail(m, "moist_get: tdt_rad needs do_rrtm_radiation"); src = m->tdt_rad.p; n = n3; break;
    case 32: src = m->dt_u.p; n = n3; break;
    case 33: src = m->dt_v.p; n = n3; break;
    case 34: src = m->dt_t.p; n = n3; break;
    case 35: src = m->dt_q.p; n = n3; break;
    case 36: src = m->diff_m.p; n = n3; break;
    case 37: src = m->diff_t.p; n = n3; break;
    case 9: {
      std::vector<int> f(nc);
      MCK(cudaMemcpyAsync(f.data(), reinterpret_cast<int*>(m->iwork.p) + nc, nc * sizeof(int), cudaMemcpyDeviceToHost, v.st));
      MCK(cudaStreamSynchronize(v.st));
      for (size_t i = 0; i < nc; ++i) host[i] = (double)f[i];
      return 0;
    }
    default: return mfail(m, "moist_get: unknown field id");
  }
  MCK(cudaMemcpyAsync(host, src, n * sizeof(double), cudaMemcpyDeviceToHost, v.st));
  MCK(cudaStreamSynchronize(v.st));
  return 0;
}

int isca_b200_moist_set_t_surf(IscaMoist m, const double* host) {
  if (!m) return mfail(nullptr, "null handle");
  if (!host) return mfail(m, "null input array");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  MCK(cudaMemcpyAsync(m->t_surf.p, host, m->nc * sizeof(double), cudaMemcpyHostToDevice, v.st));
  MCK(cudaStreamSynchronize(v.st));
  return 0;
}

int isca_b200_moist_set_sst(IscaMoist m, const double* host) {
  if (!m) return mfail(nullptr, "null handle");
  if (isca_b200_mixed_layer_set_sst(m->phy, host)) return mfail(m, isca_b200_physics_last_error(m->phy));
  return 0;
}

int isca_b200_moist_set_surface(IscaMoist m, int id, const double* host) {
  if (!m) return mfail(nullptr, "null handle");
  if (!host) return mfail(m, "null input array");
  if (!m->initialized) return mfail(m, "moist_set_surface: call isca_b200_moist_init first");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  double* dst = nullptr;
  switch (id) {
    case 20: dst = m->albedo.p; break;
    case 21: dst = m->rough_m.p; break;
    case 22: dst = m->rough_h.p; break;
    case 23: dst = m->rough_q.p; break;
    case 24: dst = m->phy->state[ST_ML_HEAT_CAP].p; break;
    case 25: dst = m->dts.p; break;                              // staged, converted to the int mask below
    default: return mfail(m, "moist_set_surface: unknown field id");
  }
  MCK(cudaMemcpyAsync(dst, host, m->nc * sizeof(double), cudaMemcpyHostToDevice, v.st));
  if (id == 25) {
    mask_to_int_kernel<<<nblk(m->nc), 256, 0, v.st>>>(m->nc, m->dts.p, reinterpret_cast<int*>(m->iwork.p));
    MCK(cudaMemsetAsync(m->dts.p, 0, m->nc * sizeof(double), v.st));
  }
  MCK(cudaStreamSynchronize(v.st));
  return 0;
}

int isca_b200_moist_set_dry_convection(IscaMoist m, double tau, double gamma) {
  if (!m) return mfail(nullptr, "null handle");
  if (!(tau > 0.0)) return mfail(m, "dry_convection: tau must be positive");
  m->dry_tau = tau; m->dry_gamma = gamma;
  return 0;
}

int isca_b200_moist_set_co2(IscaMoist m, double carbon_conc) {
  if (!m) return mfail(nullptr, "null handle");
  if (m->rr) return mfail(m, "moist_set_co2: do_read_co2 belongs to two_stream_gray_rad (RRTMG: co2ppmv of rrtm_radiation_nml)");
  if (isca_b200_two_stream_gray_rad_set_co2(m->phy, carbon_conc)) return mfail(m, isca_b200_physics_last_error(m->phy));
  return 0;
}

int isca_b200_moist_set_betts_miller(IscaMoist m, const IscaBettsMillerConfig* cfg) {
  if (!m) return mfail(nullptr, "null handle");
  if (m->mc.convection_scheme != 3) return mfail(m, "moist_set_betts_miller: convection_scheme is not 'FULL_BETTS_MILLER'");
  if (isca_b200_betts_miller_init(m->phy, cfg)) return mfail(m, isca_b200_physics_last_error(m->phy));
  return 0;
}

int isca_b200_moist_set_ocean_qflux(IscaMoist m, const double* host) {
  if (!m) return mfail(nullptr, "null handle");
  if (!host) return mfail(m, "null input array");
  if (!m->initialized) return mfail(m, "moist_set_ocean_qflux: call isca_b200_moist_init first");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  MCK(cudaMemcpyAsync(m->phy->state[ST_ML_QFLUX].p, host, m->nc * sizeof(double), cudaMemcpyHostToDevice, v.st));
  MCK(cudaStreamSynchronize(v.st));
  return 0;
}

int isca_b200_moist_timing(IscaMoist m, double* ms_step, double* ms_physics) {
  if (!m) return mfail(nullptr, "null handle");
  if (ms_step) *ms_step = m->ms_step;
  if (ms_physics) *ms_physics = m->ms_phys;
  return 0;
}

int isca_b200_moist_use_rrtm(IscaMoist m, const IscaRrtmConfig* rc, const IscaRrtmDriverConfig* dc, const char* table_path) {
  if (!m) return mfail(nullptr, "null handle");
  if (!rc || !dc || !table_path) return mfail(m, "moist_use_rrtm: null argument");
  if (dc->abi_version != 1) return mfail(m, "IscaRrtmDriverConfig abi_version mismatch");
  if (m->initialized) return mfail(m, "moist_use_rrtm must be called before isca_b200_moist_init");
  if (m->rr) return mfail(m, "moist_use_rrtm: RRTMG is already enabled");
  if (m->seasonal) return mfail(m, "moist_use_rrtm: do_seasonal of two_stream_gray_rad is enabled on this handle");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  if (dc->num_angles < 1 || dc->day_in_s <= 0.0 || dc->year_in_s <= 0.0) return mfail(m, "moist_use_rrtm: bad astronomy / calendar values");
  if (dc->dt_rad > 0) {                                        // rrtm_radiation_init (rrtm_radiation.F90:384-402)
    const long step = (long)v.dt_atmos;
    if (dc->dt_rad > step && dc->dt_rad % step != 0) return mfail(m, "rrtm_gases_init: dt_rad must be an integer multiple of dt_atmos");
  }
  IscaRrtmConfig c = *rc;
  c.num_lon = m->I; c.num_lat = m->J; c.num_levels = m->K;
  if (isca_b200_rrtm_create(&c, table_path, &m->rr)) { m->rr = nullptr; return mfail(m, std::string("rrtm: ") + isca_b200_rrtm_last_error(nullptr)); }
  m->rdc = *dc;
  if (m->rdc.dt_rad <= 0) m->rdc.dt_rad = (int)(long)v.dt_atmos;               // rrtm_radiation_init (rrtm_radiation.F90:378-380)
  if (m->rdc.dt_rad_avg <= 0) m->rdc.dt_rad_avg = m->rdc.dt_rad;              // :405
  m->orb_angle = isca_rrtm_orbit(*dc);
  bool ok = m->tdt_rad.ensure(m->n3) && m->coszen.ensure(m->nc) && m->lon2d.ensure(m->nc) && m->olr.ensure(m->nc) && m->toa_sw.ensure(m->nc);
  if (!ok) return mfail(m, "cudaMalloc failed");
  return 0;
}

int isca_b200_moist_set_seasonal(IscaMoist m, const IscaRrtmDriverConfig* dc) {
  if (!m) return mfail(nullptr, "null handle");
  if (!dc) return mfail(m, "moist_set_seasonal: null argument");
  if (dc->abi_version != 1) return mfail(m, "IscaRrtmDriverConfig abi_version mismatch");
  if (m->initialized) return mfail(m, "moist_set_seasonal must be called before isca_b200_moist_init");
  if (m->rr) return mfail(m, "moist_set_seasonal: do_seasonal belongs to two_stream_gray_rad; RRTMG has its own zenith angle");
  if (dc->num_angles < 1 || dc->day_in_s <= 0.0 || dc->year_in_s <= 0.0) return mfail(m, "moist_set_seasonal: bad astronomy / calendar values");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  m->sdc = *dc;
  if (m->sdc.dt_rad_avg <= 0) m->sdc.dt_rad_avg = (int)(long)v.dt_atmos;      // two_stream_gray_rad_init :207
  m->s_orb = isca_rrtm_orbit(*dc);
  if (!m->coszen.ensure(m->nc) || !m->lon2d.ensure(m->nc) || !m->phy->insol.ensure(m->nc)) return mfail(m, "cudaMalloc failed");
  m->phy->pc.insol_dev = m->phy->insol.p;
  m->seasonal = true;
  return 0;
}

int isca_b200_moist_set_ozone(IscaMoist m, const double* o3) {
  if (!m) return mfail(nullptr, "null handle");
  if (!m->rr) return mfail(m, "moist_set_ozone: do_rrtm_radiation is not enabled");
  if (!o3) { m->have_o3 = false; return 0; }
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return mfail(m, isca_b200_last_error(m->dyn));
  if (!m->o3.ensure(m->n3)) return mfail(m, "cudaMalloc failed");
  MCK(cudaMemcpyAsync(m->o3.p, o3, m->n3 * sizeof(double), cudaMemcpyHostToDevice, v.st));
  MCK(cudaStreamSynchronize(v.st));
  m->have_o3 = true;
  return 0;
}

int isca_b200_moist_set_time(IscaMoist m, long long days, int seconds) {
  if (!m) return mfail(nullptr, "null handle");
  if (days < 0 || seconds < 0) return mfail(m, "moist_set_time: negative time");
  m->time_s = (double)days * 86400.0 + (double)seconds;
  return 0;
}

}  // extern "C"
