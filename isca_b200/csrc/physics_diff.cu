// physics_diff.cu -- implicit vertical diffusion split around the surface (vert_diff_mod) and the slab mixed layer.
//
//   gcm_vert_diff_down / uv_vert_diff / vert_diff_down_2 / compute_e,f,mu,nu / explicit_tend / diff_surface / vert_diff_up
//                         atmos_param/vert_diff/vert_diff.F90:270-467, 556-617, 806-1087
//   mixed_layer           atmos_spectral/driver/solo/mixed_layer.F90:568-745
//
// One thread per column, loads software-pipelined three levels ahead of the serial recurrences.  The momentum system is eliminated, closed by the surface stress and back-substituted inside the
// kernel (its factors never leave thread-local storage); the temperature / humidity elimination stores e, f_t, f_q (the
// reference's module state) for gcm_vert_diff_up, which runs after the surface has been updated.
#include "physics_common.h"

using namespace isca_phys;

namespace {

struct Surf { double mu_delt_n, nu_n, e_n1, f1_delt_n1, f2_delt_n1, delta1_n, delta2_n; };

// diff_surface (vert_diff.F90:866-888)
__device__ __forceinline__ void diff_surface(double mu_delt, double nu, double e_n1, double f_delt_n1, double dflux_datmos,
                                             double& flux, double factor, double& delta_xi) {
  double fff = 1.0 / factor;
  double dflux = -nu * (1.0 - e_n1);
  delta_xi = delta_xi + mu_delt * nu * f_delt_n1;
  delta_xi = (delta_xi + mu_delt * flux * fff) / (1.0 - mu_delt * (dflux + dflux_datmos * fff));
  flux = flux + dflux_datmos * delta_xi;
}

struct DiffArgs {
  int ncol, K; double delt, grav, rdgas, cp_air, d608; int conserve, use_virtual;
  const double *u, *v, *t, *q, *diff_m, *diff_t, *p_half, *z_full, *dtau_du, *dtau_dv, *dt_q;
  double *tau_u, *tau_v, *dt_u, *dt_v, *dt_t, *diss;
  double *e_g, *ft_g, *fq_g, *tri_delta_t, *tri_dflux_t, *tri_delta_q, *tri_dflux_q, *tri_dtmass, *tri_delta_u, *tri_delta_v;
};

// Raw inputs of one level of a down sweep.  The sweeps are latency-bound (one thread per column, a serial recurrence over the
// levels), so the loads of level k+3 are issued while level k is eliminated: three levels of raw values travel in registers.
struct DiffRaw { double ph, t, q, df, z, x1, x2, d1, d2; };

// vert_diff_down_2 (explicit_tend + compute_e + compute_f + compute_mu + compute_nu) for two fields sharing mu, nu, software
// pipelined.  TEMP: the fields are (t + z*grav/cp, q) instead of (X1, X2).
template <bool TEMP>
__device__ __forceinline__ Surf down2_pipe(const DiffArgs& a, int col, const double* __restrict__ diff, const double* __restrict__ X1,
                                           const double* __restrict__ X2, const double* __restrict__ D1, const double* __restrict__ D2) {
  const int K = a.K; const size_t nc = a.ncol;
  const double delt = a.delt, gcp = a.grav / a.cp_air;
  auto load = [&](int kk) {
    DiffRaw r; r.ph = r.t = r.q = r.df = r.z = r.x1 = r.x2 = r.d1 = r.d2 = 0.0;
    if (kk <= K) {
      const size_t o = (size_t)kk * nc + col;
      r.ph = a.p_half[o];
      if (kk < K) {
        r.t = a.t[o]; r.df = diff[o]; r.z = a.z_full[o]; r.d1 = D1[o]; r.d2 = D2[o];
        if (TEMP || a.use_virtual) r.q = a.q[o];
        if (!TEMP) { r.x1 = X1[o]; r.x2 = X2[o]; }
      }
    }
    return r;
  };
  auto tv_of = [&](const DiffRaw& r) { double tv = r.t; if (a.use_virtual) tv = tv * (1.0 + a.d608 * r.q); return tv; };
  auto x1_of = [&](const DiffRaw& r) { return TEMP ? r.t + r.z * gcp : r.x1; };
  auto x2_of = [&](const DiffRaw& r) { return TEMP ? r.q : r.x2; };
  Surf s;
  DiffRaw r0 = load(0), r1 = load(1), r2 = load(2);
  double x1 = x1_of(r0), x2 = x2_of(r0);
  double fl1 = 0.0, fl2 = 0.0;                    // fluxx(k)
  double e_prev = 0.0, f1_prev = 0.0, f2_prev = 0.0;
  double nu_k = 0.0;                              // nu(1) is never referenced by the reference (c(1) = 0)
  for (int k = 0; k < K; ++k) {
    const DiffRaw r3 = load(k + 3);
    const double mu_k = a.grav / (r1.ph - r0.ph);
    double nu_k1 = 0.0;
    if (k < K - 1) {
      const double rho_half = 2.0 * r1.ph / (a.rdgas * (tv_of(r1) + tv_of(r0)));
      nu_k1 = rho_half * r1.df / (r0.z - r1.z);
    }
    double d1, d2, aa = 0.0, x1p = 0.0, x2p = 0.0, fl1p = 0.0, fl2p = 0.0;
    if (k < K - 1) {
      x1p = x1_of(r1); x2p = x2_of(r1);
      fl1p = nu_k1 * (x1p - x1); fl2p = nu_k1 * (x2p - x2);
      d1 = r0.d1 + mu_k * (fl1p - fl1);
      d2 = r0.d2 + mu_k * (fl2p - fl2);
      aa = -mu_k * nu_k1 * delt;
    } else {
      d1 = r0.d1 - mu_k * fl1;
      d2 = r0.d2 - mu_k * fl2;
    }
    double c = k > 0 ? -mu_k * nu_k * delt : 0.0;
    double b = 1.0 - aa - c;
    if (k < K - 1) {
      double e, f1, f2;
      if (k == 0) { e = -aa / b; f1 = d1 / b; f2 = d2 / b; }
      else {
        double g = 1.0 / (b + c * e_prev);
        e = -aa * g; f1 = (d1 - c * f1_prev) * g; f2 = (d2 - c * f2_prev) * g;
      }
      const size_t o = (size_t)k * nc + col;
      a.e_g[o] = e; a.ft_g[o] = f1; a.fq_g[o] = f2;
      e_prev = e; f1_prev = f1; f2_prev = f2;
    } else {
      s.mu_delt_n = mu_k * delt; s.nu_n = nu_k;
      s.e_n1 = e_prev; s.f1_delt_n1 = f1_prev * delt; s.f2_delt_n1 = f2_prev * delt;
      s.delta1_n = d1 * delt; s.delta2_n = d2 * delt;
    }
    x1 = x1p; x2 = x2p; fl1 = fl1p; fl2 = fl2p; nu_k = nu_k1;
    r0 = r1; r1 = r2; r2 = r3;
  }
  return s;
}

// bytes/column: read u,v,t,q,diff_m,diff_t,z_full,dt_u,dt_v,dt_t,dt_q (11K) + p_half (K+1) + 4; write dt_u,dt_v,dt_t,diss,
// e,f_t,f_q (7K) + 9  =  (19K + 14) * 8
__global__ void __launch_bounds__(128, ISCA_VDD_MINB) vert_diff_down_kernel(DiffArgs a) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.ncol) return;
  const int K = a.K; const size_t nc = a.ncol;
  // uv_vert_diff: the factors of the momentum system use e_global, f_t_global, f_q_global as scratch (they are rewritten by
  // the temperature / humidity elimination below)
  Surf s = down2_pipe<false>(a, col, a.diff_m, a.u, a.v, a.dt_u, a.dt_v);
  double tau_u = a.tau_u[col], tau_v = a.tau_v[col];
  double delta_u_n = s.delta1_n, delta_v_n = s.delta2_n;
  diff_surface(s.mu_delt_n, s.nu_n, s.e_n1, s.f1_delt_n1, a.dtau_du[col], tau_u, 1.0, delta_u_n);
  diff_surface(s.mu_delt_n, s.nu_n, s.e_n1, s.f2_delt_n1, a.dtau_dv[col], tau_v, 1.0, delta_v_n);
  a.tau_u[col] = tau_u; a.tau_v[col] = tau_v;
  {
    // back substitution bottom-up, loads two levels ahead of the recurrence
    struct Back { double e, f1, f2, du, dv, u, v, dtt; };
    auto loadb = [&](int kk) {
      Back r; r.e = r.f1 = r.f2 = r.du = r.dv = r.u = r.v = r.dtt = 0.0;
      if (kk >= 0) {
        const size_t o = (size_t)kk * nc + col;
        if (kk < K - 1) { r.e = a.e_g[o]; r.f1 = a.ft_g[o]; r.f2 = a.fq_g[o]; }
        if (a.conserve) { r.du = a.dt_u[o]; r.dv = a.dt_v[o]; r.u = a.u[o]; r.v = a.v[o]; r.dtt = a.dt_t[o]; }
      }
      return r;
    };
    const double half_delt = 0.5 * a.delt, cp_inv = 1.0 / a.cp_air;
    double nu_ = delta_u_n / a.delt, nv_ = delta_v_n / a.delt;
    Back b0 = loadb(K - 1), b1 = loadb(K - 2), b2 = loadb(K - 3);
    for (int k = K - 1; k >= 0; --k) {
      const Back b3 = loadb(k - 3);
      size_t o = (size_t)k * nc + col;
      if (k < K - 1) { nu_ = b0.e * nu_ + b0.f1; nv_ = b0.e * nv_ + b0.f2; }
      double heat = 0.0;
      if (a.conserve) {
        double du = nu_ - b0.du, dv = nv_ - b0.dv;
        heat = -cp_inv * ((b0.u + half_delt * du) * du + (b0.v + half_delt * dv) * dv);
        a.dt_t[o] = b0.dtt + heat;
      }
      a.dt_u[o] = nu_; a.dt_v[o] = nv_; a.diss[o] = heat;
      b0 = b1; b1 = b2; b2 = b3;
    }
  }
  // compute_nu(diff_t), vert_diff_down_2(tt, q)
  s = down2_pipe<true>(a, col, a.diff_t, nullptr, nullptr, a.dt_t, a.dt_q);
  a.tri_delta_t[col] = s.delta1_n + s.mu_delt_n * s.nu_n * s.f1_delt_n1;
  a.tri_dflux_t[col] = -s.nu_n * (1.0 - s.e_n1);
  a.tri_delta_q[col] = s.delta2_n + s.mu_delt_n * s.nu_n * s.f2_delt_n1;
  a.tri_dflux_q[col] = -s.nu_n * (1.0 - s.e_n1);
  a.tri_dtmass[col] = s.mu_delt_n;
  a.tri_delta_u[col] = delta_u_n;
  a.tri_delta_v[col] = delta_v_n;
}

// bytes/column: read e, f_t, f_q (3(K-1)) + 2, write dt_t, dt_q (2K)  ~ 5K * 8
__global__ void __launch_bounds__(128) vert_diff_up_kernel(int ncol, int K, double delt, const double* __restrict__ e,
    const double* __restrict__ ft, const double* __restrict__ fq, const double* __restrict__ delta_t, const double* __restrict__ delta_q,
    double* __restrict__ dt_t, double* __restrict__ dt_q) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  double xt = delta_t[col] / delt, xq = delta_q[col] / delt;
  size_t o = (size_t)(K - 1) * ncol + col;
  dt_t[o] = xt; dt_q[o] = xq;
  // factors of four levels are loaded before the (serial) back substitution consumes them
  for (int k0 = K - 2; k0 >= 0; k0 -= 4) {
    double ee[4], f1[4], f2[4];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (k0 - d >= 0) { const size_t oo = (size_t)(k0 - d) * ncol + col; ee[d] = e[oo]; f1[d] = ft[oo]; f2[d] = fq[oo]; }
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (k0 - d >= 0) {
        const size_t oo = (size_t)(k0 - d) * ncol + col;
        xt = ee[d] * xt + f1[d]; xq = ee[d] * xq + f2[d];
        dt_t[oo] = xt; dt_q[oo] = xq;
      }
  }
}

struct MixedArgs {
  int ncol; double dt, cp_air, hlv; int evaporation;
  const double *flux_t, *flux_q, *flux_r, *sw, *lw, *dhdt_surf, *dedt_surf, *dedq_surf, *drdt_surf, *dhdt_atm, *dedq_atm,
               *heat_cap, *qflux, *dtmass, *dflux_t, *dflux_q;
  double *t_surf, *delta_t, *delta_q, *delta_t_surf;
  const double* sst_new;             // do_sc_sst: the prescribed SST of the time stepped to; nullptr = slab ocean
};

// mixed_layer.F90:629-745; returns a non-finite increment when eff_heat_capacity == 0 (the reference aborts)
__global__ void mixed_layer_kernel(MixedArgs a, int* err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.ncol) return;
  const double inv_cp = 1.0 / a.cp_air;
  double dtmass = a.dtmass[i], dhdt_atm = a.dhdt_atm[i], dedq_atm = a.dedq_atm[i];
  double gamma_t = 1.0 / (1.0 - dtmass * (a.dflux_t[i] + dhdt_atm * inv_cp));
  double gamma_q = 1.0 / (1.0 - dtmass * (a.dflux_q[i] + dedq_atm));
  double flux_t = a.flux_t[i], flux_q = a.flux_q[i];
  double fn_t = gamma_t * (a.delta_t[i] + dtmass * flux_t * inv_cp);
  double fn_q = gamma_q * (a.delta_q[i] + dtmass * flux_q);
  double en_t = gamma_t * dtmass * a.dhdt_surf[i] * inv_cp;
  double en_q = gamma_q * dtmass * a.dedt_surf[i];
  double alpha_t = flux_t * inv_cp + dhdt_atm * inv_cp * fn_t;
  double alpha_q = flux_q + dedq_atm * fn_q;
  double alpha_lw = a.flux_r[i];
  double beta_t = a.dhdt_surf[i] * inv_cp + dhdt_atm * inv_cp * en_t;
  double beta_q = a.dedt_surf[i] + dedq_atm * en_q;
  double beta_lw = a.drdt_surf[i];
  double corrected_flux = -a.sw[i] - a.lw[i] + alpha_t * a.cp_air + alpha_lw - a.qflux[i];
  double t_surf_dependence = beta_t * a.cp_air + beta_lw;
  if (a.evaporation) {
    corrected_flux = corrected_flux + alpha_q * a.hlv;
    t_surf_dependence = t_surf_dependence + beta_q * a.hlv;
  }
  double d;
  if (a.sst_new) {                    // do_sc_sst (mixed_layer.F90:681-691; do_calc_eff_heat_cap = .false., :495-502): no slab update
    d = a.sst_new[i] - a.t_surf[i];
  } else {
    double eff = a.heat_cap[i] + t_surf_dependence * a.dt;
    if (eff == 0.0) atomicExch(err, 2);
    d = -corrected_flux * a.dt / eff;
  }
  a.t_surf[i] = a.t_surf[i] + d;
  a.delta_t[i] = fn_t + en_t * d;
  if (a.evaporation) a.delta_q[i] = fn_q + en_q * d;
  if (a.delta_t_surf) a.delta_t_surf[i] = d;
}

int ensure_state(IscaPhysics p) {
  size_t nc = p->ncol, n3 = nc * p->K;
  for (int i = ST_E_GLOBAL; i <= ST_F_Q_GLOBAL; ++i) if (!p->state[i].ensure(n3)) return fail(p, "cudaMalloc failed");
  for (int i = ST_TRI_DELTA_T; i <= ST_TRI_DELTA_V; ++i) if (!p->state[i].ensure(nc)) return fail(p, "cudaMalloc failed");
  return 0;
}

}  // namespace

namespace isca_phys {

// device-pointer launch used by the host-array entry point and by the timing hook
void launch_vert_diff_down(IscaPhysics p, double delt, const double* u, const double* v, const double* t, const double* q,
                           const double* diff_m, const double* diff_t, const double* p_half, const double* z_full, double* tau_u,
                           double* tau_v, const double* dtau_du, const double* dtau_dv, double* dt_u, double* dt_v, double* dt_t,
                           const double* dt_q, double* diss) {
  DiffArgs a;
  a.ncol = (int)p->ncol; a.K = p->K; a.delt = delt; a.grav = p->cfg.grav; a.rdgas = p->cfg.rdgas; a.cp_air = p->cfg.cp_air;
  a.d608 = (p->cfg.rvgas - p->cfg.rdgas) / p->cfg.rdgas;
  a.conserve = p->cfg.vert_diff_do_conserve_energy; a.use_virtual = p->cfg.use_virtual_temp_vert_diff;
  a.u = u; a.v = v; a.t = t; a.q = q; a.diff_m = diff_m; a.diff_t = diff_t; a.p_half = p_half; a.z_full = z_full;
  a.dtau_du = dtau_du; a.dtau_dv = dtau_dv; a.dt_q = dt_q; a.tau_u = tau_u; a.tau_v = tau_v; a.dt_u = dt_u; a.dt_v = dt_v;
  a.dt_t = dt_t; a.diss = diss;
  a.e_g = p->state[ST_E_GLOBAL].p; a.ft_g = p->state[ST_F_T_GLOBAL].p; a.fq_g = p->state[ST_F_Q_GLOBAL].p;
  a.tri_delta_t = p->state[ST_TRI_DELTA_T].p; a.tri_dflux_t = p->state[ST_TRI_DFLUX_T].p; a.tri_delta_q = p->state[ST_TRI_DELTA_Q].p;
  a.tri_dflux_q = p->state[ST_TRI_DFLUX_Q].p; a.tri_dtmass = p->state[ST_TRI_DTMASS].p; a.tri_delta_u = p->state[ST_TRI_DELTA_U].p;
  a.tri_delta_v = p->state[ST_TRI_DELTA_V].p;
  vert_diff_down_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>(a);
}

void launch_vert_diff_up(IscaPhysics p, double delt, double* dt_t, double* dt_q) {
  vert_diff_up_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>((int)p->ncol, p->K, delt, p->state[ST_E_GLOBAL].p, p->state[ST_F_T_GLOBAL].p,
      p->state[ST_F_Q_GLOBAL].p, p->state[ST_TRI_DELTA_T].p, p->state[ST_TRI_DELTA_Q].p, dt_t, dt_q);
}

int prepare_vert_diff_state(IscaPhysics p) { return ensure_state(p); }

void launch_mixed_layer(IscaPhysics p, double dt, double* t_surf, const double* flux_t, const double* flux_q, const double* flux_r,
                        const double* net_surf_sw_down, const double* surf_lw_down, const double* dhdt_surf, const double* dedt_surf,
                        const double* dedq_surf, const double* drdt_surf, const double* dhdt_atm, const double* dedq_atm, double* delta_t_surf) {
  const size_t nc = p->ncol;
  MixedArgs a;
  a.ncol = (int)nc; a.dt = dt; a.cp_air = p->cfg.cp_air; a.hlv = p->cfg.hlv; a.evaporation = p->cfg.evaporation;
  a.t_surf = t_surf; a.flux_t = flux_t; a.flux_q = flux_q; a.flux_r = flux_r; a.sw = net_surf_sw_down; a.lw = surf_lw_down;
  a.dhdt_surf = dhdt_surf; a.dedt_surf = dedt_surf; a.dedq_surf = dedq_surf; a.drdt_surf = drdt_surf; a.dhdt_atm = dhdt_atm;
  a.dedq_atm = dedq_atm;
  a.heat_cap = p->state[ST_ML_HEAT_CAP].p; a.qflux = p->state[ST_ML_QFLUX].p; a.dtmass = p->state[ST_TRI_DTMASS].p;
  a.dflux_t = p->state[ST_TRI_DFLUX_T].p; a.dflux_q = p->state[ST_TRI_DFLUX_Q].p; a.delta_t = p->state[ST_TRI_DELTA_T].p;
  a.delta_q = p->state[ST_TRI_DELTA_Q].p; a.delta_t_surf = delta_t_surf;
  a.sst_new = p->sc_sst ? p->state[ST_ML_SST].p : nullptr;
  mixed_layer_kernel<<<(int)((nc + 255) / 256), 256, 0, p->st>>>(a, p->d_err);
}

}  // namespace isca_phys

extern "C" {

int isca_b200_gcm_vert_diff_down(IscaPhysics p, double delt, const double* u, const double* v, const double* t,
                                 const double* q, const double* diff_m, const double* diff_t, const double* p_half,
                                 const double* p_full, const double* z_full, double* tau_u, double* tau_v,
                                 const double* dtau_du, const double* dtau_dv, double* dt_u, double* dt_v,
                                 double* dt_t, const double* dt_q, double* dissipative_heat) {
  if (!p) return fail(nullptr, "null handle");
  if (p->K < 2) return fail(p, "gcm_vert_diff_down needs at least 2 levels");
  if (!p_full) return fail(p, "null input array");                 // only used by do_mcm_plev, kept for the reference argument list
  size_t nc = p->ncol, n3 = nc * p->K;
  if (ensure_state(p)) return 1;
  Dev* b = p->buf;
  if (up(p, b[0], u, n3) || up(p, b[1], v, n3) || up(p, b[2], t, n3) || up(p, b[3], q, n3) || up(p, b[4], diff_m, n3) ||
      up(p, b[5], diff_t, n3) || up(p, b[6], p_half, n3 + nc) || up(p, b[7], z_full, n3) || up(p, b[8], tau_u, nc) ||
      up(p, b[9], tau_v, nc) || up(p, b[10], dtau_du, nc) || up(p, b[11], dtau_dv, nc) || up(p, b[12], dt_u, n3) ||
      up(p, b[13], dt_v, n3) || up(p, b[14], dt_t, n3) || up(p, b[15], dt_q, n3)) return 1;
  if (!b[16].ensure(n3)) return fail(p, "cudaMalloc failed");
  launch_vert_diff_down(p, delt, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p, b[5].p, b[6].p, b[7].p, b[8].p, b[9].p, b[10].p, b[11].p,
                        b[12].p, b[13].p, b[14].p, b[15].p, b[16].p);
  if (down(p, b[8], tau_u, nc) || down(p, b[9], tau_v, nc) || down(p, b[12], dt_u, n3) || down(p, b[13], dt_v, n3) ||
      down(p, b[14], dt_t, n3) || down(p, b[16], dissipative_heat, n3)) return 1;
  if (finish(p, "gcm_vert_diff_down")) return 1;
  p->vert_diff_down_done = true;
  return 0;
}

int isca_b200_get_tri_surf(IscaPhysics p, int id, double* host) {
  if (!p) return fail(nullptr, "null handle");
  if (!p->vert_diff_down_done) return fail(p, "get_tri_surf: gcm_vert_diff_down has not been called");
  size_t nc = p->ncol, n3 = nc * p->K;
  if (id >= 0 && id <= 6) { if (down(p, p->state[ST_TRI_DELTA_T + id], host, nc)) return 1; }
  else if (id >= 16 && id <= 18) {
    // levels 0..K-2 hold factors; the last level is never written by the reference either: report zeros
    if (!host) return fail(p, "null output array");
    PCK(cudaMemsetAsync(p->state[ST_E_GLOBAL + id - 16].p + (n3 - nc), 0, nc * sizeof(double), p->st));
    if (down(p, p->state[ST_E_GLOBAL + id - 16], host, n3)) return 1;
  } else return fail(p, "get_tri_surf: unknown id");
  return finish(p, "get_tri_surf");
}

int isca_b200_mixed_layer_init(IscaPhysics p, const double* heat_capacity, const double* ocean_qflux) {
  if (!p) return fail(nullptr, "null handle");
  if (up(p, p->state[ST_ML_HEAT_CAP], heat_capacity, p->ncol) || up(p, p->state[ST_ML_QFLUX], ocean_qflux, p->ncol)) return 1;
  return finish(p, "mixed_layer_init");
}

int isca_b200_mixed_layer_set_sst(IscaPhysics p, const double* sst) {
  if (!p) return fail(nullptr, "null handle");
  if (!sst) { p->sc_sst = false; return 0; }
  if (up(p, p->state[ST_ML_SST], sst, p->ncol)) return 1;
  PCK(cudaStreamSynchronize(p->st));
  p->sc_sst = true;
  return 0;
}

int isca_b200_mixed_layer(IscaPhysics p, double dt, double* t_surf, const double* flux_t, const double* flux_q,
                          const double* flux_r, const double* net_surf_sw_down, const double* surf_lw_down,
                          const double* dhdt_surf, const double* dedt_surf, const double* dedq_surf,
                          const double* drdt_surf, const double* dhdt_atm, const double* dedq_atm, double* delta_t_surf) {
  if (!p) return fail(nullptr, "null handle");
  if (!p->vert_diff_down_done) return fail(p, "mixed_layer: Tri_surf is not defined (gcm_vert_diff_down has not been called)");
  if (!p->state[ST_ML_HEAT_CAP].p) return fail(p, "mixed_layer: mixed_layer module is not initialized");
  size_t nc = p->ncol;
  Dev* b = p->buf;
  const double* in[12] = {t_surf, flux_t, flux_q, flux_r, net_surf_sw_down, surf_lw_down, dhdt_surf, dedt_surf, dedq_surf, drdt_surf,
                          dhdt_atm, dedq_atm};
  for (int i = 0; i < 12; ++i) if (up(p, b[i], in[i], nc)) return 1;
  if (!b[12].ensure(nc)) return fail(p, "cudaMalloc failed");
  launch_mixed_layer(p, dt, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p, b[5].p, b[6].p, b[7].p, b[8].p, b[9].p, b[10].p, b[11].p, b[12].p);
  if (down(p, b[0], t_surf, nc)) return 1;
  if (delta_t_surf && down(p, b[12], delta_t_surf, nc)) return 1;
  int e = 0;
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaStreamSynchronize(p->st));
  if (e) {
    PCK(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->st));
    return fail(p, "mixed_layer: Avoiding division by zero (eff_heat_capacity == 0)");
  }
  return 0;
}

int isca_b200_gcm_vert_diff_up(IscaPhysics p, double delt, double* dt_t, double* dt_q) {
  if (!p) return fail(nullptr, "null handle");
  if (!p->vert_diff_down_done) return fail(p, "gcm_vert_diff_up: gcm_vert_diff_down has not been called");
  size_t n3 = p->ncol * p->K;
  if (!p->buf[0].ensure(n3) || !p->buf[1].ensure(n3)) return fail(p, "cudaMalloc failed");
  launch_vert_diff_up(p, delt, p->buf[0].p, p->buf[1].p);
  if (down(p, p->buf[0], dt_t, n3) || down(p, p->buf[1], dt_q, n3)) return 1;
  return finish(p, "gcm_vert_diff_up");
}

}  // extern "C"
