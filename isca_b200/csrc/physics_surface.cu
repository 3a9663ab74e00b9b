// physics_surface.cu -- Monin-Obukhov similarity and the bulk surface fluxes, one thread per surface point.
//
//   monin_obukhov_drag_1d / solve_zeta / derivative_m,t / integral_m,tq / profile_1d / stable_mix / diff
//                         atmos_param/monin_obukhov/monin_obukhov_kernel.F90:35-868
//   surface_flux_1d       coupler/surface_flux.F90:338-700 (bucket = .false., all points available)
//
// The reference iterates whole rows until the slowest point has converged but freezes each point once its own correction
// is below `error`, so the result is a per-point Newton iteration -- which is what each thread runs here.
#include "physics_mo.cuh"

using namespace isca_phys;

namespace {

__global__ void mo_drag_kernel(MoConst c, int n, const double* pt, const double* pt0, const double* z, const double* z0, const double* zt,
                               const double* zq, const double* speed, double* drag_m, double* drag_t, double* drag_q, double* u_star, double* b_star) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mo_drag_point(c, pt[i], pt0[i], z[i], z0[i], zt[i], zq[i], speed[i], drag_m[i], drag_t[i], drag_q[i], u_star[i], b_star[i]);
}
__global__ void mo_profile_kernel(MoConst c, int n, double zref, double zref_t, const double* z, const double* z0, const double* zt,
                                  const double* zq, const double* u_star, const double* b_star, double* del_m, double* del_t, double* del_q) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mo_profile_point(c, zref, zref_t, z[i], z0[i], zt[i], zq[i], u_star[i], b_star[i], del_m[i], del_t[i], del_q[i]);
}
__global__ void stable_mix_kernel(MoConst c, int n, const double* rich, double* mix) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mix[i] = mo_stable_mix_point(c, rich[i]);
}
__global__ void mo_diff_kernel(MoConst c, int n, int nk, const double* z, const double* u_star, const double* b_star, double* k_m, double* k_h) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double us = u_star[i], bs = b_star[i];
  for (int k = 0; k < nk; ++k) { size_t o = (size_t)k * n + i; mo_diff_point(c, z[o], us, bs, k_m[o], k_h[o]); }
}

struct SfConst {
  double rdgas, rvgas, cp_air, stefan;
  int no_neg_q, use_virtual_temp, alt_gustiness, old_dtaudv, use_mixing_ratio, do_simple;
  double gust_const, gust_min, land_humidity_prefactor, land_evap_prefactor;
};

// bytes/point: 17 inputs + 29 outputs = 46 * 8 (one 2-D sweep; negligible next to the 3-D kernels)
__global__ void __launch_bounds__(128, ISCA_SF_MINB) surface_flux_kernel(SvpDev s, MoConst mc, SfConst c, int n, IscaSurfaceFluxArgs a, int* err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double del_temp = 0.1, del_temp_inv = 1.0 / del_temp;
  const double d622 = c.rdgas / c.rvgas, d378 = 1.0 - d622, d608 = c.use_virtual_temp ? d378 / d622 : 0.0, kappa = c.rdgas / c.cp_air;
  const bool land = a.land[i] != 0;
  double t_surf = a.t_surf[i], p_surf = a.p_surf[i];
  double t_surf0 = land ? a.t_ca[i] : t_surf;
  double t_surf1 = t_surf0 + del_temp;
  double e_sat, e_sat1, des;
  bool ok = svp_lookup(s, t_surf0, e_sat, des);
  ok &= svp_lookup(s, t_surf1, e_sat1, des);
  double q_sat, q_sat1;
  if (c.use_mixing_ratio) { q_sat = d622 * e_sat / (p_surf - e_sat); q_sat1 = d622 * e_sat1 / (p_surf - e_sat1); }
  else if (c.do_simple) { q_sat = d622 * e_sat / p_surf; q_sat1 = d622 * e_sat1 / p_surf; }
  else { q_sat = d622 * e_sat / (p_surf - d378 * e_sat); q_sat1 = d622 * e_sat1 / (p_surf - d378 * e_sat1); }
  double q_surf0 = q_sat;
  double q_atm = a.q_atm[i];
  if (c.no_neg_q && q_atm < 0.0) q_atm = 0.0;
  double t_atm = a.t_atm[i], p_atm = a.p_atm[i], z_atm = a.z_atm[i], u_atm = a.u_atm[i], v_atm = a.v_atm[i];
  double p_ratio = pow(p_surf / p_atm, kappa);
  double tv_atm = t_atm * (1.0 + d608 * q_atm);
  double th_atm = t_atm * p_ratio;
  double thv_atm = tv_atm * p_ratio;
  double thv_surf = t_surf0 * (1.0 + d608 * q_surf0);
  double u_dif = a.u_surf[i] - u_atm, v_dif = a.v_surf[i] - v_atm;
  double w_atm, dw_atmdu, dw_atmdv;
  if (c.alt_gustiness) {
    w_atm = fmax(sqrt(u_dif * u_dif + v_dif * v_dif), c.gust_const);
    if (w_atm > c.gust_const) { dw_atmdu = u_dif / w_atm; dw_atmdv = v_dif / w_atm; } else { dw_atmdu = 0.0; dw_atmdv = 0.0; }
  } else {
    double w_gust = a.gust[i];
    if (c.gust_min > 0.0) w_gust = fmax(w_gust, c.gust_min);
    w_atm = sqrt(u_dif * u_dif + v_dif * v_dif + w_gust * w_gust);
    dw_atmdu = u_dif / w_atm; dw_atmdv = v_dif / w_atm;
  }
  double rough_mom = a.rough_mom[i], rough_heat = a.rough_heat[i], rough_moist = a.rough_moist[i];
  double cd_m, cd_t, cd_q, u_star, b_star;
  mo_drag_point(mc, thv_atm, thv_surf, z_atm, rough_mom, rough_heat, rough_moist, w_atm, cd_m, cd_t, cd_q, u_star, b_star);
  double ex_del_m, ex_del_h, ex_del_q;
  mo_profile_point(mc, 10.0, 2.0, z_atm, rough_mom, rough_heat, rough_moist, u_star, b_star, ex_del_m, ex_del_h, ex_del_q);
  double q_surf_in = a.q_surf[i];
  double temp_2m = t_surf + (t_atm - t_surf) * ex_del_h;
  double q_2m = q_surf_in + (q_atm - q_surf_in) * ex_del_q;
  double e_sat_2m;
  ok &= svp_lookup(s, temp_2m, e_sat_2m, des);
  double q_sat_2m;
  if (c.use_mixing_ratio) q_sat_2m = d622 * e_sat_2m / (p_surf - e_sat_2m);
  else if (c.do_simple) q_sat_2m = d622 * e_sat_2m / p_surf;
  else q_sat_2m = d622 * e_sat_2m / (p_surf - d378 * e_sat);          // e_sat, not e_sat_2m: as the reference (:550)
  double lr = log(z_atm / rough_mom + 1.0) / log(z_atm / a.rough_scale[i] + 1.0);
  cd_m = cd_m * (lr * lr);
  double drag_t = cd_t * w_atm, drag_q = cd_q * w_atm, drag_m = cd_m * w_atm;
  double rho = p_atm / (c.rdgas * tv_atm);
  double rho_drag = c.cp_air * drag_t * rho;
  double flux_t = rho_drag * (t_surf0 - th_atm);
  a.flux_t[i] = flux_t; a.dhdt_surf[i] = rho_drag; a.dhdt_atm[i] = -rho_drag * p_ratio;
  rho_drag = drag_q * rho;
  double flux_q, dedt_surf;
  if (land) {
    flux_q = rho_drag * c.land_evap_prefactor * (c.land_humidity_prefactor * q_surf0 - q_atm);
    dedt_surf = rho_drag * c.land_evap_prefactor * (c.land_humidity_prefactor * q_sat1 - q_sat) * del_temp_inv;
  } else {
    flux_q = rho_drag * (q_surf0 - q_atm);
    dedt_surf = rho_drag * (q_sat1 - q_sat) * del_temp_inv;
  }
  a.flux_q[i] = flux_q; a.dedt_surf[i] = dedt_surf; a.dedq_surf[i] = 0.0; a.dedq_atm[i] = -rho_drag;
  a.q_star[i] = flux_q / (u_star * rho);
  a.q_surf[i] = q_atm + flux_q / (rho * cd_q * w_atm);
  double ts2 = t_surf * t_surf;
  a.flux_r[i] = c.stefan * (ts2 * ts2);
  a.drdt_surf[i] = 4.0 * c.stefan * (ts2 * t_surf);
  rho_drag = drag_m * rho;
  a.flux_u[i] = rho_drag * u_dif; a.flux_v[i] = rho_drag * v_dif;
  if (c.old_dtaudv) { a.dtaudu_atm[i] = -rho_drag; a.dtaudv_atm[i] = -rho_drag; }
  else {
    a.dtaudu_atm[i] = -cd_m * rho * (dw_atmdu * u_dif + w_atm);
    a.dtaudv_atm[i] = -cd_m * rho * (dw_atmdv * v_dif + w_atm);
  }
  a.cd_m[i] = cd_m; a.cd_t[i] = cd_t; a.cd_q[i] = cd_q; a.w_atm[i] = w_atm; a.u_star[i] = u_star; a.b_star[i] = b_star;
  a.ex_del_m[i] = ex_del_m; a.ex_del_h[i] = ex_del_h; a.ex_del_q[i] = ex_del_q;
  a.temp_2m[i] = temp_2m; a.u_10m[i] = u_atm * ex_del_m; a.v_10m[i] = v_atm * ex_del_m; a.q_2m[i] = q_2m; a.rh_2m[i] = q_2m / q_sat_2m;
  if (!ok) atomicExch(err, 1);
}

}  // namespace

namespace isca_phys {
void launch_surface_flux(IscaPhysics p, const IscaSurfaceFluxArgs& dev) {
  SfConst c;
  c.rdgas = p->cfg.rdgas; c.rvgas = p->cfg.rvgas; c.cp_air = p->cfg.cp_air; c.stefan = p->cfg.stefan;
  c.no_neg_q = p->cfg.no_neg_q; c.use_virtual_temp = p->cfg.use_virtual_temp; c.alt_gustiness = p->cfg.alt_gustiness;
  c.old_dtaudv = p->cfg.old_dtaudv; c.use_mixing_ratio = p->cfg.use_mixing_ratio; c.do_simple = p->cfg.surface_flux_do_simple;
  c.gust_const = p->cfg.gust_const; c.gust_min = p->cfg.gust_min; c.land_humidity_prefactor = p->cfg.land_humidity_prefactor;
  c.land_evap_prefactor = p->cfg.land_evap_prefactor;
  surface_flux_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>(p->svp, mo_const(p), c, (int)p->ncol, dev, p->d_err);
}
}  // namespace isca_phys

extern "C" {

int isca_b200_mo_drag(IscaPhysics p, int n, const double* pt, const double* pt0, const double* z, const double* z0,
                      const double* zt, const double* zq, const double* speed, double* drag_m, double* drag_t,
                      double* drag_q, double* u_star, double* b_star) {
  if (!p) return fail(nullptr, "null handle");
  if (n <= 0) return 0;
  Dev* b = p->buf;
  const double* in[7] = {pt, pt0, z, z0, zt, zq, speed};
  double* out[5] = {drag_m, drag_t, drag_q, u_star, b_star};
  for (int i = 0; i < 7; ++i) if (up(p, b[i], in[i], n)) return 1;
  for (int i = 0; i < 5; ++i) if (!b[7 + i].ensure(n)) return fail(p, "cudaMalloc failed");
  mo_drag_kernel<<<(n + 127) / 128, 128, 0, p->st>>>(mo_const(p), n, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p, b[5].p, b[6].p, b[7].p, b[8].p,
                                                     b[9].p, b[10].p, b[11].p);
  for (int i = 0; i < 5; ++i) if (down(p, b[7 + i], out[i], n)) return 1;
  return finish(p, "mo_drag");
}

int isca_b200_mo_profile(IscaPhysics p, int n, double zref, double zref_t, const double* z, const double* z0,
                         const double* zt, const double* zq, const double* u_star, const double* b_star,
                         double* del_m, double* del_t, double* del_q) {
  if (!p) return fail(nullptr, "null handle");
  if (n <= 0) return 0;
  Dev* b = p->buf;
  const double* in[6] = {z, z0, zt, zq, u_star, b_star};
  double* out[3] = {del_m, del_t, del_q};
  for (int i = 0; i < 6; ++i) if (up(p, b[i], in[i], n)) return 1;
  for (int i = 0; i < 3; ++i) if (!b[6 + i].ensure(n)) return fail(p, "cudaMalloc failed");
  mo_profile_kernel<<<(n + 127) / 128, 128, 0, p->st>>>(mo_const(p), n, zref, zref_t, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p, b[5].p,
                                                        b[6].p, b[7].p, b[8].p);
  for (int i = 0; i < 3; ++i) if (down(p, b[6 + i], out[i], n)) return 1;
  return finish(p, "mo_profile");
}

int isca_b200_stable_mix(IscaPhysics p, int n, const double* rich, double* mix) {
  if (!p) return fail(nullptr, "null handle");
  if (n <= 0) return 0;
  if (up(p, p->buf[0], rich, n)) return 1;
  if (!p->buf[1].ensure(n)) return fail(p, "cudaMalloc failed");
  stable_mix_kernel<<<(n + 255) / 256, 256, 0, p->st>>>(mo_const(p), n, p->buf[0].p, p->buf[1].p);
  if (down(p, p->buf[1], mix, n)) return 1;
  return finish(p, "stable_mix");
}

int isca_b200_mo_diff(IscaPhysics p, int n, int nk, const double* z, const double* u_star, const double* b_star,
                      double* k_m, double* k_h) {
  if (!p) return fail(nullptr, "null handle");
  if (n <= 0 || nk <= 0) return 0;
  size_t n3 = (size_t)n * nk;
  Dev* b = p->buf;
  if (up(p, b[0], z, n3) || up(p, b[1], u_star, n) || up(p, b[2], b_star, n)) return 1;
  if (!b[3].ensure(n3) || !b[4].ensure(n3)) return fail(p, "cudaMalloc failed");
  mo_diff_kernel<<<(n + 127) / 128, 128, 0, p->st>>>(mo_const(p), n, nk, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p);
  if (down(p, b[3], k_m, n3) || down(p, b[4], k_h, n3)) return 1;
  return finish(p, "mo_diff");
}

int isca_b200_surface_flux(IscaPhysics p, const IscaSurfaceFluxArgs* a) {
  if (!p) return fail(nullptr, "null handle");
  if (!a) return fail(p, "null argument");
  size_t nc = p->ncol;
  const double* in[16] = {a->t_atm, a->q_atm, a->u_atm, a->v_atm, a->p_atm, a->z_atm, a->p_surf, a->t_surf, a->t_ca, a->u_surf, a->v_surf,
                          a->rough_mom, a->rough_heat, a->rough_moist, a->rough_scale, a->gust};
  double* out[28] = {a->flux_t, a->flux_q, a->flux_r, a->flux_u, a->flux_v, a->cd_m, a->cd_t, a->cd_q, a->w_atm, a->u_star, a->b_star,
                     a->q_star, a->dhdt_surf, a->dedt_surf, a->dedq_surf, a->drdt_surf, a->dhdt_atm, a->dedq_atm, a->dtaudu_atm,
                     a->dtaudv_atm, a->ex_del_m, a->ex_del_h, a->ex_del_q, a->temp_2m, a->u_10m, a->v_10m, a->q_2m, a->rh_2m};
  if (!a->land || !a->q_surf) return fail(p, "null input array");
  for (int i = 0; i < 28; ++i) if (!out[i]) return fail(p, "null output array");
  // one staging block: 16 inputs | q_surf | 28 outputs | land (ints, in a double-sized slot)
  Dev& blk = p->buf[0];
  if (!blk.ensure(46 * nc)) return fail(p, "cudaMalloc failed");
  for (int i = 0; i < 16; ++i) {
    if (!in[i]) return fail(p, "null input array");
    PCK(cudaMemcpyAsync(blk.p + i * nc, in[i], nc * sizeof(double), cudaMemcpyHostToDevice, p->st));
  }
  PCK(cudaMemcpyAsync(blk.p + 16 * nc, a->q_surf, nc * sizeof(double), cudaMemcpyHostToDevice, p->st));
  PCK(cudaMemcpyAsync(blk.p + 45 * nc, a->land, nc * sizeof(int), cudaMemcpyHostToDevice, p->st));
  IscaSurfaceFluxArgs d;
  const double** din[16] = {&d.t_atm, &d.q_atm, &d.u_atm, &d.v_atm, &d.p_atm, &d.z_atm, &d.p_surf, &d.t_surf, &d.t_ca, &d.u_surf, &d.v_surf,
                            &d.rough_mom, &d.rough_heat, &d.rough_moist, &d.rough_scale, &d.gust};
  double** dout[28] = {&d.flux_t, &d.flux_q, &d.flux_r, &d.flux_u, &d.flux_v, &d.cd_m, &d.cd_t, &d.cd_q, &d.w_atm, &d.u_star, &d.b_star,
                       &d.q_star, &d.dhdt_surf, &d.dedt_surf, &d.dedq_surf, &d.drdt_surf, &d.dhdt_atm, &d.dedq_atm, &d.dtaudu_atm,
                       &d.dtaudv_atm, &d.ex_del_m, &d.ex_del_h, &d.ex_del_q, &d.temp_2m, &d.u_10m, &d.v_10m, &d.q_2m, &d.rh_2m};
  for (int i = 0; i < 16; ++i) *din[i] = blk.p + i * nc;
  d.q_surf = blk.p + 16 * nc;
  for (int i = 0; i < 28; ++i) *dout[i] = blk.p + (17 + i) * nc;
  d.land = reinterpret_cast<const int*>(blk.p + 45 * nc);
  launch_surface_flux(p, d);
  PCK(cudaMemcpyAsync(a->q_surf, d.q_surf, nc * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  for (int i = 0; i < 28; ++i) PCK(cudaMemcpyAsync(out[i], *dout[i], nc * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  return finish(p, "surface_flux");
}

}  // extern "C"
