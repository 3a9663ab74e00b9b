// physics_mo.cuh -- Monin-Obukhov similarity functions for one point (monin_obukhov_kernel.F90:35-868); shared by the
// surface-flux and the diffusivity kernels.
#pragma once
#include "physics_common.h"

namespace isca_phys {

__device__ __forceinline__ double mo_phi_stable(const MoConst& c, double zeta) {
  const double b_stab = 1.0 / c.rich_crit;
  if (c.stable_option == 1) return 1.0 + zeta * (5.0 + b_stab * zeta) / (1.0 + zeta);
  const double lambda = 1.0 + (5.0 - b_stab) * c.zeta_trans;
  return zeta < c.zeta_trans ? 1.0 + 5.0 * zeta : lambda + b_stab * zeta;
}
// monin_obukhov_derivative_m :456-494
__device__ __forceinline__ double mo_phi_m(const MoConst& c, double zeta) {
  if (zeta < 0.0) return sqrt(1.0 / sqrt(1.0 - 16.0 * zeta));
  return mo_phi_stable(c, zeta);
}
// monin_obukhov_derivative_t :415-452
__device__ __forceinline__ double mo_phi_t(const MoConst& c, double zeta) {
  if (zeta < 0.0) return 1.0 / sqrt(1.0 - 16.0 * zeta);
  return mo_phi_stable(c, zeta);
}
__device__ __forceinline__ double mo_psi_stable(const MoConst& c, double ln, double zeta, double zeta_0) {
  const double b_stab = 1.0 / c.rich_crit;
  if (c.stable_option == 1) return ln + (5.0 - b_stab) * log((1.0 + zeta) / (1.0 + zeta_0)) + b_stab * (zeta - zeta_0);
  const double lambda = 1.0 + (5.0 - b_stab) * c.zeta_trans;
  if (zeta <= c.zeta_trans) return ln + 5.0 * (zeta - zeta_0);
  double x = (lambda - 1.0) * log(zeta / c.zeta_trans) + b_stab * (zeta - c.zeta_trans);
  if (zeta_0 <= c.zeta_trans) return ln + x + 5.0 * (c.zeta_trans - zeta_0);
  return lambda * ln + b_stab * (zeta - zeta_0);
}
// monin_obukhov_integral_m :644-715
__device__ __forceinline__ double mo_psi_m(const MoConst& c, double zeta, double zeta_0, double ln_z_z0) {
  if (zeta < 0.0) {
    double x = sqrt(sqrt(1.0 - 16.0 * zeta)), x_0 = sqrt(sqrt(1.0 - 16.0 * zeta_0));
    double x1 = 1.0 + x, x1_0 = 1.0 + x_0;
    double num = x1 * x1 * (1.0 + x * x), denom = x1_0 * x1_0 * (1.0 + x_0 * x_0);
    double y = atan(x) - atan(x_0);
    return ln_z_z0 - log(num / denom) + 2.0 * y;
  }
  return mo_psi_stable(c, ln_z_z0, zeta, zeta_0);
}
// monin_obukhov_integral_tq :719-806
__device__ __forceinline__ void mo_psi_tq(const MoConst& c, double zeta, double zeta_t, double zeta_q, double ln_z_zt, double ln_z_zq,
                                          double& psi_t, double& psi_q) {
  if (zeta < 0.0) {
    double x = sqrt(1.0 - 16.0 * zeta), x_t = sqrt(1.0 - 16.0 * zeta_t), x_q = sqrt(1.0 - 16.0 * zeta_q);
    psi_t = ln_z_zt - 2.0 * log((1.0 + x) / (1.0 + x_t));
    psi_q = ln_z_zq - 2.0 * log((1.0 + x) / (1.0 + x_q));
    return;
  }
  psi_t = mo_psi_stable(c, ln_z_zt, zeta, zeta_t);
  psi_q = mo_psi_stable(c, ln_z_zq, zeta, zeta_q);
}

// monin_obukhov_solve_zeta :245-411 for one point
__device__ inline void mo_solve_zeta(const MoConst& c, double rich, double z, double z0, double zt, double zq, double& f_m, double& f_t, double& f_q) {
  const double error = 1.0e-04, zeta_min = 1.0e-06;
  const int max_iter = 20;
  double z_z0 = z / z0, z_zt = z / zt, z_zq = z / zq;
  double ln_z_z0 = log(z_z0), ln_z_zt = log(z_zt), ln_z_zq = log(z_zq);
  double zeta = rich * ln_z_z0 * ln_z_z0 / ln_z_zt;
  if (rich >= 0.0) zeta = zeta / (1.0 - rich / c.rich_crit);
  f_m = 0.0; f_t = 0.0; f_q = 0.0;
  for (int iter = 0; iter < max_iter; ++iter) {
    if (fabs(zeta) < zeta_min) { f_m = ln_z_z0; f_t = ln_z_zt; f_q = ln_z_zq; return; }
    double rzeta = 1.0 / zeta;
    double zeta_0 = zeta / z_z0, zeta_t = zeta / z_zt, zeta_q = zeta / z_zq;
    double phi_m = mo_phi_m(c, zeta), phi_m_0 = mo_phi_m(c, zeta_0);
    double phi_t = mo_phi_t(c, zeta), phi_t_0 = mo_phi_t(c, zeta_t);
    f_m = mo_psi_m(c, zeta, zeta_0, ln_z_z0);
    mo_psi_tq(c, zeta, zeta_t, zeta_q, ln_z_zt, ln_z_zq, f_t, f_q);
    double df_m = (phi_m - phi_m_0) * rzeta;
    double df_t = (phi_t - phi_t_0) * rzeta;
    double rich_1 = zeta * f_t / (f_m * f_m);
    double d_rich = rich_1 * (rzeta + df_t / f_t - 2.0 * df_m / f_m);
    double correction = (rich - rich_1) / d_rich;
    double corr = fmin(fabs(correction), fabs(correction / zeta));
    if (corr > error) zeta = zeta + correction;       // NaN corrections stop the point, as `corr > error` is false
    else return;
  }
}

// monin_obukhov_drag_1d :122-241 for one point
__device__ inline void mo_drag_point(const MoConst& c, double pt, double pt0, double z, double z0, double zt, double zq, double speed,
                              double& drag_m, double& drag_t, double& drag_q, double& u_star, double& b_star) {
  const double small = 1.0e-04;
  const double r_crit = 0.95 * c.rich_crit;
  const double sqrt_drag_min = c.drag_min != 0.0 ? sqrt(c.drag_min) : 0.0;
  double delta_b = c.grav * (pt0 - pt) / pt0;
  double rich = -z * delta_b / (speed * speed + small);
  double zz = fmax(fmax(z, z0), fmax(zt, zq));
  double us, bs, qs;
  if (c.neutral) {
    us = c.vonkarm / log(zz / z0); bs = c.vonkarm / log(zz / zt); qs = c.vonkarm / log(zz / zq);
    drag_m = us * us; drag_t = us * bs; drag_q = us * qs;
  } else if (rich >= r_crit) {
    drag_m = drag_t = drag_q = c.drag_min;
    us = bs = sqrt_drag_min;
  } else {
    double fm, ft, fq;
    mo_solve_zeta(c, rich, zz, z0, zt, zq, fm, ft, fq);
    us = fmax(c.vonkarm / fm, sqrt_drag_min);
    bs = fmax(c.vonkarm / ft, sqrt_drag_min);
    qs = fmax(c.vonkarm / fq, sqrt_drag_min);
    drag_m = us * us; drag_t = us * bs; drag_q = us * qs;
  }
  u_star = us * speed;
  b_star = bs * delta_b;
}

// monin_obukhov_profile_1d :498-640 for one point
__device__ inline void mo_profile_point(const MoConst& c, double zref, double zref_t, double z, double z0, double zt, double zq, double u_star,
                                 double b_star, double& del_m, double& del_t, double& del_q) {
  double ln_z_z0 = log(z / z0), ln_z_zt = log(z / zt), ln_z_zq = log(z / zq), ln_z_zref = log(z / zref), ln_z_zref_t = log(z / zref_t);
  if (c.neutral) {
    del_m = 1.0 - ln_z_zref / ln_z_z0; del_t = 1.0 - ln_z_zref_t / ln_z_zt; del_q = 1.0 - ln_z_zref_t / ln_z_zq;
    return;
  }
  double mo_length_inv = u_star > 0.0 ? -c.vonkarm * b_star / (u_star * u_star) : 0.0;
  double zeta = z * mo_length_inv, zeta_0 = z0 * mo_length_inv, zeta_t = zt * mo_length_inv, zeta_q = zq * mo_length_inv;
  double zeta_ref = zref * mo_length_inv, zeta_ref_t = zref_t * mo_length_inv;
  double f_m = mo_psi_m(c, zeta, zeta_0, ln_z_z0);
  double f_m_ref = mo_psi_m(c, zeta, zeta_ref, ln_z_zref);
  double f_t, f_q, f_t_ref, f_q_ref;
  mo_psi_tq(c, zeta, zeta_t, zeta_q, ln_z_zt, ln_z_zq, f_t, f_q);
  mo_psi_tq(c, zeta, zeta_ref_t, zeta_ref_t, ln_z_zref_t, ln_z_zref_t, f_t_ref, f_q_ref);
  del_m = 1.0 - f_m_ref / f_m; del_t = 1.0 - f_t_ref / f_t; del_q = 1.0 - f_q_ref / f_q;
}

// monin_obukhov_stable_mix :810-868 for one point
__device__ inline double mo_stable_mix_point(const MoConst& c, double rich) {
  const double b_stab = 1.0 / c.rich_crit;
  if (c.stable_option == 1) {
    if (!(rich > 0.0 && rich < c.rich_crit)) return 0.0;
    double r = 1.0 / rich, a = r - b_stab, b = r - (1.0 + 5.0), cc = -1.0;
    double zeta = (-b + sqrt(b * b - 4.0 * a * cc)) / (2.0 * a);
    double phi = 1.0 + b_stab * zeta + (5.0 - b_stab) * zeta / (1.0 + zeta);
    return 1.0 / (phi * phi);
  }
  const double rich_trans = c.zeta_trans / (1.0 + 5.0 * c.zeta_trans), lambda = 1.0 + (5.0 - b_stab) * c.zeta_trans;
  if (rich > 0.0 && rich <= rich_trans) { double m = 1.0 - 5.0 * rich; return m * m; }
  if (rich > rich_trans && rich < c.rich_crit) { double m = (1.0 - b_stab * rich) / lambda; return m * m; }
  return 0.0;
}

// monin_obukhov_diff :35-118 for one (point, level)
__device__ inline void mo_diff_point(const MoConst& c, double z, double u_star, double b_star, double& k_m, double& k_h) {
  const double ustar_min = 1.0e-10;
  double uss = fmax(u_star, ustar_min);
  if (c.neutral) { k_m = c.vonkarm * uss * z; k_h = k_m; return; }
  double zeta = -c.vonkarm * b_star * z / (uss * uss);
  k_m = c.vonkarm * uss * z / mo_phi_m(c, zeta);
  k_h = c.vonkarm * uss * z / mo_phi_t(c, zeta);
}

inline MoConst mo_const(IscaPhysics p) {
  MoConst c;
  c.rich_crit = p->cfg.rich_crit; c.drag_min = p->cfg.drag_min; c.zeta_trans = p->cfg.zeta_trans; c.vonkarm = p->cfg.vonkarm;
  c.grav = p->cfg.grav; c.neutral = p->cfg.neutral; c.stable_option = p->cfg.stable_option;
  return c;
}

}  // namespace isca_phys
