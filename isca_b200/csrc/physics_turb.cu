// physics_turb.cu -- K-profile boundary-layer diffusivities, one thread per column.
//
//   diffusivity / pbl_depth / diffusivity_pbl / diffusivity_entr    atmos_param/diffusivity/diffusivity.F90:263-530, 732-750
//   (the do_diffusivity branch of vert_turb_driver, vert_turb_driver.F90:277-292)
//
// With use_pog_bug_fix = .true. (the default) the reference's domain-wide min/max in diffusivity_pbl only skips work: the
// result is a function of the column alone, which is what is computed here.
#include "physics_mo.cuh"

using namespace isca_phys;

namespace {

struct TurbConst {
  double grav, cp_air, d608, vonkarm;
  int fixed_depth, do_entrain, do_simple;
  double depth_0, frac_inner, rich_crit_pbl, entr_ratio, parcel_buoy, znom, background_m, background_t;
  int free_atm_diff, free_atm_skyhi_diff, ampns;
  double rich_crit_diff, mix_len, rich_prandtl, ampns_max;
};

// bytes/column: read t, q, u, v, z_full (5K; with free_atm_diff = .false. only up to the boundary-layer top) + z_half (K+1)
// + k_m, k_t (2K, add_input) + 2; write k_m, k_t (2K) + 1  ~ (10K + 4) * 8
// tau1.tdt != NULL: vert_turb_driver_nml use_tau = .false. (vert_turb_driver.F90:209-213) -- the scheme sees the variables at time
// tau + 1, x + dt * dx/dt, formed here from the previous-level fields and the tendencies instead of by a separate pass over four 3-D
// fields.  add_input = 0: the diffusivities are written, not added to the incoming arrays (the caller would have zeroed them).
struct Tau1 { const double *tdt, *qdt, *udt, *vdt; double dt; };
__global__ void __launch_bounds__(128, ISCA_COL_MINB) diffusivity_kernel(MoConst mc, TurbConst c, int ncol, int K, const double* __restrict__ t_,
    const double* __restrict__ q_, const double* __restrict__ u_, const double* __restrict__ v_, Tau1 tau1, int add_input,
    const double* __restrict__ z_full,
    const double* __restrict__ z_half, const double* __restrict__ u_star, const double* __restrict__ b_star, double* __restrict__ h_out,
    double* __restrict__ k_m, double* __restrict__ k_t) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const double small = 1.0e-04, gcp = c.grav / c.cp_air;
  const size_t nc = ncol;
  const bool t1 = tau1.tdt != nullptr;
  auto T = [&](size_t o) { return t1 ? t_[o] + tau1.dt * tau1.tdt[o] : t_[o]; };
  auto Q = [&](size_t o) { return t1 ? q_[o] + tau1.dt * tau1.qdt[o] : q_[o]; };
  auto U = [&](size_t o) { return t1 ? u_[o] + tau1.dt * tau1.udt[o] : u_[o]; };
  auto V = [&](size_t o) { return t1 ? v_[o] + tau1.dt * tau1.vdt[o] : v_[o]; };
  const double z_surf = z_half[(size_t)K * nc + col];
  const double us = u_star[col], bs = b_star[col];
  auto svcp_at = [&](int k, double& zag) {
    size_t o = (size_t)k * nc + col;
    zag = z_full[o] - z_surf;
    double tt = T(o);
    return c.do_simple ? tt + gcp * zag : tt * (1.0 + c.d608 * Q(o)) + gcp * zag;
  };
  double h;
  if (c.fixed_depth) h = c.depth_0;
  else {                                                     // pbl_depth :358-456
    double h1, tbot = svcp_at(K - 1, h1);
    h = h1;
    if (bs <= 0.0 || c.do_simple) {
      size_t o = (size_t)(K - 1) * nc + col;
      double ub = U(o), vb = V(o);
      double rich1 = h1 * c.grav * (tbot - tbot) / tbot / (ub * ub + vb * vb + small);
      for (int k = K - 2; k >= 0; --k) {
        double h2, t2 = svcp_at(k, h2);
        o -= nc;
        const double uk = U(o), vk = V(o);
        double rich2 = h2 * c.grav * (t2 - tbot) / tbot / (uk * uk + vk * vk + small);
        if (rich2 > c.rich_crit_pbl) { h = h2 + (h1 - h2) * (rich2 - c.rich_crit_pbl) / (rich2 - rich1); break; }
        rich1 = rich2; h1 = h2;
      }
    } else {
      double h_in = c.frac_inner * c.znom, ws, kt_dummy;
      mo_diff_point(mc, h_in, us, bs, ws, kt_dummy);
      ws = fmax(small, ws / c.vonkarm / h_in);
      double svp = tbot * (1.0 + (c.parcel_buoy * us * bs / c.grav / ws));
      double t1 = tbot;
      for (int k = K - 2; k >= 0; --k) {
        double h2, t2 = svcp_at(k, h2);
        if (t2 > svp) { h = h2 + (h1 - h2) * (t2 - svp) / (t2 - t1); break; }
        h1 = h2; t1 = t2;
      }
    }
  }
  h_out[col] = h;
  // diffusivity_pbl :458-526, then + saved input, diffusivity_entr :732-750, background floors
  const double h_inner = c.frac_inner * h;
  double km_ref, kt_ref;
  mo_diff_point(mc, h_inner, us, bs, km_ref, kt_ref);
  const bool entr = c.entr_ratio > 0.0 && !c.fixed_depth && c.do_entrain && bs > 0.0;
  double zag_prev = 0.0, sv_prev = 0.0;
  if (entr) sv_prev = svcp_at(0, zag_prev);
  for (int k = 0; k < K; ++k) {
    size_t o = (size_t)k * nc + col;
    double nm = 0.0, nt = 0.0;
    if (k > 0) {
      double zm = z_half[o] - z_surf;
      if (zm >= h_inner && zm < h) {
        double r = 1.0 - (zm - h_inner) / (h - h_inner);
        double factor = (zm / h_inner) * (r * r);
        nm = km_ref * factor; nt = kt_ref * factor;
      } else if (zm < h_inner) mo_diff_point(mc, zm, us, bs, nm, nt);
    }
    if (c.free_atm_diff && k > 0) {                          // diffusivity_free :604-697 (overwrites the boundary-layer values)
      double zag_k, zag_km1;
      const double sv_k = svcp_at(k, zag_k), sv_km1 = svcp_at(k - 1, zag_km1);
      const double dz = zag_km1 - zag_k;
      const double b = c.grav * (sv_km1 - sv_k) / sv_k;
      const double du = U(o - nc) - U(o), dv = V(o - nc) - V(o);
      const double speed2 = du * du + dv * dv;
      double rich = b * dz / (speed2 + small);
      rich = fmax(rich, 0.0);
      double fri2 = 0.0;
      if (c.free_atm_skyhi_diff && rich < c.rich_crit_diff) { const double a = 1.0 - rich / c.rich_crit_diff; fri2 = a * a; }
      double dz15 = 0.0;
      if (c.ampns) { dz15 = pow(dz, 1.5); rich = rich / fmin(1.0 + 1.0e-04 * dz15, c.ampns_max); }
      const double af = 1.0 - rich / c.rich_crit_diff, fri = af * af;
      if (rich < c.rich_crit_diff && (z_half[o] - z_surf) > h) {
        if (c.free_atm_skyhi_diff) {
          nm = c.ampns ? c.mix_len * c.mix_len * sqrt(speed2) * fri * (1.0 + 1.0e-04 * dz15) / dz
                       : c.mix_len * c.mix_len * sqrt(speed2) * fri / dz;
          nt = nm * (0.1 + 0.9 * fri2);
        } else {
          nt = c.mix_len * c.mix_len * sqrt(speed2) * fri / dz;
          nm = nt * c.rich_prandtl;
        }
      }
    }
    if (add_input) { nm = nm + k_m[o]; nt = nt + k_t[o]; }
    if (entr) {
      double zag, sv = svcp_at(k, zag);
      if (k > 0 && zag_prev > h && zag <= h) {
        nt = (zag_prev - zag) * c.entr_ratio * sv * us * bs / c.grav / fmax(small, sv_prev - sv);
        nm = nt;
      }
      zag_prev = zag; sv_prev = sv;
    }
    if (c.background_m > 0.0) nm = fmax(nm, c.background_m);
    if (c.background_t > 0.0) nt = fmax(nt, c.background_t);
    k_m[o] = nm; k_t[o] = nt;
  }
}

}  // namespace

namespace isca_phys {
void launch_diffusivity(IscaPhysics p, const double* t, const double* q, const double* u, const double* v, const double* z_full,
                        const double* z_half, const double* u_star, const double* b_star, double* h, double* k_m, double* k_t,
                        const double* tdt, const double* qdt, const double* udt, const double* vdt, double dt, int add_input) {
  TurbConst c;
  c.grav = p->cfg.grav; c.cp_air = p->cfg.cp_air; c.d608 = (p->cfg.rvgas - p->cfg.rdgas) / p->cfg.rdgas; c.vonkarm = p->cfg.vonkarm;
  c.fixed_depth = p->cfg.fixed_depth; c.do_entrain = p->cfg.diffusivity_do_entrain; c.do_simple = p->cfg.diffusivity_do_simple;
  c.depth_0 = p->cfg.depth_0; c.frac_inner = p->cfg.frac_inner; c.rich_crit_pbl = p->cfg.rich_crit_pbl; c.entr_ratio = p->cfg.entr_ratio;
  c.parcel_buoy = p->cfg.parcel_buoy; c.znom = p->cfg.znom; c.background_m = p->cfg.background_m; c.background_t = p->cfg.background_t;
  c.free_atm_diff = p->cfg.free_atm_diff; c.free_atm_skyhi_diff = p->cfg.free_atm_skyhi_diff; c.ampns = p->cfg.ampns;
  c.rich_crit_diff = p->cfg.rich_crit_diff; c.mix_len = p->cfg.mix_len; c.rich_prandtl = p->cfg.rich_prandtl; c.ampns_max = p->cfg.ampns_max;
  diffusivity_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>(mo_const(p), c, (int)p->ncol, p->K, t, q, u, v, Tau1{tdt, qdt, udt, vdt, dt}, add_input,
                                                            z_full, z_half, u_star, b_star, h, k_m, k_t);
}
}  // namespace isca_phys

extern "C" int isca_b200_diffusivity(IscaPhysics p, const double* t, const double* q, const double* u, const double* v,
                                     const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                                     const double* u_star, const double* b_star, double* h, double* k_m, double* k_t) {
  if (!p) return fail(nullptr, "null handle");
  if (!p_full || !p_half) return fail(p, "null input array");      // pbl_mcm only; kept for the reference argument list
  size_t nc = p->ncol, n3 = nc * p->K;
  Dev* b = p->buf;
  if (up(p, b[0], t, n3) || up(p, b[1], q, n3) || up(p, b[2], u, n3) || up(p, b[3], v, n3) || up(p, b[4], z_full, n3) ||
      up(p, b[5], z_half, n3 + nc) || up(p, b[6], u_star, nc) || up(p, b[7], b_star, nc) || up(p, b[8], k_m, n3) || up(p, b[9], k_t, n3)) return 1;
  if (!b[10].ensure(nc)) return fail(p, "cudaMalloc failed");
  launch_diffusivity(p, b[0].p, b[1].p, b[2].p, b[3].p, b[4].p, b[5].p, b[6].p, b[7].p, b[10].p, b[8].p, b[9].p, nullptr, nullptr, nullptr, nullptr, 0.0, 1);
  if (down(p, b[10], h, nc) || down(p, b[8], k_m, n3) || down(p, b[9], k_t, n3)) return 1;
  return finish(p, "diffusivity");
}
