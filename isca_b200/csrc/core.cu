// core.cu -- handle, device-resident model state, step orchestration and the C ABI.
//
// Host-side mirror of atmosphere_mod / spectral_dynamics_mod
// (atmos_spectral/driver/solo/atmosphere.F90:120-399, model/spectral_dynamics.F90:230-1034).
#include "device.h"
#include "spectral.h"
#include "grid.h"
#include "nccl_dyn.h"
#include "tracer.h"
#include "core_internal.h"
#include <cstring>
#include <map>
#include <algorithm>

using namespace isca;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +   \
                               __FILE__ + ":" + std::to_string(__LINE__));                       \
  } while (0)

static thread_local std::string g_create_error;

// Stream-ordered host<->device copies.  A plain cudaMemcpy from pageable memory may return before
// the DMA has landed and is not ordered against work on a non-blocking stream, so every copy goes
// through the library's stream and is followed by a stream synchronize.
static void h2d_on(cudaStream_t st, void* dst, const void* src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
}
static void d2h_on(cudaStream_t st, void* dst, const void* src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
}
static void d2d_on(cudaStream_t st, void* dst, const void* src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
}

namespace {

template <class T>
struct DBuf {                       // owning device buffer
  T* p = nullptr; size_t n = 0;
  void alloc(size_t count, bool zero = true) {
    release(); n = count;
    if (count) { CK(cudaMalloc(&p, count * sizeof(T))); if (zero) CK(cudaMemset(p, 0, count * sizeof(T))); }
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
  void upload(const std::vector<T>& h) { alloc(h.size(), false); if (!h.empty()) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); CK(cudaDeviceSynchronize()); } }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DBuf() { release(); }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct IscaHandle_t {
  IscaConfig cfg;
  Geometry g;
  HostTables ht;
  DevTables dt;
  Params pr;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;              // second stream: Legendre of sub-batch i+1 overlaps the FFT of sub-batch i
  cudaEvent_t ev_pipe[8] = {nullptr};
  int pipe_parts = 1;                      // >1 measured slower on B200 (profiles/r01_experiments.md); ISCA_B200_PIPE overrides
  std::string err;
  long long launches = 0;
  long long steps = 0;
  double last_step_ms = 0.0;
  int previous = 0, current = 0;

  // ---- device tables
  DBuf<int> d_m_of, d_off, d_pos, d_row_m, d_row_n, d_owner, d_lidx, d_nm_rank, d_roff;
  DBuf<double*> d_peerA, d_peerB;
  std::vector<void*> ipc_opened;
  DBuf<double> barrier_token;
  bool p2p = false;
  DBuf<double> d_sin_lat, d_cos_lat, d_cosm_lat, d_wts_lat, d_coriolis, d_rad_lat, d_pk, d_bk, d_dpk, d_dbk;
  DBuf<double> d_leg, d_legw, d_ln_bk;
  DBuf<double> d_sg[14];
  DBuf<double> d_eigen, d_uvm, d_uvc, d_uvp, d_alpm, d_alpp, d_dym, d_dx, d_dyp, d_mask, d_damp, d_dampv, d_dampd, d_eddy, d_zmu, d_zmv;
  DBuf<double> d_rlh, d_rlf, d_rt, d_h;
  DBuf<double> d_twiddle;
  std::map<double, DBuf<double>*> wave_cache;          // xi -> wave matrices

  // ---- state
  DBuf<double2> vors[2], divs[2], ts[2], lnps[2];
  DBuf<double> u[2], v[2], T[2], ps[2];
  DBuf<double> vorg, divg, phis, wg_full;
  // ---- grid tracer (sphum)
  DBuf<double> q[2], tr1, wg, tr_wpart;
  DBuf<double> tr_halo;      // nranks > 1: [halo_s | halo_n | send_s | send_n] x [3][K][2][I]
  DBuf<double> d_fv_c, d_fv_cc, d_fv_dy, d_fv_dyy, d_fv_dyp, d_fv_dym, d_fv_rdxc, d_fv_rdyy, d_fv_rcdy, d_fv_rdy;
  FvTables fv;
  DBuf<int> ops_sum3;
  // ---- work
  DBuf<double2> dt_vors, w_div, w_T, w_lnps, k_dt_vors, k_dt_divs, k_dt_ts, k_dt_lnps;
  DBuf<double2> specA, specB, specC;
  DBuf<double> four;         // Fourier buffer, m-owner layout (A)
  DBuf<double> fourB;        // Fourier buffer, lat-owner layout (B); only allocated when P > 1
  NcclApi nccl; NcclComm comm = nullptr;
  double* four_lat() { return g.P > 1 ? fourB.p : four.p; }
  // The lat-owner buffer carries a tail that the peers address through the same CUDA-IPC mapping: arrival flags of the device-side
  // inter-GPU barriers (SYNC_FLAGS doubles: channel c, source rank r at [32 c + r]) and the receive buffers of the tracer halo
  // (halo_s, halo_n: 3*K*2*I doubles each).  sync_off = offset of the tail in doubles; equal on every rank.
  static constexpr size_t SYNC_FLAGS = 512;
  size_t fourB_cap = 0, sync_off = 0;
  size_t halo_block() const { return (size_t)3 * g.K * 2 * g.I; }
  std::vector<double*> peerA_host, peerB_host;
  DBuf<unsigned long long> bar_count;      // [3] epochs of the barrier channels (forward transpose, inverse transpose, tracer halo)
  void ensure_four(int Lp) {
    // m-owner buffer: nm*J rows (sized with nm_max so that every rank's buffer is equally large for peer writes)
    const size_t needA = (size_t)g.nm_max * g.J * 2 * Lp, needB = (size_t)(g.M + 1) * g.Jloc * 2 * Lp;
    if (p2p && (needA > four.n || needB > fourB_cap)) throw std::runtime_error("transform batch too large for the peer-mapped Fourier buffers");
    if (needA > four.n) four.alloc(needA);
    if (g.P > 1 && needB > fourB_cap) {
      const size_t tail = SYNC_FLAGS + 2 * halo_block();
      fourB.alloc(needB + tail);
      fourB_cap = needB; sync_off = needB;
      if (cudaMemset(fourB.p + sync_off, 0, tail * sizeof(double)) != cudaSuccess) throw std::runtime_error("cudaMemset(sync tail) failed");
    }
  }
  DBuf<double> gradA;        // [2K+2] planes: dxT, dyT, dxlnps, dylnps
  DBuf<double> gridB;        // [4K+1] planes: dt_T, A, B, Phi, dt_lnps
  DBuf<double> ext_tend;     // [3K+1] planes for externally supplied tendencies
  DBuf<LevDesc> levsA, levsB, levsC[2];
  DBuf<unsigned char> truncB;
  DBuf<double> part, scal, red_tmp;
  DBuf<int> ops_sum2, ops_sum1, ops_fix;
  int LpA = 0, LpB = 0, LpC = 0;
  int keep_tend = 0;
  bool grad_valid = false;   // gradA planes hold the gradients of the current level (computed by the previous step)
  // ---- CUDA graphs of the step, one per (current slot, physics) variant of the leapfrog step
  struct StepGraph { int uses = 0; cudaGraphExec_t exec = nullptr; long long launches = 0; };
  StepGraph graphs[4];
  bool use_graph = true;
  // ---- per-kernel-group profiling with CUDA events (isca_b200_profile_step)
  bool profiling = false;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  void mark(const char* name) {
    if (!profiling) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
    marks.emplace_back(name, e);
  }
  // ---- scratch for the transforms_mod-level API
  DBuf<double2> x_rect, x_spec, x_fr, x_imp[9];
  struct DiagAcc { DBuf<double>* sum = nullptr; long long count = 0; size_t n = 0; };
  std::map<int, DiagAcc> diag;          // device-side time averages (isca_b200_diag_accumulate / _fetch)
  DBuf<double> x_grid, x_four, x_pz, x_der;          // x_der: derived diagnostic fields (field_on_device)
  DBuf<LevDesc> x_levs;
  DBuf<unsigned char> x_trunc;

  size_t nplane() const { return (size_t)g.Jloc * g.I; }
  size_t n3() const { return (size_t)g.K * g.Jloc * g.I; }
  size_t nspec3() const { return (size_t)g.T * g.K; }
  double denom() const { return ht.global_sum_of_wts * g.I; }
  int owns_m0() const { return g.owner[0] == g.rank ? 1 : 0; }
};
typedef IscaHandle_t H;

// ---------------------------------------------------------------------------------------------
static void upload_tables(H& h) {
  const Geometry& g = h.g; const HostTables& t = h.ht;
  h.d_m_of.upload(g.m_of); h.d_off.upload(g.off); h.d_pos.upload(g.pos);
  h.d_row_m.upload(t.row_m); h.d_row_n.upload(t.row_n);
  {
    std::vector<int> lidx(g.M + 1, 0), cnt(g.P, 0);
    for (int m = 0; m <= g.M; ++m) lidx[m] = cnt[g.owner[m]]++;
    h.d_owner.upload(g.owner); h.d_lidx.upload(lidx); h.d_nm_rank.upload(g.nm_rank); h.d_roff.upload(g.roff);
  }
  h.d_sin_lat.upload(t.sin_lat); h.d_cos_lat.upload(t.cos_lat); h.d_cosm_lat.upload(t.cosm_lat);
  h.d_wts_lat.upload(t.wts_lat); h.d_coriolis.upload(t.coriolis); h.d_rad_lat.upload(t.rad_lat);
  h.d_pk.upload(t.pk); h.d_bk.upload(t.bk); h.d_dpk.upload(t.dpk); h.d_dbk.upload(t.dbk);
  h.d_leg.upload(t.leg); h.d_legw.upload(t.legw);
  {
    std::vector<double> lb(t.bk.size(), 0.0);
    for (size_t k = 0; k < lb.size(); ++k) if (t.bk[k] > 0.0) lb[k] = std::log(t.bk[k]);
    h.d_ln_bk.upload(lb);
  }
  h.d_eigen.upload(t.eigen); h.d_uvm.upload(t.coef_uvm); h.d_uvc.upload(t.coef_uvc); h.d_uvp.upload(t.coef_uvp);
  h.d_alpm.upload(t.coef_alpm); h.d_alpp.upload(t.coef_alpp); h.d_dym.upload(t.coef_dym); h.d_dx.upload(t.coef_dx);
  h.d_dyp.upload(t.coef_dyp); h.d_mask.upload(t.trunc_mask);
  h.d_damp.upload(t.damping); h.d_dampv.upload(t.damping_vor); h.d_dampd.upload(t.damping_div);
  h.d_eddy.upload(t.eddy_sponge); h.d_zmu.upload(t.zmu_sponge); h.d_zmv.upload(t.zmv_sponge);
  h.d_rlh.upload(t.ref_ln_p_half); h.d_rlf.upload(t.ref_ln_p_full); h.d_rt.upload(t.ref_t); h.d_h.upload(t.h);
  h.d_twiddle.upload(t.twiddle);
  DevTables& d = h.dt;
  d.g.I = g.I; d.g.J = g.J; d.g.K = g.K; d.g.M = g.M; d.g.N = g.N; d.g.Jh = g.Jh; d.g.P = g.P; d.g.rank = g.rank;
  d.g.Jloc = g.Jloc; d.g.j0 = g.j0; d.g.nm = g.nm; d.g.T = g.T;
  d.g.m_of = h.d_m_of.p; d.g.off = h.d_off.p; d.g.pos = h.d_pos.p; d.g.row_m = h.d_row_m.p;
  d.g.p2p = 0; d.g.symmetric = h.cfg.make_symmetric ? 1 : 0; d.g.owner = h.d_owner.p; d.g.lidx = h.d_lidx.p; d.g.nm_rank = h.d_nm_rank.p; d.g.roff = h.d_roff.p;
  d.g.peerA = nullptr; d.g.peerB = nullptr;
  d.row_n = h.d_row_n.p;
  d.sin_lat = h.d_sin_lat.p; d.cos_lat = h.d_cos_lat.p; d.cosm_lat = h.d_cosm_lat.p; d.wts_lat = h.d_wts_lat.p;
  d.coriolis = h.d_coriolis.p; d.rad_lat = h.d_rad_lat.p;
  d.pk = h.d_pk.p; d.bk = h.d_bk.p; d.dpk = h.d_dpk.p; d.dbk = h.d_dbk.p; d.ln_bk = h.d_ln_bk.p;
  d.leg = h.d_leg.p; d.legw = h.d_legw.p;
  d.eigen = h.d_eigen.p; d.coef_uvm = h.d_uvm.p; d.coef_uvc = h.d_uvc.p; d.coef_uvp = h.d_uvp.p;
  d.coef_alpm = h.d_alpm.p; d.coef_alpp = h.d_alpp.p; d.coef_dym = h.d_dym.p; d.coef_dx = h.d_dx.p; d.coef_dyp = h.d_dyp.p;
  d.trunc_mask = h.d_mask.p; d.damping = h.d_damp.p; d.damping_vor = h.d_dampv.p; d.damping_div = h.d_dampd.p;
  d.eddy_sponge = h.d_eddy.p; d.zmu_sponge = h.d_zmu.p; d.zmv_sponge = h.d_zmv.p;
  d.ref_ln_p_half = h.d_rlh.p; d.ref_ln_p_full = h.d_rlf.p; d.ref_t = h.d_rt.p; d.h_impl = h.d_h.p;
  d.wave_matrix = nullptr;
  d.twiddle = reinterpret_cast<const double2*>(h.d_twiddle.p);
}

static void set_params(H& h) {
  const IscaConfig& c = h.cfg; Params& p = h.pr;
  p.rdgas = c.rdgas; p.kappa = c.rdgas / (c.rdgas / c.kappa); p.cp_air = c.rdgas / c.kappa; p.grav = c.grav;
  p.radius = c.radius; p.omega = c.omega; p.ref_ps = c.reference_sea_level_press; p.dt_atmos = c.dt_atmos;
  p.robert_coeff = c.robert_coeff; p.raw_filter_coeff = c.raw_filter_coeff; p.virtual_factor = 0.0;
  p.tka = c.ka < 0. ? -1. / (86400 * c.ka) : c.ka;          // hs_forcing.F90:392-404
  p.tks = c.ks < 0. ? -1. / (86400 * c.ks) : c.ks;
  p.vkf = c.kf < 0. ? -1. / (86400 * c.kf) : c.kf;
  p.sigma_b = c.sigma_b; p.t_zero = c.t_zero; p.t_strat = c.t_strat; p.delh = c.delh; p.delv = c.delv; p.eps = c.eps;
  p.P00 = c.P00; p.do_conserve_energy = c.do_conserve_energy; p.no_forcing = c.no_forcing; p.physics_on = 1;
  p.pk0_zero = (h.ht.pk[0] == 0.0); p.pkbk0_zero = (h.ht.pk[0] == 0.0 && h.ht.bk[0] == 0.0);
  p.vr_tmin = c.valid_range_t[0]; p.vr_tmax = c.valid_range_t[1];
  p.pure_sigma = 1;
  for (double v : h.ht.pk) if (v != 0.0) p.pure_sigma = 0;
  for (size_t k = 1; k < h.ht.bk.size(); ++k) if (!(h.ht.bk[k] > 0.0)) p.pure_sigma = 0;
  p.xi = 0; p.delta_t = 0; p.first_step = 1;
  // pure-sigma fast path tables (see grid.cu)
  const std::vector<double>& bk = h.ht.bk;
  const int K = h.g.K;
  p.sigma_fast = (p.pure_sigma && p.pkbk0_zero && bk[K] == 1.0 && std::getenv("ISCA_B200_NO_SIGMA_FAST") == nullptr) ? 1 : 0;
  if (p.sigma_fast) {
    std::vector<double> db(K), rdb(K), al(K), d3(K), lf(K), pf(K), pfk(K), x1c(K), lnb(K + 1, 0.0);
    for (int k = 1; k <= K; ++k) lnb[k] = std::log(bk[k]);
    for (int k = 0; k < K; ++k) {
      db[k] = bk[k + 1] - bk[k];
      rdb[k] = 1.0 / db[k];
      double d2;
      if (k == 0) { al[k] = 1.0; d3[k] = 0.0; lf[k] = lnb[1] - 1.0; d2 = 0.0; }      // ln_p_full(1) = ln_p_half(2) - 1 ; bk(1) = 0
      else {
        d3[k] = lnb[k + 1] - lnb[k];
        al[k] = 1.0 - bk[k] * d3[k] / db[k];
        lf[k] = lnb[k + 1] - al[k];
        d2 = lf[k] - lnb[k];
      }
      pf[k] = std::exp(lf[k]);
      pfk[k] = std::exp(p.kappa * lf[k]);
      x1c[k] = (bk[k + 1] * al[k] + bk[k] * d2) / db[k];
    }
    // PPM weights with dz = db (vert_advection.F90:504-563 slope_z, :567-629 compute_weights)
    std::vector<double> c1(K, 0.0), c2(K, 0.0), z1(K, 0.0), z2(K, 0.0), z3(K, 0.0);
    for (int k = 1; k <= K - 2; ++k) {
      const double dm = db[k - 1], d0 = db[k], dp = db[k + 1];
      c1[k] = (2. * dm + d0) / (dp + d0) * d0 / (dm + d0 + dp);
      c2[k] = (2. * dp + d0) / (d0 + dm) * d0 / (dm + d0 + dp);
    }
    for (int k = 2; k <= K - 2; ++k) {
      const double e2 = db[k - 2], e1 = db[k - 1], e0 = db[k], ep = db[k + 1];
      const double denom1 = 1.0 / (e1 + e0), denom2 = 1.0 / (e2 + e1 + e0 + ep), denom3 = 1.0 / (2 * e1 + e0), denom4 = 1.0 / (e1 + 2 * e0);
      const double num3 = e2 + e1, num4 = e0 + ep;
      const double x = num3 * denom3 - num4 * denom4, y = 2.0 * e1 * e0, z0 = e1 * denom1;
      z1[k] = z0 + x * y * denom1 * denom2;
      z2[k] = e1 * num3 * denom3 * denom2;
      z3[k] = e0 * num4 * denom4 * denom2;
    }
    const std::vector<double>* src[14] = {&bk, &db, &rdb, &al, &d3, &lf, &pf, &pfk, &x1c, &c1, &c2, &z1, &z2, &z3};
    for (int q = 0; q < 14; ++q) h.d_sg[q].upload(*src[q]);
    SigmaTables& sg = h.dt.sig;
    sg.b = h.d_sg[0].p; sg.db = h.d_sg[1].p; sg.rdb = h.d_sg[2].p; sg.al = h.d_sg[3].p; sg.d3 = h.d_sg[4].p;
    sg.lf = h.d_sg[5].p; sg.pf = h.d_sg[6].p; sg.pfk = h.d_sg[7].p; sg.x1c = h.d_sg[8].p;
    sg.ppm_c1 = h.d_sg[9].p; sg.ppm_c2 = h.d_sg[10].p; sg.ppm_z1 = h.d_sg[11].p; sg.ppm_z2 = h.d_sg[12].p; sg.ppm_z3 = h.d_sg[13].p;
  }
}

static void alloc_state(H& h) {
  const Geometry& g = h.g; const int K = g.K;
  for (int s = 0; s < 2; ++s) {
    h.vors[s].alloc(h.nspec3()); h.divs[s].alloc(h.nspec3()); h.ts[s].alloc(h.nspec3()); h.lnps[s].alloc(g.T);
    h.u[s].alloc(h.n3()); h.v[s].alloc(h.n3()); h.T[s].alloc(h.n3()); h.ps[s].alloc(h.nplane());
  }
  h.vorg.alloc(h.n3()); h.divg.alloc(h.n3()); h.phis.alloc(h.nplane()); h.wg_full.alloc(h.n3());
  if (h.cfg.num_tracers > 0) {
    for (int s = 0; s < 2; ++s) h.q[s].alloc(h.n3());
    h.tr1.alloc(h.n3()); h.tr_wpart.alloc(4 * h.nplane());
    h.wg.alloc(h.n3() + h.nplane());
    if (g.P > 1) { h.tr_halo.alloc((size_t)12 * K * 2 * g.I); CK(cudaMemset(h.tr_halo.p, 0, (size_t)12 * K * 2 * g.I * sizeof(double))); }
    const HostTables& t = h.ht;
    h.d_fv_c.upload(t.fv_c); h.d_fv_cc.upload(t.fv_cc); h.d_fv_dy.upload(t.fv_dy); h.d_fv_dyy.upload(t.fv_dyy);
    h.d_fv_dyp.upload(t.fv_dy_plus); h.d_fv_dym.upload(t.fv_dy_minus);
    h.fv.c = h.d_fv_c.p; h.fv.cc = h.d_fv_cc.p; h.fv.dy = h.d_fv_dy.p; h.fv.dyy = h.d_fv_dyy.p;
    h.fv.dy_plus = h.d_fv_dyp.p; h.fv.dy_minus = h.d_fv_dym.p; h.fv.dx = t.fv_dx;
    {
      const int J = g.J;
      std::vector<double> rdxc(J), rdyy(J + 1), rcdy(J), rdy(J + 4);
      for (int j = 0; j < J; ++j) { rdxc[j] = 1.0 / (t.fv_dx * t.fv_c[j]); rcdy[j] = 1.0 / (t.fv_c[j] * t.fv_dy[j + 2]); }
      for (int j = 0; j <= J; ++j) rdyy[j] = 1.0 / t.fv_dyy[j];
      for (int j = 0; j < J + 4; ++j) rdy[j] = 1.0 / t.fv_dy[j];
      h.d_fv_rdxc.upload(rdxc); h.d_fv_rdyy.upload(rdyy); h.d_fv_rcdy.upload(rcdy); h.d_fv_rdy.upload(rdy);
      h.fv.rdxc = h.d_fv_rdxc.p; h.fv.rdyy = h.d_fv_rdyy.p; h.fv.rcdy = h.d_fv_rcdy.p; h.fv.rdy = h.d_fv_rdy.p;
    }
    h.ops_sum3.upload({0, 0, 0});
  }
  h.dt_vors.alloc(h.nspec3()); h.w_div.alloc(h.nspec3()); h.w_T.alloc(h.nspec3()); h.w_lnps.alloc(g.T);
  h.k_dt_vors.alloc(h.nspec3()); h.k_dt_divs.alloc(h.nspec3()); h.k_dt_ts.alloc(h.nspec3()); h.k_dt_lnps.alloc(g.T);
  h.LpA = round_up(2 * K + 2, 16); h.LpB = round_up(4 * K + 1, 16); h.LpC = round_up(7 * K + 3, 16);
  h.specA.alloc((size_t)g.T * h.LpA); h.specB.alloc((size_t)g.T * h.LpB); h.specC.alloc((size_t)g.T * h.LpC);
  const int Lmax = std::max(h.LpA, std::max(h.LpB, h.LpC));
  h.ensure_four(Lmax);
  h.gradA.alloc((size_t)(2 * K + 2) * h.nplane());
  h.gridB.alloc((size_t)(4 * K + 1) * h.nplane());
  h.part.alloc(9 * h.nplane()); h.scal.alloc(SC_COUNT); h.red_tmp.alloc(9 * 128);
  h.ops_sum2.upload({0, 0}); h.ops_sum1.upload({0}); h.ops_fix.upload({0, 0, 0, 0, 0, 0, 0, 2, 2});
  const size_t pl = h.nplane();
  // level descriptors
  std::vector<LevDesc> la(2 * K + 2), lb(4 * K + 1);
  for (int k = 0; k < K; ++k) {
    la[k] = {h.gradA.p + (size_t)k * pl, 1, 0};                 // dx T  -> divide_by_cos
    la[K + k] = {h.gradA.p + (size_t)(K + k) * pl, 1, 0};       // dy T
  }
  la[2 * K] = {h.gradA.p + (size_t)(2 * K) * pl, 0, 0};         // dx ln ps (cos division after *psg, in grid_step)
  la[2 * K + 1] = {h.gradA.p + (size_t)(2 * K + 1) * pl, 0, 0};
  h.levsA.upload(la);
  for (int i = 0; i < 4 * K + 1; ++i) lb[i] = {h.gridB.p + (size_t)i * pl, 0, 0};
  h.levsB.upload(lb);
  for (int f = 0; f < 2; ++f) {
    std::vector<LevDesc> lc(7 * K + 3);
    for (int k = 0; k < K; ++k) {
      lc[k] = {h.vorg.p + (size_t)k * pl, 0, 0};
      lc[K + k] = {h.divg.p + (size_t)k * pl, 0, 0};
      lc[2 * K + k] = {h.u[f].p + (size_t)k * pl, 1, 0};
      lc[3 * K + k] = {h.v[f].p + (size_t)k * pl, 1, 0};
      lc[4 * K + k] = {h.T[f].p + (size_t)k * pl, 0, 0};
    }
    lc[5 * K] = {h.ps[f].p, 2, 0};                               // psg = exp(ln_psg)
    for (int i = 0; i < 2 * K + 2; ++i) lc[5 * K + 1 + i] = la[i];  // gradients for the next step (same planes/ops as batch A)
    h.levsC[f].upload(lc);
  }
  std::vector<unsigned char> tb(h.LpB, 0);
  for (int k = 0; k < K; ++k) { tb[k] = 1; tb[3 * K + k] = 1; }  // dt_T, Phi+KE truncated; A,B not (transforms.F90:766-770)
  tb[4 * K] = 1;
  h.truncB.upload(tb);
}

static void ensure_wave_matrix(H& h, double xi) {
  auto it = h.wave_cache.find(xi);
  if (it == h.wave_cache.end()) {
    std::vector<double> wm;
    build_wave_matrices(h.cfg, h.g, h.ht, xi, wm);
    DBuf<double>* b = new DBuf<double>();
    b->upload(wm);
    it = h.wave_cache.emplace(xi, b).first;
  }
  h.dt.wave_matrix = it->second->p;
}

// Device-side inter-GPU barrier over peer memory (replaces a one-element ncclAllReduce, ~4x the latency): lane r stores this rank's new
// epoch into rank r's flag array (release, system scope; the producing kernel's remote stores precede it in stream order and are fenced)
// and spins until rank r's epoch has arrived in this rank's array (acquire).  Epochs only grow, so no reset is needed; they live in
// device memory so that a replayed CUDA graph advances them.
__global__ void peer_barrier_kernel(double* const* __restrict__ peerB, size_t flag_off, int P, int rank, int chan, unsigned long long* counter) {
  __shared__ unsigned long long ep;
  if (threadIdx.x == 0) ep = ++counter[chan];
  __syncthreads();
  const unsigned long long epoch = ep;
  const int r = threadIdx.x;
  if (r < P && r != rank) {
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(peerB[r] + flag_off) + chan * 32 + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peerB[rank] + flag_off) + chan * 32 + r;
    unsigned long long v;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v < epoch);
  }
}
static void peer_barrier(H& h, int chan) {
  peer_barrier_kernel<<<1, 32, 0, h.st>>>(h.d_peerB.p, h.sync_off, h.g.P, h.g.rank, chan, h.bar_count.p);
  h.launches++;
}

// exchange of the Fourier buffer between the lat-owner and m-owner layouts (transpose_fourier /
// reverse_transpose_fourier, tools/transforms.F90:970-1056).  One rank: the two layouts coincide.
//   direction 0 (spectral -> grid): layout A (mine: [dest s][mi][jl][C]) -> layout B ([pos[m]][jl][C])
//   direction 1 (grid -> spectral): layout B -> layout A
// Every per-peer block is contiguous on both sides (the kernels index straight into the per-peer slabs),
// so the whole transpose is one grouped NCCL send/recv per batch -- no pack/unpack kernels, no barrier.
static void exchange_fourier(H& h, int direction, int Lp) {
  const Geometry& g = h.g;
  if (g.P == 1) return;
  if (h.p2p) {
    // the producing kernel already stored into the peers' buffers; the flag barrier completes on a rank only after every rank's
    // producer kernel (earlier in its stream) has finished
    peer_barrier(h, direction);
    h.mark(direction == 0 ? "exchange_inv" : "exchange_fwd");
    return;
  }
  const size_t C = 2 * (size_t)Lp;
  const NcclApi& n = h.nccl;
  n.ck(n.GroupStart(), "ncclGroupStart");
  for (int r = 0; r < g.P; ++r) {
    double* a_blk = h.four.p + (size_t)r * g.nm * g.Jloc * C;                 // my m's, rank r's latitudes
    double* b_blk = h.fourB.p + (size_t)g.roff[r] * g.Jloc * C;               // rank r's m's, my latitudes
    const size_t na = (size_t)g.nm * g.Jloc * C, nb = (size_t)g.nm_rank[r] * g.Jloc * C;
    if (direction == 0) {
      n.ck(n.Send(a_blk, na, NCCL_FLOAT64, r, h.comm, h.st), "ncclSend");
      n.ck(n.Recv(b_blk, nb, NCCL_FLOAT64, r, h.comm, h.st), "ncclRecv");
    } else {
      n.ck(n.Send(b_blk, nb, NCCL_FLOAT64, r, h.comm, h.st), "ncclSend");
      n.ck(n.Recv(a_blk, na, NCCL_FLOAT64, r, h.comm, h.st), "ncclRecv");
    }
  }
  n.ck(n.GroupEnd(), "ncclGroupEnd");
  h.mark(direction == 0 ? "exchange_inv" : "exchange_fwd");
}
// latitude halo of the grid tracer step (fv_advection.F90:161-162): the two edge rows of (tr0, u, v), packed by
// tracer_halo_pack_kernel, go to the southern / northern neighbour in one grouped send/recv
static void exchange_tracer_halo(H& h, const TracerArgs& ta) {
  const Geometry& g = h.g;
  if (g.P == 1) return;
  launch_tracer_halo_pack(h.dt, h.pr, ta, h.st); h.launches++;
  if (h.p2p) {                             // the pack kernel stored the edge rows straight into the neighbours' halo buffers
    peer_barrier(h, 2);
    h.mark("tracer_halo");
    return;
  }
  const size_t n = (size_t)3 * g.K * 2 * g.I;
  const NcclApi& nc = h.nccl;
  nc.ck(nc.GroupStart(), "ncclGroupStart");
  if (g.rank > 0) {
    nc.ck(nc.Send(ta.send_s, n, NCCL_FLOAT64, g.rank - 1, h.comm, h.st), "ncclSend(halo)");
    nc.ck(nc.Recv(ta.halo_s, n, NCCL_FLOAT64, g.rank - 1, h.comm, h.st), "ncclRecv(halo)");
  }
  if (g.rank < g.P - 1) {
    nc.ck(nc.Send(ta.send_n, n, NCCL_FLOAT64, g.rank + 1, h.comm, h.st), "ncclSend(halo)");
    nc.ck(nc.Recv(ta.halo_n, n, NCCL_FLOAT64, g.rank + 1, h.comm, h.st), "ncclRecv(halo)");
  }
  nc.ck(nc.GroupEnd(), "ncclGroupEnd");
  h.mark("tracer_halo");
}
// global reductions across ranks of `count` device scalars (area_weighted_global_mean's mpp_global_field +
// sum, tools/transforms.F90:1059-1077, becomes a local fixed-order reduction + one small all-reduce)
static void allreduce_scalars(H& h, double* dev, int count, int op) {
  if (h.g.P == 1) return;
  h.nccl.ck(h.nccl.AllReduce(dev, dev, (size_t)count, NCCL_FLOAT64, op, h.comm, h.st), "ncclAllReduce");
}

// ---------------------------------------------------------------------------------------------
// generic batched transforms on device data
// ---------------------------------------------------------------------------------------------
static void dev_inverse(H& h, const double2* spec, int Lp, const LevDesc* levs, int nlev, const char* tag = "") {
  const int nct = Lp / 16;                               // 32-column tiles of the batch
  const int np = (h.g.P == 1 && !h.profiling) ? std::min(h.pipe_parts, nct) : 1;
  if (np <= 1) {
    launch_legendre_inv(h.dt, spec, h.four.p, Lp, h.st); h.launches++;
    if (h.profiling) h.mark((std::string("legendre_inv") + tag).c_str());
    exchange_fourier(h, 0, Lp);
    launch_fft_inv(h.dt, h.four_lat(), levs, nlev, Lp, h.st); h.launches++;
    if (h.profiling) h.mark((std::string("fft_inv") + tag).c_str());
    return;
  }
  // software pipeline over np sub-batches on two streams: Legendre(i) runs beside FFT(i-1)
  CK(cudaEventRecord(h.ev_pipe[0], h.st));
  CK(cudaStreamWaitEvent(h.st2, h.ev_pipe[0], 0));
  int ct = 0;
  for (int i = 0; i < np; ++i) {
    const int cnt = (nct - ct + (np - i) - 1) / (np - i);
    cudaStream_t s = (i & 1) ? h.st2 : h.st;
    if (i > 0) CK(cudaStreamWaitEvent(s, h.ev_pipe[1 + ((i - 1) & 1)], 0));      // stagger: after Legendre(i-1)
    launch_legendre_inv(h.dt, spec, h.four.p, Lp, s, ct, cnt);
    CK(cudaEventRecord(h.ev_pipe[1 + (i & 1)], s));
    const int lev_begin = ct * 16, lev_end = std::min((ct + cnt) * 16, nlev);
    if (lev_begin < lev_end) { launch_fft_inv(h.dt, h.four.p, levs, lev_end, Lp, s, lev_begin); h.launches++; }
    h.launches++;
    ct += cnt;
  }
  CK(cudaEventRecord(h.ev_pipe[3], h.st2));
  CK(cudaStreamWaitEvent(h.st, h.ev_pipe[3], 0));
}
static void dev_forward(H& h, const LevDesc* levs, int nlev, double2* spec, int Lp, const unsigned char* trunc,
                        const char* tag = "") {
  const int nct = Lp / 16;
  const int np = (h.g.P == 1 && !h.profiling) ? std::min(h.pipe_parts, nct) : 1;
  if (np <= 1) {
    launch_fft_fwd(h.dt, h.four_lat(), levs, nlev, Lp, h.st); h.launches++;
    if (h.profiling) h.mark((std::string("fft_fwd") + tag).c_str());
    exchange_fourier(h, 1, Lp);
    launch_legendre_fwd(h.dt, h.four.p, spec, Lp, trunc, h.st); h.launches++;
    if (h.profiling) h.mark((std::string("legendre_fwd") + tag).c_str());
    return;
  }
  CK(cudaEventRecord(h.ev_pipe[0], h.st));
  CK(cudaStreamWaitEvent(h.st2, h.ev_pipe[0], 0));
  int ct = 0;
  for (int i = 0; i < np; ++i) {
    const int cnt = (nct - ct + (np - i) - 1) / (np - i);
    cudaStream_t s = (i & 1) ? h.st2 : h.st;
    if (i > 0) CK(cudaStreamWaitEvent(s, h.ev_pipe[1 + ((i - 1) & 1)], 0));      // stagger: after FFT(i-1)
    const int lev_begin = ct * 16, lev_end = std::min((ct + cnt) * 16, nlev);
    if (lev_begin < lev_end) { launch_fft_fwd(h.dt, h.four.p, levs, lev_end, Lp, s, lev_begin); h.launches++; }
    CK(cudaEventRecord(h.ev_pipe[1 + (i & 1)], s));
    launch_legendre_fwd(h.dt, h.four.p, spec, Lp, trunc, s, ct, cnt); h.launches++;
    ct += cnt;
  }
  CK(cudaEventRecord(h.ev_pipe[3], h.st2));
  CK(cudaStreamWaitEvent(h.st, h.ev_pipe[3], 0));
}

// ---------------------------------------------------------------------------------------------
// one time step: atmosphere(Time) (atmosphere.F90:276-352)
// ---------------------------------------------------------------------------------------------
static void step_once(H& h, int physics_on, const double* dtu_in, const double* dtv_in, const double* dtt_in, const double* dtq_in = nullptr) {
  const Geometry& g = h.g; const int K = g.K;
  const int prev = h.previous, cur = h.current, fut = 1 - cur;
  const double delta_t = (prev == cur) ? h.cfg.dt_atmos : 2 * h.cfg.dt_atmos;     // atmosphere.F90:292-296
  Params& pr = h.pr;
  pr.delta_t = delta_t; pr.first_step = (prev == cur); pr.physics_on = physics_on;
  pr.xi = delta_t * h.cfg.alpha_implicit;
  if (h.cfg.use_implicit) ensure_wave_matrix(h, pr.xi);                          // implicit.F90:260-264
  const size_t pl = h.nplane();
  cudaStream_t st = h.st;

  // gradients of T(current), ln ps(current)  (horizontal_advection, compute_pressure_gradient): normally
  // produced by the previous step's inverse batch; computed here after a (re)start
  if (!h.grad_valid) {
    launch_spec_gradient(h.dt, h.ts[cur].p, K, K, h.specA.p, h.LpA, 0, K, st);
    launch_spec_gradient(h.dt, h.lnps[cur].p, 1, 1, h.specA.p, h.LpA, 2 * K, 2 * K + 1, st);
    h.launches += 2;
    h.mark("spec_gradient");
    dev_inverse(h, h.specA.p, h.LpA, h.levsA.p, 2 * K + 2, "_grad");
  }

  GridStepArgs ga;
  ga.u_cur = h.u[cur].p; ga.v_cur = h.v[cur].p; ga.t_cur = h.T[cur].p;
  ga.u_prev = h.u[prev].p; ga.v_prev = h.v[prev].p; ga.t_prev = h.T[prev].p;
  ga.vor_cur = h.vorg.p; ga.div_cur = h.divg.p; ga.ps_cur = h.ps[cur].p; ga.ps_prev = h.ps[prev].p; ga.phis = h.phis.p;
  ga.dx_t = h.gradA.p; ga.dy_t = h.gradA.p + (size_t)K * pl;
  ga.dx_lnps = h.gradA.p + (size_t)(2 * K) * pl; ga.dy_lnps = h.gradA.p + (size_t)(2 * K + 1) * pl;
  ga.dt_u_in = dtu_in; ga.dt_v_in = dtv_in; ga.dt_t_in = dtt_in;
  ga.out_T = h.gridB.p; ga.out_A = h.gridB.p + (size_t)K * pl; ga.out_B = h.gridB.p + (size_t)(2 * K) * pl;
  ga.out_phi = h.gridB.p + (size_t)(3 * K) * pl; ga.dt_lnps = h.gridB.p + (size_t)(4 * K) * pl;
  ga.wg_full = h.wg_full.p; ga.part = h.part.p;
  ga.scal = h.scal.p; ga.slot_cur = cur; ga.slot_prev = prev;
  ga.wg = (h.cfg.num_tracers > 0) ? h.wg.p : nullptr;
  launch_grid_step(h.dt, pr, ga, st); h.launches++;
  h.mark("grid_step");
  launch_reduce(h.part.p, pl, 2, h.ops_sum2.p, h.scal.p + SC_SUM_PS_PREV, h.red_tmp.p, st); h.launches += 2;
  h.mark("corr_reduce_prev");

  // ---- grid tracer: update_tracers (spectral_dynamics.F90:1116-1188); needs only the `current` winds and wg
  TracerArgs ta;
  if (h.cfg.num_tracers > 0) {
    const IscaConfig& c = h.cfg;
    ta.q_prev = h.q[prev].p; ta.q_cur = h.q[cur].p; ta.q_cur_w = h.q[cur].p; ta.q_fut = h.q[fut].p;
    ta.u_cur = h.u[cur].p; ta.v_cur = h.v[cur].p; ta.ps_cur = h.ps[cur].p; ta.ps_prev = h.ps[prev].p; ta.ps_fut = h.ps[fut].p;
    ta.wg = h.wg.p; ta.tr1 = h.tr1.p; ta.part = h.part.p; ta.wpart = h.tr_wpart.p;
    ta.delta_t = delta_t; ta.trflux = c.trflux;
    const double sink_s = c.trsink < 0. ? -86400. * c.trsink : c.trsink;       // hs_forcing.F90:408-409, 697-699
    ta.trdamp = sink_s > 0. ? 1. / sink_s : 0.;
    ta.robert_coeff = c.tracer_robert_coeff < 0. ? c.robert_coeff : c.tracer_robert_coeff;
    ta.raw_filter_coeff = c.raw_filter_coeff; ta.water_limit = c.water_correction_limit; ta.physics_on = physics_on;
    ta.dt_q_in = dtq_in;
    ta.halo_s = ta.halo_n = ta.send_s = ta.send_n = nullptr;
    if (g.P > 1) {
      const size_t hb = (size_t)3 * K * 2 * g.I;
      double* p = h.tr_halo.p;
      ta.halo_s = p; ta.halo_n = p + hb; ta.send_s = p + 2 * hb; ta.send_n = p + 3 * hb;
      if (h.p2p) {                         // receive in the tail of the lat-owner buffer, send = the neighbours' receive buffers
        const size_t o = h.sync_off + H::SYNC_FLAGS;
        ta.halo_s = h.fourB.p + o; ta.halo_n = h.fourB.p + o + hb;
        if (g.rank > 0) ta.send_s = h.peerB_host[g.rank - 1] + o + hb;          // my southern rows are the northern halo of rank - 1
        if (g.rank < g.P - 1) ta.send_n = h.peerB_host[g.rank + 1] + o;         // my northern rows are the southern halo of rank + 1
      }
    }
    exchange_tracer_halo(h, ta);
    launch_tracer_horiz(h.dt, h.fv, pr, ta, st);
    h.mark("tracer_horiz");
    launch_tracer_ppm(h.dt, pr, ta, st);
    h.mark("tracer_ppm");
    launch_reduce(h.part.p, pl, 1, h.ops_sum1.p, h.scal.p + SC_W_PREV, h.red_tmp.p, st);   // summed over the ranks with the fixers' sums
    h.launches += 4;
    h.mark("tracer_reduce");
  }

  dev_forward(h, h.levsB.p, 4 * K + 1, h.specB.p, h.LpB, h.truncB.p, "_tend");

  SpecStepArgs sa;
  sa.specB = h.specB.p; sa.LpB = h.LpB; sa.oT = 0; sa.oA = K; sa.oB = 2 * K; sa.oPhi = 3 * K; sa.oLnps = 4 * K;
  sa.vors_prev = h.vors[prev].p; sa.divs_prev = h.divs[prev].p; sa.ts_prev = h.ts[prev].p; sa.lnps_prev = h.lnps[prev].p;
  sa.vors_cur = h.vors[cur].p; sa.divs_cur = h.divs[cur].p; sa.ts_cur = h.ts[cur].p; sa.lnps_cur = h.lnps[cur].p;
  sa.vors_cur_w = h.vors[cur].p; sa.divs_cur_w = h.divs[cur].p; sa.ts_cur_w = h.ts[cur].p; sa.lnps_cur_w = h.lnps[cur].p;
  sa.vors_fut = h.vors[fut].p; sa.divs_fut = h.divs[fut].p; sa.ts_fut = h.ts[fut].p; sa.lnps_fut = h.lnps[fut].p;
  sa.dt_vors = h.dt_vors.p; sa.w_div = h.w_div.p; sa.w_T = h.w_T.p; sa.w_lnps = h.w_lnps.p;
  sa.specC = h.specC.p; sa.LpC = h.LpC; sa.cVor = 0; sa.cDiv = K; sa.cU = 2 * K; sa.cV = 3 * K; sa.cT = 4 * K; sa.cLnps = 5 * K;
  sa.cDxT = 5 * K + 1; sa.cDyT = 6 * K + 1; sa.cDxL = 7 * K + 1; sa.cDyL = 7 * K + 2;
  sa.fuse_robert_b = 1;
  sa.use_implicit = h.cfg.use_implicit;
  sa.keep_tend = h.keep_tend; sa.k_dt_vors = h.k_dt_vors.p; sa.k_dt_divs = h.k_dt_divs.p; sa.k_dt_ts = h.k_dt_ts.p;
  sa.k_dt_lnps = h.k_dt_lnps.p;
  launch_spec_step(h.dt, pr, sa, st); h.launches += (h.cfg.use_implicit ? 4 : 3);
  h.mark("spec_step");

  dev_inverse(h, h.specC.p, h.LpC, h.levsC[fut].p, 7 * K + 3, "_state");
  h.grad_valid = true;

  // compute_corrections (spectral_dynamics.F90:1213-1302): one column pass, one reduction, ONE all-reduce (the previous-level sums of
  // initialize_corrections, the previous-level water and the water integrals of the future level ride in the same SUM), one apply
  launch_colsum_fixers(h.dt, pr, h.u[fut].p, h.v[fut].p, h.T[fut].p, h.ps[fut].p, h.cfg.num_tracers > 0 ? h.tr_wpart.p : nullptr, h.part.p, st);
  launch_reduce(h.part.p, pl, 9, h.ops_fix.p, h.scal.p + SC_SUM_PS_FUT, h.red_tmp.p, st);
  allreduce_scalars(h, h.scal.p, SC_NSUM, NCCL_SUM);
  const double rc_raw = h.cfg.robert_coeff * h.cfg.raw_filter_coeff;
  launch_apply_fixers(h.dt, pr, fut, h.ps[fut].p, h.lnps[fut].p, h.lnps[cur].p, h.ts[fut].p, h.ts[cur].p, rc_raw, h.scal.p, h.denom(),
                      h.owns_m0(), h.cfg.do_mass_correction, h.cfg.do_energy_correction, st);
  h.launches += 4;
  h.mark("corr_mass_energy");
  if (h.cfg.num_tracers > 0) {
    launch_tracer_water_apply(h.dt, pr, ta, h.scal.p, h.denom(), h.cfg.do_water_correction, st);
    h.launches += 1;
    h.mark("tracer_water_fixer");
  }

  // time-level swap.  complete_robert_filter -> leapfrog_2level_B (a(previous) += rc*a(current)*raw) is fused into
  // spec_update (all coefficients) and apply_fixers (the fixers' (0,0) increments).
  h.previous = cur; h.current = fut;
  h.steps++;
}

static void check_t_flag(H& h) {
  double flag[1];
  CK(cudaMemcpyAsync(flag, h.scal.p + SC_T_FLAG, sizeof(double), cudaMemcpyDeviceToHost, h.st));
  CK(cudaStreamSynchronize(h.st));
  if (flag[0] != 0.0) throw std::runtime_error("spectral_dynamics: temperatures out of valid range");
}

// ---------------------------------------------------------------------------------------------
// transforms_mod-level helpers on host arrays (reference rectangular layout)
// ---------------------------------------------------------------------------------------------
static void x_prepare(H& h, int nlev, int nfields) {
  const Geometry& g = h.g;
  const int Lp = round_up(nlev * nfields, 16);
  h.x_rect.ensure((size_t)nlev * nfields * (g.N + 1) * (g.M + 1));
  h.x_spec.ensure((size_t)g.T * Lp);
  h.x_grid.ensure((size_t)nlev * nfields * h.nplane());
  h.ensure_four(Lp);
  h.x_levs.ensure(nlev * nfields);
  h.x_trunc.ensure(Lp);
}
static void x_set_levs(H& h, int ntot, int op_from, int op) {
  std::vector<LevDesc> l(ntot);
  for (int i = 0; i < ntot; ++i) l[i] = {h.x_grid.p + (size_t)i * h.nplane(), (i >= op_from) ? op : 0, 0};
  CK(cudaMemcpyAsync(h.x_levs.p, l.data(), sizeof(LevDesc) * ntot, cudaMemcpyHostToDevice, h.st));
  CK(cudaStreamSynchronize(h.st));
}

// ---------------------------------------------------------------------------------------------
// cold start (spectral_initialize_fields.F90:45-135)
// ---------------------------------------------------------------------------------------------
static void cold_start(H& h) {
  const Geometry& g = h.g; const int K = g.K; const IscaConfig& c = h.cfg;
  const size_t pl = h.nplane();
  cudaStream_t st = h.st;
  x_prepare(h, K, 2);
  const int Lp = round_up(2 * K, 16);
  // initial vorticity perturbation (:87-107)
  std::vector<double2> sp((size_t)g.T * Lp, make_double2(0, 0));
  const int pert_m[4] = {1, 5, 1, 5}, pert_n[4] = {3, 3, 2, 2};
  for (int q = 0; q < 4; ++q) {
    int m = pert_m[q], n = pert_n[q];
    if (m > g.M || n > g.N || g.owner[m] != g.rank) continue;
    int mi = (int)(std::find(g.m_of.begin(), g.m_of.end(), m) - g.m_of.begin());
    if (n >= g.M - m + 2) continue;
    for (int k = K - 3; k < K; ++k) if (k >= 0) sp[(size_t)(g.off[mi] + n) * Lp + k].x = 1.e-7;
  }
  h2d_on(h.st, h.x_spec.p, sp.data(), sp.size() * sizeof(double2));
  // uv_grid_from_vor_div(vors, 0): levels [vor | div | ucos | vcos] needs 4K levels -> reuse specC
  CK(cudaMemsetAsync(h.specC.p, 0, h.specC.n * sizeof(double2), st));
  // copy vor into specC[.., 0..K)
  {
    std::vector<double2> sc((size_t)g.T * h.LpC, make_double2(0, 0));
    for (int p = 0; p < g.T; ++p) for (int k = 0; k < K; ++k) sc[(size_t)p * h.LpC + k] = sp[(size_t)p * Lp + k];
    h2d_on(h.st, h.specC.p, sc.data(), sc.size() * sizeof(double2));
  }
  // T = initial_temperature, ln ps = log(p0) - phis/(rd*T0)  -> spectral and back (:113-119)
  {
    std::vector<double> tg(h.n3(), c.initial_temperature), phis(pl), lnps(pl);
    d2h_on(h.st, phis.data(), h.phis.p, pl * sizeof(double));
    for (size_t i = 0; i < pl; ++i) lnps[i] = std::log(c.reference_sea_level_press) - phis[i] / (c.rdgas * c.initial_temperature);
    h2d_on(h.st, h.T[0].p, tg.data(), tg.size() * sizeof(double));
    h2d_on(h.st, h.ps[0].p, lnps.data(), pl * sizeof(double));
  }
  // forward: [T(K) | lnps(1)] with truncation
  {
    std::vector<LevDesc> l(K + 1);
    for (int k = 0; k < K; ++k) l[k] = {h.T[0].p + (size_t)k * pl, 0, 0};
    l[K] = {h.ps[0].p, 0, 0};
    DBuf<LevDesc> dl; dl.upload(l);
    const int LpT = round_up(K + 1, 16);
    DBuf<double2> tmp; tmp.alloc((size_t)g.T * LpT);
    DBuf<unsigned char> tr; tr.upload(std::vector<unsigned char>(LpT, 1));
    dev_forward(h, dl.p, K + 1, tmp.p, LpT, tr.p);
    // back to grid: T plain, lnps -> exp
    l[K].op = 2; dl.upload(l);
    dev_inverse(h, tmp.p, LpT, dl.p, K + 1);
    // store spectral ts, ln_ps (slot 0)
    std::vector<double2> ht_((size_t)g.T * LpT);
    CK(cudaStreamSynchronize(st));
    d2h_on(h.st, ht_.data(), tmp.p, ht_.size() * sizeof(double2));
    std::vector<double2> ts((size_t)g.T * K), ln((size_t)g.T);
    for (int p = 0; p < g.T; ++p) { for (int k = 0; k < K; ++k) ts[(size_t)p * K + k] = ht_[(size_t)p * LpT + k]; ln[p] = ht_[(size_t)p * LpT + K]; }
    h2d_on(h.st, h.ts[0].p, ts.data(), ts.size() * sizeof(double2));
    h2d_on(h.st, h.lnps[0].p, ln.data(), ln.size() * sizeof(double2));
  }
  // u, v from the perturbation vorticity; then vor/div from (u,v); then u,v,vorg,divg from those (:108-123)
  {
    launch_spec_ucos_vcos(h.dt, h.specC.p, h.LpC, K, 0, K, 2 * K, 3 * K, st);
    dev_inverse(h, h.specC.p, h.LpC, h.levsC[0].p, 4 * K);      // vorg, divg, u[0], v[0] (op 1 = /cos)
    // vor_div_from_uv_grid: divide_by_cos, forward without truncation, alpha operators, truncate
    launch_divide_by_cos(h.dt, h.u[0].p, K, st);
    launch_divide_by_cos(h.dt, h.v[0].p, K, st);
    std::vector<LevDesc> l(2 * K);
    for (int k = 0; k < K; ++k) { l[k] = {h.u[0].p + (size_t)k * pl, 0, 0}; l[K + k] = {h.v[0].p + (size_t)k * pl, 0, 0}; }
    DBuf<LevDesc> dl; dl.upload(l);
    DBuf<unsigned char> tr; tr.upload(std::vector<unsigned char>(Lp, 0));
    dev_forward(h, dl.p, 2 * K, h.x_spec.p, Lp, tr.p);
    launch_spec_vor_div(h.dt, h.x_spec.p, Lp, K, 0, K, h.specC.p, h.LpC, 0, K, st);
    launch_spec_ucos_vcos(h.dt, h.specC.p, h.LpC, K, 0, K, 2 * K, 3 * K, st);
    dev_inverse(h, h.specC.p, h.LpC, h.levsC[0].p, 4 * K);
    CK(cudaStreamSynchronize(st));
    std::vector<double2> sc((size_t)g.T * h.LpC);
    d2h_on(h.st, sc.data(), h.specC.p, sc.size() * sizeof(double2));
    std::vector<double2> vo((size_t)g.T * K), di((size_t)g.T * K);
    for (int p = 0; p < g.T; ++p) for (int k = 0; k < K; ++k) { vo[(size_t)p * K + k] = sc[(size_t)p * h.LpC + k]; di[(size_t)p * K + k] = sc[(size_t)p * h.LpC + K + k]; }
    h2d_on(h.st, h.vors[0].p, vo.data(), vo.size() * sizeof(double2));
    h2d_on(h.st, h.divs[0].p, di.data(), di.size() * sizeof(double2));
  }
  // both time levels identical (spectral_dynamics.F90:616-624)
  d2d_on(h.st, h.vors[1].p, h.vors[0].p, h.nspec3() * sizeof(double2));
  d2d_on(h.st, h.divs[1].p, h.divs[0].p, h.nspec3() * sizeof(double2));
  d2d_on(h.st, h.ts[1].p, h.ts[0].p, h.nspec3() * sizeof(double2));
  d2d_on(h.st, h.lnps[1].p, h.lnps[0].p, (size_t)g.T * sizeof(double2));
  d2d_on(h.st, h.u[1].p, h.u[0].p, h.n3() * sizeof(double));
  d2d_on(h.st, h.v[1].p, h.v[0].p, h.n3() * sizeof(double));
  d2d_on(h.st, h.T[1].p, h.T[0].p, h.n3() * sizeof(double));
  d2d_on(h.st, h.ps[1].p, h.ps[0].p, pl * sizeof(double));
  if (h.cfg.num_tracers > 0) {
    std::vector<double> q0(h.n3(), c.initial_sphum);                // spectral_dynamics.F90:584-590
    for (int s2 = 0; s2 < 2; ++s2) h2d_on(st, h.q[s2].p, q0.data(), q0.size() * sizeof(double));
  }
  CK(cudaMemsetAsync(h.scal.p + SC_TSHIFT0, 0, 2 * sizeof(double), st));
  CK(cudaStreamSynchronize(st));
  h.previous = 0; h.current = 0;
  h.grad_valid = false;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// stage-level ABI helpers (single rank): reference Fourier layout (0:M, lat, lev) complex <-> the library's Fourier buffer
// [(pos[m]*J + j)][2*Lp]
// ---------------------------------------------------------------------------------------------
__global__ void four_layout_kernel(GeomDev g, double2* ref, double* four, int nlev, int Lp, int to_internal) {
  const size_t n = (size_t)nlev * g.J * (g.M + 1);
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int m = (int)(idx % (g.M + 1));
  const int j = (int)((idx / (g.M + 1)) % g.J);
  const int lev = (int)(idx / ((size_t)(g.M + 1) * g.J));
  double2* f = reinterpret_cast<double2*>(four + ((size_t)g.pos[m] * g.J + j) * 2 * Lp) + lev;
  if (to_internal) *f = ref[idx]; else ref[idx] = *f;
}
// derived fields of spectral_diagnostics (spectral_dynamics.F90:1747-1821): mode 0: a*b, 1: sqrt(a^2 + b^2)
__global__ void derived_product_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, size_t n, int mode) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = mode ? sqrt(a[i] * a[i] + b[i] * b[i]) : a[i] * b[i];
}
// sea-level pressure (spectral_dynamics.F90:1823-1835): first level with p_full/p_surf > 0.8; -1 flags a column without one
__global__ void slp_kernel(double* __restrict__ slp, const double* __restrict__ t, const double* __restrict__ p_full, const double* __restrict__ ps,
                           const double* __restrict__ phis, size_t nc, int K, double expf, double gamma, double grav, int* __restrict__ err) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const double p0 = ps[i];
  int k = 0;
  while (k < K && !(p_full[(size_t)k * nc + i] / p0 > 0.8)) ++k;
  if (k == K) { *err = 1; slp[i] = 0.0; return; }
  const double t_low = t[(size_t)k * nc + i] * pow(p_full[(size_t)k * nc + i] / p0, -expf);
  slp[i] = p0 * pow((t_low + gamma * phis[i] / grav) / t_low, 1.0 / expf);
}
__global__ void diag_axpy_kernel(double* __restrict__ acc, const double* __restrict__ x, size_t n, double scale, int accumulate) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) acc[i] = accumulate ? acc[i] + x[i] : x[i] * scale;
}

#define API_BEGIN(h) if (!(h)) return 1; try {
#define API_END(h) } catch (const std::exception& e) { (h)->err = e.what(); return 2; } return 0;

extern "C" {

void isca_b200_default_config(IscaConfig* c) {
  std::memset(c, 0, sizeof(*c));
  c->abi_version = ISCA_B200_ABI_VERSION;
  c->lon_max = 128; c->lat_max = 64; c->num_fourier = 42; c->num_spherical = 43; c->num_levels = 18; c->dt_atmos = 600.;
  c->damping_order = 2; c->damping_order_vor = -1; c->damping_order_div = -1;
  c->damping_coeff = 1.15740741e-4; c->damping_coeff_vor = -1.; c->damping_coeff_div = -1.;
  c->do_mass_correction = 1; c->do_energy_correction = 1; c->do_water_correction = 1;
  c->use_virtual_temperature = 0; c->use_implicit = 1; c->make_symmetric = 0;
  c->robert_coeff = .04; c->raw_filter_coeff = 1.0; c->alpha_implicit = .5;
  c->vert_coord_option = 0; c->scale_heights = 4.; c->surf_res = .1; c->exponent = 2.5; c->p_press = .1; c->p_sigma = .3;
  c->reference_sea_level_press = 101325.; c->initial_sphum = 0.; c->water_correction_limit = 0.;
  c->valid_range_t[0] = 100.; c->valid_range_t[1] = 500.; c->initial_temperature = 264.;
  c->num_tracers = 0; c->tracer_robert_coeff = -1.;
  c->no_forcing = 0; c->do_conserve_energy = 1;
  c->t_zero = 315.; c->t_strat = 200.; c->delh = 60.; c->delv = 10.; c->eps = 0.; c->sigma_b = 0.7; c->P00 = 1.e5;
  c->ka = -40.; c->ks = -4.; c->kf = -1.; c->trflux = 1.e-5; c->trsink = -4.;
  c->radius = 6376.0e3; c->omega = 7.2921150e-5; c->grav = 9.80; c->rdgas = 287.04; c->kappa = 2. / 7.;
}

const char* isca_b200_last_error(IscaHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int isca_b200_nccl_unique_id(void* out128) {
  try {
    static NcclApi api;
    api.load();
    NcclUniqueId id;
    api.ck(api.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out128, id.internal, 128);
  } catch (const std::exception& e) { g_create_error = e.what(); return 3; }
  return 0;
}

int isca_b200_ipc_handles(IscaHandle h, void* out128) {
  API_BEGIN(h)
  if (h->g.P < 2) throw std::runtime_error("peer access needs nranks > 1");
  cudaIpcMemHandle_t ha, hb;
  CK(cudaIpcGetMemHandle(&ha, h->four.p));
  CK(cudaIpcGetMemHandle(&hb, h->fourB.p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(out128, &ha, 64); std::memcpy((char*)out128 + 64, &hb, 64);
  API_END(h)
}

int isca_b200_set_peer_handles(IscaHandle h, const void* all_handles) {
  API_BEGIN(h)
  const int P = h->g.P;
  if (P < 2) throw std::runtime_error("peer access needs nranks > 1");
  std::vector<double*> pa(P, nullptr), pb(P, nullptr);
  for (int r = 0; r < P; ++r) {
    if (r == h->g.rank) { pa[r] = h->four.p; pb[r] = h->fourB.p; continue; }
    cudaIpcMemHandle_t ha, hb;
    std::memcpy(&ha, (const char*)all_handles + (size_t)r * 128, 64);
    std::memcpy(&hb, (const char*)all_handles + (size_t)r * 128 + 64, 64);
    void *qa = nullptr, *qb = nullptr;
    CK(cudaIpcOpenMemHandle(&qa, ha, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&qb, hb, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(qa); h->ipc_opened.push_back(qb);
    pa[r] = (double*)qa; pb[r] = (double*)qb;
  }
  h->d_peerA.upload(pa); h->d_peerB.upload(pb);
  h->peerA_host = pa; h->peerB_host = pb;
  h->dt.g.peerA = h->d_peerA.p; h->dt.g.peerB = h->d_peerB.p; h->dt.g.p2p = 1;
  h->p2p = true;
  h->bar_count.alloc(4);
  CK(cudaMemset(h->bar_count.p, 0, 4 * sizeof(unsigned long long)));
  for (auto& sg : h->graphs) if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; sg.uses = 0; }
  API_END(h)
}

int isca_b200_decomposition(const IscaConfig* cfg, int rank, int nranks, int* lat_start, int* lat_count, int* num_m,
                            int* m_list, int* owner, int* pos) {
  try {
    Geometry g;
    build_geometry(*cfg, rank, nranks, g);
    *lat_start = g.j0; *lat_count = g.Jloc; *num_m = g.nm;
    for (int i = 0; i < g.nm; ++i) m_list[i] = g.m_of[i];
    for (int m = 0; m <= g.M; ++m) { owner[m] = g.owner[m]; pos[m] = g.pos[m]; }
  } catch (const std::exception& e) { g_create_error = e.what(); return 2; }
  return 0;
}

int isca_b200_host_table(const IscaConfig* cfg, int id, double* host, int count, int* count_out) {
  try {
    if (!cfg) throw std::runtime_error("null argument");
    if (cfg->abi_version != ISCA_B200_ABI_VERSION) throw std::runtime_error("IscaConfig.abi_version mismatch");
    Geometry g;
    HostTables t;
    build_geometry(*cfg, 0, 1, g);
    build_tables(*cfg, g, t);
    std::vector<double> tmp;
    const std::vector<double>* v = nullptr;
    switch (id) {
      case ISCA_TB_SIN_LAT: v = &t.sin_lat; break;
      case ISCA_TB_WTS_LAT: v = &t.wts_lat; break;
      case ISCA_TB_DEG_LAT: v = &t.deg_lat; break;
      case ISCA_TB_DEG_LON: v = &t.deg_lon; break;
      case ISCA_TB_PK: v = &t.pk; break;
      case ISCA_TB_BK: v = &t.bk; break;
      case ISCA_TB_ROW_M: for (int r : t.row_m) tmp.push_back((double)g.m_of[r]); v = &tmp; break;     // zonal wavenumber of the row
      case ISCA_TB_ROW_N: tmp.assign(t.row_n.begin(), t.row_n.end()); v = &tmp; break;
      case ISCA_TB_LEGENDRE: v = &t.leg; break;
      case ISCA_TB_EIGEN_LAPLACIAN: v = &t.eigen; break;
      case ISCA_TB_DAMPING: v = &t.damping; break;
      case ISCA_TB_REF_T: v = &t.ref_t; break;
      case ISCA_TB_IMPLICIT_H: v = &t.h; break;
      case ISCA_TB_DIV_MAT: v = &t.div_mat; break;
      case ISCA_TB_WAVE_MATRIX: build_wave_matrices(*cfg, g, t, 2 * cfg->dt_atmos * cfg->alpha_implicit, tmp); v = &tmp; break;
      default: throw std::runtime_error("unknown table id");
    }
    if (count_out) *count_out = (int)v->size();
    if (host) {
      if ((size_t)count != v->size()) throw std::runtime_error("table size mismatch");
      std::memcpy(host, v->data(), v->size() * sizeof(double));
    }
  } catch (const std::exception& e) { g_create_error = e.what(); return 2; }
  return 0;
}

int isca_b200_create(const IscaConfig* cfg, int rank, int nranks, const void* nccl_unique_id, IscaHandle* out) {
  if (!cfg || !out) { g_create_error = "null argument"; return 1; }
  *out = nullptr;
  H* h = nullptr;
  try {
    if (cfg->abi_version != ISCA_B200_ABI_VERSION) throw std::runtime_error("IscaConfig.abi_version mismatch");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw std::runtime_error("no CUDA device: isca_b200 has no CPU fallback");
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("invalid rank / nranks");
    if (nranks > 1 && !nccl_unique_id) throw std::runtime_error("nranks > 1 needs the shared ncclUniqueId");
    // unsupported namelist values fail loudly
    if (cfg->raw_filter_coeff != 1.0) throw std::runtime_error("raw_filter_coeff /= 1 is not supported");
    if (cfg->vert_advect_uv != 0 || cfg->vert_advect_t != 0) throw std::runtime_error("only second_centered vertical advection of u,v,T is supported");
    if (cfg->use_virtual_temperature) throw std::runtime_error("use_virtual_temperature is not supported");
    if (cfg->num_tracers < 0 || cfg->num_tracers > 1) throw std::runtime_error("only 0 or 1 (grid, finite_volume_parabolic sphum) tracers are supported");
    if (cfg->num_tracers == 1 && nranks > 1 && cfg->lat_max / nranks < 4) throw std::runtime_error("the grid tracer needs at least 4 latitude rows per rank (2-row halos)");
    if (cfg->num_tracers == 1 && cfg->num_levels < 4) throw std::runtime_error("the PPM tracer advection needs num_levels >= 4");
    if (cfg->do_water_correction && cfg->num_tracers == 0) throw std::runtime_error("do_water_correction must be .false. in a dry model (spectral_dynamics.F90:1264)");
    if ((cfg->do_energy_correction || cfg->do_water_correction) && !cfg->do_mass_correction) throw std::runtime_error("energy/water correction requires mass correction (spectral_dynamics.F90:409-415)");
    h = new H();
    h->cfg = *cfg;
    CK(cudaSetDevice(rank % ndev));
    build_geometry(*cfg, rank, nranks, h->g);
    build_tables(*cfg, h->g, h->ht);
    h->cfg.pk = nullptr; h->cfg.bk = nullptr;
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    h->use_graph = (std::getenv("ISCA_B200_NO_GRAPH") == nullptr);
    CK(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
    for (auto& e : h->ev_pipe) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char* e = std::getenv("ISCA_B200_PIPE")) h->pipe_parts = std::max(1, std::atoi(e));
    if (nranks > 1) {
      h->nccl.load();
      NcclUniqueId id; std::memcpy(id.internal, nccl_unique_id, 128);
      h->nccl.ck(h->nccl.CommInitRank(&h->comm, nranks, id, rank), "ncclCommInitRank");
      // the multi-rank step (peer-memory transposes with device-side flag barriers + one ncclAllReduce) is replayed from a CUDA graph
      // like the single-rank one; ISCA_B200_NO_GRAPH_MULTI=1 issues it eagerly
      if (std::getenv("ISCA_B200_NO_GRAPH_MULTI") != nullptr) h->use_graph = false;
    }
    upload_tables(*h);
    set_params(*h);
    alloc_state(*h);
    CK(cudaStreamSynchronize(h->st));
    *out = h;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete h;
    return 2;
  }
  return 0;
}

int isca_b200_destroy(IscaHandle h) {
  if (!h) return 1;
  cudaStreamSynchronize(h->st);
  for (auto& kv : h->wave_cache) delete kv.second;
  for (auto& sg : h->graphs) if (sg.exec) cudaGraphExecDestroy(sg.exec);
  for (void* q : h->ipc_opened) cudaIpcCloseMemHandle(q);
  if (h->comm) h->nccl.CommDestroy(h->comm);
  if (h->st) cudaStreamDestroy(h->st);
  if (h->st2) cudaStreamDestroy(h->st2);
  for (auto& e : h->ev_pipe) if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

int isca_b200_cold_start(IscaHandle h) { API_BEGIN(h) cold_start(*h); CK(cudaStreamSynchronize(h->st)); API_END(h) }

int isca_b200_set_surf_geopotential(IscaHandle h, const double* sg) {
  API_BEGIN(h) h2d_on(h->st, h->phis.p, sg, h->nplane() * sizeof(double)); API_END(h)
}

int isca_b200_set_grid_state(IscaHandle h, int slot, const double* ug, const double* vg, const double* tg,
                             const double* psg, const double* tracers) {
  API_BEGIN(h)
  if (slot < 0 || slot > 1) throw std::runtime_error("slot must be 0 or 1");
  if (ug) h2d_on(h->st, h->u[slot].p, ug, h->n3() * sizeof(double));
  if (vg) h2d_on(h->st, h->v[slot].p, vg, h->n3() * sizeof(double));
  if (tg) {
    h2d_on(h->st, h->T[slot].p, tg, h->n3() * sizeof(double));
    CK(cudaMemsetAsync(h->scal.p + SC_TSHIFT0 + slot, 0, sizeof(double), h->st));
  }
  if (psg) h2d_on(h->st, h->ps[slot].p, psg, h->nplane() * sizeof(double));
  if (tracers) {
    if (h->cfg.num_tracers < 1) throw std::runtime_error("no tracer configured");
    h2d_on(h->st, h->q[slot].p, tracers, h->n3() * sizeof(double));
  }
  h->grad_valid = false;
  API_END(h)
}

static void set_spec(H& h, DBuf<double2>& dst, const double* src, int nlev) {
  const Geometry& g = h.g;
  h.x_rect.ensure((size_t)nlev * (g.N + 1) * (g.M + 1));
  h2d_on(h.st, h.x_rect.p, src, (size_t)nlev * (g.N + 1) * (g.M + 1) * sizeof(double2));
  launch_pack_spec(h.dt, h.x_rect.p, dst.p, nlev, nlev, 0, h.st);
  CK(cudaStreamSynchronize(h.st));
}
static void get_spec(H& h, const double2* src, int Ls, int lev0, double* dst, int nlev) {
  const Geometry& g = h.g;
  h.x_rect.ensure((size_t)nlev * (g.N + 1) * (g.M + 1));
  CK(cudaMemsetAsync(h.x_rect.p, 0, (size_t)nlev * (g.N + 1) * (g.M + 1) * sizeof(double2), h.st));
  launch_unpack_spec(h.dt, src, h.x_rect.p, nlev, Ls, lev0, h.st);
  CK(cudaStreamSynchronize(h.st));
  d2h_on(h.st, dst, h.x_rect.p, (size_t)nlev * (g.N + 1) * (g.M + 1) * sizeof(double2));
}

int isca_b200_set_spectral_state(IscaHandle h, int slot, const double* vors, const double* divs, const double* ts,
                                 const double* ln_ps) {
  API_BEGIN(h)
  if (slot < 0 || slot > 1) throw std::runtime_error("slot must be 0 or 1");
  if (vors) set_spec(*h, h->vors[slot], vors, h->g.K);
  if (divs) set_spec(*h, h->divs[slot], divs, h->g.K);
  if (ts) set_spec(*h, h->ts[slot], ts, h->g.K);
  if (ln_ps) set_spec(*h, h->lnps[slot], ln_ps, 1);
  h->grad_valid = false;
  API_END(h)
}

int isca_b200_set_vor_div_grid(IscaHandle h, const double* vorg, const double* divg) {
  API_BEGIN(h)
  if (vorg) h2d_on(h->st, h->vorg.p, vorg, h->n3() * sizeof(double));
  if (divg) h2d_on(h->st, h->divg.p, divg, h->n3() * sizeof(double));
  API_END(h)
}

int isca_b200_set_time_pointers(IscaHandle h, int previous_slot, int current_slot) {
  API_BEGIN(h)
  if (previous_slot < 0 || previous_slot > 1 || current_slot < 0 || current_slot > 1) throw std::runtime_error("slots must be 0 or 1");
  h->previous = previous_slot; h->current = current_slot;
  h->grad_valid = false;
  API_END(h)
}
int isca_b200_get_time_pointers(IscaHandle h, int* p, int* c) { API_BEGIN(h) *p = h->previous; *c = h->current; API_END(h) }

// One model step, replayed from a captured CUDA graph when possible.  The first (forward) step and
// the first occurrence of every variant run eagerly (they also perform lazy one-time setup).
static void step_graphed(H& h, int physics) {
  const bool first = (h.previous == h.current);
  if (!h.use_graph || first || h.keep_tend || !h.grad_valid) { step_once(h, physics, nullptr, nullptr, nullptr); return; }
  H::StepGraph& sg = h.graphs[h.current + 2 * (physics ? 1 : 0)];
  sg.uses++;
  if (sg.uses == 1) { step_once(h, physics, nullptr, nullptr, nullptr); return; }
  if (!sg.exec) {
    if (h.cfg.use_implicit) ensure_wave_matrix(h, 2 * h.cfg.dt_atmos * h.cfg.alpha_implicit);   // no uploads inside capture
    const int prev = h.previous, cur = h.current;
    const long long l0 = h.launches, s0 = h.steps;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(h.st, cudaStreamCaptureModeThreadLocal));
    try { step_once(h, physics, nullptr, nullptr, nullptr); }
    catch (...) { cudaStreamEndCapture(h.st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
    CK(cudaStreamEndCapture(h.st, &graph));
    CK(cudaGraphInstantiate(&sg.exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    sg.launches = h.launches - l0;
    h.launches = l0; h.steps = s0; h.previous = prev; h.current = cur;   // capture did not execute anything
  }
  CK(cudaGraphLaunch(sg.exec, h.st));
  const int cur = h.current;
  h.previous = cur; h.current = 1 - cur;
  h.launches += sg.launches; h.steps++;
}

static int run_steps(IscaHandle h, int n, int physics) {
  API_BEGIN(h)
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, h->st));
  for (int i = 0; i < n; ++i) step_graphed(*h, physics);
  CK(cudaEventRecord(e1, h->st));
  check_t_flag(*h);
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  h->last_step_ms = n > 0 ? ms / n : 0.0;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  API_END(h)
}
int isca_b200_step(IscaHandle h, int n_steps) { return run_steps(h, n_steps, 1); }
int isca_b200_step_dynamics_only(IscaHandle h, int n_steps) { return run_steps(h, n_steps, 0); }

int isca_b200_spectral_dynamics(IscaHandle h, const double* dt_psg, const double* dt_ug, const double* dt_vg,
                                const double* dt_tg, double* psg_final, double* ug_final, double* vg_final,
                                double* tg_final, double* wg_full, double* p_full) {
  return isca_b200_spectral_dynamics_tracers(h, dt_psg, dt_ug, dt_vg, dt_tg, nullptr, psg_final, ug_final, vg_final, tg_final, nullptr,
                                             wg_full, p_full);
}

int isca_b200_spectral_dynamics_tracers(IscaHandle h, const double* dt_psg, const double* dt_ug, const double* dt_vg,
                                        const double* dt_tg, const double* dt_tracers, double* psg_final, double* ug_final,
                                        double* vg_final, double* tg_final, double* grid_tracers_final, double* wg_full,
                                        double* p_full) {
  API_BEGIN(h)
  if (dt_psg) throw std::runtime_error("non-zero dt_psg is not supported (the solo driver always passes zero, atmosphere.F90:289)");
  if ((dt_tracers || grid_tracers_final) && h->cfg.num_tracers < 1) throw std::runtime_error("no tracer configured");
  const size_t n3 = h->n3();
  const int nf = dt_tracers ? 4 : 3;
  h->ext_tend.ensure(nf * n3);
  const double* src[4] = {dt_ug, dt_vg, dt_tg, dt_tracers};
  for (int f = 0; f < nf; ++f) {
    if (src[f]) CK(cudaMemcpyAsync(h->ext_tend.p + f * n3, src[f], n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    else CK(cudaMemsetAsync(h->ext_tend.p + f * n3, 0, n3 * sizeof(double), h->st));
  }
  step_once(*h, 0, h->ext_tend.p, h->ext_tend.p + n3, h->ext_tend.p + 2 * n3, dt_tracers ? h->ext_tend.p + 3 * n3 : nullptr);
  const int c = h->current;
  if (psg_final) CK(cudaMemcpyAsync(psg_final, h->ps[c].p, h->nplane() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (ug_final) CK(cudaMemcpyAsync(ug_final, h->u[c].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (vg_final) CK(cudaMemcpyAsync(vg_final, h->v[c].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (tg_final) {
    launch_materialize_t(h->dt, h->T[c].p, h->scal.p, c, h->st);
    CK(cudaMemcpyAsync(tg_final, h->T[c].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  }
  if (grid_tracers_final) CK(cudaMemcpyAsync(grid_tracers_final, h->q[c].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (wg_full) CK(cudaMemcpyAsync(wg_full, h->wg_full.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (p_full) {
    h->x_grid.ensure(n3);
    launch_materialize_t(h->dt, h->T[h->previous].p, h->scal.p, h->previous, h->st);
    launch_press_heights(h->dt, h->pr, h->T[h->previous].p, h->ps[h->previous].p, h->phis.p, h->x_grid.p, nullptr, nullptr, nullptr, h->st);
    CK(cudaMemcpyAsync(p_full, h->x_grid.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  }
  check_t_flag(*h);
  API_END(h)
}

// ---- internal (C++) interface used by the moist-model driver (moist_model.cu); see core_internal.h
int isca_core_view(IscaHandle h, IscaCoreView* v) {
  API_BEGIN(h)
  for (int s = 0; s < 2; ++s) {
    launch_materialize_t(h->dt, h->T[s].p, h->scal.p, s, h->st);
    v->u[s] = h->u[s].p; v->v[s] = h->v[s].p; v->T[s] = h->T[s].p; v->q[s] = h->q[s].p; v->ps[s] = h->ps[s].p;
  }
  v->phis = h->phis.p; v->wg_full = h->wg_full.p; v->rad_lat = h->d_rad_lat.p + h->g.j0;
  v->I = h->g.I; v->Jloc = h->g.Jloc; v->K = h->g.K; v->j0 = h->g.j0; v->previous = h->previous; v->current = h->current;
  v->st = h->st; v->dt_atmos = h->cfg.dt_atmos; v->grav = h->cfg.grav; v->num_tracers = h->cfg.num_tracers; v->nranks = h->g.P;
  API_END(h)
}
int isca_core_press_heights(IscaHandle h, int slot, double* p_full, double* p_half, double* z_full, double* z_half) {
  API_BEGIN(h)                                             // T[slot] must be materialized (isca_core_view)
  launch_press_heights(h->dt, h->pr, h->T[slot].p, h->ps[slot].p, h->phis.p, p_full, p_half, z_full, z_half, h->st);
  API_END(h)
}
int isca_core_step_ext(IscaHandle h, const double* dtu, const double* dtv, const double* dtt, const double* dtq) {
  API_BEGIN(h)
  if (dtq && h->cfg.num_tracers < 1) throw std::runtime_error("no tracer configured");
  step_once(*h, 0, dtu, dtv, dtt, dtq);
  API_END(h)
}
int isca_core_check(IscaHandle h) {
  API_BEGIN(h)
  check_t_flag(*h);
  API_END(h)
}
void isca_core_profile_begin(IscaHandle h) {
  if (!h) return;
  for (auto& m : h->marks) cudaEventDestroy(m.second);
  h->marks.clear();
  h->profiling = true;
  h->mark("start");
}
void isca_core_mark(IscaHandle h, const char* name) { if (h) h->mark(name); }
int isca_core_profile_end(IscaHandle h, std::vector<std::string>& order, std::map<std::string, double>& acc) {
  if (!h) return 1;
  h->profiling = false;
  int rc = cudaStreamSynchronize(h->st) == cudaSuccess ? 0 : 1;
  for (size_t q = 1; q < h->marks.size() && !rc; ++q) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->marks[q - 1].second, h->marks[q].second) != cudaSuccess) { rc = 1; break; }
    if (!acc.count(h->marks[q].first)) order.push_back(h->marks[q].first);
    acc[h->marks[q].first] += ms;
  }
  for (auto& m : h->marks) cudaEventDestroy(m.second);
  h->marks.clear();
  return rc;
}

static int slot_of(H& h, int level) {
  if (level == ISCA_LEVEL_CURRENT) return h.current;
  if (level == ISCA_LEVEL_PREVIOUS) return h.previous;
  if (level == 0 || level == 1) return level;
  throw std::runtime_error("invalid time level selector");
}

// device pointer and element count of a grid field (materialised on the library's stream where needed)
static size_t field_on_device(H& h, int id, int level, const double** ptr) {
  const int s = slot_of(h, level);
  const size_t n3 = h.n3(), pl = h.nplane();
  switch (id) {
    case ISCA_F_PS: *ptr = h.ps[s].p; return pl;
    case ISCA_F_U: *ptr = h.u[s].p; return n3;
    case ISCA_F_V: *ptr = h.v[s].p; return n3;
    case ISCA_F_T: launch_materialize_t(h.dt, h.T[s].p, h.scal.p, s, h.st); *ptr = h.T[s].p; return n3;
    case ISCA_F_VOR: *ptr = h.vorg.p; return n3;
    case ISCA_F_DIV: *ptr = h.divg.p; return n3;
    case ISCA_F_WG_FULL: *ptr = h.wg_full.p; return n3;
    case ISCA_F_TRACER0:
      if (h.cfg.num_tracers < 1) throw std::runtime_error("no tracer configured");
      *ptr = h.q[s].p; return n3;
    case ISCA_F_P_FULL: case ISCA_F_P_HALF: case ISCA_F_Z_FULL: case ISCA_F_Z_HALF: {
      const size_t nh = n3 + pl;
      h.x_pz.ensure(nh);
      double* pf = (id == ISCA_F_P_FULL) ? h.x_pz.p : nullptr;
      double* ph = (id == ISCA_F_P_HALF) ? h.x_pz.p : nullptr;
      double* zf = (id == ISCA_F_Z_FULL) ? h.x_pz.p : nullptr;
      double* zh = (id == ISCA_F_Z_HALF) ? h.x_pz.p : nullptr;
      launch_materialize_t(h.dt, h.T[s].p, h.scal.p, s, h.st);
      launch_press_heights(h.dt, h.pr, h.T[s].p, h.ps[s].p, h.phis.p, pf, ph, zf, zh, h.st);
      *ptr = h.x_pz.p;
      return (id == ISCA_F_P_HALF || id == ISCA_F_Z_HALF) ? nh : n3;
    }
    default: break;
  }
  // ---- derived fields of spectral_diagnostics (spectral_dynamics.F90:1747-1835), current / requested level
  auto base = [&](int fid) { const double* q = nullptr; field_on_device(h, fid, level, &q); return q; };
  int fa = -1, fb = -1, mode = 0;
  switch (id) {
    case ISCA_F_WSPD: fa = ISCA_F_U; fb = ISCA_F_V; mode = 1; break;
    case ISCA_F_UU: fa = fb = ISCA_F_U; break;
    case ISCA_F_VV: fa = fb = ISCA_F_V; break;
    case ISCA_F_UV: fa = ISCA_F_U; fb = ISCA_F_V; break;
    case ISCA_F_V_VOR: fa = ISCA_F_V; fb = ISCA_F_VOR; break;
    case ISCA_F_TT: fa = fb = ISCA_F_T; break;
    case ISCA_F_OMEGA_OMEGA: fa = fb = ISCA_F_WG_FULL; break;
    case ISCA_F_OMEGA_T: fa = ISCA_F_WG_FULL; fb = ISCA_F_T; break;
    case ISCA_F_UW: fa = ISCA_F_U; fb = ISCA_F_WG_FULL; break;
    case ISCA_F_VW: fa = ISCA_F_WG_FULL; fb = ISCA_F_V; break;
    case ISCA_F_UT: fa = ISCA_F_U; fb = ISCA_F_T; break;
    case ISCA_F_VT: fa = ISCA_F_T; fb = ISCA_F_V; break;
    case ISCA_F_UZ: fa = ISCA_F_U; fb = ISCA_F_Z_FULL; break;
    case ISCA_F_VZ: fa = ISCA_F_V; fb = ISCA_F_Z_FULL; break;
    case ISCA_F_OMEGA_Z: fa = ISCA_F_WG_FULL; fb = ISCA_F_Z_FULL; break;
    case ISCA_F_UTR0: fa = ISCA_F_TRACER0; fb = ISCA_F_U; break;
    case ISCA_F_VTR0: fa = ISCA_F_TRACER0; fb = ISCA_F_V; break;
    case ISCA_F_WTR0: fa = ISCA_F_TRACER0; fb = ISCA_F_WG_FULL; break;
    case ISCA_F_SLP: {
      const double* t = base(ISCA_F_T);
      const double* pf = base(ISCA_F_P_FULL);                       // into x_pz
      h.x_der.ensure(pl + 1);
      int* flag = reinterpret_cast<int*>(h.x_der.p + pl);
      CK(cudaMemsetAsync(flag, 0, sizeof(int), h.st));
      const double gamma = 0.006, expf = h.cfg.rdgas * gamma / h.cfg.grav;          // spectral_dynamics.F90:1686-1688
      slp_kernel<<<(unsigned)((pl + 127) / 128), 128, 0, h.st>>>(h.x_der.p, t, pf, h.ps[s].p, h.phis.p, pl, h.g.K, expf, gamma, h.cfg.grav, flag);
      int e = 0;
      CK(cudaMemcpyAsync(&e, flag, sizeof(int), cudaMemcpyDeviceToHost, h.st));
      CK(cudaStreamSynchronize(h.st));
      if (e) throw std::runtime_error("spectral_diagnostics: No sigma values .gt. 0.8  Cannot compute slp");
      *ptr = h.x_der.p;
      return pl;
    }
    default: throw std::runtime_error("unknown field id");
  }
  const double* a = base(fa);
  const double* b = fb == fa ? a : base(fb);
  h.x_der.ensure(n3);
  derived_product_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, h.st>>>(h.x_der.p, a, b, n3, mode);
  *ptr = h.x_der.p;
  return n3;
}

int isca_core_field_device(IscaHandle h, int field_id, int level, const double** ptr, size_t* count) {
  API_BEGIN(h)
  *count = field_on_device(*h, field_id, level, ptr);
  API_END(h)
}

int isca_b200_get_field(IscaHandle h, int id, int level, double* host) {
  API_BEGIN(h)
  const double* src = nullptr;
  const size_t n = field_on_device(*h, id, level, &src);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, host, src, n * sizeof(double));
  API_END(h)
}

int isca_b200_get_spectral(IscaHandle h, int id, int level, double* host) {
  API_BEGIN(h)
  const int K = h->g.K;
  if (id >= 8 && id <= 11) {       // final spectral tendencies of the last step (keep_tend)
    const double2* src[4] = {h->k_dt_vors.p, h->k_dt_divs.p, h->k_dt_ts.p, h->k_dt_lnps.p};
    get_spec(*h, src[id - 8], id == 11 ? 1 : K, 0, host, id == 11 ? 1 : K);
  } else {
    const int s = slot_of(*h, level);
    switch (id) {
      case ISCA_S_VOR: get_spec(*h, h->vors[s].p, K, 0, host, K); break;
      case ISCA_S_DIV: get_spec(*h, h->divs[s].p, K, 0, host, K); break;
      case ISCA_S_T: get_spec(*h, h->ts[s].p, K, 0, host, K); break;
      case ISCA_S_LNPS: get_spec(*h, h->lnps[s].p, 1, 0, host, 1); break;
      default: throw std::runtime_error("unknown spectral field id");
    }
  }
  API_END(h)
}

int isca_b200_get_scalar(IscaHandle h, int id, double* value) {
  API_BEGIN(h)
  double sc[SC_COUNT];
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, sc, h->scal.p, sizeof(sc));
  switch (id) {
    case ISCA_SC_MEAN_PS: *value = sc[SC_MEAN_PS_PREV]; break;
    case ISCA_SC_MEAN_ENERGY: *value = sc[SC_MEAN_EN_PREV]; break;
    case ISCA_SC_T_MIN: *value = sc[SC_TMIN]; break;
    case ISCA_SC_T_MAX: *value = sc[SC_TMAX]; break;
    case ISCA_SC_STEP_COUNT: *value = (double)h->steps; break;
    case ISCA_SC_KERNEL_LAUNCHES: *value = (double)h->launches; break;
    case ISCA_SC_LAST_STEP_MS: *value = h->last_step_ms; break;
    case 100: h->keep_tend = 1; *value = 1; break;      // enable tendency capture (tests)
    default: throw std::runtime_error("unknown scalar id");
  }
  API_END(h)
}

int isca_b200_get_table(IscaHandle h, int id, double* host, int count) {
  API_BEGIN(h)
  const std::vector<double>* v = nullptr;
  switch (id) {
    case ISCA_TB_SIN_LAT: v = &h->ht.sin_lat; break;
    case ISCA_TB_WTS_LAT: v = &h->ht.wts_lat; break;
    case ISCA_TB_DEG_LAT: v = &h->ht.deg_lat; break;
    case ISCA_TB_DEG_LON: v = &h->ht.deg_lon; break;
    case ISCA_TB_PK: v = &h->ht.pk; break;
    case ISCA_TB_BK: v = &h->ht.bk; break;
    default: throw std::runtime_error("unknown table id");
  }
  if ((size_t)count != v->size()) throw std::runtime_error("table size mismatch");
  std::memcpy(host, v->data(), v->size() * sizeof(double));
  API_END(h)
}

// ---- transforms_mod level ---------------------------------------------------------------------
int isca_b200_spherical_to_grid(IscaHandle h, const double* spec, double* grid, int nlev) {
  API_BEGIN(h)
  const Geometry& g = h->g;
  x_prepare(*h, nlev, 1);
  const int Lp = round_up(nlev, 16);
  h2d_on(h->st, h->x_rect.p, spec, (size_t)nlev * (g.N + 1) * (g.M + 1) * sizeof(double2));
  CK(cudaMemsetAsync(h->x_spec.p, 0, (size_t)g.T * Lp * sizeof(double2), h->st));
  launch_pack_spec(h->dt, h->x_rect.p, h->x_spec.p, nlev, Lp, 0, h->st);
  x_set_levs(*h, nlev, 0, 0);
  dev_inverse(*h, h->x_spec.p, Lp, h->x_levs.p, nlev);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, grid, h->x_grid.p, (size_t)nlev * h->nplane() * sizeof(double));
  API_END(h)
}

int isca_b200_grid_to_spherical(IscaHandle h, const double* grid, double* spec, int nlev, int do_truncation) {
  API_BEGIN(h)
  x_prepare(*h, nlev, 1);
  const int Lp = round_up(nlev, 16);
  h2d_on(h->st, h->x_grid.p, grid, (size_t)nlev * h->nplane() * sizeof(double));
  CK(cudaMemsetAsync(h->x_trunc.p, do_truncation ? 1 : 0, Lp, h->st));
  x_set_levs(*h, nlev, 0, 0);
  dev_forward(*h, h->x_levs.p, nlev, h->x_spec.p, Lp, h->x_trunc.p);
  get_spec(*h, h->x_spec.p, Lp, 0, spec, nlev);
  API_END(h)
}

// ---- stage-level entry points (secondary ABI of SURVEY section 8b; single rank) -----------------------------------------
static void stage_prepare(H& h, int nlev) {
  if (h.g.P != 1) throw std::runtime_error("the stage-level entry points (fft_*, legendre_*) are single-rank: with nranks > 1 the Fourier transpose sits between them");
  if (nlev < 1) throw std::runtime_error("nlev must be positive");
  x_prepare(h, nlev, 1);
  h.x_fr.ensure((size_t)nlev * h.g.J * (h.g.M + 1));
}
static void four_convert(H& h, int nlev, int Lp, int to_internal) {
  const size_t n = (size_t)nlev * h.g.J * (h.g.M + 1);
  four_layout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h.st>>>(h.dt.g, h.x_fr.p, h.four.p, nlev, Lp, to_internal);
}

int isca_b200_fft_c2r(IscaHandle h, const double* fourier, double* grid, int nlev) {
  API_BEGIN(h)
  stage_prepare(*h, nlev);
  const int Lp = round_up(nlev, 16);
  const size_t nf = (size_t)nlev * h->g.J * (h->g.M + 1);
  h2d_on(h->st, h->x_fr.p, fourier, nf * sizeof(double2));
  CK(cudaMemsetAsync(h->four.p, 0, (size_t)(h->g.M + 1) * h->g.J * 2 * Lp * sizeof(double), h->st));
  four_convert(*h, nlev, Lp, 1);
  x_set_levs(*h, nlev, 0, 0);
  launch_fft_inv(h->dt, h->four.p, h->x_levs.p, nlev, Lp, h->st);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, grid, h->x_grid.p, (size_t)nlev * h->nplane() * sizeof(double));
  API_END(h)
}

int isca_b200_fft_r2c(IscaHandle h, const double* grid, double* fourier, int nlev) {
  API_BEGIN(h)
  stage_prepare(*h, nlev);
  const int Lp = round_up(nlev, 16);
  const size_t nf = (size_t)nlev * h->g.J * (h->g.M + 1);
  h2d_on(h->st, h->x_grid.p, grid, (size_t)nlev * h->nplane() * sizeof(double));
  x_set_levs(*h, nlev, 0, 0);
  launch_fft_fwd(h->dt, h->four.p, h->x_levs.p, nlev, Lp, h->st);
  four_convert(*h, nlev, Lp, 0);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, fourier, h->x_fr.p, nf * sizeof(double2));
  API_END(h)
}

int isca_b200_legendre_inv(IscaHandle h, const double* spec, double* fourier, int nlev) {
  API_BEGIN(h)
  stage_prepare(*h, nlev);
  const Geometry& g = h->g;
  const int Lp = round_up(nlev, 16);
  h2d_on(h->st, h->x_rect.p, spec, (size_t)nlev * (g.N + 1) * (g.M + 1) * sizeof(double2));
  CK(cudaMemsetAsync(h->x_spec.p, 0, (size_t)g.T * Lp * sizeof(double2), h->st));
  launch_pack_spec(h->dt, h->x_rect.p, h->x_spec.p, nlev, Lp, 0, h->st);
  launch_legendre_inv(h->dt, h->x_spec.p, h->four.p, Lp, h->st);
  four_convert(*h, nlev, Lp, 0);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, fourier, h->x_fr.p, (size_t)nlev * g.J * (g.M + 1) * sizeof(double2));
  API_END(h)
}

int isca_b200_legendre_fwd(IscaHandle h, const double* fourier, double* spec, int nlev, int do_truncation) {
  API_BEGIN(h)
  stage_prepare(*h, nlev);
  const Geometry& g = h->g;
  const int Lp = round_up(nlev, 16);
  h2d_on(h->st, h->x_fr.p, fourier, (size_t)nlev * g.J * (g.M + 1) * sizeof(double2));
  CK(cudaMemsetAsync(h->four.p, 0, (size_t)(g.M + 1) * g.J * 2 * Lp * sizeof(double), h->st));
  four_convert(*h, nlev, Lp, 1);
  CK(cudaMemsetAsync(h->x_trunc.p, do_truncation ? 1 : 0, Lp, h->st));
  launch_legendre_fwd(h->dt, h->four.p, h->x_spec.p, Lp, h->x_trunc.p, h->st);
  get_spec(*h, h->x_spec.p, Lp, 0, spec, nlev);
  API_END(h)
}

int isca_b200_implicit_correction(IscaHandle h, double* dt_divs, double* dt_ts, double* dt_ln_ps, const double* divs_prev,
                                  const double* divs_cur, const double* ts_prev, const double* ts_cur, const double* ln_ps_prev,
                                  const double* ln_ps_cur, double delta_t) {
  API_BEGIN(h)
  if (!h->cfg.use_implicit) throw std::runtime_error("implicit_correction: use_implicit is .false.");
  if (!dt_divs || !dt_ts || !dt_ln_ps || !divs_prev || !divs_cur || !ts_prev || !ts_cur || !ln_ps_prev || !ln_ps_cur)
    throw std::runtime_error("implicit_correction: null array");
  const Geometry& g = h->g; const int K = g.K;
  const double* src[9] = {dt_divs, dt_ts, dt_ln_ps, divs_prev, divs_cur, ts_prev, ts_cur, ln_ps_prev, ln_ps_cur};
  const int lev[9] = {K, K, 1, K, K, K, K, 1, 1};
  for (int i = 0; i < 9; ++i) { h->x_imp[i].ensure((size_t)g.T * lev[i]); set_spec(*h, h->x_imp[i], src[i], lev[i]); }
  Params pr = h->pr;
  pr.xi = delta_t * h->cfg.alpha_implicit;                    // implicit.F90:260-264: matrices follow the step passed in
  ensure_wave_matrix(*h, pr.xi);
  launch_implicit_correction(h->dt, pr, h->x_imp[0].p, h->x_imp[1].p, h->x_imp[2].p, h->x_imp[3].p, h->x_imp[4].p, h->x_imp[5].p,
                             h->x_imp[6].p, h->x_imp[7].p, h->x_imp[8].p, h->st);
  get_spec(*h, h->x_imp[0].p, K, 0, dt_divs, K);
  get_spec(*h, h->x_imp[1].p, K, 0, dt_ts, K);
  get_spec(*h, h->x_imp[2].p, 1, 0, dt_ln_ps, 1);
  API_END(h)
}

// ---- device-side time averages for diag_manager (boundary list of 8b) ----------------------------
int isca_b200_diag_accumulate(IscaHandle h, int field_id) {
  API_BEGIN(h)
  const double* src = nullptr;
  const size_t n = field_on_device(*h, field_id, ISCA_LEVEL_CURRENT, &src);
  H::DiagAcc& a = h->diag[field_id];
  if (!a.sum) { a.sum = new DBuf<double>(); a.sum->alloc(n); a.n = n; a.count = 0; CK(cudaMemsetAsync(a.sum->p, 0, n * sizeof(double), h->st)); }
  diag_axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(a.sum->p, src, n, 1.0, 1);
  a.count++;
  API_END(h)
}

int isca_b200_diag_fetch(IscaHandle h, int field_id, double* host, int reset, int* count_out) {
  API_BEGIN(h)
  auto it = h->diag.find(field_id);
  if (it == h->diag.end() || it->second.count == 0) throw std::runtime_error("diag_fetch: nothing accumulated for this field");
  H::DiagAcc& a = it->second;
  h->x_grid.ensure(a.n);
  diag_axpy_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, h->st>>>(h->x_grid.p, a.sum->p, a.n, 1.0 / (double)a.count, 0);
  CK(cudaStreamSynchronize(h->st));
  d2h_on(h->st, host, h->x_grid.p, a.n * sizeof(double));
  if (count_out) *count_out = (int)a.count;
  if (reset) { CK(cudaMemsetAsync(a.sum->p, 0, a.n * sizeof(double), h->st)); a.count = 0; }
  API_END(h)
}

int isca_b200_uv_grid_from_vor_div(IscaHandle h, const double* vors, const double* divs, double* ug, double* vg, int nlev) {
  API_BEGIN(h)
  const Geometry& g = h->g;
  x_prepare(*h, nlev, 4);
  const int Lp = round_up(4 * nlev, 16);
  const size_t nr = (size_t)nlev * (g.N + 1) * (g.M + 1);
  CK(cudaMemsetAsync(h->x_spec.p, 0, (size_t)g.T * Lp * sizeof(double2), h->st));
  h2d_on(h->st, h->x_rect.p, vors, nr * sizeof(double2));
  launch_pack_spec(h->dt, h->x_rect.p, h->x_spec.p, nlev, Lp, 0, h->st);
  CK(cudaStreamSynchronize(h->st));
  h2d_on(h->st, h->x_rect.p, divs, nr * sizeof(double2));
  launch_pack_spec(h->dt, h->x_rect.p, h->x_spec.p, nlev, Lp, nlev, h->st);
  launch_spec_ucos_vcos(h->dt, h->x_spec.p, Lp, nlev, 0, nlev, 2 * nlev, 3 * nlev, h->st);
  x_set_levs(*h, 4 * nlev, 2 * nlev, 1);
  dev_inverse(*h, h->x_spec.p, Lp, h->x_levs.p, 4 * nlev);
  CK(cudaStreamSynchronize(h->st));
  const size_t n = (size_t)nlev * h->nplane();
  d2h_on(h->st, ug, h->x_grid.p + 2 * n, n * sizeof(double));
  d2h_on(h->st, vg, h->x_grid.p + 3 * n, n * sizeof(double));
  API_END(h)
}

int isca_b200_vor_div_from_uv_grid(IscaHandle h, const double* ug, const double* vg, double* vors, double* divs, int nlev) {
  API_BEGIN(h)
  const Geometry& g = h->g;
  x_prepare(*h, nlev, 2);
  const int Lp = round_up(2 * nlev, 16);
  const size_t n = (size_t)nlev * h->nplane();
  h2d_on(h->st, h->x_grid.p, ug, n * sizeof(double));
  h2d_on(h->st, h->x_grid.p + n, vg, n * sizeof(double));
  launch_divide_by_cos(h->dt, h->x_grid.p, 2 * nlev, h->st);
  CK(cudaMemsetAsync(h->x_trunc.p, 0, Lp, h->st));
  x_set_levs(*h, 2 * nlev, 0, 0);
  dev_forward(*h, h->x_levs.p, 2 * nlev, h->x_spec.p, Lp, h->x_trunc.p);
  DBuf<double2> out; out.alloc((size_t)g.T * Lp);
  launch_spec_vor_div(h->dt, h->x_spec.p, Lp, nlev, 0, nlev, out.p, Lp, 0, nlev, h->st);
  get_spec(*h, out.p, Lp, 0, vors, nlev);
  get_spec(*h, out.p, Lp, nlev, divs, nlev);
  API_END(h)
}

int isca_b200_time_transforms(IscaHandle h, int nlev, int reps, double ms_out[4]) {
  API_BEGIN(h)
  const Geometry& g = h->g;
  x_prepare(*h, nlev, 1);
  const int Lp = round_up(nlev, 16);
  // synthetic spectrum: deterministic, triangular, decaying with total wavenumber
  std::vector<double2> sp((size_t)g.T * Lp, make_double2(0, 0));
  unsigned long long s = 1234;
  for (int p = 0; p < g.T; ++p) {
    int m = g.m_of[h->ht.row_m[p]], n = h->ht.row_n[p];
    if (m + n > g.M) continue;
    double sc = 1.0 / ((1.0 + m + n) * (1.0 + m + n));
    for (int k = 0; k < nlev; ++k) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL; double a = ((s >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
      s = s * 6364136223846793005ULL + 1442695040888963407ULL; double b = ((s >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
      sp[(size_t)p * Lp + k] = make_double2(a * sc, m == 0 ? 0.0 : b * sc);
    }
  }
  h2d_on(h->st, h->x_spec.p, sp.data(), sp.size() * sizeof(double2));
  CK(cudaMemsetAsync(h->x_trunc.p, 1, Lp, h->st));
  x_set_levs(*h, nlev, 0, 0);
  cudaEvent_t ev[5]; for (auto& e : ev) CK(cudaEventCreate(&e));
  double acc[4] = {0, 0, 0, 0};
  for (int r = -2; r < reps; ++r) {
    CK(cudaEventRecord(ev[0], h->st));
    launch_legendre_inv(h->dt, h->x_spec.p, h->four.p, Lp, h->st);
    CK(cudaEventRecord(ev[1], h->st));
    launch_fft_inv(h->dt, h->four_lat(), h->x_levs.p, nlev, Lp, h->st);
    CK(cudaEventRecord(ev[2], h->st));
    launch_fft_fwd(h->dt, h->four_lat(), h->x_levs.p, nlev, Lp, h->st);
    CK(cudaEventRecord(ev[3], h->st));
    launch_legendre_fwd(h->dt, h->four.p, h->x_spec.p, Lp, h->x_trunc.p, h->st);
    CK(cudaEventRecord(ev[4], h->st));
    CK(cudaStreamSynchronize(h->st));
    h->launches += 4;
    if (r >= 0) for (int i = 0; i < 4; ++i) { float ms; CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); acc[i] += ms; }
  }
  for (int i = 0; i < 4; ++i) ms_out[i] = acc[i] / (reps > 0 ? reps : 1);
  for (auto& e : ev) cudaEventDestroy(e);
  API_END(h)
}

int isca_b200_profile_step(IscaHandle h, int n_steps, double* ms_out, int max_groups, char* names, int capacity) {
  if (!h) return -1;
  try {
    std::vector<std::string> order;
    std::map<std::string, double> acc;
    for (int i = 0; i < n_steps; ++i) {
      h->profiling = true; h->marks.clear();
      h->mark("start");
      step_once(*h, 1, nullptr, nullptr, nullptr);
      CK(cudaStreamSynchronize(h->st));
      h->profiling = false;
      for (size_t q = 1; q < h->marks.size(); ++q) {
        float ms = 0; CK(cudaEventElapsedTime(&ms, h->marks[q - 1].second, h->marks[q].second));
        if (!acc.count(h->marks[q].first)) order.push_back(h->marks[q].first);
        acc[h->marks[q].first] += ms;
      }
      for (auto& m : h->marks) cudaEventDestroy(m.second);
      h->marks.clear();
    }
    check_t_flag(*h);
    std::string joined;
    int n = 0;
    for (auto& nm : order) {
      if (n >= max_groups) break;
      ms_out[n++] = acc[nm] / (n_steps > 0 ? n_steps : 1);
      joined += (joined.empty() ? "" : ";") + nm;
    }
    if ((int)joined.size() + 1 > capacity) throw std::runtime_error("names buffer too small");
    std::memcpy(names, joined.c_str(), joined.size() + 1);
    return n;
  } catch (const std::exception& e) { h->profiling = false; h->err = e.what(); return -1; }
}

}  // extern "C"
