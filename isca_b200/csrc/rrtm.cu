// rrtm.cu -- RRTMG clear-sky longwave and shortwave on the GPU behind include/isca_b200_rrtm.h.
//
// Mapping: one CTA per column.  Phase A: one thread per layer runs inatm + setcoef (+ the Planck interpolation) into
// shared memory; thread 0 forms the column sums (precipitable water -> diffusivity angle, laytrop -> solar source
// layers).  Phase B: one thread per g-point (140 LW / 112 SW) evaluates the gaseous optical depths layer by layer with
// the descriptor-driven generic band code of rrtm_column.h and runs the radiative-transfer sweeps (rtrnmr clear-sky /
// reftra + vrtqdr); the radiances of a level are summed over the g-points with warp shuffles + a fixed-order
// cross-warp sum (deterministic).  Phase C: one thread per level writes fluxes and heating rates.
// The absorption tables (1.9 MB, g-point fastest) stay L2-resident; lanes of one band read consecutive doubles.
// Bound: fp64 issue + L1/L2 gather latency (≈ 6k table reads and ≈ 40k fp64 operations per (column, g-point)), not HBM:
// a radiation call reads ≈ 12 3-D fields.  It runs every dt_rad (48 steps in the MiMA configuration).
#include "../../include/isca_b200_rrtm.h"
#include "rrtm_tables.h"
#include "rrtm_internal.h"
#include "rrtm_kernels.h"
#include <cuda_runtime.h>
#include <cstdlib>
#include <string>

using namespace rrtm;
using namespace rrtm_k;

namespace {

std::string& rr_thread_error() { static thread_local std::string e; return e; }

struct DevBuf {
  double* p = nullptr; size_t n = 0;
  bool ensure(size_t count) {
    if (count <= n) return true;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    if (cudaMalloc(&p, count * sizeof(double)) != cudaSuccess) return false;
    n = count; return true;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

struct IscaRrtm_t {
  IscaRrtmConfig cfg;
  double* d_arena = nullptr;
  LwBand* d_lw = nullptr;
  SwBand* d_sw = nullptr;
  Tab tab;
  cudaStream_t st = nullptr;
  bool owns_stream = true;
  std::string err;
  DevBuf buf[32];
  DevBuf lays;                       // longwave setcoef planes (scratch of isca_rrtm_lw_device)
  ColIn last_lw{}, last_sw{};
  bool have_lw = false, have_sw = false;
};

namespace {
int rfail(IscaRrtm r, const std::string& m) { if (r) r->err = m; rr_thread_error() = m; return 1; }
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return rfail(r, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

int up(IscaRrtm r, DevBuf& d, const double* h, size_t n, const double** out) {
  if (!h) { *out = nullptr; return 0; }
  if (!d.ensure(n)) return rfail(r, "cudaMalloc failed");
  RCK(cudaMemcpyAsync(d.p, h, n * sizeof(double), cudaMemcpyHostToDevice, r->st));
  *out = d.p;
  return 0;
}
double heatfac_of(double cp_air) { return GRAV * SECDY / (cp_air * 1.0e2); }

// kernel launches on device pointers
int isca_rrtm_lw_device(IscaRrtm r, const ColIn& in_) {
  ColIn in = in_;
  if (in.nlay > KMAX || in.nlay < 2) return rfail(r, "rrtmg_lw: num_levels must be 2..64");
  static const bool gpoint = std::getenv("ISCA_B200_RRTM_LW_GPOINT") != nullptr;     // development: the g-point-per-thread kernel
  if (gpoint) {
    const size_t smem = lw_smem_doubles(in.nlay) * sizeof(double);
    RCK(cudaFuncSetAttribute(rrtmg_lw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lw_smem_doubles(KMAX) * sizeof(double))));
    rrtmg_lw_kernel<<<in.ncol, LW_THREADS, smem, r->st>>>(r->d_arena, r->tab, r->d_lw, in);
  } else {
#if ISCA_LWC_PREPASS
    const size_t ps = (size_t)in.ncol * in.nlay;
    if (!r->lays.ensure(ps * LAYP_N)) return rfail(r, "cudaMalloc failed (longwave setcoef planes)");
    in.lays = r->lays.p;
    rrtmg_lw_setcoef_kernel<<<(unsigned)((ps + 127) / 128), 128, 0, r->st>>>(r->d_arena, r->tab, in);
#endif
    rrtmg_lw_col_kernel<<<(in.ncol + 31) / 32, 32 * LWC_WARPS, 0, r->st>>>(r->d_arena, r->tab, r->d_lw, in);
  }
  RCK(cudaGetLastError());
  r->last_lw = in; r->have_lw = true;
  return 0;
}
int isca_rrtm_sw_device(IscaRrtm r, const ColIn& in) {
  if (in.nlay > KMAX || in.nlay < 2) return rfail(r, "rrtmg_sw: num_levels must be 2..64");
  static const bool percol = std::getenv("ISCA_B200_RRTM_SW_COL") != nullptr;     // development: the column-per-lane kernel (measured slower)
  if (!percol) {
    const size_t smem = sw_smem_doubles(in.nlay) * sizeof(double);
    RCK(cudaFuncSetAttribute(rrtmg_sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sw_smem_doubles(KMAX) * sizeof(double))));
    rrtmg_sw_kernel<<<in.ncol, SW_THREADS, smem, r->st>>>(r->d_arena, r->tab, r->d_sw, in);
  } else {
    rrtmg_sw_col_kernel<<<(in.ncol + 31) / 32, 32 * SWC_WARPS, 0, r->st>>>(r->d_arena, r->tab, r->d_sw, in);
  }
  RCK(cudaGetLastError());
  r->last_sw = in; r->have_sw = true;
  return 0;
}

}  // namespace

// run_rrtmg on device pointers: used by the moist-model driver (moist_model.cu) -- same stream, nothing crosses PCIe

// (model layout; scratch in r->buf[16..])
int isca_rrtm_run_device(IscaRrtm r, cudaStream_t st, const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                         const double* t, const double* q, const double* o3, const double* t_surf, const double* albedo, const double* coszen,
                         double* tdt, double* tdt_rad, double* flux_sw, double* flux_lw, double* olr, double* toa_sw) {
  const IscaRrtmConfig& c = r->cfg;
  const int K = c.num_levels, ls = c.lonstep; const size_t nm = (size_t)c.num_lon * c.num_lat, nc = nm / ls;   // model / rrtm columns
  cudaStream_t keep = r->st; r->st = st;
  DevBuf* B = r->buf + 16;
  const size_t n3 = nc * K, n3h = nc * (K + 1);
  const size_t sizes[15] = {n3, n3h, n3, n3h, n3, n3, n3h, n3h, n3, n3h, n3h, n3, nc, nc, nc};
  for (int i = 0; i < 15; ++i) if (!B[i].ensure(sizes[i])) { r->st = keep; return rfail(r, "cudaMalloc failed"); }
  double *play = B[0].p, *plev = B[1].p, *tlay = B[2].p, *tlev = B[3].p, *h2o = B[4].p, *o3v = B[5].p;
  double *swu = B[6].p, *swd = B[7].p, *swhr = B[8].p, *lwu = B[9].p, *lwd = B[10].p, *lwhr = B[11].p;
  PrepArgs pa{(int)nc, K, p_full, p_half, z_full, z_half, t, q, o3, play, plev, tlay, tlev, h2o, o3v,
              (1000.0 * c.gas_constant / c.rdgas) / c.wtmh2o,
              c.input_o3_file_is_mmr ? (1000.0 * c.gas_constant / c.rdgas) / c.wtmozone : 1.0,
              c.h2o_lower_limit, c.temp_lower_limit, c.temp_upper_limit, c.convert_sphum_to_vmr,
              ls, c.num_lon, nm, t_surf, albedo, coszen, B[12].p, B[13].p, B[14].p};
  const int T = 128, G = (int)((nc + T - 1) / T), Gm = (int)((nm + T - 1) / T);
  rrtm_prepare_kernel<<<G, T, 0, st>>>(pa);
  rrtm_fix_top_kernel<<<G, T, 0, st>>>((int)nc, K, play, plev);
  ColIn in{};
  in.ncol = (int)nc; in.nlay = K; in.play = play; in.plev = plev; in.tlay = tlay; in.tlev = tlev; in.emis = nullptr;
  in.tsfc = ls > 1 ? B[12].p : t_surf; in.albedo = ls > 1 ? B[13].p : albedo; in.coszen = ls > 1 ? B[14].p : coszen;
  for (int i = 0; i < NSP; ++i) { in.gas[i] = nullptr; in.gas_c[i] = 0.0; }
  for (int i = 0; i < 4; ++i) { in.xs[i] = nullptr; in.xs_c[i] = 0.0; }
  in.gas[SP_H2O] = h2o; in.gas[SP_O3] = o3v; in.gas_c[SP_CO2] = c.co2ppmv * 1.0e-6;
  if (c.include_secondary_gases) { in.gas_c[SP_CH4] = c.ch4_val; in.gas_c[SP_N2O] = c.n2o_val; in.gas_c[SP_O2] = c.o2_val; }
  in.heatfac = heatfac_of(c.cp_air);
  ColIn sw = in; sw.uflx = swu; sw.dflx = swd; sw.hr = swhr; sw.adjflux = c.solrad * (c.solr_cnst / 1.36822e+03);
  int rc = isca_rrtm_sw_device(r, sw);
  if (!rc) {
    ColIn lw = in; lw.uflx = lwu; lw.dflx = lwd; lw.hr = lwhr;
    if (c.include_secondary_gases) { lw.xs_c[0] = c.ccl4_val; lw.xs_c[1] = c.cfc11_val; lw.xs_c[2] = c.cfc12_val; lw.xs_c[3] = c.cfc22_val; }
    rc = isca_rrtm_lw_device(r, lw);
  }
  if (!rc) {
    FinishArgs fa{(int)nm, K, swhr, lwhr, swu, swd, lwu, lwd, tdt, tdt_rad, flux_sw, flux_lw, olr, toa_sw, ls, c.num_lon};
    rrtm_finish_kernel<<<Gm, T, 0, st>>>(fa);
    if (cudaGetLastError() != cudaSuccess) rc = rfail(r, "rrtm_finish_kernel launch failed");
  }
  r->st = keep;
  return rc;
}

void isca_rrtm_set_stream(IscaRrtm r, cudaStream_t st) {
  if (r->owns_stream && r->st) cudaStreamDestroy(r->st);
  r->st = st; r->owns_stream = false;
}

// ---- astronomy_mod on the host: orbit table, angle, declination (astronomy.f90) ----
namespace {
double r_inv_squared(const IscaRrtmDriverConfig& dc, double ang) {
  const double deg_to_rad = 3.14159265358979323846 / 180.0;
  double r = (1.0 - dc.ecc * dc.ecc) / (1.0 + dc.ecc * cos(ang - dc.per * deg_to_rad));
  return 1.0 / (r * r);
}
}  // namespace

std::vector<double> isca_rrtm_orbit(const IscaRrtmDriverConfig& dc) {
  const double twopi = 2.0 * 3.14159265358979323846;
  std::vector<double> orb(dc.num_angles + 1, 0.0);
  double dt = twopi / (double)dc.num_angles;
  dt = dt * sqrt(1.0 - dc.ecc * dc.ecc);
  for (int n = 1; n <= dc.num_angles; ++n) {
    double d1 = dt * r_inv_squared(dc, orb[n - 1]);
    double d2 = dt * r_inv_squared(dc, orb[n - 1] + 0.5 * d1);
    double d3 = dt * r_inv_squared(dc, orb[n - 1] + 0.5 * d2);
    double d4 = dt * r_inv_squared(dc, orb[n - 1] + d3);
    orb[n] = orb[n - 1] + (d1 / 6.0 + d2 / 3.0 + d3 / 3.0 + d4 / 6.0);
  }
  return orb;
}

namespace {
int launch_coszen(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb, cudaStream_t st, double gmt, double time_since_ae, double dt,
                  int n, const double* lat, const double* lon, double* coszen, double* fracday, double* rrsun) {
  const double twopi = 2.0 * 3.14159265358979323846, deg_to_rad = 3.14159265358979323846 / 180.0;
  // angle(time_since_ae): linear interpolation in the orbit table
  double norm_time = time_since_ae * (double)dc.num_angles / twopi;
  long fl = (long)floor(norm_time);
  int i0 = (int)(((fl % dc.num_angles) + dc.num_angles) % dc.num_angles);
  double x = norm_time - floor(norm_time);
  double ang = (1.0 - x) * orb[i0] + x * orb[i0 + 1];
  ang = fmod(ang, twopi); if (ang < 0.0) ang += twopi;
  double dec = asin(-sin(dc.obliq * deg_to_rad) * sin(ang));
  if (rrsun) *rrsun = r_inv_squared(dc, ang);
  coszen_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, lat, lon, gmt, dec, dt, dc.frierson_solar_rad, dc.del_sol, dc.del_sw, coszen, fracday);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
}  // namespace

int isca_rrtm_coszen_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb, cudaStream_t st, double total_seconds, int n,
                            const double* lat, const double* lon, double* coszen, double* fracday) {
  const double twopi = 2.0 * 3.14159265358979323846;
  double ts = total_seconds;
  if (dc.solday > 0) ts = fmod(total_seconds, 86400.0) + (double)dc.solday * 86400.0;     // Time_loc = set_time(seconds, solday)
  double frac_of_day = ts / dc.day_in_s;
  double frac_of_year = dc.solday > 0 ? ((double)dc.solday * dc.day_in_s) / dc.year_in_s : ts / dc.year_in_s;
  double gmt = fabs(fmod(frac_of_day, 1.0)) * twopi;
  double y = fmod(frac_of_year - dc.equinox_day, 1.0); if (y < 0.0) y += 1.0;               // modulo()
  double time_since_ae = y * twopi;
  int dt_rad_avg = dc.dt_rad_avg > 0 ? dc.dt_rad_avg : dc.dt_rad;
  double dt = dc.do_rad_time_avg ? ((double)dt_rad_avg / dc.day_in_s) * twopi : 0.0;
  return launch_coszen(dc, orb, st, gmt, time_since_ae, dt, n, lat, lon, coszen, fracday, nullptr);
}

int isca_diurnal_solar_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb, cudaStream_t st, double gmt, double time_since_ae,
                              double dt, int n, const double* lat, const double* lon, double* coszen) {
  IscaRrtmDriverConfig c = dc;
  c.frierson_solar_rad = 0;
  return launch_coszen(c, orb, st, gmt, time_since_ae, dt, n, lat, lon, coszen, nullptr, nullptr);
}

int isca_gray_coszen_device(const IscaRrtmDriverConfig& dc, const std::vector<double>& orb, cudaStream_t st, double days, double seconds,
                            int n, const double* lat, const double* lon, double* coszen) {
  const double twopi = 2.0 * 3.14159265358979323846;
  const double frac_of_day = seconds / dc.day_in_s;
  const double frac_of_year = dc.solday >= 0 ? ((double)dc.solday * dc.day_in_s) / dc.year_in_s      // `if(solday .ge. 0)` (:425)
                                             : (seconds + days * dc.day_in_s) / dc.year_in_s;
  const double gmt = fabs(fmod(frac_of_day, 1.0)) * twopi;
  double y = fmod(frac_of_year - dc.equinox_day, 1.0); if (y < 0.0) y += 1.0;                        // modulo()
  const double time_since_ae = y * twopi;
  const double dt = dc.do_rad_time_avg ? ((double)dc.dt_rad_avg / dc.day_in_s) * twopi : 0.0;
  IscaRrtmDriverConfig c = dc;
  c.frierson_solar_rad = 0;
  return launch_coszen(c, orb, st, gmt, time_since_ae, dt, n, lat, lon, coszen, nullptr, nullptr);
}

extern "C" {

int isca_b200_rrtm_default_config(IscaRrtmConfig* c) {
  if (!c) return 1;
  memset(c, 0, sizeof *c);
  c->abi_version = 1;
  c->num_lon = 128; c->num_lat = 64; c->num_levels = 40;
  c->rdgas = 287.04; c->cp_air = 287.04 / (2.0 / 7.0); c->gas_constant = 8.314;
  c->wtmh2o = 2.896440E+01 * (287.04 / 461.50); c->wtmozone = 47.99820;
  c->co2ppmv = 300.0; c->h2o_lower_limit = 2.0e-7; c->temp_lower_limit = 100.0; c->temp_upper_limit = 370.0;
  c->solrad = 1.0; c->solr_cnst = 1368.22;
  c->include_secondary_gases = 0;
  c->convert_sphum_to_vmr = 1; c->input_o3_file_is_mmr = 1; c->lonstep = 1;
  return 0;
}

const char* isca_b200_rrtm_last_error(IscaRrtm r) { return r ? r->err.c_str() : rr_thread_error().c_str(); }

int isca_b200_rrtm_create(const IscaRrtmConfig* cfg, const char* table_path, IscaRrtm* out) {
  IscaRrtm r = nullptr;
  if (!cfg || !out || !table_path) return rfail(nullptr, "rrtm_create: null argument");
  if (cfg->abi_version != 1) return rfail(nullptr, "rrtm_create: abi_version mismatch");
  if (cfg->lonstep < 1 || (cfg->lonstep > 1 && cfg->num_lon % cfg->lonstep != 0))
    return rfail(nullptr, "rrtm_create: lonstep must be >= 1 and divide the number of longitudes");
  if (cfg->num_levels < 2 || cfg->num_levels > KMAX) return rfail(nullptr, "rrtm_create: num_levels must be 2..64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return rfail(nullptr, "rrtm_create: no CUDA device (there is no CPU fallback)");
  HostTables ht;
  if (!ht.build(table_path)) return rfail(nullptr, "rrtm_create: " + ht.err);
  r = new IscaRrtm_t();
  r->cfg = *cfg; r->tab = ht.tab;
  if (cudaStreamCreateWithFlags(&r->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&r->d_arena, ht.arena.size() * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&r->d_lw, sizeof ht.lw) != cudaSuccess || cudaMalloc(&r->d_sw, sizeof ht.sw) != cudaSuccess ||
      cudaMemcpy(r->d_arena, ht.arena.data(), ht.arena.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(r->d_lw, ht.lw, sizeof ht.lw, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(r->d_sw, ht.sw, sizeof ht.sw, cudaMemcpyHostToDevice) != cudaSuccess) {
    std::string m = std::string("rrtm_create: ") + cudaGetErrorString(cudaGetLastError());
    isca_b200_rrtm_destroy(r);
    return rfail(nullptr, m);
  }
  *out = r;
  return 0;
}

int isca_b200_rrtm_destroy(IscaRrtm r) {
  if (!r) return 0;
  if (r->st && r->owns_stream) { cudaStreamSynchronize(r->st); cudaStreamDestroy(r->st); }
  if (r->d_arena) cudaFree(r->d_arena);
  if (r->d_lw) cudaFree(r->d_lw);
  if (r->d_sw) cudaFree(r->d_sw);
  delete r;
  return 0;
}

int isca_b200_rrtmg_lw(IscaRrtm r, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                       const double* tlev, const double* tsfc, const double* h2ovmr, const double* o3vmr,
                       const double* co2vmr, const double* ch4vmr, const double* n2ovmr, const double* o2vmr,
                       const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr, const double* ccl4vmr,
                       const double* emis, double* uflx, double* dflx, double* hr) {
  if (!r) return rfail(nullptr, "rrtmg_lw: null handle");
  if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !uflx || !dflx || !hr) return rfail(r, "rrtmg_lw: null array");
  if (ncol < 1 || nlay < 2 || nlay > KMAX) return rfail(r, "rrtmg_lw: bad ncol / nlay");
  const size_t n3 = (size_t)ncol * nlay, n3h = (size_t)ncol * (nlay + 1);
  ColIn in{};
  in.ncol = ncol; in.nlay = nlay;
  DevBuf* B = r->buf;
  if (up(r, B[0], play, n3, &in.play) || up(r, B[1], plev, n3h, &in.plev) || up(r, B[2], tlay, n3, &in.tlay) || up(r, B[3], tlev, n3h, &in.tlev) ||
      up(r, B[4], tsfc, ncol, &in.tsfc) || up(r, B[5], emis, (size_t)ncol * NB_LW, &in.emis)) return 1;
  const double* g[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  for (int i = 0; i < NSP; ++i) { in.gas_c[i] = 0.0; if (up(r, B[6 + i], g[i], n3, &in.gas[i])) return 1; }
  const double* x[4] = {ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr};
  for (int i = 0; i < 4; ++i) { in.xs_c[i] = 0.0; if (up(r, B[13 + i], x[i], n3, &in.xs[i])) return 1; }
  if (!r->buf[17].ensure(n3h) || !r->buf[18].ensure(n3h) || !r->buf[19].ensure(n3)) return rfail(r, "cudaMalloc failed");
  in.uflx = r->buf[17].p; in.dflx = r->buf[18].p; in.hr = r->buf[19].p;
  in.heatfac = heatfac_of(r->cfg.cp_air);
  if (isca_rrtm_lw_device(r, in)) return 1;
  RCK(cudaMemcpyAsync(uflx, in.uflx, n3h * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(dflx, in.dflx, n3h * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(hr, in.hr, n3 * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaStreamSynchronize(r->st));
  return 0;
}

int isca_b200_rrtmg_sw(IscaRrtm r, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                       const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                       const double* n2ovmr, const double* o2vmr, const double* albedo, const double* coszen,
                       double adjes, double scon, double* swuflx, double* swdflx, double* swhr) {
  if (!r) return rfail(nullptr, "rrtmg_sw: null handle");
  if (!play || !plev || !tlay || !h2ovmr || !albedo || !coszen || !swuflx || !swdflx || !swhr) return rfail(r, "rrtmg_sw: null array");
  if (ncol < 1 || nlay < 2 || nlay > KMAX) return rfail(r, "rrtmg_sw: bad ncol / nlay");
  const size_t n3 = (size_t)ncol * nlay, n3h = (size_t)ncol * (nlay + 1);
  ColIn in{};
  in.ncol = ncol; in.nlay = nlay;
  DevBuf* B = r->buf;
  if (up(r, B[0], play, n3, &in.play) || up(r, B[1], plev, n3h, &in.plev) || up(r, B[2], tlay, n3, &in.tlay) ||
      up(r, B[4], albedo, ncol, &in.albedo) || up(r, B[5], coszen, ncol, &in.coszen)) return 1;
  const double* g[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  for (int i = 0; i < NSP; ++i) { in.gas_c[i] = 0.0; if (up(r, B[6 + i], g[i], n3, &in.gas[i])) return 1; }
  if (!r->buf[17].ensure(n3h) || !r->buf[18].ensure(n3h) || !r->buf[19].ensure(n3)) return rfail(r, "cudaMalloc failed");
  in.uflx = r->buf[17].p; in.dflx = r->buf[18].p; in.hr = r->buf[19].p;
  in.heatfac = heatfac_of(r->cfg.cp_air);
  in.adjflux = adjes * (scon / 1.36822e+03);       // adjflux = adjflx * scon / rrsw_scon (inatm_sw; dyofyr = 0)
  if (isca_rrtm_sw_device(r, in)) return 1;
  RCK(cudaMemcpyAsync(swuflx, in.uflx, n3h * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(swdflx, in.dflx, n3h * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(swhr, in.hr, n3 * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaStreamSynchronize(r->st));
  return 0;
}

int isca_b200_run_rrtmg(IscaRrtm r, const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                        const double* t, const double* q, const double* o3, const double* t_surf, const double* albedo,
                        const double* coszen, double* tdt, double* tdt_rad, double* flux_sw, double* flux_lw, double* olr,
                        double* toa_sw) {
  if (!r) return rfail(nullptr, "run_rrtmg: null handle");
  if (!p_full || !p_half || !z_full || !z_half || !t || !q || !t_surf || !albedo || !coszen || !tdt || !flux_sw || !flux_lw)
    return rfail(r, "run_rrtmg: null array");
  const int K = r->cfg.num_levels; const size_t nc = (size_t)r->cfg.num_lon * r->cfg.num_lat;
  const size_t n3 = nc * K, n3h = nc * (K + 1);
  DevBuf* B = r->buf;
  const double *d_pf, *d_ph, *d_zf, *d_zh, *d_t, *d_q, *d_o3, *d_ts, *d_alb, *d_cz, *d_tdt;
  if (up(r, B[0], p_full, n3, &d_pf) || up(r, B[1], p_half, n3h, &d_ph) || up(r, B[2], z_full, n3, &d_zf) || up(r, B[3], z_half, n3h, &d_zh) ||
      up(r, B[4], t, n3, &d_t) || up(r, B[5], q, n3, &d_q) || up(r, B[6], o3, n3, &d_o3) || up(r, B[7], t_surf, nc, &d_ts) ||
      up(r, B[8], albedo, nc, &d_alb) || up(r, B[9], coszen, nc, &d_cz) || up(r, B[10], tdt, n3, &d_tdt)) return 1;
  if (!B[11].ensure(n3) || !B[12].ensure(nc) || !B[13].ensure(nc) || !B[14].ensure(nc) || !B[15].ensure(nc)) return rfail(r, "cudaMalloc failed");
  if (isca_rrtm_run_device(r, r->st, d_pf, d_ph, d_zf, d_zh, d_t, d_q, d_o3, d_ts, d_alb, d_cz, (double*)d_tdt, B[11].p, B[12].p, B[13].p,
                           B[14].p, B[15].p)) return 1;
  RCK(cudaMemcpyAsync(tdt, d_tdt, n3 * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  if (tdt_rad) RCK(cudaMemcpyAsync(tdt_rad, B[11].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(flux_sw, B[12].p, nc * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(flux_lw, B[13].p, nc * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  if (olr) RCK(cudaMemcpyAsync(olr, B[14].p, nc * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  if (toa_sw) RCK(cudaMemcpyAsync(toa_sw, B[15].p, nc * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaStreamSynchronize(r->st));
  return 0;
}

int isca_b200_rrtm_time(IscaRrtm r, int which, int reps, double* ms) {
  if (!r || !ms || reps < 1) return rfail(r, "rrtm_time: bad argument");
  if ((which == 0 && !r->have_lw) || (which == 1 && !r->have_sw)) return rfail(r, "rrtm_time: no previous call of that kernel");
  struct Ev {                                        // destroyed on every exit path
    cudaEvent_t e = nullptr;
    ~Ev() { if (e) cudaEventDestroy(e); }
  } e0, e1;
  RCK(cudaEventCreate(&e0.e)); RCK(cudaEventCreate(&e1.e));
  for (int i = 0; i < 2; ++i) { if (which == 0 ? isca_rrtm_lw_device(r, r->last_lw) : isca_rrtm_sw_device(r, r->last_sw)) return 1; }
  RCK(cudaEventRecord(e0.e, r->st));
  for (int i = 0; i < reps; ++i) { if (which == 0 ? isca_rrtm_lw_device(r, r->last_lw) : isca_rrtm_sw_device(r, r->last_sw)) return 1; }
  RCK(cudaEventRecord(e1.e, r->st));
  RCK(cudaEventSynchronize(e1.e));
  float t = 0.f;
  RCK(cudaEventElapsedTime(&t, e0.e, e1.e));
  *ms = (double)t / reps;
  return 0;
}

int isca_b200_rrtm_driver_default_config(IscaRrtmDriverConfig* dc) {
  if (!dc) return 1;
  memset(dc, 0, sizeof *dc);
  dc->abi_version = 1;
  dc->dt_rad = 0; dc->dt_rad_avg = -1; dc->do_rad_time_avg = 1; dc->store_intermediate_rad = 1; dc->solday = 0; dc->frierson_solar_rad = 0;
  dc->equinox_day = 0.75; dc->del_sol = 0.95; dc->del_sw = 0.0;
  dc->ecc = 0.0; dc->obliq = 23.439; dc->per = 102.932; dc->num_angles = 3600;
  dc->day_in_s = 86400.0; dc->year_in_s = 360.0 * 86400.0;
  return 0;
}

int isca_b200_diurnal_solar(IscaRrtm r, const IscaRrtmDriverConfig* dc, int n, const double* lat, const double* lon, double gmt,
                            double time_since_ae, double dt, double* cosz, double* fracday, double* rrsun) {
  if (!r) return rfail(nullptr, "diurnal_solar: null handle");
  if (!dc || !lat || !lon || !cosz || !fracday || n < 1) return rfail(r, "diurnal_solar: null argument");
  const double twopi = 2.0 * 3.14159265358979323846;
  if (time_since_ae < 0.0 || time_since_ae > twopi) return rfail(r, "astronomy_mod: time_since_ae not between 0 and 2pi");
  if (gmt < 0.0 || gmt > twopi) return rfail(r, "astronomy_mod: gmt not between 0 and 2pi");
  const double *d_lat, *d_lon;
  if (up(r, r->buf[0], lat, n, &d_lat) || up(r, r->buf[1], lon, n, &d_lon)) return 1;
  if (!r->buf[2].ensure(n) || !r->buf[3].ensure(n)) return rfail(r, "cudaMalloc failed");
  IscaRrtmDriverConfig c = *dc; c.frierson_solar_rad = 0;
  if (launch_coszen(c, isca_rrtm_orbit(c), r->st, gmt, time_since_ae, dt, n, d_lat, d_lon, r->buf[2].p, r->buf[3].p, rrsun))
    return rfail(r, "coszen_kernel launch failed");
  RCK(cudaMemcpyAsync(cosz, r->buf[2].p, n * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaMemcpyAsync(fracday, r->buf[3].p, n * sizeof(double), cudaMemcpyDeviceToHost, r->st));
  RCK(cudaStreamSynchronize(r->st));
  return 0;
}

}  // extern "C"
