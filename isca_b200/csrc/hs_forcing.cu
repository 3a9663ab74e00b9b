// hs_forcing.cu -- hs_forcing_mod with the namelist options beyond the Held-Suarez default (atmos_param/hs_forcing/hs_forcing.F90)
// and the dry model driven by it (atmosphere.F90:276-352): C ABI of include/isca_b200_hs.h.
//
// One thread per column walks the K levels (hs_forcing_column.h): per level the kernel reads p_full, u, v, t (+ um, vm, zfull) and
// updates udt, vdt, tdt, i.e. it is a streaming kernel over 6-9 [K][J][I] arrays with coalesced accesses along longitude; the
// transcendental work per point (log, pow, exp) is small against the memory traffic.  The default Held-Suarez option inside
// isca_b200_step stays fused into the grid kernel of the dynamical core (grid.cu); this module is the general path.
#include "../../include/isca_b200_hs.h"
#include "hs_forcing_column.h"
#include "core_internal.h"
#include "rrtm_internal.h"
#include <cstring>
#include <string>
#include <vector>

using isca_hs::HsParams;

namespace {

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

thread_local std::string g_err;
int hfail(const std::string& m) { g_err = m; return 1; }
#define HCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return hfail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

__global__ void __launch_bounds__(128) hs_forcing_kernel(HsParams p, double dt, double dec, const double* __restrict__ lat,
                                                         const double* __restrict__ lon, const double* __restrict__ coszen,
                                                         const double* __restrict__ p_half, const double* __restrict__ p_full,
                                                         const double* __restrict__ u, const double* __restrict__ v,
                                                         const double* __restrict__ t, const double* __restrict__ um,
                                                         const double* __restrict__ vm, const double* __restrict__ zfull,
                                                         double* __restrict__ udt, double* __restrict__ vdt, double* __restrict__ tdt,
                                                         double* __restrict__ teq, double* __restrict__ tg_prev, double* __restrict__ h_trop) {
  const size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (col >= p.plane) return;
  isca_hs::hs_column(p, col, dt, lat[col], lon[col], coszen ? coszen[col] : 0.0, dec, p_half, p_full, u, v, t, um, vm, zfull, udt, vdt, tdt,
                     teq, tg_prev, h_trop);
}

__global__ void __launch_bounds__(128) hs_tracer_kernel(HsParams p, double dt, const double* __restrict__ p_half,
                                                        const double* __restrict__ rm, double* __restrict__ rdt) {
  const size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (col >= p.plane) return;
  isca_hs::hs_tracer_column(p, col, dt, p_half, rm, rdt);
}

// the spin-up loop of hs_forcing_init (:338-363), one thread per column over the host-computed declinations of the n_iter days;
// tg_prev is left at the value before the last update, as the reference's loop leaves it
__global__ void __launch_bounds__(128) hs_spinup_kernel(HsParams p, int n_iter, const double* __restrict__ dec, const double* __restrict__ lat,
                                                        double* __restrict__ tg_prev) {
  const size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (col >= p.plane) return;
  const double la = lat[col];
  double tg = 250.0, prev = 250.0;
  for (int i = 0; i < n_iter; ++i) {
    prev = tg;
    double t_trop, h_trop, t_surf;
    isca_hs::hs_radiative_surface(p, la, dec[i], t_trop, h_trop, t_surf);
    tg = isca_hs::hs_slab_update(p, 86400.0, t_surf, prev);
  }
  tg_prev[col] = prev;
}

__global__ void lat_lon_kernel(double* lat2d, double* lon2d, const double* rad_lat, int I, int J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < I) {
    lat2d[(size_t)j * I + i] = rad_lat[j];
    lon2d[(size_t)j * I + i] = (i * 360.0 / I) * (3.14159265358979323846 / 180.0);      // rad_lon = deg_lon * pi/180
  }
}

}  // namespace

struct IscaHsForcing_t {
  IscaHsForcingConfig cfg;
  HsParams hp;
  size_t nc = 0, n3 = 0;
  cudaStream_t st = nullptr;
  bool owns_stream = true;
  IscaRrtmDriverConfig astro{};                 // astronomy_nml values for diurnal_exoplanet
  std::vector<double> orb;
  double orbital_rate = 0.0;
  DevBuf tg_prev, coszen, buf[20];
};

namespace {

// update_orbit (:816-828): only the declination is used by the forcing
double hs_declination(const IscaHsForcingConfig& c, long long current_time) {
  const double PI = 3.14159265358979323846;
  const double theta = 2 * PI * (double)current_time / (c.orbital_period * 86400);
  return asin(sin(c.obliq * PI / 180) * sin(theta));
}

int upload(IscaHsForcing h, DevBuf& d, const double* src, size_t n) {
  if (!src) return hfail("hs_forcing: null input array");
  if (!d.ensure(n)) return hfail("cudaMalloc failed");
  HCK(cudaMemcpyAsync(d.p, src, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
  return 0;
}

int spin_up(IscaHsForcing h, const double* lat_dev, long long dt_integer) {
  const IscaHsForcingConfig& c = h->cfg;
  int n_iter = 0;                                              // `do ... spin_count = spin_count + 1 ... if (spin_count >= spinup_time) exit`
  while (true) { ++n_iter; if ((double)n_iter >= c.spinup_time) break; }
  std::vector<double> dec(n_iter);
  for (int i = 0; i < n_iter; ++i) { dt_integer += 86400; dec[i] = hs_declination(c, dt_integer); }
  DevBuf d;
  if (!d.ensure(n_iter) || !h->tg_prev.ensure(h->nc)) return hfail("cudaMalloc failed");
  HCK(cudaMemcpyAsync(d.p, dec.data(), n_iter * sizeof(double), cudaMemcpyHostToDevice, h->st));
  hs_spinup_kernel<<<(unsigned)((h->nc + 127) / 128), 128, 0, h->st>>>(h->hp, n_iter, d.p, lat_dev, h->tg_prev.p);
  HCK(cudaGetLastError());
  HCK(cudaStreamSynchronize(h->st));
  return 0;
}

// hs_forcing on device arrays ([K][J][I]); total_seconds = 86400*days + seconds of Time
int hs_forcing_device(IscaHsForcing h, double dt, long long total_seconds, const double* lon, const double* lat, const double* p_half,
                      const double* p_full, const double* u, const double* v, const double* t, const double* um, const double* vm,
                      const double* zfull, double* udt, double* vdt, double* tdt, const double* rm, double* rdt, int ntr,
                      double* teq, double* h_trop) {
  const IscaHsForcingConfig& c = h->cfg;
  if (c.no_forcing) return 0;
  const unsigned nb = (unsigned)((h->nc + 127) / 128);
  double dec = 0.0;
  const double* cz = nullptr;
  if (c.equilibrium_t_option == ISCA_HS_EXOPLANET) {           // diurnal_exoplanet (astronomy.f90:3672-3715)
    const double twopi = 2.0 * 3.14159265358979323846;
    const double substellar_lon = (h->orbital_rate - c.omega) * (double)total_seconds;
    double gmt = fmod(-substellar_lon, twopi); if (gmt < 0.0) gmt += twopi;                    // modulo()
    double foy = fmod(h->orbital_rate * (double)total_seconds, 1.0); if (foy < 0.0) foy += 1.0;
    if (!h->coszen.ensure(h->nc)) return hfail("cudaMalloc failed");
    if (isca_diurnal_solar_device(h->astro, h->orb, h->st, gmt, foy * twopi, 0.0, (int)h->nc, lat, lon, h->coszen.p))
      return hfail("hs_forcing: zenith-angle kernel launch failed");
    cz = h->coszen.p;
  } else if (c.equilibrium_t_option == ISCA_HS_TOP_DOWN) {
    if (!zfull) return hfail("hs_forcing: top_down needs zfull");
    if (!h->tg_prev.p) return hfail("hs_forcing: top_down has no tg_prev (create with lat, or isca_b200_hs_forcing_set_tg_prev)");
    dec = hs_declination(c, total_seconds);
  }
  hs_forcing_kernel<<<nb, 128, 0, h->st>>>(h->hp, dt, dec, lat, lon, cz, p_half, p_full, u, v, t, um, vm, zfull, udt, vdt, tdt, teq,
                                           h->tg_prev.p, h_trop);
  HCK(cudaGetLastError());
  for (int n = 0; n < ntr; ++n) {
    hs_tracer_kernel<<<nb, 128, 0, h->st>>>(h->hp, dt, p_half, rm + (size_t)n * h->n3, rdt + (size_t)n * h->n3);
    HCK(cudaGetLastError());
  }
  return 0;
}

}  // namespace

struct IscaHsModel_t {
  IscaHandle dyn = nullptr;
  IscaHsForcing hs = nullptr;
  int I = 0, J = 0, K = 0, ntr = 0;
  size_t nc = 0, n3 = 0;
  bool initialized = false;
  long long time_s = 0;                          // Time of atmosphere(Time), seconds
  DevBuf lat2d, lon2d, p_full, p_half, z_full, z_half, dt_u, dt_v, dt_t, dt_q, teq, h_trop;
};

extern "C" {

const char* isca_b200_hs_last_error(void) { return g_err.c_str(); }

int isca_b200_hs_forcing_default_config(IscaHsForcingConfig* c) {
  if (!c) return hfail("null argument");
  memset(c, 0, sizeof *c);
  c->abi_version = 1;
  c->no_forcing = 0; c->do_conserve_energy = 1;
  c->equilibrium_t_option = ISCA_HS_HELD_SUAREZ; c->stratosphere_t_option = ISCA_HS_EXTEND_TP; c->local_heating_option = 0;
  c->num_angles = 3600;
  c->t_zero = 315.; c->t_strat = 200.; c->delh = 60.; c->delv = 10.; c->eps = 0.; c->sigma_b = 0.7; c->P00 = 1.e5; c->p_trop = 1.e4;
  c->alpha = 2. / 7; c->ka = -40.; c->ks = -4.; c->kf = -1.; c->trflux = 1.e-5; c->trsink = -4.;
  c->local_heating_srfamp = 0.0; c->local_heating_xwidth = 10.; c->local_heating_ywidth = 10.; c->local_heating_xcenter = 180.;
  c->local_heating_ycenter = 45.; c->local_heating_vert_decay = 1.e4;
  c->peri_time = 0.25; c->smaxis = 1.5e6; c->albedo = 0.3; c->lapse = 6.5; c->h_a = 2; c->tau_s = 5; c->heat_capacity = 4.2e6;
  c->ml_depth = 1; c->spinup_time = 10800.;
  c->kappa = 2. / 7.; c->rdgas = 287.04; c->grav = 9.80; c->stefan = 5.6734e-8; c->solar_const = 1368.22; c->omega = 7.2921150e-5;
  c->orbital_period = 365.25 * 86400.0; c->orbital_rate = 0.0;
  c->ecc = 0.0; c->obliq = 23.439; c->per = 102.932;
  return 0;
}

int isca_b200_hs_forcing_destroy(IscaHsForcing h) {
  if (!h) return 0;
  if (h->st && h->owns_stream) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

int isca_b200_hs_forcing_create(const IscaHsForcingConfig* cfg, const double* lat, long long days, int seconds, IscaHsForcing* out) {
  if (!cfg || !out) return hfail("null argument");
  if (cfg->abi_version != 1) return hfail("IscaHsForcingConfig abi_version mismatch");
  if (cfg->num_lon < 1 || cfg->num_lat < 1 || cfg->num_levels < 1) return hfail("hs_forcing: bad dimensions");
  if (cfg->equilibrium_t_option < 0 || cfg->equilibrium_t_option > 3)
    return hfail("hs_forcing_nml: not a valid value for equilibrium_t_option ('from_file' reads a netCDF file through interpolator_mod and is not built)");
  if (cfg->stratosphere_t_option < 0 || cfg->stratosphere_t_option > 3) return hfail("hs_forcing_nml: bad stratosphere_t_option");
  if (cfg->local_heating_option < 0 || cfg->local_heating_option > 1)
    return hfail("hs_forcing_nml: not a valid value for local_heating_option ('from_file' is not built)");
  if (cfg->sigma_b == 1.0) return hfail("hs_forcing_nml: sigma_b = 1 divides by zero");
  if (cfg->equilibrium_t_option == ISCA_HS_EXOPLANET && cfg->num_angles < 1) return hfail("astronomy_nml: num_angles must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return hfail("no CUDA device: hs_forcing has no CPU path");
  IscaHsForcing h = new IscaHsForcing_t();
  h->cfg = *cfg;
  const IscaHsForcingConfig& c = h->cfg;
  h->nc = (size_t)c.num_lon * c.num_lat; h->n3 = h->nc * c.num_levels;
  HsParams& p = h->hp;
  const double PI = 3.14159265358979323846, twopi = 2 * PI;
  p.K = c.num_levels; p.plane = h->nc;
  p.do_conserve_energy = c.do_conserve_energy; p.eq_opt = c.equilibrium_t_option; p.strat_opt = c.stratosphere_t_option;
  p.local_heating = c.local_heating_option;
  p.tka = c.ka < 0. ? -1. / (86400 * c.ka) : c.ka;             // :392-406
  p.tks = c.ks < 0. ? -1. / (86400 * c.ks) : c.ks;
  p.vkf = c.kf < 0. ? -1. / (86400 * c.kf) : c.kf;
  p.sigma_b = c.sigma_b; p.t_zero = c.t_zero; p.t_strat = c.t_strat; p.delh = c.delh; p.delv = c.delv; p.eps = c.eps; p.P00 = c.P00;
  p.p_trop = c.p_trop; p.alpha = c.alpha; p.kappa = c.kappa; p.cp_air = c.rdgas / c.kappa;
  p.xwidth = c.local_heating_xwidth * PI / 180.; p.ywidth = c.local_heating_ywidth * PI / 180.;          // :370-381
  p.xcenter = c.local_heating_xcenter * PI / 180.; p.ycenter = c.local_heating_ycenter * PI / 180.;
  p.xcenter = p.xcenter - twopi * floor(p.xcenter / twopi);
  p.srfamp = c.local_heating_srfamp / 86400.0; p.vert_decay = c.local_heating_vert_decay;
  p.lapse = c.lapse; p.h_a = c.h_a; p.tau_s = c.tau_s; p.stefan = c.stefan; p.solar_const = c.solar_const; p.albedo = c.albedo;
  p.ml_heat = c.ml_depth * c.heat_capacity;
  p.trflux = c.trflux;
  { double rd = c.trsink; if (rd < 0.) rd = -86400. * rd; if (rd > 0.) rd = 1. / rd; p.rdamp = rd; }    // tracer_source_sink :697-699
  h->orbital_rate = c.orbital_rate > 0.0 ? c.orbital_rate : twopi / c.orbital_period;
  isca_b200_rrtm_driver_default_config(&h->astro);
  h->astro.ecc = c.ecc; h->astro.obliq = c.obliq; h->astro.per = c.per; h->astro.num_angles = c.num_angles > 0 ? c.num_angles : 1;
  if (c.equilibrium_t_option == ISCA_HS_EXOPLANET) h->orb = isca_rrtm_orbit(h->astro);
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { delete h; return hfail("cudaStreamCreate failed"); }
  if (c.equilibrium_t_option == ISCA_HS_TOP_DOWN && !c.no_forcing && lat) {
    if (!(c.spinup_time >= 0.0 && c.spinup_time < 1.0e7)) { isca_b200_hs_forcing_destroy(h); return hfail("hs_forcing_nml: spinup_time out of range"); }
    if (upload(h, h->buf[0], lat, h->nc) || spin_up(h, h->buf[0].p, 86400 * days + seconds)) { isca_b200_hs_forcing_destroy(h); return 1; }
  }
  *out = h;
  return 0;
}

int isca_b200_hs_forcing(IscaHsForcing h, double dt, long long days, int seconds, const double* lon, const double* lat,
                         const double* p_half, const double* p_full, const double* u, const double* v, const double* t,
                         const double* r, const double* um, const double* vm, const double* tm, const double* rm,
                         double* udt, double* vdt, double* tdt, double* rdt, const double* zfull, int num_tracers,
                         double* teq, double* h_trop) {
  (void)r; (void)tm;                                           // not read by the reference either (shape only)
  if (!h) return hfail("null handle");
  if (num_tracers < 0 || (num_tracers > 0 && (!rm || !rdt))) return hfail("hs_forcing: tracer arrays missing");
  if (h->cfg.no_forcing) return 0;
  const size_t nc = h->nc, n3 = h->n3;
  DevBuf* B = h->buf;
  const bool td = h->cfg.equilibrium_t_option == ISCA_HS_TOP_DOWN;
  if (td && !zfull) return hfail("hs_forcing: top_down needs zfull");
  if (upload(h, B[0], lat, nc) || upload(h, B[1], lon, nc) || upload(h, B[2], p_half, n3 + nc) || upload(h, B[3], p_full, n3) ||
      upload(h, B[4], u, n3) || upload(h, B[5], v, n3) || upload(h, B[6], t, n3) || upload(h, B[7], um, n3) || upload(h, B[8], vm, n3) ||
      upload(h, B[9], udt, n3) || upload(h, B[10], vdt, n3) || upload(h, B[11], tdt, n3)) return 1;
  if (td && upload(h, B[12], zfull, n3)) return 1;
  if (num_tracers > 0 && (upload(h, B[13], rm, n3 * num_tracers) || upload(h, B[14], rdt, n3 * num_tracers))) return 1;
  if (!B[15].ensure(n3) || !B[16].ensure(nc)) return hfail("cudaMalloc failed");
  if (hs_forcing_device(h, dt, 86400 * days + seconds, B[1].p, B[0].p, B[2].p, B[3].p, B[4].p, B[5].p, B[6].p, B[7].p, B[8].p,
                        td ? B[12].p : nullptr, B[9].p, B[10].p, B[11].p, B[13].p, B[14].p, num_tracers, B[15].p, B[16].p)) return 1;
  HCK(cudaMemcpyAsync(udt, B[9].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  HCK(cudaMemcpyAsync(vdt, B[10].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  HCK(cudaMemcpyAsync(tdt, B[11].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (num_tracers > 0) HCK(cudaMemcpyAsync(rdt, B[14].p, n3 * num_tracers * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (teq) HCK(cudaMemcpyAsync(teq, B[15].p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (h_trop && td) HCK(cudaMemcpyAsync(h_trop, B[16].p, nc * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  HCK(cudaStreamSynchronize(h->st));
  return 0;
}

int isca_b200_hs_forcing_get_tg_prev(IscaHsForcing h, double* tg_prev) {
  if (!h || !tg_prev) return hfail("null argument");
  if (!h->tg_prev.p) return hfail("hs_forcing: tg_prev exists only with equilibrium_t_option = 'top_down'");
  HCK(cudaMemcpyAsync(tg_prev, h->tg_prev.p, h->nc * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  HCK(cudaStreamSynchronize(h->st));
  return 0;
}

int isca_b200_hs_forcing_set_tg_prev(IscaHsForcing h, const double* tg_prev) {
  if (!h || !tg_prev) return hfail("null argument");
  if (h->cfg.equilibrium_t_option != ISCA_HS_TOP_DOWN) return hfail("hs_forcing: tg_prev exists only with equilibrium_t_option = 'top_down'");
  if (!h->tg_prev.ensure(h->nc)) return hfail("cudaMalloc failed");
  HCK(cudaMemcpyAsync(h->tg_prev.p, tg_prev, h->nc * sizeof(double), cudaMemcpyHostToDevice, h->st));
  HCK(cudaStreamSynchronize(h->st));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// the dry model with the general forcing
// ------------------------------------------------------------------------------------------------------------------------
int isca_b200_hs_model_destroy(IscaHsModel m) {
  if (!m) return 0;
  if (m->hs) isca_b200_hs_forcing_destroy(m->hs);
  if (m->dyn) isca_b200_destroy(m->dyn);
  delete m;
  return 0;
}

IscaHandle isca_b200_hs_model_dycore(IscaHsModel m) { return m ? m->dyn : nullptr; }

int isca_b200_hs_model_create(const IscaConfig* dyn, const IscaHsForcingConfig* hs, IscaHsModel* out) {
  if (!dyn || !hs || !out) return hfail("null argument");
  if (dyn->num_tracers < 0 || dyn->num_tracers > 1) return hfail("hs_model: num_tracers must be 0 or 1");
  IscaHsModel m = new IscaHsModel_t();
  if (isca_b200_create(dyn, 0, 1, nullptr, &m->dyn)) { std::string e = isca_b200_last_error(nullptr); delete m; return hfail("dynamical core: " + e); }
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) { std::string e = isca_b200_last_error(m->dyn); isca_b200_hs_model_destroy(m); return hfail(e); }
  m->I = v.I; m->J = v.Jloc; m->K = v.K; m->nc = (size_t)v.I * v.Jloc; m->n3 = m->nc * v.K; m->ntr = dyn->num_tracers;
  IscaHsForcingConfig c = *hs;
  c.num_lon = v.I; c.num_lat = v.Jloc; c.num_levels = v.K; c.kappa = dyn->kappa; c.rdgas = dyn->rdgas; c.grav = dyn->grav;
  // the top_down spin-up needs the model latitudes: done in isca_b200_hs_model_init (lat = NULL here)
  if (isca_b200_hs_forcing_create(&c, nullptr, 0, 0, &m->hs)) { isca_b200_hs_model_destroy(m); return 1; }
  cudaStreamDestroy(m->hs->st);
  m->hs->st = v.st; m->hs->owns_stream = false;
  const size_t nc = m->nc, n3 = m->n3;
  bool ok = m->lat2d.ensure(nc) && m->lon2d.ensure(nc) && m->p_full.ensure(n3) && m->p_half.ensure(n3 + nc) && m->z_full.ensure(n3) &&
            m->z_half.ensure(n3 + nc) && m->dt_u.ensure(n3) && m->dt_v.ensure(n3) && m->dt_t.ensure(n3) && m->teq.ensure(n3) && m->h_trop.ensure(nc);
  if (m->ntr) ok = ok && m->dt_q.ensure(n3);
  if (!ok) { isca_b200_hs_model_destroy(m); return hfail("cudaMalloc failed"); }
  *out = m;
  return 0;
}

int isca_b200_hs_model_set_time(IscaHsModel m, long long days, int seconds) {
  if (!m) return hfail("null handle");
  if (days < 0 || seconds < 0) return hfail("hs_model_set_time: negative time");
  m->time_s = 86400 * days + seconds;
  return 0;
}

int isca_b200_hs_model_set_tg_prev(IscaHsModel m, const double* tg_prev) {
  if (!m) return hfail("null handle");
  return isca_b200_hs_forcing_set_tg_prev(m->hs, tg_prev);
}

int isca_b200_hs_model_init(IscaHsModel m) {
  if (!m) return hfail("null handle");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return hfail(isca_b200_last_error(m->dyn));
  dim3 g2((m->I + 127) / 128, m->J);
  lat_lon_kernel<<<g2, 128, 0, v.st>>>(m->lat2d.p, m->lon2d.p, v.rad_lat, m->I, m->J);
  HCK(cudaGetLastError());
  HCK(cudaMemsetAsync(m->teq.p, 0, m->n3 * sizeof(double), v.st));
  HCK(cudaMemsetAsync(m->h_trop.p, 0, m->nc * sizeof(double), v.st));
  HCK(cudaMemsetAsync(m->dt_t.p, 0, m->n3 * sizeof(double), v.st));
  const IscaHsForcingConfig& c = m->hs->cfg;
  if (c.equilibrium_t_option == ISCA_HS_TOP_DOWN && !c.no_forcing && !m->hs->tg_prev.p) {     // no restart handed over: spin up (:338-363)
    if (!(c.spinup_time >= 0.0 && c.spinup_time < 1.0e7)) return hfail("hs_forcing_nml: spinup_time out of range");
    if (spin_up(m->hs, m->lat2d.p, m->time_s)) return 1;
  }
  HCK(cudaStreamSynchronize(v.st));
  m->initialized = true;
  return 0;
}

int isca_b200_hs_model_step(IscaHsModel m, int n_steps) {
  if (!m) return hfail("null handle");
  if (!m->initialized) return hfail("hs_forcing: hs_forcing_init has not been called (isca_b200_hs_model_init)");
  for (int i = 0; i < n_steps; ++i) {
    IscaCoreView v;
    if (isca_core_view(m->dyn, &v)) return hfail(std::string("dynamical core: ") + isca_b200_last_error(m->dyn));
    const int prev = v.previous, cur = v.current;
    const double delta_t = (prev == cur) ? v.dt_atmos : 2 * v.dt_atmos;                        // atmosphere.F90:292-296
    if (isca_core_press_heights(m->dyn, cur, m->p_full.p, m->p_half.p, m->z_full.p, m->z_half.p))
      return hfail(std::string("dynamical core: ") + isca_b200_last_error(m->dyn));
    HCK(cudaMemsetAsync(m->dt_u.p, 0, m->n3 * sizeof(double), v.st));
    HCK(cudaMemsetAsync(m->dt_v.p, 0, m->n3 * sizeof(double), v.st));
    HCK(cudaMemsetAsync(m->dt_t.p, 0, m->n3 * sizeof(double), v.st));
    if (m->ntr) HCK(cudaMemsetAsync(m->dt_q.p, 0, m->n3 * sizeof(double), v.st));
    const long long time_next = m->time_s + (long long)v.dt_atmos;                            // Time_next = Time + Time_step (:298)
    if (hs_forcing_device(m->hs, delta_t, time_next, m->lon2d.p, m->lat2d.p, m->p_half.p, m->p_full.p, v.u[prev], v.v[prev], v.T[prev],
                          v.u[prev], v.v[prev], m->z_full.p, m->dt_u.p, m->dt_v.p, m->dt_t.p, m->ntr ? v.q[prev] : nullptr,
                          m->ntr ? m->dt_q.p : nullptr, m->ntr, m->teq.p, m->h_trop.p)) return 1;
    if (isca_core_step_ext(m->dyn, m->dt_u.p, m->dt_v.p, m->dt_t.p, m->ntr ? m->dt_q.p : nullptr))
      return hfail(std::string("spectral_dynamics: ") + isca_b200_last_error(m->dyn));
    m->time_s = time_next;
  }
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return hfail(isca_b200_last_error(m->dyn));
  HCK(cudaStreamSynchronize(v.st));
  if (isca_core_check(m->dyn)) return hfail(isca_b200_last_error(m->dyn));
  return 0;
}

int isca_b200_hs_model_get(IscaHsModel m, int id, double* host) {
  if (!m || !host) return hfail("null argument");
  IscaCoreView v;
  if (isca_core_view(m->dyn, &v)) return hfail(isca_b200_last_error(m->dyn));
  const double* src = nullptr; size_t n = 0;
  switch (id) {
    case 0: src = m->teq.p; n = m->n3; break;
    case 1: src = m->h_trop.p; n = m->nc; break;
    case 2: if (!m->hs->tg_prev.p) return hfail("hs_model_get: tg_prev exists only with top_down"); src = m->hs->tg_prev.p; n = m->nc; break;
    case 3: src = m->dt_t.p; n = m->n3; break;
    default: return hfail("hs_model_get: unknown id");
  }
  HCK(cudaMemcpyAsync(host, src, n * sizeof(double), cudaMemcpyDeviceToHost, v.st));
  HCK(cudaStreamSynchronize(v.st));
  return 0;
}

}  // extern "C"
