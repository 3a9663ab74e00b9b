// TEST INFRASTRUCTURE ONLY -- CPU thread emulation of the RRTMG CUDA kernels (isca_b200/csrc/rrtm_kernels.h).
//
// The kernel bodies are compiled unchanged by g++ behind a small shim: one OS thread per CUDA thread of a block,
// `__syncthreads()` = std::barrier over the block, `__shfl_xor_sync` = a per-warp exchange buffer with a warp barrier,
// `__shared__` = a function-local static (one block runs at a time), threadIdx / blockIdx / blockDim thread-local.
// "Device" pointers are host pointers.  This checks what the serial host build (rrtm_host.cpp) cannot: the phase structure of
// the kernels, the shared-memory staging and indexing, the warp-shuffle + cross-warp reductions with padding lanes, the early
// exit of night columns, and the glue kernels of run_rrtmg (layout, interp_temp, lonstep) and the zenith-angle kernel.
// Built by tests/test_rrtm_host.py into tests/host/_build/; never linked into the product library.
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

// ---- CUDA shim ------------------------------------------------------------------------------------------------
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
static std::barrier<>* g_block_barrier = nullptr;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
static double g_shfl[64][32];

#define __global__
#define __device__
#define __constant__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_shfl[w][l] = v;
  g_warp_barrier[w]->arrive_and_wait();
  double r = g_shfl[w][l ^ o];
  g_warp_barrier[w]->arrive_and_wait();
  return r;
}

static inline int atomicExch(int* a, int v) { int o = __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); return o; }

#define ISCA_RRTM_EMU 1
#include "../../isca_b200/csrc/rrtm_tables.h"
#include "../../isca_b200/csrc/rrtm_kernels.h"
#include "../../isca_b200/csrc/physics_dry_kernels.h"

using namespace rrtm;
using namespace rrtm_k;

namespace {
// run `kernel(args...)` for a grid of `nblocks` blocks of `nthreads` threads (nthreads a multiple of 32 when shuffles are used)
template <class F>
void launch(int nblocks, int nthreads, F body) {
  blockDim.x = nthreads; gridDim.x = nblocks;
  std::barrier<> bar(nthreads);
  g_block_barrier = &bar;
  g_warp_barrier.clear();
  for (int w = 0; w < (nthreads + 31) / 32; ++w) g_warp_barrier.emplace_back(new std::barrier<>(32));
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=, &bar]() {
      threadIdx.x = t;
      for (int b = 0; b < nblocks; ++b) {
        blockIdx.x = b;
        body();
        bar.arrive_and_wait();            // a block finishes before the next one reuses the "shared memory"
      }
    });
  for (auto& x : th) x.join();
}

std::unique_ptr<HostTables> g_tab;
std::string g_err;
bool ensure_tables(const char* path) {
  if (g_tab) return true;
  g_tab.reset(new HostTables());
  if (!g_tab->build(path)) { g_err = g_tab->err; g_tab.reset(); return false; }
  return true;
}
void fill_common(ColIn& in, int ncol, int nlay, const double* play, const double* plev, const double* tlay, const double* const* gas,
                 double cp_air) {
  std::memset(&in, 0, sizeof in);
  in.ncol = ncol; in.nlay = nlay; in.play = play; in.plev = plev; in.tlay = tlay;
  for (int i = 0; i < NSP; ++i) { in.gas[i] = gas[i]; in.gas_c[i] = 0.0; }
  in.heatfac = GRAV * SECDY / (cp_air * 1.0e2);
}
}  // namespace

extern "C" {

const char* rrtm_emu_error() { return g_err.c_str(); }

int rrtm_emu_lw(const char* table_path, double cp_air, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                const double* tlev, const double* tsfc, const double* h2ovmr, const double* o3vmr, const double* co2vmr,
                const double* ch4vmr, const double* n2ovmr, const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                const double* cfc22vmr, const double* ccl4vmr, const double* emis, double* uflx, double* dflx, double* hr) {
  if (!ensure_tables(table_path)) return 1;
  HostTables& H = *g_tab;
  const double* gas[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  ColIn in;
  fill_common(in, ncol, nlay, play, plev, tlay, gas, cp_air);
  in.tlev = tlev; in.tsfc = tsfc; in.emis = emis;
  const double* xs[4] = {ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr};
  for (int i = 0; i < 4; ++i) in.xs[i] = xs[i];
  in.uflx = uflx; in.dflx = dflx; in.hr = hr;
  const double* A = H.arena.data();
  const LwBand* bands = H.lw;
  Tab tb = H.tab;
  if (std::getenv("RRTM_EMU_LW_GPOINT")) launch(ncol, LW_THREADS, [=]() { rrtmg_lw_kernel(A, tb, bands, in); });
  else {
    std::vector<double> lays((size_t)LAYP_N * ncol * nlay);
    in.lays = lays.data();
    launch((ncol * nlay + 127) / 128, 128, [=]() { rrtmg_lw_setcoef_kernel(A, tb, in); });
    launch((ncol + 31) / 32, 32 * LWC_WARPS, [=]() { rrtmg_lw_col_kernel(A, tb, bands, in); });
  }
  return 0;
}

int rrtm_emu_sw(const char* table_path, double cp_air, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                const double* o2vmr, const double* albedo, const double* coszen, double adjes, double scon, double* swuflx,
                double* swdflx, double* swhr) {
  if (!ensure_tables(table_path)) return 1;
  HostTables& H = *g_tab;
  const double* gas[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  ColIn in;
  fill_common(in, ncol, nlay, play, plev, tlay, gas, cp_air);
  in.albedo = albedo; in.coszen = coszen; in.adjflux = adjes * (scon / 1.36822e+03);
  in.uflx = swuflx; in.dflx = swdflx; in.hr = swhr;
  const double* A = H.arena.data();
  const SwBand* bands = H.sw;
  Tab tb = H.tab;
  if (!std::getenv("RRTM_EMU_SW_COL")) launch(ncol, SW_THREADS, [=]() { rrtmg_sw_kernel(A, tb, bands, in); });
  else launch((ncol + 31) / 32, 32 * SWC_WARPS, [=]() { rrtmg_sw_col_kernel(A, tb, bands, in); });
  return 0;
}

// the kernel sequence of isca_rrtm_run_device (rrtm.cu) on host arrays in the model layout
int rrtm_emu_run_rrtmg(const char* table_path, int I, int J, int K, int lonstep, double cp_air, double rdgas, double gas_constant,
                       double wtmh2o, double wtmozone, double co2ppmv, double h2o_lower_limit, double t_lo, double t_hi, double solrad,
                       double solr_cnst, const double* p_full, const double* p_half, const double* z_full, const double* z_half,
                       const double* t, const double* q, const double* o3, const double* t_surf, const double* albedo, const double* coszen,
                       double* tdt, double* tdt_rad, double* flux_sw, double* flux_lw, double* olr, double* toa_sw) {
  if (!ensure_tables(table_path)) return 1;
  HostTables& H = *g_tab;
  const int ls = lonstep; const size_t nm = (size_t)I * J, nc = nm / ls;
  const size_t n3 = nc * K, n3h = nc * (K + 1);
  std::vector<double> play(n3), plev(n3h), tlay(n3), tlev(n3h), h2o(n3), o3v(n3), swu(n3h), swd(n3h), swhr(n3), lwu(n3h), lwd(n3h), lwhr(n3),
      ts_s(nc), al_s(nc), cz_s(nc);
  PrepArgs pa{(int)nc, K, p_full, p_half, z_full, z_half, t, q, o3, play.data(), plev.data(), tlay.data(), tlev.data(), h2o.data(), o3v.data(),
              (1000.0 * gas_constant / rdgas) / wtmh2o, (1000.0 * gas_constant / rdgas) / wtmozone, h2o_lower_limit, t_lo, t_hi, 1,
              ls, I, nm, t_surf, albedo, coszen, ts_s.data(), al_s.data(), cz_s.data()};
  const int T = 128, G = (int)((nc + T - 1) / T), Gm = (int)((nm + T - 1) / T);
  launch(G, T, [=]() { rrtm_prepare_kernel(pa); });
  double* plev_p = plev.data(); const double* play_p = play.data();
  launch(G, T, [=]() { rrtm_fix_top_kernel((int)nc, K, play_p, plev_p); });
  ColIn in;
  std::memset(&in, 0, sizeof in);
  in.ncol = (int)nc; in.nlay = K; in.play = play.data(); in.plev = plev.data(); in.tlay = tlay.data(); in.tlev = tlev.data();
  in.tsfc = ls > 1 ? ts_s.data() : t_surf; in.albedo = ls > 1 ? al_s.data() : albedo; in.coszen = ls > 1 ? cz_s.data() : coszen;
  in.gas[SP_H2O] = h2o.data(); in.gas[SP_O3] = o3v.data(); in.gas_c[SP_CO2] = co2ppmv * 1.0e-6;
  in.heatfac = GRAV * SECDY / (cp_air * 1.0e2);
  const double* A = H.arena.data();
  Tab tb = H.tab;
  ColIn sw = in; sw.uflx = swu.data(); sw.dflx = swd.data(); sw.hr = swhr.data(); sw.adjflux = solrad * (solr_cnst / 1.36822e+03);
  const SwBand* sb = H.sw;
  launch((int)nc, SW_THREADS, [=]() { rrtmg_sw_kernel(A, tb, sb, sw); });
  ColIn lw = in; lw.uflx = lwu.data(); lw.dflx = lwd.data(); lw.hr = lwhr.data();
  const LwBand* lb = H.lw;
  std::vector<double> lays((size_t)LAYP_N * nc * K);
  lw.lays = lays.data();
  launch(((int)nc * K + 127) / 128, 128, [=]() { rrtmg_lw_setcoef_kernel(A, tb, lw); });
  launch(((int)nc + 31) / 32, 32 * LWC_WARPS, [=]() { rrtmg_lw_col_kernel(A, tb, lb, lw); });
  FinishArgs fa{(int)nm, K, swhr.data(), lwhr.data(), swu.data(), swd.data(), lwu.data(), lwd.data(), tdt, tdt_rad, flux_sw, flux_lw, olr, toa_sw,
                ls, I};
  launch(Gm, T, [=]() { rrtm_finish_kernel(fa); });
  return 0;
}

// dry_convection_kernel (physics_dry_kernels.h) on host planes [K][ncol]
int emu_dry_convection(int ncol, int K, double tau, double gamma, double cons1, double rdgas, const double* tg, const double* p_full,
                       const double* p_half, double* dt_tg, double* cape, double* cin, int* lzb, int* lcl, int* err) {
  std::vector<double> tp((size_t)ncol * K);
  double* tpp = tp.data();
  launch((ncol + 127) / 128, 128, [=]() {
    dryconv_k::dry_convection_kernel(ncol, K, tau, gamma, cons1, rdgas, tg, p_full, p_half, tpp, dt_tg, cape, cin, lzb, lcl, err);
  });
  return 0;
}

int rrtm_emu_coszen(int n, const double* lat, const double* lon, double gmt, double dec, double dt, int frierson, double del_sol, double del_sw,
                    double* cosz, double* fracday) {
  launch((n + 127) / 128, 128, [=]() { coszen_kernel(n, lat, lon, gmt, dec, dt, frierson, del_sol, del_sw, cosz, fracday); });
  return 0;
}

}  // extern "C"
