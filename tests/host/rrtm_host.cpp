// TEST INFRASTRUCTURE ONLY -- serial host build of the RRTMG column arithmetic of isca_b200/csrc/rrtm_column.h.
//
// The CUDA kernels of rrtm.cu call the `__host__ __device__` functions of rrtm_column.h from (column, g-point) threads;
// this file calls the very same functions (and the same descriptor builder, rrtm_tables.h) in plain loops, so that
// `pytest -m "not gpu"` can check the device arithmetic against the NumPy oracle on a machine without a GPU.
// It is compiled by tests/test_rrtm_host.py with g++ into tests/host/_build/ and is never linked into the product library
// (isca_b200/lib/libisca_b200.so fails at create time when no CUDA device exists).  Built with -fopenmp (oracle/_build/, by
// __graft_entry__.build() and bench.py) the column loops run on all host cores: that build is bench.py's C++/OpenMP CPU baseline
// of one RRTMG radiation call (`cpu_baseline`, `--impl reference`) -- the same arithmetic, term lists built once per (layer, band).
#include "../../isca_b200/csrc/rrtm_tables.h"
#include <memory>
#include <string>

using namespace rrtm;

namespace {
struct HostRed {                    // sequential sum over the g-points (the reference's order within a band)
  double* u; double* d;
  void up(int lev, double v) { u[lev] += v; }
  void down(int lev, double v) { d[lev] += v; }
};
std::unique_ptr<HostTables> g_tab;
std::string g_err;
bool ensure_tables(const char* path) {
  if (g_tab) return true;
  g_tab.reset(new HostTables());
  if (!g_tab->build(path)) { g_err = g_tab->err; g_tab.reset(); return false; }
  return true;
}
const double FLUXFAC = 3.14159265358979323846 * 2.0e4;
}  // namespace

extern "C" {

const char* rrtm_host_error() { return g_err.c_str(); }

// same argument meaning as isca_b200_rrtmg_lw; gas pointers may be NULL (= 0)
int rrtm_host_lw(const char* table_path, double cp_air, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                 const double* tlev, const double* tsfc, const double* h2ovmr, const double* o3vmr, const double* co2vmr,
                 const double* ch4vmr, const double* n2ovmr, const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                 const double* cfc22vmr, const double* ccl4vmr, const double* emis, double* uflx, double* dflx, double* hr,
                 double* taug_out, double* fracs_out) {
  if (!ensure_tables(table_path)) return 1;
  if (nlay > KMAX) { g_err = "nlay > KMAX"; return 1; }
  HostTables& H = *g_tab;
  const double* A = H.arena.data();
  const double* gas[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  const double* xsp[4] = {ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr};
  const double heatfac = GRAV * SECDY / (cp_air * 1.0e2);
  const double delwave[NB_LW] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
  const int PS = KMAX + 1;
  const size_t nc = ncol;
#pragma omp parallel for schedule(dynamic, 32)
  for (int col = 0; col < ncol; ++col) {
    std::vector<Layer> lay(KMAX);
    std::vector<double> planklay(NB_LW * PS), planklev(NB_LW * PS);
    std::vector<LwRec> recs(KMAX);
    double pz[KMAX + 1], semiss[NB_LW], plankbnd[NB_LW], secdiff[NB_LW];
    for (int l = 0; l < nlay; ++l) {
      double vmr[NSP], xs[4];
      for (int i = 0; i < NSP; ++i) vmr[i] = gas[i] ? gas[i][col + nc * l] : 0.0;
      for (int i = 0; i < 4; ++i) xs[i] = xsp[i] ? xsp[i][col + nc * l] : 0.0;
      double pb = plev[col + nc * l], pa = plev[col + nc * (l + 1)];
      double tav = tlay[col + nc * l];
      lw_setcoef_layer(A, H.tab, play[col + nc * l], tav, coldry_of(pb, pa, vmr[0]), vmr, xs, lay[l]);
      planck16(A, H.tab, tav, planklay.data() + l, PS);
      planck16(A, H.tab, tlev[col + nc * (l + 1)], planklev.data() + l + 1, PS);
      pz[l + 1] = pa;
      if (l == 0) pz[0] = pb;
    }
    for (int ib = 0; ib < NB_LW; ++ib) semiss[ib] = emis ? emis[col + nc * ib] : 1.0;
    planck16(A, H.tab, tlev[col], planklev.data(), PS);
    double pb16[NB_LW];
    planck16(A, H.tab, tsfc[col], pb16, 1);
    for (int ib = 0; ib < NB_LW; ++ib) plankbnd[ib] = semiss[ib] * pb16[ib];
    double amttl = 0.0, wvttl = 0.0;
    for (int l = 0; l < nlay; ++l) {
      double wv = lay[l].col[SP_H2O] * 1.0e20;
      amttl += lay[l].coldry + wv;
      wvttl += wv;
    }
    double wvsh = (AMW * wvttl) / (AMD * amttl);
    double pwvcm = wvsh * (1.0e3 * pz[0]) / (1.0e2 * GRAV);
    for (int ib = 0; ib < NB_LW; ++ib) secdiff[ib] = lw_secdiff(ib, pwvcm);
    double u[KMAX + 1] = {0}, d[KMAX + 1] = {0};
    HostRed red{u, d};
    for (int ib = 0; ib < NB_LW; ++ib) {
      lw_band_recs(A, H.tab, H.lw[ib], nlay, lay.data(), recs.data());
      for (int g = 0; g < H.lw[ib].ng; ++g) {
        lw_gpoint_recs(A, H.tab, recs.data(), ib, g, nlay, planklay.data(), planklev.data(), PS, plankbnd[ib], semiss[ib],
                       secdiff[ib], 0.5 * delwave[ib], red);
        if (taug_out)
          for (int l = 0; l < nlay; ++l) {
            double tau, fr;
            LwRec rec;
            lw_terms(A, H.tab, H.lw[ib], lay[l], rec);          // the term lists the kernels build per (layer, band) ...
            lw_tau_rec(A, rec, g, tau, fr);                       // ... and the per-g-point dot product
            taug_out[((size_t)col * nlay + l) * NG_LW + H.lw[ib].g0 + g] = tau;
            fracs_out[((size_t)col * nlay + l) * NG_LW + H.lw[ib].g0 + g] = fr;
          }
      }
    }
    double fnet[KMAX + 1];
    for (int lev = 0; lev <= nlay; ++lev) {
      uflx[col + nc * lev] = u[lev] * FLUXFAC;
      dflx[col + nc * lev] = d[lev] * FLUXFAC;
      fnet[lev] = uflx[col + nc * lev] - dflx[col + nc * lev];
    }
    for (int l = 0; l < nlay; ++l) hr[col + nc * l] = heatfac * (fnet[l] - fnet[l + 1]) / (pz[l] - pz[l + 1]);
  }
  return 0;
}

int rrtm_host_sw(const char* table_path, double cp_air, int ncol, int nlay, const double* play, const double* plev, const double* tlay,
                 const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                 const double* o2vmr, const double* albedo, const double* coszen, double adjes, double scon, double* swuflx,
                 double* swdflx, double* swhr, double* taug_out, double* taur_out, double* sflux_out) {
  if (!ensure_tables(table_path)) return 1;
  if (nlay > KMAX) { g_err = "nlay > KMAX"; return 1; }
  HostTables& H = *g_tab;
  const double* A = H.arena.data();
  const double* gas[NSP] = {h2ovmr, co2vmr, o3vmr, n2ovmr, nullptr, ch4vmr, o2vmr};
  const double heatfac = GRAV * SECDY / (cp_air * 1.0e2);
  const double adjflux = adjes * (scon / 1.36822e+03);
  const size_t nc = ncol;
#pragma omp parallel for schedule(dynamic, 32)
  for (int col = 0; col < ncol; ++col) {
    std::vector<Layer> lay(KMAX);
    std::vector<SwRec> recs(KMAX);
    if (coszen[col] < 1.0e-10) {
      for (int lev = 0; lev <= nlay; ++lev) { swuflx[col + nc * lev] = 0.0; swdflx[col + nc * lev] = 0.0; }
      for (int l = 0; l < nlay; ++l) swhr[col + nc * l] = 0.0;
      continue;
    }
    double pz[KMAX + 1];
    int laytrop = 0;
    for (int l = 0; l < nlay; ++l) {
      double vmr[NSP];
      for (int i = 0; i < NSP; ++i) vmr[i] = gas[i] ? gas[i][col + nc * l] : 0.0;
      double pb = plev[col + nc * l], pa = plev[col + nc * (l + 1)];
      sw_setcoef_layer(A, H.tab, play[col + nc * l], tlay[col + nc * l], coldry_of(pb, pa, vmr[0]), vmr, lay[l]);
      pz[l + 1] = pa;
      if (l == 0) pz[0] = pb;
      laytrop += lay[l].lower;
    }
    double u[KMAX + 1] = {0}, d[KMAX + 1] = {0};
    HostRed red{u, d};
    for (int ib = 0; ib < NB_SW; ++ib) {
      int lsol = sw_laysolfr(H.sw[ib], lay.data(), nlay, laytrop);
      sw_band_recs(A, H.sw[ib], nlay, lay.data(), recs.data());
      for (int g = 0; g < H.sw[ib].ng; ++g) {
        sw_gpoint_recs(A, H.tab, H.sw[ib], recs.data(), g, nlay, lsol, coszen[col], albedo[col], adjflux, 1.0, red);
        if (taug_out)
          for (int l = 0; l < nlay; ++l) {
            double tg, tr, src;
            SwRec rec;
            sw_terms(A, H.sw[ib], lay[l], rec);
            sw_tau_rec(A, rec, g, tg, tr);
            src = sw_src_rec(A, H.sw[ib], rec, g);
            taug_out[((size_t)col * nlay + l) * NG_SW + H.sw[ib].g0 + g] = tg;
            taur_out[((size_t)col * nlay + l) * NG_SW + H.sw[ib].g0 + g] = tr;
            if (l + 1 == lsol) sflux_out[(size_t)col * NG_SW + H.sw[ib].g0 + g] = src;
          }
      }
    }
    double fnet[KMAX + 1];
    for (int lev = 0; lev <= nlay; ++lev) {
      swuflx[col + nc * lev] = u[lev];
      swdflx[col + nc * lev] = d[lev];
      fnet[lev] = d[lev] - u[lev];
    }
    for (int l = 0; l < nlay; ++l) swhr[col + nc * l] = l == nlay - 1 ? 0.0 : (fnet[l + 1] - fnet[l]) * heatfac / (pz[l] - pz[l + 1]);
  }
  return 0;
}

}  // extern "C"
