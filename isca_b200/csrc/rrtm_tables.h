// rrtm_tables.h -- host side of the RRTMG tables: reads the coefficient file written by tools/make_rrtmg_tables.py
// (the data of rrtmg_lw_k_g.f90 / rrtmg_sw_k_g.f90 after the g-point reduction of rrtmg_lw_ini / rrtmg_sw_ini), re-tiles
// every table with the g-point index fastest into one arena, builds the exp / tfn lookup tables of rrtmg_lw_ini
// (rrtmg_lw_init.f90:118-138) and the per-band descriptors that replace taugb1..16 / taumol16..29 (see rrtm_column.h).
// Plain C++ (no CUDA) so that the test-only host build can use it too.
#pragma once
#include "rrtm_column.h"
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace rrtm {

struct RawTable { int ndim; int dims[6]; std::vector<double> v; };

struct HostTables {
  std::vector<double> arena;
  Tab tab;
  LwBand lw[NB_LW];
  SwBand sw[NB_SW];
  std::map<std::string, RawTable> raw;
  std::string err;

  bool read_file(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open RRTMG table file ") + path; return false; }
    char magic[8]; int32_t ver = 0, n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "ISCARRTM", 8) != 0 || fread(&ver, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || ver != 1) {
      fclose(f); err = "bad RRTMG table file header"; return false;
    }
    struct Ent { char name[48]; int32_t ndim; int32_t dims[6]; int64_t off; };
    std::vector<Ent> ents(n);
    for (int i = 0; i < n; ++i) {
      Ent& e = ents[i];
      if (fread(e.name, 1, 48, f) != 48 || fread(&e.ndim, 4, 1, f) != 1 || fread(e.dims, 4, 6, f) != 6 || fread(&e.off, 8, 1, f) != 1) {
        fclose(f); err = "truncated RRTMG table directory"; return false;
      }
    }
    long base = ftell(f);
    for (int i = 0; i < n; ++i) {
      const Ent& e = ents[i];
      RawTable t; t.ndim = e.ndim; size_t sz = 1;
      for (int d = 0; d < 6; ++d) { t.dims[d] = e.dims[d]; if (d < e.ndim) sz *= (size_t)e.dims[d]; }
      t.v.resize(sz);
      fseek(f, base + (long)e.off * 8, SEEK_SET);
      if (fread(t.v.data(), 8, sz, f) != sz) { fclose(f); err = "truncated RRTMG table payload"; return false; }
      char nm[49]; memcpy(nm, e.name, 48); nm[48] = 0;
      raw[nm] = std::move(t);
    }
    fclose(f);
    return true;
  }

  const RawTable* get(const std::string& name) {
    auto it = raw.find(name);
    if (it == raw.end()) { if (err.empty()) err = "RRTMG table missing: " + name; return nullptr; }
    return &it->second;
  }
  // table whose LAST Fortran dimension is the g-point: (rows..., ng) -> arena [row][g]
  int add_kg(const std::string& name) {
    const RawTable* t = get(name);
    if (!t) return -1;
    int ng = t->dims[t->ndim - 1];
    size_t rows = t->v.size() / ng;
    if (arena.size() & 1) arena.push_back(0.0);               // every table starts on a 16-byte boundary (128-bit row reads)
    int off = (int)arena.size();
    arena.resize(arena.size() + t->v.size());
    for (size_t r = 0; r < rows; ++r)
      for (int g = 0; g < ng; ++g) arena[off + r * ng + g] = t->v[r + rows * g];
    return off;
  }
  // table already g-fastest ((ng) or (ng, nsp)) or shared (copied verbatim)
  int add_raw(const std::string& name) {
    const RawTable* t = get(name);
    if (!t) return -1;
    if (arena.size() & 1) arena.push_back(0.0);
    int off = (int)arena.size();
    arena.insert(arena.end(), t->v.begin(), t->v.end());
    return off;
  }
  int add_vec(const std::vector<double>& v) {
    if (arena.size() & 1) arena.push_back(0.0);
    int off = (int)arena.size(); arena.insert(arena.end(), v.begin(), v.end()); return off;
  }
  double chi(int sp1, int lev1) { const RawTable* t = get("lw_chi_mls"); return t ? t->v[(sp1 - 1) + 7 * (lev1 - 1)] : 0.0; }

  static std::string nm(const char* fam, int band, const char* what) {
    char b[64]; snprintf(b, sizeof b, "%s%02d_%s", fam, band, what); return b;
  }

  // ---- LW descriptor helpers ----
  void lw_region(LwRegion& R, int band, bool lower, int major, int spA, int spB, bool self, bool forc, int frac2d, double refrat_planck) {
    memset(&R, 0, sizeof R);
    R.major = major; R.spA = spA; R.spB = spB;
    R.nsp = major == 2 ? (lower ? 9 : 5) : 1;
    R.k_off = major ? add_kg(nm("lw", band, lower ? "ka" : "kb")) : -1;
    R.self_off = self ? add_kg(nm("lw", band, "selfref")) : -1;
    R.for_off = forc ? add_kg(nm("lw", band, "forref")) : -1;
    R.frac2d = frac2d; R.refrat_planck = refrat_planck;
    R.frac_off = -1; R.gscale_off = -1;
  }
  void lw_frac(LwRegion& R, int band, const char* which) { R.frac_off = add_raw(nm("lw", band, which)); }
  void lw_minor(LwRegion& R, int band, const char* tbl, int binary, double refrat, int scale, int sp, double thresh = 0, double base = 0,
                double expo = 0, double chiref = 0) {
    Minor& M = R.minor[R.nminor++];
    M.k_off = add_kg(nm("lw", band, tbl)); M.binary = binary; M.refrat = refrat; M.scale = scale; M.sp = sp;
    M.thresh = thresh; M.base = base; M.expo = expo; M.chiref = chiref;
  }
  void lw_cfc(LwRegion& R, int band, const char* tbl, int wx) { R.cfc_wx[R.ncfc] = wx; R.cfc_off[R.ncfc] = add_raw(nm("lw", band, tbl)); R.ncfc++; }

  bool build(const char* path) {
    if (!read_file(path)) return false;
    arena.clear();
    tab.preflog = add_raw("lw_preflog"); tab.tref = add_raw("lw_tref"); tab.chi = add_raw("lw_chi_mls");
    tab.totplnk = add_raw("lw_totplnk"); tab.sw_preflog = add_raw("sw_preflog"); tab.sw_tref = add_raw("sw_tref");
    {   // rrtmg_lw_init.f90:118-138 (the SW table, rrtmg_sw_init.f90, is the same exp_tbl)
      std::vector<double> ex(NTBL + 1), tfn(NTBL + 1);
      const double expeps = 1.0e-20;
      ex[0] = 1.0; ex[NTBL] = expeps; tfn[0] = 0.0; tfn[NTBL] = 1.0;
      for (int itr = 1; itr < NTBL; ++itr) {
        double t = (double)itr / (double)NTBL;
        double tau = BPADE * t / (1.0 - t);
        double e = exp(-tau);
        if (e <= expeps) e = expeps;
        ex[itr] = e;
        tfn[itr] = tau < 0.06 ? tau / 6.0 : 1.0 - 2.0 * ((1.0 / tau) - (e / (1.0 - e)));
      }
      tab.exp_tbl = add_vec(ex); tab.tfn_tbl = add_vec(tfn);
      std::vector<double> both(2 * (NTBL + 1));
      for (int itr = 0; itr <= NTBL; ++itr) { both[2 * itr] = ex[itr]; both[2 * itr + 1] = tfn[itr]; }
      if (arena.size() & 1) arena.push_back(0.0);            // 16-byte alignment of the pairs
      tab.exptfn = add_vec(both);
    }
    const int lw_ngc[NB_LW] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
    const int sw_ngc[NB_SW] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
    int g0 = 0;
    for (int b = 0; b < NB_LW; ++b) { lw[b].ng = lw_ngc[b]; lw[b].g0 = g0; g0 += lw_ngc[b]; }
    g0 = 0;
    for (int b = 0; b < NB_SW; ++b) { sw[b].ng = sw_ngc[b]; sw[b].g0 = g0; g0 += sw_ngc[b]; }
    const int H = SP_H2O, C = SP_CO2, O3 = SP_O3, N = SP_N2O, CO = SP_CO, M4 = SP_CH4, O2 = SP_O2;
    auto CH = [&](int sp0, int lev) { return chi(sp0 + 1, lev); };     // chi_mls(species, level), species 0-based here
    LwRegion* R;
    // band 1 (taugb1): h2o; minor n2
    R = &lw[0].r[0]; lw_region(*R, 1, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 1, "fracrefa"); R->corr = 1;
    lw_minor(*R, 1, "ka_mn2", 0, 0, SC_BRD_N2, 0);
    R = &lw[0].r[1]; lw_region(*R, 1, false, 1, H, -1, false, true, 0, 0); lw_frac(*R, 1, "fracrefb"); R->corr = 2;
    lw_minor(*R, 1, "kb_mn2", 0, 0, SC_BRD_N2, 0);
    // band 2: h2o
    R = &lw[1].r[0]; lw_region(*R, 2, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 2, "fracrefa"); R->corr = 3;
    R = &lw[1].r[1]; lw_region(*R, 2, false, 1, H, -1, false, true, 0, 0); lw_frac(*R, 2, "fracrefb");
    // band 3: (h2o, co2); minor n2o
    R = &lw[2].r[0]; lw_region(*R, 3, true, 2, H, C, true, true, 1, CH(H, 9) / CH(C, 9)); lw_frac(*R, 3, "fracrefa");
    lw_minor(*R, 3, "ka_mn2o", 1, CH(H, 3) / CH(C, 3), SC_ADJ, N, 1.5, 0.5, 0.65);
    R = &lw[2].r[1]; lw_region(*R, 3, false, 2, H, C, false, true, 1, CH(H, 13) / CH(C, 13)); lw_frac(*R, 3, "fracrefb");
    lw_minor(*R, 3, "kb_mn2o", 1, CH(H, 13) / CH(C, 13), SC_ADJ, N, 1.5, 0.5, 0.65);
    // band 4: (h2o, co2) / (o3, co2)
    R = &lw[3].r[0]; lw_region(*R, 4, true, 2, H, C, true, true, 1, CH(H, 11) / CH(C, 11)); lw_frac(*R, 4, "fracrefa");
    R = &lw[3].r[1]; lw_region(*R, 4, false, 2, O3, C, false, false, 1, CH(O3, 13) / CH(C, 13)); lw_frac(*R, 4, "fracrefb");
    {
      std::vector<double> s(14, 1.0);
      const double f[7] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.88, 0.943};
      for (int i = 0; i < 7; ++i) s[7 + i] = f[i];
      R->gscale_off = add_vec(s);
    }
    // band 5: (h2o, co2; minor o3, ccl4) / (o3, co2; ccl4)
    R = &lw[4].r[0]; lw_region(*R, 5, true, 2, H, C, true, true, 1, CH(H, 5) / CH(C, 5)); lw_frac(*R, 5, "fracrefa");
    lw_minor(*R, 5, "ka_mo3", 1, CH(H, 7) / CH(C, 7), SC_COL, O3); lw_cfc(*R, 5, "ccl4", 0);
    R = &lw[4].r[1]; lw_region(*R, 5, false, 2, O3, C, false, false, 1, CH(O3, 43) / CH(C, 43)); lw_frac(*R, 5, "fracrefb");
    lw_cfc(*R, 5, "ccl4", 0);
    // band 6: h2o; minor co2, cfc11, cfc12 / cfc only
    R = &lw[5].r[0]; lw_region(*R, 6, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 6, "fracrefa");
    lw_minor(*R, 6, "ka_mco2", 0, 0, SC_ADJ, C, 3.0, 2.0, 0.77); lw_cfc(*R, 6, "cfc11adj", 1); lw_cfc(*R, 6, "cfc12", 2);
    R = &lw[5].r[1]; lw_region(*R, 6, false, 0, H, -1, false, false, 0, 0); lw_frac(*R, 6, "fracrefa");
    lw_cfc(*R, 6, "cfc11adj", 1); lw_cfc(*R, 6, "cfc12", 2);
    // band 7: (h2o, o3; minor co2) / o3; minor co2
    R = &lw[6].r[0]; lw_region(*R, 7, true, 2, H, O3, true, true, 1, CH(H, 3) / CH(O3, 3)); lw_frac(*R, 7, "fracrefa");
    lw_minor(*R, 7, "ka_mco2", 1, CH(H, 3) / CH(O3, 3), SC_ADJ, C, 3.0, 3.0, 0.79);
    R = &lw[6].r[1]; lw_region(*R, 7, false, 1, O3, -1, false, false, 0, 0); lw_frac(*R, 7, "fracrefb");
    lw_minor(*R, 7, "kb_mco2", 0, 0, SC_ADJ, C, 3.0, 2.0, 0.79);
    {
      std::vector<double> s(12, 1.0);
      const double f[6] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.855};
      for (int i = 0; i < 6; ++i) s[5 + i] = f[i];
      R->gscale_off = add_vec(s);
    }
    // band 8: h2o; minor co2, o3, n2o, cfc12, cfc22 / o3; minor co2, n2o, cfcs
    R = &lw[7].r[0]; lw_region(*R, 8, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 8, "fracrefa");
    lw_minor(*R, 8, "ka_mco2", 0, 0, SC_ADJ, C, 3.0, 2.0, 0.65); lw_minor(*R, 8, "ka_mo3", 0, 0, SC_COL, O3);
    lw_minor(*R, 8, "ka_mn2o", 0, 0, SC_COL, N); lw_cfc(*R, 8, "cfc12", 2); lw_cfc(*R, 8, "cfc22adj", 3);
    R = &lw[7].r[1]; lw_region(*R, 8, false, 1, O3, -1, false, false, 0, 0); lw_frac(*R, 8, "fracrefb");
    lw_minor(*R, 8, "kb_mco2", 0, 0, SC_ADJ, C, 3.0, 2.0, 0.65); lw_minor(*R, 8, "kb_mn2o", 0, 0, SC_COL, N);
    lw_cfc(*R, 8, "cfc12", 2); lw_cfc(*R, 8, "cfc22adj", 3);
    // band 9: (h2o, ch4; minor n2o) / ch4; minor n2o
    R = &lw[8].r[0]; lw_region(*R, 9, true, 2, H, M4, true, true, 1, CH(H, 9) / CH(M4, 9)); lw_frac(*R, 9, "fracrefa");
    lw_minor(*R, 9, "ka_mn2o", 1, CH(H, 3) / CH(M4, 3), SC_ADJ, N, 1.5, 0.5, 0.65);
    R = &lw[8].r[1]; lw_region(*R, 9, false, 1, M4, -1, false, false, 0, 0); lw_frac(*R, 9, "fracrefb");
    lw_minor(*R, 9, "kb_mn2o", 0, 0, SC_ADJ, N, 1.5, 0.5, 0.65);
    // band 10: h2o
    R = &lw[9].r[0]; lw_region(*R, 10, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 10, "fracrefa");
    R = &lw[9].r[1]; lw_region(*R, 10, false, 1, H, -1, false, true, 0, 0); lw_frac(*R, 10, "fracrefb");
    // band 11: h2o; minor o2
    R = &lw[10].r[0]; lw_region(*R, 11, true, 1, H, -1, true, true, 0, 0); lw_frac(*R, 11, "fracrefa");
    lw_minor(*R, 11, "ka_mo2", 0, 0, SC_O2, O2);
    R = &lw[10].r[1]; lw_region(*R, 11, false, 1, H, -1, false, true, 0, 0); lw_frac(*R, 11, "fracrefb");
    lw_minor(*R, 11, "kb_mo2", 0, 0, SC_O2, O2);
    // band 12: (h2o, co2) / nothing
    R = &lw[11].r[0]; lw_region(*R, 12, true, 2, H, C, true, true, 1, CH(H, 10) / CH(C, 10)); lw_frac(*R, 12, "fracrefa");
    R = &lw[11].r[1]; lw_region(*R, 12, false, 0, H, -1, false, false, 0, 0);
    // band 13: (h2o, n2o; minor co2, co) / minor o3
    R = &lw[12].r[0]; lw_region(*R, 13, true, 2, H, N, true, true, 1, CH(H, 5) / CH(N, 5)); lw_frac(*R, 13, "fracrefa");
    lw_minor(*R, 13, "ka_mco2", 1, CH(H, 1) / CH(N, 1), SC_ADJ, C, 3.0, 2.0, 0.68, 3.55e-4);
    lw_minor(*R, 13, "ka_mco", 1, CH(H, 3) / CH(N, 3), SC_COL, CO);
    R = &lw[12].r[1]; lw_region(*R, 13, false, 0, H, -1, false, false, 0, 0); lw_frac(*R, 13, "fracrefb");
    lw_minor(*R, 13, "kb_mo3", 0, 0, SC_COL, O3);
    // band 14: co2
    R = &lw[13].r[0]; lw_region(*R, 14, true, 1, C, -1, true, true, 0, 0); lw_frac(*R, 14, "fracrefa");
    R = &lw[13].r[1]; lw_region(*R, 14, false, 1, C, -1, false, false, 0, 0); lw_frac(*R, 14, "fracrefb");
    // band 15: (n2o, co2; minor n2) / nothing
    R = &lw[14].r[0]; lw_region(*R, 15, true, 2, N, C, true, true, 1, CH(N, 1) / CH(C, 1)); lw_frac(*R, 15, "fracrefa");
    lw_minor(*R, 15, "ka_mn2", 1, CH(N, 1) / CH(C, 1), SC_BRD, 0);
    R = &lw[14].r[1]; lw_region(*R, 15, false, 0, H, -1, false, false, 0, 0);
    // band 16: (h2o, ch4) / ch4
    R = &lw[15].r[0]; lw_region(*R, 16, true, 2, H, M4, true, true, 1, CH(H, 6) / CH(M4, 6)); lw_frac(*R, 16, "fracrefa");
    R = &lw[15].r[1]; lw_region(*R, 16, false, 1, M4, -1, false, false, 0, 0); lw_frac(*R, 16, "fracrefb");

    // ---- SW (taumol16..29) ----
    auto swr = [&](SwRegion& S, int band, bool lower, int major, int spA, int spB, double strrat, double kscale, bool self, bool forc,
                   const char* rayl, int rayl_mode) {
      memset(&S, 0, sizeof S);
      S.major = major; S.spA = spA; S.spB = spB; S.strrat = strrat; S.kscale = kscale;
      S.nsp = major == 2 ? (lower ? 9 : 5) : 1;
      S.k_off = major ? add_kg(nm("sw", band, lower ? "ka" : "kb")) : -1;
      S.self_off = self ? add_kg(nm("sw", band, "selfref")) : -1;
      S.for_off = forc ? add_kg(nm("sw", band, "forref")) : -1;
      S.rayl_off = add_raw(nm("sw", band, rayl)); S.rayl_mode = rayl_mode;
    };
    auto extra = [&](SwRegion& S, int band, const char* tbl, int sp) { S.extra_sp[S.nextra] = sp; S.extra_off[S.nextra] = add_raw(nm("sw", band, tbl)); S.nextra++; };
    auto src = [&](int band, int sflux2d, int upper, int layreffr, double scale) {
      SwBand& B = sw[band - 16];
      B.sflux_off = add_raw(nm("sw", band, "sfluxref")); B.sflux2d = sflux2d; B.sflux_upper = upper; B.layreffr = layreffr; B.sflux_scale = scale;
    };
    SwBand* B;
    B = &sw[0]; swr(B->r[0], 16, true, 2, H, M4, 252.131, 1, true, true, "rayl", 0); swr(B->r[1], 16, false, 1, M4, -1, 0, 1, false, false, "rayl", 0);
    src(16, 0, 1, 18, 1.0);
    B = &sw[1]; swr(B->r[0], 17, true, 2, H, C, 0.364641, 1, true, true, "rayl", 0); swr(B->r[1], 17, false, 2, H, C, 0.364641, 1, false, true, "rayl", 0);
    src(17, 1, 1, 30, 1.0);
    B = &sw[2]; swr(B->r[0], 18, true, 2, H, M4, 38.9589, 1, true, true, "rayl", 0); swr(B->r[1], 18, false, 1, M4, -1, 0, 1, false, false, "rayl", 0);
    src(18, 1, 0, 6, 1.0);
    B = &sw[3]; swr(B->r[0], 19, true, 2, H, C, 5.49281, 1, true, true, "rayl", 0); swr(B->r[1], 19, false, 1, C, -1, 0, 1, false, false, "rayl", 0);
    src(19, 1, 0, 3, 1.0);
    B = &sw[4]; swr(B->r[0], 20, true, 1, H, -1, 0, 1, true, true, "rayl", 0); extra(B->r[0], 20, "absch4", M4);
    swr(B->r[1], 20, false, 1, H, -1, 0, 1, false, true, "rayl", 0); extra(B->r[1], 20, "absch4", M4);
    src(20, 0, 0, 3, 1.0);
    B = &sw[5]; swr(B->r[0], 21, true, 2, H, C, 0.0045321, 1, true, true, "rayl", 0); swr(B->r[1], 21, false, 2, H, C, 0.0045321, 1, false, true, "rayl", 0);
    src(21, 1, 0, 8, 1.0);
    B = &sw[6]; swr(B->r[0], 22, true, 2, H, O2, 1.6 * 0.022708, 1, true, true, "rayl", 0); B->r[0].o2cont = 1;
    swr(B->r[1], 22, false, 1, O2, -1, 0, 1.6, false, false, "rayl", 0); B->r[1].o2cont = 1;
    src(22, 1, 0, 2, 1.0);
    B = &sw[7]; swr(B->r[0], 23, true, 1, H, -1, 0, 1.029, true, true, "rayl", 1); swr(B->r[1], 23, false, 0, H, -1, 0, 1, false, false, "rayl", 1);
    src(23, 0, 0, 6, 1.0);
    B = &sw[8]; swr(B->r[0], 24, true, 2, H, O2, 0.124692, 1, true, true, "rayla", 2); extra(B->r[0], 24, "abso3a", O3);
    swr(B->r[1], 24, false, 1, O2, -1, 0, 1, false, false, "raylb", 1); extra(B->r[1], 24, "abso3b", O3);
    src(24, 1, 0, 1, 1.0);
    B = &sw[9]; swr(B->r[0], 25, true, 1, H, -1, 0, 1, false, false, "rayl", 1); extra(B->r[0], 25, "abso3a", O3);
    swr(B->r[1], 25, false, 0, H, -1, 0, 1, false, false, "rayl", 1); extra(B->r[1], 25, "abso3b", O3);
    src(25, 0, 0, 2, 1.0);
    B = &sw[10]; swr(B->r[0], 26, true, 0, H, -1, 0, 1, false, false, "rayl", 1); swr(B->r[1], 26, false, 0, H, -1, 0, 1, false, false, "rayl", 1);
    src(26, 0, 0, 0, 1.0);
    B = &sw[11]; swr(B->r[0], 27, true, 1, O3, -1, 0, 1, false, false, "rayl", 1); swr(B->r[1], 27, false, 1, O3, -1, 0, 1, false, false, "rayl", 1);
    src(27, 0, 1, 32, 50.15 / 48.37);
    B = &sw[12]; swr(B->r[0], 28, true, 2, O3, O2, 6.67029e-07, 1, false, false, "rayl", 0); swr(B->r[1], 28, false, 2, O3, O2, 6.67029e-07, 1, false, false, "rayl", 0);
    src(28, 1, 1, 58, 1.0);
    B = &sw[13]; swr(B->r[0], 29, true, 1, H, -1, 0, 1, true, true, "rayl", 0); extra(B->r[0], 29, "absco2", C);
    swr(B->r[1], 29, false, 1, C, -1, 0, 1, false, false, "rayl", 0); extra(B->r[1], 29, "absh2o", H);
    src(29, 0, 1, 49, 1.0);
    raw.clear();
    return err.empty();
  }
};

}  // namespace rrtm
