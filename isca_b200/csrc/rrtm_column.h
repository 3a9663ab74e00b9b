// rrtm_column.h -- RRTMG clear-sky column arithmetic, written once for the CUDA kernels of rrtm.cu.
//
// Every function is `__host__ __device__`: the kernels call them per (column, g-point) thread, and the test-only host
// build (tests/host/rrtm_host.cpp, compiled by g++) calls the very same functions in a serial loop so that the device
// arithmetic can be checked against an independent NumPy restatement on a machine without a GPU.  The product never runs the host build.
//
// Replaces (paths relative to /root/reference/src/atmos_param/rrtm_radiation):
//   rrtmg_lw/gcm_model/src/rrtmg_lw_setcoef.f90:setcoef, rrtmg_lw_taumol.f90:taugb1..16,
//   rrtmg_lw_rtrnmr.f90:rtrnmr (clear layers), rrtmg_sw/gcm_model/src/rrtmg_sw_setcoef.f90:setcoef_sw,
//   rrtmg_sw_taumol.f90:taumol16..29, rrtmg_sw_reftra.f90, rrtmg_sw_vrtqdr.f90, rrtmg_sw_spcvrt.f90 (clear sky).
//
// Design: the sixteen (fourteen) hand-unrolled band routines of the reference become ONE generic evaluation driven by a
// per-band / per-region (below / above the ~100 hPa switch) descriptor: which key species, single or binary species-ratio
// interpolation, which continua, which minor gases and how their column amount is scaled.  All lanes of a warp run the
// same code whatever band their g-point belongs to.  Tables are re-tiled with the g-point index fastest so that the
// lanes of a band read consecutive doubles.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define RR_HD __host__ __device__ __forceinline__
#else
#define RR_HD inline
#endif

namespace rrtm {

constexpr int KMAX = 64;          // maximum number of model layers
constexpr int NB_LW = 16, NG_LW = 140, NB_SW = 14, NG_SW = 112;
constexpr int NTBL = 10000;
constexpr double TBLINT = 10000.0, BPADE = 1.0 / 0.278, ONEMINUS = 1.0 - 1.0e-6, STPFAC = 296.0 / 1013.0;
constexpr double GRAV = 9.8066, AVOGAD = 6.02214199e+23, SECDY = 8.6400e4, AMD = 28.9660, AMW = 18.0160;

// species order of wkl / chi_mls: h2o co2 o3 n2o co ch4 o2
enum { SP_H2O = 0, SP_CO2, SP_O3, SP_N2O, SP_CO, SP_CH4, SP_O2, NSP };
// how a minor gas' column amount is formed
enum { SC_COL = 0, SC_ADJ, SC_BRD_N2, SC_BRD, SC_O2 };

struct Minor {
  int k_off, binary, scale, sp;             // table [19 or nsp*19][ng]; binary: interpolated in the species ratio too
  double refrat, thresh, base, expo, chiref;  // chiref > 0: fixed reference mixing ratio (band 13), else chi_mls(sp, jp+1)
};
struct LwRegion {
  int major, spA, spB, k_off, nsp;          // major: 0 none, 1 single key species, 2 binary (spA, spB)
  int self_off, for_off, frac_off, frac2d;  // -1 = absent; frac2d: Planck fraction interpolated in the species ratio
  int nminor, ncfc, corr, gscale_off;
  double refrat_planck;
  Minor minor[3];
  int cfc_wx[2], cfc_off[2];
};
struct LwBand { int ng, g0; LwRegion r[2]; };

struct SwRegion {
  int major, spA, spB, k_off, nsp;
  double strrat, kscale;
  int self_off, for_off, nextra, o2cont, rayl_off, rayl_mode;   // rayl_mode 0 scalar, 1 per g, 2 per (g, js)
  int extra_sp[2], extra_off[2];
};
struct SwBand { int ng, g0; SwRegion r[2]; int sflux_off, sflux2d, sflux_upper, layreffr; double sflux_scale; };

struct Tab {                     // offsets (in doubles) into the arena
  int preflog, tref, chi, totplnk, exp_tbl, tfn_tbl, sw_preflog, sw_tref, exptfn;     // exptfn: (exp, tfn) pairs interleaved
};

struct Layer {                   // setcoef / setcoef_sw output of one layer
  double pavel, fac00, fac01, fac10, fac11, selffac, selffrac, forfac, forfrac, minorfrac, scaleminor, scaleminorn2,
         colbrd, coldry, colmol, col[NSP], wx[4];
  int jp, jt, jt1, indself, indfor, indminor, lower;
};

RR_HD double chi_mls(const double* A, const Tab& tb, int sp, int lev1) { return A[tb.chi + (lev1 - 1) * 7 + sp]; }

// jp, jt, jt1, fac00..fac11: identical in setcoef (rrtmg_lw_setcoef.f90:251-287,395-400) and setcoef_sw
RR_HD void pt_indices(const double* preflog, const double* tref, double pavel, double tavel, double& plog, Layer& L) {
  plog = log(pavel);
  int jp = (int)(36.0 - 5.0 * (plog + 0.04));
  jp = jp < 1 ? 1 : (jp > 58 ? 58 : jp);
  double fp = 5.0 * (preflog[jp - 1] - plog);
  int jt = (int)(3.0 + (tavel - tref[jp - 1]) / 15.0);
  jt = jt < 1 ? 1 : (jt > 4 ? 4 : jt);
  double ft = ((tavel - tref[jp - 1]) / 15.0) - (double)(jt - 3);
  int jt1 = (int)(3.0 + (tavel - tref[jp]) / 15.0);
  jt1 = jt1 < 1 ? 1 : (jt1 > 4 ? 4 : jt1);
  double ft1 = ((tavel - tref[jp]) / 15.0) - (double)(jt1 - 3);
  double compfp = 1.0 - fp;
  L.jp = jp; L.jt = jt; L.jt1 = jt1;
  L.fac10 = compfp * ft; L.fac00 = compfp * (1.0 - ft); L.fac11 = fp * ft1; L.fac01 = fp * (1.0 - ft1);
}

// inatm (rrtmg_lw_rad.nomcica.f90:774-812): dry-air column density of a layer
RR_HD double coldry_of(double pz_below, double pz_above, double h2ovmr) {
  double amm = (1.0 - h2ovmr) * AMD + h2ovmr * AMW;
  return (pz_below - pz_above) * 1.0e3 * AVOGAD / (1.0e2 * GRAV * amm * (1.0 + h2ovmr));
}

// setcoef of one layer (rrtmg_lw_setcoef.f90:251-402).  vmr: h2o co2 o3 n2o co ch4 o2; xs: ccl4 cfc11 cfc12 cfc22
RR_HD void lw_setcoef_layer(const double* A, const Tab& tb, double pavel, double tavel, double coldry, const double* vmr,
                            const double* xs, Layer& L) {
  double plog;
  pt_indices(A + tb.preflog, A + tb.tref, pavel, tavel, plog, L);
  L.pavel = pavel; L.coldry = coldry;
  double summol = 0.0;
  for (int i = 1; i < NSP; ++i) summol += vmr[i];
  double wbroad = coldry * (1.0 - summol);
  double wkl[NSP];
  for (int i = 0; i < NSP; ++i) wkl[i] = coldry * vmr[i];
  for (int i = 0; i < 4; ++i) L.wx[i] = coldry * xs[i] * 1.0e-20;
  double water = wkl[0] / coldry;
  double scalefac = pavel * STPFAC / tavel;
  L.lower = plog > 4.56;
  L.forfac = scalefac / (1.0 + water);
  L.selffac = water * L.forfac;
  if (L.lower) {
    double factor = (332.0 - tavel) / 36.0;
    int i = (int)factor; L.indfor = i < 1 ? 1 : (i > 2 ? 2 : i);
    L.forfrac = factor - (double)L.indfor;
  } else {
    L.indfor = 3;
    L.forfrac = (tavel - 188.0) / 36.0 - 1.0;
  }
  {
    double factor = (tavel - 188.0) / 7.2;
    int i = (int)factor - 7; L.indself = i < 1 ? 1 : (i > 9 ? 9 : i);
    L.selffrac = factor - (double)(L.indself + 7);
  }
  L.scaleminor = pavel / tavel;
  L.scaleminorn2 = (pavel / tavel) * (wbroad / (coldry + wkl[0]));
  {
    double factor = (tavel - 180.8) / 7.2;
    int i = (int)factor; L.indminor = i < 1 ? 1 : (i > 18 ? 18 : i);
    L.minorfrac = factor - (double)L.indminor;
  }
  for (int i = 0; i < NSP; ++i) L.col[i] = 1.0e-20 * wkl[i];
  // `if (colco2(lay) .eq. 0._rb) colco2(lay) = 1.e-32_rb * coldry(lay)` for co2, o3, n2o, co, ch4
  const int guard[5] = {SP_CO2, SP_O3, SP_N2O, SP_CO, SP_CH4};
  for (int i = 0; i < 5; ++i) if (L.col[guard[i]] == 0.0) L.col[guard[i]] = 1.0e-32 * coldry;
  L.colbrd = 1.0e-20 * wbroad;
  L.colmol = 0.0;
  L.selffac = L.col[SP_H2O] * L.selffac;
  L.forfac = L.col[SP_H2O] * L.forfac;
}

// setcoef_sw of one layer (rrtmg_sw_setcoef.f90)
RR_HD void sw_setcoef_layer(const double* A, const Tab& tb, double pavel, double tavel, double coldry, const double* vmr, Layer& L) {
  double plog;
  pt_indices(A + tb.sw_preflog, A + tb.sw_tref, pavel, tavel, plog, L);
  L.pavel = pavel; L.coldry = coldry;
  double wkl[NSP];
  for (int i = 0; i < NSP; ++i) wkl[i] = coldry * vmr[i];
  double water = wkl[0] / coldry;
  double scalefac = pavel * STPFAC / tavel;
  L.lower = plog > 4.56;
  L.forfac = scalefac / (1.0 + water);
  if (L.lower) {
    double factor = (332.0 - tavel) / 36.0;
    int i = (int)factor; L.indfor = i < 1 ? 1 : (i > 2 ? 2 : i);
    L.forfrac = factor - (double)L.indfor;
    L.selffac = water * L.forfac;
    factor = (tavel - 188.0) / 7.2;
    i = (int)factor - 7; L.indself = i < 1 ? 1 : (i > 9 ? 9 : i);
    L.selffrac = factor - (double)(L.indself + 7);
  } else {
    L.indfor = 3;
    L.forfrac = (tavel - 188.0) / 36.0 - 1.0;
    L.selffac = 0.0; L.selffrac = 0.0; L.indself = 1;
  }
  for (int i = 0; i < NSP; ++i) L.col[i] = 1.0e-20 * wkl[i];
  L.colmol = 1.0e-20 * coldry + L.col[SP_H2O];
  const int guard[4] = {SP_CO2, SP_N2O, SP_CH4, SP_O2};
  for (int i = 0; i < 4; ++i) if (L.col[guard[i]] == 0.0) L.col[guard[i]] = 1.0e-32 * coldry;
  L.minorfrac = L.scaleminor = L.scaleminorn2 = L.colbrd = 0.0; L.indminor = 1;
  for (int i = 0; i < 4; ++i) L.wx[i] = 0.0;
}

// index and fraction of the totplnk interpolation of setcoef (:154-250) for one temperature; planck_at = the band's value
RR_HD void planck_index(double t, int& ind, double& frac) {
  ind = (int)(t - 159.0);
  ind = ind < 1 ? 1 : (ind > 180 ? 180 : ind);
  frac = t - 159.0 - (double)ind;
}
RR_HD double planck_at(const double* A, const Tab& tb, int ib, int ind, double frac) {
  const double* p = A + tb.totplnk + ib * 181;
  return p[ind - 1] + frac * (p[ind] - p[ind - 1]);
}
// Planck function of one temperature for all 16 bands: the totplnk interpolation of setcoef (:154-250)
RR_HD void planck16(const double* A, const Tab& tb, double t, double* out, int stride) {
  int ind = (int)(t - 159.0);
  ind = ind < 1 ? 1 : (ind > 180 ? 180 : ind);
  double frac = t - 159.0 - (double)ind;
  for (int ib = 0; ib < NB_LW; ++ib) {
    const double* p = A + tb.totplnk + ib * 181;
    out[ib * stride] = p[ind - 1] + frac * (p[ind] - p[ind - 1]);
  }
}

struct Spec { double speccomb, specparm, fs; int js; };
RR_HD Spec specparm_of(double ca, double cb, double rat, double mult) {
  Spec s;
  s.speccomb = ca + rat * cb;
  s.specparm = ca / s.speccomb;
  if (s.specparm >= ONEMINUS) s.specparm = ONEMINUS;
  double specmult = mult * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = specmult - floor(specmult);         // = mod(specmult, 1.) exactly (specmult >= 0)
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Term lists (round 2).  Everything lw_tau / sw_tau do that does not depend on the g-point -- species ratios and their
// divisions, the three-branch interpolation weights, `pow` of the adjusted minor-gas amounts, continuum and minor-gas
// interpolation weights, the band's pressure correction -- is evaluated ONCE per (layer, band) into a list of
// (table row, weight) pairs; a g-point thread then only forms  tau = sum_i w_i * A[off_i + g]  (consecutive doubles across
// the lanes of a band).  Same formulas as lw_tau / sw_tau with the products re-associated (differences at the 1e-16 level).
// ---------------------------------------------------------------------------------------------------------------
constexpr int NT_LW = 24, NT_SW = 16;      // capacity of a term list (LW: multiple of 8, SW: of 4, see rec_dot)
struct alignas(16) LwRec { double w[NT_LW]; int off[NT_LW]; int n, gs_off, f0, f1; double ffs, pad_; };     // 320 bytes
struct alignas(16) SwRec { double w[NT_SW]; int off[NT_SW]; int n, r0, r1, rg, js, pad; double rc0, rc1, tconst, fs; };   // 248 -> 256 bytes

template <class Rec>
RR_HD void rec_add(Rec& r, int off, double w) { r.off[r.n] = off; r.w[r.n] = w; ++r.n; }

template <class Rec>
RR_HD void major3_terms(Rec& r, int K, int ng, int row, const Spec& s, double fa, double fb) {
  const double sc = s.speccomb;
  if (s.specparm < 0.125) {
    double p = s.fs - 1.0, p4 = p * p * p * p, fk0 = p4, fk1 = 1.0 - p - 2.0 * p4, fk2 = p + p4;
    rec_add(r, K + row * ng, sc * (fk0 * fa)); rec_add(r, K + (row + 1) * ng, sc * (fk1 * fa)); rec_add(r, K + (row + 2) * ng, sc * (fk2 * fa));
    rec_add(r, K + (row + 9) * ng, sc * (fk0 * fb)); rec_add(r, K + (row + 10) * ng, sc * (fk1 * fb)); rec_add(r, K + (row + 11) * ng, sc * (fk2 * fb));
  } else if (s.specparm > 0.875) {
    double p = -s.fs, p4 = p * p * p * p, fk0 = p4, fk1 = 1.0 - p - 2.0 * p4, fk2 = p + p4;
    rec_add(r, K + (row - 1) * ng, sc * (fk2 * fa)); rec_add(r, K + row * ng, sc * (fk1 * fa)); rec_add(r, K + (row + 1) * ng, sc * (fk0 * fa));
    rec_add(r, K + (row + 8) * ng, sc * (fk2 * fb)); rec_add(r, K + (row + 9) * ng, sc * (fk1 * fb)); rec_add(r, K + (row + 10) * ng, sc * (fk0 * fb));
  } else {
    rec_add(r, K + row * ng, sc * ((1.0 - s.fs) * fa)); rec_add(r, K + (row + 1) * ng, sc * (s.fs * fa));
    rec_add(r, K + (row + 9) * ng, sc * ((1.0 - s.fs) * fb)); rec_add(r, K + (row + 10) * ng, sc * (s.fs * fb));
  }
}
template <class Rec>
RR_HD void major2_terms(Rec& r, int K, int ng, int row, double sc, double fs, double fa, double fb, int stride) {
  rec_add(r, K + row * ng, sc * ((1.0 - fs) * fa)); rec_add(r, K + (row + 1) * ng, sc * (fs * fa));
  rec_add(r, K + (row + stride) * ng, sc * ((1.0 - fs) * fb)); rec_add(r, K + (row + stride + 1) * ng, sc * (fs * fb));
}
template <class Rec>
RR_HD void itab_terms(Rec& r, int T, int ng, int ind1, double frac, double amount) {   // amount * (T(ind) + frac*(T(ind+1)-T(ind)))
  rec_add(r, T + (ind1 - 1) * ng, amount * (1.0 - frac)); rec_add(r, T + ind1 * ng, amount * frac);
}

// the g-independent part of taugbNN for one (layer, band)
RR_HD void lw_terms(const double* A, const Tab& tb, const LwBand& B, const Layer& L, LwRec& rec) {
  const LwRegion& R = B.r[L.lower ? 0 : 1];
  const int ng = B.ng, low = L.lower;
  const double mult = low ? 8.0 : 4.0;
  rec.n = 0;
  double ca = 0.0, cb = 0.0;
  if (R.major == 1) {
    int i0 = low ? (L.jp - 1) * 5 + (L.jt - 1) : (L.jp - 13) * 5 + (L.jt - 1);
    int i1 = low ? L.jp * 5 + (L.jt1 - 1) : (L.jp - 12) * 5 + (L.jt1 - 1);
    const double c = L.col[R.spA];
    rec_add(rec, R.k_off + i0 * ng, c * L.fac00); rec_add(rec, R.k_off + (i0 + 1) * ng, c * L.fac10);
    rec_add(rec, R.k_off + i1 * ng, c * L.fac01); rec_add(rec, R.k_off + (i1 + 1) * ng, c * L.fac11);
  }
  if (R.spB >= 0) { ca = L.col[R.spA]; cb = L.col[R.spB]; }
  if (R.major == 2) {
    Spec s0 = specparm_of(ca, cb, chi_mls(A, tb, R.spA, L.jp) / chi_mls(A, tb, R.spB, L.jp), mult);
    Spec s1 = specparm_of(ca, cb, chi_mls(A, tb, R.spA, L.jp + 1) / chi_mls(A, tb, R.spB, L.jp + 1), mult);
    if (low) {
      int i0 = ((L.jp - 1) * 5 + (L.jt - 1)) * 9 + s0.js - 1;
      int i1 = (L.jp * 5 + (L.jt1 - 1)) * 9 + s1.js - 1;
      major3_terms(rec, R.k_off, ng, i0, s0, L.fac00, L.fac10);
      major3_terms(rec, R.k_off, ng, i1, s1, L.fac01, L.fac11);
    } else {
      int i0 = ((L.jp - 13) * 5 + (L.jt - 1)) * 5 + s0.js - 1;
      int i1 = ((L.jp - 12) * 5 + (L.jt1 - 1)) * 5 + s1.js - 1;
      major2_terms(rec, R.k_off, ng, i0, s0.speccomb, s0.fs, L.fac00, L.fac10, 5);
      major2_terms(rec, R.k_off, ng, i1, s1.speccomb, s1.fs, L.fac01, L.fac11, 5);
    }
  }
  if (R.self_off >= 0) itab_terms(rec, R.self_off, ng, L.indself, L.selffrac, L.selffac);
  if (R.for_off >= 0) itab_terms(rec, R.for_off, ng, L.indfor, L.forfrac, L.forfac);
  for (int m = 0; m < R.nminor; ++m) {
    const Minor& M = R.minor[m];
    double amount;
    if (M.scale == SC_COL) amount = L.col[M.sp];
    else if (M.scale == SC_ADJ) {
      double chiref = M.chiref > 0.0 ? M.chiref : chi_mls(A, tb, M.sp, L.jp + 1);
      double rat = 1.0e20 * (L.col[M.sp] / L.coldry) / chiref;
      amount = rat > M.thresh ? (M.base + pow(rat - M.base, M.expo)) * chiref * L.coldry * 1.0e-20 : L.col[M.sp];
    } else if (M.scale == SC_BRD_N2) amount = L.colbrd * L.scaleminorn2;
    else if (M.scale == SC_BRD) amount = L.colbrd * L.scaleminor;
    else amount = L.col[SP_O2] * L.scaleminor;
    if (M.binary) {
      Spec s = specparm_of(ca, cb, M.refrat, mult);
      int nsp = low ? 9 : 5;
      int r0 = (L.indminor - 1) * nsp + s.js - 1, r1 = L.indminor * nsp + s.js - 1;
      const double a0 = amount * (1.0 - L.minorfrac), a1 = amount * L.minorfrac;
      rec_add(rec, M.k_off + r0 * ng, a0 * (1.0 - s.fs)); rec_add(rec, M.k_off + (r0 + 1) * ng, a0 * s.fs);
      rec_add(rec, M.k_off + r1 * ng, a1 * (1.0 - s.fs)); rec_add(rec, M.k_off + (r1 + 1) * ng, a1 * s.fs);
    } else {
      itab_terms(rec, M.k_off, ng, L.indminor, L.minorfrac, amount);
    }
  }
  for (int c = 0; c < R.ncfc; ++c) rec_add(rec, R.cfc_off[c], L.wx[R.cfc_wx[c]]);
  double corr = 1.0;
  if (R.corr == 1) corr = (L.pavel < 250.0 ? 1.0 - 0.15 * (250.0 - L.pavel) / 154.4 : 1.0);
  else if (R.corr == 2) corr = 1.0 - 0.15 * (L.pavel / 95.6);
  else if (R.corr == 3) corr = 1.0 - 0.05 * (L.pavel - 100.0) / 900.0;
  if (corr != 1.0) for (int i = 0; i < rec.n; ++i) rec.w[i] *= corr;
  while (rec.n & 7) rec_add(rec, 0, 0.0);
  rec.gs_off = R.gscale_off;
  rec.f0 = rec.f1 = -1; rec.ffs = 0.0;
  if (R.frac_off >= 0) {
    if (R.frac2d) {
      Spec s = specparm_of(ca, cb, R.refrat_planck, mult);
      rec.f0 = R.frac_off + (s.js - 1) * ng; rec.f1 = R.frac_off + s.js * ng; rec.ffs = s.fs;
    } else rec.f0 = R.frac_off;
  }
}
// the g-dependent part: a dot product over the term list
// W independent table reads in flight per step (the reads mostly hit in L2: their latency, not their number, is what costs); the lists
// are padded with zero-weight terms to whole groups of W (8 for the longwave lists of up to 24 terms, 4 for the shorter shortwave ones)
template <int W, class Rec>
RR_HD double rec_dot(const double* A, const Rec& r, int g) {
  const double* Ag = A + g;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  for (int i = 0; i < r.n; i += W) {
    double a[W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = Ag[r.off[i + j]];
    t0 += r.w[i] * a[0]; t1 += r.w[i + 1] * a[1]; t2 += r.w[i + 2] * a[2]; t3 += r.w[i + 3] * a[3];
    if (W == 8) { t0 += r.w[i + 4] * a[4]; t1 += r.w[i + 5] * a[5]; t2 += r.w[i + 6] * a[6]; t3 += r.w[i + 7] * a[7]; }
  }
  return (t0 + t1) + (t2 + t3);
}
RR_HD void lw_tau_rec(const double* A, const LwRec& r, int g, double& tau, double& frac) {
  double t = rec_dot<8>(A, r, g);
  if (r.gs_off >= 0) t *= A[r.gs_off + g];
  tau = t;
  if (r.f0 < 0) frac = 0.0;
  else {
    double a = A[r.f0 + g];
    frac = r.f1 >= 0 ? a + r.ffs * (A[r.f1 + g] - a) : a;
  }
}

// Two consecutive doubles of a table with one 128-bit read (every table of the arena starts on a 16-byte boundary and the g-point
// counts are even).  ISCA_TAB2_MODE 2 marks the read evict-last in L1 -- measured harmful (the thread-local transmittance arrays of
// the sweeps lose their L1 lines: DRAM write traffic doubles), kept for the record; 1 = plain 128-bit read; 0 = two scalar reads.
#ifndef ISCA_TAB2_MODE
#define ISCA_TAB2_MODE 0
#endif
RR_HD void tab2(const double* p, double& a, double& b) {
#if defined(__CUDA_ARCH__) && ISCA_TAB2_MODE == 2
  asm("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(__cvta_generic_to_global(p)));
#elif defined(__CUDA_ARCH__) && ISCA_TAB2_MODE == 1
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(__cvta_generic_to_global(p)));
#else
  a = p[0]; b = p[1];
#endif
}

// one layer of rtrnmr's clear-sky downward sweep for one g-point (rrtmg_lw_rtrnmr.f90:390-480): transmittance from the Pade
// table (exptfn = interleaved exp / tfn tables), layer emission towards both sides, update of the downward radiance
RR_HD void lw_layer(const double* exptfn, double secdiff, double tau, double plfrac, double blay, double plev_up, double plev_dn,
                    double& radld, double& atrans, double& bbugas) {
  double dplankup = plev_up - blay;
  double dplankdn = plev_dn - blay;
  double odepth = secdiff * tau;
  if (odepth < 0.0) odepth = 0.0;
  double at, tfac;
  if (odepth <= 0.06) {
    at = odepth - 0.5 * odepth * odepth;
    tfac = 0.166667 * odepth;
  } else {
    double tblind = odepth / (BPADE + odepth);
    int itr = (int)(TBLINT * tblind + 0.5);
    double e;
    tab2(exptfn + 2 * itr, e, tfac);
    at = 1.0 - e;
  }
  double bbd = plfrac * (blay + tfac * dplankdn);
  bbugas = plfrac * (blay + tfac * dplankup);
  atrans = at;
  radld = radld + (bbd - radld) * at;
}

// One g-point through rtrnmr's clear-sky sweeps, serially (rrtmg_lw_rtrnmr.f90:390-480 down, :560-640 up): the order of
// operations of rrtmg_lw_kernel's phase B written as one loop nest, on the term lists `recs` of the band's layers.  Used by the
// host build (tests/host/rrtm_host.cpp: CPU check of this arithmetic, and the C++/OpenMP CPU baseline of bench.py).
// lay[] = the column's setcoef output, planklay/planklev [16][stride], w = wtdiff*delwave(band).  red.down(lev, v) / red.up(lev, v)
// receive the weighted radiances of level lev = 0..nl (0 = surface).
template <class Red>
RR_HD void lw_gpoint_recs(const double* A, const Tab& tb, const LwRec* recs, int ib, int g, int nl,
                          const double* planklay, const double* planklev, int pstride, double plankbnd, double semiss,
                          double secdiff, double w, Red& red) {
  double atrans[KMAX], bbugas[KMAX];
  const double* exptfn = A + tb.exptfn;
  double radld = 0.0, frac1 = 0.0;
  red.down(nl, 0.0);
  for (int lev = nl; lev >= 1; --lev) {
    double tau, plfrac;
    lw_tau_rec(A, recs[lev - 1], g, tau, plfrac);
    if (lev == 1) frac1 = plfrac;
    lw_layer(exptfn, secdiff, tau, plfrac, planklay[ib * pstride + lev - 1], planklev[ib * pstride + lev], planklev[ib * pstride + lev - 1],
             radld, atrans[lev - 1], bbugas[lev - 1]);
    red.down(lev - 1, radld * w);
  }
  double rad0 = frac1 * plankbnd;
  double radlu = rad0 + (1.0 - semiss) * radld;
  red.up(0, radlu * w);
  for (int lev = 1; lev <= nl; ++lev) {
    radlu = radlu + (bbugas[lev - 1] - radlu) * atrans[lev - 1];
    red.up(lev, radlu * w);
  }
}
// the term lists of one band for all layers of a column (built once per band, shared by the band's g-points)
RR_HD void lw_band_recs(const double* A, const Tab& tb, const LwBand& B, int nl, const Layer* lay, LwRec* recs) {
  for (int l = 0; l < nl; ++l) lw_terms(A, tb, B, lay[l], recs[l]);
}

// secant of the diffusivity angle per band (rrtmg_lw_rtrnmr.f90:262-272)
RR_HD double lw_secdiff(int ib, double pwvcm) {
  const double a0[9] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668};
  const double a1[9] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006};
  const double a2[9] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414};
  if (ib == 0 || ib == 3 || ib >= 9) return 1.66;
  double s = a0[ib] + a1[ib] * exp(a2[ib] * pwvcm);
  return s > 1.80 ? 1.80 : (s < 1.50 ? 1.50 : s);
}

// ---------------------------------------------------------------------------------------------------------------
// shortwave
// ---------------------------------------------------------------------------------------------------------------
// the g-independent part of taumolNN for one (layer, band) (see the term lists above)
RR_HD void sw_terms(const double* A, const SwBand& B, const Layer& L, SwRec& rec) {
  const SwRegion& R = B.r[L.lower ? 0 : 1];
  const int ng = B.ng, low = L.lower;
  rec.n = 0; rec.js = 1; rec.fs = 0.0; rec.pad = 0;
  if (R.major == 1) {
    int i0 = low ? (L.jp - 1) * 5 + (L.jt - 1) : (L.jp - 13) * 5 + (L.jt - 1);
    int i1 = low ? L.jp * 5 + (L.jt1 - 1) : (L.jp - 12) * 5 + (L.jt1 - 1);
    const double c = L.col[R.spA] * R.kscale;
    rec_add(rec, R.k_off + i0 * ng, c * L.fac00); rec_add(rec, R.k_off + (i0 + 1) * ng, c * L.fac10);
    rec_add(rec, R.k_off + i1 * ng, c * L.fac01); rec_add(rec, R.k_off + (i1 + 1) * ng, c * L.fac11);
  } else if (R.major == 2) {
    const int nsp = low ? 9 : 5;
    Spec s = specparm_of(L.col[R.spA], L.col[R.spB], R.strrat, low ? 8.0 : 4.0);
    rec.js = s.js; rec.fs = s.fs;
    int i0 = low ? ((L.jp - 1) * 5 + (L.jt - 1)) * nsp + s.js - 1 : ((L.jp - 13) * 5 + (L.jt - 1)) * nsp + s.js - 1;
    int i1 = low ? (L.jp * 5 + (L.jt1 - 1)) * nsp + s.js - 1 : ((L.jp - 12) * 5 + (L.jt1 - 1)) * nsp + s.js - 1;
    major2_terms(rec, R.k_off, ng, i0, s.speccomb, s.fs, L.fac00, L.fac10, nsp);
    major2_terms(rec, R.k_off, ng, i1, s.speccomb, s.fs, L.fac01, L.fac11, nsp);
  }
  if (R.self_off >= 0) itab_terms(rec, R.self_off, ng, L.indself, L.selffrac, L.col[SP_H2O] * L.selffac);
  if (R.for_off >= 0) itab_terms(rec, R.for_off, ng, L.indfor, L.forfrac, L.col[SP_H2O] * L.forfac);
  for (int e = 0; e < R.nextra; ++e) rec_add(rec, R.extra_off[e], L.col[R.extra_sp[e]]);
  while (rec.n & 3) rec_add(rec, 0, 0.0);
  rec.tconst = R.o2cont ? 4.35e-4 * L.col[SP_O2] / (350.0 * 2.0) : 0.0;
  rec.r1 = -1; rec.rc1 = 0.0;
  if (R.rayl_mode == 0) { rec.r0 = R.rayl_off; rec.rg = 0; rec.rc0 = L.colmol; }
  else if (R.rayl_mode == 1) { rec.r0 = R.rayl_off; rec.rg = 1; rec.rc0 = L.colmol; }
  else { rec.r0 = R.rayl_off + (rec.js - 1) * ng; rec.r1 = R.rayl_off + rec.js * ng; rec.rg = 1; rec.rc0 = L.colmol; rec.rc1 = rec.fs; }
}
RR_HD void sw_tau_rec(const double* A, const SwRec& r, int g, double& taug, double& taur) {
  taug = rec_dot<4>(A, r, g) + r.tconst;
  double a = A[r.r0 + r.rg * g];
  taur = r.rc0 * (r.r1 >= 0 ? a + r.rc1 * (A[r.r1 + g] - a) : a);
}
RR_HD double sw_src_rec(const double* A, const SwBand& B, const SwRec& r, int g) {       // solar source term of the band's laysolfr layer
  if (B.sflux2d) {
    const double* F = A + B.sflux_off;
    return F[(r.js - 1) * B.ng + g] + r.fs * (F[r.js * B.ng + g] - F[(r.js - 1) * B.ng + g]);
  }
  return B.sflux_scale * A[B.sflux_off + g];
}

// sw_laysolfr from the pressure indices alone (the column-per-lane kernel keeps jp of every layer, not the whole setcoef output)
RR_HD int sw_laysolfr_jp(const SwBand& B, const int* jp, int nl, int laytrop) {
  if (B.sflux_upper) {
    int ls = nl;
    for (int l = (laytrop + 1 > 2 ? laytrop + 1 : 2); l <= nl; ++l)
      if (jp[l - 2] < B.layreffr && jp[l - 1] >= B.layreffr) ls = l;
    return ls > laytrop ? ls : 0;
  }
  int ls = laytrop;
  for (int l = 1; l <= laytrop; ++l) {
    int jpn = l < nl ? jp[l] : 0;
    if (jp[l - 1] < B.layreffr && jpn >= B.layreffr) ls = (l + 1 < laytrop ? l + 1 : laytrop);
  }
  return ls;
}

// the layer (1-based) whose species ratio defines the band's solar source: the `laysolfr` logic of taumol16..29
RR_HD int sw_laysolfr(const SwBand& B, const Layer* lay, int nl, int laytrop) {
  if (B.sflux_upper) {
    int ls = nl;
    for (int l = (laytrop + 1 > 2 ? laytrop + 1 : 2); l <= nl; ++l)
      if (lay[l - 2].jp < B.layreffr && lay[l - 1].jp >= B.layreffr) ls = l;
    return ls > laytrop ? ls : 0;          // the source is only set inside the upper-atmosphere loop
  }
  int ls = laytrop;
  for (int l = 1; l <= laytrop; ++l) {
    int jpn = l < nl ? lay[l].jp : 0;
    if (lay[l - 1].jp < B.layreffr && jpn >= B.layreffr) ls = (l + 1 < laytrop ? l + 1 : laytrop);
  }
  return ls;
}

RR_HD double sw_exp(const double* exp_tbl, double ze) {            // exp(-ze) of spcvrt_sw / reftra_sw
  if (ze <= 0.06) return 1.0 - ze + 0.5 * ze * ze;
  double tblind = ze / (BPADE + ze);
  int itind = (int)(TBLINT * tblind + 0.5);
  return exp_tbl[itind];
}

// reftra_sw for one layer (kmodts = 2) with asymmetry factor g = 0 (clear sky without aerosols: Rayleigh scattering only).  These are
// the reference's expressions with zg = 0 substituted -- zg3 = 0, zgamma3 = zgamma4 = 1/2, zwo = zw / (1 - (1 - zw) * 0) = zw -- which
// leaves every retained operation and its operands unchanged (bit-identical results), but spares a division and the dead products.
// ze_dir = zto1 / prmuz (the caller needs it for the direct beam as well).
RR_HD void sw_reftra(const double* exp_tbl, double prmuz, double ze_dir, double zto1, double zw, double& pref, double& prefd,
                     double& ptra, double& ptrad) {
  const double zwcrit = 0.9999995, eps = 1.0e-08;
  const double zgamma1 = (8.0 - zw * 5.0) * 0.25;
  const double zgamma2 = 3.0 * zw * 0.25;
  const double zgamma3 = 0.5, zgamma4 = 0.5;
  if (zw >= zwcrit) {
    double za = zgamma1 * prmuz, za1 = za - zgamma3, zgt = zgamma1 * zto1;
    double ze1 = ze_dir; if (ze1 > 500.0) ze1 = 500.0;
    double ze2 = sw_exp(exp_tbl, ze1);
    pref = (zgt - za1 * (1.0 - ze2)) / (1.0 + zgt);
    ptra = 1.0 - pref;
    prefd = zgt / (1.0 + zgt);
    ptrad = 1.0 - prefd;
    if (ze2 == 1.0) { pref = 0.0; ptra = 1.0; prefd = 0.0; ptrad = 1.0; }
  } else {
    double za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
    double za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
    double zrk = sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
    double zrp = zrk * prmuz, zrp1 = 1.0 + zrp, zrm1 = 1.0 - zrp, zrk2 = 2.0 * zrk, zrpp = 1.0 - zrp * zrp;
    double zrkg = zrk + zgamma1;
    double zr1 = zrm1 * (za2 + zrk * zgamma3), zr2 = zrp1 * (za2 - zrk * zgamma3), zr3 = zrk2 * (zgamma3 - za2 * prmuz);
    double zr4 = zrpp * zrkg, zr5 = zrpp * (zrk - zgamma1);
    double zt1 = zrp1 * (za1 + zrk * zgamma4), zt2 = zrm1 * (za1 - zrk * zgamma4), zt3 = zrk2 * (zgamma4 + za1 * prmuz);
    double zt4 = zr4, zt5 = zr5;
    double zbeta = (zgamma1 - zrk) / zrkg;
    double ze1 = zrk * zto1; if (ze1 > 500.0) ze1 = 500.0;
    double ze2 = ze_dir; if (ze2 > 500.0) ze2 = 500.0;
    double zem1 = sw_exp(exp_tbl, ze1), zep1 = 1.0 / zem1;
    double zem2 = sw_exp(exp_tbl, ze2), zep2 = 1.0 / zem2;
    double zdenr = zr4 * zep1 + zr5 * zem1, zdent = zt4 * zep1 + zt5 * zem1;
    if (zdenr >= -eps && zdenr <= eps) { pref = eps; ptra = zem2; }
    else {
      pref = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
      ptra = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
    }
    double zemm = zem1 * zem1;
    double zdend = 1.0 / ((1.0 - zbeta * zemm) * zrkg);
    prefd = zgamma2 * (1.0 - zemm) * zdend;
    ptrad = zrk2 * zem1 * zdend;
  }
}

// one layer of spcvrt_sw for one g-point (clear sky, no aerosol: asymmetry 0): two-stream reflectance / transmittance (reftra_sw),
// direct-beam transmittance, and vrtqdr_sw's upward combination with everything below (rup_below, rupd_below)
RR_HD void sw_layer(const double* exp_tbl, double prmu0, double taug, double taur, double rup_below, double rupd_below,
                    double& zref, double& zrefd, double& ztra, double& ztrad, double& zdbt, double& zrup, double& zrupd) {
  double ztauc = taur + taug;               // + aerosol (none)
  double zomcc = taur / ztauc;              // single-scattering albedo
  const double ze_dir = ztauc / prmu0;      // optical path of the direct beam: reftra's ze1 / ze2 (clamped there) and the argument of zdbt
  sw_reftra(exp_tbl, prmu0, ze_dir, ztauc, zomcc, zref, zrefd, ztra, ztrad);
  zdbt = sw_exp(exp_tbl, ze_dir);
  double zreflect = 1.0 / (1.0 - rupd_below * zrefd);
  zrup = zref + (ztrad * ((ztra - zdbt) * rupd_below + zdbt * rup_below)) * zreflect;
  zrupd = zrefd + ztrad * ztrad * rupd_below * zreflect;
}
// one level of vrtqdr_sw's downward combination + the fluxes of spcvrt_sw at that level (jk = 0 top .. nl surface); tdbt = ztdbt(jk),
// tdbt_prev = ztdbt(jk-1); layer values are those of layer j = jk-1
RR_HD void sw_level(int jk, double tdbt, double tdbt_prev, double zref_j, double zrefd_j, double ztra_j, double ztrad_j, double zrup, double zrupd,
                    double& ztdn, double& zrdnd, double& fu, double& fd) {
  if (jk == 1) { ztdn = ztra_j; zrdnd = zrefd_j; }
  else if (jk >= 2) {
    double zr = 1.0 / (1.0 - zrefd_j * zrdnd);
    double tdn_new = tdbt_prev * ztra_j + (ztrad_j * ((ztdn - tdbt_prev) + tdbt_prev * zref_j * zrdnd)) * zr;
    double rdnd_new = zrefd_j + ztrad_j * ztrad_j * zrdnd * zr;
    ztdn = tdn_new; zrdnd = rdnd_new;
  }
  double zreflect = 1.0 / (1.0 - zrdnd * zrupd);
  fu = (tdbt * zrup + (ztdn - tdbt) * zrupd) * zreflect;
  fd = tdbt + (ztdn - tdbt + tdbt * zrup * zrdnd) * zreflect;
}

// One g-point through spcvrt_sw + vrtqdr_sw, serially: the order of operations of rrtmg_sw_kernel written as one loop nest.  Used by
// the test-only host build.  lsol = laysolfr of the band; incoming flux zincflx = adjflux * sfluxzen * prmu0.
// red.up(lev, v) / red.down(lev, v): lev = 0 surface .. nl top of atmosphere.
template <class Red>
RR_HD void sw_gpoint_recs(const double* A, const Tab& tb, const SwBand& B, const SwRec* recs, int g, int nl, int lsol,
                          double prmu0, double albedo, double adjflux, double w, Red& red) {
  const double* exp_tbl = A + tb.exp_tbl;
  // layer arrays ordered top (0) to bottom (nl-1), as jk = 1..klev of spcvrt_sw
  double zref[KMAX], zrefd[KMAX], ztra[KMAX], ztrad[KMAX], zdbt[KMAX], zrup[KMAX + 1], zrupd[KMAX + 1];
  double sflux = 0.0;
  zrup[nl] = albedo; zrupd[nl] = albedo;
  for (int l1 = 1; l1 <= nl; ++l1) {          // bottom-up: layer optics + upward combination
    const int jk = nl - l1;
    const SwRec& rec = recs[l1 - 1];
    double taug, taur;
    sw_tau_rec(A, rec, g, taug, taur);
    if (l1 == lsol) sflux = sw_src_rec(A, B, rec, g);
    sw_layer(exp_tbl, prmu0, taug, taur, zrup[jk + 1], zrupd[jk + 1], zref[jk], zrefd[jk], ztra[jk], ztrad[jk], zdbt[jk], zrup[jk], zrupd[jk]);
  }
  const double zinc = adjflux * sflux * prmu0 * w;
  double ztdn = 1.0, zrdnd = 0.0, tdbt = 1.0, tdbt_prev = 1.0;
  for (int jk = 0; jk <= nl; ++jk) {          // top-down: downward combination + fluxes
    if (jk >= 1) { tdbt_prev = tdbt; tdbt = zdbt[jk - 1] * tdbt_prev; }
    const int j = jk >= 1 ? jk - 1 : 0;
    double fu, fd;
    sw_level(jk, tdbt, tdbt_prev, zref[j], zrefd[j], ztra[j], ztrad[j], zrup[jk], zrupd[jk], ztdn, zrdnd, fu, fd);
    red.up(nl - jk, zinc * fu);
    red.down(nl - jk, zinc * fd);
  }
}
RR_HD void sw_band_recs(const double* A, const SwBand& B, int nl, const Layer* lay, SwRec* recs) {
  for (int l = 0; l < nl; ++l) sw_terms(A, B, lay[l], recs[l]);
}

}  // namespace rrtm
