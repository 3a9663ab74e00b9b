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
  int preflog, tref, chi, totplnk, exp_tbl, tfn_tbl, sw_preflog, sw_tref;
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
  s.fs = fmod(specmult, 1.0);
  return s;
}

// three-branch species-ratio interpolation of the lower-atmosphere binary bands (e.g. taugb3 :370-470); row = 0-based
// row of absa(ind0, :)
RR_HD double major3(const double* K, int ng, int g, int row, const Spec& s, double fa, double fb) {
  double r;
  if (s.specparm < 0.125) {
    double p = s.fs - 1.0, p4 = p * p * p * p, fk0 = p4, fk1 = 1.0 - p - 2.0 * p4, fk2 = p + p4;
    r = fk0 * fa * K[row * ng + g] + fk1 * fa * K[(row + 1) * ng + g] + fk2 * fa * K[(row + 2) * ng + g]
      + fk0 * fb * K[(row + 9) * ng + g] + fk1 * fb * K[(row + 10) * ng + g] + fk2 * fb * K[(row + 11) * ng + g];
  } else if (s.specparm > 0.875) {
    double p = -s.fs, p4 = p * p * p * p, fk0 = p4, fk1 = 1.0 - p - 2.0 * p4, fk2 = p + p4;
    r = fk2 * fa * K[(row - 1) * ng + g] + fk1 * fa * K[row * ng + g] + fk0 * fa * K[(row + 1) * ng + g]
      + fk2 * fb * K[(row + 8) * ng + g] + fk1 * fb * K[(row + 9) * ng + g] + fk0 * fb * K[(row + 10) * ng + g];
  } else {
    r = (1.0 - s.fs) * fa * K[row * ng + g] + s.fs * fa * K[(row + 1) * ng + g]
      + (1.0 - s.fs) * fb * K[(row + 9) * ng + g] + s.fs * fb * K[(row + 10) * ng + g];
  }
  return s.speccomb * r;
}
RR_HD double major2(const double* K, int ng, int g, int row, double fs, double fa, double fb, int stride) {
  return (1.0 - fs) * fa * K[row * ng + g] + fs * fa * K[(row + 1) * ng + g]
       + (1.0 - fs) * fb * K[(row + stride) * ng + g] + fs * fb * K[(row + stride + 1) * ng + g];
}
RR_HD double itab(const double* T, int ng, int g, int ind1, double frac) {     // T(ind,g) + frac*(T(ind+1,g) - T(ind,g))
  double a = T[(ind1 - 1) * ng + g];
  return a + frac * (T[ind1 * ng + g] - a);
}

// gaseous optical depth and Planck fraction of one (layer, g-point): taugbNN of rrtmg_lw_taumol.f90
RR_HD void lw_tau(const double* A, const Tab& tb, const LwBand& B, const Layer& L, int g, double& tau, double& frac) {
  const LwRegion& R = B.r[L.lower ? 0 : 1];
  const int ng = B.ng, low = L.lower;
  const double mult = low ? 8.0 : 4.0;
  double t = 0.0;
  double ca = 0.0, cb = 0.0;
  if (R.major == 1) {
    const double* K = A + R.k_off;
    int i0 = low ? (L.jp - 1) * 5 + (L.jt - 1) : (L.jp - 13) * 5 + (L.jt - 1);
    int i1 = low ? L.jp * 5 + (L.jt1 - 1) : (L.jp - 12) * 5 + (L.jt1 - 1);
    t = L.col[R.spA] * (L.fac00 * K[i0 * ng + g] + L.fac10 * K[(i0 + 1) * ng + g]
                        + L.fac01 * K[i1 * ng + g] + L.fac11 * K[(i1 + 1) * ng + g]);
  }
  if (R.spB >= 0) { ca = L.col[R.spA]; cb = L.col[R.spB]; }
  if (R.major == 2) {
    const double* K = A + R.k_off;
    Spec s0 = specparm_of(ca, cb, chi_mls(A, tb, R.spA, L.jp) / chi_mls(A, tb, R.spB, L.jp), mult);
    Spec s1 = specparm_of(ca, cb, chi_mls(A, tb, R.spA, L.jp + 1) / chi_mls(A, tb, R.spB, L.jp + 1), mult);
    if (low) {
      int i0 = ((L.jp - 1) * 5 + (L.jt - 1)) * 9 + s0.js - 1;
      int i1 = (L.jp * 5 + (L.jt1 - 1)) * 9 + s1.js - 1;
      t = major3(K, ng, g, i0, s0, L.fac00, L.fac10) + major3(K, ng, g, i1, s1, L.fac01, L.fac11);
    } else {
      int i0 = ((L.jp - 13) * 5 + (L.jt - 1)) * 5 + s0.js - 1;
      int i1 = ((L.jp - 12) * 5 + (L.jt1 - 1)) * 5 + s1.js - 1;
      t = s0.speccomb * major2(K, ng, g, i0, s0.fs, L.fac00, L.fac10, 5)
        + s1.speccomb * major2(K, ng, g, i1, s1.fs, L.fac01, L.fac11, 5);
    }
  }
  if (R.self_off >= 0) t += L.selffac * itab(A + R.self_off, ng, g, L.indself, L.selffrac);
  if (R.for_off >= 0) t += L.forfac * itab(A + R.for_off, ng, g, L.indfor, L.forfrac);
  for (int m = 0; m < R.nminor; ++m) {
    const Minor& M = R.minor[m];
    const double* K = A + M.k_off;
    double ab;
    if (M.binary) {
      Spec s = specparm_of(ca, cb, M.refrat, mult);
      int nsp = low ? 9 : 5;
      // k(js, indm, g): row = (indm-1)*nsp + js-1
      int r0 = (L.indminor - 1) * nsp + s.js - 1, r1 = L.indminor * nsp + s.js - 1;
      double m1 = K[r0 * ng + g] + s.fs * (K[(r0 + 1) * ng + g] - K[r0 * ng + g]);
      double m2 = K[r1 * ng + g] + s.fs * (K[(r1 + 1) * ng + g] - K[r1 * ng + g]);
      ab = m1 + L.minorfrac * (m2 - m1);
    } else {
      ab = itab(K, ng, g, L.indminor, L.minorfrac);
    }
    double amount;
    if (M.scale == SC_COL) amount = L.col[M.sp];
    else if (M.scale == SC_ADJ) {
      double chiref = M.chiref > 0.0 ? M.chiref : chi_mls(A, tb, M.sp, L.jp + 1);
      double rat = 1.0e20 * (L.col[M.sp] / L.coldry) / chiref;
      amount = rat > M.thresh ? (M.base + pow(rat - M.base, M.expo)) * chiref * L.coldry * 1.0e-20 : L.col[M.sp];
    } else if (M.scale == SC_BRD_N2) amount = L.colbrd * L.scaleminorn2;
    else if (M.scale == SC_BRD) amount = L.colbrd * L.scaleminor;
    else amount = L.col[SP_O2] * L.scaleminor;
    t += amount * ab;
  }
  for (int c = 0; c < R.ncfc; ++c) t += L.wx[R.cfc_wx[c]] * A[R.cfc_off[c] + g];
  if (R.corr == 1) t *= (L.pavel < 250.0 ? 1.0 - 0.15 * (250.0 - L.pavel) / 154.4 : 1.0);
  else if (R.corr == 2) t *= 1.0 - 0.15 * (L.pavel / 95.6);
  else if (R.corr == 3) t *= 1.0 - 0.05 * (L.pavel - 100.0) / 900.0;
  if (R.gscale_off >= 0) t *= A[R.gscale_off + g];
  tau = t;
  if (R.frac_off < 0) frac = 0.0;
  else if (R.frac2d) {
    Spec s = specparm_of(ca, cb, R.refrat_planck, mult);
    const double* F = A + R.frac_off;
    frac = F[(s.js - 1) * ng + g] + s.fs * (F[s.js * ng + g] - F[(s.js - 1) * ng + g]);
  } else frac = A[R.frac_off + g];
}

// One g-point through rtrnmr's clear-sky sweeps (rrtmg_lw_rtrnmr.f90:390-480 down, :560-640 up).  lay[] = the column's
// setcoef output, planklay/planklev [16][stride], w = wtdiff*delwave(band) (0 for padding lanes).  red.down(lev, v) /
// red.up(lev, v) receive the weighted radiances of level lev = 0..nl (0 = surface).
template <class Red>
RR_HD void lw_gpoint(const double* A, const Tab& tb, const LwBand& B, int ib, int g, int nl, const Layer* lay,
                     const double* planklay, const double* planklev, int pstride, double plankbnd, double semiss,
                     double secdiff, double w, Red& red) {
  double atrans[KMAX], bbugas[KMAX];
  const double* exp_tbl = A + tb.exp_tbl;
  const double* tfn_tbl = A + tb.tfn_tbl;
  double radld = 0.0, frac1 = 0.0;
  red.down(nl, 0.0);
  for (int lev = nl; lev >= 1; --lev) {
    double tau, plfrac;
    lw_tau(A, tb, B, lay[lev - 1], g, tau, plfrac);
    if (lev == 1) frac1 = plfrac;
    double blay = planklay[ib * pstride + lev - 1];
    double dplankup = planklev[ib * pstride + lev] - blay;
    double dplankdn = planklev[ib * pstride + lev - 1] - blay;
    double odepth = secdiff * tau;
    if (odepth < 0.0) odepth = 0.0;
    double at, tfac;
    if (odepth <= 0.06) {
      at = odepth - 0.5 * odepth * odepth;
      tfac = 0.166667 * odepth;
    } else {
      double tblind = odepth / (BPADE + odepth);
      int itr = (int)(TBLINT * tblind + 0.5);
      at = 1.0 - exp_tbl[itr];
      tfac = tfn_tbl[itr];
    }
    double bbd = plfrac * (blay + tfac * dplankdn);
    bbugas[lev - 1] = plfrac * (blay + tfac * dplankup);
    atrans[lev - 1] = at;
    radld = radld + (bbd - radld) * at;
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
struct SwOptics { double taug, taur, src; };

// taumolNN of rrtmg_sw_taumol.f90 for one (layer, g-point); src = solar source if this layer is the band's laysolfr
RR_HD void sw_tau(const double* A, const SwBand& B, const Layer& L, int g, double& taug, double& taur, double& src) {
  const SwRegion& R = B.r[L.lower ? 0 : 1];
  const int ng = B.ng, low = L.lower;
  double t = 0.0;
  Spec s; s.js = 1; s.fs = 0.0; s.speccomb = 0.0; s.specparm = 0.0;
  if (R.major == 1) {
    const double* K = A + R.k_off;
    int i0 = low ? (L.jp - 1) * 5 + (L.jt - 1) : (L.jp - 13) * 5 + (L.jt - 1);
    int i1 = low ? L.jp * 5 + (L.jt1 - 1) : (L.jp - 12) * 5 + (L.jt1 - 1);
    t = L.col[R.spA] * R.kscale * (L.fac00 * K[i0 * ng + g] + L.fac10 * K[(i0 + 1) * ng + g]
                                   + L.fac01 * K[i1 * ng + g] + L.fac11 * K[(i1 + 1) * ng + g]);
  } else if (R.major == 2) {
    const double* K = A + R.k_off;
    const int nsp = low ? 9 : 5;
    s = specparm_of(L.col[R.spA], L.col[R.spB], R.strrat, low ? 8.0 : 4.0);
    int i0 = low ? ((L.jp - 1) * 5 + (L.jt - 1)) * nsp + s.js - 1 : ((L.jp - 13) * 5 + (L.jt - 1)) * nsp + s.js - 1;
    int i1 = low ? (L.jp * 5 + (L.jt1 - 1)) * nsp + s.js - 1 : ((L.jp - 12) * 5 + (L.jt1 - 1)) * nsp + s.js - 1;
    t = s.speccomb * (major2(K, ng, g, i0, s.fs, L.fac00, L.fac10, nsp) + major2(K, ng, g, i1, s.fs, L.fac01, L.fac11, nsp));
  }
  double cont = 0.0;
  if (R.self_off >= 0) cont += L.selffac * itab(A + R.self_off, ng, g, L.indself, L.selffrac);
  if (R.for_off >= 0) cont += L.forfac * itab(A + R.for_off, ng, g, L.indfor, L.forfrac);
  t += L.col[SP_H2O] * cont;
  for (int e = 0; e < R.nextra; ++e) t += L.col[R.extra_sp[e]] * A[R.extra_off[e] + g];
  if (R.o2cont) t += 4.35e-4 * L.col[SP_O2] / (350.0 * 2.0);
  taug = t;
  if (R.rayl_mode == 0) taur = L.colmol * A[R.rayl_off];
  else if (R.rayl_mode == 1) taur = L.colmol * A[R.rayl_off + g];
  else {
    const double* Ra = A + R.rayl_off;
    taur = L.colmol * (Ra[(s.js - 1) * ng + g] + s.fs * (Ra[s.js * ng + g] - Ra[(s.js - 1) * ng + g]));
  }
  if (B.sflux2d) {
    const double* F = A + B.sflux_off;
    src = F[(s.js - 1) * ng + g] + s.fs * (F[s.js * ng + g] - F[(s.js - 1) * ng + g]);
  } else src = B.sflux_scale * A[B.sflux_off + g];
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

// reftra_sw for one layer (kmodts = 2)
RR_HD void sw_reftra(const double* exp_tbl, double zg, double prmuz, double zto1, double zw, double& pref, double& prefd,
                     double& ptra, double& ptrad) {
  const double zwcrit = 0.9999995, eps = 1.0e-08;
  double zg3 = 3.0 * zg;
  double zgamma1 = (8.0 - zw * (5.0 + zg3)) * 0.25;
  double zgamma2 = 3.0 * (zw * (1.0 - zg)) * 0.25;
  double zgamma3 = (2.0 - zg3 * prmuz) * 0.25;
  double zgamma4 = 1.0 - zgamma3;
  double q = zg / (1.0 - zg);
  double zwo = zw / (1.0 - (1.0 - zw) * q * q);
  if (zwo >= zwcrit) {
    double za = zgamma1 * prmuz, za1 = za - zgamma3, zgt = zgamma1 * zto1;
    double ze1 = zto1 / prmuz; if (ze1 > 500.0) ze1 = 500.0;
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
    double ze2 = zto1 / prmuz; if (ze2 > 500.0) ze2 = 500.0;
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

// One g-point through spcvrt_sw (clear sky, no aerosol) + vrtqdr_sw.  lsol[ib] = laysolfr of the band; incoming flux
// zincflx = adjflux * sfluxzen * prmu0.  red.up(lev, v) / red.down(lev, v): lev = 0 surface .. nl top of atmosphere.
template <class Red>
RR_HD void sw_gpoint(const double* A, const Tab& tb, const SwBand& B, int g, int nl, const Layer* lay, int lsol,
                     double prmu0, double albedo, double adjflux, double w, Red& red) {
  const double* exp_tbl = A + tb.exp_tbl;
  // layer arrays ordered top (0) to bottom (nl-1), as jk = 1..klev of spcvrt_sw
  double zref[KMAX + 1], zrefd[KMAX + 1], ztra[KMAX], ztrad[KMAX], zdbt[KMAX], ztdbt[KMAX + 1], zrup[KMAX + 1], zrupd[KMAX + 1];
  double sflux = 0.0;
  ztdbt[0] = 1.0;
  for (int jk = 0; jk < nl; ++jk) {
    int ikl = nl - 1 - jk;
    double taug, taur, src;
    sw_tau(A, B, lay[ikl], g, taug, taur, src);
    if (ikl + 1 == lsol) sflux = src;
    double ztauc = taur + taug;               // + aerosol (none)
    double zomcc = taur / ztauc;              // single-scattering albedo; asymmetry 0 (Rayleigh only)
    sw_reftra(exp_tbl, 0.0, prmu0, ztauc, zomcc, zref[jk], zrefd[jk], ztra[jk], ztrad[jk]);
    zdbt[jk] = sw_exp(exp_tbl, ztauc / prmu0);
    ztdbt[jk + 1] = zdbt[jk] * ztdbt[jk];
  }
  zref[nl] = albedo; zrefd[nl] = albedo; zrup[nl] = albedo; zrupd[nl] = albedo;
  // vrtqdr_sw: bottom-up combined reflectances
  {
    int k = nl - 1;
    double zreflect = 1.0 / (1.0 - zrefd[nl] * zrefd[k]);
    zrup[k] = zref[k] + (ztrad[k] * ((ztra[k] - zdbt[k]) * zrefd[nl] + zdbt[k] * zref[nl])) * zreflect;
    zrupd[k] = zrefd[k] + ztrad[k] * ztrad[k] * zrefd[nl] * zreflect;
    for (int ikx = nl - 2; ikx >= 0; --ikx) {
      int ikp = ikx + 1;
      zreflect = 1.0 / (1.0 - zrupd[ikp] * zrefd[ikx]);
      zrup[ikx] = zref[ikx] + (ztrad[ikx] * ((ztra[ikx] - zdbt[ikx]) * zrupd[ikp] + zdbt[ikx] * zrup[ikp])) * zreflect;
      zrupd[ikx] = zrefd[ikx] + ztrad[ikx] * ztrad[ikx] * zrupd[ikp] * zreflect;
    }
  }
  // top-down transmittances and the fluxes at every level
  double zinc = adjflux * sflux * prmu0 * w;
  double ztdn = 1.0, zrdnd = 0.0;
  for (int jk = 0; jk <= nl; ++jk) {
    if (jk == 1) { ztdn = ztra[0]; zrdnd = zrefd[0]; }
    else if (jk >= 2) {
      int j = jk - 1;
      double zr = 1.0 / (1.0 - zrefd[j] * zrdnd);
      double tdn_new = ztdbt[j] * ztra[j] + (ztrad[j] * ((ztdn - ztdbt[j]) + ztdbt[j] * zref[j] * zrdnd)) * zr;
      double rdnd_new = zrefd[j] + ztrad[j] * ztrad[j] * zrdnd * zr;
      ztdn = tdn_new; zrdnd = rdnd_new;
    }
    double zreflect = 1.0 / (1.0 - zrdnd * zrupd[jk]);
    double fu = (ztdbt[jk] * zrup[jk] + (ztdn - ztdbt[jk]) * zrupd[jk]) * zreflect;
    double fd = ztdbt[jk] + (ztdn - ztdbt[jk] + ztdbt[jk] * zrup[jk] * zrdnd) * zreflect;
    red.up(nl - jk, zinc * fu);
    red.down(nl - jk, zinc * fd);
  }
}

}  // namespace rrtm
