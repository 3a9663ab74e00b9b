// physics_bm_column.h -- the full Betts-Miller convection scheme (atmos_param/betts_miller/betts_miller.f90) for one column as a
// `__host__ __device__` function: physics_bm.cu runs it from one CUDA thread per column, tests/host/bm_host.cpp (test
// infrastructure) runs the same function in a serial loop on a machine without a GPU.
//
//   betts_miller      :86-438    relaxation to the parcel's moist adiabat / rhbm-scaled humidity, energy-conserving deep convection
//                                (do_simp or the uniform temperature-reference shift), shallow convection (do_shallower, do_changeqref)
//   capecalcnew       :444-776   CAPE, CIN, LCL, LZB of the lowest-level parcel (avgbl = .false., as betts_miller calls it)
//   lcltabl           :779-845   LCL temperature from the 127-entry table (bm_lcltable.h)
//   escomp                        lookup_es of sat_vapor_pres_mod (table + 2nd-order Taylor, sat_vapor_pres_k.F90:1132-1158)
//
// Level indices follow the Fortran (1-based, klzb / klcl = 0: none); arrays are addressed as a[(k-1)*stride + col].
// do_taucape is not built (the reference overwrites the module's tau_bm inside the grid loop: the result depends on the order of the
// columns).  Two out-of-bounds reads of the reference get a defined meaning: a parcel still buoyant at the model top (klzb left 0,
// `do k=klzb,kx`) -> klzb = 1; lcltabl at value >= -10.4 (lcltable(128) with weight 0) -> the last table entry.
#pragma once
#include <cmath>
#include <cstddef>
#include "bm_lcltable.h"

#if defined(__CUDACC__)
#define BM_HD __host__ __device__ __forceinline__
#else
#define BM_HD inline
#endif

namespace isca_bm {

struct BmSvp { const double *tab, *dtab, *d2tab; double tminl, dtinvl, tepsl, dtres; int n; };

struct BmConst {
  double tau_bm, rhbm, buoyancy_kick;
  int do_simp, do_shallower, do_changeqref, do_envsat;
  double rdgas, rvgas, cp_air, hlv, kappa, grav, es0;
  const double* lcltable;                     // 127 entries (device or host copy of ISCA_BM_LCLTABLE_VALUES)
};

BM_HD double bm_escomp(const BmSvp& s, double T, int& bad) {
  const double tmp = T - s.tminl;
  const double x = s.dtinvl * (tmp + s.tepsl);
  if (!(x > -1.0 && x < (double)s.n)) { bad |= 1; return 0.0; }
  const int ind = (int)x;
  const double dl = tmp - s.dtres * (double)ind;
  return s.tab[ind] + dl * (s.dtab[ind] + dl * s.d2tab[ind]);
}

BM_HD double bm_lcltabl(const BmConst& c, double value) {
  double v1 = value;
  if (value < -23.0) v1 = -23.0;
  if (value > -10.4) v1 = -10.4;
  const int ival = (int)floor(10. * (v1 + 23.0));
  const double v2 = -230. + ival;
  v1 = 10. * v1;
  if (ival + 1 > 126) return c.lcltable[126];
  return (v2 + 1.0 - v1) * c.lcltable[ival] + (v1 - v2) * c.lcltable[ival + 1];
}

// tp, rp: work arrays of kx doubles owned by the caller (thread-local), 0-based level index.
// Outputs: tdel, qdel, q_ref, t_ref (strided like the inputs); the scalars through `out`.
struct BmOut { double rain, cape, cin, invtau_t, invtau_q; int bmflag, klzb, klcl, bad; };

BM_HD void bm_column(const BmConst& c, const BmSvp& s, int kx, size_t stride, size_t col, double dt, const double* tin_, const double* qin_,
                     const double* pfull_, const double* phalf_, double* tdel_, double* qdel_, double* qref_, double* tref_, double* tp,
                     double* rp, BmOut& out) {
  // 1-based accessors
  auto tin = [&](int k) { return tin_[(size_t)(k - 1) * stride + col]; };
  auto qin = [&](int k) { return qin_[(size_t)(k - 1) * stride + col]; };
  auto rin = [&](int k) { const double q = qin_[(size_t)(k - 1) * stride + col]; return q / (1.0 - q); };
  auto p = [&](int k) { return pfull_[(size_t)(k - 1) * stride + col]; };
  auto ph = [&](int k) { return phalf_[(size_t)(k - 1) * stride + col]; };
  auto dlnph = [&](int k) { return log(ph(k + 1) / ph(k)); };
#define TP(k) tp[(k) - 1]
#define RP(k) rp[(k) - 1]
#define TDEL(k) tdel_[(size_t)((k) - 1) * stride + col]
#define QDEL(k) qdel_[(size_t)((k) - 1) * stride + col]
#define QREF(k) qref_[(size_t)((k) - 1) * stride + col]
#define TREF(k) tref_[(size_t)((k) - 1) * stride + col]
  const double kappa = c.kappa, rdgas = c.rdgas, rvgas = c.rvgas, hlv = c.hlv, cp_air = c.cp_air, grav = c.grav;
  const double pstar = 1.e5, small = 1.e-10;
  int bad = 0;

  // ------------------------------------------------------------------ capecalcnew
  bool nocape = true;
  double cape = 0., cin = 0.;
  int klcl = 0, klzb = 0;
  for (int k = 1; k <= kx; ++k) { TP(k) = tin(k); RP(k) = rin(k); }
  {
    const double t0 = tin(kx) + c.buoyancy_kick;
    const double r0 = rin(kx);
    double es = bm_escomp(s, t0, bad);
    const double rs = rdgas / rvgas * es / p(kx);
    bool ascend = true;                        // false = `go to 11`
    if (r0 >= rs) {
      klcl = kx;
      TP(kx) = t0 + (r0 - rs) / (cp_air / (hlv + small) + hlv * rs / rvgas / (t0 * t0));
      es = bm_escomp(s, TP(kx), bad);
      RP(kx) = rdgas / rvgas * es / p(kx);
    } else {
      const double theta0 = t0 * pow(pstar / p(kx), kappa);
      double plcl, tlcl;
      if (r0 > 0.) {
        const double value = log(pow(theta0, -1 / kappa) * r0 * pstar * rvgas / rdgas / c.es0);
        tlcl = bm_lcltabl(c, value);
        plcl = pstar * pow(tlcl / theta0, 1 / kappa);
        if (plcl < p(1)) { plcl = p(1); tlcl = theta0 * pow(plcl / pstar, kappa); }
        int k = kx;
        while (k >= 1 && p(k) > plcl) {        // adiabatic ascent below the LCL (k >= 1 always holds: plcl >= p(1))
          TP(k) = theta0 * pow(p(k) / pstar, kappa);
          es = bm_escomp(s, TP(k), bad);
          RP(k) = rdgas / rvgas * es / p(k);
          cin = cin + rdgas * (tin(k) - TP(k)) * dlnph(k);
          k = k - 1;
        }
        if (k < 1) k = 1;
        klcl = k;
        if (klcl == 1) klcl = 2;
        double a = kappa * tlcl + hlv / cp_air * r0;
        double b = hlv * hlv * r0 / cp_air / rvgas / (tlcl * tlcl);
        double dtdlnp = a / (1. + b);
        TP(klcl) = tlcl + dtdlnp * log(p(klcl) / plcl) / 2.;
        if (TP(klcl) < 173.16 && nocape) ascend = false;
        if (ascend) {
          es = bm_escomp(s, TP(klcl), bad);
          RP(klcl) = rdgas / rvgas * es / (p(klcl) + plcl) * 2.;
          a = kappa * TP(klcl) + hlv / cp_air * RP(klcl);
          b = hlv * hlv / cp_air / rvgas * RP(klcl) / (TP(klcl) * TP(klcl));
          dtdlnp = a / (1. + b);
          TP(klcl) = tlcl + dtdlnp * log(p(klcl) / plcl);
          if (TP(klcl) < 173.16 && nocape) ascend = false;
        }
        if (ascend) {
          es = bm_escomp(s, TP(klcl), bad);
          RP(klcl) = rdgas / rvgas * es / p(klcl);
          if (TP(klcl) < tin(klcl) && nocape) {
            cin = cin + rdgas * (tin(klcl) - TP(klcl)) * dlnph(klcl);
          } else {
            cape = cape + rdgas * (TP(klcl) - tin(klcl)) * dlnph(klcl);
            if (nocape) nocape = false;
          }
        }
      } else {                                 // dry parcel: LCL at the top level, no moist ascent
        for (int k = 1; k <= kx; ++k) {
          TP(k) = theta0 * pow(p(k) / pstar, kappa);
          RP(k) = 0.;
          cin = cin + rdgas * (tin(k) - TP(k)) * dlnph(k);
        }
        ascend = false;
      }
    }
    if (ascend) {
      for (int k = klcl - 1; k >= 1; --k) {    // moist adiabatic ascent, RK2 in ln p
        double a = kappa * TP(k + 1) + hlv / cp_air * RP(k + 1);
        double b = hlv * hlv / cp_air / rvgas * RP(k + 1) / (TP(k + 1) * TP(k + 1));
        double dtdlnp = a / (1. + b);
        const double dl = log(p(k) / p(k + 1));
        TP(k) = TP(k + 1) + dtdlnp * dl / 2.;
        if (TP(k) < 173.16 && nocape) break;
        es = bm_escomp(s, TP(k), bad);
        RP(k) = rdgas / rvgas * es / (p(k) + p(k + 1)) * 2.;
        a = kappa * TP(k) + hlv / cp_air * RP(k);
        b = hlv * hlv / cp_air / rvgas * RP(k) / (TP(k) * TP(k));
        dtdlnp = a / (1. + b);
        TP(k) = TP(k + 1) + dtdlnp * dl;
        if (TP(k) < 173.16 && nocape) break;
        es = bm_escomp(s, TP(k), bad);
        RP(k) = rdgas / rvgas * es / p(k);
        if (TP(k) < tin(k) && nocape) {
          cin = cin + rdgas * (tin(k) - TP(k)) * dlnph(k);
        } else if (TP(k) < tin(k) && !nocape) {
          klzb = k + 1;
          break;
        } else {
          cape = cape + rdgas * (TP(k) - tin(k)) * dlnph(k);
          if (nocape) nocape = false;
        }
      }
    }
    if (nocape) {                              // label 11
      klzb = 0; cin = 0.;
      for (int k = 1; k <= kx; ++k) { TP(k) = tin(k); RP(k) = rin(k); }
    }
  }

  // ------------------------------------------------------------------ betts_miller
  out.cape = cape; out.cin = cin; out.klcl = klcl;
  out.invtau_t = 0.; out.invtau_q = 0.;
  int bmflag = 0;
  double precip = 0.;
  const double tau_bm = c.tau_bm;
  auto no_adjustment = [&]() {
    for (int k = 1; k <= kx; ++k) { TDEL(k) = 0.0; QDEL(k) = 0.0; QREF(k) = qin(k); TREF(k) = tin(k); }
    precip = 0.0; out.invtau_t = 0.; out.invtau_q = 0.;
  };
  if (cape > 0.) {
    if (klzb == 0) klzb = 1;
    bmflag = 1;
    for (int k = 1; k <= kx; ++k) TREF(k) = TP(k);
    for (int k = klzb; k <= kx; ++k) {
      if (c.do_envsat) {
        const double es = bm_escomp(s, tin(k), bad) * c.rhbm;
        RP(k) = rdgas / rvgas * es / p(k);
      } else {
        RP(k) = c.rhbm * RP(k);
      }
      QREF(k) = RP(k) / (1 + RP(k));
    }
    const int kz = klzb - 1 > 1 ? klzb - 1 : 1;
    for (int k = 1; k <= kz; ++k) { QDEL(k) = 0.0; TDEL(k) = 0.0; QREF(k) = qin(k); TREF(k) = tin(k); }
    double precip_t = 0.;
    for (int k = klzb; k <= kx; ++k) {
      const double td = -(tin(k) - TREF(k)) / tau_bm * dt;
      const double qd = -(qin(k) - QREF(k)) / tau_bm * dt;
      TDEL(k) = td; QDEL(k) = qd;
      precip = precip - qd * (ph(k + 1) - ph(k)) / grav;
      precip_t = precip_t + cp_air / (hlv + small) * td * (ph(k + 1) - ph(k)) / grav;
    }
    if (precip > 0. && precip_t > 0.) {
      bmflag = 2;
      if (precip > precip_t) {
        out.invtau_q = precip_t / precip / tau_bm;
        for (int k = klzb; k <= kx; ++k) QDEL(k) = tau_bm * out.invtau_q * QDEL(k);
        precip = precip_t;
        out.invtau_t = 1. / tau_bm;
      } else if (c.do_simp) {
        out.invtau_t = precip / precip_t / tau_bm;
        for (int k = klzb; k <= kx; ++k) TDEL(k) = tau_bm * out.invtau_t * TDEL(k);
        out.invtau_q = 1. / tau_bm;
      } else {
        double deltak = 0.;
        for (int k = klzb; k <= kx; ++k) deltak = deltak - (TDEL(k) + hlv / cp_air * QDEL(k)) * (ph(k + 1) - ph(k));
        deltak = deltak / (ph(kx + 1) - ph(klzb));
        for (int k = klzb; k <= kx; ++k) { TREF(k) = TREF(k) + deltak * tau_bm / dt; TDEL(k) = TDEL(k) + deltak; }
      }
    } else if (precip_t > 0.) {
      if (c.do_shallower) {
        int ktop = klzb;
        while (precip < 0. && ktop <= kx) {
          precip = precip - QDEL(ktop) * (ph(ktop) - ph(ktop + 1)) / grav;
          ktop = ktop + 1;
        }
        ktop = ktop - 1;
        if (ktop > klzb)
          for (int k = klzb; k <= ktop - 1; ++k) { QDEL(k) = 0.; QREF(k) = qin(k); TDEL(k) = 0.; TREF(k) = tin(k); }
        if (precip > 0.) {
          const double ptopfrac = precip / (QDEL(ktop) * (ph(ktop + 1) - ph(ktop))) * grav;
          QDEL(ktop) = ptopfrac * QDEL(ktop);
          precip = 0.;
          TDEL(ktop) = ptopfrac * TDEL(ktop);
          if (ktop < kx) {
            double deltak = 0.;
            for (int k = ktop; k <= kx; ++k) deltak = deltak + TDEL(k) * (ph(k) - ph(k + 1));
            deltak = deltak / (ph(kx + 1) - ph(ktop));
            for (int k = ktop; k <= kx; ++k) { TDEL(k) = TDEL(k) + deltak; TREF(k) = TREF(k) + deltak * tau_bm / dt; }
          }
        } else {
          precip = 0.;
          QDEL(kx) = 0.; QREF(kx) = qin(kx); TDEL(kx) = 0.; TREF(kx) = tin(kx);
          out.invtau_t = 0.; out.invtau_q = 0.;
        }
      } else if (c.do_changeqref) {
        double deltak = 0., deltaq = 0., qrefint = 0.;
        for (int k = klzb; k <= kx; ++k) {
          const double dp = ph(k) - ph(k + 1);
          deltaq = deltaq - QDEL(k) * tau_bm / dt * dp;
          deltak = deltak + TDEL(k) * dp;
          qrefint = qrefint - QREF(k) * dp;
        }
        deltak = deltak / (ph(kx + 1) - ph(klzb));
        const double deltaqfrac = 1. - deltaq / qrefint;
        const double deltaqfrac2 = -deltaq / qrefint * dt / tau_bm;
        precip = 0.0;
        for (int k = klzb; k <= kx; ++k) {
          QDEL(k) = QDEL(k) + deltaqfrac2 * QREF(k);
          QREF(k) = deltaqfrac * QREF(k);
          TDEL(k) = TDEL(k) + deltak;
          TREF(k) = TREF(k) + deltak * tau_bm / dt;
        }
      } else {
        precip = 0.;
        for (int k = 1; k <= kx; ++k) { TDEL(k) = 0.; QDEL(k) = 0.; }
        out.invtau_t = 0.; out.invtau_q = 0.;
      }
    } else {
      no_adjustment();
    }
  } else {
    no_adjustment();
  }
  out.rain = precip; out.bmflag = bmflag; out.klzb = klzb; out.bad = bad;
#undef TP
#undef RP
#undef TDEL
#undef QDEL
#undef QREF
#undef TREF
}

}  // namespace isca_bm
