"""TEST INFRASTRUCTURE -- restatement of betts_miller_mod (atmos_param/betts_miller/betts_miller.f90), the full Betts-Miller
convection scheme of `convection_scheme = 'FULL_BETTS_MILLER'` (idealized_moist_phys.F90:889-916).

Written column by column with scalar loops exactly as the reference (small test sizes only).  1-based level indices as in the
Fortran (arrays are padded with an unused element 0) so that klzb / klcl / ktop read as they do there.

Pinned part: the reference ships the LCL lookup table `lcltable` (betts_miller.f90:800-833, 127 temperatures) -- known-answer data
for the relation value = log(es(T)/T**(1/kappa)) (es = the do_simple Clausius-Clapeyron form with es0 = 1): the table is committed as
tests/golden/bm_lcltable.py and tests/test_betts_miller.py checks that relation to the 8 printed digits.  Everything else: parity
unpinned (no known-answer data in the reference); `capecalcnew` is cross-checked against the independent restatement of
the simplified scheme's CAPE routine in oracle/physics.py where the two coincide.

Two places where the reference indexes out of bounds are given a defined meaning (same in the CUDA code): a parcel that is still
buoyant at the model top leaves klzb = 0 and `do k=klzb,kx` reads level 0 -> klzb = 1; `lcltabl` at value >= -10.4 reads
lcltable(128) with weight zero -> the last entry.

Not restated: do_taucape (the reference overwrites the module variable tau_bm inside the grid loop, so that the result depends on the
order in which columns are visited, betts_miller.f90:237-240), the optional `mask` / `conv` arguments (unused by idealized_moist_phys).
"""
from __future__ import annotations
import math
from dataclasses import dataclass
import numpy as np

from .physics import SatVaporPres, KAPPA, RDGAS, RVGAS, CP_AIR, HLV, GRAV


@dataclass
class BettsMillerConfig:
    """betts_miller_nml (betts_miller.f90:56-70)"""
    tau_bm: float = 7200.0
    rhbm: float = 0.8
    do_simp: bool = True
    do_shallower: bool = False
    do_changeqref: bool = False
    do_envsat: bool = False
    do_taucape: bool = False
    capetaubm: float = 900.0
    tau_min: float = 2400.0
    buoyancy_kick: float = 0.0


class BettsMiller:
    def __init__(self, svp: SatVaporPres, cfg: BettsMillerConfig = None, lcltable=None, es0=1.0):
        self.svp, self.c, self.es0 = svp, cfg or BettsMillerConfig(), es0
        if self.c.do_taucape:
            raise ValueError("do_taucape is not restated (order-dependent in the reference)")
        if lcltable is None:
            import os, runpy
            here = os.path.dirname(os.path.abspath(__file__))
            lcltable = runpy.run_path(os.path.join(here, "..", "tests", "golden", "bm_lcltable.py"))["LCLTABLE"]
        self.lcltable = np.asarray(lcltable, dtype=float)
        assert self.lcltable.size == 127

    # ---- escomp = lookup_es of sat_vapor_pres_mod (table + 2nd-order Taylor)
    def escomp(self, T):
        s = self.svp
        tmp = T - s.tminl
        x = s.dtinvl * (tmp + s.tepsl)
        ind = int(x)
        if ind < 0 or ind >= s.table_siz or not (x > -1.0):
            raise FloatingPointError("escomp: temperature out of the table range")
        dl = tmp - s.dtres * ind
        return s.TABLE[ind] + dl * (s.DTABLE[ind] + dl * s.D2TABLE[ind])

    # ---- lcltabl (:779-845)
    def lcltabl(self, value):
        v1 = value
        if value < -23.0:
            v1 = -23.0
        if value > -10.4:
            v1 = -10.4
        ival = int(math.floor(10. * (v1 + 23.0)))
        v2 = -230. + ival
        v1 = 10. * v1
        t = self.lcltable
        if ival + 1 > 126:                       # value >= -10.4: the Fortran reads lcltable(128), one past the table, with weight
            return t[126]                        # 0 -- the last entry is the defined part of that expression
        return (v2 + 1.0 - v1) * t[ival] + (v1 - v2) * t[ival + 1]      # lcltable(ival+1), lcltable(ival+2) 1-based

    # ---- capecalcnew (:444-776), avgbl = .false. (betts_miller sets it so, :175)
    def capecalcnew(self, kx, p, phalf, tin, rin):
        """p, tin, rin: [kx+1] 1-based; phalf: [kx+2] 1-based -> cape, cin, tp, rp, klzb, klcl"""
        c = self.c
        kappa, rdgas, rvgas, hlv, cp_air = KAPPA, RDGAS, RVGAS, HLV, CP_AIR
        pstar, small = 1.e5, 1.e-10
        nocape = True
        cape = cin = 0.
        klcl = klzb = 0
        tp, rp = tin.copy(), rin.copy()
        t0 = tin[kx] + c.buoyancy_kick
        r0 = rin[kx]
        es = self.escomp(t0)
        rs = rdgas / rvgas * es / p[kx]

        def finish():
            nonlocal cin, klzb, tp, rp
            if nocape:
                klzb = 0
                cin = 0.
                tp, rp = tin.copy(), rin.copy()
            return cape, cin, tp, rp, klzb, klcl

        if r0 >= rs:
            klcl = kx
            tp[kx] = t0 + (r0 - rs) / (cp_air / (hlv + small) + hlv * rs / rvgas / t0 ** 2.)
            es = self.escomp(tp[kx])
            rp[kx] = rdgas / rvgas * es / p[kx]
        else:
            theta0 = t0 * (pstar / p[kx]) ** kappa
            if r0 > 0.:
                value = math.log(theta0 ** (-1 / kappa) * r0 * pstar * rvgas / rdgas / self.es0)
                tlcl = self.lcltabl(value)
                plcl = pstar * (tlcl / theta0) ** (1 / kappa)
                if plcl < p[1]:
                    plcl = p[1]
                    tlcl = theta0 * (plcl / pstar) ** kappa
                k = kx
            else:
                plcl = p[1]
                tlcl = theta0 * (plcl / pstar) ** kappa
                for k in range(1, kx + 1):
                    tp[k] = theta0 * (p[k] / pstar) ** kappa
                    rp[k] = 0.
                    cin = cin + rdgas * (tin[k] - tp[k]) * math.log(phalf[k + 1] / phalf[k])
                return finish()
            while p[k] > plcl:
                tp[k] = theta0 * (p[k] / pstar) ** kappa
                es = self.escomp(tp[k])
                rp[k] = rdgas / rvgas * es / p[k]
                cin = cin + rdgas * (tin[k] - tp[k]) * math.log(phalf[k + 1] / phalf[k])
                k = k - 1
            klcl = k
            if klcl == 1:
                klcl = 2
            a = kappa * tlcl + hlv / cp_air * r0
            b = hlv ** 2. * r0 / cp_air / rvgas / tlcl ** 2.
            dtdlnp = a / (1. + b)
            tp[klcl] = tlcl + dtdlnp * math.log(p[klcl] / plcl) / 2.
            if tp[klcl] < 173.16 and nocape:
                return finish()
            es = self.escomp(tp[klcl])
            rp[klcl] = rdgas / rvgas * es / (p[klcl] + plcl) * 2.
            a = kappa * tp[klcl] + hlv / cp_air * rp[klcl]
            b = hlv ** 2. / cp_air / rvgas * rp[klcl] / tp[klcl] ** 2.
            dtdlnp = a / (1. + b)
            tp[klcl] = tlcl + dtdlnp * math.log(p[klcl] / plcl)
            if tp[klcl] < 173.16 and nocape:
                return finish()
            es = self.escomp(tp[klcl])
            rp[klcl] = rdgas / rvgas * es / p[klcl]
            if tp[klcl] < tin[klcl] and nocape:
                cin = cin + rdgas * (tin[klcl] - tp[klcl]) * math.log(phalf[klcl + 1] / phalf[klcl])
            else:
                cape = cape + rdgas * (tp[klcl] - tin[klcl]) * math.log(phalf[klcl + 1] / phalf[klcl])
                if nocape:
                    nocape = False
        for k in range(klcl - 1, 0, -1):
            a = kappa * tp[k + 1] + hlv / cp_air * rp[k + 1]
            b = hlv ** 2. / cp_air / rvgas * rp[k + 1] / tp[k + 1] ** 2.
            dtdlnp = a / (1. + b)
            tp[k] = tp[k + 1] + dtdlnp * math.log(p[k] / p[k + 1]) / 2.
            if tp[k] < 173.16 and nocape:
                return finish()
            es = self.escomp(tp[k])
            rp[k] = rdgas / rvgas * es / (p[k] + p[k + 1]) * 2.
            a = kappa * tp[k] + hlv / cp_air * rp[k]
            b = hlv ** 2. / cp_air / rvgas * rp[k] / tp[k] ** 2.
            dtdlnp = a / (1. + b)
            tp[k] = tp[k + 1] + dtdlnp * math.log(p[k] / p[k + 1])
            if tp[k] < 173.16 and nocape:
                return finish()
            es = self.escomp(tp[k])
            rp[k] = rdgas / rvgas * es / p[k]
            if tp[k] < tin[k] and nocape:
                cin = cin + rdgas * (tin[k] - tp[k]) * math.log(phalf[k + 1] / phalf[k])
            elif tp[k] < tin[k] and not nocape:
                klzb = k + 1
                return finish()
            else:
                cape = cape + rdgas * (tp[k] - tin[k]) * math.log(phalf[k + 1] / phalf[k])
                if nocape:
                    nocape = False
        return finish()

    # ---- betts_miller (:86-438) for one column; inputs 0-based [kx] (phalf [kx+1]), outputs 0-based
    def column(self, dt, tin0, qin0, pfull0, phalf0):
        c = self.c
        kx = tin0.size
        pad = lambda a: np.concatenate([[0.0], np.asarray(a, dtype=float)])
        tin, qin, pfull, phalf = pad(tin0), pad(qin0), pad(pfull0), pad(phalf0)
        grav, cp_air, hlv, rdgas, rvgas = GRAV, CP_AIR, HLV, RDGAS, RVGAS
        small = 1.e-10
        tau_bm = c.tau_bm
        rin = qin / (1.0 - qin)
        with np.errstate(divide="ignore", invalid="ignore"):      # log(phalf(2)/phalf(1)) with phalf(1) = 0 in the dry-parcel branch
            cape1, cin1, tpc, rpc, klzb, klcl = self.capecalcnew(kx, pfull, phalf, tin, rin)
        if cape1 > 0. and klzb == 0:
            klzb = 1          # parcel buoyant up to the model top: the reference then indexes level 0 (out of bounds); top level here
        tdel, qdel = np.zeros(kx + 1), np.zeros(kx + 1)
        q_ref, t_ref = np.zeros(kx + 1), np.zeros(kx + 1)
        bmflag = 0
        precip = 0.
        invtau_bm_t = invtau_bm_q = 0.        # left unset by the reference on some branches; 0 here and in the CUDA code

        def none():
            nonlocal precip, invtau_bm_t, invtau_bm_q
            tdel[:] = 0.0; qdel[:] = 0.0
            precip = 0.0
            q_ref[:] = qin; t_ref[:] = tin
            invtau_bm_t = invtau_bm_q = 0.

        if cape1 > 0.:
            bmflag = 1
            t_ref[:] = tpc
            for k in range(klzb, kx + 1):
                if c.do_envsat:
                    es = self.escomp(tin[k]) * c.rhbm
                    rpc[k] = rdgas / rvgas * es / pfull[k]
                    q_ref[k] = rpc[k] / (1 + rpc[k])
                else:
                    rpc[k] = c.rhbm * rpc[k]
                    q_ref[k] = rpc[k] / (1 + rpc[k])
            for k in range(1, max(klzb - 1, 1) + 1):
                qdel[k] = 0.0; tdel[k] = 0.0
                q_ref[k] = qin[k]; t_ref[k] = tin[k]
            precip = 0.
            precip_t = 0.
            for k in range(klzb, kx + 1):
                tdel[k] = - (tin[k] - t_ref[k]) / tau_bm * dt
                qdel[k] = - (qin[k] - q_ref[k]) / tau_bm * dt
                precip = precip - qdel[k] * (phalf[k + 1] - phalf[k]) / grav
                precip_t = precip_t + cp_air / (hlv + small) * tdel[k] * (phalf[k + 1] - phalf[k]) / grav
            if precip > 0. and precip_t > 0.:
                bmflag = 2
                if precip > precip_t:
                    invtau_bm_q = precip_t / precip / tau_bm
                    qdel[klzb:kx + 1] = tau_bm * invtau_bm_q * qdel[klzb:kx + 1]
                    precip = precip_t
                    invtau_bm_t = 1. / tau_bm
                else:
                    if c.do_simp:
                        invtau_bm_t = precip / precip_t / tau_bm
                        tdel[klzb:kx + 1] = tau_bm * invtau_bm_t * tdel[klzb:kx + 1]
                        invtau_bm_q = 1. / tau_bm
                    else:
                        deltak = 0.
                        for k in range(klzb, kx + 1):
                            deltak = deltak - (tdel[k] + hlv / cp_air * qdel[k]) * (phalf[k + 1] - phalf[k])
                        deltak = deltak / (phalf[kx + 1] - phalf[klzb])
                        t_ref[klzb:kx + 1] = t_ref[klzb:kx + 1] + deltak * tau_bm / dt
                        tdel[klzb:kx + 1] = tdel[klzb:kx + 1] + deltak
            elif precip_t > 0.:
                if c.do_shallower:
                    ktop = klzb
                    while precip < 0. and ktop <= kx:
                        precip = precip - qdel[ktop] * (phalf[ktop] - phalf[ktop + 1]) / grav
                        ktop = ktop + 1
                    ktop = ktop - 1
                    if ktop > klzb:
                        qdel[klzb:ktop] = 0.
                        q_ref[klzb:ktop] = qin[klzb:ktop]
                        tdel[klzb:ktop] = 0.
                        t_ref[klzb:ktop] = tin[klzb:ktop]
                    if precip > 0.:
                        ptopfrac = precip / (qdel[ktop] * (phalf[ktop + 1] - phalf[ktop])) * grav
                        qdel[ktop] = ptopfrac * qdel[ktop]
                        precip = 0.
                        tdel[ktop] = ptopfrac * tdel[ktop]
                        deltak = 0.
                        if ktop < kx:
                            for k in range(ktop, kx + 1):
                                deltak = deltak + tdel[k] * (phalf[k] - phalf[k + 1])
                            deltak = deltak / (phalf[kx + 1] - phalf[ktop])
                            for k in range(ktop, kx + 1):
                                tdel[k] = tdel[k] + deltak
                                t_ref[k] = t_ref[k] + deltak * tau_bm / dt
                    else:
                        precip = 0.
                        qdel[kx] = 0.
                        q_ref[kx] = qin[kx]
                        tdel[kx] = 0.
                        t_ref[kx] = tin[kx]
                        invtau_bm_t = invtau_bm_q = 0.
                elif c.do_changeqref:
                    deltak = deltaq = qrefint = 0.
                    for k in range(klzb, kx + 1):
                        deltaq = deltaq - qdel[k] * tau_bm / dt * (phalf[k] - phalf[k + 1])
                        deltak = deltak + tdel[k] * (phalf[k] - phalf[k + 1])
                        qrefint = qrefint - q_ref[k] * (phalf[k] - phalf[k + 1])
                    deltak = deltak / (phalf[kx + 1] - phalf[klzb])
                    deltaqfrac = 1. - deltaq / qrefint
                    deltaqfrac2 = - deltaq / qrefint * dt / tau_bm
                    precip = 0.0
                    for k in range(klzb, kx + 1):
                        qdel[k] = qdel[k] + deltaqfrac2 * q_ref[k]
                        q_ref[k] = deltaqfrac * q_ref[k]
                        tdel[k] = tdel[k] + deltak
                        t_ref[k] = t_ref[k] + deltak * tau_bm / dt
                else:
                    precip = 0.
                    tdel[:] = 0.; qdel[:] = 0.
                    invtau_bm_t = invtau_bm_q = 0.
            else:
                none()
        else:
            none()
        return dict(rain=precip, tdel=tdel[1:], qdel=qdel[1:], q_ref=q_ref[1:], t_ref=t_ref[1:], bmflag=bmflag, klzb=klzb, klcl=klcl,
                    cape=cape1, cin=cin1, invtau_bm_t=invtau_bm_t, invtau_bm_q=invtau_bm_q)

    def __call__(self, dt, tin, qin, p_full, p_half):
        """arrays [K, J, I] (p_half [K+1, J, I]) -> dict with the names SBMConvection uses (deltaT, deltaq, rain, convflag, CAPE, ...)"""
        K, J, I = tin.shape
        out3 = {n: np.zeros((K, J, I)) for n in ("tdel", "qdel", "q_ref", "t_ref")}
        out2 = {n: np.zeros((J, I)) for n in ("rain", "cape", "cin", "invtau_bm_t", "invtau_bm_q")}
        outi = {n: np.zeros((J, I), dtype=int) for n in ("bmflag", "klzb", "klcl")}
        for j in range(J):
            for i in range(I):
                o = self.column(dt, tin[:, j, i], qin[:, j, i], p_full[:, j, i], p_half[:, j, i])
                for n in out3:
                    out3[n][:, j, i] = o[n]
                for n in out2:
                    out2[n][j, i] = o[n]
                for n in outi:
                    outi[n][j, i] = o[n]
        return dict(deltaT=out3["tdel"], deltaq=out3["qdel"], qref=out3["q_ref"], Tref=out3["t_ref"], rain=out2["rain"], CAPE=out2["cape"],
                    CIN=out2["cin"], invtau_t=out2["invtau_bm_t"], invtau_q=out2["invtau_bm_q"], convflag=outi["bmflag"], kLZB=outi["klzb"],
                    kLCL=outi["klcl"])
