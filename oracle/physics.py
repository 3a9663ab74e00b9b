"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the column-physics rows built so far:

  SatVaporPres + compute_qs   shared/sat_vapor_pres/sat_vapor_pres_k.F90:161-266 (do_simple tables), :1132-1158, :457-540
  lscale_cond + precip_evap   atmos_param/lscale_cond/lscale_cond.F90:79-255
  two_stream_gray_rad (Frierson) down / up   atmos_param/two_stream_gray_rad/two_stream_gray_rad.F90:386-655, 659-776
  rayleigh sponge             atmos_param/damping_driver/damping_driver.f90:404-420, 594-636

Arrays are [lev, lat, lon] (== Fortran (lon, lat, lev)), k = 0 is the model top.  Parity unpinned by the
reference (no known-answer tests for these routines); pinned by conservation properties in tests/test_oracle_physics.py."""
from __future__ import annotations
from dataclasses import dataclass
import numpy as np

GRAV = 9.80
RDGAS = 287.04
KAPPA = 2.0 / 7.0
CP_AIR = RDGAS / KAPPA
RVGAS = 461.50
HLV = 2.500e6
TFREEZE = 273.16
STEFAN = 5.6734e-8
PSTD_MKS = 101325.0
PI = 3.14159265358979323846


class SatVaporPres:
    """sat_vapor_pres_init_k with do_simple=.true. (tcmin=-173, tcmax=350, esres=10) and the lookup
    (2nd-order Taylor within a 0.1 K bin)."""

    def __init__(self, es0: float = 1.0):
        tcmin, tcmax, esres = -173, 350, 10
        n = (tcmax - tcmin) * esres + 1
        self.table_siz = n
        self.dtres = (tcmax - tcmin) / (n - 1)
        self.tminl = float(tcmin) + TFREEZE
        self.dtinvl = 1.0 / self.dtres
        self.tepsl = 0.5 * self.dtres
        tem = self.tminl + self.dtres * np.arange(n, dtype=np.float64)
        self.TABLE = es0 * 610.78 * np.exp(-HLV / RVGAS * (1.0 / tem - 1.0 / TFREEZE))
        self.DTABLE = HLV * self.TABLE / RVGAS / tem ** 2.0
        d2 = np.zeros(n)
        d2[1:-1] = 0.25 * self.dtinvl * (self.DTABLE[2:] - self.DTABLE[:-2])
        d2[0] = 0.50 * self.dtinvl * (self.DTABLE[1] - self.DTABLE[0])
        d2[-1] = 0.50 * self.dtinvl * (self.DTABLE[-1] - self.DTABLE[-2])
        self.D2TABLE = d2

    def lookup_es_des(self, temp):
        tmp = temp - self.tminl
        ind = np.trunc(self.dtinvl * (tmp + self.tepsl)).astype(np.int64)
        if np.any((ind < 0) | (ind >= self.table_siz)):
            raise FloatingPointError("lookup_es: temperature out of the table range")
        dl = tmp - self.dtres * ind
        es = self.TABLE[ind] + dl * (self.DTABLE[ind] + dl * self.D2TABLE[ind])
        des = self.DTABLE[ind] + 2.0 * dl * self.D2TABLE[ind]
        return es, des

    def compute_qs(self, temp, press, hc=1.0):
        """compute_qs_k_3d without q (use_exact_qs irrelevant): qs and dqs/dT."""
        eps = RDGAS / RVGAS
        es, des = self.lookup_es_des(temp)
        des = des * hc
        es = es * hc
        denom = press - (1.0 - eps) * es
        qs = np.where(denom > 0.0, eps * es / np.where(denom > 0.0, denom, 1.0), eps)
        dqs = eps * press * des / denom ** 2
        return qs, dqs


def lscale_cond(svp: SatVaporPres, tin, qin, pfull, phalf, hc=1.0, do_evap=True):
    """lscale_cond (do_simple=.true.: no snow, hlcp = HLv/cp everywhere). Returns rain, tdel, qdel."""
    hlcp = HLV / CP_AIR
    qsat, dqsat = svp.compute_qs(tin, pfull, hc)
    adj = (qin - qsat) * qsat > 0.0
    qdel = np.where(adj, (qsat - qin) / (1.0 + hlcp * dqsat), 0.0)
    tdel = np.where(adj, -hlcp * qdel, 0.0)
    pmass = (phalf[1:] - phalf[:-1]) / GRAV
    if do_evap:                                               # precip_evap :216-255
        exq = np.zeros(tin.shape[1:])
        for k in range(tin.shape[0]):
            neg = qdel[k] < 0.0
            exq = np.where(neg, exq - qdel[k] * pmass[k], exq)
            ev = (qdel[k] >= 0.0) & (exq > 0.0)
            exq_l = exq / pmass[k]
            d = (qsat[k] - qin[k]) / (1.0 + hlcp * dqsat[k])
            d = np.minimum(np.maximum(d, 0.0), exq_l)
            qdel[k] = np.where(ev, qdel[k] + d, qdel[k])
            tdel[k] = np.where(ev, tdel[k] - d * hlcp, tdel[k])
            exq = np.where(ev, (exq_l - d) * pmass[k], exq)
    precip = np.zeros(tin.shape[1:])
    for k in range(tin.shape[0]):
        precip = precip - pmass[k] * qdel[k]
    rain = np.maximum(precip, 0.0)
    return rain, tdel, qdel


@dataclass
class GreyRadConfig:
    """two_stream_gray_rad_nml defaults (two_stream_gray_rad.F90:72-113), rad_scheme='frierson', do_seasonal=.false."""
    solar_constant: float = 1360.0
    del_sol: float = 1.4
    del_sw: float = 0.0
    ir_tau_eq: float = 6.0
    ir_tau_pole: float = 1.5
    atm_abs: float = 0.0
    sw_diff: float = 0.0
    linear_tau: float = 0.1
    wv_exponent: float = 4.0
    solar_exponent: float = 4.0
    odp: float = 1.0
    diabatic_acce: float = 1.0


class GreyRadiation:
    def __init__(self, cfg: GreyRadConfig):
        self.c = cfg

    def down(self, lat, p_half, t):
        """two_stream_gray_rad_down: lat [lat, lon] (radians), p_half [K+1,..], t [K,..]."""
        c = self.c
        n = t.shape[0]
        p2 = (1.0 - 3.0 * np.sin(lat) ** 2) / 4.0
        insolation = 0.25 * c.solar_constant * (1.0 + c.del_sol * p2 + c.del_sw * np.sin(lat))
        sw_tau_0 = (1.0 - c.sw_diff * np.sin(lat) ** 2) * c.atm_abs
        sw_tau = sw_tau_0[None] * (p_half / PSTD_MKS) ** c.solar_exponent
        sw_down = insolation[None] * np.exp(-sw_tau)
        b = STEFAN * t ** 4
        lw_tau_0 = c.ir_tau_eq + (c.ir_tau_pole - c.ir_tau_eq) * np.sin(lat) ** 2
        lw_tau_0 = lw_tau_0 * c.odp
        lw_tau = lw_tau_0[None] * (c.linear_tau * p_half / PSTD_MKS + (1.0 - c.linear_tau) * (p_half / PSTD_MKS) ** c.wv_exponent)
        lw_dtrans = np.exp(-(lw_tau[1:] - lw_tau[:-1]))
        lw_down = np.zeros_like(p_half)
        for k in range(n):
            lw_down[k + 1] = lw_down[k] * lw_dtrans[k] + b[k] * (1.0 - lw_dtrans[k])
        self._st = dict(sw_down=sw_down, lw_down=lw_down, lw_dtrans=lw_dtrans, b=b)
        return dict(surf_lw_down=lw_down[n], sw_down_surf=sw_down[n])

    def up(self, t_surf, albedo, p_half, tdt):
        """two_stream_gray_rad_up; net_surf_sw_down = (1-albedo)*sw_down(surface) (:647)."""
        c, st = self.c, self._st
        n = st["b"].shape[0]
        lw_up = np.zeros_like(p_half)
        lw_up[n] = STEFAN * t_surf ** 4
        for k in range(n - 1, -1, -1):
            lw_up[k] = lw_up[k + 1] * st["lw_dtrans"][k] + st["b"][k] * (1.0 - st["lw_dtrans"][k])
        sw_up = albedo[None] * st["sw_down"][n][None] + 0 * p_half
        lw_flux = lw_up - st["lw_down"]
        sw_flux = sw_up - st["sw_down"]
        rad_flux = lw_flux + sw_flux
        tdt_rad = c.diabatic_acce * (rad_flux[1:] - rad_flux[:-1]) * GRAV / (CP_AIR * (p_half[1:] - p_half[:-1]))
        return tdt + tdt_rad, dict(olr=lw_up[0], net_lw_surf=lw_flux[n], rad_flux=rad_flux)


def rayleigh_sponge(dt, p_full, u, v, pref, sponge_pbottom=5000.0, trayfric=-0.25, do_conserve_energy=True):
    """damping_driver 'rayleigh' (damping_driver.f90:404-420, 594-636). pref: reference full pressures + surface."""
    nlev = int(np.argmin(np.abs(pref - 2 * sponge_pbottom))) + 1
    rfactr = (1.0 / trayfric) if trayfric > 0.0 else (1.0 / abs(trayfric)) * (1.0 / 86400.0)
    udt = np.zeros_like(u); vdt = np.zeros_like(v); tdt = np.zeros_like(u)
    for k in range(nlev):
        m = p_full[k] < sponge_pbottom
        fact = rfactr * (sponge_pbottom - p_full[k]) ** 2 / (sponge_pbottom) ** 2
        udt[k] = np.where(m, -u[k] * fact, 0.0)
        vdt[k] = np.where(m, -v[k] * fact, 0.0)
    if do_conserve_energy:
        for k in range(nlev):
            tdt[k] = -((u[k] + 0.5 * dt * udt[k]) * udt[k] + (v[k] + 0.5 * dt * vdt[k]) * vdt[k]) / CP_AIR
    return udt, vdt, tdt, nlev
