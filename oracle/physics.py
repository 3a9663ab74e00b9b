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


def compute_es(tem, tfreeze=None):
    """compute_es_k (sat_vapor_pres_k.F90:331-381): Goff-Gratch saturation vapour pressure over ice below freezing, over water above
    -20 C, linearly blended in between (Smithsonian Meteorological Tables p. 350)"""
    tfreeze = TFREEZE if tfreeze is None else tfreeze
    tem = np.asarray(tem, dtype=np.float64)
    ESBASW, ESBASI = 101324.60, 610.71
    TBASW, TBASI = tfreeze + 100., tfreeze
    x = -9.09718 * (TBASI / tem - 1.0) - 3.56654 * np.log10(TBASI / tem) + 0.876793 * (1.0 - tem / TBASI) + np.log10(ESBASI)
    esice = np.where(tem < TBASI, 10. ** x, 0.)
    x = -7.90298 * (TBASW / tem - 1) + 5.02808 * np.log10(TBASW / tem) - 1.3816e-07 * (10 ** ((1 - tem / TBASW) * 11.344) - 1) \
        + 8.1328e-03 * (10 ** ((TBASW / tem - 1) * (-3.49149)) - 1) + np.log10(ESBASW)
    esh2o = np.where(tem > -20. + TBASI, 10. ** x, 0.)
    blend = 0.05 * ((TBASI - tem) * esice + (tem - TBASI + 20.) * esh2o)
    return np.where(tem <= -20. + TBASI, esice, np.where(tem >= TBASI, esh2o, blend))


class SatVaporPres:
    """sat_vapor_pres_init_k (tcmin=-173, tcmax=350, esres=10; do_simple=.true. unless told otherwise) and the lookup
    (2nd-order Taylor within a 0.1 K bin)."""

    def __init__(self, es0: float = 1.0, do_simple: bool = True):
        tcmin, tcmax, esres = -173, 350, 10
        n = (tcmax - tcmin) * esres + 1
        self.table_siz = n
        self.dtres = (tcmax - tcmin) / (n - 1)
        self.tminl = float(tcmin) + TFREEZE
        self.dtinvl = 1.0 / self.dtres
        self.tepsl = 0.5 * self.dtres
        tem = self.tminl + self.dtres * np.arange(n, dtype=np.float64)
        if do_simple:
            self.TABLE = es0 * 610.78 * np.exp(-HLV / RVGAS * (1.0 / tem - 1.0 / TFREEZE))
            self.DTABLE = HLV * self.TABLE / RVGAS / tem ** 2.0
        else:                                                    # :238-246: centred difference over +-0.1*dtres, tfact = 5*dtinvl
            tinrc, tfact = .1 * self.dtres, 5 * self.dtinvl
            self.TABLE = compute_es(tem)
            self.DTABLE = (compute_es(tem + tinrc) - compute_es(tem - tinrc)) * tfact
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


PSTD_MKS_EARTH = 101325.0        # constants.F90:252


@dataclass
class GreyRadConfig:
    """two_stream_gray_rad_nml defaults (two_stream_gray_rad.F90:72-125), do_seasonal=.false.;
    rad_scheme in {'frierson', 'byrne', 'geen', 'schneider'} (:214-230)."""
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
    rad_scheme: str = "frierson"
    # Geen et al. 2016 two-band scheme (:96-104)
    ir_tau_co2_win: float = 0.2150
    ir_tau_wv_win1: float = 147.11
    ir_tau_wv_win2: float = 1.0814e4
    ir_tau_co2: float = 0.1
    ir_tau_wv1: float = 23.8
    ir_tau_wv2: float = 254.0
    window: float = 0.3732
    carbon_conc: float = 360.0
    # Schneider & Liu 2009 giant-planet scheme (:106-113)
    single_albedo: float = 0.8
    back_scatter: float = 0.398
    lw_tau_0_gp: float = 80.0
    sw_tau_0_gp: float = 3.0
    lw_tau_exponent_gp: float = 2.0
    sw_tau_exponent_gp: float = 1.0
    # Byrne & O'Gorman 2013 (:116-118)
    bog_a: float = 0.8678
    bog_b: float = 1997.9
    bog_mu: float = 1.0


class GreyRadiation:
    def __init__(self, cfg: GreyRadConfig):
        self.c = cfg
        self.scheme = cfg.rad_scheme.upper()
        if self.scheme not in ("FRIERSON", "BYRNE", "GEEN", "SCHNEIDER"):
            raise ValueError(f'"{cfg.rad_scheme}" is not a valid radiation scheme.')       # :228
        if self.scheme == "SCHNEIDER":                                                     # :233-238
            self.g_asym = 1 - 2.0 * cfg.back_scatter
            r1, r2 = np.sqrt(1.0 - self.g_asym * cfg.single_albedo), np.sqrt(1.0 - cfg.single_albedo)
            self.gp_albedo = (r1 - r2) / (r1 + r2)
            self.Ga_asym = 2.0 * np.sqrt((1.0 - cfg.single_albedo) * (1.0 - self.g_asym * cfg.single_albedo))

    def down(self, lat, p_half, t, q=None, albedo=None, insolation=None, carbon_conc=None):
        """two_stream_gray_rad_down (:386-655): lat [lat, lon] (radians), p_half [K+1,..], t [K,..], q [K,..] (BYRNE, GEEN);
        insolation [lat, lon]: the do_seasonal value `solar_constant * coszen` (:417-447, `seasonal_insolation` below)."""
        c, sch = self.c, self.scheme
        n = t.shape[0]
        cc_sw = getattr(self, "_cc_prev", c.carbon_conc)         # do_read_co2: the shortwave (:466) runs before the file is read (:519)
        cc_lw = c.carbon_conc if carbon_conc is None else carbon_conc
        self._cc_prev = cc_lw
        if insolation is not None:                                                         # do_seasonal wins over the scheme (:417)
            insolation = np.asarray(insolation, dtype=float)
        elif sch == "SCHNEIDER":
            insolation = (c.solar_constant / np.pi) * np.cos(lat)                          # :450-451
        else:
            p2 = (1.0 - 3.0 * np.sin(lat) ** 2) / 4.0
            insolation = 0.25 * c.solar_constant * (1.0 + c.del_sol * p2 + c.del_sw * np.sin(lat))
        # ---- shortwave (:458-508)
        sw_down = np.zeros_like(p_half)
        if sch == "GEEN":
            sw_tau_k = np.zeros_like(lat)
            sw_down[0] = insolation
            for k in range(n):
                sw_wv = sw_tau_k + 0.5194
                sw_wv = np.exp(0.01887 / (sw_tau_k + 0.009522) + 1.603 / (sw_wv * sw_wv))
                del_sol_tau = (0.0596 + 0.0029 * np.log(cc_sw / 360.0) + sw_wv * q[k]) * (p_half[k + 1] - p_half[k]) / p_half[n]
                sw_dtrans = np.exp(-del_sol_tau)
                sw_tau_k = sw_tau_k + del_sol_tau
                sw_down[k + 1] = sw_down[k] * sw_dtrans
        elif sch in ("FRIERSON", "BYRNE"):
            sw_tau_0 = (1.0 - c.sw_diff * np.sin(lat) ** 2) * c.atm_abs
            sw_tau = sw_tau_0[None] * (p_half / PSTD_MKS) ** c.solar_exponent
            sw_down = insolation[None] * np.exp(-sw_tau)
        else:
            sw_tau = c.sw_tau_0_gp * (p_half / PSTD_MKS) ** c.sw_tau_exponent_gp
            sw_down = insolation[None] * (1.0 - self.gp_albedo) * np.exp(-self.Ga_asym * sw_tau)
        # ---- longwave (:512-616)
        b = STEFAN * t ** 4
        st = {}
        if sch == "GEEN":
            dp = p_half[1:] - p_half[:-1]
            lw_del_tau = (c.ir_tau_co2 + 0.2023 * np.log(cc_lw / 360.0) + c.ir_tau_wv1 * np.log(c.ir_tau_wv2 * q + 1)) * dp / PSTD_MKS_EARTH
            lw_dtrans = np.exp(-lw_del_tau)
            lw_del_tau_win = (c.ir_tau_co2_win + 0.0954 * np.log(cc_lw / 360.0) + c.ir_tau_wv_win1 * q
                              + c.ir_tau_wv_win2 * q * q) * dp / PSTD_MKS_EARTH
            lw_dtrans_win = np.exp(-lw_del_tau_win)
            b_win = c.window * b
            b = (1.0 - c.window) * b
            lw_down = np.zeros_like(p_half); lw_down_win = np.zeros_like(p_half)
            for k in range(n):
                lw_down[k + 1] = lw_down[k] * lw_dtrans[k] + b[k] * (1.0 - lw_dtrans[k])
                lw_down_win[k + 1] = lw_down_win[k] * lw_dtrans_win[k] + b_win[k] * (1.0 - lw_dtrans_win[k])
            lw_down = lw_down + lw_down_win
            st.update(b_win=b_win, lw_dtrans_win=lw_dtrans_win)
        else:
            if sch == "BYRNE":
                lw_del_tau = (c.bog_a * c.bog_mu + 0.17 * np.log(cc_lw / 360.0) + c.bog_b * q) * ((p_half[1:] - p_half[:-1]) / PSTD_MKS_EARTH)
                lw_dtrans = np.exp(-lw_del_tau)
            elif sch == "FRIERSON":
                lw_tau_0 = c.ir_tau_eq + (c.ir_tau_pole - c.ir_tau_eq) * np.sin(lat) ** 2
                lw_tau_0 = lw_tau_0 * c.odp
                lw_tau = lw_tau_0[None] * (c.linear_tau * p_half / PSTD_MKS + (1.0 - c.linear_tau) * (p_half / PSTD_MKS) ** c.wv_exponent)
                lw_dtrans = np.exp(-(lw_tau[1:] - lw_tau[:-1]))
            else:
                lw_tau = c.lw_tau_0_gp * (p_half / PSTD_MKS) ** c.lw_tau_exponent_gp
                lw_dtrans = np.exp(-(lw_tau[1:] - lw_tau[:-1]))
            lw_down = np.zeros_like(p_half)
            for k in range(n):
                lw_down[k + 1] = lw_down[k] * lw_dtrans[k] + b[k] * (1.0 - lw_dtrans[k])
        st.update(sw_down=sw_down, lw_down=lw_down, lw_dtrans=lw_dtrans, b=b)
        if sch == "SCHNEIDER":                                   # :626-628 (needs the albedo of the down call)
            alb = albedo if albedo is not None else np.zeros_like(lat)
            st["b_surf_gp"] = lw_down[n] + sw_down[n] * (1.0 - alb)
        self._st = st
        return dict(surf_lw_down=lw_down[n], sw_down_surf=sw_down[n])

    def up(self, t_surf, albedo, p_half, tdt):
        """two_stream_gray_rad_up (:659-776); net_surf_sw_down = (1-albedo)*sw_down(surface) (:647)."""
        c, st, sch = self.c, self._st, self.scheme
        n = st["b"].shape[0]
        b_surf = STEFAN * t_surf ** 4
        lw_up = np.zeros_like(p_half)
        if sch == "GEEN":
            lw_up_win = np.zeros_like(p_half)
            lw_up[n] = b_surf * (1 - c.window)
            lw_up_win[n] = b_surf * c.window
            for k in range(n - 1, -1, -1):
                lw_up[k] = lw_up[k + 1] * st["lw_dtrans"][k] + st["b"][k] * (1.0 - st["lw_dtrans"][k])
                lw_up_win[k] = lw_up_win[k + 1] * st["lw_dtrans_win"][k] + st["b_win"][k] * (1.0 - st["lw_dtrans_win"][k])
            lw_up = lw_up + lw_up_win
        else:
            lw_up[n] = st["b_surf_gp"] if sch == "SCHNEIDER" else b_surf
            for k in range(n - 1, -1, -1):
                lw_up[k] = lw_up[k + 1] * st["lw_dtrans"][k] + st["b"][k] * (1.0 - st["lw_dtrans"][k])
        sw_up = albedo[None] * st["sw_down"][n][None] + 0 * p_half
        lw_flux = lw_up - st["lw_down"]
        sw_flux = sw_up - st["sw_down"]
        rad_flux = lw_flux + sw_flux
        tdt_rad = c.diabatic_acce * (rad_flux[1:] - rad_flux[:-1]) * GRAV / (CP_AIR * (p_half[1:] - p_half[:-1]))
        return tdt + tdt_rad, dict(olr=lw_up[0], net_lw_surf=lw_flux[n], rad_flux=rad_flux)


def seasonal_insolation(cfg: GreyRadConfig, astro, days, seconds, lat, lon, solday=-10, equinox_day=0.75,
                        use_time_average_coszen=False, dt_rad_avg=None, day_in_s=86400.0, year_in_s=360 * 86400.0):
    """do_seasonal branch of two_stream_gray_rad_down (two_stream_gray_rad.F90:417-447): `days`, `seconds` = get_time(Time_diag);
    astro: an astronomy_mod restatement with diurnal_solar(lat, lon, gmt, time_since_ae, dt) (oracle/rrtmg.py Astronomy);
    dt_rad_avg: seconds (two_stream_gray_rad_init sets it to dt_atmos when the namelist value is <= 0, :207)."""
    frac_of_day = float(seconds) / day_in_s
    if solday >= 0:
        frac_of_year = (float(solday) * day_in_s) / year_in_s
    else:
        frac_of_year = (float(seconds) + float(days) * day_in_s) / year_in_s
    gmt = abs(np.fmod(frac_of_day, 1.0)) * 2.0 * np.pi
    time_since_ae = ((frac_of_year - equinox_day) % 1.0) * 2.0 * np.pi
    dt = (float(dt_rad_avg) / day_in_s) * 2.0 * np.pi if use_time_average_coszen else None
    coszen = astro.diurnal_solar(lat, lon, gmt, time_since_ae, dt)[0]
    return cfg.solar_constant * coszen


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


# ------------------------------------------------------------------------------------------------------------------
# vertical diffusion (atmos_param/vert_diff/vert_diff.F90) and the slab mixed layer
# (atmos_spectral/driver/solo/mixed_layer.F90:568-745)
D608 = (RVGAS - RDGAS) / RDGAS


def compute_mu(p_half):
    """vert_diff.F90:1031-1046"""
    return GRAV / (p_half[1:] - p_half[:-1])


def compute_nu(diff, p_half, p_full, z_full, t, q, use_virtual_temp=True):
    """vert_diff.F90:1050-1087 (do_mcm_plev = .false.); nu(1) is never defined by the reference: 0 here."""
    tt = t * (1.0 + D608 * q) if use_virtual_temp else t
    nu = np.zeros_like(t)
    rho_half = 2.0 * p_half[1:-1] / (RDGAS * (tt[1:] + tt[:-1]))
    nu[1:] = rho_half * diff[1:] / (z_full[:-1] - z_full[1:])
    return nu


def explicit_tend(mu, nu, xi, dt_xi):
    """vert_diff.F90:1005-1027; returns the updated tendency."""
    fl = np.zeros_like(xi)
    fl[1:] = nu[1:] * (xi[1:] - xi[:-1])
    out = dt_xi.copy()
    out[:-1] = out[:-1] + mu[:-1] * (fl[1:] - fl[:-1])
    out[-1] = out[-1] - mu[-1] * fl[-1]
    return out


def compute_e(delt, mu, nu):
    """vert_diff.F90:951-980"""
    K = mu.shape[0]
    a = np.zeros_like(mu); c = np.zeros_like(mu); g = np.zeros_like(mu); e = np.zeros_like(mu)
    a[:-1] = -mu[:-1] * nu[1:] * delt
    c[1:] = -mu[1:] * nu[1:] * delt
    b = 1.0 - a - c
    e[0] = -a[0] / b[0]
    for k in range(1, K - 1):
        g[k] = 1.0 / (b[k] + c[k] * e[k - 1])
        e[k] = -a[k] * g[k]
    return e, a, b, c, g


def compute_f(dt_xi, b, c, g):
    """vert_diff.F90:984-1001"""
    f = np.zeros_like(dt_xi)
    f[0] = dt_xi[0] / b[0]
    for k in range(1, dt_xi.shape[0] - 1):
        f[k] = (dt_xi[k] - c[k] * f[k - 1]) * g[k]
    return f


def vert_diff_down_2(delt, mu, nu, xi_1, xi_2, dt_xi_1, dt_xi_2):
    """vert_diff.F90:806-862 (no kbot)"""
    d1 = explicit_tend(mu, nu, xi_1, dt_xi_1)
    d2 = explicit_tend(mu, nu, xi_2, dt_xi_2)
    e, a, b, c, g = compute_e(delt, mu, nu)
    f1 = compute_f(d1, b, c, g)
    f2 = compute_f(d2, b, c, g)
    return dict(e=e, f_1=f1, f_2=f2, mu_delt_n=mu[-1] * delt, nu_n=nu[-1], e_n1=e[-2],
                f_1_delt_n1=f1[-2] * delt, f_2_delt_n1=f2[-2] * delt, delta_1_n=d1[-1] * delt, delta_2_n=d2[-1] * delt)


def diff_surface(mu_delt, nu, e_n1, f_delt_n1, dflux_datmos, flux, factor, delta_xi):
    """vert_diff.F90:866-888; returns (flux, delta_xi)."""
    fff = 1.0 / factor
    dflux = -nu * (1.0 - e_n1)
    delta_xi = delta_xi + mu_delt * nu * f_delt_n1
    delta_xi = (delta_xi + mu_delt * flux * fff) / (1.0 - mu_delt * (dflux + dflux_datmos * fff))
    flux = flux + dflux_datmos * delta_xi
    return flux, delta_xi


def vert_diff_up(delt, e, f, delta_xi_n):
    """vert_diff.F90:892-947 (no kbot)"""
    K = e.shape[0]
    out = np.zeros_like(e)
    out[K - 1] = delta_xi_n / delt
    for k in range(K - 2, -1, -1):
        out[k] = e[k] * out[k + 1] + f[k]
    return out


def gcm_vert_diff_down(delt, u, v, t, q, diff_m, diff_t, p_half, p_full, z_full, tau_u, tau_v, dtau_du, dtau_dv,
                       dt_u, dt_v, dt_t, dt_q, do_conserve_energy=True, use_virtual_temp=False):
    """vert_diff.F90:270-402 with sphum the only tracer (diffused with temperature by vert_diff_down_2).
    Returns the updated dt_u, dt_v, dt_t, tau_u, tau_v, dissipative_heat and the module state / Tri_surf dict."""
    gcp = GRAV / CP_AIR
    tt = t + z_full * gcp
    mu = compute_mu(p_half)
    nu = compute_nu(diff_m, p_half, p_full, z_full, t, q, use_virtual_temp)
    # uv_vert_diff :558-617
    r = vert_diff_down_2(delt, mu, nu, u, v, dt_u, dt_v)
    tau_u, delta_u_n = diff_surface(r["mu_delt_n"], r["nu_n"], r["e_n1"], r["f_1_delt_n1"], dtau_du, tau_u, 1.0, r["delta_1_n"])
    tau_v, delta_v_n = diff_surface(r["mu_delt_n"], r["nu_n"], r["e_n1"], r["f_2_delt_n1"], dtau_dv, tau_v, 1.0, r["delta_2_n"])
    new_u = vert_diff_up(delt, r["e"], r["f_1"], delta_u_n)
    new_v = vert_diff_up(delt, r["e"], r["f_2"], delta_v_n)
    if do_conserve_energy:
        du, dv = new_u - dt_u, new_v - dt_v
        heat = -(1.0 / CP_AIR) * ((u + 0.5 * delt * du) * du + (v + 0.5 * delt * dv) * dv)
        dt_t = dt_t + heat
    else:
        heat = np.zeros_like(t)
    nu = compute_nu(diff_t, p_half, p_full, z_full, t, q, use_virtual_temp)
    r = vert_diff_down_2(delt, mu, nu, tt, q, dt_t, dt_q)
    tri = dict(delta_t=r["delta_1_n"] + r["mu_delt_n"] * r["nu_n"] * r["f_1_delt_n1"],
               dflux_t=-r["nu_n"] * (1.0 - r["e_n1"]),
               delta_q=r["delta_2_n"] + r["mu_delt_n"] * r["nu_n"] * r["f_2_delt_n1"],
               dflux_q=-r["nu_n"] * (1.0 - r["e_n1"]),
               dtmass=r["mu_delt_n"], delta_u=delta_u_n, delta_v=delta_v_n,
               e_global=r["e"], f_t_global=r["f_1"], f_q_global=r["f_2"])
    return dict(dt_u=new_u, dt_v=new_v, dt_t=dt_t, tau_u=tau_u, tau_v=tau_v, dissipative_heat=heat, tri=tri)


def mixed_layer(tri, dt, t_surf, flux_t, flux_q, flux_r, net_surf_sw_down, surf_lw_down, dhdt_surf, dedt_surf, dedq_surf,
                drdt_surf, dhdt_atm, dedq_atm, heat_capacity, ocean_qflux, evaporation=True, sst_new=None):
    """mixed_layer.F90:568-745: the slab update (do_calc_eff_heat_cap path; no ice, no flux anomalies) or, with sst_new, do_sc_sst
    (:681-691 with do_calc_eff_heat_cap = .false., :495-502): t_surf moves to the prescribed SST of the time stepped to.
    Returns the new t_surf and the updated Tri_surf delta_t / delta_q."""
    inv_cp = 1.0 / CP_AIR
    gamma_t = 1.0 / (1.0 - tri["dtmass"] * (tri["dflux_t"] + dhdt_atm * inv_cp))
    gamma_q = 1.0 / (1.0 - tri["dtmass"] * (tri["dflux_q"] + dedq_atm))
    fn_t = gamma_t * (tri["delta_t"] + tri["dtmass"] * flux_t * inv_cp)
    fn_q = gamma_q * (tri["delta_q"] + tri["dtmass"] * flux_q)
    en_t = gamma_t * tri["dtmass"] * dhdt_surf * inv_cp
    en_q = gamma_q * tri["dtmass"] * dedt_surf
    alpha_t = flux_t * inv_cp + dhdt_atm * inv_cp * fn_t
    alpha_q = flux_q + dedq_atm * fn_q
    alpha_lw = flux_r
    beta_t = dhdt_surf * inv_cp + dhdt_atm * inv_cp * en_t
    beta_q = dedt_surf + dedq_atm * en_q
    beta_lw = drdt_surf
    corrected_flux = -net_surf_sw_down - surf_lw_down + alpha_t * CP_AIR + alpha_lw - ocean_qflux
    t_surf_dependence = beta_t * CP_AIR + beta_lw
    if evaporation:
        corrected_flux = corrected_flux + alpha_q * HLV
        t_surf_dependence = t_surf_dependence + beta_q * HLV
    if sst_new is not None:
        delta_t_surf = sst_new - t_surf
    else:
        eff = heat_capacity + t_surf_dependence * dt
        delta_t_surf = -corrected_flux * dt / eff
    out = dict(tri)
    out["delta_t"] = fn_t + en_t * delta_t_surf
    if evaporation:
        out["delta_q"] = fn_q + en_q * delta_t_surf
    return t_surf + delta_t_surf, out, delta_t_surf


def gcm_vert_diff_up(delt, tri):
    """vert_diff.F90:406-467 -> dt_t, dt_q"""
    return (vert_diff_up(delt, tri["e_global"], tri["f_t_global"], tri["delta_t"]),
            vert_diff_up(delt, tri["e_global"], tri["f_q_global"], tri["delta_q"]))


# ------------------------------------------------------------------------------------------------------------------
# Monin-Obukhov similarity (atmos_param/monin_obukhov/monin_obukhov_kernel.F90) and the bulk surface fluxes
# (coupler/surface_flux.F90:338-700).  PINNED against the reference's own known-answer self-test
# (monin_obukhov_kernel.F90:905-1120, CHKSUM_DRAG / CHKSUM_STABLE_MIX / CHKSUM_DIFF / CHKSUM_PROFILE).
VONKARM = 0.40


@dataclass
class MOConfig:
    """monin_obukhov_nml defaults (monin_obukhov.F90:76-83)"""
    rich_crit: float = 2.0
    drag_min: float = 1.0e-05
    neutral: bool = False
    stable_option: int = 1
    zeta_trans: float = 0.5


def mo_derivative_m(c: MOConfig, zeta):
    """monin_obukhov_derivative_m :456-494"""
    b_stab = 1.0 / c.rich_crit
    uns = zeta < 0.0
    zu = np.where(uns, zeta, 0.0)
    zs = np.where(uns, 0.0, zeta)
    x = (1 - 16.0 * zu) ** (-0.5)
    phi_u = np.sqrt(x)
    if c.stable_option == 1:
        phi_s = 1.0 + zs * (5.0 + b_stab * zs) / (1.0 + zs)
    else:
        lam = 1.0 + (5.0 - b_stab) * c.zeta_trans
        phi_s = np.where(zs < c.zeta_trans, 1 + 5.0 * zs, lam + b_stab * zs)
    return np.where(uns, phi_u, phi_s)


def mo_derivative_t(c: MOConfig, zeta):
    """monin_obukhov_derivative_t :415-452"""
    b_stab = 1.0 / c.rich_crit
    uns = zeta < 0.0
    zu = np.where(uns, zeta, 0.0)
    zs = np.where(uns, 0.0, zeta)
    phi_u = (1 - 16.0 * zu) ** (-0.5)
    if c.stable_option == 1:
        phi_s = 1.0 + zs * (5.0 + b_stab * zs) / (1.0 + zs)
    else:
        lam = 1.0 + (5.0 - b_stab) * c.zeta_trans
        phi_s = np.where(zs < c.zeta_trans, 1 + 5.0 * zs, lam + b_stab * zs)
    return np.where(uns, phi_u, phi_s)


def _psi_stable(c, ln, zeta, zeta_0):
    b_stab = 1.0 / c.rich_crit
    if c.stable_option == 1:
        return ln + (5.0 - b_stab) * np.log((1.0 + zeta) / (1.0 + zeta_0)) + b_stab * (zeta - zeta_0)
    lam = 1.0 + (5.0 - b_stab) * c.zeta_trans
    weak = zeta <= c.zeta_trans
    with np.errstate(divide="ignore", invalid="ignore"):
        x = (lam - 1.0) * np.log(np.where(weak, c.zeta_trans, zeta) / c.zeta_trans) + b_stab * (zeta - c.zeta_trans)
    strong = np.where(zeta_0 <= c.zeta_trans, ln + x + 5.0 * (c.zeta_trans - zeta_0), lam * ln + b_stab * (zeta - zeta_0))
    return np.where(weak, ln + 5.0 * (zeta - zeta_0), strong)


def mo_integral_m(c: MOConfig, zeta, zeta_0, ln_z_z0):
    """monin_obukhov_integral_m :644-715"""
    uns = zeta < 0.0
    zu, z0u = np.where(uns, zeta, 0.0), np.where(uns, zeta_0, 0.0)
    x = np.sqrt(np.sqrt(1 - 16.0 * zu))
    x_0 = np.sqrt(np.sqrt(1 - 16.0 * z0u))
    x1, x1_0 = 1.0 + x, 1.0 + x_0
    num = x1 * x1 * (1.0 + x * x)
    denom = x1_0 * x1_0 * (1.0 + x_0 * x_0)
    y = np.arctan(x) - np.arctan(x_0)
    psi_u = ln_z_z0 - np.log(num / denom) + 2 * y
    psi_s = _psi_stable(c, ln_z_z0, np.where(uns, 0.0, zeta), np.where(uns, 0.0, zeta_0))
    return np.where(uns, psi_u, psi_s)


def mo_integral_tq(c: MOConfig, zeta, zeta_t, zeta_q, ln_z_zt, ln_z_zq):
    """monin_obukhov_integral_tq :719-806"""
    uns = zeta < 0.0
    zu = np.where(uns, zeta, 0.0)
    x = np.sqrt(1 - 16.0 * zu)
    x_t = np.sqrt(1 - 16.0 * np.where(uns, zeta_t, 0.0))
    x_q = np.sqrt(1 - 16.0 * np.where(uns, zeta_q, 0.0))
    pt_u = ln_z_zt - 2.0 * np.log((1.0 + x) / (1.0 + x_t))
    pq_u = ln_z_zq - 2.0 * np.log((1.0 + x) / (1.0 + x_q))
    zs = np.where(uns, 0.0, zeta)
    pt_s = _psi_stable(c, ln_z_zt, zs, np.where(uns, 0.0, zeta_t))
    pq_s = _psi_stable(c, ln_z_zq, zs, np.where(uns, 0.0, zeta_q))
    return np.where(uns, pt_u, pt_s), np.where(uns, pq_u, pq_s)


def mo_solve_zeta(c: MOConfig, rich, z, z0, zt, zq, mask, error=1e-4, zeta_min=1e-6, max_iter=20):
    """monin_obukhov_solve_zeta :245-411 (Newton iteration on zeta, per-point freeze once converged)."""
    z_z0, z_zt, z_zq = z / z0, z / zt, z / zq
    ln_z_z0, ln_z_zt, ln_z_zq = np.log(z_z0), np.log(z_zt), np.log(z_zq)
    f_m, f_t, f_q = np.zeros_like(rich), np.zeros_like(rich), np.zeros_like(rich)
    corr = np.zeros_like(rich)
    mask_1 = mask.copy()
    zeta = np.where(mask_1, rich * ln_z_z0 * ln_z_z0 / ln_z_zt, 0.0)
    zeta = np.where(mask_1 & (rich >= 0.0), zeta / (1.0 - rich / c.rich_crit), zeta)
    for _ in range(max_iter):
        small_z = mask_1 & (np.abs(zeta) < zeta_min)
        zeta = np.where(small_z, 0.0, zeta)
        f_m = np.where(small_z, ln_z_z0, f_m); f_t = np.where(small_z, ln_z_zt, f_t); f_q = np.where(small_z, ln_z_zq, f_q)
        mask_1 = mask_1 & ~small_z
        zs = np.where(mask_1, zeta, 1.0)
        rzeta = 1.0 / zs
        zeta_0 = np.where(mask_1, zeta / z_z0, 0.0)
        zeta_t = np.where(mask_1, zeta / z_zt, 0.0)
        zeta_q = np.where(mask_1, zeta / z_zq, 0.0)
        zm = np.where(mask_1, zeta, 0.0)
        phi_m, phi_m_0 = mo_derivative_m(c, zm), mo_derivative_m(c, zeta_0)
        phi_t, phi_t_0 = mo_derivative_t(c, zm), mo_derivative_t(c, zeta_t)
        fm_new = mo_integral_m(c, zm, zeta_0, ln_z_z0)
        ft_new, fq_new = mo_integral_tq(c, zm, zeta_t, zeta_q, ln_z_zt, ln_z_zq)
        f_m = np.where(mask_1, fm_new, f_m); f_t = np.where(mask_1, ft_new, f_t); f_q = np.where(mask_1, fq_new, f_q)
        df_m = (phi_m - phi_m_0) * rzeta
        df_t = (phi_t - phi_t_0) * rzeta
        with np.errstate(divide="ignore", invalid="ignore"):
            rich_1 = zs * f_t / (f_m * f_m)
            d_rich = rich_1 * (rzeta + df_t / f_t - 2.0 * df_m / f_m)
            correction = (rich - rich_1) / d_rich
            cnew = np.minimum(np.abs(correction), np.abs(correction / zs))
        corr = np.where(mask_1, cnew, corr)
        if corr.max(initial=0.0) > error:
            mask_1 = mask_1 & (corr > error)
            zeta = np.where(mask_1, zeta + correction, zeta)
        else:
            break
    return f_m, f_t, f_q


def mo_drag(c: MOConfig, pt, pt0, z, z0, zt, zq, speed, small=1e-4):
    """monin_obukhov_drag_1d :122-241 -> drag_m, drag_t, drag_q, u_star, b_star"""
    r_crit = 0.95 * c.rich_crit
    sqrt_drag_min = np.sqrt(c.drag_min) if c.drag_min != 0.0 else 0.0
    delta_b = GRAV * (pt0 - pt) / pt0
    rich = -z * delta_b / (speed * speed + small)
    zz = np.maximum(np.maximum(z, z0), np.maximum(zt, zq))
    if c.neutral:
        us, bs, qs = VONKARM / np.log(zz / z0), VONKARM / np.log(zz / zt), VONKARM / np.log(zz / zq)
        return us * us, us * bs, us * qs, us * speed, bs * delta_b
    m1 = rich < r_crit
    fm, ft, fq = mo_solve_zeta(c, rich, zz, z0, zt, zq, m1)
    with np.errstate(divide="ignore", invalid="ignore"):
        us = np.where(m1, np.maximum(VONKARM / fm, sqrt_drag_min), sqrt_drag_min)
        bs = np.where(m1, np.maximum(VONKARM / ft, sqrt_drag_min), sqrt_drag_min)
        qs = np.where(m1, np.maximum(VONKARM / fq, sqrt_drag_min), sqrt_drag_min)
    drag_m = np.where(m1, us * us, c.drag_min)
    drag_t = np.where(m1, us * bs, c.drag_min)
    drag_q = np.where(m1, us * qs, c.drag_min)
    return drag_m, drag_t, drag_q, us * speed, bs * delta_b


def mo_profile(c: MOConfig, zref, zref_t, z, z0, zt, zq, u_star, b_star):
    """monin_obukhov_profile_1d :498-640 -> del_m, del_t, del_q"""
    ln_z_z0, ln_z_zt, ln_z_zq = np.log(z / z0), np.log(z / zt), np.log(z / zq)
    ln_z_zref, ln_z_zref_t = np.log(z / zref), np.log(z / zref_t)
    if c.neutral:
        return 1.0 - ln_z_zref / ln_z_z0, 1.0 - ln_z_zref_t / ln_z_zt, 1.0 - ln_z_zref_t / ln_z_zq
    pos = u_star > 0.0
    mo_length_inv = np.where(pos, -VONKARM * b_star / np.where(pos, u_star * u_star, 1.0), 0.0)
    zeta, zeta_0, zeta_t, zeta_q = z * mo_length_inv, z0 * mo_length_inv, zt * mo_length_inv, zq * mo_length_inv
    zeta_ref, zeta_ref_t = zref * mo_length_inv, zref_t * mo_length_inv
    f_m = mo_integral_m(c, zeta, zeta_0, ln_z_z0)
    f_m_ref = mo_integral_m(c, zeta, zeta_ref, ln_z_zref)
    f_t, f_q = mo_integral_tq(c, zeta, zeta_t, zeta_q, ln_z_zt, ln_z_zq)
    f_t_ref, f_q_ref = mo_integral_tq(c, zeta, zeta_ref_t, zeta_ref_t, ln_z_zref_t, ln_z_zref_t)
    return 1.0 - f_m_ref / f_m, 1.0 - f_t_ref / f_t, 1.0 - f_q_ref / f_q


def mo_stable_mix(c: MOConfig, rich):
    """monin_obukhov_stable_mix :810-868"""
    b_stab = 1.0 / c.rich_crit
    mix = np.zeros_like(rich)
    if c.stable_option == 1:
        ok = (rich > 0.0) & (rich < c.rich_crit)
        r = 1.0 / np.where(ok, rich, 1.0)
        a = r - b_stab
        b = r - (1.0 + 5.0)
        zeta = (-b + np.sqrt(b * b - 4.0 * a * (-1.0))) / (2.0 * a)
        phi = 1.0 + b_stab * zeta + (5.0 - b_stab) * zeta / (1.0 + zeta)
        return np.where(ok, 1.0 / (phi * phi), 0.0)
    rich_trans = c.zeta_trans / (1.0 + 5.0 * c.zeta_trans)
    lam = 1.0 + (5.0 - b_stab) * c.zeta_trans
    mix = np.where((rich > 0.0) & (rich <= rich_trans), (1.0 - 5.0 * rich) ** 2, mix)
    mix = np.where((rich > rich_trans) & (rich < c.rich_crit), ((1.0 - b_stab * rich) / lam) ** 2, mix)
    return mix


def mo_diff(c: MOConfig, z, u_star, b_star, ustar_min=1e-10):
    """monin_obukhov_diff :35-118; z [K,...], u_star/b_star [...] -> k_m, k_h"""
    uss = np.maximum(u_star, ustar_min)
    if c.neutral:
        k_m = VONKARM * uss * z
        return k_m, k_m.copy()
    zeta = -VONKARM * b_star * z / (uss * uss)
    return VONKARM * uss * z / mo_derivative_m(c, zeta), VONKARM * uss * z / mo_derivative_t(c, zeta)


@dataclass
class SurfaceFluxConfig:
    """surface_flux_nml defaults (surface_flux.F90:225-239); bucket, ncar_ocean_flux, raoult_sat_vap not restated."""
    no_neg_q: bool = False
    use_virtual_temp: bool = True
    alt_gustiness: bool = False
    old_dtaudv: bool = False
    use_mixing_ratio: bool = False
    gust_const: float = 1.0
    gust_min: float = 0.0
    do_simple: bool = False
    land_humidity_prefactor: float = 1.0
    land_evap_prefactor: float = 1.0


def surface_flux(svp: SatVaporPres, mo: MOConfig, c: SurfaceFluxConfig, t_atm, q_atm_in, u_atm, v_atm, p_atm, z_atm, p_surf,
                 t_surf, t_ca, q_surf, u_surf, v_surf, rough_mom, rough_heat, rough_moist, rough_scale, gust, land):
    """surface_flux_1d (surface_flux.F90:338-700), bucket = .false., every point available.  Returns a dict of the outputs."""
    del_temp = 0.1; del_temp_inv = 1.0 / del_temp
    d622 = RDGAS / RVGAS; d378 = 1.0 - d622
    d608 = d378 / d622 if c.use_virtual_temp else 0.0
    kappa = RDGAS / CP_AIR
    t_surf0 = np.where(land, t_ca, t_surf)
    t_surf1 = t_surf0 + del_temp
    e_sat, _ = svp.lookup_es_des(t_surf0)
    e_sat1, _ = svp.lookup_es_des(t_surf1)
    if c.use_mixing_ratio:
        q_sat, q_sat1 = d622 * e_sat / (p_surf - e_sat), d622 * e_sat1 / (p_surf - e_sat1)
    elif c.do_simple:
        q_sat, q_sat1 = d622 * e_sat / p_surf, d622 * e_sat1 / p_surf
    else:
        q_sat, q_sat1 = d622 * e_sat / (p_surf - d378 * e_sat), d622 * e_sat1 / (p_surf - d378 * e_sat1)
    q_surf0 = q_sat
    q_atm = np.where(q_atm_in < 0.0, 0.0, q_atm_in) if c.no_neg_q else q_atm_in
    p_ratio = (p_surf / p_atm) ** kappa
    tv_atm = t_atm * (1.0 + d608 * q_atm)
    th_atm = t_atm * p_ratio
    thv_atm = tv_atm * p_ratio
    thv_surf = t_surf0 * (1.0 + d608 * q_surf0)
    u_dif, v_dif = u_surf - u_atm, v_surf - v_atm
    if c.alt_gustiness:
        w_atm = np.maximum(np.sqrt(u_dif ** 2 + v_dif ** 2), c.gust_const)
        big = w_atm > c.gust_const
        dw_atmdu, dw_atmdv = np.where(big, u_dif / w_atm, 0.0), np.where(big, v_dif / w_atm, 0.0)
    else:
        w_gust = np.maximum(gust, c.gust_min) if c.gust_min > 0.0 else gust
        w_atm = np.sqrt(u_dif * u_dif + v_dif * v_dif + w_gust * w_gust)
        dw_atmdu, dw_atmdv = u_dif / w_atm, v_dif / w_atm
    cd_m, cd_t, cd_q, u_star, b_star = mo_drag(mo, thv_atm, thv_surf, z_atm, rough_mom, rough_heat, rough_moist, w_atm)
    ex_del_m, ex_del_h, ex_del_q = mo_profile(mo, 10.0, 2.0, z_atm, rough_mom, rough_heat, rough_moist, u_star, b_star)
    temp_2m = t_surf + (t_atm - t_surf) * ex_del_h
    u_10m, v_10m = u_atm * ex_del_m, v_atm * ex_del_m
    q_2m = q_surf + (q_atm - q_surf) * ex_del_q                       # q_surf: the value on entry (intent inout)
    e_sat_2m, _ = svp.lookup_es_des(temp_2m)
    if c.use_mixing_ratio:
        q_sat_2m = d622 * e_sat_2m / (p_surf - e_sat_2m)
    elif c.do_simple:
        q_sat_2m = d622 * e_sat_2m / p_surf
    else:
        q_sat_2m = d622 * e_sat_2m / (p_surf - d378 * e_sat)          # sic: e_sat, as in the reference (:550)
    rh_2m = q_2m / q_sat_2m
    cd_m = cd_m * (np.log(z_atm / rough_mom + 1) / np.log(z_atm / rough_scale + 1)) ** 2
    drag_t, drag_q, drag_m = cd_t * w_atm, cd_q * w_atm, cd_m * w_atm
    rho = p_atm / (RDGAS * tv_atm)
    rho_drag = CP_AIR * drag_t * rho
    flux_t = rho_drag * (t_surf0 - th_atm)
    dhdt_surf = rho_drag
    dhdt_atm = -rho_drag * p_ratio
    rho_drag = drag_q * rho
    flux_q = np.where(land, rho_drag * c.land_evap_prefactor * (c.land_humidity_prefactor * q_surf0 - q_atm), rho_drag * (q_surf0 - q_atm))
    dedt_surf = np.where(land, rho_drag * c.land_evap_prefactor * (c.land_humidity_prefactor * q_sat1 - q_sat) * del_temp_inv,
                         rho_drag * (q_sat1 - q_sat) * del_temp_inv)
    dedq_surf = np.zeros_like(flux_q)
    dedq_atm = -rho_drag
    q_star = flux_q / (u_star * rho)
    q_surf_out = q_atm + flux_q / (rho * cd_q * w_atm)
    flux_r = STEFAN * t_surf ** 4
    drdt_surf = 4 * STEFAN * t_surf ** 3
    rho_drag = drag_m * rho
    flux_u, flux_v = rho_drag * u_dif, rho_drag * v_dif
    if c.old_dtaudv:
        dtaudu_atm = dtaudv_atm = -rho_drag
    else:
        dtaudu_atm = -cd_m * rho * (dw_atmdu * u_dif + w_atm)
        dtaudv_atm = -cd_m * rho * (dw_atmdv * v_dif + w_atm)
    return dict(flux_t=flux_t, flux_q=flux_q, flux_r=flux_r, flux_u=flux_u, flux_v=flux_v, cd_m=cd_m, cd_t=cd_t, cd_q=cd_q,
                w_atm=w_atm, u_star=u_star, b_star=b_star, q_star=q_star, dhdt_surf=dhdt_surf, dedt_surf=dedt_surf,
                dedq_surf=dedq_surf, drdt_surf=drdt_surf, dhdt_atm=dhdt_atm, dedq_atm=dedq_atm, dtaudu_atm=dtaudu_atm,
                dtaudv_atm=dtaudv_atm, ex_del_m=ex_del_m, ex_del_h=ex_del_h, ex_del_q=ex_del_q, temp_2m=temp_2m, u_10m=u_10m,
                v_10m=v_10m, q_2m=q_2m, rh_2m=rh_2m, q_surf=q_surf_out)


# ------------------------------------------------------------------------------------------------------------------
# K-profile boundary layer diffusivities (atmos_param/diffusivity/diffusivity.F90:263-530, 732-750), as called by
# vert_turb_driver with do_diffusivity = .true. (vert_turb_driver.F90:277-292)
@dataclass
class DiffusivityConfig:
    """diffusivity_nml defaults (diffusivity.F90:124-147); pbl_mcm, use_pog_bug_fix=.false. not restated."""
    fixed_depth: bool = False
    depth_0: float = 5000.0
    frac_inner: float = 0.1
    rich_crit_pbl: float = 1.0
    entr_ratio: float = 0.2
    parcel_buoy: float = 2.0
    znom: float = 1000.0
    background_m: float = 0.0
    background_t: float = 0.0
    do_entrain: bool = True
    do_simple: bool = False
    free_atm_diff: bool = False
    free_atm_skyhi_diff: bool = False
    rich_crit_diff: float = 0.25
    mix_len: float = 30.0
    rich_prandtl: float = 1.0
    ampns: bool = False
    ampns_max: float = 1.0e20


def diffusivity_free(c: DiffusivityConfig, t, u, v, z, zz, h, k_m, k_t, small=1e-4):
    """diffusivity.F90:604-697: Richardson-number mixing-length diffusivities above the boundary layer (zz > h) overwrite the
    boundary-layer values where the local Richardson number is sub-critical.  t = (virtual) dry static energy / cp."""
    K = t.shape[0]
    k_m, k_t = k_m.copy(), k_t.copy()
    for k in range(1, K):
        dz = z[k - 1] - z[k]
        b = GRAV * (t[k - 1] - t[k]) / t[k]
        speed2 = (u[k - 1] - u[k]) ** 2 + (v[k - 1] - v[k]) ** 2
        rich = b * dz / (speed2 + small)
        rich = np.maximum(rich, 0.0)
        fri2 = None
        if c.free_atm_skyhi_diff:
            fri2 = np.where(rich >= c.rich_crit_diff, 0.0, (1.0 - rich / c.rich_crit_diff) ** 2)
        if c.ampns:
            alpz = np.minimum(1.0 + 1.0e-04 * (dz ** 1.5), c.ampns_max)
            rich = rich / alpz
        fri = (1.0 - rich / c.rich_crit_diff) ** 2
        m = (rich < c.rich_crit_diff) & (zz[k] > h)
        if c.free_atm_skyhi_diff:
            if c.ampns:
                km = c.mix_len * c.mix_len * np.sqrt(speed2) * fri * (1.0 + 1.0e-04 * (dz ** 1.5)) / dz
            else:
                km = c.mix_len * c.mix_len * np.sqrt(speed2) * fri / dz
            kt = km * (0.1 + 0.9 * fri2)
        else:
            kt = c.mix_len * c.mix_len * np.sqrt(speed2) * fri / dz
            km = kt * c.rich_prandtl
        k_m[k] = np.where(m, km, k_m[k]); k_t[k] = np.where(m, kt, k_t[k])
    return k_m, k_t


def pbl_depth(c: DiffusivityConfig, mo: MOConfig, t, u, v, z, u_star, b_star, small=1e-4):
    """diffusivity.F90:358-456 (no kbot); t is the (virtual) dry static energy / cp."""
    K = t.shape[0]
    tbot = t[K - 1]
    rich = z * GRAV * (t - tbot[None]) / tbot[None] / (u * u + v * v + small)
    h_inner = np.full_like(tbot, c.frac_inner * c.znom)
    ws, _ = mo_diff(mo, h_inner[None], u_star, b_star)
    ws = np.maximum(small, ws[0] / VONKARM / h_inner)
    h = z[K - 1].copy()
    done = np.zeros(tbot.shape, bool)
    rich_branch = (b_star <= 0.0) | c.do_simple
    # Richardson-number search
    h1, r1 = z[K - 1].copy(), rich[K - 1].copy()
    for k in range(K - 2, -1, -1):
        r2, h2 = rich[k], z[k]
        hit = rich_branch & ~done & (r2 > c.rich_crit_pbl)
        with np.errstate(divide="ignore", invalid="ignore"):
            h = np.where(hit, h2 + (h1 - h2) * (r2 - c.rich_crit_pbl) / (r2 - r1), h)
        done |= hit
        r1, h1 = r2, h2
    # parcel search
    with np.errstate(divide="ignore", invalid="ignore"):
        svp = tbot * (1.0 + (c.parcel_buoy * u_star * b_star / GRAV / ws))
    h1, t1 = z[K - 1].copy(), tbot.copy()
    for k in range(K - 2, -1, -1):
        h2, t2 = z[k], t[k]
        hit = ~rich_branch & ~done & (t2 > svp)
        with np.errstate(divide="ignore", invalid="ignore"):
            h = np.where(hit, h2 + (h1 - h2) * (t2 - svp) / (t2 - t1), h)
        done |= hit
        h1, t1 = h2, t2
    return h


def diffusivity(c: DiffusivityConfig, mo: MOConfig, t, q, u, v, p_full, p_half, z_full, z_half, u_star, b_star, k_m, k_t, small=1e-4):
    """diffusivity.F90:263-354 -> h, k_m, k_t (k_m, k_t in: values to be added, as vert_turb_driver passes zeros)."""
    K = t.shape[0]
    gcp = GRAV / CP_AIR
    z_surf = z_half[K]
    z_full_ag = z_full - z_surf[None]
    z_half_ag = z_half - z_surf[None]
    svcp = (t + gcp * z_full_ag) if c.do_simple else (t * (1.0 + D608 * q) + gcp * z_full_ag)
    if c.fixed_depth:
        h = np.full_like(u_star, c.depth_0)
    else:
        h = pbl_depth(c, mo, svcp, u, v, z_full_ag, u_star, b_star)
    # diffusivity_pbl :458-526 (use_pog_bug_fix = .true.: the result is a per-column function)
    zm = z_half_ag
    h_inner = c.frac_inner * h
    km_ref, kt_ref = mo_diff(mo, h_inner[None], u_star, b_star)
    km_ref, kt_ref = km_ref[0], kt_ref[0]
    km_sl, kt_sl = mo_diff(mo, np.maximum(zm[:K], 0.0), u_star, b_star)
    new_m, new_t = np.zeros_like(t), np.zeros_like(t)
    for k in range(1, K):
        inner = zm[k] < h_inner
        mid = (zm[k] >= h_inner) & (zm[k] < h)
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = (zm[k] / h_inner) * (1.0 - (zm[k] - h_inner) / (h - h_inner)) ** 2
        new_m[k] = np.where(mid, km_ref * factor, np.where(inner, km_sl[k], 0.0))
        new_t[k] = np.where(mid, kt_ref * factor, np.where(inner, kt_sl[k], 0.0))
    if c.free_atm_diff:
        new_m, new_t = diffusivity_free(c, svcp, u, v, z_full_ag, z_half_ag, h, new_m, new_t, small)
    k_m, k_t = new_m + k_m, new_t + k_t
    if c.entr_ratio > 0.0 and not c.fixed_depth and c.do_entrain:                      # diffusivity_entr :732-750
        for k in range(1, K):
            m = (b_star > 0.0) & (z_full_ag[k - 1] > h) & (z_full_ag[k] <= h)
            val = (z_full_ag[k - 1] - z_full_ag[k]) * c.entr_ratio * svcp[k] * u_star * b_star / GRAV / np.maximum(small, svcp[k - 1] - svcp[k])
            k_t[k] = np.where(m, val, k_t[k]); k_m[k] = np.where(m, val, k_m[k])
    if c.background_m > 0.0:
        k_m = np.maximum(k_m, c.background_m)
    if c.background_t > 0.0:
        k_t = np.maximum(k_t, c.background_t)
    return h, k_m, k_t


# ------------------------------------------------------------------------------------------------------------------
# Simplified Betts-Miller convection (atmos_param/qe_moist_convection/qe_moist_convection.F90:77-1179).
# Written column by column with scalar loops exactly as the reference (small test sizes only).
import math


class SBMConvection:
    """qe_moist_convection_mod: namelist tau_bm, rhbm, Tmin, Tmax, val_inc (:61-75) + the LCL temperature table (:113-154)."""
    SMALL = 1.0e-10
    PREF = 1.0e5

    def __init__(self, svp: SatVaporPres, tau_bm=7200.0, rhbm=0.8, Tmin=173.0, Tmax=335.0, val_inc=0.01):
        self.svp, self.tau_bm, self.rhbm, self.Tmin, self.Tmax, self.val_inc = svp, tau_bm, rhbm, Tmin, Tmax, val_inc
        self.val_min = math.log(self.es(Tmin) / (Tmin ** (1.0 / KAPPA)))
        self.val_max = math.log(self.es(Tmax) / (Tmax ** (1.0 / KAPPA)))
        n = int(math.ceil((self.val_max - self.val_min) / val_inc))
        tab, guess = np.zeros(n), Tmin
        for k in range(n):
            tab[k] = self._lcl_temp(self.val_min + k * val_inc, guess)
            guess = tab[k]
        self.lcl_temp_table = tab

    def es(self, T):
        s = self.svp
        tmp = T - s.tminl
        ind = int(s.dtinvl * (tmp + s.tepsl))
        if ind < 0 or ind >= s.table_siz or not (s.dtinvl * (tmp + s.tepsl) > -1.0):
            raise FloatingPointError("escomp: temperature out of the table range")
        dl = tmp - s.dtres * ind
        return s.TABLE[ind] + dl * (s.DTABLE[ind] + dl * s.D2TABLE[ind])

    def _lcl_temp(self, value, guess):
        """lcl_temp :1090-1117 (Newton)"""
        T, dT, it = guess, 1.0e-7 + 1.0, 0
        while abs(dT) > 1.0e-7 and it < 100:
            f = value - math.log(self.es(T) * T ** (-1 / KAPPA))
            df = 1 / KAPPA * T ** (-1) - HLV / RVGAS * T ** (-2)
            dT = f / df
            T = T - dT
            it += 1
        if not dT < 1.0e-7:
            raise FloatingPointError("qe_moist_convection: LCL calculation did not converge")
        return T

    def get_lcl_temp(self, value):
        """:1053-1082"""
        if value < self.val_min:
            raise FloatingPointError("get_lcl_temp: value to low.")
        if value > self.val_max:
            raise FloatingPointError("get_lcl_temp: value too high.")
        iv = int(math.floor((value - self.val_min) / self.val_inc)) + 1
        if iv + 1 > self.lcl_temp_table.size:
            raise FloatingPointError("get_lcl_temp: value too high.")     # the reference reads past the table here
        w_floor = self.val_min + (iv - 1) * self.val_inc
        w_ceil = (value - w_floor) / self.val_inc
        return self.lcl_temp_table[iv] * w_ceil - self.lcl_temp_table[iv - 1] * (w_ceil - 1)

    @staticmethod
    def mixing_ratio(e, p):
        return RDGAS * e / RVGAS / (p - e)

    @staticmethod
    def virtual_temp(T, r):
        q = r / (1.0 + r)
        return T * (1.0 + q * (RVGAS / RDGAS - 1.0))

    def cape_calculation(self, pf, ph, Tin, rin):
        """CAPE_calculation / CAPE_below_LCL / CAPE_above_LCL :373-690.  Levels are 0-based here; kLZB, kLCL are returned
        1-based (0 = not found) as the reference reports them."""
        K = Tin.size
        ks = K - 1
        small = self.SMALL
        Tp, rp = Tin.copy(), rin.copy()
        st = dict(nocape=True, CAPE=0.0, CIN=0.0, kLFC=0, kLZB=0)
        Tv = np.array([self.virtual_temp(Tin[k], rin[k]) for k in range(K)])
        T0, r0 = Tin[ks], rin[ks]
        rs = self.mixing_ratio(self.es(T0), pf[ks])
        saturated = r0 >= rs

        def nocape_reset():
            st["kLZB"] = 0; st["kLFC"] = 0; st["CIN"] = 0.0
            Tp[:] = Tin; rp[:] = rin

        skip = False
        kLCL = 0                                          # 1-based
        if saturated:
            kLCL = K
            Tp[ks] = T0 + (r0 - rs) / ((CP_AIR / (HLV + small)) + (HLV * rs) / RVGAS / T0 ** 2)
            rp[ks] = self.mixing_ratio(self.es(Tp[ks]), pf[ks])
        else:
            theta0 = Tin[ks] * (self.PREF / pf[ks]) ** KAPPA
            if r0 <= 0:
                skip = True
            else:
                value = math.log(theta0 ** (-1 / KAPPA) * self.PREF * r0 / (RDGAS / RVGAS + r0))
                TLCL = self.get_lcl_temp(value)
                pLCL = self.PREF * (TLCL / theta0) ** (1.0 / KAPPA)
                if pLCL < pf[0]:
                    pLCL = pf[0]
                    TLCL = theta0 * (pLCL / self.PREF) ** KAPPA
                k = ks
                st["CIN"] = 0.0
                while pf[k] > pLCL:
                    Tp[k] = theta0 * (pf[k] / self.PREF) ** KAPPA
                    rp[k] = self.mixing_ratio(self.es(Tp[k]), pf[k])
                    st["CIN"] = st["CIN"] + RDGAS * (Tv[k] - self.virtual_temp(Tp[k], r0)) * math.log(ph[k + 1] / ph[k])
                    k -= 1
                kLCL = k + 1
                a = KAPPA * TLCL + (HLV / CP_AIR) * r0
                b = (HLV ** 2) * r0 / (CP_AIR * RVGAS * TLCL ** 2)
                dtdlnp = a / (1.0 + b)
                Tp[k] = TLCL + dtdlnp * math.log(pf[k] / pLCL) / 2
                if Tp[k] < self.Tmin and st["nocape"]:
                    skip = True
                    nocape_reset()
                else:
                    rp[k] = self.mixing_ratio(self.es(Tp[k]), (pf[k] + pLCL) / 2)
                    a = KAPPA * Tp[k] + (HLV / CP_AIR) * rp[k]
                    b = (HLV ** 2) * rp[k] / (CP_AIR * RVGAS * Tp[k] ** 2)
                    dtdlnp = a / (1.0 + b)
                    Tp[k] = TLCL + dtdlnp * math.log(pf[k] / pLCL)
                    if Tp[k] < self.Tmin and st["nocape"]:
                        skip = True
                        nocape_reset()
                    else:
                        rp[k] = self.mixing_ratio(self.es(Tp[k]), pf[k])
                        tvp = self.virtual_temp(Tp[k], rp[k])
                        if tvp < Tv[k] and st["nocape"]:
                            st["CIN"] = st["CIN"] + RDGAS * (Tv[k] - tvp) * math.log(ph[k + 1] / ph[k])
                        else:
                            st["CAPE"] = st["CAPE"] + RDGAS * (tvp - Tv[k]) * math.log(ph[k + 1] / ph[k])
                            if st["nocape"]:
                                st["nocape"] = False
                                st["kLFC"] = k + 1
        # CAPE_above_LCL
        if skip:
            if st["nocape"]:
                nocape_reset()
        else:
            for k in range(kLCL - 2, -1, -1):
                a = KAPPA * Tp[k + 1] + (HLV / CP_AIR) * rp[k + 1]
                b = (HLV ** 2) * rp[k + 1] / (CP_AIR * RVGAS * Tp[k + 1] ** 2)
                dtdlnp = a / (1.0 + b)
                Tp[k] = Tp[k + 1] + dtdlnp * math.log(pf[k] / pf[k + 1]) / 2
                if Tp[k] < self.Tmin and st["nocape"]:
                    nocape_reset()
                    break
                rp[k] = self.mixing_ratio(self.es(Tp[k]), (pf[k] + pf[k + 1]) / 2)
                a = KAPPA * Tp[k] + (HLV / CP_AIR) * rp[k]
                b = (HLV ** 2) * rp[k] / (CP_AIR * RVGAS * Tp[k] ** 2)
                dtdlnp = a / (1.0 + b)
                Tp[k] = Tp[k + 1] + dtdlnp * math.log(pf[k] / pf[k + 1])
                if Tp[k] < self.Tmin and st["nocape"]:
                    nocape_reset()
                    break
                rp[k] = self.mixing_ratio(self.es(Tp[k]), pf[k])
                tvp = self.virtual_temp(Tp[k], rp[k])
                if tvp < Tv[k] and st["nocape"]:
                    st["CIN"] = st["CIN"] + RDGAS * (Tv[k] - tvp) * math.log(ph[k + 1] / ph[k])
                elif tvp < Tv[k] and not st["nocape"]:
                    st["kLZB"] = k + 2
                    break
                else:
                    st["CAPE"] = st["CAPE"] + RDGAS * (tvp - Tv[k]) * math.log(ph[k + 1] / ph[k])
                    if st["nocape"]:
                        st["nocape"] = False
                        st["kLFC"] = k + 1
        return st["kLZB"], kLCL, Tp, rp, st["CAPE"], st["CIN"]

    def column(self, dt, Tin, qin, pf, ph):
        """one column of SBM_convection_scheme :257-369"""
        K = Tin.size
        small = self.SMALL
        rin = qin / (1.0 - qin)
        kLZB, kLCL, Tp, rp, cape, cin = self.cape_calculation(pf, ph, Tin, rin)
        deltaq, deltaT = np.zeros(K), np.zeros(K)
        qref, Tref = np.zeros(K), np.zeros(K)
        convflag, Pq, itq, itt = 0, 0.0, 0.0, 0.0

        def full(k1, k2):                                   # set_profiles_to_full_model_values, 1-based inclusive
            Tref[k1 - 1:k2] = Tin[k1 - 1:k2]; qref[k1 - 1:k2] = qin[k1 - 1:k2]
            deltaT[k1 - 1:k2] = 0.0; deltaq[k1 - 1:k2] = 0.0

        if cape > 0:
            convflag = 1
            lz = max(kLZB, 1)                               # kLZB = 0 (LZB above the model top) indexes out of bounds in the reference
            Tref[:] = Tp                                    # set_reference_profiles :768-796
            for k in range(lz - 1, K):
                eref = self.rhbm * pf[k] * rp[k] / (rp[k] + (RDGAS / RVGAS))
                rp[k] = self.mixing_ratio(eref, pf[k])
                qref[k] = rp[k] / (1 + rp[k])
            full(1, max(lz - 1, 1))
            Pq = 0.0                                        # Pq_calculation :715-736
            for k in range(lz - 1, K):
                deltaq[k] = -(qin[k] - qref[k]) * dt / self.tau_bm
                Pq = Pq + deltaq[k] * (ph[k] - ph[k + 1])
            Pq = Pq / GRAV
            Pt = 0.0                                        # Pt_calculation :739-764
            for k in range(lz - 1, K):
                deltaT[k] = -(Tin[k] - Tref[k]) * dt / self.tau_bm
                Pt = Pt + (CP_AIR / (HLV + small)) * deltaT[k] * (ph[k + 1] - ph[k])
            Pt = Pt / GRAV
            if Pq > 0 and Pt > 0:
                convflag = 2
                if Pq > Pt:                                 # do_change_time_scale_deepconv :1001-1017
                    itq = Pt / Pq / self.tau_bm
                    deltaq[lz - 1:K] = self.tau_bm * itq * deltaq[lz - 1:K]
                    Pq = Pt
                    itt = 1.0 / self.tau_bm
                else:                                       # do_change_Tref_deepconv :968-999
                    deltak = 0.0
                    for k in range(lz - 1, K):
                        deltak = deltak - (deltaT[k] + (HLV / CP_AIR) * deltaq[k]) * (ph[k + 1] - ph[k])
                    deltak = deltak / (ph[K] - ph[lz - 1])
                    for k in range(lz - 1, K):
                        Tref[k] = Tref[k] + deltak * self.tau_bm / dt
                        deltaT[k] = deltaT[k] + deltak
            elif Pt > 0:                                    # do_shallow_convection :800-929
                k = lz
                while Pq < 0.0 and k <= K:                  # level_of_zero_precip
                    Pq = Pq - deltaq[k - 1] * (ph[k - 1] - ph[k]) / GRAV
                    k += 1
                k_top = k - 1
                found = Pq > 0.0
                if k_top > lz:
                    full(lz, k_top - 1)
                if found:                                   # change_Tref_LZB_shallowconv
                    c = Pq * GRAV / (deltaq[k_top - 1] * (ph[k_top] - ph[k_top - 1]))
                    deltaq[k_top - 1] = deltaq[k_top - 1] * c
                    deltaT[k_top - 1] = deltaT[k_top - 1] * c
                    deltak = 0.0
                    for kk in range(k_top - 1, K):
                        deltak = deltak + deltaT[kk] * (ph[kk] - ph[kk + 1])
                    deltak = deltak / (ph[K] - ph[k_top - 1])
                    if k_top != K:
                        deltaT[k_top - 1:K] = deltaT[k_top - 1:K] + deltak
                        Tref[k_top - 1:K] = Tref[k_top - 1:K] + deltak * self.tau_bm / dt
                else:
                    if k_top == lz:
                        full(K, K)
                    else:
                        full(lz, k_top)
                Pq = 0.0
            else:
                Pq = 0.0
                full(1, K)
        else:
            Pq = 0.0
            full(1, K)
        return dict(rain=Pq, deltaT=deltaT, deltaq=deltaq, qref=qref, Tref=Tref, convflag=convflag, kLZB=kLZB, kLCL=kLCL,
                    CAPE=cape, CIN=cin, invtau_q=itq, invtau_t=itt)

    def __call__(self, dt, Tin, qin, p_full, p_half):
        """qe_moist_convection :157-186 over [K, J, I] arrays."""
        K, J, I = Tin.shape
        out = dict(rain=np.zeros((J, I)), snow=np.zeros((J, I)), CAPE=np.zeros((J, I)), CIN=np.zeros((J, I)),
                   deltaT=np.zeros_like(Tin), deltaq=np.zeros_like(Tin), qref=np.zeros_like(Tin), Tref=np.zeros_like(Tin),
                   convflag=np.zeros((J, I), np.int32), kLZBs=np.zeros((J, I), np.int32), kLCLs=np.zeros((J, I), np.int32),
                   invtau_q_relaxation=np.zeros((J, I)), invtau_t_relaxation=np.zeros((J, I)))
        for j in range(J):
            for i in range(I):
                r = self.column(dt, Tin[:, j, i].copy(), qin[:, j, i].copy(), p_full[:, j, i], p_half[:, j, i])
                out["rain"][j, i] = r["rain"]; out["CAPE"][j, i] = r["CAPE"]; out["CIN"][j, i] = r["CIN"]
                for n in ("deltaT", "deltaq", "qref", "Tref"):
                    out[n][:, j, i] = r[n]
                out["convflag"][j, i] = r["convflag"]; out["kLZBs"][j, i] = r["kLZB"]; out["kLCLs"][j, i] = r["kLCL"]
                # the reference zeroes the whole relaxation-rate arrays inside the column loop (:283-284): only the last
                # column processed (i = I, j = J, the loop runs i outer / j inner) keeps its value
                if i == I - 1 and j == J - 1:
                    out["invtau_q_relaxation"][j, i] = r["invtau_q"]; out["invtau_t_relaxation"][j, i] = r["invtau_t"]
        return out


# ------------------------------------------------------------------------------------------------------------------
# idealized_moist_phys dispatcher (atmos_spectral/driver/solo/idealized_moist_phys.F90:819-1395) for the grey-radiation /
# slab-ocean aquaplanet configuration of the Frierson test case (exp/test_cases/frierson/frierson_test_case.py):
# convection_scheme SIMPLE_BETTS_MILLER or NONE, lscale_cond, two_stream_gray, surface_flux, mixed_layer_bc, turb
# (vert_turb_driver with do_diffusivity), gcm_vert_diff; optional rayleigh damping_driver.  No bucket, no land, no clouds.
DENS_H2O = 1000.0
RHO_CP = 1.035e3 * 3989.24495292815


def dry_convection(tg, p_full, p_half, tau, gamma, rdgas=RDGAS, cp_air=CP_AIR):
    """dry_convection + capecalc (atmos_param/dry_convection/dry_convection.f90:105-299): arrays [K, ...] top-down.
    -> dt_tg, cape, cin, lzb, lcl (1-based level indices, btm = K)"""
    K = tg.shape[0]
    shp = tg.shape[1:]
    cons1 = rdgas / cp_air
    btm = K
    tp = tg.copy()
    for k in range(btm - 1, 0, -1):                                  # 1-based k
        zdpkpk = np.exp(cons1 * np.log(p_full[k - 1] / p_full[k]))
        tp[k - 1] = tp[k] + gamma * (tp[k] * zdpkpk - tp[k])
    cape, cin = np.zeros(shp), np.zeros(shp)
    lzb = np.full(shp, btm, dtype=np.int64)
    lcl = np.full(shp, btm, dtype=np.int64)
    for k in range(btm - 1, 0, -1):
        with np.errstate(divide="ignore", invalid="ignore"):         # p_half(1) = 0 in sigma coordinates: log(inf), as in the Fortran
            lg = np.log(p_half[k] / p_half[k - 1])
        unst = tp[k - 1] > tg[k - 1]
        nocloud = lzb == btm
        a = unst & nocloud
        with np.errstate(invalid="ignore"):
            cape = np.where(a, cape + rdgas * (tp[k - 1] - tg[k - 1]) * lg, cape)
        lcl = np.where(a & (tp[k] < tg[k]), k, lcl)
        above = (tp[k - 2] < tg[k - 2]) if k > 1 else np.ones(shp, bool)
        lzb = np.where(a & above, k, lzb)
        tp[k - 1] = np.where(unst & ~nocloud, tg[k - 1], tp[k - 1])
        stab = tp[k - 1] <= tg[k - 1]                                # re-evaluated after the reset above, as the Fortran does
        nocloud = lzb == btm
        with np.errstate(invalid="ignore"):
            cin = np.where(stab & nocloud & (lcl == btm), cin - rdgas * (tp[k - 1] - tg[k - 1]) * lg, cin)
        tp[k - 1] = np.where(stab & ~nocloud, tg[k - 1], tp[k - 1])
    tp = np.where((cin > cape)[None], tg, tp)
    if ((lcl != btm) & (lzb == btm)).any():
        raise FloatingPointError("dry_convection: LCL defined, LZB not defined")
    if (lcl < lzb).any():
        raise FloatingPointError("dry_convection: LCL above LZB")
    none = (lcl == btm) & (lzb == btm)
    cape, cin = np.where(none, 0.0, cape), np.where(none, 0.0, cin)
    ener_int, dp = np.zeros(shp), np.zeros(shp)
    for k in range(1, K + 1):
        inside = (k >= lzb) & (k <= btm)
        dph = p_half[k] - p_half[k - 1]
        ener_int = np.where(inside, ener_int + dph * (tg[k - 1] - tp[k - 1]), ener_int)
        dp = np.where(inside, dp + dph, dp)
        tp[k - 1] = np.where(inside, tp[k - 1], tg[k - 1])
    ener_int = ener_int / dp
    for k in range(btm, 0, -1):
        tp[k - 1] = np.where(k >= lzb, tp[k - 1] + ener_int, tp[k - 1])
    return (tp - tg) / tau, cape, cin, lzb, lcl


@dataclass
class MoistPhysConfig:
    convection_scheme: str = "SIMPLE_BETTS_MILLER"
    do_damping: bool = False
    roughness_mom: float = 0.05
    roughness_heat: float = 0.05
    roughness_moist: float = 0.05
    do_virtual: bool = False
    # mixed_layer_nml
    depth: float = 40.0
    albedo_value: float = 0.06
    evaporation: bool = True
    # lscale_cond_nml
    hc: float = 1.0
    do_evap: bool = False
    # vert_turb_driver_nml
    constant_gust: float = 1.0
    use_tau: bool = True
    # dry_convection_nml (convection_scheme = "DRY"; no defaults in the reference)
    dry_tau: float = 0.0
    dry_gamma: float = 0.0
    # damping_driver_nml
    trayfric: float = 0.0
    sponge_pbottom: float = 50.0


class IdealizedMoistPhys:
    def __init__(self, cfg: MoistPhysConfig, dt_atmos, rad_lat, z_surf, t_surf_init, pref=None, svp=None, rad=None, mo=None,
                 sflux=None, diff=None, sbm=None):
        self.c, self.dt_real = cfg, float(dt_atmos)
        self.svp = svp or SatVaporPres()
        self.rad = GreyRadiation(rad or GreyRadConfig())
        self.mo, self.sflux, self.diffc = mo or MOConfig(), sflux or SurfaceFluxConfig(), diff or DiffusivityConfig()
        self.sbm = sbm or SBMConvection(self.svp)
        self.rad_lat, self.z_surf, self.pref = rad_lat, z_surf, pref
        shp = rad_lat.shape
        self.t_surf = t_surf_init + 1.0                                   # idealized_moist_phys.F90:643
        self.q_surf = np.zeros(shp); self.gust = np.ones(shp)
        self.albedo = np.full(shp, cfg.albedo_value)
        self.heat_capacity = np.full(shp, cfg.depth * RHO_CP)
        self.ocean_qflux = np.zeros(shp)
        # land options (idealized_moist_phys_init / mixed_layer_init): per-column roughness, land mask (None = aquaplanet values)
        self.rough_mom = self.rough_heat = self.rough_moist = self.land = None
        self.diag = {}
        # do_rrtm_radiation (idealized_moist_phys.F90:1167-1177): an oracle.rrtmg.RrtmRadiation instead of the grey scheme
        self.rrtm = None
        self.bm = None                                                    # convection_scheme = "FULL_BETTS_MILLER": an oracle.betts_miller.BettsMiller
        self.time_s = 0.0
        # two_stream_gray_rad_nml do_seasonal: dict(astro=, lon=, solday=, equinox_day=, use_time_average_coszen=, dt_rad_avg=,
        # day_in_s=, year_in_s=) -> seasonal_insolation at Time (idealized_moist_phys.F90:1054 passes Time, not Time + Time_step)
        self.seasonal = None
        self.co2 = None                                                   # do_read_co2: carbon_conc (ppmv) of co2_file at Time, set per step

    def __call__(self, core, delta_t):
        """core: the dynamical core state (ug, vg, tg, grid_tracers, p_half, p_full, z_half, z_full at two time levels)."""
        c = self.c
        prev, cur = core.previous, core.current
        K = core.tg.shape[1]
        tg_p, q_p, ug_p, vg_p = core.tg[prev], core.grid_tracers[prev, 0], core.ug[prev], core.vg[prev]
        dt_ug, dt_vg, dt_tg, dt_q = (np.zeros_like(tg_p) for _ in range(4))
        zero2 = np.zeros_like(self.t_surf)
        if c.convection_scheme == "SIMPLE_BETTS_MILLER":
            o = self.sbm(delta_t, tg_p, q_p, core.p_full[prev], core.p_half[prev])
            conv_dt_tg, conv_dt_qg, rain = o["deltaT"], o["deltaq"], o["rain"]
            tg_tmp, qg_tmp = conv_dt_tg + tg_p, conv_dt_qg + q_p
            conv_dt_tg, conv_dt_qg = conv_dt_tg / delta_t, conv_dt_qg / delta_t
            rain = rain / delta_t
            precip = rain
            self.diag.update(convflag=o["convflag"], cape=o["CAPE"])
        elif c.convection_scheme == "FULL_BETTS_MILLER":                 # idealized_moist_phys.F90:889-916 (self.bm: oracle.betts_miller)
            o = self.bm(delta_t, tg_p, q_p, core.p_full[prev], core.p_half[prev])
            conv_dt_tg, conv_dt_qg, rain = o["deltaT"], o["deltaq"], o["rain"]
            tg_tmp, qg_tmp = conv_dt_tg + tg_p, conv_dt_qg + q_p
            conv_dt_tg, conv_dt_qg = conv_dt_tg / delta_t, conv_dt_qg / delta_t
            rain = rain / delta_t
            precip = rain
            self.diag.update(convflag=o["convflag"], cape=o["CAPE"])
        elif c.convection_scheme == "DRY":                               # idealized_moist_phys.F90:918-928
            conv_dt_tg, cape, cin, lzb, lcl = dry_convection(tg_p, core.p_full[prev], core.p_half[prev], c.dry_tau, c.dry_gamma)
            conv_dt_qg = np.zeros_like(tg_p)
            tg_tmp, qg_tmp = conv_dt_tg * delta_t + tg_p, q_p
            precip = zero2
            self.diag.update(cape=cape, cin=cin, lzb=lzb)
        elif c.convection_scheme == "NONE":
            conv_dt_tg, conv_dt_qg = np.zeros_like(tg_p), np.zeros_like(tg_p)
            tg_tmp, qg_tmp = tg_p, q_p
            precip = zero2
        else:
            raise ValueError("convection scheme not restated")
        dt_tg = dt_tg + conv_dt_tg
        dt_q = dt_q + conv_dt_qg
        if c.convection_scheme != "DRY":                                   # `if (r_conv_scheme .ne. DRY_CONV)` (:977)
            rain, cond_dt_tg, cond_dt_qg = lscale_cond(self.svp, tg_tmp, qg_tmp, core.p_full[prev], core.p_half[prev], hc=c.hc, do_evap=c.do_evap)
            cond_dt_tg, cond_dt_qg = cond_dt_tg / delta_t, cond_dt_qg / delta_t
            rain = rain / delta_t
            precip = precip + rain
            dt_tg = dt_tg + cond_dt_tg
            dt_q = dt_q + cond_dt_qg
        if self.rrtm is None:
            insol = None
            if self.seasonal is not None:
                kw = dict(self.seasonal)
                astro, lon = kw.pop("astro"), kw.pop("lon")
                if kw.get("dt_rad_avg") is None or kw["dt_rad_avg"] <= 0:
                    kw["dt_rad_avg"] = self.dt_real                      # two_stream_gray_rad_init :207
                days = int(self.time_s // 86400.0)
                insol = seasonal_insolation(self.rad.c, astro, days, self.time_s - 86400.0 * days, self.rad_lat, lon, **kw)
                self.diag["insolation"] = insol
            d = self.rad.down(self.rad_lat, core.p_half[cur], tg_p, q=q_p, albedo=self.albedo, insolation=insol, carbon_conc=self.co2)
            net_surf_sw_down = (1.0 - self.albedo) * d["sw_down_surf"]
            surf_lw_down = d["surf_lw_down"]
        sf = surface_flux(self.svp, self.mo, self.sflux, t_atm=tg_p[K - 1], q_atm_in=q_p[K - 1], u_atm=ug_p[K - 1], v_atm=vg_p[K - 1],
                          p_atm=core.p_full[cur][K - 1], z_atm=core.z_full[cur][K - 1] - self.z_surf, p_surf=core.p_half[cur][K],
                          t_surf=self.t_surf, t_ca=self.t_surf, q_surf=self.q_surf, u_surf=zero2, v_surf=zero2,
                          rough_mom=np.full_like(zero2, c.roughness_mom) if self.rough_mom is None else self.rough_mom,
                          rough_heat=np.full_like(zero2, c.roughness_heat) if self.rough_heat is None else self.rough_heat,
                          rough_moist=np.full_like(zero2, c.roughness_moist) if self.rough_moist is None else self.rough_moist,
                          rough_scale=np.full_like(zero2, c.roughness_mom) if self.rough_mom is None else self.rough_mom,
                          gust=self.gust, land=np.zeros(zero2.shape, bool) if self.land is None else self.land)
        self.q_surf = sf["q_surf"]
        if self.rrtm is None:
            dt_tg, _ = self.rad.up(self.t_surf, self.albedo, core.p_half[cur], dt_tg)
        else:
            dt_tg, net_surf_sw_down, surf_lw_down = self.rrtm(self.time_s, core.p_full[cur], core.p_half[cur], core.z_full[cur],
                                                              core.z_half[cur], tg_p, q_p, self.t_surf, self.albedo, dt_tg)
        self.time_s += self.dt_real
        if c.do_damping:
            udt, vdt, tdt, _ = rayleigh_sponge(delta_t, core.p_full[cur], ug_p, vg_p, self.pref, c.sponge_pbottom, c.trayfric, True)
            dt_ug, dt_vg, dt_tg = dt_ug + udt, dt_vg + vdt, dt_tg + tdt
        # vert_turb_driver (do_diffusivity) on the `current` fields (use_tau = .true.)
        if c.use_tau:
            tt, qq, uu, vv = core.tg[cur], core.grid_tracers[cur, 0], core.ug[cur], core.vg[cur]
        else:                                   # vert_turb_driver.F90:209-213: variables at time tau+1
            uu, vv, tt, qq = ug_p + delta_t * dt_ug, vg_p + delta_t * dt_vg, tg_p + delta_t * dt_tg, q_p + delta_t * dt_q
        z_pbl, diff_m, diff_t = diffusivity(self.diffc, self.mo, tt, qq, uu, vv,
                                            core.p_full[cur], core.p_half[cur], core.z_full[cur], core.z_half[cur], sf["u_star"], sf["b_star"],
                                            np.zeros_like(tg_p), np.zeros_like(tg_p))
        self.gust = np.full_like(zero2, c.constant_gust)
        r = gcm_vert_diff_down(delta_t, ug_p, vg_p, tg_p, q_p, diff_m, diff_t, core.p_half[cur], core.p_full[cur], core.z_full[cur],
                               sf["flux_u"], sf["flux_v"], sf["dtaudu_atm"], sf["dtaudv_atm"], dt_ug, dt_vg, dt_tg, dt_q,
                               do_conserve_energy=True, use_virtual_temp=c.do_virtual)
        dt_ug, dt_vg = r["dt_u"], r["dt_v"]
        self.t_surf, tri, dts = mixed_layer(r["tri"], self.dt_real, self.t_surf, sf["flux_t"], sf["flux_q"], sf["flux_r"], net_surf_sw_down,
                                            surf_lw_down, sf["dhdt_surf"], sf["dedt_surf"], sf["dedq_surf"], sf["drdt_surf"], sf["dhdt_atm"],
                                            sf["dedq_atm"], self.heat_capacity, self.ocean_qflux, evaporation=c.evaporation,
                                            sst_new=getattr(self, "sst_new", None))
        dt_tg, dt_q = gcm_vert_diff_up(delta_t, tri)
        self.diag.update(precip=precip, z_pbl=z_pbl, flux_t=sf["flux_t"], flux_q=sf["flux_q"], delta_t_surf=dts, diff_m=diff_m, diff_t=diff_t,
                         net_surf_sw_down=net_surf_sw_down, surf_lw_down=surf_lw_down)
        return dt_ug, dt_vg, dt_tg, [dt_q]
