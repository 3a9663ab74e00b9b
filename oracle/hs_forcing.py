"""TEST INFRASTRUCTURE -- NumPy restatement of hs_forcing_mod with every option that needs no input file
(atmos_param/hs_forcing/hs_forcing.F90).  The default Held-Suarez path is also restated in oracle/isca_oracle.py (HSForcing);
this module adds what hs_forcing_nml can switch on beyond it:

  equilibrium_t_option = 'Held_Suarez' | 'EXOPLANET' | 'EXOPLANET2'   (newtonian_damping, :508-603)
  equilibrium_t_option = 'top_down' with stratosphere_t_option        (top_down_newtonian_damping, :894-1026; spin-up :331-366)
  local_heating_option = 'Isidoro'                                    (local_heating, :727-765)
  tracer_source_sink                                                  (:683-724)

Not restated (they read netCDF files through interpolator_mod): equilibrium_t_option = 'from_file',
local_heating_option = 'from_file', relax_to_specified_wind.

Parity unpinned: the reference holds no known-answer data for hs_forcing; the default option is cross-checked against
the independent restatement in oracle/isca_oracle.py, the others through analytic properties (tests/test_oracle_hs.py).
Arrays are [lev, lat, lon] (level 0 = model top), 2-D arrays [lat, lon].
"""
from __future__ import annotations
from dataclasses import dataclass
import numpy as np


@dataclass
class HsConfig:
    # hs_forcing_nml defaults (hs_forcing.F90:74-107)
    no_forcing: bool = False
    t_zero: float = 315.0
    t_strat: float = 200.0
    delh: float = 60.0
    delv: float = 10.0
    eps: float = 0.0
    sigma_b: float = 0.7
    P00: float = 1.0e5
    p_trop: float = 1.0e4
    alpha: float = 2.0 / 7
    ka: float = -40.0
    ks: float = -4.0
    kf: float = -1.0
    do_conserve_energy: bool = True
    trflux: float = 1.0e-5
    trsink: float = -4.0
    local_heating_option: str = ""
    local_heating_srfamp: float = 0.0
    local_heating_xwidth: float = 10.0
    local_heating_ywidth: float = 10.0
    local_heating_xcenter: float = 180.0
    local_heating_ycenter: float = 45.0
    local_heating_vert_decay: float = 1.0e4
    equilibrium_t_option: str = "Held_Suarez"
    stratosphere_t_option: str = "extend_tp"
    peri_time: float = 0.25
    smaxis: float = 1.5e6
    albedo: float = 0.3
    lapse: float = 6.5
    h_a: float = 2.0
    tau_s: float = 5.0
    heat_capacity: float = 4.2e6
    ml_depth: float = 1.0
    spinup_time: float = 10800.0
    # constants_mod (shared/constants/constants.F90) and astronomy_nml values the module uses
    kappa: float = 2.0 / 7
    rdgas: float = 287.04
    grav: float = 9.80
    stefan: float = 5.6734e-8
    solar_const: float = 1368.22
    omega: float = 7.2921150e-5
    orbital_period: float = 365.25 * 86400.0      # constants_mod value (s); update_orbit multiplies it by 86400 (:822-828)
    orbital_rate: float = None                    # 2 pi / orbital_period (constants_init :307); None = that
    ecc: float = 0.0
    obliq: float = 23.439

    @property
    def cp_air(self):
        return self.rdgas / self.kappa


class HsForcing:
    def __init__(self, cfg: HsConfig, lat2d=None, days=0, seconds=0, astronomy=None):
        """hs_forcing_init (:276-470).  lat2d, (days, seconds) = Time: needed by the top_down spin-up; astronomy: an
        oracle.rrtmg.Astronomy for the EXOPLANET options (diurnal_exoplanet -> diurnal_solar_2d)."""
        c = self.c = cfg
        self.tka = -1.0 / (86400 * c.ka) if c.ka < 0.0 else c.ka                  # :392-406
        self.tks = -1.0 / (86400 * c.ks) if c.ks < 0.0 else c.ks
        self.vkf = -1.0 / (86400 * c.kf) if c.kf < 0.0 else c.kf
        self.opt = c.equilibrium_t_option if c.equilibrium_t_option in ("Held_Suarez", "top_down") else c.equilibrium_t_option.upper()
        if self.opt not in ("Held_Suarez", "top_down", "EXOPLANET", "EXOPLANET2"):
            raise ValueError(f'"{c.equilibrium_t_option}"  is not a valid value for equilibrium_t_option')
        if c.local_heating_option not in ("", "Isidoro"):
            raise ValueError(f'"{c.local_heating_option}"  is not a valid value for local_heating_option')
        self.astro = astronomy
        self.orbital_rate = 2 * np.pi / c.orbital_period if c.orbital_rate is None else c.orbital_rate
        twopi = 2 * np.pi
        self.xwidth, self.ywidth = np.deg2rad(c.local_heating_xwidth), np.deg2rad(c.local_heating_ywidth)     # :370-373
        self.xcenter, self.ycenter = np.deg2rad(c.local_heating_xcenter), np.deg2rad(c.local_heating_ycenter)
        self.xcenter = self.xcenter - twopi * np.floor(self.xcenter / twopi)                                  # :377
        self.srfamp = c.local_heating_srfamp / 86400.0                                                         # :381 (deg/day -> deg/s)
        self.tg_prev = None
        self.diag = {}
        if self.opt == "top_down" and not c.no_forcing:
            self.tg_prev = self.spin_up(lat2d, 86400 * int(days) + int(seconds))

    # ---- orbit of the top-down forcing (:816-858)
    def update_orbit(self, current_time):
        c = self.c
        theta = 2 * np.pi * current_time / (c.orbital_period * 86400)
        return np.arcsin(np.sin(c.obliq * np.pi / 180) * np.sin(theta))

    @staticmethod
    def calc_hour_angle(lat, dec):
        return np.arccos(np.clip(-np.tan(lat) * np.tan(dec), -1.0, 1.0))

    def _radiative_surface(self, lat, dec):
        """s, t_radbal, t_trop, h_trop, t_surf of :347-357 / :946-958"""
        c = self.c
        ha = self.calc_hour_angle(lat, dec)
        s = c.solar_const / np.pi * (ha * np.sin(lat) * np.sin(dec) + np.cos(lat) * np.cos(dec) * np.sin(ha))
        t_radbal = ((1 - c.albedo) * s / c.stefan) ** 0.25
        t_trop = t_radbal / (2 ** 0.25)
        h_trop = 1.0 / (16 * c.lapse) * (1.3863 * t_trop + np.sqrt((1.3863 * t_trop) ** 2 + 32 * c.lapse * c.tau_s * c.h_a * t_trop))
        return t_trop, h_trop, t_trop + h_trop * c.lapse

    def spin_up(self, lat, dt_integer):
        """the spin-up loop of hs_forcing_init (:338-363): tg_prev is left at the value BEFORE the last update (the loop assigns
        tg_prev = tg at its top and exits after computing a new tg that is then discarded)"""
        c = self.c
        tg = np.full(lat.shape, 250.0)
        spin_count, step_days = 0, 1
        while True:
            tg_prev = tg
            dt_integer = dt_integer + 86400 * step_days
            spin_count += 1
            dec = self.update_orbit(dt_integer)
            _, _, t_surf = self._radiative_surface(lat, dec)
            tg = c.stefan * 86400 * step_days / (c.ml_depth * c.heat_capacity) * (t_surf ** 4 - tg_prev ** 4) + tg_prev
            if spin_count >= c.spinup_time:
                break
        return tg_prev

    # ---- rayleigh_damping (:607-679), relax_to_specified_wind = .false.
    def rayleigh_damping(self, ps, p_full, u, v):
        c = self.c
        vcoeff = -self.vkf / (1.0 - c.sigma_b)
        sigma = p_full * (1.0 / ps)[None]
        act = (sigma <= 1.0) & (sigma > c.sigma_b)
        vfactr = vcoeff * (sigma - c.sigma_b)
        return np.where(act, vfactr * u, 0.0), np.where(act, vfactr * v, 0.0)

    def diurnal_exoplanet(self, lat, lon, total_seconds):
        """astronomy.f90:3672-3715: `get_time(Time, seconds)` without days = the whole time in seconds"""
        c = self.c
        substellar_lon = (self.orbital_rate - c.omega) * total_seconds
        gmt = np.mod(-substellar_lon, 2.0 * np.pi)
        frac_of_year = np.mod(self.orbital_rate * total_seconds, 1.0)
        return self.astro.diurnal_solar(lat, lon, gmt, frac_of_year * 2.0 * np.pi)[0]

    def _tdamp(self, lat, ps, p_full):
        c = self.c
        cos_lat_2 = 1.0 - np.sin(lat) ** 2
        cos_lat_4 = cos_lat_2 * cos_lat_2
        tcoeff = (self.tks - self.tka) / (1.0 - c.sigma_b)
        sigma = p_full * (1.0 / ps)[None]
        act = (sigma <= 1.0) & (sigma > c.sigma_b)
        return np.where(act, self.tka + cos_lat_4[None] * (tcoeff * (sigma - c.sigma_b)), self.tka)

    # ---- newtonian_damping (:508-603)
    def newtonian_damping(self, total_seconds, lat, lon, ps, p_full, t):
        c = self.c
        sin_lat, cos_lat = np.sin(lat), np.cos(lat)
        cos_lat_2 = 1.0 - sin_lat * sin_lat
        t_star = c.t_zero - c.delh * sin_lat * sin_lat - c.eps * sin_lat
        tstr = c.t_strat - c.eps * sin_lat
        if self.opt == "Held_Suarez":
            p_norm = p_full / c.P00
            the = t_star[None] - c.delv * cos_lat_2[None] * np.log(p_norm)
            teq = np.maximum(the * p_norm ** c.kappa, tstr[None])
        elif self.opt == "EXOPLANET":
            coszen = self.diurnal_exoplanet(lat, lon, total_seconds)
            self.diag["coszen"] = coszen
            t_star = c.t_zero - c.delh * (1 - coszen) - c.eps * sin_lat
            p_norm = p_full / c.P00
            the = t_star[None] - c.delv * coszen[None] * np.log(p_norm)
            teq = np.maximum(the * p_norm ** c.kappa, tstr[None])
        else:                                                              # EXOPLANET2
            p_norm = p_full / c.p_trop
            teq = np.maximum(c.t_strat * cos_lat[None] * p_norm ** c.alpha, c.t_strat)
        tdamp = self._tdamp(lat, ps, p_full)
        return -tdamp * (t - teq), teq

    # ---- top_down_newtonian_damping (:894-1026)
    def top_down_newtonian_damping(self, total_seconds, lat, ps, p_full, t, dt, zfull):
        c = self.c
        dec = self.update_orbit(int(total_seconds))
        t_trop, h_trop, t_surf = self._radiative_surface(lat, dec)
        tg = c.stefan * dt / (c.ml_depth * c.heat_capacity) * (t_surf ** 4 - self.tg_prev ** 4) + self.tg_prev
        self.tg_prev = tg
        t_trop = tg - h_trop * c.lapse
        tstr = c.t_strat - c.eps * np.sin(lat)
        teq = t_trop[None] + c.lapse * (h_trop[None] - zfull / 1000)
        above = zfull / 1000 >= h_trop[None]
        so = c.stratosphere_t_option
        if so == "c_above_tp":
            teq = np.where(above, tstr[None] + 0 * teq, teq)
        elif so == "hs_like":
            teq = np.maximum(teq, tstr[None])
        elif so == "extend_tp":
            teq = np.where(above, t_trop[None] + 0 * teq, teq)
        else:
            teq = np.maximum(teq, 0.0)
        tdamp = self._tdamp(lat, ps, p_full)
        return -tdamp * (t - teq), teq, h_trop

    # ---- local_heating 'Isidoro' (:727-765)
    def local_heating(self, lon, lat, ps, p_full):
        c = self.c
        twopi = 2 * np.pi
        lon_temp = lon - twopi * np.floor(lon / twopi)
        lon_factor = np.exp(-.5 * ((lon_temp - self.xcenter) / self.xwidth) ** 2)
        lat_factor = np.exp(-.5 * ((lat - self.ycenter) / self.ywidth) ** 2)
        p_factor = np.exp((p_full - ps[None]) / c.local_heating_vert_decay)
        return self.srfamp * lon_factor[None] * lat_factor[None] * p_factor

    # ---- tracer_source_sink (:683-724), no kbot
    @staticmethod
    def tracer_source_sink(flux, damp, p_half, r):
        rdamp = damp
        if rdamp < 0.0:
            rdamp = -86400.0 * rdamp
        if rdamp > 0.0:
            rdamp = 1.0 / rdamp
        source = np.zeros_like(r)
        K = r.shape[0]
        source[K - 1] = flux / (p_half[K] - p_half[K - 1])
        return source - rdamp * r

    # ---- hs_forcing (:148-272)
    def __call__(self, dt, total_seconds, lon, lat, p_half, p_full, u, v, t, r, um, vm, tm, rm, udt, vdt, tdt, rdt, zfull=None):
        """total_seconds = 86400*days + seconds of Time (atmosphere.F90:304 passes Time_next).  r, rm, rdt: lists of tracer arrays.
        -> udt, vdt, tdt, rdt (new arrays); self.diag holds teq, h_trop, tdt_ndamp, local_heating, tdt_diss."""
        c = self.c
        if c.no_forcing:
            return udt, vdt, tdt, rdt
        ps = p_half[-1]
        utnd, vtnd = self.rayleigh_damping(ps, p_full, u, v)
        if c.do_conserve_energy:
            ttnd = -((um + .5 * utnd * dt) * utnd + (vm + .5 * vtnd * dt) * vtnd) / c.cp_air
            tdt = tdt + ttnd
            self.diag["tdt_diss"] = ttnd
        udt = udt + utnd
        vdt = vdt + vtnd
        if self.opt == "top_down":
            ttnd, teq, h_trop = self.top_down_newtonian_damping(total_seconds, lat, ps, p_full, t, dt, zfull)
            self.diag["h_trop"] = h_trop
        else:
            ttnd, teq = self.newtonian_damping(total_seconds, lat, lon, ps, p_full, t)
        tdt = tdt + ttnd
        self.diag.update(teq=teq, tdt_ndamp=ttnd)
        if c.local_heating_option != "":
            ttnd = self.local_heating(lon, lat, ps, p_full)
            tdt = tdt + ttnd
            self.diag["local_heating"] = ttnd
        new = []
        for n in range(len(rdt)):
            rst = rm[n] + dt * rdt[n]
            new.append(rdt[n] + self.tracer_source_sink(c.trflux, c.trsink, p_half, rst))
        return udt, vdt, tdt, new


class CoreHsForcing:
    """adapter with the call signature SpectralCore.step uses for its `hs` member (oracle/isca_oracle.py:1127), keeping the model
    clock: atmosphere.F90:298-311 calls hs_forcing with Time_next = Time + Time_step, u = um = ug(previous), ... and z_full(current)"""

    def __init__(self, hs: HsForcing, core, lon2d, lat2d, time_s=0.0):
        self.hs, self.core, self.lon, self.lat, self.time_s = hs, core, lon2d, lat2d, float(time_s)

    def __call__(self, dt, p_half, p_full, u, v, t, r, udt, vdt, tdt, rdt):
        core = self.core
        time_next = self.time_s + core.cfg.dt_atmos
        out = self.hs(dt, time_next, self.lon, self.lat, p_half, p_full, u, v, t, r, u, v, t, r, udt, vdt, tdt, rdt,
                      zfull=core.z_full[core.current])
        self.time_s = time_next
        return out
