"""The reference's shallow-water model (src/atmos_spectral_shallow/) on the transform-level C ABI.

`ShallowAtmosphere` mirrors that model's `atmosphere_mod` (atmosphere.F90), `shallow_dynamics_mod` (shallow_dynamics.F90:200-514) and
`shallow_physics_mod` (shallow_physics.F90).  Every spherical-harmonic transform of a step -- vor_div_from_uv_grid, three
trans_grid_to_spherical, four trans_spherical_to_grid, three uv_grid_from_vor_div -- runs in the CUDA library through the
`transforms_mod`-level entry points of include/isca_b200.h; the per-coefficient algebra (semi-implicit gravity-wave correction, implicit
damping, leapfrog with the RAW filter) stays on the host as in the Fortran.  There is no CPU transform path.

Not built (rejected at construction): grid_tracer = .true., add_initial_vortex_pair, initial_condition_from_input_file, stirring,
damping_option = 'exponential_cutoff', restarts."""
from __future__ import annotations
import numpy as np
from .api import IscaError
from .barotropic import _make_engine

DYN_DEFAULTS = dict(num_lon=256, num_lat=128, num_fourier=85, num_spherical=86, robert_coeff=0.04, raw_filter_coeff=1.0,
                    damping_option="resolution_dependent", damping_order=4, damping_coeff=1.0e-04, h_0=3.0e04, u_deep_mag=0.0,
                    n_merid_deep_flow=3.0, u_upper_mag_init=0.0, spec_tracer=True, grid_tracer=False, add_initial_vortex_pair=False,
                    initial_condition_from_input_file=False, valid_range_v=(-1.0e3, 1.0e3))          # shallow_dynamics_nml (:110-172)
PHYS_DEFAULTS = dict(fric_damp_time=-20.0, therm_damp_time=-10.0, del_h=1.0e04, h_0=3.0e04, h_amp=2.0e04, h_lon=90.0, h_lat=25.0,
                     h_width=15.0, h_itcz=1.0e05, itcz_width=4.0)                                      # shallow_physics_nml (:58-70)


class ShallowAtmosphere:
    def __init__(self, dt_atmos: float, physics_nml: dict | None = None, **nml):
        bad = set(nml) - set(DYN_DEFAULTS)
        if bad:
            raise IscaError(f"unknown shallow_dynamics_nml variable(s) {sorted(bad)}")
        badp = set(physics_nml or {}) - set(PHYS_DEFAULTS)
        if badp:
            raise IscaError(f"unknown shallow_physics_nml variable(s) {sorted(badp)}")
        self.nml = n = dict(DYN_DEFAULTS, **nml)
        self.pnml = pn = dict(PHYS_DEFAULTS, **(physics_nml or {}))
        for k in ("grid_tracer", "add_initial_vortex_pair", "initial_condition_from_input_file"):
            if n[k]:
                raise IscaError(f"shallow_dynamics: {k} = .true. is not built")
        if n["damping_option"] not in ("resolution_dependent", "resolution_independent"):
            raise IscaError('spectral_damping_init: "%s" is an invalid (or not built) value for damping_option' % n["damping_option"])
        self.dt_real = float(dt_atmos)
        self.eng, self.radius, self.omega = _make_engine(n)
        from . import api
        sin_lat = self.eng.get_table(api.TB_SIN_LAT)
        self.wts_lat = self.eng.get_table(api.TB_WTS_LAT)
        self.deg_lat = lat = self.eng.get_table(api.TB_DEG_LAT)
        self.deg_lon = self.eng.get_table(api.TB_DEG_LON)
        self.coriolis = 2 * self.omega * sin_lat
        M, N = n["num_fourier"], n["num_spherical"]
        L = np.arange(N + 1)[:, None] + np.arange(M + 1)[None, :]
        self.eigen = L * (L + 1.0) / (self.radius * self.radius)
        if n["damping_option"] == "resolution_dependent":
            self.damping = n["damping_coeff"] * (self.eigen / self.eigen[N - 1, 0]) ** n["damping_order"]
        else:
            self.damping = n["damping_coeff"] * self.eigen ** n["damping_order"]
        J, I = n["num_lat"], n["num_lon"]
        self.spec = {k: np.zeros((2, N + 1, M + 1), dtype=np.complex128) for k in ("vor", "div", "h", "trs")}
        self.grid = {k: np.zeros((2, J, I)) for k in ("u", "v", "vor", "div", "h", "trs")}
        self._w = self.wts_lat[:, None] / (self.wts_lat.sum() * I)             # area_weighted_global_mean weights
        d2r, nm = np.pi / 180.0, n["n_merid_deep_flow"]
        dg = -2. * self.omega * n["u_deep_mag"] * self.radius * (1. / (1. - nm ** 2.)) * (
            -np.cos(nm * d2r * lat) * np.cos(d2r * lat) - nm * (np.sin(nm * d2r * lat) * np.sin(d2r * lat) - np.sin(nm * (2. * np.arctan(1.)))))
        dg = np.repeat(dg[:, None], I, 1)
        self.deep_geopot = dg - float((self._w * dg).sum())
        g, s = self.grid, self.spec
        g["h"][0] = n["h_0"] - self.deep_geopot
        g["vor"][0] = (-((n["u_upper_mag_init"] * nm) / self.radius) * np.sin(d2r * lat))[:, None]
        for k in ("vor", "div", "h"):
            s[k][0] = self.eng.trans_grid_to_spherical(g[k][0])
        g["u"][0], g["v"][0] = self._uv(s["vor"][0], s["div"][0])
        if n["spec_tracer"]:
            t = np.zeros((J, I))
            t[(lat > 10.0) & (lat < 20.0)] = 1.0
            t[lat > 70.0] = -1.0
            g["trs"][0] = t
            s["trs"][0] = self.eng.trans_grid_to_spherical(t)
        fd = -pn["fric_damp_time"] * 86400 if pn["fric_damp_time"] < 0 else pn["fric_damp_time"]
        td = -pn["therm_damp_time"] * 86400 if pn["therm_damp_time"] < 0 else pn["therm_damp_time"]
        self.kappa_m = 1.0 / fd if fd != 0.0 else 0.0
        self.kappa_t = 1.0 / td if td != 0.0 else 0.0
        xx = (self.deg_lon[None, :] - pn["h_lon"]) / (pn["h_width"] * 2.0)
        yy = (lat[:, None] - pn["h_lat"]) / pn["h_width"]
        self.h_eq = pn["h_0"] + pn["h_amp"] * np.maximum(1.0e-10, np.exp(-(xx * xx + yy * yy)))
        yy = lat / pn["itcz_width"]
        self.h_eq = self.h_eq + (pn["h_itcz"] * np.exp(-yy * yy))[:, None]
        self.previous = self.current = 0
        self._first = True

    def _uv(self, vor, div):
        u, v = self.eng.uv_grid_from_vor_div(vor[None], div[None])
        return u[0], v[0]

    def _grad(self, s):
        return self._uv(np.zeros_like(s), -self.eigen * s)         # gradient = irrotational wind of the potential s

    def _leap(self, a, tend, prev, cur, fut, delta_t):
        rc, raw = self.nml["robert_coeff"], self.nml["raw_filter_coeff"]
        tend = (tend - self.damping * a[prev]) / (1.0 + self.damping * delta_t)
        part = a[prev] - 2.0 * a[cur]
        if prev == cur:
            a[fut] = a[prev] + delta_t * tend
            a[cur] = a[cur] + rc * (part + a[fut]) * raw
        else:
            a[cur] = a[cur] + rc * part * raw
            a[fut] = a[prev] + delta_t * tend
            a[cur] = a[cur] + rc * a[fut] * raw
        a[fut] = a[fut] + rc * (part + a[fut]) * (raw - 1.0)

    def atmosphere(self, n_steps: int = 1):
        g, s, e, h_0 = self.grid, self.spec, self.eng, self.nml["h_0"]
        for _ in range(n_steps):
            if self._first:
                delta_t, fut = self.dt_real, 1
            else:
                delta_t, fut = 2.0 * self.dt_real, self.previous
            prev, cur = self.previous, self.current
            av = g["vor"][cur] + self.coriolis[:, None]
            tu = -self.kappa_m * g["u"][prev] + av * g["v"][cur]
            tv = -self.kappa_m * g["v"][prev] - av * g["u"][cur]
            dvor, ddiv = e.vor_div_from_uv_grid(tu[None], tv[None])
            dvor, ddiv = dvor[0], ddiv[0]
            hx, hy = self._grad(s["h"][cur])
            th = -self.kappa_t * (g["h"][prev] - self.h_eq) - g["u"][cur] * hx - g["v"][cur] * hy - g["h"][cur] * g["div"][cur]
            dh = e.trans_grid_to_spherical(th)
            bern = g["h"][cur] + self.deep_geopot + 0.5 * (g["u"][cur] ** 2 + g["v"][cur] ** 2)
            ddiv = ddiv + self.eigen * e.trans_grid_to_spherical(bern)            # - laplacian(bs)
            mu = 0.5 * delta_t                                                     # implicit_correction (:493-514), xi = 0.5
            dh = dh + h_0 * (s["div"][cur] - s["div"][prev])
            ddiv = ddiv - self.eigen * (s["h"][cur] - s["h"][prev])
            ddiv = (ddiv + mu * self.eigen * dh) / (1.0 + mu * mu * self.eigen * h_0)
            dh = dh - mu * h_0 * ddiv
            self._leap(s["vor"], dvor, prev, cur, fut, delta_t)
            self._leap(s["div"], ddiv, prev, cur, fut, delta_t)
            self._leap(s["h"], dh, prev, cur, fut, delta_t)
            for k in ("vor", "div", "h"):
                g[k][fut] = e.trans_spherical_to_grid(s[k][fut])
            g["u"][fut], g["v"][fut] = self._uv(s["vor"][fut], s["div"][fut])
            lo, hi = self.nml["valid_range_v"]
            if g["v"].min() < lo or g["v"].max() > hi:
                raise IscaError("shallow_dynamics: meridional wind out of valid range")
            if self.nml["spec_tracer"]:
                tx, ty = self._grad(s["trs"][cur])
                dt = e.trans_grid_to_spherical(-g["u"][cur] * tx - g["v"][cur] * ty)
                self._leap(s["trs"], dt, prev, cur, fut, delta_t)
                g["trs"][fut] = e.trans_spherical_to_grid(s["trs"][fut])
            self.previous, self.current = cur, fut
            self._first = False

    def global_diag(self):
        """enstrophy, div_squared, max_Froude as atmosphere.F90 global_diag prints them"""
        k, g = self.current, self.grid
        sp = g["u"][k] ** 2 + g["v"][k] ** 2
        return float((self._w * g["vor"][k] ** 2).sum()), float((self._w * g["div"][k] ** 2).sum()), float((sp / g["h"][k]).max())

    def atmosphere_end(self):
        if self.eng is not None:
            self.eng.atmosphere_end()
            self.eng = None
