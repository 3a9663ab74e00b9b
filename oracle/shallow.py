"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's shallow-water model (src/atmos_spectral_shallow/:
atmosphere.F90, shallow_dynamics.F90:200-514, shallow_physics.F90) on the oracle's own transforms.  Only tests/ may import it.
Parity unpinned by the reference; pinned by properties (steady geostrophic state, mass conservation, gravity-wave speed) in
tests/test_shallow.py.  The vortex-pair and input-file initial conditions and the finite-volume grid tracer are not restated."""
from dataclasses import dataclass

import numpy as np

from .isca_oracle import Config, Tables, Transforms


@dataclass
class ShallowConfig:
    num_lon: int = 256
    num_lat: int = 128
    num_fourier: int = 85
    num_spherical: int = 86
    robert_coeff: float = 0.04
    raw_filter_coeff: float = 1.0
    damping_option: str = "resolution_dependent"
    damping_order: int = 4
    damping_coeff: float = 1.0e-04
    h_0: float = 3.0e04
    u_deep_mag: float = 0.0
    n_merid_deep_flow: float = 3.0
    u_upper_mag_init: float = 0.0
    spec_tracer: bool = True
    valid_range_v: tuple = (-1.0e3, 1.0e3)
    # shallow_physics_nml (shallow_physics.F90:58-70); negative damping times are in days
    fric_damp_time: float = -20.0
    therm_damp_time: float = -10.0
    h_amp: float = 2.0e04
    h_lon: float = 90.0
    h_lat: float = 25.0
    h_width: float = 15.0
    h_itcz: float = 1.0e05
    itcz_width: float = 4.0
    phys_h_0: float = 3.0e04
    dt_atmos: float = 1200.0
    radius: float = Config.radius
    omega: float = Config.omega


class ShallowModel:
    def __init__(self, c: ShallowConfig):
        self.c = c
        cfg = Config(lon_max=c.num_lon, lat_max=c.num_lat, num_fourier=c.num_fourier, num_spherical=c.num_spherical, radius=c.radius,
                     omega=c.omega)
        self.tb = tb = Tables(cfg)
        self.tr = tr = Transforms(tb)
        self.eigen = eigen = tb.eigen_laplacian
        N = c.num_spherical
        if c.damping_option == "resolution_dependent":
            self.damping = c.damping_coeff * (eigen / eigen[N - 1, 0]) ** c.damping_order
        elif c.damping_option == "resolution_independent":
            self.damping = c.damping_coeff * eigen ** c.damping_order
        else:
            raise ValueError("damping_option not restated")
        J, I = c.num_lat, c.num_lon
        ss = (2, N + 1, c.num_fourier + 1)
        self.vors, self.divs, self.hs, self.trs_s = (np.zeros(ss, dtype=np.complex128) for _ in range(4))
        self.u, self.v, self.vorg, self.divg, self.h, self.trs_g = (np.zeros((2, J, I)) for _ in range(6))
        d2r = np.pi / 180.0
        lat = tb.deg_lat
        nm = c.n_merid_deep_flow
        dg = -2. * c.omega * c.u_deep_mag * c.radius * (1. / (1. - nm ** 2.)) * (
            -np.cos(nm * d2r * lat) * np.cos(d2r * lat) - nm * (np.sin(nm * d2r * lat) * np.sin(d2r * lat) - np.sin(nm * (2. * np.arctan(1.)))))
        dg = np.repeat(dg[:, None], I, 1)
        self.deep_geopot = dg - tr.area_weighted_global_mean(dg)
        self.h[0] = c.h_0 - self.deep_geopot
        self.vorg[0] = (-((c.u_upper_mag_init * nm) / c.radius) * np.sin(d2r * lat))[:, None]
        self.vors[0] = tr.grid_to_spherical(self.vorg[0])
        self.divs[0] = tr.grid_to_spherical(self.divg[0])
        self.hs[0] = tr.grid_to_spherical(self.h[0])
        self.u[0], self.v[0] = tr.uv_grid_from_vor_div(self.vors[0], self.divs[0])
        if c.spec_tracer:
            g = np.zeros((J, I))
            g[(lat > 10.0) & (lat < 20.0)] = 1.0
            g[lat > 70.0] = -1.0
            self.trs_g[0] = g
            self.trs_s[0] = tr.grid_to_spherical(g)
        # shallow_physics_init
        fd = -c.fric_damp_time * 86400 if c.fric_damp_time < 0 else c.fric_damp_time
        td = -c.therm_damp_time * 86400 if c.therm_damp_time < 0 else c.therm_damp_time
        self.kappa_m = 1.0 / fd if fd != 0.0 else 0.0
        self.kappa_t = 1.0 / td if td != 0.0 else 0.0
        xx = (tb.deg_lon[None, :] - c.h_lon) / (c.h_width * 2.0)
        yy = (lat[:, None] - c.h_lat) / c.h_width
        self.h_eq = c.phys_h_0 + c.h_amp * np.maximum(1.0e-10, np.exp(-(xx * xx + yy * yy)))
        yy = lat / c.itcz_width
        self.h_eq = self.h_eq + (c.h_itcz * np.exp(-yy * yy))[:, None]
        self.coriolis = 2 * c.omega * tb.sin_lat
        self.previous = self.current = 0
        self.first = True

    def _damp(self, spec_prev, dt_spec, delta_t):
        return (dt_spec - self.damping * spec_prev) / (1.0 + self.damping * delta_t)

    @staticmethod
    def _leapfrog(a, dt_a, previous, current, future, delta_t, robert, raw):
        part = a[previous] - 2.0 * a[current]
        if previous == current:
            a[future] = a[previous] + delta_t * dt_a
            a[current] = a[current] + robert * (part + a[future]) * raw
        else:
            a[current] = a[current] + robert * part * raw
            a[future] = a[previous] + delta_t * dt_a
            a[current] = a[current] + robert * a[future] * raw
        a[future] = a[future] + robert * (part + a[future]) * (raw - 1.0)

    def step(self):
        c, tr = self.c, self.tr
        if self.first:
            delta_t, future = c.dt_atmos, 1
        else:
            delta_t, future = 2.0 * c.dt_atmos, self.previous
        p, cur = self.previous, self.current
        # shallow_physics (:113-128)
        tend_u = -self.kappa_m * self.u[p]
        tend_v = -self.kappa_m * self.v[p]
        tend_h = -self.kappa_t * (self.h[p] - self.h_eq)
        # shallow_dynamics (:408-480)
        vorg = self.vorg[cur] + self.coriolis[:, None]
        tend_u = tend_u + vorg * self.v[cur]
        tend_v = tend_v - vorg * self.u[cur]
        dt_vors, dt_divs = tr.vor_div_from_uv_grid(tend_u, tend_v)
        tend_h = tr.horizontal_advection(self.hs[cur], self.u[cur], self.v[cur], tend_h)
        tend_h = tend_h - self.h[cur] * self.divg[cur]
        dt_hs = tr.grid_to_spherical(tend_h)
        bg = self.h[cur] + self.deep_geopot + 0.5 * (self.u[cur] ** 2 + self.v[cur] ** 2)
        dt_divs = dt_divs - tr.laplacian(tr.grid_to_spherical(bg))
        # implicit_correction (:493-514)
        mu = 0.5 * delta_t
        dt_hs = dt_hs + c.h_0 * (self.divs[cur] - self.divs[p])
        dt_divs = dt_divs - self.eigen * (self.hs[cur] - self.hs[p])
        dt_divs = (dt_divs + mu * self.eigen * dt_hs) / (1.0 + mu * mu * self.eigen * c.h_0)
        dt_hs = dt_hs - mu * c.h_0 * dt_divs
        dt_vors = self._damp(self.vors[p], dt_vors, delta_t)
        dt_divs = self._damp(self.divs[p], dt_divs, delta_t)
        dt_hs = self._damp(self.hs[p], dt_hs, delta_t)
        for a, d in ((self.vors, dt_vors), (self.divs, dt_divs), (self.hs, dt_hs)):
            self._leapfrog(a, d, p, cur, future, delta_t, c.robert_coeff, c.raw_filter_coeff)
        self.vorg[future] = tr.spherical_to_grid(self.vors[future])
        self.divg[future] = tr.spherical_to_grid(self.divs[future])
        self.u[future], self.v[future] = tr.uv_grid_from_vor_div(self.vors[future], self.divs[future])
        self.h[future] = tr.spherical_to_grid(self.hs[future])
        if self.v.min() < c.valid_range_v[0] or self.v.max() > c.valid_range_v[1]:
            raise FloatingPointError("shallow_dynamics: meridional wind out of valid range")
        if c.spec_tracer:
            dt_tr = tr.horizontal_advection(self.trs_s[cur], self.u[cur], self.v[cur], np.zeros_like(self.u[cur]))
            dt_trs = self._damp(self.trs_s[p], tr.grid_to_spherical(dt_tr), delta_t)
            self._leapfrog(self.trs_s, dt_trs, p, cur, future, delta_t, c.robert_coeff, c.raw_filter_coeff)
            self.trs_g[future] = tr.spherical_to_grid(self.trs_s[future])
        self.previous, self.current = cur, future
        self.first = False

    def global_diag(self):
        """atmosphere.F90 global_diag: enstrophy, div_squared, max_Froude of the current level"""
        k = self.current
        sp = self.u[k] ** 2 + self.v[k] ** 2
        return (self.tr.area_weighted_global_mean(self.vorg[k] ** 2), self.tr.area_weighted_global_mean(self.divg[k] ** 2),
                float((sp / self.h[k]).max()))
