"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's barotropic vorticity-equation model
(src/atmos_spectral_barotropic/: atmosphere.F90:96-205, barotropic_dynamics.F90:166-437) on the oracle's own transforms.
Only tests/ may import this module.  Parity unpinned by the reference (no golden vectors); pinned by conservation properties
(energy, enstrophy of the undamped equation) in tests/test_oracle_barotropic.py.

Namelist defaults are those of barotropic_dynamics_nml (:110-146).  `stirring` (amplitude = 0 by default) and
`barotropic_physics` (empty) do nothing in the reference's default configuration and are not restated; the finite-volume grid tracer
uses oracle/fv_advection.py."""
from dataclasses import dataclass

import numpy as np

from .isca_oracle import Config, Tables, Transforms


@dataclass
class BarotropicConfig:
    num_lon: int = 256
    num_lat: int = 128
    num_fourier: int = 85
    num_spherical: int = 86
    robert_coeff: float = 0.04
    raw_filter_coeff: float = 1.0
    damping_option: str = "resolution_dependent"
    damping_order: int = 4
    damping_coeff: float = 1.0e-04
    damping_coeff_r: float = 0.0
    cutoff_wn: int = 30
    zeta_0: float = 8.0e-05
    m_0: int = 4
    eddy_width: float = 15.0
    eddy_lat: float = 45.0
    spec_tracer: bool = True
    initial_zonal_wind: str = "two_jets"
    valid_range_v: tuple = (-1.0e3, 1.0e3)
    dt_atmos: float = 1200.0
    radius: float = Config.radius            # constants.F90: RADIUS, OMEGA (the oracle core's defaults)
    omega: float = Config.omega


class BarotropicModel:
    def __init__(self, c: BarotropicConfig):
        self.c = c
        cfg = Config(lon_max=c.num_lon, lat_max=c.num_lat, num_fourier=c.num_fourier, num_spherical=c.num_spherical, radius=c.radius,
                     omega=c.omega)
        self.tb = Tables(cfg)
        self.tr = Transforms(self.tb)
        tb = self.tb
        eigen = tb.eigen_laplacian                                   # [n, m]
        N = c.num_spherical
        # spectral_damping_init (spectral_damping.F90:56-168); eigen(0, num_spherical-1) is m = 0, n = N-1
        if c.damping_option == "resolution_dependent":
            self.damping = c.damping_coeff * (eigen / eigen[N - 1, 0]) ** c.damping_order
        elif c.damping_option == "resolution_independent":
            self.damping = c.damping_coeff * eigen ** c.damping_order
        else:
            raise ValueError("damping_option not restated")
        self.damping = self.damping + c.damping_coeff_r
        J, I = c.num_lat, c.num_lon
        shape_s = (N + 1, c.num_fourier + 1)
        self.vors = np.zeros((2,) + shape_s, dtype=np.complex128)
        self.u, self.v, self.vorg = (np.zeros((2, J, I)) for _ in range(3))
        self.trs_s = np.zeros((2,) + shape_s, dtype=np.complex128)
        self.trs_g = np.zeros((2, J, I))
        self.stream = np.zeros((J, I))
        cos_lat, sin_lat, deg_lat = tb.cos_lat, tb.sin_lat, tb.deg_lat
        rad_lon = tb.deg_lon * np.arctan(1.0) / 45.0
        if c.initial_zonal_wind == "zero":
            u0 = np.zeros(J)
        elif c.initial_zonal_wind == "two_jets":
            u0 = 25.0 * cos_lat - 30.0 * cos_lat ** 3 + 300.0 * sin_lat ** 2 * cos_lat ** 6
        else:
            raise ValueError("not a valid value of initial_zonal_wind")
        self.u[0] = u0[:, None]
        vor, div = self.tr.vor_div_from_uv_grid(self.u[0], self.v[0])
        self.vors[0] = vor
        self.vorg[0] = self.tr.spherical_to_grid(vor)
        yy = (deg_lat - c.eddy_lat) / c.eddy_width
        self.vorg[0] = self.vorg[0] + 0.5 * c.zeta_0 * cos_lat[:, None] * np.exp(-yy * yy)[:, None] * np.cos(c.m_0 * rad_lon)[None, :]
        self.vors[0] = self.tr.grid_to_spherical(self.vorg[0])
        self.u[0], self.v[0] = self.tr.uv_grid_from_vor_div(self.vors[0], np.zeros_like(self.vors[0]))
        if c.spec_tracer:
            g = np.zeros((J, I))
            g[(deg_lat > 10.0) & (deg_lat < 20.0)] = 1.0
            g[deg_lat > 70.0] = -1.0
            self.trs_g[0] = g
            self.trs_s[0] = self.tr.grid_to_spherical(g)
        self.previous, self.current = 0, 0                          # forward step first (atmosphere.F90:140-143)
        self.first = True
        self.coriolis = 2 * c.omega * sin_lat

    def _damp(self, spec_prev, dt_spec, delta_t):                   # compute_spectral_damping (spectral_damping.F90:172-195)
        coeff = 1.0 / (1.0 + self.damping * delta_t)
        return coeff * (dt_spec - self.damping * spec_prev)

    @staticmethod
    def _leapfrog(a, dt_a, previous, current, future, delta_t, robert, raw):   # leapfrog_3d_complex (leapfrog.F90:217-247)
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
        """atmosphere(Time) (atmosphere.F90:150-205) -> (energy, enstrophy) as the reference prints them"""
        c, tr = self.c, self.tr
        if self.first:
            delta_t, future = c.dt_atmos, 1
        else:
            delta_t, future = 2.0 * c.dt_atmos, self.previous
        p, cur = self.previous, self.current
        pv = self.vorg[cur] + self.coriolis[:, None]
        tend_u = pv * self.v[cur]
        tend_v = -pv * self.u[cur]
        dt_vors, _ = tr.vor_div_from_uv_grid(tend_u, tend_v)
        dt_vors = self._damp(self.vors[p], dt_vors, delta_t)
        self._leapfrog(self.vors, dt_vors, p, cur, future, delta_t, c.robert_coeff, c.raw_filter_coeff)
        self.vorg[future] = tr.spherical_to_grid(self.vors[future])
        self.u[future], self.v[future] = tr.uv_grid_from_vor_div(self.vors[future], np.zeros_like(self.vors[future]))
        if self.v.min() < c.valid_range_v[0] or self.v.max() > c.valid_range_v[1]:
            raise FloatingPointError("barotropic_dynamics:  Meridional wind out of valid range.")
        if c.spec_tracer:                                           # update_spec_tracer (:393-412)
            dt_tr = tr.horizontal_advection(self.trs_s[cur], self.u[cur], self.v[cur], np.zeros_like(self.u[cur]))
            dt_trs = self._damp(self.trs_s[p], tr.grid_to_spherical(dt_tr), delta_t)
            self._leapfrog(self.trs_s, dt_trs, p, cur, future, delta_t, c.robert_coeff, c.raw_filter_coeff)
            self.trs_g[future] = tr.spherical_to_grid(self.trs_s[future])
        eigen = self.tb.eigen_laplacian
        with np.errstate(divide="ignore"):
            factor = np.where(eigen != 0.0, 1.0 / np.where(eigen != 0.0, -eigen, 1.0), 0.0)   # compute_laplacian(., -1)
        self.stream = tr.spherical_to_grid(self.vors[cur] * factor)
        self.previous, self.current = cur, future
        self.first = False
        enstrophy = tr.area_weighted_global_mean(self.vorg[self.current] * self.vorg[self.previous])
        energy = -tr.area_weighted_global_mean(self.stream * self.vorg[self.previous])
        return energy, enstrophy
