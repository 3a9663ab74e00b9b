"""The reference's barotropic vorticity-equation model (src/atmos_spectral_barotropic/) on the transform-level C ABI.

`BarotropicAtmosphere` mirrors `atmosphere_mod` of that model (atmosphere.F90:96-238: atmosphere_init / atmosphere /
atmosphere_end) and `barotropic_dynamics_mod` (barotropic_dynamics.F90:166-437).  Every spherical-harmonic transform of a step --
vor_div_from_uv_grid, trans_spherical_to_grid (x3), uv_grid_from_vor_div (x2), trans_grid_to_spherical -- runs in the CUDA library
through the `transforms_mod`-level entry points of include/isca_b200.h; the per-coefficient algebra
between them (implicit spectral damping, leapfrog with the Robert-Asselin-Williams filter, inverse Laplacian) is a few vector
operations on (num_spherical+1) x (num_fourier+1) complex arrays and stays on the host, as the reference keeps it in Fortran.
There is no CPU transform path: constructing the model needs the CUDA library and a device.

Not built (rejected at construction): the finite-volume grid tracer (`grid_tracer = .true.` needs a_grid_horiz_advection as a stand-alone
entry point), `stirring` with non-zero amplitude, `damping_option = 'exponential_cutoff'`, restarts."""
from __future__ import annotations
import numpy as np
from .api import IscaError

NML_DEFAULTS = dict(num_lon=256, num_lat=128, num_fourier=85, num_spherical=86, robert_coeff=0.04, raw_filter_coeff=1.0,
                    damping_option="resolution_dependent", damping_order=4, damping_coeff=1.0e-04, damping_coeff_r=0.0, cutoff_wn=30,
                    zeta_0=8.0e-05, m_0=4, eddy_width=15.0, eddy_lat=45.0, spec_tracer=True, grid_tracer=False,
                    initial_zonal_wind="two_jets", valid_range_v=(-1.0e3, 1.0e3))   # barotropic_dynamics_nml (:110-146)


def _make_engine(nml):
    """the transform engine: a dynamical-core handle of the same horizontal resolution (its transforms_mod-level entry points)"""
    from . import api
    cfg = api.make_config(lon_max=nml["num_lon"], lat_max=nml["num_lat"], num_fourier=nml["num_fourier"], num_spherical=nml["num_spherical"],
                          num_levels=10, do_water_correction=False)   # no tracer in the engine (spectral_dynamics.F90:1264)
    return api.Atmosphere(cfg), cfg.radius, cfg.omega


class BarotropicAtmosphere:
    def __init__(self, dt_atmos: float, **nml):
        bad = set(nml) - set(NML_DEFAULTS)
        if bad:
            raise IscaError(f"unknown barotropic_dynamics_nml variable(s) {sorted(bad)}")
        self.nml = n = dict(NML_DEFAULTS, **nml)
        if n["grid_tracer"]:
            raise IscaError("barotropic_dynamics: grid_tracer = .true. is not built (no stand-alone a_grid_horiz_advection entry point)")
        if n["damping_option"] not in ("resolution_dependent", "resolution_independent"):
            raise IscaError('spectral_damping_init: "%s" is an invalid (or not built) value for damping_option' % n["damping_option"])
        if n["initial_zonal_wind"] not in ("zero", "two_jets"):
            raise IscaError("barotropic_dynamics_init: %s is not a valid value of initial_zonal_wind" % n["initial_zonal_wind"])
        self.dt_real = float(dt_atmos)
        self.eng, self.radius, self.omega = _make_engine(n)
        from . import api
        sin_lat = self.eng.get_table(api.TB_SIN_LAT)
        self.wts_lat = self.eng.get_table(api.TB_WTS_LAT)
        self.deg_lat = self.eng.get_table(api.TB_DEG_LAT)
        self.deg_lon = self.eng.get_table(api.TB_DEG_LON)
        self.sin_lat, self.cos_lat = sin_lat, np.sqrt(1.0 - sin_lat * sin_lat)
        self.coriolis = 2 * self.omega * sin_lat
        M, N = n["num_fourier"], n["num_spherical"]
        L = np.arange(N + 1)[:, None] + np.arange(M + 1)[None, :]              # total wavenumber, [n, m]
        self.eigen = L * (L + 1.0) / (self.radius * self.radius)
        ref = self.eigen[N - 1, 0]                                             # eigen(0, num_spherical-1)
        if n["damping_option"] == "resolution_dependent":
            self.damping = n["damping_coeff"] * (self.eigen / ref) ** n["damping_order"]
        else:
            self.damping = n["damping_coeff"] * self.eigen ** n["damping_order"]
        self.damping = self.damping + n["damping_coeff_r"]
        self.inv_lap = np.zeros_like(self.eigen)
        self.inv_lap[self.eigen != 0.0] = -1.0 / self.eigen[self.eigen != 0.0]
        J, I = n["num_lat"], n["num_lon"]
        self.vor_spec = np.zeros((2, N + 1, M + 1), dtype=np.complex128)       # Dyn%Spec%vor(:,:,1:2)
        self.u, self.v, self.vor = np.zeros((2, J, I)), np.zeros((2, J, I)), np.zeros((2, J, I))
        self.trs_spec = np.zeros((2, N + 1, M + 1), dtype=np.complex128)
        self.trs = np.zeros((2, J, I))
        self.stream = np.zeros((J, I))
        self._initial_state()
        self.previous = self.current = 0                                       # forward step first (atmosphere.F90:140-143)
        self._first = True
        self.energy = self.enstrophy = float("nan")

    # ---- transforms (all on the GPU) ----
    def _s2g(self, s):
        return self.eng.trans_spherical_to_grid(s)

    def _g2s(self, g):
        return self.eng.trans_grid_to_spherical(g)

    def _uv(self, vor, div=None):
        z = np.zeros_like(vor) if div is None else div
        u, v = self.eng.uv_grid_from_vor_div(vor[None], z[None])
        return u[0], v[0]

    def _vordiv(self, u, v):
        vor, div = self.eng.vor_div_from_uv_grid(u[None], v[None])
        return vor[0], div[0]

    def _grad(self, s):
        """horizontal gradient of a spectral scalar: the irrotational wind of the velocity potential s, i.e.
        uv_grid_from_vor_div(0, laplacian(s)) (transforms.F90:808-831 evaluates the same derivatives with compute_gradient_cos)"""
        return self._uv(np.zeros_like(s), -self.eigen * s)

    def _initial_state(self):
        n = self.nml
        if n["initial_zonal_wind"] == "two_jets":
            c, s = self.cos_lat, self.sin_lat
            u0 = 25.0 * c - 30.0 * c ** 3 + 300.0 * s ** 2 * c ** 6
        else:
            u0 = np.zeros_like(self.sin_lat)
        u = np.repeat(u0[:, None], n["num_lon"], 1)
        vor, _ = self._vordiv(u, np.zeros_like(u))
        g = self._s2g(vor)
        yy = (self.deg_lat - n["eddy_lat"]) / n["eddy_width"]
        rad_lon = self.deg_lon * np.arctan(1.0) / 45.0
        g = g + 0.5 * n["zeta_0"] * self.cos_lat[:, None] * np.exp(-yy * yy)[:, None] * np.cos(n["m_0"] * rad_lon)[None, :]
        self.vor[0] = g
        self.vor_spec[0] = self._g2s(g)
        self.u[0], self.v[0] = self._uv(self.vor_spec[0])
        if n["spec_tracer"]:
            t = np.zeros_like(g)
            t[(self.deg_lat > 10.0) & (self.deg_lat < 20.0)] = 1.0
            t[self.deg_lat > 70.0] = -1.0
            self.trs[0] = t
            self.trs_spec[0] = self._g2s(t)

    def _advance(self, a, tend, prev, cur, fut, delta_t):
        """implicit spectral damping (spectral_damping.F90:172-195) + leapfrog with the RAW filter (leapfrog.F90:217-247)"""
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
        for _ in range(n_steps):
            if self._first:
                delta_t, fut = self.dt_real, 1
            else:
                delta_t, fut = 2.0 * self.dt_real, self.previous
            prev, cur = self.previous, self.current
            pv = self.vor[cur] + self.coriolis[:, None]
            dt_vors, _ = self._vordiv(pv * self.v[cur], -pv * self.u[cur])
            self._advance(self.vor_spec, dt_vors, prev, cur, fut, delta_t)
            self.vor[fut] = self._s2g(self.vor_spec[fut])
            self.u[fut], self.v[fut] = self._uv(self.vor_spec[fut])
            lo, hi = self.nml["valid_range_v"]
            if self.v.min() < lo or self.v.max() > hi:
                raise IscaError("barotropic_dynamics:  Meridional wind out of valid range.")
            if self.nml["spec_tracer"]:
                dx, dy = self._grad(self.trs_spec[cur])
                dt_trs = self._g2s(-self.u[cur] * dx - self.v[cur] * dy)
                self._advance(self.trs_spec, dt_trs, prev, cur, fut, delta_t)
                self.trs[fut] = self._s2g(self.trs_spec[fut])
            self.stream = self._s2g(self.vor_spec[cur] * self.inv_lap)
            self.previous, self.current = cur, fut
            self._first = False
        w = self.wts_lat[:, None] / (self.wts_lat.sum() * self.nml["num_lon"])   # area_weighted_global_mean (transforms.F90:1059-1077)
        self.enstrophy = float((w * self.vor[self.current] * self.vor[self.previous]).sum())
        self.energy = float(-(w * self.stream * self.vor[self.previous]).sum())

    def atmosphere_end(self):
        if self.eng is not None:
            self.eng.atmosphere_end()
            self.eng = None
