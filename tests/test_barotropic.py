"""The barotropic vorticity-equation model (src/atmos_spectral_barotropic) on the transform-level ABI.

CPU: the host driver isca_b200/barotropic.py (leapfrog + RAW filter, implicit damping, tracer, stream function, first forward step)
against the independent restatement oracle/barotropic.py, with the driver's GPU transform engine replaced -- in the test only -- by
an adapter around the checker's transforms; conservation properties of the restatement.  GPU: the real engine."""
import numpy as np
import pytest

from oracle.barotropic import BarotropicConfig, BarotropicModel

T21 = dict(num_lon=64, num_lat=32, num_fourier=21, num_spherical=22)


class _CheckerEngine:
    """test-only stand-in for api.Atmosphere's transforms_mod-level methods"""

    def __init__(self, nml):
        from oracle.isca_oracle import Config, Tables, Transforms
        cfg = Config(lon_max=nml["num_lon"], lat_max=nml["num_lat"], num_fourier=nml["num_fourier"], num_spherical=nml["num_spherical"])
        self.tb = Tables(cfg)
        self.tr = Transforms(self.tb)
        self.radius, self.omega = cfg.radius, cfg.omega

    def get_table(self, tid):
        from isca_b200 import api
        return {api.TB_SIN_LAT: self.tb.sin_lat, api.TB_WTS_LAT: self.tb.wts_lat, api.TB_DEG_LAT: self.tb.deg_lat,
                api.TB_DEG_LON: self.tb.deg_lon}[tid].copy()

    def trans_spherical_to_grid(self, s):
        return self.tr.spherical_to_grid(s)

    def trans_grid_to_spherical(self, g):
        return self.tr.grid_to_spherical(g)

    def uv_grid_from_vor_div(self, vor, div):
        return self.tr.uv_grid_from_vor_div(vor, div)

    def vor_div_from_uv_grid(self, u, v):
        return self.tr.vor_div_from_uv_grid(u, v)

    def atmosphere_end(self):
        pass


def _compare(m, o, tol):
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(m.vor_spec[m.current], o.vors[o.current]) < tol
    assert rel(m.u[m.current], o.u[o.current]) < tol and rel(m.v[m.current], o.v[o.current]) < tol
    assert rel(m.vor[m.previous], o.vorg[o.previous]) < tol
    assert rel(m.trs[m.current], o.trs_g[o.current]) < tol
    assert rel(m.stream, o.stream) < tol


@pytest.mark.parametrize("kw", [{}, dict(raw_filter_coeff=0.53, damping_option="resolution_independent", damping_coeff=1e30, damping_order=4,
                                         damping_coeff_r=1.0e-7, initial_zonal_wind="zero", zeta_0=2.0e-4, m_0=3)])
def test_driver_logic_matches_restatement(monkeypatch, kw):
    from isca_b200 import barotropic
    monkeypatch.setattr(barotropic, "_make_engine", lambda nml: (lambda e: (e, e.radius, e.omega))(_CheckerEngine(nml)))
    m = barotropic.BarotropicAtmosphere(1800.0, **T21, **kw)
    o = BarotropicModel(BarotropicConfig(dt_atmos=1800.0, **T21, **kw))
    _compare(m, o, 1e-12)
    for step in range(30):
        m.atmosphere(1)
        e, z = o.step()
        assert (m.previous, m.current) == (o.previous, o.current)
        _compare(m, o, 1e-10)
        assert abs(m.energy - e) < 1e-10 * abs(e) + 1e-300 and abs(m.enstrophy - z) < 1e-10 * abs(z)
    with pytest.raises(barotropic.IscaError):
        barotropic.BarotropicAtmosphere(1800.0, grid_tracer=True)
    with pytest.raises(barotropic.IscaError):
        barotropic.BarotropicAtmosphere(1800.0, not_a_namelist_variable=1)


def test_restatement_conserves_energy_and_enstrophy_without_damping():
    """the undamped barotropic vorticity equation conserves kinetic energy and enstrophy; the leapfrog + weak Robert filter keeps
    both to 1e-4 over two days at T21 (the reference prints exactly these two diagnostics every print_interval)"""
    o = BarotropicModel(BarotropicConfig(dt_atmos=900.0, damping_coeff=0.0, robert_coeff=0.01, spec_tracer=False, **T21))
    e0, z0 = o.step()
    for _ in range(191):
        e, z = o.step()
    assert abs(e - e0) / e0 < 1e-4 and abs(z - z0) / z0 < 2e-3
    # solid-body rotation is a steady solution: zero tendency
    s = BarotropicModel(BarotropicConfig(dt_atmos=900.0, damping_coeff=0.0, zeta_0=0.0, initial_zonal_wind="zero", spec_tracer=False, **T21))
    s.step()
    assert np.abs(s.u).max() < 1e-12
