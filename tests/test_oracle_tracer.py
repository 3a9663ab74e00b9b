"""Oracle pins for the grid-tracer path (fv_advection.F90 Lin-Rood A-grid advection, PPM vertical
advection, Held-Suarez tracer source/sink, water fixer): properties the schemes guarantee."""
import numpy as np
import pytest
from oracle.isca_oracle import SpectralCore, held_suarez_config, vert_advection, FINITE_VOLUME_PARABOLIC
from oracle.fv_advection import FVGrid, a_grid_horiz_advection, _integer_flux_x


@pytest.fixture(scope="module")
def core():
    cfg = held_suarez_config("T21", 8, 1200.0, num_tracers=1)
    cfg.initial_sphum = 2.0e-3
    c = SpectralCore(cfg)
    c.cold_start()
    for _ in range(40):
        c.step()
    return c


def test_constant_is_preserved_and_scheme_is_shift_invariant(core):
    g = core.fv
    ua, va = core.ug[core.current], core.vg[core.current]
    dq = a_grid_horiz_advection(g, ua, va, np.full_like(ua, 3.0), 1200.0, np.zeros_like(ua))
    assert np.abs(dq).max() < 1e-18                       # q*div term cancels the flux divergence of a constant
    rng = np.random.default_rng(0)
    q = 1.0 + 0.2 * rng.random(ua.shape)
    d1 = a_grid_horiz_advection(g, ua, va, q, 1200.0, np.zeros_like(q))
    d2 = a_grid_horiz_advection(g, ua, va, q + 5.0, 1200.0, np.zeros_like(q))
    assert np.abs(d1 - d2).max() < 1e-15                  # adding a constant changes nothing


def test_monotone_no_new_extrema(core):
    g = core.fv
    ua, va = core.ug[core.current], core.vg[core.current]
    rng = np.random.default_rng(1)
    q = rng.random(ua.shape)
    dt = 1200.0
    q_new = q + dt * a_grid_horiz_advection(g, 5 * ua, 5 * va, q, dt, np.zeros_like(q))
    assert q_new.min() > -0.05 and q_new.max() < 1.05     # van Leer limited: no significant over/undershoot


def test_integer_flux_large_courant():
    nx = 16
    q = np.arange(nx, dtype=float)[None, None, :] ** 2
    c = np.full((1, 1, nx), 2.5)
    f = _integer_flux_x(c, q)
    i = 5                                                  # 0-based; Fortran i = 6: sum(q(i-2:i-1))
    assert f[0, 0, i] == q[0, 0, i - 2] + q[0, 0, i - 1]
    c = np.full((1, 1, nx), -1.5)
    f = _integer_flux_x(c, q)
    assert f[0, 0, i] == -q[0, 0, i]


def test_ppm_bounds():
    rng = np.random.default_rng(2)
    K = 12
    dz = 900.0 + 300 * rng.random((K, 2, 3))
    w = np.zeros((K + 1, 2, 3)); w[1:K] = 1.5 * rng.standard_normal((K - 1, 2, 3))
    r = rng.random((K, 2, 3))
    dt = 200.0
    r_new = r + dt * vert_advection(dt, w, dz, r, FINITE_VOLUME_PARABOLIC)
    assert r_new.min() > -0.02 and r_new.max() < 1.02


def test_hs_tracer_source_sink_and_water_fixer(core):
    cfg = core.cfg
    c = SpectralCore(cfg); c.cold_start()
    w0 = c.mass_weighted_global_integral(c.grid_tracers[c.current, 0], c.psg[c.current])
    c.step()
    w1 = c.mass_weighted_global_integral(c.grid_tracers[c.current, 0], c.psg[c.current])
    dt = cfg.dt_atmos
    expected = w0 + dt * (cfg.trflux / cfg.grav - w0 / (4 * 86400.0))      # source flux/pmass integrates to flux/g
    assert abs(w1 - expected) / w0 < 1e-10                 # the water fixer removes the advection error exactly
    assert np.isfinite(core.grid_tracers).all() and core.grid_tracers.min() >= 0.0
