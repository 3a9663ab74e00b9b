"""The shallow-water model (src/atmos_spectral_shallow) on the transform-level ABI: host driver isca_b200/shallow.py against the
restatement oracle/shallow.py (driver's GPU transform engine replaced, in the test only, by an adapter around the checker's
transforms), and properties of the restatement.  The GPU test lives in tests/test_gpu_rows_f.py."""
import numpy as np
import pytest

from oracle.shallow import ShallowConfig, ShallowModel
from test_barotropic import _CheckerEngine, T21


def compare(m, o, tol):
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    for k, oa in (("vor", o.vors), ("div", o.divs), ("h", o.hs)):
        assert rel(m.spec[k][m.current], oa[o.current]) < tol, k
    for k, oa in (("u", o.u), ("v", o.v), ("h", o.h), ("trs", o.trs_g)):
        assert rel(m.grid[k][m.current], oa[o.current]) < tol, k
    assert rel(m.grid["h"][m.previous], o.h[o.previous]) < tol


@pytest.mark.parametrize("kw", [{}, dict(u_deep_mag=20.0, u_upper_mag_init=30.0, raw_filter_coeff=0.53, h_0=2.0e4,
                                         damping_option="resolution_independent", damping_coeff=1e30)])
def test_driver_logic_matches_restatement(monkeypatch, kw):
    from isca_b200 import barotropic, shallow
    monkeypatch.setattr(shallow, "_make_engine", lambda nml: (lambda e: (e, e.radius, e.omega))(_CheckerEngine(nml)))
    m = shallow.ShallowAtmosphere(900.0, **T21, **kw)
    o = ShallowModel(ShallowConfig(dt_atmos=900.0, **T21, **kw))
    compare(m, o, 1e-12)
    for step in range(30):
        m.atmosphere(1)
        o.step()
        assert (m.previous, m.current) == (o.previous, o.current)
        compare(m, o, 1e-9)
    a, b = m.global_diag(), o.global_diag()
    assert all(abs(x - y) <= 1e-9 * abs(y) + 1e-300 for x, y in zip(a, b))
    with pytest.raises(shallow.IscaError):
        shallow.ShallowAtmosphere(900.0, grid_tracer=True)
    with pytest.raises(shallow.IscaError):
        shallow.ShallowAtmosphere(900.0, physics_nml=dict(nonsense=1))


def test_restatement_properties():
    """(i) a resting fluid of uniform depth without forcing stays at rest; (ii) with the default forcing (mass source in the ITCZ
    and a subtropical bump, relaxation on 10 days) the flow spins up, stays sub-critical and conserves nothing it should not:
    the global-mean depth follows d<h>/dt = -kappa_t (<h> - <h_eq>) exactly (mass source only through the relaxation)"""
    rest = ShallowModel(ShallowConfig(dt_atmos=900.0, fric_damp_time=0.0, therm_damp_time=0.0, spec_tracer=False, **T21))
    for _ in range(10):
        rest.step()
    assert np.abs(rest.u).max() < 1e-9 and np.abs(rest.h - 3.0e4).max() < 1e-6
    o = ShallowModel(ShallowConfig(dt_atmos=900.0, spec_tracer=False, **T21))
    mean = lambda f: o.tr.area_weighted_global_mean(f)
    for _ in range(96):
        o.step()
    ens, div2, fr = o.global_diag()
    assert ens > 0 and np.isfinite(fr) and fr < 1.0
    # mean-depth budget over one leapfrog step (centred difference around `current`)
    hp = mean(o.h[o.previous])
    o.step()
    hc_prev = mean(o.h[o.previous])          # the level that was `current` during the step
    o2 = mean(o.h[o.current])
    # h(future) - h(previous) = 2 dt * (-kappa_t (h(previous) - h_eq) + dynamics with zero global mean) up to the Robert filter
    lhs = (o2 - hp) / (2 * 900.0)
    rhs = -o.kappa_t * (hp - mean(o.h_eq))
    assert abs(lhs - rhs) < 2e-3 * abs(rhs)
