"""GPU parity of the column-physics kernels (through the C ABI, include/isca_b200_physics.h) against the CPU oracle
(oracle/physics.py) on the same seeded columns.  Tolerance: 1e-12 relative to the field maximum (the kernels follow
the reference operation order; the residual is fused-multiply-add contraction and libm exp/pow differences)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def columns(K, J, I, seed):
    rng = np.random.default_rng(seed)
    ps = 1.0e5 + 2.0e3 * rng.standard_normal((J, I))
    sig_h = np.linspace(0.0, 1.0, K + 1) ** 1.5
    ph = sig_h[:, None, None] * ps[None]
    pf = 0.5 * (ph[1:] + ph[:-1])
    t = 210.0 + 85.0 * (pf / 1.0e5) + 3.0 * rng.standard_normal((K, J, I))
    lat = np.repeat(np.linspace(-1.5, 1.5, J)[:, None], I, 1)
    return rng, ps, ph, pf, t, lat


@pytest.fixture(scope="module")
def mods(lib_built):
    from isca_b200 import physics
    from oracle import physics as O
    return physics, O


def test_sat_vapor_pres_tables(mods):
    physics, O = mods
    cp = physics.ColumnPhysics(16, 8, 10)
    s = O.SatVaporPres()
    T = np.random.default_rng(0).uniform(101.0, 620.0, 100001)
    T[:5] = [s.tminl, s.tminl + 0.05, 273.16, 623.1, 623.16]            # bin edges and the last table entry
    es, des = cp.lookup_es_des(T)
    eo, do = s.lookup_es_des(T)
    assert np.max(np.abs(es / eo - 1)) < 1e-13 and np.max(np.abs(des / do - 1)) < 1e-13
    pr = np.random.default_rng(1).uniform(100.0, 1.05e5, T.shape)
    qs, dqs = cp.compute_qs(T, pr)
    qo, dqo = s.compute_qs(T, pr)
    assert np.max(np.abs(qs / qo - 1)) < 1e-12 and np.max(np.abs(dqs / dqo - 1)) < 1e-12
    with pytest.raises(physics.IscaError) as e:                        # reference: table overflow is FATAL
        cp.lookup_es_des(np.array([300.0, 90.0]))
    assert "table" in str(e.value)
    with pytest.raises(physics.IscaError):
        cp.lookup_es_des(np.array([700.0]))
    with pytest.raises(physics.IscaError):
        cp.lookup_es_des(np.array([np.nan]))
    es2, _ = cp.lookup_es_des(T[:10])                                 # the handle stays usable after an error
    assert np.array_equal(es2, es[:10])


@pytest.mark.parametrize("K,J,I,evap,hc", [(12, 6, 8, 0, 1.0), (25, 32, 64, 1, 1.0), (40, 64, 128, 1, 0.8), (1, 3, 5, 1, 1.0)])
def test_lscale_cond(mods, K, J, I, evap, hc):
    physics, O = mods
    rng, ps, ph, pf, t, lat = columns(K, J, I, K)
    s = O.SatVaporPres()
    qs, _ = s.compute_qs(t, pf, hc)
    q = qs * rng.uniform(0.3, 1.4, size=t.shape)
    q[:, 0, 0] = 0.0
    q[:, -1, -1] = 0.999 * qs[:, -1, -1]                               # just below saturation: no adjustment
    cp = physics.ColumnPhysics(I, J, K, do_evap=evap, hc=hc)
    rain, tdel, qdel = cp.lscale_cond(t, q, pf, ph)
    ro, to, qo = O.lscale_cond(s, t, q, pf, ph, hc=hc, do_evap=bool(evap))
    assert rel(tdel, to) < TOL and rel(qdel, qo) < TOL and rel(rain, ro) < TOL
    assert np.array_equal(qdel == 0.0, qo == 0.0)                      # same layers adjusted
    assert np.all(qdel[:, -1, -1] == 0.0)
    with pytest.raises(physics.IscaError):
        bad = t.copy(); bad[K // 2, 1, 1] = 20.0
        cp.lscale_cond(bad, q, pf, ph)
    with pytest.raises(physics.IscaError):
        cp.lscale_cond(t[:, :, :-1], q, pf, ph)


@pytest.mark.parametrize("K,J,I", [(10, 8, 16), (40, 64, 128), (80, 4, 32)])
@pytest.mark.parametrize("nml", [dict(), dict(atm_abs=0.2, sw_diff=0.1, del_sw=0.05, odp=1.3, linear_tau=0.2, wv_exponent=3.5, solar_exponent=2.0,
                                                  diabatic_acce=2.0)])
def test_two_stream_gray_rad(mods, K, J, I, nml):
    physics, O = mods
    rng, ps, ph, pf, t, lat = columns(K, J, I, 7 + K)
    alb = rng.uniform(0.1, 0.4, (J, I))
    ts = t[-1] + rng.uniform(-3, 3, (J, I))
    tdt0 = 1e-5 * rng.standard_normal(t.shape)
    cp = physics.ColumnPhysics(I, J, K, **nml)
    g = O.GreyRadiation(O.GreyRadConfig(**nml))
    d = g.down(lat, ph, t)
    sw, lw = cp.two_stream_gray_rad_down(lat, ph, t, alb)
    assert rel(lw, d["surf_lw_down"]) < TOL and rel(sw, (1 - alb) * d["sw_down_surf"]) < TOL
    tdt, olr = cp.two_stream_gray_rad_up(lat, ph, t, ts, alb, tdt0)
    to, o = g.up(ts, alb, ph, tdt0)
    assert rel(olr, o["olr"]) < TOL
    assert rel(tdt - tdt0, to - tdt0) < 1e-11                          # difference of O(100 W/m2) fluxes
    col = ((tdt - tdt0) / g.c.diabatic_acce * O.CP_AIR * (ph[1:] - ph[:-1]) / O.GRAV).sum(0)
    assert np.allclose(col, o["rad_flux"][-1] - o["rad_flux"][0], rtol=1e-9)


@pytest.mark.parametrize("conserve", [1, 0])
def test_rayleigh_damping(mods, conserve):
    physics, O = mods
    K, J, I = 30, 16, 32
    rng = np.random.default_rng(3)
    pref = np.append(1.0e5 * ((np.arange(K) + 0.5) / K) ** 3, 1.0e5)
    pf = pref[:K, None, None] * rng.uniform(0.9, 1.1, (K, J, I))
    u, v = 30 * rng.standard_normal((K, J, I)), 10 * rng.standard_normal((K, J, I))
    for tray in (-0.25, 3600.0):
        cp = physics.ColumnPhysics(I, J, K, trayfric=tray, sponge_pbottom=5000.0, do_conserve_energy=conserve)
        udt, vdt, tdt = cp.rayleigh_damping(600.0, pf, u, v, pref)
        uo, vo, to, nlev = O.rayleigh_sponge(600.0, pf, u, v, pref, 5000.0, tray, bool(conserve))
        assert 0 < nlev < K
        assert rel(udt, uo) < 1e-14 and rel(vdt, vo) < 1e-14
        assert (rel(tdt, to) < 1e-13) if conserve else np.all(tdt == 0)
        assert np.array_equal(udt == 0, uo == 0)


def test_physics_kernels_full_size_timing(mods):
    """T170 window (512 x 256 x 40): the kernels run and report a plausible streaming rate."""
    physics, _ = mods
    cp = physics.ColumnPhysics(512, 256, 40, do_evap=1)
    for which in range(4):
        ms, by = cp.time_kernel(which, reps=10)
        assert ms > 0 and by / (ms * 1e-3) / 1e9 > 50.0, (which, ms, by)
