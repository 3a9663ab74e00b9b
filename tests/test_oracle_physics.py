"""CPU checks of the column-physics oracle (oracle/physics.py) through properties the schemes guarantee:
Clausius-Clapeyron tables, enthalpy/water conservation of the large-scale condensation, flux-divergence form of the
grey radiation, sign/limits of the Rayleigh sponge."""
import numpy as np
import pytest
from oracle import physics as P


def columns(K=12, J=6, I=8, seed=0):
    rng = np.random.default_rng(seed)
    ps = 1.0e5 + 2.0e3 * rng.standard_normal((J, I))
    sig_h = np.linspace(0.0, 1.0, K + 1) ** 1.5
    ph = sig_h[:, None, None] * ps[None]
    pf = 0.5 * (ph[1:] + ph[:-1])
    t = 210.0 + 85.0 * (pf / 1.0e5) + 3.0 * rng.standard_normal((K, J, I))
    return rng, ps, ph, pf, t


def test_sat_vapor_tables_clausius_clapeyron():
    s = P.SatVaporPres()
    assert s.table_siz == 5231 and abs(s.dtres - 0.1) < 1e-15
    T = np.linspace(150.0, 600.0, 977)
    es, des = s.lookup_es_des(T)
    exact = 610.78 * np.exp(-P.HLV / P.RVGAS * (1.0 / T - 1.0 / P.TFREEZE))
    assert np.max(np.abs(es / exact - 1.0)) < 1e-6          # 2nd-order Taylor inside a 0.1 K bin
    assert np.max(np.abs(des / (P.HLV * exact / P.RVGAS / T ** 2) - 1.0)) < 1e-4
    assert abs(s.lookup_es_des(np.array([P.TFREEZE]))[0][0] - 610.78) < 1e-9
    with pytest.raises(FloatingPointError):
        s.lookup_es_des(np.array([50.0]))
    with pytest.raises(FloatingPointError):
        s.lookup_es_des(np.array([700.0]))


def test_lscale_cond_conserves_enthalpy_and_water():
    rng, ps, ph, pf, t = columns()
    s = P.SatVaporPres()
    qs, _ = s.compute_qs(t, pf)
    q = qs * rng.uniform(0.3, 1.4, size=t.shape)            # sub- and super-saturated layers
    for evap in (False, True):
        rain, tdel, qdel = P.lscale_cond(s, t, q, pf, ph, do_evap=evap)
        assert np.allclose(P.CP_AIR * tdel + P.HLV * qdel, 0.0, atol=1e-9)
        pmass = (ph[1:] - ph[:-1]) / P.GRAV
        assert np.allclose(rain, np.maximum(-(pmass * qdel).sum(0), 0.0), rtol=1e-13)
        assert np.all(rain >= 0.0)
        sup = q > qs
        assert np.all(qdel[sup & (qdel != 0)] < 0.0) if not evap else True
        # one Newton step towards saturation: the adjusted state is much closer to saturation than the input
        qs2, _ = s.compute_qs(t + tdel, pf)
        m = sup if not evap else (qdel < 0)
        assert np.max(np.abs((q + qdel)[m] / qs2[m] - 1.0)) < 0.02
    r0 = P.lscale_cond(s, t, q, pf, ph, do_evap=False)[0]
    r1 = P.lscale_cond(s, t, q, pf, ph, do_evap=True)[0]
    assert np.all(r1 <= r0 + 1e-15) and np.any(r1 < r0)     # re-evaporation can only reduce the surface rain


def test_grey_radiation_flux_divergence_and_limits():
    rng, ps, ph, pf, t = columns()
    J, I = ps.shape
    lat = np.repeat(np.linspace(-1.4, 1.4, J)[:, None], I, 1)
    cfg = P.GreyRadConfig(atm_abs=0.2, sw_diff=0.1, del_sw=0.05)
    g = P.GreyRadiation(cfg)
    d = g.down(lat, ph, t)
    ts = t[-1] + 2.0
    alb = np.full((J, I), 0.3)
    tdt, o = g.up(ts, alb, ph, np.zeros_like(t))
    # column-integrated heating == net flux through the boundaries
    col = (tdt * P.CP_AIR * (ph[1:] - ph[:-1]) / P.GRAV).sum(0)
    assert np.allclose(col, o["rad_flux"][-1] - o["rad_flux"][0], rtol=1e-12)
    assert np.all(d["surf_lw_down"] > 0) and np.all(d["surf_lw_down"] < P.STEFAN * t.max() ** 4)
    assert np.all(o["olr"] > 0)
    # isothermal column with the surface at the same temperature emits exactly sigma T^4 upward
    tiso = np.full_like(t, 250.0)
    g.down(lat, ph, tiso)
    _, o2 = g.up(np.full((J, I), 250.0), alb, ph, np.zeros_like(t))
    assert np.allclose(o2["olr"], P.STEFAN * 250.0 ** 4, rtol=1e-13)
    # transparent atmosphere: no shortwave heating
    g0 = P.GreyRadiation(P.GreyRadConfig(ir_tau_eq=0.0, ir_tau_pole=0.0))
    g0.down(lat, ph, t)
    tdt0, _ = g0.up(ts, alb, ph, np.zeros_like(t))
    assert np.max(np.abs(tdt0)) < 1e-18


def test_rayleigh_sponge_levels_and_energy():
    K, J, I = 20, 4, 6
    rng = np.random.default_rng(3)
    pref = np.append(1.0e5 * (np.arange(K) + 0.5) / K * 0.2, 1.0e5)     # top-heavy reference column
    pf = np.repeat(np.repeat(pref[:K, None, None], J, 1), I, 2) * rng.uniform(0.9, 1.1, (K, J, I))
    u, v = 30 * rng.standard_normal((K, J, I)), 10 * rng.standard_normal((K, J, I))
    dt = 600.0
    udt, vdt, tdt, nlev = P.rayleigh_sponge(dt, pf, u, v, pref, sponge_pbottom=5000.0, trayfric=-0.25)
    assert nlev == int(np.argmin(np.abs(pref - 1.0e4))) + 1
    assert np.all(udt[nlev:] == 0) and np.all(tdt[nlev:] == 0)
    assert np.all(udt[pf >= 5000.0] == 0)
    damped = (pf < 5000.0) & (np.arange(K)[:, None, None] < nlev)
    assert np.all((udt * u)[damped] <= 0)
    # kinetic energy lost over the step reappears as heat
    ke0 = 0.5 * (u ** 2 + v ** 2)
    ke1 = 0.5 * ((u + dt * udt) ** 2 + (v + dt * vdt) ** 2)
    assert np.allclose((ke1 - ke0)[damped] / dt, -(P.CP_AIR * tdt)[damped], rtol=1e-11)
