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


@pytest.mark.parametrize("scheme", ["byrne", "geen", "schneider"])
def test_grey_radiation_variants_flux_divergence_and_limits(scheme):
    """rad_scheme = byrne | geen | schneider (two_stream_gray_rad.F90:458-700): flux-divergence form, limits, scheme features."""
    rng, ps, ph, pf, t = columns()
    J, I = ps.shape
    lat = np.repeat(np.linspace(-1.4, 1.4, J)[:, None], I, 1)
    q = 0.02 * (pf / ps[None]) ** 3 * rng.uniform(0.2, 1.0, t.shape)
    alb = np.full((J, I), 0.3)
    ts = t[-1] + 2.0
    g = P.GreyRadiation(P.GreyRadConfig(rad_scheme=scheme, atm_abs=0.2))
    d = g.down(lat, ph, t, q=q, albedo=alb)
    tdt, o = g.up(ts, alb, ph, np.zeros_like(t))
    col = (tdt * P.CP_AIR * (ph[1:] - ph[:-1]) / P.GRAV).sum(0)
    assert np.allclose(col, o["rad_flux"][-1] - o["rad_flux"][0], rtol=1e-12)
    assert np.all(d["surf_lw_down"] > 0) and np.all(o["olr"] > 0) and np.all(d["sw_down_surf"] > 0)
    if scheme in ("byrne", "geen"):
        # more water vapour -> more back radiation, less outgoing longwave (surface warmer than the air above)
        g2 = P.GreyRadiation(P.GreyRadConfig(rad_scheme=scheme, atm_abs=0.2))
        d2 = g2.down(lat, ph, t, q=2.0 * q, albedo=alb)
        assert np.all(d2["surf_lw_down"] > d["surf_lw_down"])
        # isothermal column + surface at the same temperature: blackbody emission whatever the optical depths
        tiso = np.full_like(t, 250.0)
        g.down(lat, ph, tiso, q=q, albedo=alb)
        _, o2 = g.up(np.full((J, I), 250.0), alb, ph, np.zeros_like(t))
        assert np.allclose(o2["olr"], P.STEFAN * 250.0 ** 4, rtol=1e-13)
    if scheme == "geen":
        # the shortwave beam is attenuated by water vapour level by level: monotone, and weaker with more vapour
        assert np.all(np.diff(g2._st["sw_down"], axis=0) < 0) and np.all(d2["sw_down_surf"] < d["sw_down_surf"])
    if scheme == "schneider":
        # giant planet: the lower boundary returns exactly what it receives (no surface energy budget), insolation ~ cos(lat)
        assert np.allclose(o["rad_flux"][-1], 0.0, atol=1e-9)
        assert np.allclose(g._st["sw_down"][0], 1360.0 / np.pi * np.cos(lat) * (1 - g.gp_albedo), rtol=1e-13)


def test_grey_radiation_rejects_unknown_scheme():
    with pytest.raises(ValueError):
        P.GreyRadiation(P.GreyRadConfig(rad_scheme="rrtm"))


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


def diff_case(K=14, J=3, I=5, seed=5):
    rng, ps, ph, pf, t = columns(K, J, I, seed)
    q = 5e-3 * (pf / 1e5) ** 2 * rng.uniform(0.5, 1.5, t.shape)
    dlnp = np.log(ph[1:] / np.maximum(ph[:-1], 0.3 * ph[1]))
    zh = np.concatenate([np.cumsum((P.RDGAS * t * dlnp / P.GRAV)[::-1], 0)[::-1], np.zeros((1, J, I))])
    z = 0.5 * (zh[1:] + zh[:-1])
    u, v = 10 * rng.standard_normal(t.shape), 5 * rng.standard_normal(t.shape)
    dm, dh = rng.uniform(0.0, 30.0, t.shape), rng.uniform(0.0, 30.0, t.shape)
    dm[: K // 3] = 0.0                                     # no mixing aloft
    return rng, ph, pf, t, q, z, u, v, dm, dh


def test_vert_diff_is_the_implicit_tridiagonal_solve_and_conserves():
    rng, ph, pf, t, q, z, u, v, dm, dh = diff_case()
    K, J, I = t.shape
    zero2, zero3 = np.zeros((J, I)), np.zeros_like(t)
    delt = 900.0
    r = P.gcm_vert_diff_down(delt, u, v, t, q, dm, dh, ph, pf, z, zero2, zero2, zero2, zero2, zero3, zero3, zero3, zero3)
    dp = ph[1:] - ph[:-1]
    # zero surface stress: column momentum is conserved, dissipated kinetic energy reappears as heat
    assert np.abs((r["dt_u"] * dp).sum(0)).max() < 1e-12 * np.abs(u * dp).sum(0).max()
    ke0 = 0.5 * (u ** 2 + v ** 2)
    ke1 = 0.5 * ((u + delt * r["dt_u"]) ** 2 + (v + delt * r["dt_v"]) ** 2)
    assert np.allclose(((ke1 - ke0) / delt * dp).sum(0), -(P.CP_AIR * r["dissipative_heat"] * dp).sum(0), rtol=1e-10)
    assert np.all((r["dissipative_heat"] * dp).sum(0) >= 0)
    # dense solve of (I - delt*D) x = dt_explicit for u in one column
    j, i = 1, 2
    mu = P.compute_mu(ph)[:, j, i]; nu = P.compute_nu(dm, ph, pf, z, t, q, False)[:, j, i]
    A = np.zeros((K, K))
    for k in range(K):
        if k > 0:
            A[k, k - 1] += mu[k] * nu[k]; A[k, k] -= mu[k] * nu[k]
        if k < K - 1:
            A[k, k + 1] += mu[k] * nu[k + 1]; A[k, k] -= mu[k] * nu[k + 1]
    x = np.linalg.solve(np.eye(K) - delt * A, A @ u[:, j, i])
    assert np.allclose(r["dt_u"][:, j, i], x, rtol=1e-9, atol=1e-14)
    # closing the T/q system with zero surface flux (mixed layer bypassed) conserves column dry static energy and water
    tri = r["tri"]
    _, dT = P.diff_surface(tri["dtmass"], 0 * zero2, 0 * zero2, 0 * zero2, zero2, zero2, 1.0, tri["delta_t"].copy())
    dt_t, dt_q = P.gcm_vert_diff_up(delt, tri)
    # (delta_t already contains the nu*f term; with dflux + dflux_datmos = -nu(1-e) the closure divides by 1 - mu*dflux)
    closed = dict(tri)
    closed["delta_t"] = tri["delta_t"] / (1.0 - tri["dtmass"] * tri["dflux_t"])
    closed["delta_q"] = tri["delta_q"] / (1.0 - tri["dtmass"] * tri["dflux_q"])
    dt_t, dt_q = P.gcm_vert_diff_up(delt, closed)
    assert np.abs((dt_q * dp).sum(0)).max() < 1e-12 * np.abs(q * dp).sum(0).max()
    assert np.abs(((dt_t - r["dissipative_heat"]) * dp).sum(0)).max() < 1e-11 * np.abs(t * dp).sum(0).max() / delt


def test_mixed_layer_energy_balance():
    rng, ph, pf, t, q, z, u, v, dm, dh = diff_case()
    K, J, I = t.shape
    zero2, zero3 = np.zeros((J, I)), np.zeros_like(t)
    delt = 900.0
    tri = P.gcm_vert_diff_down(delt, u, v, t, q, dm, dh, ph, pf, z, zero2, zero2, zero2, zero2, zero3, zero3, zero3, zero3)["tri"]
    f = lambda lo, hi: rng.uniform(lo, hi, (J, I))
    ts = f(280, 300)
    args = dict(flux_t=f(-20, 60), flux_q=f(0, 1e-4), flux_r=f(350, 450), net_surf_sw_down=f(0, 300), surf_lw_down=f(250, 400),
                dhdt_surf=f(5, 20), dedt_surf=f(1e-6, 5e-6), dedq_surf=f(0, 1e-2), drdt_surf=f(4, 6), dhdt_atm=f(-20, -5),
                dedq_atm=f(-1e-2, -1e-3))
    cap = np.full((J, I), 40.0 * 1.035e3 * 3989.24495292815)
    ts2, tri2, d = P.mixed_layer(tri, 450.0, ts, heat_capacity=cap, ocean_qflux=zero2, **args)
    assert np.allclose(ts2 - ts, d)
    # implicit surface energy balance: C dTs/dt = SW + LW_down - LW_up(new) - SH(new) - LH(new), fluxes linearised about the old state
    dT1, dq1 = tri2["delta_t"], tri2["delta_q"]               # lowest-level increments consistent with the new surface
    sh = args["flux_t"] + args["dhdt_surf"] * d + args["dhdt_atm"] * dT1
    lh = P.HLV * (args["flux_q"] + args["dedt_surf"] * d + args["dedq_atm"] * dq1)
    lw = args["flux_r"] + args["drdt_surf"] * d
    assert np.allclose(cap * d / 450.0, args["net_surf_sw_down"] + args["surf_lw_down"] - lw - sh - lh, rtol=1e-9)
    # a huge heat capacity pins the surface
    ts3, _, d3 = P.mixed_layer(tri, 450.0, ts, heat_capacity=cap * 1e12, ocean_qflux=zero2, **args)
    assert np.max(np.abs(d3)) < 1e-9


def _kat():
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "monin_obukhov_kat.py")
    spec = importlib.util.spec_from_file_location("monin_obukhov_kat", path)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_monin_obukhov_matches_reference_self_test_checksums():
    """PIN: the reference's own known-answer test (monin_obukhov_kernel.F90:905-1120).  The restatement reproduces the
    Intel checksums exactly (distance 0 in units of the last bit)."""
    K = _kat()
    c = P.MOConfig(**K.NML)
    assert K.distance(K.checksum(P.mo_drag(c, K.PT, K.PT0, K.Z, K.Z0, K.ZT, K.ZQ, K.SPEED)), K.CHKSUM_DRAG) <= 2
    assert K.distance(K.checksum([P.mo_stable_mix(c, K.RICH)]), K.CHKSUM_STABLE_MIX) <= 2
    km, kh = P.mo_diff(c, np.array([[K.DIFF_Z]]), np.array([K.DIFF_USTAR]), np.array([K.DIFF_BSTAR]))
    assert K.distance(K.checksum([km, kh]), K.CHKSUM_DIFF) <= 2
    assert K.distance(K.checksum(P.mo_profile(c, K.ZREF, K.ZREF_T, K.Z, K.Z0, K.ZT, K.ZQ, K.U_STAR, K.B_STAR)), K.CHKSUM_PROFILE) <= 2


def surface_case(J=6, I=9, seed=2):
    rng = np.random.default_rng(seed)
    f = lambda lo, hi: rng.uniform(lo, hi, (J, I))
    t_surf = f(260, 305)
    d = dict(t_atm=t_surf + f(-6, 4), q_atm=f(1e-4, 1.5e-2), u_atm=f(-15, 15), v_atm=f(-10, 10), p_surf=f(9.6e4, 1.03e5),
             z_atm=f(15, 60), t_surf=t_surf, t_ca=t_surf + f(-1, 1), u_surf=np.zeros((J, I)), v_surf=np.zeros((J, I)),
             rough_mom=f(1e-4, 0.1), rough_heat=f(1e-4, 0.1), rough_moist=f(1e-4, 0.1), gust=np.ones((J, I)))
    d["p_atm"] = d["p_surf"] * f(0.985, 0.998)
    d["rough_scale"] = d["rough_mom"].copy()
    land = rng.uniform(size=(J, I)) < 0.3
    q_surf = f(1e-3, 2e-2)
    return d, land, q_surf


def test_surface_flux_consistency():
    d, land, q_surf = surface_case()
    s, mo = P.SatVaporPres(), P.MOConfig()
    o = P.surface_flux(s, mo, P.SurfaceFluxConfig(), q_atm_in=d["q_atm"], q_surf=q_surf, land=land,
                       **{k: v for k, v in d.items() if k != "q_atm"})
    # bulk formulae close: flux = rho * C * |V| * difference; derivatives are the analytic ones
    tv = d["t_atm"] * (1 + P.D608 * d["q_atm"])
    rho = d["p_atm"] / (P.RDGAS * tv)
    th = d["t_atm"] * (d["p_surf"] / d["p_atm"]) ** P.KAPPA
    t0 = np.where(land, d["t_ca"], d["t_surf"])
    assert np.allclose(o["flux_t"], P.CP_AIR * rho * o["cd_t"] * o["w_atm"] * (t0 - th), rtol=1e-13)
    assert np.allclose(o["flux_u"], -rho * o["cd_m"] * o["w_atm"] * d["u_atm"], rtol=1e-13)
    assert np.allclose(o["dhdt_surf"], o["flux_t"] / (t0 - th), rtol=1e-10)
    assert np.all(o["cd_m"] > 0) and np.all(o["cd_m"] < 0.1) and np.all(o["u_star"] > 0)
    assert np.all((o["ex_del_h"] > 0) & (o["ex_del_h"] < 1)) and np.all((o["ex_del_m"] > 0) & (o["ex_del_m"] < 1))
    assert np.allclose(o["flux_r"], P.STEFAN * d["t_surf"] ** 4) and np.allclose(o["drdt_surf"] * d["t_surf"], 4 * o["flux_r"])
    # stable (surface colder) columns have smaller exchange coefficients than unstable ones with the same geometry
    d2 = {k: np.full_like(v, v.flat[0]) for k, v in d.items()}
    d2["t_atm"] = d2["t_surf"] + np.linspace(-5, 5, d2["t_surf"].size).reshape(d2["t_surf"].shape)
    o2 = P.surface_flux(s, mo, P.SurfaceFluxConfig(), q_atm_in=d2["q_atm"], q_surf=q_surf, land=np.zeros_like(land),
                        **{k: v for k, v in d2.items() if k != "q_atm"})
    assert np.all(np.diff(o2["cd_t"].ravel()) <= 1e-15)


def turb_case(K=20, J=5, I=7, seed=4):
    rng = np.random.default_rng(seed)
    ps = 1.0e5 + 1.0e3 * rng.standard_normal((J, I))
    sig_h = np.linspace(0.0, 1.0, K + 1) ** 2.5                      # thin layers near the surface
    ph = sig_h[:, None, None] * ps[None]
    pf = 0.5 * (ph[1:] + ph[:-1])
    theta = 290.0 + rng.uniform(-1.0, 6.0, (J, I))[None] * (1.0 - pf / ps) * 20 + 0.3 * rng.standard_normal((K, J, I))
    t = theta * (pf / 1.0e5) ** P.KAPPA
    q = 8e-3 * (pf / 1e5) ** 3
    dlnp = np.log(ph[1:] / np.maximum(ph[:-1], 0.3 * ph[1]))
    zh = np.concatenate([np.cumsum((P.RDGAS * t * dlnp / P.GRAV)[::-1], 0)[::-1], np.zeros((1, J, I))]) + 50.0 * rng.uniform(size=(J, I))
    zf = 0.5 * (zh[1:] + zh[:-1])
    u, v = 8 + 3 * rng.standard_normal((K, J, I)), 2 * rng.standard_normal((K, J, I))
    u_star = rng.uniform(0.05, 0.6, (J, I))
    b_star = rng.uniform(-0.01, 0.02, (J, I))
    return t, q, u, v, pf, ph, zf, zh, u_star, b_star


def test_diffusivity_profile_properties():
    t, q, u, v, pf, ph, zf, zh, us, bs = turb_case()
    K = t.shape[0]
    mo = P.MOConfig()
    for c in (P.DiffusivityConfig(), P.DiffusivityConfig(do_simple=True, do_entrain=False), P.DiffusivityConfig(fixed_depth=True, depth_0=1500.0)):
        h, km, kt = P.diffusivity(c, mo, t, q, u, v, pf, ph, zf, zh, us, bs, np.zeros_like(t), np.zeros_like(t))
        zag = zh - zh[K]
        assert np.all(h >= zf[K - 1] - zh[K] - 1e-9)                         # at least the lowest full level
        assert np.all(km >= 0) and np.all(kt >= 0) and np.all(km[0] == 0)
        above = zag[:K] >= h[None]
        if not (c.do_entrain and not c.fixed_depth):
            assert np.all(km[above] == 0) and np.all(kt[above] == 0)          # use_pog_bug_fix: nothing above the PBL top
        # surface layer: K = k u* z / phi(z/L)
        inner = (zag[:K] < c.frac_inner * h[None]) & (np.arange(K)[:, None, None] > 0)
        kref, _ = P.mo_diff(mo, zag[:K], us, bs)
        assert np.allclose(km[inner], kref[inner], rtol=1e-14)
        if c.fixed_depth:
            assert np.all(h == 1500.0)
    # the input diffusivities are added; the floors apply afterwards
    c = P.DiffusivityConfig(do_entrain=False, background_m=0.5, background_t=0.25)
    h0, km0, kt0 = P.diffusivity(P.DiffusivityConfig(do_entrain=False), mo, t, q, u, v, pf, ph, zf, zh, us, bs, np.ones_like(t), 2 * np.ones_like(t))
    h1, km1, kt1 = P.diffusivity(P.DiffusivityConfig(do_entrain=False), mo, t, q, u, v, pf, ph, zf, zh, us, bs, np.zeros_like(t), np.zeros_like(t))
    assert np.allclose(km0, km1 + 1) and np.allclose(kt0, kt1 + 2)
    _, km2, kt2 = P.diffusivity(c, mo, t, q, u, v, pf, ph, zf, zh, us, bs, np.zeros_like(t), np.zeros_like(t))
    assert np.array_equal(km2, np.maximum(km1, 0.5)) and np.array_equal(kt2, np.maximum(kt1, 0.25))


def test_diffusivity_free_atmosphere_properties():
    """diffusivity_nml free_atm_diff (diffusivity_free, diffusivity.F90:604-697): above the boundary layer the diffusivity is the
    mixing-length value mix_len^2 |dU/dz| (1 - Ri/Ri_c)^2 where Ri < Ri_c, and untouched elsewhere"""
    t, q, u, v, pf, ph, zf, zh, us, bs = turb_case()
    K = t.shape[0]
    mo = P.MOConfig()
    z0 = np.zeros_like(t)
    base = P.DiffusivityConfig(do_entrain=False)
    h0, km0, kt0 = P.diffusivity(base, mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    c = P.DiffusivityConfig(do_entrain=False, free_atm_diff=True, rich_crit_diff=0.5, mix_len=40.0, rich_prandtl=0.8)
    h1, km1, kt1 = P.diffusivity(c, mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    assert np.array_equal(h0, h1)
    zag_h, zag_f = zh - zh[K], zf - zh[K]
    gcp = P.GRAV / P.CP_AIR
    svcp = t * (1 + P.D608 * q) + gcp * zag_f
    changed = (km1 != km0) | (kt1 != kt0)
    assert changed.any() and not changed[0].any()
    for k in range(1, K):
        dz = zag_f[k - 1] - zag_f[k]
        sp2 = (u[k - 1] - u[k]) ** 2 + (v[k - 1] - v[k]) ** 2
        ri = np.maximum(P.GRAV * (svcp[k - 1] - svcp[k]) / svcp[k] * dz / (sp2 + 1e-4), 0.0)
        on = (ri < 0.5) & (zag_h[k] > h1)
        want = 40.0 * 40.0 * np.sqrt(sp2) * (1 - ri / 0.5) ** 2 / dz
        assert np.allclose(kt1[k][on], want[on], rtol=1e-14) and np.allclose(km1[k][on], 0.8 * want[on], rtol=1e-14)
        assert np.array_equal(km1[k][~on], km0[k][~on]) and np.array_equal(kt1[k][~on], kt0[k][~on])
    # skyhi form: k_t / k_m = 0.1 + 0.9 (1 - Ri/Ri_c)^2 with the Richardson number before the ampns factor; ampns scales Ri down
    cs = P.DiffusivityConfig(do_entrain=False, free_atm_diff=True, free_atm_skyhi_diff=True, rich_crit_diff=0.5)
    _, kms, kts = P.diffusivity(cs, mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    ca = P.DiffusivityConfig(do_entrain=False, free_atm_diff=True, free_atm_skyhi_diff=True, rich_crit_diff=0.5, ampns=True, ampns_max=3.0)
    _, kma, kta = P.diffusivity(ca, mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    on = kms != km0
    assert on.any() and np.all(kts[on] <= kms[on] * (1 + 1e-14)) and np.all(kts[on] >= 0.1 * kms[on] * (1 - 1e-14))
    assert np.count_nonzero(kma != km0) >= np.count_nonzero(on)          # a smaller Ri is sub-critical at least as often
    # the boundary-layer entrainment value is written after the free-atmosphere values (diffusivity.F90:347-350)
    ce = P.DiffusivityConfig(free_atm_diff=True)
    _, kme, kte = P.diffusivity(ce, mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    _, kmb, ktb = P.diffusivity(P.DiffusivityConfig(), mo, t, q, u, v, pf, ph, zf, zh, us, bs, z0, z0)
    for k in range(1, K):
        ent = (bs > 0.0) & (zag_f[k - 1] > h1) & (zag_f[k] <= h1)
        assert np.array_equal(kte[k][ent], ktb[k][ent])


def conv_case(K, J, I, seed):
    svp = P.SatVaporPres()
    rng = np.random.default_rng(seed)
    ps = 1e5 + 1e3 * rng.standard_normal((J, I))
    sig = np.linspace(0, 1, K + 1) ** 1.2
    ph = sig[:, None, None] * ps
    pf = 0.5 * (ph[1:] + ph[:-1]); pf[0] = ph[1] / np.e
    Ts = rng.uniform(270, 305, (J, I))
    lapse = (0.17 + 0.09 * rng.uniform(size=(J, I)))[None]
    t = np.maximum(Ts[None] * (pf / ps) ** lapse, 200 + 5 * rng.uniform(size=(K, J, I)))
    qs, _ = svp.compute_qs(t, pf)
    rh = rng.uniform(0.1, 1.05, (J, I))[None] * rng.uniform(0.6, 1.0, (K, J, I))
    q = np.minimum(qs * rh, 0.05)
    q[:, 0, 0] = 0.0                                            # dry column: r0 <= 0 branch
    q[-1, 0, 1] = 1.2 * qs[-1, 0, 1]                            # saturated lowest level
    return svp, t, q, pf, ph


def test_sbm_convection_conservation_and_branches():
    svp, t, q, pf, ph = conv_case(25, 12, 20, 1)
    sbm = P.SBMConvection(svp, rhbm=0.7, Tmin=160.0, Tmax=350.0)
    assert sbm.lcl_temp_table.size == 1564 and abs(sbm.lcl_temp_table[0] - 160.0) < 1e-6
    # the table inverts value(T) = log(es(T) T^(-1/kappa))
    for T in (180.0, 250.0, 300.0, 340.0):
        v = np.log(sbm.es(T) * T ** (-1 / P.KAPPA))
        assert abs(sbm.get_lcl_temp(v) - T) < 2e-3
    dt = 600.0
    o = sbm(dt, t, q, pf, ph)
    dp = ph[1:] - ph[:-1]
    flags = np.bincount(o["convflag"].ravel(), minlength=3)
    assert flags[0] > 0 and flags[1] > 0 and flags[2] > 0
    assert np.all(o["snow"] == 0) and np.all(o["rain"] >= 0)
    assert np.array_equal(o["rain"] > 0, o["convflag"] == 2)
    # water: rain = -int dq dp/g ; enthalpy: int (cp dT + L dq) dp = 0 for every column
    assert np.abs(o["rain"] + (o["deltaq"] * dp).sum(0) / P.GRAV).max() < 1e-14
    ent = ((P.CP_AIR * o["deltaT"] + P.HLV * o["deltaq"]) * dp).sum(0) / P.GRAV
    assert np.abs(ent).max() < 1e-9 * np.abs(P.CP_AIR * o["deltaT"] * dp).sum(0).max() / P.GRAV + 1e-12
    # no convection: reference profiles are the model profiles, increments vanish
    none = o["convflag"] == 0
    assert np.all(o["deltaT"][:, none] == 0) and np.all(o["Tref"][:, none] == t[:, none]) and np.all(o["qref"][:, none] == q[:, none])
    assert o["convflag"][0, 0] == 0 and o["kLCLs"][0, 0] == 0           # dry column
    assert o["kLCLs"][0, 1] == 25                                        # saturated lowest level: LCL there
    # increments relax towards the reference profiles on tau_bm where deep convection changed only the T reference
    deep = o["convflag"] == 2
    k = np.arange(25)[:, None, None] + 1 >= o["kLZBs"][None]
    sel = deep[None] & k
    assert np.allclose(o["deltaT"][sel], -(t - o["Tref"])[sel] * dt / sbm.tau_bm, rtol=1e-10, atol=1e-13)
    # only the last column keeps its relaxation rates (array assignment inside the reference's column loop)
    assert np.count_nonzero(o["invtau_q_relaxation"].ravel()[:-1]) == 0


# ------------------------------------------------------------------------------------------------------------------
# two_stream_gray_rad_nml do_seasonal (two_stream_gray_rad.F90:417-447)
# ------------------------------------------------------------------------------------------------------------------
def _gauss_grid(J, I):
    x, w = np.polynomial.legendre.leggauss(J)
    lat = np.repeat(np.arcsin(x)[:, None], I, 1)
    lon = np.repeat((np.arange(I) * 2 * np.pi / I)[None], J, 0)
    return lat, lon, w


def test_seasonal_insolation_global_mean_is_a_quarter_of_the_solar_constant():
    """the sunlit disc: area mean of solar_constant*coszen = solar_constant/4 at any date and time of day (ecc = 0)"""
    from oracle import physics as P
    from oracle.rrtmg import Astronomy
    lat, lon, w = _gauss_grid(96, 256)
    c = P.GreyRadConfig()
    for days, seconds in ((0, 0.0), (93, 21600.0), (181, 80000.0), (275, 43200.0)):
        ins = P.seasonal_insolation(c, Astronomy(), days, seconds, lat, lon)
        mean = (ins.mean(1) * w).sum() / w.sum()
        assert abs(mean / c.solar_constant - 0.25) < 2e-4, (days, seconds)
        assert ins.min() == 0.0 and ins.max() <= c.solar_constant


def test_seasonal_insolation_equinox_solstice_and_perpetual_day():
    from oracle import physics as P
    from oracle.rrtmg import Astronomy
    lat, lon, w = _gauss_grid(32, 64)
    c, a = P.GreyRadConfig(solar_constant=1000.0), Astronomy()
    # equinox_day = 0.75 of a 360-day year: day 270 is the autumn equinox, declination 0, noon at lon = pi - gmt
    ins = P.seasonal_insolation(c, a, 270, 0.0, lat, lon)
    t = lon - np.pi
    expect = 1000.0 * np.maximum(np.cos(lat) * np.cos(t), 0.0)
    assert np.allclose(ins, expect, atol=1e-9)
    # a quarter year later: northern winter solstice, the daily mean vanishes poleward of 90 - obliq in the north
    daily = np.mean([P.seasonal_insolation(c, a, 0, s, lat, lon) for s in np.arange(0, 86400, 3600.0)], 0)
    north_polar = lat[:, 0] > np.deg2rad(90 - 23.439 + 1.0)
    assert north_polar.any() and np.all(daily[north_polar] == 0.0) and np.all(daily[lat[:, 0] < -np.deg2rad(70)] > 300.0)
    # solday >= 0: the date is frozen, the time of day still runs
    p1 = P.seasonal_insolation(c, a, 5, 1000.0, lat, lon, solday=90)
    p2 = P.seasonal_insolation(c, a, 200, 1000.0, lat, lon, solday=90)
    p3 = P.seasonal_insolation(c, a, 200, 40000.0, lat, lon, solday=90)
    assert np.array_equal(p1, p2) and not np.allclose(p2, p3)
    # use_time_average_coszen over a whole day = the daily mean insolation (h sin(lat) sin(dec) + cos(lat) cos(dec) sin h) / pi
    avg = P.seasonal_insolation(c, a, 45, 0.0, lat, lon, use_time_average_coszen=True, dt_rad_avg=86400.0)
    ang = a.angle(((45 * 86400.0 / (360 * 86400.0) - 0.75) % 1.0) * 2 * np.pi)
    dec = a.declination(ang)
    h = a.half_day(lat, dec)
    want = 1000.0 * (h * np.sin(lat) * np.sin(dec) + np.cos(lat) * np.cos(dec) * np.sin(h)) / np.pi
    assert np.allclose(avg, np.maximum(want, 0.0), rtol=1e-9, atol=1e-9)


def test_grey_radiation_with_insolation_override():
    """the do_seasonal insolation replaces the scheme's profile, schneider included (`if (do_seasonal) ... else if (B_SCHNEIDER_LIU)`);
    handing the analytic profile over reproduces the default result exactly"""
    from oracle import physics as P
    rng = np.random.default_rng(3)
    K, J, I = 12, 8, 16
    lat = np.repeat(np.linspace(-1.4, 1.4, J)[:, None], I, 1)
    ph = np.linspace(0, 1e5, K + 1)[:, None, None] * np.ones((1, J, I))
    t = 220 + 70 * (0.5 * (ph[1:] + ph[:-1]) / 1e5) ** 0.3 + rng.standard_normal((K, J, I))
    for scheme in ("frierson", "schneider"):
        g = P.GreyRadiation(P.GreyRadConfig(rad_scheme=scheme, atm_abs=0.2))
        d0 = g.down(lat, ph, t, albedo=np.zeros((J, I)))
        c = g.c
        if scheme == "schneider":
            prof = (c.solar_constant / np.pi) * np.cos(lat)
        else:
            prof = 0.25 * c.solar_constant * (1.0 + c.del_sol * (1.0 - 3.0 * np.sin(lat) ** 2) / 4.0 + c.del_sw * np.sin(lat))
        d1 = g.down(lat, ph, t, albedo=np.zeros((J, I)), insolation=prof)
        assert np.array_equal(d0["sw_down_surf"], d1["sw_down_surf"])
        d2 = g.down(lat, ph, t, albedo=np.zeros((J, I)), insolation=0.5 * prof)
        assert np.allclose(d2["sw_down_surf"], 0.5 * d0["sw_down_surf"], rtol=1e-14)
        assert np.array_equal(d2["surf_lw_down"], d0["surf_lw_down"])


# ------------------------------------------------------------------------------------------------------------------
# sat_vapor_pres tables: the product's host table builder (no GPU needed) against the oracle, do_simple and compute_es_k
# ------------------------------------------------------------------------------------------------------------------
def test_sat_vapor_pres_tables_of_the_library_match_the_oracle(lib_built):
    from isca_b200 import physics
    from oracle import physics as P
    for simple in (True, False):
        t, d, d2 = physics.sat_vapor_pres_tables(sat_vapor_pres_do_simple=int(simple))
        o = P.SatVaporPres(do_simple=simple)
        assert t.size == o.table_siz == physics.SVP_TABLE_SIZE
        assert np.abs(t / o.TABLE - 1).max() < 1e-13, simple
        assert np.abs(d / o.DTABLE - 1).max() < (1e-13 if simple else 1e-8), simple      # centred difference of nearly equal numbers
        assert np.abs(d2 - o.D2TABLE).max() < 1e-7 * np.abs(o.D2TABLE).max(), simple
    t2, _, _ = physics.sat_vapor_pres_tables(es0=1.5)
    assert np.allclose(t2, 1.5 * P.SatVaporPres().TABLE, rtol=1e-14)
    with pytest.raises(physics.IscaError):
        physics.sat_vapor_pres_tables(nonsense=1)


def test_compute_es_known_values():
    """compute_es_k: its anchor points by construction and the Smithsonian / Goff-Gratch table values it encodes"""
    from oracle import physics as P
    tf = P.TFREEZE
    assert abs(P.compute_es(tf + 100.) / 101324.60 - 1) < 1e-13              # ESBASW at the steam point
    assert abs(P.compute_es(tf - 1e-9) / 610.71 - 1) < 1e-3                  # ESBASI just below the ice point (blend weight ~0)
    assert abs(P.compute_es(tf - 20.) - P.compute_es(tf - 20. - 1e-9)) < 1e-6          # continuous at the ends of the blend
    assert abs(P.compute_es(tf) - P.compute_es(tf - 1e-9)) < 1e-3
    # Goff-Gratch values (WMO / Smithsonian tables): 20 C over water 2338 Pa, -10 C over ice 259.9 Pa, -40 C over ice 12.84 Pa
    assert abs(P.compute_es(293.15) / 2338.0 - 1) < 3e-3
    assert abs(P.compute_es(233.15) / 12.84 - 1) < 5e-3
    es = P.compute_es(np.linspace(150.0, 400.0, 2000))
    assert np.all(np.diff(es) > 0)
    # the simple Clausius-Clapeyron form agrees with it within a few per cent in the troposphere's warm range
    s, f = P.SatVaporPres(do_simple=True), P.SatVaporPres(do_simple=False)
    i0, i1 = int((273.16 - s.tminl) / s.dtres), int((303.16 - s.tminl) / s.dtres)
    assert np.abs(s.TABLE[i0:i1] / f.TABLE[i0:i1] - 1).max() < 0.03
    # lookup with the full tables: 2nd-order Taylor reproduces compute_es between the nodes
    T = np.array([215.37, 262.91, 288.04, 301.55])
    es_l, des_l = f.lookup_es_des(T)
    assert np.abs(es_l / P.compute_es(T) - 1).max() < 1e-6


def test_grey_radiation_do_read_co2_semantics():
    """do_read_co2: carbon_conc is replaced by the file value between the shortwave and the longwave part of
    two_stream_gray_rad_down (:466 vs :519-521): the geen shortwave lags by one call, the longwave does not"""
    from oracle import physics as P
    rng = np.random.default_rng(8)
    K, J, I = 10, 4, 8
    lat = np.repeat(np.linspace(-1.2, 1.2, J)[:, None], I, 1)
    ph = np.linspace(0, 1e5, K + 1)[:, None, None] * np.ones((1, J, I))
    pf = 0.5 * (ph[1:] + ph[:-1])
    t = 220 + 70 * (pf / 1e5) ** 0.3 + rng.standard_normal((K, J, I))
    q = 0.015 * (pf / 1e5) ** 3
    ref = {c: P.GreyRadiation(P.GreyRadConfig(rad_scheme="geen", carbon_conc=c)).down(lat, ph, t, q=q) for c in (360.0, 720.0)}
    g = P.GreyRadiation(P.GreyRadConfig(rad_scheme="geen", carbon_conc=360.0))
    d1 = g.down(lat, ph, t, q=q, carbon_conc=720.0)                      # first call with the new value: SW old, LW new
    assert np.array_equal(d1["sw_down_surf"], ref[360.0]["sw_down_surf"]) and np.array_equal(d1["surf_lw_down"], ref[720.0]["surf_lw_down"])
    d2 = g.down(lat, ph, t, q=q, carbon_conc=720.0)                      # second call: both new
    assert np.array_equal(d2["sw_down_surf"], ref[720.0]["sw_down_surf"]) and np.array_equal(d2["surf_lw_down"], ref[720.0]["surf_lw_down"])
    assert ref[720.0]["surf_lw_down"].mean() > ref[360.0]["surf_lw_down"].mean()       # more CO2, more back radiation
    b = P.GreyRadiation(P.GreyRadConfig(rad_scheme="byrne"))
    assert b.down(lat, ph, t, q=q, carbon_conc=1440.0)["surf_lw_down"].mean() > b.down(lat, ph, t, q=q, carbon_conc=360.0)["surf_lw_down"].mean()
