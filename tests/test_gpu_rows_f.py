"""GPU tests of the rows widened after the core path that use only kernels already verified on a B200 plus small
new ones: the moist test-case options (use_tau = .false., q-flux, sponge), the dry convection scheme, the land surface properties,
the barotropic model on the transform-level ABI.  Runs after the established suites and before the RRTMG tests."""
import numpy as np
import pytest

from rrtm_cases import rrtm_setup, unstable_boundary_layer

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def run_test_case_options(radiation):
    """what every shipped moist test case sets and the default does not: vert_turb_driver_nml use_tau = .false. (diffusivity from
    previous + delta_t * tendencies), constant_gust = 0, the MiMA roughness lengths, the Rayleigh sponge, a tropical ocean q-flux
    (qflux_mod) under the slab -- with grey and with RRTMG radiation, three steps against the oracle"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=9, damping=True)
    unstable_boundary_layer(core, mp)                       # non-zero diffusivities above the lowest level
    Kk, J, I = core.tg[0].shape
    mp.c.use_tau, mp.c.constant_gust = False, 0.0
    mp.c.roughness_mom = mp.c.roughness_heat = mp.c.roughness_moist = 3.21e-05
    qf = moist.qflux(moist.lat_boundaries(J), I)
    mp.ocean_qflux = qf.copy()
    phys = dict(FRIERSON_PHYS, trayfric=-0.5, sponge_pbottom=5000.0)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=phys, convection_scheme="SIMPLE_BETTS_MILLER",
                              mixed_layer_depth=2.5, albedo_value=0.31, do_damping=1, use_tau=0, constant_gust=0.0,
                              roughness_mom=3.21e-05, roughness_heat=3.21e-05, roughness_moist=3.21e-05)
    if radiation == "rrtm":
        rrtm_setup(core, mp, cfg, 1800, None)
        m.use_rrtm(dict(co2ppmv=360.0, solr_cnst=1360.0), dt_rad=1800)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    m.set_ocean_qflux(qf)
    for step in range(3):
        core.step(physics=True)
        m.atmosphere(1)
        if step == 0:
            assert mp.diag["diff_t"].max() > 0.0
            # with RRTMG the K-profile sees T + dt * tdt (use_tau = .false.) with heating rates that are differences of fluxes
            # (1e-10 relative between the two implementations): the Richardson-number dependence amplifies that
            tol_d = 1e-8 if radiation == "rrtm" else 1e-9
            assert rel(m.get("diff_t"), mp.diag["diff_t"]) < tol_d and rel(m.get("z_pbl"), mp.diag["z_pbl"]) < tol_d
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(atm.get_field(api.F_U), core.ug[core.current]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    m.atmosphere_end()


def test_moist_model_test_case_options_grey(lib_built):
    run_test_case_options("grey")


def test_dry_convection_parity(lib_built):
    """dry_convection (Schneider & Walker adjustment) through the C ABI: indices bit-exact, tendencies 1e-12"""
    from isca_b200 import api, physics
    from oracle import physics as PH
    from test_rrtm_host import dry_columns
    J, I, K = 8, 32, 25
    tg, pf, ph = dry_columns(J, I, K, 11)
    cp = physics.ColumnPhysics(I, J, K)
    o = cp.dry_convection(14400.0, 0.7, tg, pf, ph)
    dt, cape, cin, lzb, lcl = PH.dry_convection(tg, pf, ph, 14400.0, 0.7)
    assert np.array_equal(o["lzb"], lzb) and np.array_equal(o["lcl"], lcl)
    assert rel(o["dt_tg"], dt) < 1e-12 and rel(o["cape"], cape) < 1e-12
    with pytest.raises(api.IscaError):
        cp.dry_convection(0.0, 0.7, tg, pf, ph)          # dry_convection_nml has no defaults: tau must be given
    cp.close()


def test_moist_model_dry_convection_scheme(lib_built):
    """convection_scheme = 'DRY': dry adjustment, no large-scale condensation (idealized_moist_phys.F90:918-928, 977)"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    cfg, core, mp = build("T21", 25, 900.0, "NONE", seed=4)
    unstable_boundary_layer(core, mp, amp=14.0)             # dry-adiabatically unstable lower troposphere
    mp.c.convection_scheme, mp.c.dry_tau, mp.c.dry_gamma = "DRY", 7200.0, 0.7
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS), convection_scheme="DRY",
                              mixed_layer_depth=2.5, albedo_value=0.31)
    m.set_dry_convection(7200.0, 0.7)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    for step in range(3):
        core.step(physics=True)
        m.atmosphere(1)
        if step == 0:
            assert (mp.diag["lzb"] < 25).any() and rel(m.get("cape"), mp.diag["cape"]) < 1e-11
            assert np.abs(m.get("precip")).max() == 0.0
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    m.atmosphere_end()
    m2 = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS), convection_scheme="DRY")
    m2.core.cold_start()
    m2.idealized_moist_phys_init()
    with pytest.raises(api.IscaError):
        m2.atmosphere(1)                                     # tau / gamma not set
    m2.atmosphere_end()


def test_moist_model_with_land_surface_properties(lib_built):
    """the land options of idealized_moist_phys_init / mixed_layer_init as per-column fields: a land mask (surface_flux humidity and
    evaporation prefactors), land heat capacity, albedo and roughness; three steps against the oracle"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=12)
    unstable_boundary_layer(core, mp)
    Kk, J, I = core.tg[0].shape
    rng = np.random.default_rng(12)
    land = np.zeros((J, I), bool)
    land[J // 4: J // 2, I // 8: I // 2] = True
    land[3 * J // 5: 4 * J // 5, 5 * I // 8:] = True
    albedo = np.where(land, 0.31 * 1.3, 0.31)
    heat_cap = np.where(land, 0.1, 1.0) * mp.heat_capacity
    rough = np.where(land, 10.0, 1.0) * 3.21e-05
    mp.albedo, mp.heat_capacity, mp.land = albedo.copy(), heat_cap.copy(), land.copy()
    mp.rough_mom = mp.rough_heat = mp.rough_moist = rough.copy()
    mp.sflux.land_humidity_prefactor, mp.sflux.land_evap_prefactor = 0.7, 0.6
    phys = dict(FRIERSON_PHYS, land_humidity_prefactor=0.7, land_evap_prefactor=0.6)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=phys, convection_scheme="SIMPLE_BETTS_MILLER",
                              mixed_layer_depth=2.5, albedo_value=0.31)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    for name, f in (("albedo", albedo), ("heat_capacity", heat_cap), ("land", land.astype(float)), ("rough_mom", rough), ("rough_heat", rough),
                    ("rough_moist", rough)):
        m.set_surface(name, f)
    for step in range(3):
        core.step(physics=True)
        m.atmosphere(1)
        assert rel(m.get("flux_q"), mp.diag["flux_q"]) < TOL and rel(m.get("flux_t"), mp.diag["flux_t"]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
    # the land columns did behave differently: faster surface temperature response of the small heat capacity
    assert np.abs(m.get("delta_t_surf")[land]).mean() > 2.0 * np.abs(m.get("delta_t_surf")[~land]).mean()
    with pytest.raises(api.IscaError):
        m.set_surface("albedo", np.zeros((J, I + 1)))
    m.atmosphere_end()


def test_barotropic_model_on_the_gpu_transforms(lib_built):
    from isca_b200 import barotropic
    from oracle.barotropic import BarotropicConfig, BarotropicModel
    from test_barotropic import _compare, T21
    m = barotropic.BarotropicAtmosphere(1800.0, **T21)
    o = BarotropicModel(BarotropicConfig(dt_atmos=1800.0, **T21))
    _compare(m, o, 1e-11)
    for step in range(24):
        m.atmosphere(1)
        o.step()
        _compare(m, o, 1e-9)
    m.atmosphere_end()
    big = barotropic.BarotropicAtmosphere(1200.0)                  # the reference's default T85 (256 x 128)
    big.atmosphere(36)
    assert np.isfinite(big.energy) and 100.0 < big.energy < 2000.0 and np.abs(big.v).max() < 100.0
    big.atmosphere_end()



def test_shallow_water_model_on_the_gpu_transforms(lib_built):
    from isca_b200 import shallow
    from oracle.shallow import ShallowConfig, ShallowModel
    from test_barotropic import T21
    from test_shallow import compare
    m = shallow.ShallowAtmosphere(900.0, **T21)
    o = ShallowModel(ShallowConfig(dt_atmos=900.0, **T21))
    compare(m, o, 1e-11)
    for step in range(24):
        m.atmosphere(1)
        o.step()
        compare(m, o, 1e-8)
    m.atmosphere_end()
    big = shallow.ShallowAtmosphere(600.0)                         # the reference's default T85 (256 x 128)
    big.atmosphere(48)
    ens, div2, fr = big.global_diag()
    assert np.isfinite(ens) and ens > 0 and fr < 1.0
    big.atmosphere_end()


@pytest.mark.parametrize("scheme", ["frierson", "geen", "schneider"])
def test_two_stream_gray_rad_seasonal_insolation(lib_built, scheme):
    """do_seasonal (two_stream_gray_rad.F90:417-447): insolation = solar_constant * coszen handed to the down / up kernels"""
    from isca_b200 import physics
    from oracle import physics as O
    from oracle.rrtmg import Astronomy
    from test_gpu_physics import columns, TOL
    K, J, I = 20, 16, 32
    rng, ps, ph, pf, t, lat = columns(K, J, I, 31)
    lon = np.repeat((np.arange(I) * 2 * np.pi / I)[None], J, 0)
    q = 0.02 * (pf / ps[None]) ** 3 * rng.uniform(0.2, 1.0, t.shape)
    alb = rng.uniform(0.1, 0.4, (J, I))
    ts = t[-1] + rng.uniform(-3, 3, (J, I))
    tdt0 = 1e-5 * rng.standard_normal(t.shape)
    cp = physics.ColumnPhysics(I, J, K, rad_scheme=scheme, atm_abs=0.2)
    g = O.GreyRadiation(O.GreyRadConfig(rad_scheme=scheme, atm_abs=0.2))
    ins = O.seasonal_insolation(g.c, Astronomy(), 100, 30000.0, lat, lon, use_time_average_coszen=True, dt_rad_avg=3600.0)
    assert (ins == 0).any() and ins.max() > 800.0                       # a night side and a sub-solar region
    cp.two_stream_gray_rad_set_insolation(ins)
    d = g.down(lat, ph, t, q=q, albedo=alb, insolation=ins)
    sw, lw = cp.two_stream_gray_rad_down(lat, ph, t, alb, q=q)
    assert rel(lw, d["surf_lw_down"]) < TOL and rel(sw, (1 - alb) * d["sw_down_surf"]) < TOL
    tdt, olr = cp.two_stream_gray_rad_up(lat, ph, t, ts, alb, tdt0, q=q)
    to, o = g.up(ts, alb, ph, tdt0)
    assert rel(olr, o["olr"]) < TOL and rel(tdt - tdt0, to - tdt0) < 1e-11
    cp.two_stream_gray_rad_set_insolation(None)                        # back to the analytic profile
    d0 = g.down(lat, ph, t, q=q, albedo=alb)
    sw0, _ = cp.two_stream_gray_rad_down(lat, ph, t, alb, q=q)
    assert rel(sw0, (1 - alb) * d0["sw_down_surf"]) < TOL and not np.allclose(sw0, sw)


def test_moist_model_seasonal_grey_radiation(lib_built):
    """the moist model with two_stream_gray_rad_nml do_seasonal = .true.: diurnal + seasonal insolation from the model clock"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    from oracle.rrtmg import Astronomy
    cfg, core, mp = build("T21", 12, 900.0, "SIMPLE_BETTS_MILLER", seed=4)
    Kk, J, I = core.tg[0].shape
    lon = np.repeat((np.arange(I) * 2 * np.pi / I)[None], J, 0)
    nml = dict(solday=-10, equinox_day=0.75, use_time_average_coszen=True, dt_rad_avg=-1)
    mp.seasonal = dict(nml, astro=Astronomy(), lon=lon, day_in_s=86400.0, year_in_s=360 * 86400.0)
    mp.time_s = 47 * 86400.0 + 20000.0
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS), convection_scheme="SIMPLE_BETTS_MILLER",
                              mixed_layer_depth=2.5, albedo_value=0.31)
    m.set_seasonal(**nml)
    with pytest.raises(api.IscaError):
        m.use_rrtm(dict(co2ppmv=360.0))                                # do_seasonal belongs to the grey scheme
    m.set_time(47, 20000)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    for step in range(3):
        core.step()
        m.atmosphere(1)
        cz = m.get("coszen")
        assert rel(cz * 1360.0, mp.diag["insolation"]) < 1e-12, step
        assert (cz == 0).any() and cz.max() > 0.8
        assert rel(m.get("net_surf_sw_down"), mp.diag["net_surf_sw_down"]) < TOL, step
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    m.atmosphere_end()


# ---- hs_forcing_mod beyond the Held-Suarez default (include/isca_b200_hs.h); the column arithmetic of these kernels is checked on the
# ---- CPU in tests/test_hs_host.py, here the launch geometry, the zenith-angle / spin-up kernels and the model driver
def _hs_cases():
    from test_hs_host import CASES
    return CASES


@pytest.mark.parametrize("idx", range(9))
def test_hs_forcing_options_parity(lib_built, idx):
    from isca_b200 import hs as HS
    from oracle import hs_forcing as H
    from oracle.rrtmg import Astronomy
    from test_hs_host import case, run_oracle
    nml = _hs_cases()[idx]
    cfg = H.HsConfig(**nml)
    g = case(K=20, J=32, I=64, seed=idx)
    o = H.HsForcing(cfg, g["lat"], 3, 100, astronomy=Astronomy(ecc=cfg.ecc, obliq=cfg.obliq))
    f = HS.HsForcing(64, 32, 20, lat=g["lat"], time=(3, 100), **nml)
    if cfg.equilibrium_t_option == "top_down":
        assert rel(f.tg_prev, o.tg_prev) < 1e-13                       # the spin-up kernel
    ts = 86400 * 12 + 4321
    _, (udt, vdt, tdt, rdt) = run_oracle(cfg, g, 1800.0, ts, hs=o, ntr=2)
    z = np.zeros_like(g["t"])
    gu, gv, gt, gr, d = f.hs_forcing(1800.0, (12, 4321), g["lon"], g["lat"], g["p_half"], g["p_full"], g["u"], g["v"], g["t"], z + 1e-6, z - 1e-6,
                                     z + 1e-5, um=g["u"] * 0.9, vm=g["v"] * 1.1, rm=np.stack([g["r"]] * 2), rdt=np.stack([1e-9 * g["r"]] * 2),
                                     zfull=g["zfull"])
    assert rel(gu, udt) < 1e-12 and rel(gv, vdt) < 1e-12 and rel(gt, tdt) < 1e-12
    assert rel(d["teq"], o.diag["teq"]) < 1e-12
    for n in range(2):
        assert rel(gr[n], rdt[n]) < 1e-12
    if cfg.equilibrium_t_option == "top_down":
        assert rel(d["h_trop"], o.diag["h_trop"]) < 1e-13 and rel(f.tg_prev, o.tg_prev) < 1e-13
        f.tg_prev = o.tg_prev + 1.0                                    # restart hand-over
        assert rel(f.tg_prev, o.tg_prev + 1.0) < 1e-15
    f.hs_forcing_end()


@pytest.mark.parametrize("nml,ntr", [(dict(equilibrium_t_option="EXOPLANET", obliq=30.0), 0),
                                     (dict(equilibrium_t_option="top_down", stratosphere_t_option="hs_like", spinup_time=40.0,
                                           orbital_period=360.0, local_heating_option="Isidoro", local_heating_srfamp=2.0), 1),
                                     (dict(), 1)])
def test_dry_model_with_general_hs_forcing(lib_built, nml, ntr):
    """atmosphere() of the dry model with the general forcing against the oracle core driven by oracle/hs_forcing.py, three steps
    from a developed state; with the default options the result must also equal the fused Held-Suarez path of isca_b200_step"""
    from isca_b200 import api, hs as HS
    from oracle import hs_forcing as H
    from oracle.isca_oracle import SpectralCore, held_suarez_config
    from oracle.rrtmg import Astronomy
    cfg = held_suarez_config("T21", 15, 1200.0, num_tracers=ntr)
    core = SpectralCore(cfg)
    core.cold_start()
    for _ in range(40):
        core.step()
    Kk, J, I = core.tg[0].shape
    lat = np.repeat(core.tb.rad_lat[:, None], I, 1)
    lon = np.repeat((np.arange(I) * 2 * np.pi / I)[None], J, 0)
    hc = H.HsConfig(kappa=cfg.kappa, rdgas=cfg.rdgas, grav=cfg.grav, **nml)
    t0 = 5 * 86400 + 600
    o = H.HsForcing(hc, lat, 5, 600, astronomy=Astronomy(ecc=hc.ecc, obliq=hc.obliq))
    default = not nml
    fused = None
    if default:                                                          # the fused path, same start
        fused = api.Atmosphere(api.config_from_namelist_object(cfg))
    core.hs = H.CoreHsForcing(o, core, lon, lat, time_s=t0)
    m = HS.HsAtmosphere(api.config_from_namelist_object(cfg), **nml)
    for a in ([m.core, fused] if fused else [m.core]):
        for slot in (0, 1):
            tr = (core.grid_tracers[slot, 0],) if ntr else ()
            a.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], *tr)
            a.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
        a.set_vor_div_grid(core.vorg, core.divg)
        a.set_time_pointers(core.previous, core.current)
    m.set_time(5, 600)
    m.hs_forcing_init()
    if hc.equilibrium_t_option == "top_down":
        assert rel(m.get("tg_prev"), o.tg_prev) < 1e-13
    for step in range(3):
        core.step()
        m.atmosphere(1)
        assert rel(m.get("teq"), o.diag["teq"]) < 1e-11, step
        for name, fid in (("ug", api.F_U), ("vg", api.F_V), ("tg", api.F_T)):
            assert rel(m.core.get_field(fid), getattr(core, name)[core.current]) < 1e-10, (step, name)
        assert rel(m.core.get_field(api.F_PS), core.psg[core.current]) < 1e-10, step
        if ntr:
            assert rel(m.core.get_field(api.F_TRACER0), core.grid_tracers[core.current, 0]) < 1e-10, step
        if fused:
            fused.atmosphere(1)
            assert rel(m.core.get_field(api.F_T), fused.get_field(api.F_T)) < 1e-12, step
            assert rel(m.core.get_field(api.F_U), fused.get_field(api.F_U)) < 1e-12, step
    if hc.equilibrium_t_option == "top_down":
        assert rel(m.get("tg_prev"), o.tg_prev) < 1e-13 and rel(m.get("h_trop"), o.diag["h_trop"]) < 1e-13
    m.atmosphere_end()
    if fused:
        fused.atmosphere_end()


# ---- betts_miller_mod, the full Betts-Miller scheme (physics_bm.cu); column code checked on the CPU in tests/test_betts_miller.py
@pytest.mark.parametrize("idx", range(6))
def test_betts_miller_parity(lib_built, idx):
    from isca_b200 import physics
    from oracle.betts_miller import BettsMiller, BettsMillerConfig
    from test_betts_miller import OPTION_SETS, columns
    nml = OPTION_SETS[idx]
    svp, t, q, pf, ph = columns(768, K=24, seed=20 + idx)
    K, J, I = 24, 12, 64
    r = lambda a: np.ascontiguousarray(a.reshape(a.shape[0], J, I))
    t, q, pf, ph = r(t), r(q), r(pf), r(ph)
    o = BettsMiller(svp, BettsMillerConfig(**nml))(1200.0, t, q, pf, ph)
    cp = physics.ColumnPhysics(I, J, K)
    cp.betts_miller_init(**nml)
    g = cp.betts_miller(1200.0, t, q, pf, ph)
    assert np.array_equal(g["convflag"], o["convflag"]) and np.array_equal(g["kLZBs"], o["kLZB"]) and np.array_equal(g["kLCLs"], o["kLCL"])
    # do_shallower (betts_miller.f90:307-366): when the shallow-convection search removes every layer, what is left of the precipitation
    # integral is round-off (+-1e-20) and the reference branches on its SIGN: either the lowest layer keeps its reference profile with a
    # tendency scaled by ~1e-17, or it is reset to the environment with a zero tendency.  Both give the same tendencies to 1e-17; only the
    # diagnostic q_ref / t_ref of that one layer differ, and fused multiply-adds decide differently than the host arithmetic.  Those layers
    # (shallow columns whose lowest-layer tendency is round-off) must carry one of the two outcomes: the oracle's value, or the environment
    # value where the oracle kept the reference profile / the reference profile where the oracle reset to the environment.
    sc = max(np.abs(o["deltaq"]).max(), 1e-300)
    degenerate = (o["convflag"] == 1) & (np.abs(o["deltaq"][-1]) < 1e-13 * sc) & (np.abs(g["deltaq"][-1]) < 1e-13 * sc)
    if not nml.get("do_shallower", False):
        degenerate[:] = False
    for a, env in (("qref", q), ("Tref", t)):
        gd, od, ed = g[a][-1][degenerate], o[a][-1][degenerate], env[-1][degenerate]
        assert np.all((np.abs(gd - od) <= 1e-11 * np.abs(od)) | (gd == ed) | (od == ed)), a
    for a, b in (("rain", "rain"), ("deltaT", "deltaT"), ("deltaq", "deltaq"), ("qref", "qref"), ("Tref", "Tref"), ("CAPE", "CAPE"), ("CIN", "CIN"),
                 ("invtau_t_relaxation", "invtau_t"), ("invtau_q_relaxation", "invtau_q")):
        ga, ob = g[a], o[b]
        if a in ("qref", "Tref"):
            ga, ob = ga.copy(), ob.copy()
            ga[-1][degenerate] = ob[-1][degenerate]
        assert rel(ga, ob) < 1e-11, a
    assert np.all(g["snow"] == 0) and np.all(g["capeflag"] == 0)
    assert np.bincount(o["convflag"].ravel(), minlength=3).min() > 5
    with pytest.raises(physics.IscaError):
        cp.betts_miller_init(do_taucape=True)                          # order-dependent in the reference: rejected
    with pytest.raises(physics.IscaError):
        bad = t.copy(); bad[-1, 0, 0] = 900.0
        cp.betts_miller(1200.0, bad, q, pf, ph)                        # saturation vapour pressure table overflow


@pytest.mark.parametrize("nml", [dict(), dict(do_simp=False, do_shallower=True)])
def test_moist_model_full_betts_miller(lib_built, nml):
    """convection_scheme = 'FULL_BETTS_MILLER' in the moist model, three steps against the oracle"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    from oracle.betts_miller import BettsMiller, BettsMillerConfig
    cfg, core, mp = build("T21", 12, 900.0, "FULL_BETTS_MILLER", seed=6)
    mp.bm = BettsMiller(mp.svp, BettsMillerConfig(rhbm=0.7, **nml))
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS), convection_scheme="FULL_BETTS_MILLER",
                              mixed_layer_depth=2.5, albedo_value=0.31)
    m.set_betts_miller(rhbm=0.7, **nml)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    for step in range(3):
        core.step()
        m.atmosphere(1)
        assert np.array_equal(m.get("convflag").astype(int), mp.diag["convflag"]), step
        assert rel(m.get("precip"), mp.diag["precip"]) < 1e-9 or np.abs(mp.diag["precip"]).max() < 1e-12, step
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(atm.get_field(api.F_TRACER0), core.grid_tracers[core.current, 0]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    flags = np.bincount(mp.diag["convflag"].ravel(), minlength=3)
    assert flags[1] + flags[2] > 0
    m.atmosphere_end()


def test_full_saturation_vapour_pressure_tables(lib_built):
    """sat_vapor_pres_nml do_simple = .false.: compute_es_k tables behind lookup_es_des / compute_qs / lscale_cond"""
    from isca_b200 import physics
    from oracle import physics as O
    from test_gpu_physics import columns, TOL
    K, J, I = 16, 8, 32
    rng, ps, ph, pf, t, lat = columns(K, J, I, 5)
    s = O.SatVaporPres(do_simple=False)
    cp = physics.ColumnPhysics(I, J, K, sat_vapor_pres_do_simple=0, do_evap=1)
    T = np.linspace(180.0, 330.0, 997)
    es, des = cp.lookup_es_des(T)
    eo, do_ = s.lookup_es_des(T)
    assert rel(es, eo) < 1e-13 and rel(des, do_) < 1e-8
    qs, _ = s.compute_qs(t, pf)
    q = qs * rng.uniform(0.3, 1.4, size=t.shape)
    rain, tdel, qdel = cp.lscale_cond(t, q, pf, ph)
    ro, to, qo = O.lscale_cond(s, t, q, pf, ph, hc=1.0, do_evap=True)
    assert rel(tdel, to) < 1e-9 and rel(qdel, qo) < 1e-9 and rel(rain, ro) < 1e-9
    assert np.array_equal(qdel == 0.0, qo == 0.0)
    simple = physics.ColumnPhysics(I, J, K, do_evap=1)
    r2, _, _ = simple.lscale_cond(t, q, pf, ph)
    assert not np.allclose(r2, rain)


@pytest.mark.parametrize("scheme", ["geen", "byrne"])
def test_two_stream_gray_rad_do_read_co2(lib_built, scheme):
    """do_read_co2: a CO2 value that changes from call to call (the variable_co2_concentration test case), incl. the one-call lag
    of the geen shortwave"""
    from isca_b200 import physics
    from oracle import physics as O
    from test_gpu_physics import columns, TOL
    K, J, I = 18, 8, 32
    rng, ps, ph, pf, t, lat = columns(K, J, I, 41)
    q = 0.02 * (pf / ps[None]) ** 3 * rng.uniform(0.2, 1.0, t.shape)
    alb = rng.uniform(0.1, 0.4, (J, I))
    ts = t[-1] + rng.uniform(-3, 3, (J, I))
    tdt0 = np.zeros_like(t)
    cp = physics.ColumnPhysics(I, J, K, rad_scheme=scheme, atm_abs=0.2)
    g = O.GreyRadiation(O.GreyRadConfig(rad_scheme=scheme, atm_abs=0.2))
    prev = None
    for co2 in (360.0, 500.0, 500.0, 800.0):
        cp.two_stream_gray_rad_set_co2(co2)
        d = g.down(lat, ph, t, q=q, albedo=alb, carbon_conc=co2)
        sw, lw = cp.two_stream_gray_rad_down(lat, ph, t, alb, q=q)
        assert rel(lw, d["surf_lw_down"]) < TOL and rel(sw, (1 - alb) * d["sw_down_surf"]) < TOL, co2
        tdt, olr = cp.two_stream_gray_rad_up(lat, ph, t, ts, alb, tdt0, q=q)
        to, o = g.up(ts, alb, ph, tdt0)
        assert rel(olr, o["olr"]) < TOL and rel(tdt, to) < 1e-11, co2
        if prev is not None and co2 != prev[0]:
            assert not np.allclose(lw, prev[1])
        prev = (co2, lw)
    with pytest.raises(physics.IscaError):
        cp.two_stream_gray_rad_set_co2(0.0)


def test_moist_model_variable_co2_test_case(lib_built):
    """the options of exp/test_cases/variable_co2_concentration: rad_scheme = 'byrne', do_seasonal, do_read_co2"""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    from isca_b200 import api, moist
    from oracle.rrtmg import Astronomy
    cfg, core, mp = build("T21", 12, 900.0, "SIMPLE_BETTS_MILLER", seed=11, rad_scheme="byrne")
    Kk, J, I = core.tg[0].shape
    lon = np.repeat((np.arange(I) * 2 * np.pi / I)[None], J, 0)
    mp.seasonal = dict(solday=-10, equinox_day=0.75, use_time_average_coszen=True, dt_rad_avg=86400.0, astro=Astronomy(), lon=lon,
                       day_in_s=86400.0, year_in_s=360 * 86400.0)
    mp.time_s = 200 * 86400.0
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS, rad_scheme="byrne"),
                              convection_scheme="SIMPLE_BETTS_MILLER", mixed_layer_depth=2.5, albedo_value=0.31)
    m.set_seasonal(solday=-10, equinox_day=0.75, use_time_average_coszen=True, dt_rad_avg=86400)
    m.set_time(200, 0)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    m.set_t_surf(mp.t_surf)
    for step, co2 in enumerate((360.0, 540.0, 720.0)):
        mp.co2 = co2
        m.set_co2(co2)
        core.step()
        m.atmosphere(1)
        assert rel(m.get("surf_lw_down"), mp.diag["surf_lw_down"]) < TOL, step
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    m.atmosphere_end()


def test_top_down_model_restart_round_trip(lib_built):
    """tg_prev of RESTART/hs_forcing.res together with the dynamical core's restart variables: a run continued from the dump equals
    the uninterrupted run (1e-13: the fresh core recomputes the gradient batch the resident one carries over)"""
    from isca_b200 import api, hs as HS
    from oracle.isca_oracle import held_suarez_config
    cfg = api.config_from_namelist_object(held_suarez_config("T21", 12, 1200.0))
    nml = dict(equilibrium_t_option="top_down", spinup_time=30.0, orbital_period=360.0, ml_depth=2.0)
    a = HS.HsAtmosphere(cfg, **nml)
    a.core.cold_start()
    a.hs_forcing_init()
    a.atmosphere(6)
    c = a.core
    prev, cur = c.get_time_pointers()
    dump = {slot: dict(ug=c.get_field(api.F_U, slot), vg=c.get_field(api.F_V, slot), tg=c.get_field(api.F_T, slot),
                       psg=c.get_field(api.F_PS, slot), vors=c.get_spectral(api.S_VOR, slot), divs=c.get_spectral(api.S_DIV, slot),
                       ts=c.get_spectral(api.S_T, slot), ln_ps=c.get_spectral(api.S_LNPS, slot)) for slot in (0, 1)}
    vorg, divg = c.get_field(api.F_VOR), c.get_field(api.F_DIV)
    tg = a.get("tg_prev")
    a.atmosphere(5)
    ref_t, ref_u, ref_tg = c.get_field(api.F_T), c.get_field(api.F_U), a.get("tg_prev")
    b = HS.HsAtmosphere(cfg, **nml)
    for slot in (0, 1):
        d = dump[slot]
        b.core.set_grid_state(slot, d["ug"], d["vg"], d["tg"], d["psg"])
        b.core.set_spectral_state(slot, d["vors"], d["divs"], d["ts"], d["ln_ps"])
    b.core.set_vor_div_grid(vorg, divg)
    b.core.set_time_pointers(prev, cur)
    b.set_time(0, 6 * 1200)
    b.set_tg_prev(tg)
    b.hs_forcing_init()
    assert np.array_equal(b.get("tg_prev"), tg)                        # no spin-up after a restart
    b.atmosphere(5)
    assert rel(b.core.get_field(api.F_T), ref_t) < 1e-13 and rel(b.core.get_field(api.F_U), ref_u) < 1e-13
    assert rel(b.get("tg_prev"), ref_tg) < 1e-14
    assert not np.array_equal(ref_tg, tg)
    a.atmosphere_end(); b.atmosphere_end()
