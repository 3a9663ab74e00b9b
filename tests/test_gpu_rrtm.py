"""RRTMG on the GPU through the C ABI (include/isca_b200_rrtm.h) against the NumPy oracle (oracle/rrtmg.py) on the same
seeded columns.  Tolerances: fluxes 1e-11, heating rates 1e-9 relative to the field maximum (the device sums the g-points in
a warp-shuffle tree, the oracle sequentially; the exp-table index `int(tblint*x + 0.5)` is a discontinuity a last-bit
difference can cross once in ~1e7 evaluations: a single crossing changes a flux by < 1e-6 relative, hence the few-column
sizes here and the dedicated statistics in the full-size test)."""
import numpy as np
import pytest

from rrtm_cases import columns, mls_column, model_columns, rrtm_setup, unstable_boundary_layer, zero_if_none as z

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rr(lib_built):
    from isca_b200 import rrtm
    r = rrtm.Rrtm(num_lon=16, num_lat=8, num_levels=40)
    yield r
    r.close()


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("secondary,K,seed", [(False, 40, 1), (True, 40, 2), (True, 25, 3), (False, 60, 4)])
def test_rrtmg_lw_parity(rr, secondary, K, seed):
    from oracle import rrtmg as R
    g = columns(96, K, seed, secondary=secondary)
    u, d, hr = rr.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], g["ch4"], g["n2o"],
                           g["o2"], g["cfc11"], g["cfc12"], g["cfc22"], g["ccl4"])
    ou, od, ohr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]), z(g["n2o"]),
                             z(g["o2"]), z(g["cfc11"]), z(g["cfc12"]), z(g["cfc22"]), z(g["ccl4"]))
    assert rel(u, ou) < 1e-11 and rel(d, od) < 1e-11 and rel(hr, ohr) < 1e-9


@pytest.mark.parametrize("secondary,K,seed", [(False, 40, 5), (True, 40, 6), (True, 30, 7)])
def test_rrtmg_sw_parity(rr, secondary, K, seed):
    from oracle import rrtmg as R
    g = columns(96, K, seed, secondary=secondary)
    rng = np.random.default_rng(seed)
    alb = rng.uniform(0.0, 0.9, 96)
    cz = rng.uniform(-0.2, 1.0, 96)
    cz[:4] = [1e-11, 1e-9, 1.0, 0.01]
    su, sd, shr = rr.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], g["ch4"], g["n2o"], g["o2"], alb, cz, 1.03, 1360.0)
    ou, od, ohr = R.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]), z(g["n2o"]), z(g["o2"]), alb, cz,
                             1.03, 1360.0)
    night = cz < 1e-10
    assert (su[night] == 0).all() and (sd[night] == 0).all() and (shr[night] == 0).all()
    assert rel(su, ou) < 1e-11 and rel(sd, od) < 1e-11 and rel(shr, ohr) < 1e-9


def test_emissivity_and_known_magnitudes(rr):
    from oracle import rrtmg as R
    g = mls_column(nc=4)
    emis = np.tile(np.linspace(0.85, 1.0, 16), (4, 1))
    u, d, hr = rr.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], emis=emis)
    ou, od, ohr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], emis=emis)
    assert rel(u, ou) < 1e-11 and rel(hr, ohr) < 1e-9
    u, d, hr = rr.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"])
    assert 278.0 < u[0, -1] < 287.0 and 340.0 < d[0, 0] < 355.0        # published clear-sky MLS range


def test_run_rrtmg_parity(lib_built):
    """interp_temp + run_rrtmg on model-layout fields: heating added to tdt, surface and TOA fluxes"""
    from isca_b200 import rrtm
    from oracle import rrtmg as R
    I, J, K = 16, 8, 40
    m = model_columns(I, J, K, 21)
    r = rrtm.Rrtm(num_lon=I, num_lat=J, num_levels=K, co2ppmv=360.0, solr_cnst=1360.0)
    tdt0 = np.random.default_rng(3).normal(0, 1e-5, (K, J, I))
    tdt = tdt0.copy()
    out = r.run_rrtmg(m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], m["coszen"], tdt, o3=m["o3"])
    r.close()
    col = lambda a: a.reshape(a.shape[0], -1).T                      # [K][J][I] -> [ncol, K]
    th = R.interp_temp(col(m["z_full"]), col(m["z_half"]), col(m["t"]))
    o3v = col(m["o3"]) * (1000.0 * R.GAS_CONSTANT / R.RDGAS) / R.WTMOZONE
    o = R.run_rrtmg_columns(col(m["p_full"]), col(m["p_half"]), col(m["t"]), th, col(m["q"]), m["t_surf"].ravel(), m["albedo"].ravel(),
                            m["coszen"].ravel(), o3vmr=o3v, co2ppmv=360.0, solr_cnst=1360.0)
    back = lambda a: a.T.reshape(K, J, I)
    assert rel(out["tdt_rad"], back(o["tdt_rad"])) < 1e-9
    assert np.abs((tdt - tdt0) - out["tdt_rad"]).max() < 1e-18 + 1e-12 * np.abs(out["tdt_rad"]).max()
    for k in ("flux_sw", "flux_lw", "olr", "toa_sw"):
        assert rel(out[k].ravel(), o[k]) < 1e-11, k


def test_full_size_properties(lib_built):
    """T170 grid (131072 columns, 40 levels): size-independent properties -- every flux finite, surface emission = sigma T^4
    (black surface, to the Planck-table accuracy), heating = flux divergence, TOA incoming SW = S0 cos(zenith), night columns
    zero, two identical halves of the batch give bitwise identical results (deterministic reductions)."""
    from isca_b200 import rrtm
    g = columns(4096, 40, 99, secondary=False)
    rep = 16                                     # 65536 columns: half of T170 keeps the host arrays small; two halves compared
    big = {k: (None if v is None else np.tile(v, (rep,) + (1,) * (v.ndim - 1))) for k, v in g.items()}
    r = rrtm.Rrtm(num_lon=16, num_lat=8, num_levels=40)
    u, d, hr = r.rrtmg_lw(big["play"], big["plev"], big["tlay"], big["tlev"], big["tsfc"], big["h2o"], big["o3"], big["co2"])
    assert np.isfinite(u).all() and np.isfinite(d).all() and np.isfinite(hr).all()
    assert np.array_equal(u[:4096], u[4096 * (rep - 1):]) and np.array_equal(hr[:4096], hr[4096 * (rep - 1):])
    sb = 5.6704e-8 * big["tsfc"] ** 4
    assert np.abs(u[:, 0] - sb).max() / sb.max() < 1e-3
    hf = 9.8066 * 8.64e4 / (287.04 / (2.0 / 7.0) * 1.0e2)
    fnet = u - d
    assert np.allclose(hr, hf * (fnet[:, :-1] - fnet[:, 1:]) / (big["plev"][:, :-1] - big["plev"][:, 1:]), rtol=1e-12, atol=1e-12)
    cz = np.tile(np.linspace(-0.5, 1.0, 4096), rep)
    su, sd, shr = r.rrtmg_sw(big["play"], big["plev"], big["tlay"], big["h2o"], big["o3"], big["co2"], albedo=0.25, coszen=cz)
    day = cz >= 1e-10
    assert (sd[~day] == 0).all() and np.allclose(sd[day, -1], 1368.22 * cz[day], rtol=2e-3)
    assert np.allclose(su[:, 0], 0.25 * sd[:, 0], rtol=1e-12)
    ms_lw, ms_sw = r.time_kernel(0, 3), r.time_kernel(1, 3)
    assert ms_lw > 0 and ms_sw > 0
    r.close()


@pytest.mark.parametrize("dt", [None, 7200.0 / 86400.0 * 2 * np.pi, 2 * np.pi])
def test_diurnal_solar_parity(rr, dt):
    """astronomy_mod diurnal_solar (instantaneous and time-averaged zenith angle) at several times of day / year"""
    from oracle import rrtmg as R
    lat = np.repeat(np.linspace(-np.pi / 2, np.pi / 2, 33)[:, None], 64, 1)
    lon = np.repeat(np.linspace(0, 2 * np.pi, 64, endpoint=False)[None, :], 33, 0)
    for ecc in (0.0, 0.0167):
        a = R.Astronomy(ecc=ecc)
        for gmt, tsae in ((0.0, 0.0), (1.0, 0.3), (4.5, 2.0), (6.2, 5.5)):
            c, f, r = rr.diurnal_solar(lat, lon, gmt, tsae, dt, ecc=ecc)
            oc, of, orr = a.diurnal_solar(lat, lon, gmt, tsae, dt)
            assert np.abs(c - oc).max() < 1e-13 and np.abs(f - of).max() < 1e-12 and abs(r - orr) < 1e-14


@pytest.mark.parametrize("dt_rad_steps,kw", [(2, {}), (1, dict(frierson_solar_rad=True)), (2, dict(solday=90, do_rad_time_avg=False)),
                                             (0, {})])    # dt_rad = 0: the namelist default, dt_rad = dt_rad_avg = dt_atmos (rrtm_radiation.F90:378-405)
def test_moist_model_with_rrtm_radiation(lib_built, dt_rad_steps, kw):
    """do_rrtm_radiation = .true. (the MiMA configuration, idealized_moist_phys.F90:1167-1177): RRTMG on a radiation step, the
    stored heating rates / surface fluxes in between (dt_rad = 2 dt_atmos), diurnal-mean zenith angle from astronomy_mod; four
    steps of the whole model against the oracle."""
    from test_gpu_moist import build, FRIERSON_PHYS, TOL
    cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=5)
    Kk, J, I = core.tg[0].shape
    _, _, pf, _ = core.pg.compute_pressures_and_heights(core.tg[1], core.psg[1], core.surf_geopotential, None)
    o3 = np.where(pf < 1.0e4, 1.2e-5 * np.exp(-((np.log(pf) - np.log(1.0e3)) ** 2) / 2), 6e-8)      # mass mixing ratio
    dt_rad = int(dt_rad_steps * cfg.dt_atmos)
    rrtm_setup(core, mp, cfg, dt_rad, o3, **kw)
    from isca_b200 import api, moist
    phys = dict(FRIERSON_PHYS)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=phys, convection_scheme="SIMPLE_BETTS_MILLER",
                              mixed_layer_depth=2.5, albedo_value=0.31)
    m.use_rrtm(dict(co2ppmv=360.0, solr_cnst=1360.0), dt_rad=dt_rad, **{k: int(v) if isinstance(v, bool) else v for k, v in kw.items()})
    m.set_ozone(o3)
    atm = m.core
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot], core.grid_tracers[slot, 0])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    m.idealized_moist_phys_init()
    # start in the afternoon of day 3 so that the zenith angle has day and night columns
    m.set_time(3, 43200)
    mp.time_s = 3 * 86400.0 + 43200.0
    for step in range(4):
        core.step(physics=True)
        m.atmosphere(1)
        if step == 0:
            assert rel(m.get("coszen"), mp.rrtm.coszen) < 1e-12
            assert rel(m.get("tdt_rad"), mp.rrtm.tdt_rad) < 1e-9
            assert rel(m.get("net_surf_sw_down"), mp.rrtm.sw_flux) < 1e-11 and rel(m.get("surf_lw_down"), mp.rrtm.lw_flux) < 1e-11
            assert rel(m.get("olr"), mp.rrtm.olr) < 1e-11
        assert rel(atm.get_field(api.F_T), core.tg[core.current]) < TOL, step
        assert rel(atm.get_field(api.F_PS), core.psg[core.current]) < TOL, step
        assert rel(m.get("t_surf"), mp.t_surf) < TOL, step
    assert mp.rrtm.n_rad_calls == (4 if dt_rad_steps <= 1 else 2)
    m.atmosphere_end()


def test_rrtm_driver_loud_failures(lib_built):
    from isca_b200 import api, moist
    from test_gpu_moist import build, FRIERSON_PHYS
    cfg, core, mp = build("T21", 25, 900.0, "NONE", seed=1)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=dict(FRIERSON_PHYS), convection_scheme="NONE")
    with pytest.raises(api.IscaError) as e:
        m.use_rrtm(dt_rad=1000)                        # rrtm_gases_init: dt_rad must be an integer multiple of dt_atmos
    assert "integer multiple" in str(e.value)
    with pytest.raises(api.IscaError):
        m.set_ozone(np.zeros(m.s3))                    # before use_rrtm
    with pytest.raises(api.IscaError):
        m.use_rrtm(dict(lonstep=5))                    # does not divide 64 longitudes
    m.use_rrtm()
    with pytest.raises(api.IscaError):
        m.use_rrtm()                                   # twice
    m.atmosphere_end()



def test_mima_test_case_runs(lib_built):
    """MiMA_test_case.py at T21 L40: RRTMG every 7200 s, sponge, q-flux, prescribed initial SST; 30 steps stay physical"""
    from isca_b200 import moist
    m = moist.mima_test_case("T21", 40, 900.0)
    ts0 = m.get("t_surf")
    assert abs(ts0.max() - (285.0 + 40.0 / 3.0)) < 0.2 and ts0.min() > 285.0 - 2 * 40.0 / 3.0 - 1e-9
    m.atmosphere(30)
    olr, cz, tsw = m.get("olr"), m.get("coszen"), m.get("toa_sw")
    assert np.isfinite(olr).all() and 120.0 < olr.mean() < 330.0
    assert (cz >= 0).all() and cz.max() <= 1.0 and (tsw[cz == 0] == 0).all() and tsw.max() > 100.0
    from isca_b200 import api
    t = m.core.get_field(api.F_T)
    assert np.isfinite(t).all() and 150.0 < t.min() and t.max() < 330.0
    m.atmosphere_end()


def test_axisymmetric_test_case_runs(lib_built):
    """axisymmetric_test_case.py at T21 L40: MiMA physics + make_symmetric + free_atm_diff + prescribed, seasonally moving SSTs;
    the flow stays zonally symmetric (every field equals its zonal mean to rounding) and physical"""
    from isca_b200 import moist, api
    m = moist.axisymmetric_test_case("T21", 40, 900.0)
    lat = np.repeat(np.arcsin(m.core.get_table(api.TB_SIN_LAT))[:, None], 64, 1)
    for i in range(24):
        sst = 273.0 + 27.0 * np.cos(lat - 0.1 * np.sin(2 * np.pi * i / 480.0)) ** 2        # stand-in for sn_1.000_sst.nc
        m.set_sst(sst)
        m.atmosphere(1)
    assert np.array_equal(m.get("t_surf"), sst)
    for fid in (api.F_U, api.F_V, api.F_T, api.F_TRACER0):
        f = m.core.get_field(fid)
        assert np.isfinite(f).all()
        assert np.abs(f - f.mean(axis=-1, keepdims=True)).max() <= 1e-9 * max(np.abs(f).max(), 1e-30), fid
    t = m.core.get_field(api.F_T)
    assert 150.0 < t.min() and t.max() < 330.0 and np.isfinite(m.get("olr")).all()
    m.atmosphere_end()


@pytest.mark.parametrize("lonstep", [2, 4])
def test_run_rrtmg_lonstep(lib_built, lonstep):
    """rrtm_radiation_nml lonstep: radiation on every lonstep-th longitude, linear interpolation closed around the latitude circle"""
    from isca_b200 import rrtm
    from oracle import rrtmg as R
    I, J, K = 16, 4, 30
    m = model_columns(I, J, K, 33)
    r = rrtm.Rrtm(num_lon=I, num_lat=J, num_levels=K, lonstep=lonstep)
    tdt = np.zeros((K, J, I))
    out = r.run_rrtmg(m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], m["coszen"], tdt, o3=m["o3"])
    r.close()
    lat = np.zeros((J, I))
    o = R.RrtmRadiation(lat, lat, 600.0, o3=m["o3"], lonstep=lonstep)
    o.zenith = lambda total_seconds: m["coszen"]
    t2, fsw, flw = o(0.0, m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], np.zeros((K, J, I)))
    assert rel(out["tdt_rad"], o.tdt_rad) < 1e-9 and np.array_equal(tdt, out["tdt_rad"])
    assert rel(out["flux_sw"], fsw) < 1e-11 and rel(out["flux_lw"], flw) < 1e-11 and rel(out["olr"], o.olr) < 1e-11
    # the computed longitudes carry their own columns unchanged
    full = R.RrtmRadiation(lat, lat, 600.0, o3=m["o3"])
    full.zenith = o.zenith
    full(0.0, m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], np.zeros((K, J, I)))
    assert rel(out["flux_lw"][:, ::lonstep], full.lw_flux[:, ::lonstep]) < 1e-11


def test_moist_model_test_case_options_rrtm(lib_built):
    from test_gpu_rows_f import run_test_case_options
    run_test_case_options("rrtm")
