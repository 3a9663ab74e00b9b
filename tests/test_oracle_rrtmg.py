"""RRTMG oracle (oracle/rrtmg.py) without a GPU: the coefficient file against the reference's own data where the reference
tree is present, the g-point reduction's invariants, and physical properties of the restated LW / SW schemes.
Parity unpinned by the reference (no golden vectors, no Fortran compiler): these checks are what pins the oracle."""
import os
import sys
import numpy as np
import pytest

from oracle import rrtmg as R
from rrtm_cases import columns, mls_column, zero_if_none as z

REF = "/root/reference/src/atmos_param/rrtm_radiation"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_file_complete():
    T = R.tables()
    assert T["lw_chi_mls"].shape == (7, 59) and T["lw_totplnk"].shape == (181, 16)
    ng_lw = [T["lw%02d_forref" % b].shape[-1] for b in range(1, 17)]
    assert ng_lw == R.LW_NGC and sum(ng_lw) == 140
    ng_sw = [T["sw%02d_sfluxref" % b].shape[0] for b in range(16, 30)]
    assert ng_sw == R.SW_NGC and sum(ng_sw) == 112
    for name, a in T.items():
        assert np.isfinite(a).all(), name


def test_reduction_invariants():
    """cmbgbNN keeps the band totals: Planck fractions sum to 1 per band (and per species-ratio column); the reduced solar
    source functions sum to the solar constant RRTMG_SW was built with (1368.22 W/m2) when each band is taken at its
    reference ratio; weighted k tables stay within the range of the originals (convex combinations)."""
    T = R.tables()
    for b in range(1, 17):
        for key in ("fracrefa", "fracrefb"):
            name = "lw%02d_%s" % (b, key)
            if name in T:
                s = T[name].sum(axis=0)
                assert np.allclose(s, 1.0, atol=2e-4), (name, s)
    tot = 0.0
    for b in range(16, 30):
        f = T["sw%02d_sfluxref" % b]
        tot += f.sum() if f.ndim == 1 else f[:, f.shape[1] // 2].sum()
    assert abs(tot - 1368.22) / 1368.22 < 0.01


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_tables_reproduce_reference_data():
    """(i) the parser's ORIGINAL 16-g-point key-species tables equal the reference's own netCDF copy of the same data
    (rrtmg_lw.nc / rrtmg_sw.nc, GPointSet 1) and (ii) the committed file is what the tool produces today."""
    from scipy.io import netcdf_file
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_rrtmg_tables as M
    for fam, first, nb, kg, nc in (("lw", 1, 16, "/rrtmg_lw/gcm_model/src/rrtmg_lw_k_g.f90", "/rrtmg_lw/gcm_model/data/rrtmg_lw.nc"),
                                   ("sw", 16, 14, "/rrtmg_sw/gcm_model/src/rrtmg_sw_k_g.f90", "/rrtmg_sw/gcm_model/data/rrtmg_sw.nc")):
        f = netcdf_file(REF + nc, "r", mmap=False)
        KA = f.variables["KeySpeciesAbsorptionCoefficientsLowerAtmos"].data[0]
        KB = f.variables["KeySpeciesAbsorptionCoefficientsUpperAtmos"].data[0]
        subs = M.split_subroutines(REF + kg)
        for band in range(first, first + nb):
            decls, gaxis, sc = M.parse_module_decls(REF + "/rrtmg_%s/gcm_model/modules/rr%s_kg%02d.f90" % (fam, fam, band))
            a = M.parse_assignments(subs["%s_kgb%02d" % (fam, band)], decls, sc)
            for key, ref in (("kao", KA), ("kbo", KB)):
                if key not in a:
                    continue
                k = a[key]
                r = ref[band - first, :, :, :, 0].transpose(2, 1, 0) if k.ndim == 3 else ref[band - first].transpose(3, 2, 1, 0)
                assert np.abs(k - r).max() <= 1e-15 * np.abs(r).max(), (fam, band, key)
    tables = M.OrderedDict()
    tables.update(M.build_shared())
    tables.update(M.build_family("lw", REF + "/rrtmg_lw/gcm_model/src/rrtmg_lw_k_g.f90", REF + "/rrtmg_lw/gcm_model/modules/rrlw_kg%02d.f90", range(1, 17), M.LW_NGN))
    tables.update(M.build_family("sw", REF + "/rrtmg_sw/gcm_model/src/rrtmg_sw_k_g.f90", REF + "/rrtmg_sw/gcm_model/modules/rrsw_kg%02d.f90", range(16, 30), M.SW_NGN))
    T = R.tables()
    assert set(T) == set(tables)
    for k, v in tables.items():
        assert np.array_equal(np.asarray(v), T[k]), k


def test_lw_known_magnitudes():
    """clear-sky mid-latitude-summer column: OLR, surface downward flux and tropospheric cooling in the published RRTMG
    range (OLR 280-285, surface down 343-352 W/m2 for MLS / 355 ppmv CO2); surface emission = sigma T^4 to 0.01 %"""
    g = mls_column()
    u, d, hr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"])
    assert 278.0 < u[0, -1] < 287.0
    assert 340.0 < d[0, 0] < 355.0
    assert abs(u[0, 0] - 5.6704e-8 * 294.0 ** 4) / u[0, 0] < 2e-4
    assert d[0, -1] == 0.0
    trop = g["play"][0] > 250.0
    assert (-3.5 < hr[0, trop]).all() and (hr[0, trop] < -0.8).all()
    # heating rate is the flux divergence: hr = heatfac * d(fnet)/dp
    fnet = u - d
    hf = R.heatfac(287.04 / (2.0 / 7.0))
    assert np.allclose(hr, hf * (fnet[:, :-1] - fnet[:, 1:]) / (g["plev"][:, :-1] - g["plev"][:, 1:]), rtol=1e-13)


def test_lw_responses():
    """more CO2 -> less OLR; isothermal atmosphere at the surface temperature with a black surface -> net flux ~ 0 inside
    the atmosphere's opaque bands, OLR <= sigma T^4; warmer surface -> more upward flux at every level"""
    g = mls_column(nc=3)
    g["co2"] = np.array([180e-6, 355e-6, 1420e-6])[:, None] * np.ones_like(g["play"])
    u, d, hr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"])
    assert u[0, -1] > u[1, -1] > u[2, -1]
    assert 2.0 < u[1, -1] - u[2, -1] < 8.0            # two doublings of CO2: 2 x (2.5 - 3.5) W/m2 instantaneous at the TOA
    iso = mls_column()
    iso["tlay"][:] = 280.0
    iso["tlev"][:] = 280.0
    iso["tsfc"][:] = 280.0
    u, d, hr = R.rrtmg_lw(iso["play"], iso["plev"], iso["tlay"], iso["tlev"], iso["tsfc"], iso["h2o"], iso["o3"], iso["co2"])
    sb = 5.6704e-8 * 280.0 ** 4
    assert abs(u[0, 0] - sb) / sb < 2e-4 and np.abs(u[0] - sb).max() / sb < 4e-3     # isothermal: upward flux constant in height (to the k-distribution accuracy)
    g2 = mls_column()
    u0 = R.rrtmg_lw(g2["play"], g2["plev"], g2["tlay"], g2["tlev"], g2["tsfc"], g2["h2o"], g2["o3"], g2["co2"])[0]
    u1 = R.rrtmg_lw(g2["play"], g2["plev"], g2["tlay"], g2["tlev"], g2["tsfc"] + 5.0, g2["h2o"], g2["o3"], g2["co2"])[0]
    assert (u1 > u0).all()


def test_sw_known_magnitudes_and_conservation():
    g = mls_column(nc=4)
    cz = np.array([0.5, 1.0, 0.0, 0.2])
    alb = np.array([0.2, 0.2, 0.2, 0.0])
    su, sd, hr = R.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], albedo=alb, coszen=cz)
    # incoming flux at the top = S0 * mu0: the 112 reduced solar source terms sum to the solar constant
    assert np.allclose(sd[:, -1], 1368.22 * cz, rtol=2e-3)
    assert (su[2] == 0).all() and (sd[2] == 0).all() and (hr[2] == 0).all()          # night
    # surface reflection: up = albedo * down at the surface; black surface reflects nothing
    assert np.allclose(su[:, 0], alb * sd[:, 0], rtol=1e-12, atol=1e-12)
    # clear-sky atmospheric absorption of an MLS column at mu0 = 0.5: 18-24 % of the incoming flux
    absorbed = (sd[0, -1] - su[0, -1]) - (sd[0, 0] - su[0, 0])
    assert 0.17 < absorbed / sd[0, -1] < 0.25
    # fluxes are bounded and the heating is positive
    assert (sd <= sd[:, -1:] * (1 + 1e-12)).all() and (su >= 0).all() and (hr >= -1e-12).all()
    # heating = net flux convergence (all layers but the top one, which the reference zeroes)
    net = sd - su
    hf = R.heatfac(287.04 / (2.0 / 7.0))
    ref = (net[:, 1:] - net[:, :-1]) * hf / (g["plev"][:, :-1] - g["plev"][:, 1:])
    assert np.allclose(hr[:, :-1], ref[:, :-1], rtol=1e-13, atol=1e-13) and (hr[:, -1] == 0).all()
    # scaling the solar constant scales every flux
    su2, sd2, _ = R.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], albedo=alb, coszen=cz, scon=2 * 1368.22)
    assert np.allclose(su2, 2 * su, rtol=1e-13) and np.allclose(sd2, 2 * sd, rtol=1e-13)


def test_sw_two_stream_building_blocks():
    """reftra: a conservative (w = 1) layer neither absorbs (R + T = 1) nor does a transparent one reflect; vrtqdr: a
    single non-scattering layer over a surface gives down = direct beam and up = albedo * beam * diffuse transmission"""
    tau = np.array([1e-4, 0.05, 0.5, 3.0])
    ref, refd, tra, trad = R.sw_reftra(np.zeros(4), 0.6, tau, np.ones(4))
    assert np.allclose(ref + tra, 1.0, atol=1e-12) and np.allclose(refd + trad, 1.0, atol=1e-12)
    ref, refd, tra, trad = R.sw_reftra(np.zeros(4), 0.6, tau, np.full(4, 1e-9))
    assert np.abs(ref).max() < 1e-8 and np.allclose(tra, R._sw_exp(tau / 0.6), rtol=1e-6)


def test_interp_temp():
    """linear profile in z -> half-level temperatures on the same line (interior levels), reference's extrapolations"""
    K = 12
    zf = np.linspace(30000.0, 100.0, K)[None]
    zh = np.concatenate([[0.0], 0.5 * (zf[0, :-1] + zf[0, 1:]), [0.0]])[None]
    t = 300.0 - 6.5e-3 * zf
    th = R.interp_temp(zf, zh, t)
    assert np.allclose(th[0, 1:K], 300.0 - 6.5e-3 * zh[0, 1:K], rtol=1e-13)
    assert np.isclose(th[0, 0], 0.5 * (3 * t[0, 0] - t[0, 1])) and np.isclose(th[0, K], 300.0, rtol=1e-12)


def test_random_columns_are_finite_and_consistent():
    for sec in (False, True):
        g = columns(48, 40, 11 + sec, secondary=sec)
        u, d, hr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]),
                              z(g["n2o"]), z(g["o2"]), z(g["cfc11"]), z(g["cfc12"]), z(g["cfc22"]), z(g["ccl4"]))
        assert np.isfinite(u).all() and np.isfinite(d).all() and np.isfinite(hr).all()
        assert (u > 0).all() and (d >= 0).all() and (d[:, -1] == 0).all()


def test_astronomy_time_average_is_the_mean_of_instantaneous_values():
    a = R.Astronomy()
    lat = np.repeat(np.linspace(-np.pi / 2, np.pi / 2, 17)[:, None], 16, 1)
    lon = np.repeat(np.linspace(0, 2 * np.pi, 16, endpoint=False)[None, :], 17, 0)
    dt = 7200.0 / 86400.0 * 2 * np.pi
    avg = a.diurnal_solar(lat, lon, 2.0, 0.3, dt=dt)[0]
    fine = np.mean([a.diurnal_solar(lat, lon, 2.0 + s, 0.3)[0] for s in np.linspace(0, dt, 801)], axis=0)
    assert np.abs(avg - fine).max() < 1e-3
    # equinox, daily mean: cos(lat)/pi
    daily = a.diurnal_solar(lat, lon, 0.0, 0.0, dt=2 * np.pi)[0]
    assert np.allclose(daily, np.cos(lat) / np.pi, atol=1e-9)
    # perpetual-equinox insolation integrates to S0/4 over the sphere
    w = np.cos(lat[:, 0]); w /= w.sum()
    assert abs((daily[:, 0] * w).sum() - 0.25) < 2e-3
    assert a.angle(0.0) == 0.0 and abs(a.orb_angle[-1] - 2 * np.pi) < 1e-9     # circular orbit: the table closes after one year


def test_radiation_alarm_and_storage():
    """run_rrtmg's alarm: radiation at the first call and then every dt_rad; stored heating / fluxes are returned in between"""
    rng = np.random.default_rng(0)
    from rrtm_cases import model_columns
    I, J, K = 4, 2, 20
    m = model_columns(I, J, K, 4)
    lat = np.repeat(np.linspace(-1.0, 1.0, J)[:, None], I, 1)
    lon = np.repeat(np.linspace(0, 2 * np.pi, I, endpoint=False)[None, :], J, 0)
    r = R.RrtmRadiation(lat, lon, 600.0, dt_rad=1800, o3=m["o3"])
    calls = []
    for step in range(7):
        tdt, fsw, flw = r(step * 600.0, m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"],
                          np.zeros((K, J, I)))
        calls.append(r.n_rad_calls)
        assert np.array_equal(tdt, r.tdt_rad) and np.array_equal(fsw, r.sw_flux)
    assert calls == [1, 1, 1, 2, 2, 2, 3]


def test_moist_oracle_with_rrtm_radiation_runs():
    """the oracle's idealized_moist_phys dispatcher with do_rrtm_radiation (what tests/test_gpu_moist.py compares the GPU
    model with): two steps at T21 L25 stay finite, radiation cools the troposphere, the slab receives the RRTMG surface fluxes"""
    from test_gpu_moist import build
    from rrtm_cases import rrtm_setup
    cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=5)
    _, _, pf, _ = core.pg.compute_pressures_and_heights(core.tg[1], core.psg[1], core.surf_geopotential, None)
    o3 = np.where(pf < 1.0e4, 1.2e-5 * np.exp(-((np.log(pf) - np.log(1.0e3)) ** 2) / 2), 6e-8)
    rrtm_setup(core, mp, cfg, 1800, o3)
    mp.time_s = 3 * 86400.0 + 43200.0
    for _ in range(2):
        core.step(physics=True)
    assert mp.rrtm.n_rad_calls == 1
    assert np.isfinite(core.tg[core.current]).all() and np.isfinite(mp.t_surf).all()
    assert (mp.rrtm.coszen >= 0).all() and (mp.rrtm.coszen > 0).any() and (mp.rrtm.coszen == 0).any()
    assert mp.rrtm.tdt_rad[12:].mean() < 0 and 150.0 < mp.rrtm.olr.mean() < 330.0
    assert np.array_equal(mp.diag["surf_lw_down"], mp.rrtm.lw_flux)


def test_qflux_and_lat_boundaries():
    """python helpers of the MiMA test case: lat_boundaries_global (transforms.F90:314-323) and qflux_mod"""
    from isca_b200 import moist            # pure NumPy helpers; no library call
    J = 64
    latb = moist.lat_boundaries(J)
    x, w = np.polynomial.legendre.leggauss(J)
    assert latb[0] == -np.pi / 2 and latb[-1] == np.pi / 2 and (np.diff(latb) > 0).all()
    assert ((np.arcsin(x) > latb[:-1]) & (np.arcsin(x) < latb[1:])).all()          # each Gaussian latitude inside its box
    assert np.allclose(np.sin(latb[1:]) - np.sin(latb[:-1]), w, atol=1e-13)         # box area = Gaussian weight
    q = moist.qflux(latb, 4)
    assert q.shape == (J, 4) and np.allclose(q, q[::-1], atol=1e-12)               # symmetric about the equator
    assert q[J // 2, 0] < -25.0 and q[J // 8, 0] >= 0.0                             # heat taken up at the equator, released poleward
    assert abs((q[:, 0] * w).sum()) < 0.5                                           # a transport: no net heating (W/m2, area mean)


def test_moist_oracle_use_tau_false():
    """vert_turb_driver_nml use_tau = .false. in the oracle dispatcher, on a state with an unstable boundary layer"""
    from test_gpu_moist import build
    from rrtm_cases import unstable_boundary_layer
    out = []
    for use_tau in (True, False):
        cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=9, damping=True)
        unstable_boundary_layer(core, mp)
        mp.c.use_tau = use_tau
        core.step(physics=True)
        assert np.isfinite(core.tg[core.current]).all() and mp.diag["diff_t"].max() > 0.0
        out.append(mp.diag["diff_t"].copy())
    d = np.abs(out[0] - out[1]).max()
    assert 0.0 < d <= max(out[0].max(), out[1].max())               # a different, comparable diffusivity


def test_lonstep_interpolation():
    """run_rrtmg with lonstep > 1: computed longitudes keep their own columns, the others are the linear interpolation between
    the neighbouring computed ones, closed around the latitude circle"""
    from rrtm_cases import model_columns
    I, J, K = 8, 2, 20
    m = model_columns(I, J, K, 8)
    lat = np.zeros((J, I))
    args = (0.0, m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], np.zeros((K, J, I)))
    full = R.RrtmRadiation(lat, lat, 600.0, o3=m["o3"])
    full.zenith = lambda s: m["coszen"]
    full(*args)
    sub = R.RrtmRadiation(lat, lat, 600.0, o3=m["o3"], lonstep=4)
    sub.zenith = full.zenith
    sub(*args)
    assert np.allclose(sub.lw_flux[:, ::4], full.lw_flux[:, ::4], rtol=1e-14)
    assert np.allclose(sub.tdt_rad[:, :, ::4], full.tdt_rad[:, :, ::4], rtol=1e-12, atol=1e-20)
    assert np.allclose(sub.lw_flux[:, 1], 0.25 * full.lw_flux[:, 4] + 0.75 * full.lw_flux[:, 0], rtol=1e-14)
    assert np.allclose(sub.lw_flux[:, 7], 0.75 * full.lw_flux[:, 0] + 0.25 * full.lw_flux[:, 4], rtol=1e-14)     # wraps around


def test_moist_oracle_dry_convection_scheme():
    """the oracle dispatcher with convection_scheme = 'DRY' (what tests/test_gpu_rrtm.py compares the GPU model with)"""
    from test_gpu_moist import build
    from rrtm_cases import unstable_boundary_layer
    cfg, core, mp = build("T21", 25, 900.0, "NONE", seed=4)
    unstable_boundary_layer(core, mp, amp=14.0)
    mp.c.convection_scheme, mp.c.dry_tau, mp.c.dry_gamma = "DRY", 7200.0, 0.7
    t0 = core.tg[core.previous].copy()
    core.step(physics=True)
    assert (mp.diag["lzb"] < 25).any() and (mp.diag["cape"] > 0).any()
    assert np.abs(mp.diag["precip"]).max() == 0.0 and np.isfinite(core.tg[core.current]).all()


def test_moist_oracle_land_surface_properties():
    """oracle dispatcher with per-column land properties (what the GPU land test compares with)"""
    from test_gpu_moist import build
    from rrtm_cases import unstable_boundary_layer
    cfg, core, mp = build("T21", 25, 900.0, "SIMPLE_BETTS_MILLER", seed=12)
    unstable_boundary_layer(core, mp)
    Kk, J, I = core.tg[0].shape
    land = np.zeros((J, I), bool)
    land[J // 4: J // 2, I // 8: I // 2] = True
    land[3 * J // 5: 4 * J // 5, 5 * I // 8:] = True
    mp.albedo = np.where(land, 0.31 * 1.3, 0.31)
    mp.heat_capacity = np.where(land, 0.1, 1.0) * mp.heat_capacity
    mp.land = land
    mp.rough_mom = mp.rough_heat = mp.rough_moist = np.where(land, 10.0, 1.0) * 3.21e-05
    mp.sflux.land_humidity_prefactor, mp.sflux.land_evap_prefactor = 0.7, 0.6
    for _ in range(3):
        core.step(physics=True)
    d = mp.diag["delta_t_surf"]
    assert np.isfinite(core.tg[core.current]).all()
    assert np.abs(d[land]).mean() > 2.0 * np.abs(d[~land]).mean()
