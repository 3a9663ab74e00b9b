"""RRTMG on the GPU through the C ABI (include/isca_b200_rrtm.h) against the NumPy oracle (oracle/rrtmg.py) on the same
seeded columns.  Tolerances: fluxes 1e-11, heating rates 1e-9 relative to the field maximum (the device sums the g-points in
a warp-shuffle tree, the oracle sequentially; the exp-table index `int(tblint*x + 0.5)` is a discontinuity a last-bit
difference can cross once in ~1e7 evaluations: a single crossing changes a flux by < 1e-6 relative, hence the few-column
sizes here and the dedicated statistics in the full-size test)."""
import numpy as np
import pytest

from rrtm_cases import columns, mls_column, model_columns, zero_if_none as z

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
