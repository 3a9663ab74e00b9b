"""Shared input generators of the RRTMG tests: seeded random columns spanning the tables' pressure / temperature /
composition ranges (surface pressures 600-1050 hPa, tops 0.01-0.04 hPa, CO2 180-5000 ppmv, with and without the secondary
gases), and a mid-latitude-summer-like column for the known-magnitude checks."""
import numpy as np


def columns(nc, K, seed, secondary=False, ptop=0.02):
    rng = np.random.default_rng(seed)
    ps = rng.uniform(600, 1050, nc)
    sig = np.linspace(0, 1, K + 1) ** rng.choice([1.5, 2.0, 2.5])
    ph = sig[None, ::-1] * ps[:, None]
    ph[:, -1] = ptop * rng.uniform(0.5, 2, nc)
    pl = 0.5 * (ph[:, :-1] + ph[:, 1:])
    ts = rng.uniform(230, 310, nc)
    z = 7.0 * np.log(ps[:, None] / pl)
    t = np.maximum(ts[:, None] - rng.uniform(4, 8, nc)[:, None] * z, rng.uniform(190, 225, nc)[:, None]) + rng.normal(0, 2, (nc, K))
    t = np.where(pl < 30, t + (30 - pl) / 30 * rng.uniform(10, 60, nc)[:, None], t)
    tl = np.empty((nc, K + 1))
    tl[:, 1:-1] = 0.5 * (t[:, :-1] + t[:, 1:])
    tl[:, 0] = t[:, 0] + rng.normal(0, 1, nc)
    tl[:, -1] = t[:, -1]
    h2o = np.maximum(rng.uniform(0.002, 0.03, nc)[:, None] * np.exp(-z / rng.uniform(1.5, 3, nc)[:, None]), 2e-7)
    o3 = np.where(pl < 100, rng.uniform(2e-6, 1e-5, nc)[:, None] * np.exp(-((np.log(pl) - np.log(10)) ** 2) / 2),
                  rng.uniform(1e-8, 6e-8, nc)[:, None])
    co2 = np.full((nc, K), 1e-6) * rng.choice([180., 300., 355., 1200., 5000.], nc)[:, None]
    g = dict(play=pl, plev=ph, tlay=t, tlev=tl, tsfc=ts + rng.normal(0, 2, nc), h2o=h2o, o3=o3, co2=co2)
    if secondary:
        g.update(ch4=np.full((nc, K), 1.7e-6) * rng.uniform(0, 3, nc)[:, None], n2o=np.full((nc, K), 3.2e-7) * rng.uniform(0, 8, nc)[:, None],
                 o2=np.full((nc, K), 0.209), cfc11=np.full((nc, K), 2.5e-10), cfc12=np.full((nc, K), 5e-10),
                 cfc22=np.full((nc, K), 1e-10), ccl4=np.full((nc, K), 1e-10))
    else:
        g.update(ch4=None, n2o=None, o2=None, cfc11=None, cfc12=None, cfc22=None, ccl4=None)
    return g


def mls_column(K=40, nc=1):
    """mid-latitude-summer-like clear column (294 K surface, 6.5 K/km, 1.4 % surface water vapour, ozone layer)"""
    ph = np.linspace(0, 1, K + 1) ** 2 * 1000.0
    ph[0] = 0.02
    ph = ph[::-1].copy()
    pl = 0.5 * (ph[:-1] + ph[1:])
    z = 7.0 * np.log(1013. / pl)
    t = np.maximum(294 - 6.5 * z, 216.0)
    t = np.where(pl < 50, 216 + (50 - pl) / 50 * 40, t)
    zl = 7.0 * np.log(1013. / np.maximum(ph, 0.02))
    tl = np.maximum(294 - 6.5 * zl, 216.0)
    q = np.maximum(0.014 * np.exp(-z / 2.0), 3e-6)
    h2o = q / (1 - q) * 28.9644 / 18.015
    o3 = np.where(pl < 100, 6e-6 * np.exp(-((np.log(pl) - np.log(10)) ** 2) / 2), 3e-8)
    rep = lambda a: np.tile(a, (nc, 1))
    return dict(play=rep(pl), plev=rep(ph), tlay=rep(t), tlev=rep(tl), tsfc=np.full(nc, 294.0), h2o=rep(h2o), o3=rep(o3),
                co2=np.full((nc, K), 355e-6))


def zero_if_none(a):
    return 0.0 if a is None else a


def model_columns(I, J, K, seed):
    """[K][J][I] model-layout fields (Pa, K, kg/kg, m) for run_rrtmg"""
    rng = np.random.default_rng(seed)
    bk = np.linspace(0, 1, K + 1) ** 2.2
    ps = rng.uniform(90000, 103000, (J, I))
    p_half = bk[:, None, None] * ps[None]
    p_full = np.empty((K, J, I))
    # Simmons-Burridge-like full pressures (any monotone choice serves the test)
    p_full[:] = 0.5 * (p_half[:-1] + p_half[1:])
    p_full[0] = 0.5 * p_half[1]
    ts = rng.uniform(250, 305, (J, I))
    z_full = 7000.0 * np.log(ps[None] / p_full)
    z_half = np.empty((K + 1, J, I))
    z_half[1:] = 7000.0 * np.log(ps[None] / p_half[1:])
    z_half[0] = 0.0                         # the reference's z_half(k=1) = 0 quirk (rrtm_radiation.F90:516)
    t = np.maximum(ts[None] - 6.5e-3 * z_full, 205.0) + rng.normal(0, 1.5, (K, J, I))
    q = np.maximum(rng.uniform(0.003, 0.02, (J, I))[None] * np.exp(-z_full / 2200.0), 1e-8)
    o3 = np.where(p_full < 1.0e4, 8e-6 * np.exp(-((np.log(p_full) - np.log(1.0e3)) ** 2) / 2), 5e-8)
    albedo = rng.uniform(0.05, 0.5, (J, I))
    coszen = np.clip(rng.uniform(-0.3, 1.0, (J, I)), 0.0, 1.0)
    return dict(p_full=p_full, p_half=p_half, z_full=z_full, z_half=z_half, t=t, q=q, o3=o3, t_surf=ts + rng.normal(0, 1, (J, I)),
                albedo=albedo, coszen=coszen)


def rrtm_setup(core, mp, cfg, dt_rad, o3, **kw):
    """attach an oracle RrtmRadiation (do_rrtm_radiation) to the oracle moist physics of tests/test_gpu_moist.build"""
    from oracle import rrtmg as R
    Kk, J, I = core.tg[0].shape
    lat = np.repeat(core.tb.rad_lat[:, None], I, 1)
    lon = np.repeat((np.arange(I) * 360.0 / I * np.pi / 180.0)[None, :], J, 0)
    mp.rrtm = R.RrtmRadiation(lat, lon, cfg.dt_atmos, dt_rad=dt_rad, o3=o3, co2ppmv=360.0, solr_cnst=1360.0, **kw)


def unstable_boundary_layer(core, mp, amp=8.0):
    """warm the lowest 30 % of the oracle model state of tests/test_gpu_moist.build so that the boundary layer is convectively
    unstable: the K-profile diffusivity is then non-zero above the lowest level (the unmodified state is stably stratified and has
    diff_m = diff_t = 0 everywhere)"""
    tr = core.tr
    for lev in (0, 1):
        ps = core.psg[lev]
        zf, zh, pf, ph = core.pg.compute_pressures_and_heights(core.tg[lev], ps, core.surf_geopotential, None)
        sig = pf / ps[None]
        tg = core.tg[lev] + amp * np.maximum(sig - 0.7, 0.0) / 0.3
        ts = tr.grid_to_spherical(tg)
        core.ts[lev] = ts
        core.tg[lev] = tr.spherical_to_grid(ts)
    core.finish_init()
    mp.t_surf = core.tg[core.current][-1] + 3.0
