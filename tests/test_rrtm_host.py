"""The RRTMG device arithmetic without a GPU.

isca_b200/csrc/rrtm_column.h holds every formula of the CUDA kernels as `__host__ __device__` functions and
rrtm_tables.h the band descriptors; tests/host/rrtm_host.cpp (TEST INFRASTRUCTURE, built here with g++, never part of the
product library) runs the same functions in serial loops.  Comparing it with the independent NumPy oracle checks the
descriptor-driven band code, the table re-tiling and the radiative-transfer sweeps that the GPU executes; the
`-m gpu` tests (tests/test_gpu_rrtm.py) then only have the launch geometry and the cross-thread reductions left to prove.
Tolerances: optical depths 1e-13, fluxes 1e-12, heating rates 1e-10 relative to the field maximum."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from oracle import rrtmg as R
from rrtm_cases import columns, mls_column, zero_if_none as z

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "rrtm_host.cpp")
OUT = os.path.join(HERE, "host", "_build", "librrtm_host.so")
P = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    deps = [SRC] + [os.path.join(HERE, "..", "isca_b200", "csrc", f) for f in ("rrtm_column.h", "rrtm_tables.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.rrtm_host_error.restype = C.c_char_p
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(P)


def _F(a):
    return None if a is None else np.asfortranarray(a)


def host_lw(lib, g, cp=287.04 / (2 / 7)):
    nc, K = g["play"].shape
    arr = {k: _F(v) for k, v in g.items()}
    u = np.zeros((nc, K + 1), order="F")
    d = np.zeros((nc, K + 1), order="F")
    hr = np.zeros((nc, K), order="F")
    tg = np.zeros((nc, K, 140))
    fr = np.zeros((nc, K, 140))
    rc = lib.rrtm_host_lw(R.TABLE_FILE.encode(), C.c_double(cp), nc, K, _ptr(arr["play"]), _ptr(arr["plev"]), _ptr(arr["tlay"]),
                          _ptr(arr["tlev"]), _ptr(np.ascontiguousarray(g["tsfc"])), _ptr(arr["h2o"]), _ptr(arr["o3"]), _ptr(arr["co2"]),
                          _ptr(arr["ch4"]), _ptr(arr["n2o"]), _ptr(arr["o2"]), _ptr(arr["cfc11"]), _ptr(arr["cfc12"]), _ptr(arr["cfc22"]),
                          _ptr(arr["ccl4"]), None, _ptr(u), _ptr(d), _ptr(hr), _ptr(tg), _ptr(fr))
    assert rc == 0, lib.rrtm_host_error()
    return u, d, hr, tg, fr


def host_sw(lib, g, alb, cz, adjes=1.0, scon=1368.22, cp=287.04 / (2 / 7)):
    nc, K = g["play"].shape
    arr = {k: _F(v) for k, v in g.items()}
    u = np.zeros((nc, K + 1), order="F")
    d = np.zeros((nc, K + 1), order="F")
    hr = np.zeros((nc, K), order="F")
    tg = np.zeros((nc, K, 112))
    tr = np.zeros((nc, K, 112))
    sf = np.zeros((nc, 112))
    rc = lib.rrtm_host_sw(R.TABLE_FILE.encode(), C.c_double(cp), nc, K, _ptr(arr["play"]), _ptr(arr["plev"]), _ptr(arr["tlay"]),
                          _ptr(arr["h2o"]), _ptr(arr["o3"]), _ptr(arr["co2"]), _ptr(arr["ch4"]), _ptr(arr["n2o"]), _ptr(arr["o2"]),
                          _ptr(np.ascontiguousarray(alb)), _ptr(np.ascontiguousarray(cz)), C.c_double(adjes), C.c_double(scon),
                          _ptr(u), _ptr(d), _ptr(hr), _ptr(tg), _ptr(tr), _ptr(sf))
    assert rc == 0, lib.rrtm_host_error()
    return u, d, hr, tg, tr, sf


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("secondary,K,seed", [(False, 40, 1), (True, 40, 2), (True, 25, 3), (False, 60, 4)])
def test_lw_device_arithmetic_matches_oracle(host, secondary, K, seed):
    g = columns(48, K, seed, secondary=secondary)
    u, d, hr, tg, fr = host_lw(host, g)
    ou, od, ohr, opt = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]),
                                  z(g["n2o"]), z(g["o2"]), z(g["cfc11"]), z(g["cfc12"]), z(g["cfc22"]), z(g["ccl4"]), return_optics=True)
    ngs = np.concatenate([[0], np.cumsum(R.LW_NGC)])
    for b in range(16):                                  # per band: every descriptor is exercised
        s = slice(ngs[b], ngs[b + 1])
        assert rel(tg[:, :, s], opt["taug"][:, :, s]) < 1e-13, b + 1
        assert np.abs(fr[:, :, s] - opt["fracs"][:, :, s]).max() < 1e-14, b + 1
    # both branches of every band were visited
    assert opt["sc"]["lower"].any() and (~opt["sc"]["lower"]).any()
    assert rel(u, ou) < 1e-12 and rel(d, od) < 1e-12 and rel(hr, ohr) < 1e-10


@pytest.mark.parametrize("secondary,K,seed", [(False, 40, 5), (True, 40, 6), (True, 30, 7)])
def test_sw_device_arithmetic_matches_oracle(host, secondary, K, seed):
    g = columns(48, K, seed, secondary=secondary)
    rng = np.random.default_rng(seed)
    alb = rng.uniform(0.0, 0.9, 48)
    cz = rng.uniform(-0.2, 1.0, 48)
    cz[:4] = [1e-11, 1e-9, 1.0, 0.01]
    su, sd, shr, tg, tr, sf = host_sw(host, g, alb, cz, 1.03, 1360.0)
    ou, od, ohr, opt = R.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]), z(g["n2o"]), z(g["o2"]),
                                  alb, cz, 1.03, 1360.0, return_optics=True)
    day = cz >= 1e-10
    ngs = np.concatenate([[0], np.cumsum(R.SW_NGC)])
    for b in range(14):
        s = slice(ngs[b], ngs[b + 1])
        assert rel(tg[day][:, :, s], opt["taug"][day][:, :, s] + 1e-300) < 1e-13 or np.abs(tg[day][:, :, s] - opt["taug"][day][:, :, s]).max() < 1e-300, b + 16
        assert rel(tr[day][:, :, s], opt["taur"][day][:, :, s]) < 1e-13, b + 16
        assert rel(sf[day][:, s], opt["sflux"][day][:, s]) < 1e-14, b + 16
    assert (su[~day] == 0).all() and (sd[~day] == 0).all() and (shr[~day] == 0).all()
    assert rel(su, ou) < 1e-12 and rel(sd, od) < 1e-12 and rel(shr, ohr) < 1e-10


def test_mls_column_host(host):
    g = mls_column(nc=2)
    gg = dict(g, ch4=None, n2o=None, o2=None, cfc11=None, cfc12=None, cfc22=None, ccl4=None)
    u, d, hr, _, _ = host_lw(host, gg)
    assert 278.0 < u[0, -1] < 287.0 and 340.0 < d[0, 0] < 355.0
    su, sd, shr, _, _, sf = host_sw(host, gg, np.full(2, 0.2), np.array([0.5, 1.0]))
    assert np.allclose(sd[:, -1], 1368.22 * np.array([0.5, 1.0]), rtol=2e-3)
    assert np.allclose(sf.sum(axis=1), sd[:, -1] / np.array([0.5, 1.0]), rtol=1e-12)


# ----------------------------------------------------------------------------------------------------------------
# the CUDA kernels themselves under CPU thread emulation (tests/host/rrtm_emu.cpp): same kernel bodies (rrtm_kernels.h),
# one OS thread per CUDA thread, std::barrier for __syncthreads, per-warp exchange buffers for the shuffles
# ----------------------------------------------------------------------------------------------------------------
EMU_SRC = os.path.join(HERE, "host", "rrtm_emu.cpp")
EMU_OUT = os.path.join(HERE, "host", "_build", "librrtm_emu.so")


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(HERE, "..", "isca_b200", "csrc", f) for f in ("rrtm_column.h", "rrtm_tables.h", "rrtm_kernels.h")]
    if not os.path.exists(EMU_OUT) or any(os.path.getmtime(d) > os.path.getmtime(EMU_OUT) for d in deps):
        os.makedirs(os.path.dirname(EMU_OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-o", EMU_OUT, EMU_SRC])
    return C.CDLL(EMU_OUT)


def _cp(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(P)


@pytest.mark.parametrize("secondary,K,seed,alternate", [(False, 40, 3, False), (True, 25, 4, False), (False, 40, 5, True), (True, 31, 6, True)])
def test_emulated_lw_sw_kernels_match_oracle(emu, secondary, K, seed, alternate, monkeypatch):
    """the production kernels (rrtmg_lw_setcoef_kernel + rrtmg_lw_col_kernel: a lane = a column, a warp = a band group;
    rrtmg_sw_kernel: a thread = a g-point, shared-memory staging, shuffle reductions with padding lanes, night columns leaving
    early) and, with `alternate`, the other mapping of each (rrtmg_lw_kernel, rrtmg_sw_col_kernel: selectable at run time)"""
    if alternate:
        monkeypatch.setenv("RRTM_EMU_LW_GPOINT", "1")
        monkeypatch.setenv("RRTM_EMU_SW_COL", "1")
    nc = 6
    g = columns(nc, K, seed, secondary=secondary)
    arr = {k: _F(v) for k, v in g.items()}
    cp = C.c_double(287.04 / (2 / 7))
    u = np.zeros((nc, K + 1), order="F")
    d = np.zeros((nc, K + 1), order="F")
    hr = np.zeros((nc, K), order="F")
    emis = np.asfortranarray(np.tile(np.linspace(0.9, 1.0, 16), (nc, 1)))
    rc = emu.rrtm_emu_lw(R.TABLE_FILE.encode(), cp, nc, K, _ptr(arr["play"]), _ptr(arr["plev"]), _ptr(arr["tlay"]), _ptr(arr["tlev"]),
                         _cp(g["tsfc"]), _ptr(arr["h2o"]), _ptr(arr["o3"]), _ptr(arr["co2"]), _ptr(arr["ch4"]), _ptr(arr["n2o"]), _ptr(arr["o2"]),
                         _ptr(arr["cfc11"]), _ptr(arr["cfc12"]), _ptr(arr["cfc22"]), _ptr(arr["ccl4"]), _ptr(emis), _ptr(u), _ptr(d), _ptr(hr))
    assert rc == 0
    ou, od, ohr = R.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]), z(g["n2o"]),
                             z(g["o2"]), z(g["cfc11"]), z(g["cfc12"]), z(g["cfc22"]), z(g["ccl4"]), emis=np.ascontiguousarray(emis))
    assert rel(u, ou) < 1e-12 and rel(d, od) < 1e-12 and rel(hr, ohr) < 1e-10
    alb = np.linspace(0.0, 0.8, nc)
    cz = np.array([1e-11, 1.0, 0.3, -0.1, 0.01, 0.7])
    su = np.zeros((nc, K + 1), order="F")
    sd = np.zeros((nc, K + 1), order="F")
    shr = np.zeros((nc, K), order="F")
    rc = emu.rrtm_emu_sw(R.TABLE_FILE.encode(), cp, nc, K, _ptr(arr["play"]), _ptr(arr["plev"]), _ptr(arr["tlay"]), _ptr(arr["h2o"]),
                         _ptr(arr["o3"]), _ptr(arr["co2"]), _ptr(arr["ch4"]), _ptr(arr["n2o"]), _ptr(arr["o2"]), _cp(alb), _cp(cz),
                         C.c_double(1.03), C.c_double(1360.0), _ptr(su), _ptr(sd), _ptr(shr))
    assert rc == 0
    osu, osd, oshr = R.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], g["co2"], z(g["ch4"]), z(g["n2o"]), z(g["o2"]), alb, cz,
                                1.03, 1360.0)
    assert (su[cz < 1e-10] == 0).all() and (shr[cz < 1e-10] == 0).all()
    assert rel(su, osu) < 1e-12 and rel(sd, osd) < 1e-12 and rel(shr, oshr) < 1e-10


@pytest.mark.parametrize("lonstep", [1, 2])
def test_emulated_run_rrtmg_kernels_match_oracle(emu, lonstep):
    """the kernel sequence of run_rrtmg: rrtm_prepare_kernel (layout reversal, interp_temp, units, limits, lonstep gather),
    rrtm_fix_top_kernel, SW, LW, rrtm_finish_kernel (K/day -> K/s, surface / TOA fluxes, lonstep interpolation)"""
    from rrtm_cases import model_columns
    I, J, K = 4, 2, 20
    m = model_columns(I, J, K, 5)
    keep = {k: np.ascontiguousarray(v) for k, v in m.items()}
    tdt0 = np.random.default_rng(0).normal(0, 1e-5, (K, J, I))
    tdt = tdt0.copy()
    tr, fs, fl, olr, ts = np.zeros((K, J, I)), np.zeros((J, I)), np.zeros((J, I)), np.zeros((J, I)), np.zeros((J, I))
    d = C.c_double
    rc = emu.rrtm_emu_run_rrtmg(R.TABLE_FILE.encode(), I, J, K, lonstep, d(287.04 / (2 / 7)), d(R.RDGAS), d(R.GAS_CONSTANT), d(R.WTMH2O),
                                d(R.WTMOZONE), d(300.0), d(2e-7), d(100.0), d(370.0), d(1.0), d(1368.22),
                                *[_cp(keep[k]) for k in ("p_full", "p_half", "z_full", "z_half", "t", "q", "o3", "t_surf", "albedo", "coszen")],
                                _cp(tdt), _cp(tr), _cp(fs), _cp(fl), _cp(olr), _cp(ts))
    assert rc == 0
    lat = np.zeros((J, I))
    o = R.RrtmRadiation(lat, lat, 600.0, o3=m["o3"], lonstep=lonstep)
    o.zenith = lambda s: m["coszen"]
    t2, fsw, flw = o(0.0, m["p_full"], m["p_half"], m["z_full"], m["z_half"], m["t"], m["q"], m["t_surf"], m["albedo"], tdt0.copy())
    assert rel(tr, o.tdt_rad) < 1e-10 and rel(tdt, t2) < 1e-10
    assert rel(fs, fsw) < 1e-12 and rel(fl, flw) < 1e-12 and rel(olr, o.olr) < 1e-12 and rel(ts, o.toa_sw) < 1e-12


def test_emulated_coszen_kernel_matches_oracle(emu):
    a = R.Astronomy()
    lat = np.repeat(np.linspace(-np.pi / 2, np.pi / 2, 17)[:, None], 16, 1)
    lon = np.repeat(np.linspace(0, 2 * np.pi, 16, endpoint=False)[None, :], 17, 0)
    d = C.c_double
    for dt in (None, 7200 / 86400 * 2 * np.pi, 2 * np.pi):
        for gmt, tsae in ((0.0, 0.0), (1.0, 0.3), (4.5, 2.0), (6.2, 5.5)):
            dec = a.declination(a.angle(tsae))
            cz, fr = np.zeros(lat.size), np.zeros(lat.size)
            emu.rrtm_emu_coszen(lat.size, _cp(lat.ravel()), _cp(lon.ravel()), d(gmt), d(dec), d(-1.0 if dt is None else dt), 0, d(0.95), d(0.0),
                                _cp(cz), _cp(fr))
            oc, of, _ = a.diurnal_solar(lat, lon, gmt, tsae, dt)
            assert np.abs(cz.reshape(lat.shape) - oc).max() < 1e-14 and np.abs(fr.reshape(lat.shape) - of).max() < 1e-13, (dt, gmt)
    # frierson_solar_rad
    cz = np.zeros(lat.size)
    emu.rrtm_emu_coszen(lat.size, _cp(lat.ravel()), _cp(lon.ravel()), d(0.0), d(0.0), d(-1.0), 1, d(0.95), d(0.1), _cp(cz), None)
    p2 = (1.0 - 3.0 * np.sin(lat) ** 2) / 4.0
    assert np.abs(cz.reshape(lat.shape) - 0.25 * (1.0 + 0.95 * p2 + 0.1 * np.sin(lat))).max() < 1e-15


def dry_columns(J, I, K, seed):
    """columns with unstable layers near the surface, stable caps, elevated unstable layers above a cloud top, fully stable ones"""
    rng = np.random.default_rng(seed)
    ps = rng.uniform(9.0e4, 1.03e5, (J, I))
    ph = (np.linspace(0.0, 1.0, K + 1) ** 1.3)[:, None, None] * ps[None]
    ph[0] = 10.0
    pf = 0.5 * (ph[1:] + ph[:-1])
    theta = 300.0 + rng.uniform(-1.0, 1.0, (J, I))[None] + np.cumsum(rng.normal(0.6, 2.5, (K, J, I))[::-1], axis=0)[::-1]
    tg = theta * (pf / 1.0e5) ** (2.0 / 7.0)
    tg[:, 0, 0] = 250.0 * (pf[:, 0, 0] / 1.0e5) ** 0.1            # very stable column: no convection at all
    return tg, pf, ph


def test_emulated_dry_convection_kernel_matches_oracle(emu):
    """dry_convection_kernel under thread emulation against the NumPy restatement of dry_convection / capecalc"""
    from oracle import physics as PH
    J, I, K = 6, 20, 18
    tg, pf, ph = dry_columns(J, I, K, 3)
    tau, gamma = 14400.0, 0.7
    dt, cape, cin, lzb, lcl = PH.dry_convection(tg, pf, ph, tau, gamma)
    assert (lzb < K).any() and (lzb == K).any() and (cape > 0).any()          # convecting and non-convecting columns
    nc = J * I
    o_dt, o_cape, o_cin = np.zeros((K, J, I)), np.zeros((J, I)), np.zeros((J, I))
    o_lzb, o_lcl, err = np.zeros((J, I), dtype=np.int32), np.zeros((J, I), dtype=np.int32), np.zeros(1, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    d = C.c_double
    emu.emu_dry_convection(nc, K, d(tau), d(gamma), d(PH.RDGAS / PH.CP_AIR), d(PH.RDGAS), _cp(tg), _cp(pf), _cp(ph), _cp(o_dt), _cp(o_cape),
                           _cp(o_cin), ip(o_lzb), ip(o_lcl), ip(err))
    assert err[0] == 0
    assert np.array_equal(o_lzb, lzb) and np.array_equal(o_lcl, lcl)
    assert np.abs(o_dt - dt).max() <= 1e-13 * np.abs(dt).max()
    assert np.abs(o_cape - cape).max() <= 1e-12 * cape.max() and np.abs(o_cin - cin).max() <= 1e-12 * max(cin.max(), 1e-30)
    # energy conservation of the adjustment: the mass-weighted temperature change vanishes in every column
    dp = ph[1:] - ph[:-1]
    assert np.abs((dt * dp).sum(axis=0)).max() < 1e-12 * np.abs(dt * dp).sum(axis=0).max() + 1e-9
