"""betts_miller_mod, the full Betts-Miller scheme (convection_scheme = 'FULL_BETTS_MILLER') without a GPU.

(1) pin: the LCL table the reference ships (tests/golden/bm_lcltable.py) satisfies its defining relation with the do_simple
    saturation vapour pressure, i.e. the oracle's escomp / lcltabl are consistent with the reference's own numbers;
(2) oracle/betts_miller.py: conservation properties of every energy-correction branch, agreement of capecalcnew with the independent
    restatement of the simplified scheme's CAPE routine where they coincide;
(3) the `__host__ __device__` column code of isca_b200/csrc/physics_bm_column.h -- what the CUDA kernel of physics_bm.cu executes
    per column -- built for the host (tests/host/bm_host.cpp, test infrastructure) against the oracle on 600 random columns per
    option set: integer outputs bit-exact, fields 1e-12 relative to the field maximum;
(4) the C ABI: symbols, struct layout, wrapper argument counts."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from oracle import physics as P
from oracle.betts_miller import BettsMiller, BettsMillerConfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "bm_host.cpp")
OUT = os.path.join(HERE, "host", "_build", "libbm_host.so")
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def host():
    deps = [SRC] + [os.path.join(ROOT, "isca_b200", "csrc", f) for f in ("physics_bm_column.h", "bm_lcltable.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.bm_host_lcltable.restype = DP
    return lib


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def columns(n, K=20, seed=0):
    """n random columns [K, 1, n] spanning no CAPE / shallow / deep convection (see test_branch_coverage)"""
    rng = np.random.default_rng(seed)
    svp = P.SatVaporPres()
    sig_h = np.linspace(0, 1, K + 1) ** 1.3
    ps = 1e5 + 3000.0 * rng.standard_normal(n)
    ph = sig_h[:, None] * ps[None]
    pf = 0.5 * (ph[1:] + ph[:-1])
    g, rh, ts, e = rng.uniform(0.17, 0.3, n), rng.uniform(0.2, 1.05, n), rng.uniform(270, 305, n), rng.uniform(0, 0.8, n)
    t = np.maximum(ts[None] * (pf / 1e5) ** g[None], 200.0) + 0.5 * rng.standard_normal((K, n))
    qs, _ = svp.compute_qs(t, pf)
    q = rh[None] * qs * (pf / 1e5) ** e[None]
    q[:, 0] = 0.0                                            # a completely dry column (`r0 <= 0` branch)
    t[-1, 1] = 340.0                                         # a hot surface parcel: buoyant up to high levels
    r3 = lambda a: np.ascontiguousarray(a.reshape(a.shape[0], 1, n))
    return svp, r3(t), r3(q), r3(pf), r3(ph)


# ---------------------------------------------------------------------------------------------------------------------------
def test_reference_lcl_table_is_consistent_with_escomp():
    """known-answer data: T_i of the reference's table solves value_i = log(es(T_i)/T_i**(1/kappa)), value_i = -23 + 0.1 i"""
    from golden.bm_lcltable import LCLTABLE
    T = np.array(LCLTABLE)
    assert T.size == 127 and np.all(np.diff(T) > 0)
    svp = P.SatVaporPres()
    bm = BettsMiller(svp)
    es = np.array([bm.escomp(x) for x in T])
    value = np.log(es / T ** (1 / P.KAPPA))
    assert np.abs(value - (-23.0 + 0.1 * np.arange(127))).max() < 2e-6          # 8 printed digits of T
    # lcltabl interpolates linearly in value and clamps
    assert bm.lcltabl(-23.0) == T[0] and bm.lcltabl(-30.0) == T[0] and bm.lcltabl(-10.4) == T[126] and bm.lcltabl(0.0) == T[126]
    assert abs(bm.lcltabl(-17.25) - 0.5 * (T[57] + T[58])) < 1e-9
    # and inverts the relation between the nodes to the accuracy of linear interpolation
    for v in (-22.96, -18.513, -12.0001):
        tl = bm.lcltabl(v)
        assert abs(np.log(bm.escomp(tl) / tl ** (1 / P.KAPPA)) - v) < 2e-4


def test_capecalcnew_agrees_with_the_simplified_scheme_where_they_coincide():
    """qe_moist_convection's CAPE_calculation descends from capecalcnew; without virtual-temperature effects (do_virtual = false is
    not an option there) they differ, but the LCL level and the dry-adiabatic part below it follow the same formulas"""
    svp, t, q, pf, ph = columns(80, seed=5)
    bm = BettsMiller(svp)
    sbm = P.SBMConvection(svp, Tmin=173.0, Tmax=335.0)
    same = 0
    for i in range(2, 80):
        o = bm.column(1800.0, t[:, 0, i], q[:, 0, i], pf[:, 0, i], ph[:, 0, i])
        s = sbm.column(1800.0, t[:, 0, i], q[:, 0, i], pf[:, 0, i], ph[:, 0, i])
        if "kLCL" in s and s["kLCL"] > 0 and o["klcl"] > 0:
            assert abs(int(s["kLCL"]) - o["klcl"]) <= 1, i       # table (0.1 spacing, 8 digits) vs Newton-built table (0.01)
            same += 1
    assert same > 30


OPTION_SETS = [dict(), dict(do_simp=False), dict(do_shallower=True), dict(do_changeqref=True), dict(do_envsat=True, rhbm=0.7),
               dict(do_simp=False, do_shallower=True, buoyancy_kick=1.5, tau_bm=3600.0)]


@pytest.mark.parametrize("nml", OPTION_SETS)
def test_branch_coverage_and_conservation(nml):
    svp, t, q, pf, ph = columns(240, seed=1)
    bm = BettsMiller(svp, BettsMillerConfig(**nml))
    o = bm(1800.0, t, q, pf, ph)
    flags = np.bincount(o["convflag"].ravel(), minlength=3)
    assert flags[0] > 10 and flags[1] > 10 and flags[2] > 4, flags
    dp = ph[1:] - ph[:-1]
    water = (o["deltaq"] * dp).sum(0) / P.GRAV + o["rain"]
    energy = ((P.CP_AIR * o["deltaT"] + P.HLV * o["deltaq"]) * dp).sum(0) / P.GRAV
    scale = np.abs(P.HLV * o["deltaq"] * dp).sum(0).max() / P.GRAV
    deep = o["convflag"] == 2
    assert np.abs(water[deep]).max() < 1e-12 * max(o["rain"].max(), 1e-30) + 1e-15      # rain = -int(qdel dp)/g
    assert np.abs(energy[deep]).max() < 1e-9 * scale                                     # cp dT + L dq integrates to zero
    assert np.all(o["rain"] >= 0.0) and np.all(o["rain"][~deep] == 0.0)
    shallow = o["convflag"] == 1
    if nml.get("do_shallower") or nml.get("do_changeqref"):
        adj = shallow & (np.abs(o["deltaq"]).sum(0) > 0)
        assert adj.sum() > 3
        assert np.abs(water[adj]).max() < 1e-10 * np.abs(o["deltaq"] * dp).sum(0).max() / P.GRAV     # no net condensation
        assert np.abs(energy[adj]).max() < 1e-9 * scale
    else:
        assert np.all(o["deltaq"][:, shallow] == 0.0) and np.all(o["deltaT"][:, shallow] == 0.0)
    none = o["convflag"] == 0
    assert np.all(o["qref"][:, none] == q[:, none]) and np.all(o["Tref"][:, none] == t[:, none]) and np.all(o["kLZB"][none] == 0)
    assert np.all(o["CAPE"][~none] > 0)
    # above the level of zero buoyancy nothing is changed
    for j, i in zip(*np.nonzero(~none)):
        kz = o["kLZB"][j, i]
        assert np.all(o["deltaT"][:max(kz - 1, 0), j, i] == 0.0)


def run_host(lib, svp, nml, dt, t, q, pf, ph):
    K, J, I = t.shape
    n = J * I
    c = BettsMillerConfig(**nml)
    table = np.concatenate([svp.TABLE, svp.DTABLE, svp.D2TABLE]).astype(np.float64)
    sp = np.array([svp.tminl, svp.dtinvl, svp.tepsl, svp.dtres])
    cfg = np.array([c.tau_bm, c.rhbm, c.buoyancy_kick, P.RDGAS, P.RVGAS, P.CP_AIR, P.HLV, P.KAPPA, P.GRAV, 1.0])
    flags = np.array([c.do_simp, c.do_shallower, c.do_changeqref, c.do_envsat], dtype=np.int32)
    a = [np.ascontiguousarray(x) for x in (t, q, pf, ph)]
    o3 = [np.zeros((K, J, I)) for _ in range(4)]
    o2 = [np.zeros((J, I)) for _ in range(5)]
    oi = [np.zeros((J, I), dtype=np.int32) for _ in range(3)]
    d = lambda x: x.ctypes.data_as(DP)
    bad = lib.bm_host_run(n, K, C.c_double(dt), d(table), svp.TABLE.size, d(sp), d(cfg), flags.ctypes.data_as(IP), *[d(x) for x in a],
                          d(o2[0]), *[d(x) for x in o3], *[x.ctypes.data_as(IP) for x in oi], *[d(x) for x in o2[1:]])
    assert bad == 0
    return dict(rain=o2[0], deltaT=o3[0], deltaq=o3[1], qref=o3[2], Tref=o3[3], convflag=oi[0], kLZB=oi[1], kLCL=oi[2], CAPE=o2[1], CIN=o2[2],
                invtau_t=o2[3], invtau_q=o2[4])


@pytest.mark.parametrize("nml", OPTION_SETS)
def test_device_column_code_matches_oracle(host, nml):
    assert np.array_equal(np.ctypeslib.as_array(host.bm_host_lcltable(), (127,)), BettsMiller(P.SatVaporPres()).lcltable)
    svp, t, q, pf, ph = columns(600, seed=2 + len(nml))
    o = BettsMiller(svp, BettsMillerConfig(**nml))(1200.0, t, q, pf, ph)
    h = run_host(host, svp, nml, 1200.0, t, q, pf, ph)
    for n in ("convflag", "kLZB", "kLCL"):
        assert np.array_equal(h[n], o[n]), n
    for n in ("rain", "deltaT", "deltaq", "qref", "Tref", "CAPE", "CIN", "invtau_t", "invtau_q"):
        assert rel(h[n], o[n]) < 1e-12, n
    assert np.bincount(o["convflag"].ravel(), minlength=3).min() > 5


# ---------------------------------------------------------------------------------------------------------------------------
def test_betts_miller_abi(lib_built, tmp_path):
    from isca_b200 import physics, moist
    from test_wrappers_stub import _StubLib, _nparams
    lib = physics._lib()
    for s in ("isca_b200_betts_miller_default_config", "isca_b200_betts_miller_init", "isca_b200_betts_miller", "isca_b200_moist_set_betts_miller"):
        assert hasattr(lib, s), s
    st = physics.IscaBettsMillerConfigStruct
    body = "".join(f'printf("%zu\\n", offsetof(IscaBettsMillerConfig, {n}));' for n, _ in st._fields_)
    src = tmp_path / "l.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200_physics.h"\nint main(){printf("%zu\\n", sizeof(IscaBettsMillerConfig));'
                   + body + "return 0;}")
    exe = tmp_path / "l"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(st)
    for (n, _), off in zip(st._fields_, out[1:]):
        assert getattr(st, n).offset == off, n
    cfg = physics.betts_miller_config(do_simp=False, rhbm=0.7)
    d = BettsMillerConfig()
    assert (cfg.abi_version, cfg.do_simp, cfg.rhbm) == (1, 0, 0.7)
    for n in ("tau_bm", "capetaubm", "tau_min", "buoyancy_kick"):
        assert getattr(cfg, n) == getattr(d, n)
    with pytest.raises(physics.IscaError):
        physics.betts_miller_config(nonsense=1)
    # wrapper argument counts against a recording stub
    stub = _StubLib()
    cp = physics.ColumnPhysics.__new__(physics.ColumnPhysics)
    cp._lib, cp._h = stub, C.c_void_p(1)
    cp.s2, cp.s3, cp.s3h = (4, 8), (10, 4, 8), (11, 4, 8)
    cp.betts_miller_init(do_shallower=True)
    o = cp.betts_miller(1800.0, np.ones(cp.s3), np.ones(cp.s3), np.ones(cp.s3), np.ones(cp.s3h))
    assert o["convflag"].dtype == np.int32 and o["deltaT"].shape == cp.s3 and o["rain"].shape == cp.s2
    seen = dict(stub.calls)
    assert seen["isca_b200_betts_miller"] == _nparams("isca_b200_physics.h", "isca_b200_betts_miller")
    assert seen["isca_b200_betts_miller_init"] == _nparams("isca_b200_physics.h", "isca_b200_betts_miller_init")
    m = moist.MoistAtmosphere.__new__(moist.MoistAtmosphere)
    m._lib, m._h = stub, C.c_void_p(1)
    m.set_betts_miller(do_changeqref=True)
    assert dict(stub.calls)["isca_b200_moist_set_betts_miller"] == _nparams("isca_b200_physics.h", "isca_b200_moist_set_betts_miller")
    assert moist.CONVECTION["FULL_BETTS_MILLER"] == 3
    cp._h = C.c_void_p()
    m._h = C.c_void_p()
