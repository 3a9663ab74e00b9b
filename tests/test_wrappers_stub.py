"""The Python mirrors of the RRTMG / moist-model entry points against a recording stub of the C library (no GPU, no compute):
every wrapper method can be called with arrays of the documented shapes, passes the number of arguments its C prototype has,
and rejects wrong shapes.  (The real library refuses to create handles without a CUDA device, so these code paths would otherwise
only execute on the GPU box.)"""
import ctypes as C
import re
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Fn:
    def __init__(self, name, log):
        self.name, self.log = name, log
        self.argtypes = None
        self.restype = C.c_int

    def __call__(self, *a):
        self.log.append((self.name, len(a)))
        return b"" if self.name.endswith("last_error") else 0


class _StubLib:
    def __init__(self):
        self.calls = []
        self._fns = {}

    def __getattr__(self, name):
        if name.startswith("isca_b200_"):
            return self._fns.setdefault(name, _Fn(name, self.calls))
        raise AttributeError(name)


def _nparams(header, fn):
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)
    m = re.search(r"\b" + fn + r"\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    return len([p for p in m.group(1).split(",") if p.strip()])


def test_rrtm_wrapper_methods(lib_built):
    from isca_b200 import rrtm
    stub = _StubLib()
    r = rrtm.Rrtm.__new__(rrtm.Rrtm)
    r._lib, r._h = stub, C.c_void_p(1)
    r.cfg = rrtm.default_config(num_lon=8, num_lat=4, num_levels=10)
    nc, K = 12, 10
    a = np.ones((nc, K))
    ah = np.ones((nc, K + 1))
    u, d, hr = r.rrtmg_lw(a, ah, a, ah, np.ones(nc), a, 1e-7, 3e-4, emis=np.ones((nc, 16)))
    assert u.shape == (nc, K + 1) and hr.shape == (nc, K) and u.flags.f_contiguous
    r.rrtmg_sw(a, ah, a, a, a, 3e-4, albedo=0.2, coszen=np.linspace(0, 1, nc))
    K3 = (10, 4, 8)
    f3, f3h, f2 = np.ones(K3), np.ones((11, 4, 8)), np.ones((4, 8))
    out = r.run_rrtmg(f3, f3h, f3, f3h, f3, f3, f2, f2, f2, np.zeros(K3), o3=f3)
    assert set(out) == {"tdt_rad", "flux_sw", "flux_lw", "olr", "toa_sw"}
    with pytest.raises(rrtm.IscaError):
        r.run_rrtmg(np.ones((10, 4, 9)), f3h, f3, f3h, np.ones((10, 4, 9)), f3, f2, f2, f2, np.zeros(K3))
    r.diurnal_solar(f2, f2, 1.0, 2.0, dt=0.5, ecc=0.01)
    r.time_kernel(0, 2)
    want = {"isca_b200_rrtmg_lw": "isca_b200_rrtm.h", "isca_b200_rrtmg_sw": "isca_b200_rrtm.h", "isca_b200_run_rrtmg": "isca_b200_rrtm.h",
            "isca_b200_diurnal_solar": "isca_b200_rrtm.h", "isca_b200_rrtm_time": "isca_b200_rrtm.h"}
    seen = dict(stub.calls)
    for fn, hdr in want.items():
        assert seen[fn] == _nparams(hdr, fn), fn
    r._h = C.c_void_p()          # nothing to destroy


def test_moist_wrapper_methods(lib_built, monkeypatch):
    from isca_b200 import moist, rrtm
    stub = _StubLib()
    monkeypatch.setattr(rrtm, "_lib", lambda: stub)
    # rrtm.default_config / driver_config go through rrtm._lib(): give the stub's config functions the real defaults
    real = rrtm.load_library()
    stub._fns["isca_b200_rrtm_default_config"] = real.isca_b200_rrtm_default_config
    stub._fns["isca_b200_rrtm_driver_default_config"] = real.isca_b200_rrtm_driver_default_config
    m = moist.MoistAtmosphere.__new__(moist.MoistAtmosphere)
    m._lib, m._h = stub, C.c_void_p(1)
    m.s2, m.s3 = (4, 8), (10, 4, 8)
    m.core = None
    m.use_rrtm(dict(co2ppmv=360.0), dt_rad=1800, solday=90, do_rad_time_avg=0)
    m.set_ozone(np.ones(m.s3))
    m.set_ozone(None)
    m.set_time(3, 43200)
    m.set_seasonal(solday=-10, equinox_day=0.75, use_time_average_coszen=True, dt_rad_avg=1800, obliq=25.0)
    m.set_ocean_qflux(np.ones(m.s2))
    m.set_co2(420.0)
    m.set_dry_convection(7200.0, 0.7)
    for name in moist.MoistAtmosphere.SURFACE_FIELDS:
        m.set_surface(name, np.ones(m.s2))
    for name in ("coszen", "olr", "toa_sw", "tdt_rad", "delta_t_surf", "diff_t"):
        assert m.get(name).shape in (m.s2, m.s3)
    with pytest.raises(moist.IscaError):
        m.set_surface("albedo", np.ones((4, 9)))
    with pytest.raises(moist.IscaError):
        m.set_surface("nonsense", np.ones(m.s2))
    with pytest.raises(moist.IscaError):
        m.use_rrtm(not_a_namelist_value=1)
    seen = dict(stub.calls)
    for fn, hdr in (("isca_b200_moist_use_rrtm", "isca_b200_rrtm.h"), ("isca_b200_moist_set_ozone", "isca_b200_rrtm.h"),
                    ("isca_b200_moist_set_time", "isca_b200_rrtm.h"), ("isca_b200_moist_set_seasonal", "isca_b200_rrtm.h"), ("isca_b200_moist_set_co2", "isca_b200_physics.h"),
                    ("isca_b200_moist_set_ocean_qflux", "isca_b200_physics.h"),
                    ("isca_b200_moist_set_dry_convection", "isca_b200_physics.h"), ("isca_b200_moist_set_surface", "isca_b200_physics.h"),
                    ("isca_b200_moist_get", "isca_b200_physics.h")):
        assert seen[fn] == _nparams(hdr, fn), fn
    m._h = C.c_void_p()


def test_physics_dry_convection_wrapper(lib_built):
    from isca_b200 import physics
    stub = _StubLib()
    cp = physics.ColumnPhysics.__new__(physics.ColumnPhysics)
    cp._lib, cp._h = stub, C.c_void_p(1)
    cp.s2, cp.s3, cp.s3h = (4, 8), (10, 4, 8), (11, 4, 8)
    o = cp.dry_convection(7200.0, 0.7, np.ones(cp.s3), np.ones(cp.s3), np.ones(cp.s3h))
    assert o["lzb"].dtype == np.int32 and o["dt_tg"].shape == cp.s3
    assert dict(stub.calls)["isca_b200_dry_convection"] == _nparams("isca_b200_physics.h", "isca_b200_dry_convection")
    cp.two_stream_gray_rad_set_insolation(np.ones(cp.s2))
    cp.two_stream_gray_rad_set_insolation(None)
    cp.two_stream_gray_rad_set_co2(400.0)
    assert dict(stub.calls)["isca_b200_two_stream_gray_rad_set_co2"] == _nparams("isca_b200_physics.h", "isca_b200_two_stream_gray_rad_set_co2")
    assert dict(stub.calls)["isca_b200_two_stream_gray_rad_set_insolation"] == \
        _nparams("isca_b200_physics.h", "isca_b200_two_stream_gray_rad_set_insolation")
    with pytest.raises(physics.IscaError):
        cp.two_stream_gray_rad_set_insolation(np.ones((4, 9)))
    cp._h = C.c_void_p()
