"""Host-side mirror of the RRTMG interface (include/isca_b200_rrtm.h).

Names and argument meaning follow the Fortran they replace (atmos_param/rrtm_radiation):
  rrtmg_lw_rad: rrtmg_lw     (rrtmg_lw/gcm_model/src/rrtmg_lw_rad.nomcica.f90:81)
  rrtmg_sw_rad: rrtmg_sw     (rrtmg_sw/gcm_model/src/rrtmg_sw_rad.nomcica.f90:73)
  rrtm_radiation: interp_temp + run_rrtmg   (rrtm_radiation.F90:502, 547)
`rrtmg_lw` / `rrtmg_sw` take numpy float64 arrays shaped (ncol, nlay) like the reference's dummy arguments (layer 0 =
lowest layer, hPa, volume mixing ratios); `run_rrtmg` takes the model's [lev, lat, lon] arrays (level 0 = top, Pa).
Everything runs in the CUDA library; there is no CPU path."""
from __future__ import annotations
import ctypes as C
import os
import numpy as np
from .api import load_library, IscaError

RRTM_EXPORTS = ["isca_b200_rrtm_default_config", "isca_b200_rrtm_create", "isca_b200_rrtm_destroy", "isca_b200_rrtm_last_error",
                "isca_b200_rrtmg_lw", "isca_b200_rrtmg_sw", "isca_b200_run_rrtmg", "isca_b200_rrtm_time",
                "isca_b200_rrtm_driver_default_config", "isca_b200_diurnal_solar",
                "isca_b200_moist_use_rrtm", "isca_b200_moist_set_ozone", "isca_b200_moist_set_time", "isca_b200_moist_set_seasonal"]

TABLE_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "rrtmg_tables.bin")


class IscaRrtmConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "num_lon", "num_lat", "num_levels")] + \
               [(n, C.c_double) for n in ("cp_air", "rdgas", "gas_constant", "wtmh2o", "wtmozone", "co2ppmv", "h2o_lower_limit",
                                          "temp_lower_limit", "temp_upper_limit", "solrad", "solr_cnst")] + \
               [("include_secondary_gases", C.c_int)] + \
               [(n, C.c_double) for n in ("ch4_val", "n2o_val", "o2_val", "cfc11_val", "cfc12_val", "cfc22_val", "ccl4_val")] + \
               [(n, C.c_int) for n in ("convert_sphum_to_vmr", "input_o3_file_is_mmr", "lonstep")]


class IscaRrtmDriverConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "dt_rad", "dt_rad_avg", "do_rad_time_avg", "store_intermediate_rad", "solday",
                                       "frierson_solar_rad")] + \
               [(n, C.c_double) for n in ("equinox_day", "del_sol", "del_sw", "ecc", "obliq", "per")] + \
               [("num_angles", C.c_int)] + [(n, C.c_double) for n in ("day_in_s", "year_in_s")]


_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        lib.isca_b200_rrtm_default_config.argtypes = [C.POINTER(IscaRrtmConfigStruct)]
        lib.isca_b200_rrtm_create.argtypes = [C.POINTER(IscaRrtmConfigStruct), C.c_char_p, C.POINTER(vp)]
        lib.isca_b200_rrtm_destroy.argtypes = [vp]
        lib.isca_b200_rrtm_last_error.argtypes = [vp]
        lib.isca_b200_rrtm_last_error.restype = C.c_char_p
        lib.isca_b200_rrtmg_lw.argtypes = [vp, C.c_int, C.c_int] + [dp] * 19
        lib.isca_b200_rrtmg_sw.argtypes = [vp, C.c_int, C.c_int] + [dp] * 11 + [C.c_double, C.c_double] + [dp] * 3
        lib.isca_b200_run_rrtmg.argtypes = [vp] + [dp] * 16
        lib.isca_b200_rrtm_time.argtypes = [vp, C.c_int, C.c_int, dp]
        lib.isca_b200_rrtm_driver_default_config.argtypes = [C.POINTER(IscaRrtmDriverConfigStruct)]
        lib.isca_b200_diurnal_solar.argtypes = [vp, C.POINTER(IscaRrtmDriverConfigStruct), C.c_int, dp, dp, C.c_double, C.c_double, C.c_double,
                                                dp, dp, dp]
        lib.isca_b200_moist_use_rrtm.argtypes = [vp, C.POINTER(IscaRrtmConfigStruct), C.POINTER(IscaRrtmDriverConfigStruct), C.c_char_p]
        lib.isca_b200_moist_set_ozone.argtypes = [vp, dp]
        lib.isca_b200_moist_set_time.argtypes = [vp, C.c_longlong, C.c_int]
        lib.isca_b200_moist_set_seasonal.argtypes = [vp, C.POINTER(IscaRrtmDriverConfigStruct)]
        _bound = True
    return lib


def default_config(**kw) -> IscaRrtmConfigStruct:
    """rrtm_radiation_nml defaults (rrtm_radiation.F90:117-226) + constants_mod values; keyword arguments override"""
    cfg = IscaRrtmConfigStruct()
    _lib().isca_b200_rrtm_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise IscaError(f"unknown rrtm_radiation_nml / config key {k}")
        setattr(cfg, k, v)
    return cfg


def driver_config(**kw) -> IscaRrtmDriverConfigStruct:
    """radiation time stepping / zenith angle values of rrtm_radiation_nml and astronomy_nml; keyword arguments override"""
    dc = IscaRrtmDriverConfigStruct()
    _lib().isca_b200_rrtm_driver_default_config(C.byref(dc))
    for k, v in kw.items():
        if not hasattr(dc, k):
            raise IscaError(f"unknown rrtm_radiation_nml / astronomy_nml key {k}")
        setattr(dc, k, v)
    return dc


def _f(a):
    """(ncol, nlay) array -> the reference's Fortran memory order (column index fastest)"""
    return None if a is None else np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Rrtm:
    """rrtmg_lw_ini + rrtmg_sw_ini + rrtm_radiation_init: loads the reduced coefficient tables onto the device"""

    def __init__(self, cfg: IscaRrtmConfigStruct | None = None, table_file: str = TABLE_FILE, **kw):
        self._lib = _lib()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self._h = C.c_void_p()
        rc = self._lib.isca_b200_rrtm_create(C.byref(self.cfg), table_file.encode(), C.byref(self._h))
        if rc != 0:
            raise IscaError((self._lib.isca_b200_rrtm_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            self._lib.isca_b200_rrtm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise IscaError((self._lib.isca_b200_rrtm_last_error(self._h) or b"").decode())

    def rrtmg_lw(self, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr=None, n2ovmr=None, o2vmr=None,
                 cfc11vmr=None, cfc12vmr=None, cfc22vmr=None, ccl4vmr=None, emis=None):
        """-> uflx, dflx (ncol, nlay+1) [W/m2], hr (ncol, nlay) [K/day]"""
        ncol, nlay = np.shape(play)
        full = lambda v: None if v is None else _f(np.broadcast_to(np.asarray(v, dtype=np.float64), (ncol, nlay)))
        a = [_f(play), _f(plev), _f(tlay), _f(tlev), np.ascontiguousarray(tsfc, dtype=np.float64)] + \
            [full(v) for v in (h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr)] + [_f(emis)]
        uflx = np.zeros((ncol, nlay + 1), order="F")
        dflx = np.zeros((ncol, nlay + 1), order="F")
        hr = np.zeros((ncol, nlay), order="F")
        self._check(self._lib.isca_b200_rrtmg_lw(self._h, ncol, nlay, *[_p(x) for x in a], _p(uflx), _p(dflx), _p(hr)))
        return uflx, dflx, hr

    def rrtmg_sw(self, play, plev, tlay, h2ovmr, o3vmr, co2vmr, ch4vmr=None, n2ovmr=None, o2vmr=None, albedo=0.3, coszen=0.5,
                 adjes=1.0, scon=1368.22):
        """-> swuflx, swdflx (ncol, nlay+1) [W/m2], swhr (ncol, nlay) [K/day]"""
        ncol, nlay = np.shape(play)
        full = lambda v: None if v is None else _f(np.broadcast_to(np.asarray(v, dtype=np.float64), (ncol, nlay)))
        vec = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (ncol,)))
        a = [_f(play), _f(plev), _f(tlay)] + [full(v) for v in (h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr)] + [vec(albedo), vec(coszen)]
        u = np.zeros((ncol, nlay + 1), order="F")
        d = np.zeros((ncol, nlay + 1), order="F")
        hr = np.zeros((ncol, nlay), order="F")
        self._check(self._lib.isca_b200_rrtmg_sw(self._h, ncol, nlay, *[_p(x) for x in a], float(adjes), float(scon), _p(u), _p(d), _p(hr)))
        return u, d, hr

    def run_rrtmg(self, p_full, p_half, z_full, z_half, t, q, t_surf, albedo, coszen, tdt, o3=None):
        """interp_temp + run_rrtmg on [lev, lat, lon] arrays; `tdt` is incremented in place.
        -> dict(tdt_rad, flux_sw, flux_lw, olr, toa_sw)"""
        c = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        K, J, I = np.shape(t)
        if (I, J, K) != (self.cfg.num_lon, self.cfg.num_lat, self.cfg.num_levels):
            raise IscaError("run_rrtmg: array shape does not match the configured grid")
        if not (tdt.flags.c_contiguous and tdt.dtype == np.float64):
            raise IscaError("run_rrtmg: tdt must be a C-contiguous float64 array (it is updated in place)")
        a = [c(p_full), c(p_half), c(z_full), c(z_half), c(t), c(q), c(o3), c(t_surf), c(albedo), c(coszen)]
        out = dict(tdt_rad=np.zeros((K, J, I)), flux_sw=np.zeros((J, I)), flux_lw=np.zeros((J, I)), olr=np.zeros((J, I)),
                   toa_sw=np.zeros((J, I)))
        self._check(self._lib.isca_b200_run_rrtmg(self._h, *[_p(x) for x in a], _p(tdt), _p(out["tdt_rad"]), _p(out["flux_sw"]),
                                                  _p(out["flux_lw"]), _p(out["olr"]), _p(out["toa_sw"])))
        return out

    def diurnal_solar(self, lat, lon, gmt, time_since_ae, dt=None, **astronomy_nml):
        """astronomy_mod diurnal_solar (astronomy.f90:1123) -> cosz, fracday, rrsun"""
        shape = np.shape(lat)
        la = np.ascontiguousarray(np.broadcast_to(lat, shape), dtype=np.float64).ravel()
        lo = np.ascontiguousarray(np.broadcast_to(lon, shape), dtype=np.float64).ravel()
        cosz, frac, rr = np.zeros(la.size), np.zeros(la.size), C.c_double()
        dc = driver_config(**astronomy_nml)
        self._check(self._lib.isca_b200_diurnal_solar(self._h, C.byref(dc), la.size, _p(la), _p(lo), float(gmt), float(time_since_ae),
                                                      -1.0 if dt is None else float(dt), _p(cosz), _p(frac), C.byref(rr)))
        return cosz.reshape(shape), frac.reshape(shape), rr.value

    def time_kernel(self, which: int, reps: int = 10) -> float:
        ms = C.c_double()
        self._check(self._lib.isca_b200_rrtm_time(self._h, which, reps, C.byref(ms)))
        return ms.value
