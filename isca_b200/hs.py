"""Python mirror of include/isca_b200_hs.h: hs_forcing_mod with the namelist options beyond the Held-Suarez default
(atmos_param/hs_forcing/hs_forcing.F90) and the dry model driven by it (atmosphere.F90:120-352).

    hs = HsForcing(num_lon, num_lat, num_levels, equilibrium_t_option="EXOPLANET", ...)      # hs_forcing_init
    udt, vdt, tdt, rdt, diag = hs.hs_forcing(dt, (days, seconds), lon, lat, p_half, p_full, u, v, t, ...)

    m = HsAtmosphere(dyn_config, equilibrium_t_option="top_down", ...)                        # atmosphere_init
    m.core.set_grid_state(...); m.hs_forcing_init(); m.atmosphere(n)

Arrays are [lev, lat, lon] (level 0 = model top).  There is no CPU path: the library refuses to create handles without a GPU."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .api import load_library, IscaError, IscaConfigStruct, Atmosphere

HS_EXPORTS = ["isca_b200_hs_forcing_default_config", "isca_b200_hs_last_error", "isca_b200_hs_forcing_create", "isca_b200_hs_forcing_destroy",
              "isca_b200_hs_forcing", "isca_b200_hs_forcing_get_tg_prev", "isca_b200_hs_forcing_set_tg_prev", "isca_b200_hs_model_create",
              "isca_b200_hs_model_destroy", "isca_b200_hs_model_dycore", "isca_b200_hs_model_set_time", "isca_b200_hs_model_set_tg_prev", "isca_b200_hs_model_init",
              "isca_b200_hs_model_step", "isca_b200_hs_model_get"]

EQUILIBRIUM_T = {"HELD_SUAREZ": 0, "EXOPLANET": 1, "EXOPLANET2": 2, "TOP_DOWN": 3}
STRATOSPHERE_T = {"extend_tp": 0, "c_above_tp": 1, "hs_like": 2}
LOCAL_HEATING = {"": 0, "Isidoro": 1}
MODEL_FIELDS = dict(teq=(0, 3), h_trop=(1, 2), tg_prev=(2, 2), tdt=(3, 3))


class IscaHsForcingConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "num_lon", "num_lat", "num_levels", "no_forcing", "do_conserve_energy",
                                       "equilibrium_t_option", "stratosphere_t_option", "local_heating_option", "num_angles")] + \
               [(n, C.c_double) for n in ("t_zero", "t_strat", "delh", "delv", "eps", "sigma_b", "P00", "p_trop", "alpha", "ka", "ks", "kf",
                                          "trflux", "trsink", "local_heating_srfamp", "local_heating_xwidth", "local_heating_ywidth",
                                          "local_heating_xcenter", "local_heating_ycenter", "local_heating_vert_decay", "peri_time",
                                          "smaxis", "albedo", "lapse", "h_a", "tau_s", "heat_capacity", "ml_depth", "spinup_time",
                                          "kappa", "rdgas", "grav", "stefan", "solar_const", "omega", "orbital_period", "orbital_rate",
                                          "ecc", "obliq", "per")]


_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        vp, dp, cp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(IscaHsForcingConfigStruct)
        lib.isca_b200_hs_forcing_default_config.argtypes = [cp]
        lib.isca_b200_hs_last_error.argtypes = []
        lib.isca_b200_hs_last_error.restype = C.c_char_p
        lib.isca_b200_hs_forcing_create.argtypes = [cp, dp, C.c_longlong, C.c_int, C.POINTER(vp)]
        lib.isca_b200_hs_forcing_destroy.argtypes = [vp]
        lib.isca_b200_hs_forcing.argtypes = [vp, C.c_double, C.c_longlong, C.c_int] + [dp] * 17 + [C.c_int, dp, dp]
        lib.isca_b200_hs_forcing_get_tg_prev.argtypes = [vp, dp]
        lib.isca_b200_hs_forcing_set_tg_prev.argtypes = [vp, dp]
        lib.isca_b200_hs_model_create.argtypes = [C.POINTER(IscaConfigStruct), cp, C.POINTER(vp)]
        lib.isca_b200_hs_model_destroy.argtypes = [vp]
        lib.isca_b200_hs_model_dycore.argtypes = [vp]
        lib.isca_b200_hs_model_dycore.restype = vp
        lib.isca_b200_hs_model_set_time.argtypes = [vp, C.c_longlong, C.c_int]
        lib.isca_b200_hs_model_set_tg_prev.argtypes = [vp, dp]
        lib.isca_b200_hs_model_init.argtypes = [vp]
        lib.isca_b200_hs_model_step.argtypes = [vp, C.c_int]
        lib.isca_b200_hs_model_get.argtypes = [vp, C.c_int, dp]
        _bound = True
    return lib


def hs_config(**nml) -> IscaHsForcingConfigStruct:
    """hs_forcing_nml defaults (hs_forcing.F90:74-107); keyword arguments override.  The option strings of the namelist are accepted
    (equilibrium_t_option = 'Held_Suarez' | 'EXOPLANET' | 'EXOPLANET2' | 'top_down', stratosphere_t_option, local_heating_option)."""
    cfg = IscaHsForcingConfigStruct()
    _lib().isca_b200_hs_forcing_default_config(C.byref(cfg))
    names = {f[0] for f in IscaHsForcingConfigStruct._fields_}
    for k, v in nml.items():
        if k not in names:
            raise IscaError(f"unknown hs_forcing_nml variable {k}")
        if k == "equilibrium_t_option" and isinstance(v, str):
            if v.upper() not in EQUILIBRIUM_T:
                raise IscaError(f'hs_forcing_nml: "{v}"  is not a valid value for equilibrium_t_option')
            v = EQUILIBRIUM_T[v.upper()]
        if k == "stratosphere_t_option" and isinstance(v, str):
            v = STRATOSPHERE_T.get(v, 3)                     # any other string: teq = max(teq, 0) (:1000)
        if k == "local_heating_option" and isinstance(v, str):
            if v not in LOCAL_HEATING:
                raise IscaError(f'hs_forcing_nml: "{v}"  is not a valid value for local_heating_option')
            v = LOCAL_HEATING[v]
        setattr(cfg, k, v)
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _in(a, shape, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise IscaError(f"{name} has shape {a.shape}, expected {tuple(shape)}")
    return a


class HsForcing:
    """hs_forcing_init + hs_forcing on host arrays"""

    def __init__(self, num_lon, num_lat, num_levels, lat=None, time=(0, 0), **nml):
        self._lib = _lib()
        self.cfg = hs_config(num_lon=num_lon, num_lat=num_lat, num_levels=num_levels, **nml)
        self.s2, self.s3, self.s3h = (num_lat, num_lon), (num_levels, num_lat, num_lon), (num_levels + 1, num_lat, num_lon)
        la = None if lat is None else _in(lat, self.s2, "lat")
        self._h = C.c_void_p()
        if self._lib.isca_b200_hs_forcing_create(C.byref(self.cfg), _p(la), int(time[0]), int(time[1]), C.byref(self._h)) != 0:
            raise IscaError("hs_forcing_init: " + self._lib.isca_b200_hs_last_error().decode())

    def _ck(self, rc, where):
        if rc != 0:
            raise IscaError(f"{where}: " + self._lib.isca_b200_hs_last_error().decode())

    def hs_forcing(self, dt, time, lon, lat, p_half, p_full, u, v, t, udt, vdt, tdt, um=None, vm=None, rm=None, rdt=None, zfull=None):
        """-> udt, vdt, tdt, rdt (incremented copies), dict(teq=, h_trop=).  time = (days, seconds) of Time; rm, rdt: [ntr, lev, lat, lon]"""
        lon, lat = _in(lon, self.s2, "lon"), _in(lat, self.s2, "lat")
        p_half = _in(p_half, self.s3h, "p_half")
        a3 = lambda a, n: _in(a, self.s3, n)
        p_full, u, v, t = a3(p_full, "p_full"), a3(u, "u"), a3(v, "v"), a3(t, "t")
        um = u if um is None else a3(um, "um")
        vm = v if vm is None else a3(vm, "vm")
        zf = None if zfull is None else a3(zfull, "zfull")
        udt, vdt, tdt = (np.array(a3(x, n), copy=True) for x, n in ((udt, "udt"), (vdt, "vdt"), (tdt, "tdt")))
        ntr = 0
        if rdt is not None:
            rdt = np.array(np.ascontiguousarray(rdt, dtype=np.float64), copy=True)
            rm = np.ascontiguousarray(rm, dtype=np.float64)
            if rdt.ndim != 4 or rdt.shape[1:] != self.s3 or rm.shape != rdt.shape:
                raise IscaError("rm / rdt have the wrong shape")
            ntr = rdt.shape[0]
        teq, h_trop = np.empty(self.s3), np.zeros(self.s2)
        self._ck(self._lib.isca_b200_hs_forcing(self._h, float(dt), int(time[0]), int(time[1]), _p(lon), _p(lat), _p(p_half), _p(p_full),
                                                _p(u), _p(v), _p(t), None, _p(um), _p(vm), _p(t), _p(rm) if ntr else None, _p(udt), _p(vdt),
                                                _p(tdt), _p(rdt) if ntr else None, _p(zf), ntr, _p(teq), _p(h_trop)), "hs_forcing")
        return udt, vdt, tdt, rdt, dict(teq=teq, h_trop=h_trop)

    @property
    def tg_prev(self):
        out = np.empty(self.s2)
        self._ck(self._lib.isca_b200_hs_forcing_get_tg_prev(self._h, _p(out)), "tg_prev")
        return out

    @tg_prev.setter
    def tg_prev(self, a):
        self._ck(self._lib.isca_b200_hs_forcing_set_tg_prev(self._h, _p(_in(a, self.s2, "tg_prev"))), "tg_prev")

    def hs_forcing_end(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.isca_b200_hs_forcing_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.hs_forcing_end()
        except Exception:
            pass


class HsAtmosphere:
    """atmosphere_init / atmosphere / atmosphere_end of the dry model with the general hs_forcing (state resident on the device)"""

    def __init__(self, dyn_config, **hs_nml):
        self._lib = _lib()
        self.cfg = hs_config(**hs_nml)
        self._h = C.c_void_p()
        if self._lib.isca_b200_hs_model_create(C.byref(dyn_config), C.byref(self.cfg), C.byref(self._h)) != 0:
            raise IscaError("atmosphere_init: " + self._lib.isca_b200_hs_last_error().decode())
        self.core = Atmosphere(dyn_config, _adopt_handle=self._lib.isca_b200_hs_model_dycore(self._h))
        self.s2 = (dyn_config.lat_max, dyn_config.lon_max)
        self.s3 = (dyn_config.num_levels,) + self.s2

    def _ck(self, rc, where):
        if rc != 0:
            raise IscaError(f"{where}: " + self._lib.isca_b200_hs_last_error().decode())

    def set_time(self, days, seconds):
        self._ck(self._lib.isca_b200_hs_model_set_time(self._h, int(days), int(seconds)), "set_time")

    def set_tg_prev(self, tg_prev):
        """restart of the top_down option (INPUT/hs_forcing.res.nc); before hs_forcing_init, which then skips the spin-up"""
        self._ck(self._lib.isca_b200_hs_model_set_tg_prev(self._h, _p(_in(tg_prev, self.s2, "tg_prev"))), "hs_forcing_init")

    def hs_forcing_init(self):
        """after the atmospheric state is in place and the clock is set"""
        self._ck(self._lib.isca_b200_hs_model_init(self._h), "hs_forcing_init")

    def atmosphere(self, n_steps=1):
        self._ck(self._lib.isca_b200_hs_model_step(self._h, int(n_steps)), "atmosphere")

    def get(self, name):
        if name not in MODEL_FIELDS:
            raise IscaError(f"unknown field {name}")
        i, nd = MODEL_FIELDS[name]
        out = np.empty(self.s3 if nd == 3 else self.s2)
        self._ck(self._lib.isca_b200_hs_model_get(self._h, i, _p(out)), "get")
        return out

    def atmosphere_end(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.core.h = None                      # owned by the model
            self._lib.isca_b200_hs_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.atmosphere_end()
        except Exception:
            pass
