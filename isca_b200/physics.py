"""Host-side mirror of the column-physics interface (include/isca_b200_physics.h).

Method names and argument meaning follow the Fortran modules they replace:
  sat_vapor_pres_mod: lookup_es_des, compute_qs       (shared/sat_vapor_pres/sat_vapor_pres.F90)
  lscale_cond_mod:    lscale_cond                     (atmos_param/lscale_cond/lscale_cond.F90:79)
  two_stream_gray_rad_mod: two_stream_gray_rad_down / _up  (atmos_param/two_stream_gray_rad/two_stream_gray_rad.F90:386, 659)
  damping_driver_mod: rayleigh                        (atmos_param/damping_driver/damping_driver.f90:594)
Arrays are numpy float64 [lev, lat, lon] (the Fortran (lon, lat, lev) memory order).  Everything runs in the CUDA
library; there is no CPU path."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .api import load_library, IscaError

PHYSICS_EXPORTS = [
    "isca_b200_physics_default_config", "isca_b200_physics_create", "isca_b200_physics_destroy",
    "isca_b200_physics_last_error", "isca_b200_lookup_es_des", "isca_b200_compute_qs", "isca_b200_lscale_cond",
    "isca_b200_two_stream_gray_rad_down", "isca_b200_two_stream_gray_rad_up", "isca_b200_two_stream_gray_rad_set_insolation", "isca_b200_two_stream_gray_rad_set_co2",
    "isca_b200_rayleigh_damping",
    "isca_b200_physics_time", "isca_b200_gcm_vert_diff_down", "isca_b200_get_tri_surf", "isca_b200_mixed_layer_init", "isca_b200_mixed_layer_set_sst",
    "isca_b200_mixed_layer", "isca_b200_gcm_vert_diff_up", "isca_b200_mo_drag", "isca_b200_mo_profile", "isca_b200_stable_mix",
    "isca_b200_mo_diff", "isca_b200_surface_flux", "isca_b200_diffusivity", "isca_b200_qe_moist_convection", "isca_b200_dry_convection",
    "isca_b200_sat_vapor_pres_tables", "isca_b200_betts_miller_default_config", "isca_b200_betts_miller_init", "isca_b200_betts_miller",
]

class IscaBettsMillerConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "do_simp", "do_shallower", "do_changeqref", "do_envsat", "do_taucape")] + \
               [(n, C.c_double) for n in ("tau_bm", "rhbm", "capetaubm", "tau_min", "buoyancy_kick")]


SVP_TABLE_SIZE = 5231


def sat_vapor_pres_tables(**nml):
    """sat_vapor_pres_init_k (sat_vapor_pres_k.F90:161-266) -> TABLE, DTABLE, D2TABLE as the handle builds them for the namelist values
    (tfreeze, hlv, rvgas, es0, sat_vapor_pres_do_simple).  Computed on the host: works without a GPU."""
    lib = _lib()
    cfg = IscaPhysicsConfigStruct()
    lib.isca_b200_physics_default_config(C.byref(cfg))
    for k, v in nml.items():
        if not hasattr(cfg, k):
            raise IscaError(f"unknown physics namelist variable {k}")
        setattr(cfg, k, v)
    out = np.empty(3 * SVP_TABLE_SIZE)
    if lib.isca_b200_sat_vapor_pres_tables(C.byref(cfg), SVP_TABLE_SIZE, out.ctypes.data_as(C.POINTER(C.c_double))) != 0:
        raise IscaError("sat_vapor_pres_init: " + lib.isca_b200_physics_last_error(None).decode())
    return out[:SVP_TABLE_SIZE], out[SVP_TABLE_SIZE:2 * SVP_TABLE_SIZE], out[2 * SVP_TABLE_SIZE:]


def betts_miller_config(**nml) -> "IscaBettsMillerConfigStruct":
    """betts_miller_nml defaults (betts_miller.f90:56-66); keyword arguments override"""
    cfg = IscaBettsMillerConfigStruct()
    _lib().isca_b200_betts_miller_default_config(C.byref(cfg))
    names = {f[0] for f in IscaBettsMillerConfigStruct._fields_}
    for k, v in nml.items():
        if k not in names:
            raise IscaError(f"unknown betts_miller_nml variable {k}")
        setattr(cfg, k, int(v) if k.startswith("do_") else v)
    return cfg


SURFACE_FLUX_IN = ("t_atm", "q_atm", "u_atm", "v_atm", "p_atm", "z_atm", "p_surf", "t_surf", "t_ca", "u_surf", "v_surf",
                   "rough_mom", "rough_heat", "rough_moist", "rough_scale", "gust")
SURFACE_FLUX_OUT = ("flux_t", "flux_q", "flux_r", "flux_u", "flux_v", "cd_m", "cd_t", "cd_q", "w_atm", "u_star", "b_star", "q_star",
                    "dhdt_surf", "dedt_surf", "dedq_surf", "drdt_surf", "dhdt_atm", "dedq_atm", "dtaudu_atm", "dtaudv_atm",
                    "ex_del_m", "ex_del_h", "ex_del_q", "temp_2m", "u_10m", "v_10m", "q_2m", "rh_2m")


class IscaSurfaceFluxArgsStruct(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in SURFACE_FLUX_IN] + [("land", C.POINTER(C.c_int)), ("q_surf", C.POINTER(C.c_double))] + \
               [(n, C.POINTER(C.c_double)) for n in SURFACE_FLUX_OUT]


class IscaPhysicsConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "num_lon", "num_lat", "num_levels")] + \
               [(n, C.c_double) for n in ("grav", "rdgas", "rvgas", "cp_air", "hlv", "tfreeze", "stefan", "pstd_mks", "es0", "hc")] + \
               [("do_evap", C.c_int)] + \
               [(n, C.c_double) for n in ("solar_constant", "del_sol", "del_sw", "ir_tau_eq", "ir_tau_pole", "atm_abs", "sw_diff",
                                          "linear_tau", "wv_exponent", "solar_exponent", "odp", "diabatic_acce",
                                          "trayfric", "sponge_pbottom")] + \
               [(n, C.c_int) for n in ("do_conserve_energy", "vert_diff_do_conserve_energy", "use_virtual_temp_vert_diff", "evaporation")] + \
               [(n, C.c_double) for n in ("rich_crit", "drag_min", "zeta_trans", "vonkarm")] + \
               [(n, C.c_int) for n in ("neutral", "stable_option", "no_neg_q", "use_virtual_temp", "alt_gustiness", "old_dtaudv",
                                       "use_mixing_ratio", "surface_flux_do_simple")] + \
               [(n, C.c_double) for n in ("gust_const", "gust_min", "land_humidity_prefactor", "land_evap_prefactor")] + \
               [(n, C.c_int) for n in ("fixed_depth", "diffusivity_do_entrain", "diffusivity_do_simple", "free_atm_diff", "pbl_mcm",
                                       "use_pog_bug_fix")] + \
               [(n, C.c_double) for n in ("depth_0", "frac_inner", "rich_crit_pbl", "entr_ratio", "parcel_buoy", "znom", "background_m",
                                          "background_t", "tau_bm", "rhbm", "Tmin", "Tmax", "val_inc")] + \
               [("rad_scheme", C.c_int)] + \
               [(n, C.c_double) for n in ("ir_tau_co2_win", "ir_tau_wv_win1", "ir_tau_wv_win2", "ir_tau_co2", "ir_tau_wv1", "ir_tau_wv2",
                                          "window", "carbon_conc", "single_albedo", "back_scatter", "lw_tau_0_gp", "sw_tau_0_gp",
                                          "lw_tau_exponent_gp", "sw_tau_exponent_gp", "bog_a", "bog_b", "bog_mu")] + \
               [("sat_vapor_pres_do_simple", C.c_int), ("free_atm_skyhi_diff", C.c_int), ("ampns", C.c_int)] + \
               [(n, C.c_double) for n in ("rich_crit_diff", "mix_len", "rich_prandtl", "ampns_max")]

RAD_SCHEMES = {"FRIERSON": 0, "BYRNE": 1, "GEEN": 2, "SCHNEIDER": 3}      # two_stream_gray_rad.F90:214-230


_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        lib.isca_b200_physics_default_config.argtypes = [C.POINTER(IscaPhysicsConfigStruct)]
        lib.isca_b200_physics_create.argtypes = [C.POINTER(IscaPhysicsConfigStruct), C.POINTER(vp)]
        lib.isca_b200_physics_destroy.argtypes = [vp]
        lib.isca_b200_physics_last_error.argtypes = [vp]
        lib.isca_b200_physics_last_error.restype = C.c_char_p
        lib.isca_b200_lookup_es_des.argtypes = [vp, C.c_int, dp, dp, dp]
        lib.isca_b200_compute_qs.argtypes = [vp, C.c_int, dp, dp, dp, dp]
        lib.isca_b200_lscale_cond.argtypes = [vp] + [dp] * 7
        lib.isca_b200_two_stream_gray_rad_down.argtypes = [vp] + [dp] * 7
        lib.isca_b200_two_stream_gray_rad_up.argtypes = [vp] + [dp] * 8
        lib.isca_b200_two_stream_gray_rad_set_insolation.argtypes = [vp, dp]
        lib.isca_b200_two_stream_gray_rad_set_co2.argtypes = [vp, C.c_double]
        lib.isca_b200_rayleigh_damping.argtypes = [vp, C.c_double] + [dp] * 7
        lib.isca_b200_physics_time.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        lib.isca_b200_gcm_vert_diff_down.argtypes = [vp, C.c_double] + [dp] * 18
        lib.isca_b200_get_tri_surf.argtypes = [vp, C.c_int, dp]
        lib.isca_b200_mixed_layer_init.argtypes = [vp, dp, dp]
        lib.isca_b200_mixed_layer_set_sst.argtypes = [vp, dp]
        lib.isca_b200_mixed_layer.argtypes = [vp, C.c_double] + [dp] * 13
        lib.isca_b200_gcm_vert_diff_up.argtypes = [vp, C.c_double, dp, dp]
        lib.isca_b200_mo_drag.argtypes = [vp, C.c_int] + [dp] * 12
        lib.isca_b200_mo_profile.argtypes = [vp, C.c_int, C.c_double, C.c_double] + [dp] * 9
        lib.isca_b200_stable_mix.argtypes = [vp, C.c_int, dp, dp]
        lib.isca_b200_mo_diff.argtypes = [vp, C.c_int, C.c_int] + [dp] * 5
        lib.isca_b200_surface_flux.argtypes = [vp, C.POINTER(IscaSurfaceFluxArgsStruct)]
        lib.isca_b200_diffusivity.argtypes = [vp] + [dp] * 13
        ip = C.POINTER(C.c_int)
        lib.isca_b200_qe_moist_convection.argtypes = [vp, C.c_double] + [dp] * 9 + [ip, ip] + [dp] * 5 + [ip]
        lib.isca_b200_dry_convection.argtypes = [vp, C.c_double, C.c_double] + [dp] * 6 + [ip, ip]
        lib.isca_b200_sat_vapor_pres_tables.argtypes = [C.POINTER(IscaPhysicsConfigStruct), C.c_int, dp]
        lib.isca_b200_betts_miller_default_config.argtypes = [C.POINTER(IscaBettsMillerConfigStruct)]
        lib.isca_b200_betts_miller_init.argtypes = [vp, C.POINTER(IscaBettsMillerConfigStruct)]
        lib.isca_b200_betts_miller.argtypes = [vp, C.c_double] + [dp] * 9 + [ip, ip] + [dp] * 6 + [ip]
        _bound = True
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _in(a, shape, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise IscaError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


class ColumnPhysics:
    """One handle per (num_lon, num_lat, num_levels) window; keyword arguments are the namelist variables."""

    def __init__(self, num_lon, num_lat, num_levels, **nml):
        lib = _lib()
        c = IscaPhysicsConfigStruct()
        lib.isca_b200_physics_default_config(C.byref(c))
        c.num_lon, c.num_lat, c.num_levels = num_lon, num_lat, num_levels
        names = {f[0] for f in IscaPhysicsConfigStruct._fields_}
        for k, v in nml.items():
            if k not in names:
                raise IscaError(f"unknown namelist variable {k}")
            if k == "rad_scheme" and isinstance(v, str):          # the namelist value is a string (two_stream_gray_rad.F90:89)
                if v.upper() not in RAD_SCHEMES:
                    raise IscaError(f'two_stream_gray_rad: "{v}" is not a valid radiation scheme.')
                v = RAD_SCHEMES[v.upper()]
            setattr(c, k, v)
        self.config = c
        self._h = C.c_void_p()
        if lib.isca_b200_physics_create(C.byref(c), C.byref(self._h)) != 0:
            raise IscaError("physics_create: " + lib.isca_b200_physics_last_error(None).decode())
        self._lib = lib
        self.s2 = (num_lat, num_lon)
        self.s3 = (num_levels, num_lat, num_lon)
        self.s3h = (num_levels + 1, num_lat, num_lon)

    def close(self):
        if self._h:
            self._lib.isca_b200_physics_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise IscaError(f"{what}: " + self._lib.isca_b200_physics_last_error(self._h).decode())

    def lookup_es_des(self, temp):
        t = np.ascontiguousarray(temp, dtype=np.float64)
        es, des = np.empty_like(t), np.empty_like(t)
        self._ck(self._lib.isca_b200_lookup_es_des(self._h, t.size, _p(t), _p(es), _p(des)), "lookup_es_des")
        return es, des

    def compute_qs(self, temp, press):
        t = np.ascontiguousarray(temp, dtype=np.float64)
        pr = _in(press, t.shape, "press")
        qs, dqs = np.empty_like(t), np.empty_like(t)
        self._ck(self._lib.isca_b200_compute_qs(self._h, t.size, _p(t), _p(pr), _p(qs), _p(dqs)), "compute_qs")
        return qs, dqs

    def lscale_cond(self, tin, qin, pfull, phalf):
        """-> rain [lat, lon], tdel, qdel [lev, lat, lon]"""
        tin, qin, pfull = (_in(a, self.s3, n) for a, n in ((tin, "tin"), (qin, "qin"), (pfull, "pfull")))
        phalf = _in(phalf, self.s3h, "phalf")
        rain, tdel, qdel = np.empty(self.s2), np.empty(self.s3), np.empty(self.s3)
        self._ck(self._lib.isca_b200_lscale_cond(self._h, _p(tin), _p(qin), _p(pfull), _p(phalf), _p(rain), _p(tdel), _p(qdel)),
                 "lscale_cond")
        return rain, tdel, qdel

    def two_stream_gray_rad_down(self, lat, p_half, t, albedo, q=None):
        """-> net_surf_sw_down, surf_lw_down [lat, lon]; q (specific humidity) is read by the byrne and geen schemes"""
        lat, albedo = _in(lat, self.s2, "lat"), _in(albedo, self.s2, "albedo")
        p_half, t = _in(p_half, self.s3h, "p_half"), _in(t, self.s3, "t")
        q = None if q is None else _in(q, self.s3, "q")
        sw, lw = np.empty(self.s2), np.empty(self.s2)
        self._ck(self._lib.isca_b200_two_stream_gray_rad_down(self._h, _p(lat), _p(p_half), _p(t), _p(albedo), _p(q) if q is not None else None,
                                                              _p(sw), _p(lw)), "two_stream_gray_rad_down")
        return sw, lw

    def two_stream_gray_rad_set_co2(self, carbon_conc):
        """do_read_co2 (two_stream_gray_rad.F90:519-521): carbon_conc (ppmv) read from co2_file for the following calls"""
        self._ck(self._lib.isca_b200_two_stream_gray_rad_set_co2(self._h, float(carbon_conc)), "two_stream_gray_rad_down")

    def two_stream_gray_rad_set_insolation(self, insolation):
        """do_seasonal (two_stream_gray_rad.F90:417-447): insolation [lat, lon] = solar_constant * coszen for the following down / up
        calls; None = back to the analytic annual-mean profile"""
        a = None if insolation is None else _in(insolation, self.s2, "insolation")
        self._ck(self._lib.isca_b200_two_stream_gray_rad_set_insolation(self._h, _p(a) if a is not None else None), "two_stream_gray_rad_down")

    def two_stream_gray_rad_up(self, lat, p_half, t, t_surf, albedo, tdt, q=None):
        """-> tdt + radiative heating [lev, lat, lon], olr [lat, lon]"""
        lat, albedo, t_surf = _in(lat, self.s2, "lat"), _in(albedo, self.s2, "albedo"), _in(t_surf, self.s2, "t_surf")
        p_half, t = _in(p_half, self.s3h, "p_half"), _in(t, self.s3, "t")
        q = None if q is None else _in(q, self.s3, "q")
        out = np.array(_in(tdt, self.s3, "tdt"), copy=True)
        olr = np.empty(self.s2)
        self._ck(self._lib.isca_b200_two_stream_gray_rad_up(self._h, _p(lat), _p(p_half), _p(t), _p(t_surf), _p(albedo),
                                                            _p(q) if q is not None else None, _p(out), _p(olr)), "two_stream_gray_rad_up")
        return out, olr

    def rayleigh_damping(self, delt, p_full, u, v, pref):
        """-> udt, vdt, tdt [lev, lat, lon]"""
        p_full, u, v = (_in(a, self.s3, n) for a, n in ((p_full, "p_full"), (u, "u"), (v, "v")))
        pref = _in(pref, (self.s3[0] + 1,), "pref")
        udt, vdt, tdt = np.empty(self.s3), np.empty(self.s3), np.empty(self.s3)
        self._ck(self._lib.isca_b200_rayleigh_damping(self._h, float(delt), _p(p_full), _p(u), _p(v), _p(pref), _p(udt), _p(vdt), _p(tdt)),
                 "rayleigh_damping")
        return udt, vdt, tdt

    TRI_IDS = dict(delta_t=0, dflux_t=1, delta_q=2, dflux_q=3, dtmass=4, delta_u=5, delta_v=6, e_global=16, f_t_global=17, f_q_global=18)

    def gcm_vert_diff_down(self, delt, u, v, t, q, diff_m, diff_t, p_half, p_full, z_full, tau_u, tau_v, dtau_du, dtau_dv,
                           dt_u, dt_v, dt_t, dt_q):
        """-> dict(dt_u, dt_v, dt_t, tau_u, tau_v, dissipative_heat); Tri_surf and e/f factors stay in the handle."""
        a3 = [_in(x, self.s3, n) for x, n in ((u, "u"), (v, "v"), (t, "t"), (q, "q"), (diff_m, "diff_m"), (diff_t, "diff_t"))]
        p_half = _in(p_half, self.s3h, "p_half")
        p_full, z_full = _in(p_full, self.s3, "p_full"), _in(z_full, self.s3, "z_full")
        tau_u, tau_v = np.array(_in(tau_u, self.s2, "tau_u"), copy=True), np.array(_in(tau_v, self.s2, "tau_v"), copy=True)
        dtau_du, dtau_dv = _in(dtau_du, self.s2, "dtau_du"), _in(dtau_dv, self.s2, "dtau_dv")
        dt_u, dt_v, dt_t = (np.array(_in(x, self.s3, n), copy=True) for x, n in ((dt_u, "dt_u"), (dt_v, "dt_v"), (dt_t, "dt_t")))
        dt_q = _in(dt_q, self.s3, "dt_q")
        heat = np.empty(self.s3)
        self._ck(self._lib.isca_b200_gcm_vert_diff_down(self._h, float(delt), *[_p(x) for x in a3], _p(p_half), _p(p_full), _p(z_full),
                                                        _p(tau_u), _p(tau_v), _p(dtau_du), _p(dtau_dv), _p(dt_u), _p(dt_v), _p(dt_t),
                                                        _p(dt_q), _p(heat)), "gcm_vert_diff_down")
        return dict(dt_u=dt_u, dt_v=dt_v, dt_t=dt_t, tau_u=tau_u, tau_v=tau_v, dissipative_heat=heat)

    def tri_surf(self, name):
        i = self.TRI_IDS[name]
        out = np.empty(self.s2 if i < 16 else self.s3)
        self._ck(self._lib.isca_b200_get_tri_surf(self._h, i, _p(out)), "get_tri_surf")
        return out

    def mixed_layer_init(self, heat_capacity, ocean_qflux):
        hc, qf = _in(heat_capacity, self.s2, "heat_capacity"), _in(ocean_qflux, self.s2, "ocean_qflux")
        self._ck(self._lib.isca_b200_mixed_layer_init(self._h, _p(hc), _p(qf)), "mixed_layer_init")

    def mixed_layer_set_sst(self, sst):
        """do_sc_sst: prescribed SST [lat, lon] of the time stepped to; None = slab ocean"""
        self._ck(self._lib.isca_b200_mixed_layer_set_sst(self._h, None if sst is None else _p(_in(sst, self.s2, "sst"))), "mixed_layer_set_sst")

    def mixed_layer(self, dt, t_surf, flux_t, flux_q, flux_r, net_surf_sw_down, surf_lw_down, dhdt_surf, dedt_surf, dedq_surf,
                    drdt_surf, dhdt_atm, dedq_atm):
        """-> new t_surf, delta_t_surf [lat, lon]; Tri_surf delta_t / delta_tr(sphum) are updated inside the handle."""
        ts = np.array(_in(t_surf, self.s2, "t_surf"), copy=True)
        rest = [_in(x, self.s2, "mixed_layer input") for x in (flux_t, flux_q, flux_r, net_surf_sw_down, surf_lw_down, dhdt_surf,
                                                                dedt_surf, dedq_surf, drdt_surf, dhdt_atm, dedq_atm)]
        d = np.empty(self.s2)
        self._ck(self._lib.isca_b200_mixed_layer(self._h, float(dt), _p(ts), *[_p(x) for x in rest], _p(d)), "mixed_layer")
        return ts, d

    def gcm_vert_diff_up(self, delt):
        """-> dt_t, dt_q [lev, lat, lon]"""
        dt_t, dt_q = np.empty(self.s3), np.empty(self.s3)
        self._ck(self._lib.isca_b200_gcm_vert_diff_up(self._h, float(delt), _p(dt_t), _p(dt_q)), "gcm_vert_diff_up")
        return dt_t, dt_q

    def mo_drag(self, pt, pt0, z, z0, zt, zq, speed):
        """monin_obukhov_drag_1d on flat arrays -> drag_m, drag_t, drag_q, u_star, b_star"""
        a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (pt, pt0, z, z0, zt, zq, speed)]
        n = a[0].size
        out = [np.empty(n) for _ in range(5)]
        self._ck(self._lib.isca_b200_mo_drag(self._h, n, *[_p(x) for x in a], *[_p(x) for x in out]), "mo_drag")
        return tuple(out)

    def mo_profile(self, zref, zref_t, z, z0, zt, zq, u_star, b_star):
        """monin_obukhov_profile_1d -> del_m, del_t, del_q"""
        a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (z, z0, zt, zq, u_star, b_star)]
        n = a[0].size
        out = [np.empty(n) for _ in range(3)]
        self._ck(self._lib.isca_b200_mo_profile(self._h, n, float(zref), float(zref_t), *[_p(x) for x in a], *[_p(x) for x in out]), "mo_profile")
        return tuple(out)

    def stable_mix(self, rich):
        r = np.ascontiguousarray(rich, dtype=np.float64)
        mix = np.empty_like(r)
        self._ck(self._lib.isca_b200_stable_mix(self._h, r.size, _p(r), _p(mix)), "stable_mix")
        return mix

    def mo_diff(self, z, u_star, b_star):
        """monin_obukhov_diff: z [nk, n], u_star, b_star [n] -> k_m, k_h [nk, n]"""
        z = np.ascontiguousarray(z, dtype=np.float64)
        nk, n = z.shape
        us, bs = _in(u_star, (n,), "u_star"), _in(b_star, (n,), "b_star")
        km, kh = np.empty_like(z), np.empty_like(z)
        self._ck(self._lib.isca_b200_mo_diff(self._h, n, nk, _p(z), _p(us), _p(bs), _p(km), _p(kh)), "mo_diff")
        return km, kh

    def surface_flux(self, land, q_surf, **inputs):
        """surface_flux (bucket off): keyword inputs named as the reference arguments (SURFACE_FLUX_IN), land a bool/int
        [lat, lon] mask, q_surf [lat, lon] (inout) -> dict of the outputs (SURFACE_FLUX_OUT + q_surf)."""
        a = IscaSurfaceFluxArgsStruct()
        keep = []
        for n in SURFACE_FLUX_IN:
            if n not in inputs:
                raise IscaError(f"surface_flux: missing input {n}")
            x = _in(inputs[n], self.s2, n); keep.append(x); setattr(a, n, _p(x))
        ld = np.ascontiguousarray(land, dtype=np.int32)
        if ld.shape != self.s2:
            raise IscaError("surface_flux: land has the wrong shape")
        a.land = ld.ctypes.data_as(C.POINTER(C.c_int))
        out = {n: np.empty(self.s2) for n in SURFACE_FLUX_OUT}
        out["q_surf"] = np.array(_in(q_surf, self.s2, "q_surf"), copy=True)
        a.q_surf = _p(out["q_surf"])
        for n in SURFACE_FLUX_OUT:
            setattr(a, n, _p(out[n]))
        self._ck(self._lib.isca_b200_surface_flux(self._h, C.byref(a)), "surface_flux")
        return out

    def diffusivity(self, t, q, u, v, p_full, p_half, z_full, z_half, u_star, b_star, k_m=None, k_t=None):
        """-> h [lat, lon], k_m, k_t [lev, lat, lon] (k_m, k_t in are added; default zeros as in vert_turb_driver)"""
        a = [_in(x, self.s3, n) for x, n in ((t, "t"), (q, "q"), (u, "u"), (v, "v"), (p_full, "p_full"))]
        p_half, z_half = _in(p_half, self.s3h, "p_half"), _in(z_half, self.s3h, "z_half")
        z_full = _in(z_full, self.s3, "z_full")
        us, bs = _in(u_star, self.s2, "u_star"), _in(b_star, self.s2, "b_star")
        km = np.zeros(self.s3) if k_m is None else np.array(_in(k_m, self.s3, "k_m"), copy=True)
        kt = np.zeros(self.s3) if k_t is None else np.array(_in(k_t, self.s3, "k_t"), copy=True)
        h = np.empty(self.s2)
        self._ck(self._lib.isca_b200_diffusivity(self._h, *[_p(x) for x in a], _p(p_half), _p(z_full), _p(z_half), _p(us), _p(bs),
                                                 _p(h), _p(km), _p(kt)), "diffusivity")
        return h, km, kt

    def qe_moist_convection(self, dt, Tin, qin, p_full, p_half):
        """-> dict(rain, snow, deltaT, deltaq, qref, convflag, kLZBs, CAPE, CIN, invtau_q_relaxation, invtau_t_relaxation, Tref, kLCLs)"""
        Tin, qin, p_full = (_in(x, self.s3, n) for x, n in ((Tin, "Tin"), (qin, "qin"), (p_full, "p_full")))
        p_half = _in(p_half, self.s3h, "p_half")
        o = {n: np.empty(self.s2) for n in ("rain", "snow", "CAPE", "CIN", "invtau_q_relaxation", "invtau_t_relaxation")}
        o.update({n: np.empty(self.s3) for n in ("deltaT", "deltaq", "qref", "Tref")})
        o.update({n: np.empty(self.s2, dtype=np.int32) for n in ("convflag", "kLZBs", "kLCLs")})
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self._lib.isca_b200_qe_moist_convection(self._h, float(dt), _p(Tin), _p(qin), _p(p_full), _p(p_half), _p(o["rain"]), _p(o["snow"]),
                                                         _p(o["deltaT"]), _p(o["deltaq"]), _p(o["qref"]), ip(o["convflag"]), ip(o["kLZBs"]),
                                                         _p(o["CAPE"]), _p(o["CIN"]), _p(o["invtau_q_relaxation"]),
                                                         _p(o["invtau_t_relaxation"]), _p(o["Tref"]), ip(o["kLCLs"])), "qe_moist_convection")
        return o

    def betts_miller_init(self, **nml):
        """betts_miller_nml (tau_bm, rhbm, do_simp, do_shallower, do_changeqref, do_envsat, buoyancy_kick)"""
        cfg = betts_miller_config(**nml)
        self._ck(self._lib.isca_b200_betts_miller_init(self._h, C.byref(cfg)), "betts_miller_init")

    def betts_miller(self, dt, tin, qin, pfull, phalf):
        """betts_miller (betts_miller.f90:86) -> dict with the names of qe_moist_convection: rain, snow, deltaT (tdel), deltaq (qdel),
        qref, Tref, convflag (bmflag), kLZBs, kLCLs, CAPE, CIN, invtau_t_relaxation, invtau_q_relaxation, capeflag"""
        tin, qin, pfull = (_in(x, self.s3, n) for x, n in ((tin, "tin"), (qin, "qin"), (pfull, "pfull")))
        phalf = _in(phalf, self.s3h, "phalf")
        o = {n: np.empty(self.s2) for n in ("rain", "snow", "CAPE", "CIN", "invtau_q_relaxation", "invtau_t_relaxation", "capeflag")}
        o.update({n: np.empty(self.s3) for n in ("deltaT", "deltaq", "qref", "Tref")})
        o.update({n: np.empty(self.s2, dtype=np.int32) for n in ("convflag", "kLZBs", "kLCLs")})
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self._lib.isca_b200_betts_miller(self._h, float(dt), _p(tin), _p(qin), _p(pfull), _p(phalf), _p(o["rain"]), _p(o["snow"]),
                                                  _p(o["deltaT"]), _p(o["deltaq"]), _p(o["qref"]), ip(o["convflag"]), ip(o["kLZBs"]),
                                                  _p(o["CAPE"]), _p(o["CIN"]), _p(o["Tref"]), _p(o["invtau_t_relaxation"]),
                                                  _p(o["invtau_q_relaxation"]), _p(o["capeflag"]), ip(o["kLCLs"])), "betts_miller")
        return o

    def dry_convection(self, tau, gamma, tg, p_full, p_half):
        """dry_convection (dry_convection.f90:105) -> dict(dt_tg, cape, cin, lzb, lcl)"""
        tg, p_full = (_in(x, self.s3, n) for x, n in ((tg, "tg"), (p_full, "p_full")))
        p_half = _in(p_half, self.s3h, "p_half")
        o = dict(dt_tg=np.empty(self.s3), cape=np.empty(self.s2), cin=np.empty(self.s2), lzb=np.empty(self.s2, dtype=np.int32),
                 lcl=np.empty(self.s2, dtype=np.int32))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self._lib.isca_b200_dry_convection(self._h, float(tau), float(gamma), _p(tg), _p(p_full), _p(p_half), _p(o["dt_tg"]),
                                                    _p(o["cape"]), _p(o["cin"]), ip(o["lzb"]), ip(o["lcl"])), "dry_convection")
        return o

    def time_kernel(self, which, reps=20):
        """(ms per launch, algorithmic bytes per launch) on resident synthetic columns."""
        ms, by = C.c_double(), C.c_double()
        self._ck(self._lib.isca_b200_physics_time(self._h, which, reps, C.byref(ms), C.byref(by)), "physics_time")
        return ms.value, by.value
