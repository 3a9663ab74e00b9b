"""idealized_moist_model mirror (include/isca_b200_physics.h, section idealized_moist_model): the atmosphere_mod boundary
with idealized_moist_phys as physics (atmos_spectral/driver/solo/atmosphere.F90:263-266, 300-302 and
idealized_moist_phys.F90).  The whole step -- column physics and spectral dynamics -- runs on the GPU."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .api import load_library, IscaError, Atmosphere, IscaConfigStruct
from .physics import IscaPhysicsConfigStruct, _lib as _physics_lib

MOIST_EXPORTS = ["isca_b200_moist_default_config", "isca_b200_moist_create", "isca_b200_moist_create_ranked", "isca_b200_moist_destroy", "isca_b200_moist_last_error",
                 "isca_b200_moist_dycore", "isca_b200_moist_init", "isca_b200_moist_step", "isca_b200_moist_get",
                 "isca_b200_moist_set_t_surf", "isca_b200_moist_set_sst", "isca_b200_moist_set_surface", "isca_b200_moist_set_dry_convection", "isca_b200_moist_set_betts_miller", "isca_b200_moist_set_co2", "isca_b200_moist_set_ocean_qflux", "isca_b200_moist_timing",
                 "isca_b200_moist_profile_step", "isca_b200_moist_step_io", "isca_b200_moist_io_sync", "isca_b200_moist_io_wait"]

FIELDS_2D = dict(t_surf=0, precip=1, flux_t=2, flux_q=3, z_pbl=4, net_surf_sw_down=5, surf_lw_down=6, conv_rain=7, cape=8, convflag=9,
                 q_surf=10, u_star=11, b_star=12, flux_u=13, flux_v=14, delta_t_surf=15, coszen=16, olr=17, toa_sw=18)
FIELDS_3D = dict(dt_ug=32, dt_vg=33, dt_tg=34, dt_qg=35, diff_m=36, diff_t=37, tdt_rad=38)
# declared in include/isca_b200_rrtm.h (RRTMG as the moist model's radiation)
MOIST_RRTM_EXPORTS = ["isca_b200_moist_use_rrtm", "isca_b200_moist_set_ozone", "isca_b200_moist_set_time", "isca_b200_moist_set_seasonal"]
CONVECTION = {"NONE": 0, "SIMPLE_BETTS_MILLER": 1, "DRY": 2, "FULL_BETTS_MILLER": 3}


class IscaMoistConfigStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("abi_version", "convection_scheme", "do_damping")] + \
               [(n, C.c_double) for n in ("roughness_mom", "roughness_heat", "roughness_moist", "mixed_layer_depth", "albedo_value", "rho_cp",
                                          "constant_gust")] + [("use_tau", C.c_int)]


_bound = False


def _lib():
    global _bound
    lib = _physics_lib()
    if not _bound:
        vp, dp = C.c_void_p, C.POINTER(C.c_double)
        lib.isca_b200_moist_default_config.argtypes = [C.POINTER(IscaMoistConfigStruct)]
        lib.isca_b200_moist_create.argtypes = [C.POINTER(IscaConfigStruct), C.POINTER(IscaPhysicsConfigStruct), C.POINTER(IscaMoistConfigStruct),
                                               C.POINTER(vp)]
        lib.isca_b200_moist_create_ranked.argtypes = [C.POINTER(IscaConfigStruct), C.POINTER(IscaPhysicsConfigStruct),
                                                      C.POINTER(IscaMoistConfigStruct), C.c_int, C.c_int, vp, C.POINTER(vp)]
        lib.isca_b200_moist_destroy.argtypes = [vp]
        lib.isca_b200_moist_last_error.argtypes = [vp]
        lib.isca_b200_moist_last_error.restype = C.c_char_p
        lib.isca_b200_moist_dycore.argtypes = [vp]
        lib.isca_b200_moist_dycore.restype = vp
        lib.isca_b200_moist_init.argtypes = [vp]
        lib.isca_b200_moist_step.argtypes = [vp, C.c_int]
        lib.isca_b200_moist_get.argtypes = [vp, C.c_int, dp]
        lib.isca_b200_moist_set_t_surf.argtypes = [vp, dp]
        lib.isca_b200_moist_set_sst.argtypes = [vp, dp]
        lib.isca_b200_moist_set_ocean_qflux.argtypes = [vp, dp]
        lib.isca_b200_moist_set_surface.argtypes = [vp, C.c_int, dp]
        lib.isca_b200_moist_set_dry_convection.argtypes = [vp, C.c_double, C.c_double]
        lib.isca_b200_moist_set_co2.argtypes = [vp, C.c_double]
        from .physics import IscaBettsMillerConfigStruct
        lib.isca_b200_moist_set_betts_miller.argtypes = [vp, C.POINTER(IscaBettsMillerConfigStruct)]
        lib.isca_b200_moist_timing.argtypes = [vp, dp, dp]
        lib.isca_b200_moist_profile_step.argtypes = [vp, C.c_int, dp, C.c_int, C.c_char_p, C.c_int]
        ip = C.POINTER(C.c_int)
        lib.isca_b200_moist_step_io.argtypes = [vp, vp, C.c_int, ip, ip, ip, C.POINTER(vp)]
        lib.isca_b200_moist_io_sync.argtypes = [vp]
        lib.isca_b200_moist_io_wait.argtypes = [vp, C.c_int]
        _bound = True
    return lib


class MoistAtmosphere:
    """atmosphere_init / atmosphere / atmosphere_end of the idealized moist model.

    dyn_config: IscaConfigStruct of the dynamical core (api.make_config / config_from_namelist_object, num_tracers = 1);
    physics_nml: keyword namelist values of the column schemes (IscaPhysicsConfig field names);
    convection_scheme, do_damping, roughness_*, mixed_layer_depth, albedo_value, constant_gust: idealized_moist_phys_nml /
    mixed_layer_nml / vert_turb_driver_nml values."""

    def __init__(self, dyn_config, physics_nml=None, convection_scheme="SIMPLE_BETTS_MILLER", rank=0, nranks=1, nccl_unique_id=None,
                 **moist_nml):
        lib = _lib()
        self._lib = lib
        pc = IscaPhysicsConfigStruct()
        lib.isca_b200_physics_default_config(C.byref(pc))
        names = {f[0] for f in IscaPhysicsConfigStruct._fields_}
        for k, v in (physics_nml or {}).items():
            if k not in names:
                raise IscaError(f"unknown physics namelist variable {k}")
            if k == "rad_scheme" and isinstance(v, str):
                from .physics import RAD_SCHEMES
                if v.upper() not in RAD_SCHEMES:
                    raise IscaError(f'two_stream_gray_rad: "{v}" is not a valid radiation scheme.')
                v = RAD_SCHEMES[v.upper()]
            setattr(pc, k, v)
        mc = IscaMoistConfigStruct()
        lib.isca_b200_moist_default_config(C.byref(mc))
        if convection_scheme.upper() not in CONVECTION:
            raise IscaError(f"idealized_moist_phys: {convection_scheme} is not a valid convection scheme of this build")
        mc.convection_scheme = CONVECTION[convection_scheme.upper()]
        mnames = {f[0] for f in IscaMoistConfigStruct._fields_}
        for k, v in moist_nml.items():
            if k not in mnames:
                raise IscaError(f"unknown namelist variable {k}")
            setattr(mc, k, v)
        self._h = C.c_void_p()
        uid = None
        if nccl_unique_id is not None:
            self._uid_buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            uid = C.cast(self._uid_buf, C.c_void_p)
        if lib.isca_b200_moist_create_ranked(C.byref(dyn_config), C.byref(pc), C.byref(mc), rank, nranks, uid, C.byref(self._h)) != 0:
            raise IscaError("atmosphere_init: " + lib.isca_b200_moist_last_error(None).decode())
        self.core = Atmosphere(dyn_config, _adopt_handle=lib.isca_b200_moist_dycore(self._h))
        self.core.Jloc = dyn_config.lat_max // nranks
        self.s2 = (dyn_config.lat_max // nranks, dyn_config.lon_max)
        self.s3 = (dyn_config.num_levels,) + self.s2

    def _ck(self, rc, where):
        if rc != 0:
            raise IscaError(f"{where}: " + self._lib.isca_b200_moist_last_error(self._h).decode())

    def idealized_moist_phys_init(self):
        """after the atmospheric state is in place (cold start or set_grid_state on .core)"""
        self._ck(self._lib.isca_b200_moist_init(self._h), "idealized_moist_phys_init")

    def atmosphere(self, n_steps=1):
        self._ck(self._lib.isca_b200_moist_step(self._h, n_steps), "atmosphere")

    def get(self, name, out=None):
        """host copy of a physics field; `out`: an existing C-contiguous float64 array of the field's shape (e.g. pinned memory)"""
        if name in FIELDS_2D:
            shape, i = self.s2, FIELDS_2D[name]
        elif name in FIELDS_3D:
            shape, i = self.s3, FIELDS_3D[name]
        else:
            raise IscaError(f"unknown field {name}")
        if out is None:
            out = np.empty(shape)
        elif out.shape != tuple(shape) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise IscaError(f"get({name}): out must be a C-contiguous float64 array of shape {tuple(shape)}")
        self._ck(self._lib.isca_b200_moist_get(self._h, i, out.ctypes.data_as(C.POINTER(C.c_double))), "get")
        return out

    def use_rrtm(self, rrtm_nml=None, table_file=None, **driver_nml):
        """do_rrtm_radiation = .true. (idealized_moist_phys.F90:1167-1177): RRTMG replaces two_stream_gray_rad.
        rrtm_nml: rrtm_radiation_nml gas / limit values (IscaRrtmConfig fields); driver_nml: dt_rad, dt_rad_avg, do_rad_time_avg,
        store_intermediate_rad, solday, equinox_day, frierson_solar_rad, del_sol, del_sw and the astronomy_nml / calendar values.
        Call before idealized_moist_phys_init."""
        from . import rrtm
        rc = rrtm.default_config(**(rrtm_nml or {}))
        dc = rrtm.driver_config(**driver_nml)
        lib = rrtm._lib()
        self._ck(lib.isca_b200_moist_use_rrtm(self._h, C.byref(rc), C.byref(dc), (table_file or rrtm.TABLE_FILE).encode()), "rrtm_radiation_init")

    def set_seasonal(self, solday=-10, equinox_day=0.75, use_time_average_coszen=False, dt_rad_avg=-1, **astronomy_nml):
        """two_stream_gray_rad_nml do_seasonal = .true. (two_stream_gray_rad.F90:83-87, 417-447): insolation = solar_constant *
        coszen(Time) from astronomy_mod every step.  astronomy_nml: ecc, obliq, per, num_angles, day_in_s, year_in_s.
        Call before idealized_moist_phys_init."""
        from . import rrtm
        dc = rrtm.driver_config(solday=int(solday), equinox_day=float(equinox_day), do_rad_time_avg=int(bool(use_time_average_coszen)),
                                dt_rad_avg=int(dt_rad_avg), **astronomy_nml)
        self._ck(rrtm._lib().isca_b200_moist_set_seasonal(self._h, C.byref(dc)), "two_stream_gray_rad_init")

    def set_ozone(self, o3):
        """the field read from `ozone_file` (do_read_ozone), [lev, lat, lon]; None = no ozone"""
        from . import rrtm
        lib = rrtm._lib()
        if o3 is None:
            self._ck(lib.isca_b200_moist_set_ozone(self._h, None), "set_ozone")
            return
        a = np.ascontiguousarray(o3, dtype=np.float64)
        if a.shape != self.s3:
            raise IscaError("ozone has the wrong shape")
        self._ck(lib.isca_b200_moist_set_ozone(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), "set_ozone")

    def set_time(self, days, seconds):
        from . import rrtm
        self._ck(rrtm._lib().isca_b200_moist_set_time(self._h, int(days), int(seconds)), "set_time")

    def set_t_surf(self, t_surf):
        a = np.ascontiguousarray(t_surf, dtype=np.float64)
        if a.shape != self.s2:
            raise IscaError("t_surf has the wrong shape")
        self._ck(self._lib.isca_b200_moist_set_t_surf(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), "set_t_surf")

    def set_sst(self, sst):
        """mixed_layer_nml do_sc_sst: the prescribed SST [lat, lon] the following steps move t_surf to (the field interpolator_mod
        reads from sst_file for Time_next); None = slab ocean again"""
        if sst is None:
            self._ck(self._lib.isca_b200_moist_set_sst(self._h, None), "set_sst")
            return
        a = np.ascontiguousarray(sst, dtype=np.float64)
        if a.shape != self.s2:
            raise IscaError("sst has the wrong shape")
        self._ck(self._lib.isca_b200_moist_set_sst(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), "set_sst")

    SURFACE_FIELDS = dict(albedo=20, rough_mom=21, rough_heat=22, rough_moist=23, heat_capacity=24, land=25)

    def set_surface(self, name, field):
        """per-column surface properties of the land options ([lat, lon]; `land` is a 0/1 mask); after idealized_moist_phys_init"""
        if name not in self.SURFACE_FIELDS:
            raise IscaError(f"unknown surface field {name}")
        a = np.ascontiguousarray(field, dtype=np.float64)
        if a.shape != self.s2:
            raise IscaError(f"{name} has the wrong shape")
        self._ck(self._lib.isca_b200_moist_set_surface(self._h, self.SURFACE_FIELDS[name], a.ctypes.data_as(C.POINTER(C.c_double))), "set_surface")

    def set_co2(self, carbon_conc):
        """two_stream_gray_rad_nml do_read_co2: the value of co2_file at the Time of the next atmosphere() call (ppmv)"""
        self._ck(self._lib.isca_b200_moist_set_co2(self._h, float(carbon_conc)), "two_stream_gray_rad")

    def set_betts_miller(self, **nml):
        """betts_miller_nml of convection_scheme = 'FULL_BETTS_MILLER' (tau_bm, rhbm, do_simp, do_shallower, do_changeqref, do_envsat,
        buoyancy_kick)"""
        from .physics import betts_miller_config
        cfg = betts_miller_config(**nml)
        self._ck(self._lib.isca_b200_moist_set_betts_miller(self._h, C.byref(cfg)), "betts_miller_init")

    def set_dry_convection(self, tau, gamma):
        """dry_convection_nml (convection_scheme = 'DRY'): relaxation time tau [s], lapse-rate factor gamma"""
        self._ck(self._lib.isca_b200_moist_set_dry_convection(self._h, float(tau), float(gamma)), "dry_convection_init")

    def set_ocean_qflux(self, qflux):
        """mixed_layer_init: the ocean q-flux [lat, lon] in W/m2 (see `qflux`); after idealized_moist_phys_init"""
        a = np.ascontiguousarray(qflux, dtype=np.float64)
        if a.shape != self.s2:
            raise IscaError("ocean_qflux has the wrong shape")
        self._ck(self._lib.isca_b200_moist_set_ocean_qflux(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), "set_ocean_qflux")

    def timing(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self._lib.isca_b200_moist_timing(self._h, C.byref(a), C.byref(b)), "timing")
        return a.value, b.value

    def step_io(self, ozone, outputs):
        """atmosphere(Time) with the host I/O of the step, software-pipelined (isca_b200_moist_step_io): `ozone` a C-contiguous host
        array [lev, lat, lon] or None; `outputs` a list of (kind, id, level, array) with kind 0 = dynamical-core field id of api.F_*,
        1 = physics field name / id of FIELDS_2D / FIELDS_3D.  Returns at once; the arrays are complete after io_sync()."""
        n = len(outputs)
        kinds = (C.c_int * n)(*[int(o[0]) for o in outputs])
        ids = (C.c_int * n)(*[int(FIELDS_2D.get(o[1], FIELDS_3D.get(o[1], -1)) if isinstance(o[1], str) else o[1]) for o in outputs])
        levels = (C.c_int * n)(*[int(o[2]) for o in outputs])
        ptrs = (C.c_void_p * n)(*[o[3].ctypes.data for o in outputs])
        o3 = None if ozone is None else ozone.ctypes.data_as(C.c_void_p)
        self._ck(self._lib.isca_b200_moist_step_io(self._h, o3, n, kinds, ids, levels, ptrs), "step_io")

    def io_sync(self):
        self._ck(self._lib.isca_b200_moist_io_sync(self._h), "io_sync")

    def io_wait(self, age=0):
        """wait for the downloads of the last step_io call (age 0) or of the one before it (age 1); the pipeline keeps running"""
        self._ck(self._lib.isca_b200_moist_io_wait(self._h, int(age)), "io_wait")

    def profile_step(self, n_steps=10):
        """average milliseconds per kernel group over n eager steps: physics kernels ("phys_*") and the dynamical core's groups"""
        ms = (C.c_double * 96)()
        names = C.create_string_buffer(8192)
        n = self._lib.isca_b200_moist_profile_step(self._h, int(n_steps), ms, 96, names, 8192)
        if n < 0:
            self._ck(1, "profile_step")
        return dict(zip(names.value.decode().split(";"), [ms[i] for i in range(n)]))

    def atmosphere_end(self):
        if self._h:
            self.core.h = None
            self._lib.isca_b200_moist_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.atmosphere_end()
        except Exception:
            pass


RESOLUTIONS = {"T21": (64, 32, 21), "T42": (128, 64, 42), "T85": (256, 128, 85), "T170": (512, 256, 170), "T341": (1024, 512, 341)}

# spectral_dynamics_nml of the moist test cases (frierson / MiMA / axisymmetric _test_case.py; the axisymmetric case: surf_res 0.2,
# make_symmetric).  tests/test_reference_python_pins.py holds these dicts against the reference's scripts.
TEST_CASE_DYNAMICS_NML = dict(damping_order=4, water_correction_limit=200.0e2, reference_sea_level_press=1.0e5, valid_range_t=(100.0, 800.0),
                              initial_sphum=2.0e-6, vert_coord_option="uneven_sigma", scale_heights=11.0, exponent=7.0, surf_res=0.5,
                              robert_coeff=0.03)

# exp/test_cases/frierson/frierson_test_case.py: scheme namelists of the grey-radiation aquaplanet
FRIERSON_PHYSICS_NML = dict(atm_abs=0.2,                                                 # two_stream_gray_rad_nml
                            use_virtual_temp=0, surface_flux_do_simple=1, old_dtaudv=1,  # surface_flux_nml
                            diffusivity_do_entrain=0, diffusivity_do_simple=1,           # diffusivity_nml
                            rhbm=0.7, Tmin=160.0, Tmax=350.0)                            # qe_moist_convection_nml
FRIERSON_MOIST_NML = dict(mixed_layer_depth=2.5, albedo_value=0.31)                      # mixed_layer_nml
# the remaining scheme options of frierson_test_case.py (reference_options=True): lscale_cond_nml do_evap, the Rayleigh sponge of
# damping_driver_nml, idealized_moist_phys_nml roughness lengths and do_damping, vert_turb_driver_nml constant_gust / use_tau
FRIERSON_REFERENCE_PHYSICS_NML = dict(FRIERSON_PHYSICS_NML, do_evap=1, trayfric=-0.25, sponge_pbottom=5000.0, do_conserve_energy=1)
FRIERSON_REFERENCE_MOIST_NML = dict(FRIERSON_MOIST_NML, roughness_mom=3.21e-05, roughness_heat=3.21e-05, roughness_moist=3.21e-05,
                                    constant_gust=0.0, use_tau=0, do_damping=1)


def frierson_test_case(res: str, num_levels: int, dt_atmos: float, reference_options: bool = False, **ranks) -> MoistAtmosphere:
    """The Frierson test case (frierson_test_case.py:60-170) at a given resolution: spectral_dynamics_nml with uneven sigma
    levels (scale_heights 11, exponent 7, surf_res 0.5 -- the MiMA/Frierson level distribution of SURVEY section 8d; the shipped script
    reads 25 levels from a file with vert_coord_option = 'input'), sphum as the grid tracer, SIMPLE_BETTS_MILLER convection, slab ocean
    of 2.5 m.  reference_options = True adds the script's remaining scheme options (re-evaporation in lscale_cond, the Rayleigh sponge
    above 50 hPa, roughness lengths 3.21e-5 m, constant_gust = 0, use_tau = .false.: FRIERSON_REFERENCE_*_NML); the default leaves them
    at the schemes' namelist defaults."""
    from .api import make_config
    I, J, M = RESOLUTIONS[res]
    cfg = make_config(lon_max=I, lat_max=J, num_fourier=M, num_spherical=M + 1, num_levels=num_levels, dt_atmos=dt_atmos,
                      num_tracers=1, **TEST_CASE_DYNAMICS_NML)
    phys, mnml = (FRIERSON_REFERENCE_PHYSICS_NML, FRIERSON_REFERENCE_MOIST_NML) if reference_options else (FRIERSON_PHYSICS_NML, FRIERSON_MOIST_NML)
    return MoistAtmosphere(cfg, physics_nml=phys, convection_scheme="SIMPLE_BETTS_MILLER", **ranks, **mnml)


def lat_boundaries(lat_max: int) -> np.ndarray:
    """transforms_mod lat_boundaries_global (transforms.F90:314-323): latb(j+1) = asin(sum of the Gaussian weights up to j - 1),
    south to north, radians"""
    _, w = np.polynomial.legendre.leggauss(lat_max)
    latb = np.empty(lat_max + 1)
    latb[0], latb[-1] = -0.5 * np.pi, 0.5 * np.pi
    latb[1:-1] = np.arcsin(np.clip(np.cumsum(w)[:-1] - 1.0, -1.0, 1.0))
    return latb


def qflux(latb: np.ndarray, num_lon: int, qflux_amp=30.0, qflux_width=16.0) -> np.ndarray:
    """qflux_mod qflux (atmos_param/qflux/qflux.f90:64-83) on an initially zero flux: the Merlis & Schneider tropical ocean heat
    transport, [lat, lon] in W/m2"""
    lat = 0.5 * (latb[1:] + latb[:-1])
    coslat = np.cos(lat)
    lat = lat * 180.0 / np.pi
    f = -qflux_amp * (1 - 2. * lat ** 2 / qflux_width ** 2) * np.exp(-((lat) ** 2 / (qflux_width) ** 2)) / coslat
    return np.repeat(f[:, None], num_lon, 1)


# exp/test_cases/MiMA/MiMA_test_case.py: scheme namelists of the RRTMG aquaplanet (Jucker & Gerber 2017)
MIMA_PHYSICS_NML = dict(use_virtual_temp=0, surface_flux_do_simple=1, old_dtaudv=1,      # surface_flux_nml
                        diffusivity_do_entrain=0, diffusivity_do_simple=1,               # diffusivity_nml
                        rhbm=0.7, Tmin=160.0, Tmax=350.0,                                # qe_moist_convection_nml
                        do_evap=1,                                                       # lscale_cond_nml
                        trayfric=-0.5, sponge_pbottom=50.0, do_conserve_energy=1)        # damping_driver_nml
MIMA_MOIST_NML = dict(mixed_layer_depth=100.0, albedo_value=0.205, roughness_mom=3.21e-05, roughness_heat=3.21e-05,
                      roughness_moist=3.21e-05, constant_gust=0.0, use_tau=0, do_damping=1)
MIMA_RRTM_NML = dict(solr_cnst=1360.0)
MIMA_RRTM_DRIVER_NML = dict(dt_rad=7200)


def mima_test_case(res: str, num_levels: int, dt_atmos: float, ozone=None, **ranks) -> MoistAtmosphere:
    """The MiMA test case (MiMA_test_case.py:50-175): as the Frierson case but RRTMG radiation every 7200 s, Rayleigh sponge,
    use_tau = .false., a 100 m slab with the prescribed initial SST distribution (mixed_layer.F90:347: tconst = 285, delta_T = 40) and
    the tropical q-flux (qflux_amp = 30).  `ozone`: the field of ozone_1990.nc on the model levels ([lev, lat, lon], mass mixing
    ratio) or None.  Returns the model after cold start and idealized_moist_phys_init."""
    from .api import make_config
    I, J, M = RESOLUTIONS[res]
    nranks, rank = ranks.get("nranks", 1), ranks.get("rank", 0)
    cfg = make_config(lon_max=I, lat_max=J, num_fourier=M, num_spherical=M + 1, num_levels=num_levels, dt_atmos=dt_atmos,
                      num_tracers=1, **TEST_CASE_DYNAMICS_NML)
    m = MoistAtmosphere(cfg, physics_nml=MIMA_PHYSICS_NML, convection_scheme="SIMPLE_BETTS_MILLER", **ranks, **MIMA_MOIST_NML)
    dt_rad = MIMA_RRTM_DRIVER_NML["dt_rad"]
    if dt_rad % int(dt_atmos) != 0:
        dt_rad = int(dt_atmos) * max(1, round(dt_rad / dt_atmos))
    m.use_rrtm(MIMA_RRTM_NML, dt_rad=dt_rad)
    if ozone is not None:
        m.set_ozone(ozone)
    m.core.cold_start()
    m.idealized_moist_phys_init()
    from .api import TB_SIN_LAT, TB_WTS_LAT
    Jloc = J // nranks
    sin_lat = m.core.get_table(TB_SIN_LAT)                      # global tables of the core (Gaussian latitudes, south to north)
    wts = m.core.get_table(TB_WTS_LAT)
    latb = np.empty(J + 1)
    latb[0], latb[-1] = -0.5 * np.pi, 0.5 * np.pi
    latb[1:-1] = np.arcsin(np.clip(np.cumsum(wts)[:-1] * (2.0 / wts.sum()) - 1.0, -1.0, 1.0))
    sl = slice(rank * Jloc, (rank + 1) * Jloc)
    m.set_t_surf(np.repeat((285.0 - 40.0 * ((3. * sin_lat[sl] ** 2.) - 1.) / 3.)[:, None], I, 1))
    m.set_ocean_qflux(qflux(latb, I)[sl])
    return m


def axisymmetric_test_case(res: str, num_levels: int, dt_atmos: float, ozone=None, **ranks) -> MoistAtmosphere:
    """The axisymmetric test case (exp/test_cases/axisymmetric/axisymmetric_test_case.py:52-178): the MiMA options with a zonally
    symmetric dynamical core (spectral_dynamics_nml make_symmetric), vertical diffusion in the free atmosphere (diffusivity_nml
    free_atm_diff), RRTMG every 3600 s, a sponge below 150 Pa, surf_res = 0.2 and prescribed SSTs (mixed_layer_nml do_sc_sst): the
    caller hands the SST of sst_file for the time stepped to with set_sst() before each step (or once, for a perpetual field).
    diffusivity_nml: the shipped script's dict literal repeats the key 'diffusivity_nml', so Python keeps only {free_atm_diff: True} and
    do_entrain / do_simple run at their defaults there; this helper keeps the MiMA values (do_entrain = .false., do_simple = .true.)
    that the script's first entry intends, plus free_atm_diff."""
    from .api import make_config
    I, J, M = RESOLUTIONS[res]
    cfg = make_config(lon_max=I, lat_max=J, num_fourier=M, num_spherical=M + 1, num_levels=num_levels, dt_atmos=dt_atmos,
                      num_tracers=1, **dict(TEST_CASE_DYNAMICS_NML, surf_res=0.2, make_symmetric=1))
    phys = dict(MIMA_PHYSICS_NML, sponge_pbottom=150.0, free_atm_diff=1)
    moist_nml = dict(MIMA_MOIST_NML, albedo_value=0.25)
    m = MoistAtmosphere(cfg, physics_nml=phys, convection_scheme="SIMPLE_BETTS_MILLER", **ranks, **moist_nml)
    dt_rad = 3600
    if dt_rad % int(dt_atmos) != 0:
        dt_rad = int(dt_atmos) * max(1, round(dt_rad / dt_atmos))
    m.use_rrtm(dict(), dt_rad=dt_rad)                  # rrtm_radiation_nml of the script: dt_rad only (solr_cnst at its default 1368.22)
    if ozone is not None:
        m.set_ozone(ozone)
    m.core.cold_start()
    m.idealized_moist_phys_init()
    return m
