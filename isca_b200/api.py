"""ctypes binding of the isca_b200 C ABI (include/isca_b200.h) and a host-side mirror of the
reference's Fortran module interface for the hot path.

The reference host language is Fortran; no Fortran compiler exists in this image, so the host
mirror used by the tests and by bench.py is this thin Python layer.  It keeps the reference's
names and argument meaning:

    atmosphere_mod        atmosphere_init / atmosphere / atmosphere_end
                          (atmos_spectral/driver/solo/atmosphere.F90:120,276,356)
    spectral_dynamics_mod spectral_dynamics(...)  (model/spectral_dynamics.F90:780)
    transforms_mod        trans_spherical_to_grid, trans_grid_to_spherical, uv_grid_from_vor_div,
                          vor_div_from_uv_grid     (tools/transforms.F90:134-184)

All compute happens in the CUDA library; there is no CPU fallback: if the library or a CUDA
device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field, fields
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libisca_b200.so")
ABI_VERSION = 1


class IscaConfigStruct(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("lon_max", C.c_int32), ("lat_max", C.c_int32), ("num_fourier", C.c_int32),
        ("num_spherical", C.c_int32), ("num_levels", C.c_int32),
        ("dt_atmos", C.c_double),
        ("damping_order", C.c_int32), ("damping_order_vor", C.c_int32), ("damping_order_div", C.c_int32),
        ("damping_coeff", C.c_double), ("damping_coeff_vor", C.c_double), ("damping_coeff_div", C.c_double),
        ("eddy_sponge_coeff", C.c_double), ("zmu_sponge_coeff", C.c_double), ("zmv_sponge_coeff", C.c_double),
        ("do_mass_correction", C.c_int32), ("do_energy_correction", C.c_int32), ("do_water_correction", C.c_int32),
        ("use_virtual_temperature", C.c_int32), ("use_implicit", C.c_int32), ("make_symmetric", C.c_int32),
        ("robert_coeff", C.c_double), ("raw_filter_coeff", C.c_double), ("alpha_implicit", C.c_double),
        ("vert_coord_option", C.c_int32),
        ("scale_heights", C.c_double), ("surf_res", C.c_double), ("exponent", C.c_double),
        ("p_press", C.c_double), ("p_sigma", C.c_double),
        ("vert_advect_uv", C.c_int32), ("vert_advect_t", C.c_int32),
        ("reference_sea_level_press", C.c_double), ("initial_sphum", C.c_double),
        ("water_correction_limit", C.c_double),
        ("valid_range_t", C.c_double * 2),
        ("initial_temperature", C.c_double),
        ("num_tracers", C.c_int32), ("tracer_robert_coeff", C.c_double),
        ("no_forcing", C.c_int32), ("do_conserve_energy", C.c_int32),
        ("t_zero", C.c_double), ("t_strat", C.c_double), ("delh", C.c_double), ("delv", C.c_double),
        ("eps", C.c_double), ("sigma_b", C.c_double), ("P00", C.c_double), ("ka", C.c_double),
        ("ks", C.c_double), ("kf", C.c_double), ("trflux", C.c_double), ("trsink", C.c_double),
        ("radius", C.c_double), ("omega", C.c_double), ("grav", C.c_double), ("rdgas", C.c_double),
        ("kappa", C.c_double),
        ("pk", C.POINTER(C.c_double)), ("bk", C.POINTER(C.c_double)),
    ]


EXPORTS = [
    "isca_b200_default_config", "isca_b200_create", "isca_b200_destroy", "isca_b200_last_error", "isca_b200_host_table",
    "isca_b200_nccl_unique_id", "isca_b200_cold_start", "isca_b200_set_grid_state",
    "isca_b200_set_spectral_state", "isca_b200_set_vor_div_grid", "isca_b200_set_surf_geopotential",
    "isca_b200_set_time_pointers", "isca_b200_step", "isca_b200_step_dynamics_only",
    "isca_b200_spectral_dynamics", "isca_b200_spectral_dynamics_tracers", "isca_b200_get_field", "isca_b200_get_spectral", "isca_b200_get_scalar",
    "isca_b200_get_table", "isca_b200_get_time_pointers", "isca_b200_spherical_to_grid",
    "isca_b200_grid_to_spherical", "isca_b200_uv_grid_from_vor_div", "isca_b200_vor_div_from_uv_grid",
    "isca_b200_time_transforms", "isca_b200_profile_step", "isca_b200_decomposition",
    "isca_b200_ipc_handles", "isca_b200_set_peer_handles",
    "isca_b200_fft_r2c", "isca_b200_fft_c2r", "isca_b200_legendre_inv", "isca_b200_legendre_fwd",
    "isca_b200_implicit_correction", "isca_b200_diag_accumulate", "isca_b200_diag_fetch",
]

# field / scalar ids (include/isca_b200.h)
F_PS, F_U, F_V, F_T, F_VOR, F_DIV, F_WG_FULL, F_P_FULL, F_P_HALF, F_Z_FULL, F_Z_HALF = range(11)
F_TRACER0 = 16
# derived fields of spectral_diagnostics (spectral_dynamics.F90:1747-1835), formed on the device (include/isca_b200.h)
(F_WSPD, F_UU, F_VV, F_UV, F_V_VOR, F_TT, F_OMEGA_OMEGA, F_OMEGA_T, F_UW, F_VW, F_UT, F_VT, F_UZ, F_VZ, F_OMEGA_Z) = range(32, 47)
F_UTR0, F_VTR0, F_WTR0, F_SLP = 48, 49, 50, 56
S_VOR, S_DIV, S_T, S_LNPS = range(4)
S_DT_VOR, S_DT_DIV, S_DT_T, S_DT_LNPS = 8, 9, 10, 11
LEVEL_CURRENT, LEVEL_PREVIOUS = -1, -2
SC_MEAN_PS, SC_MEAN_ENERGY, SC_T_MIN, SC_T_MAX, SC_STEP_COUNT, SC_KERNEL_LAUNCHES, SC_LAST_STEP_MS = range(7)
TB_SIN_LAT, TB_WTS_LAT, TB_DEG_LAT, TB_DEG_LON, TB_PK, TB_BK = range(6)
# isca_b200_host_table only (packed spectral rows, see include/isca_b200.h)
TB_ROW_M, TB_ROW_N, TB_LEGENDRE, TB_EIGEN_LAPLACIAN, TB_DAMPING, TB_REF_T, TB_IMPLICIT_H, TB_DIV_MAT, TB_WAVE_MATRIX = range(16, 25)


def host_table(config, table_id) -> np.ndarray:
    """The host-side table the library builds at create time for this configuration (host_tables.cpp), flat.  Needs no GPU."""
    lib = load_library()
    lib.isca_b200_host_table.argtypes = [C.POINTER(IscaConfigStruct), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    n = C.c_int(0)
    if lib.isca_b200_host_table(C.byref(config), table_id, None, 0, C.byref(n)) != 0:
        raise IscaError("host_table: " + lib.isca_b200_last_error(None).decode())
    out = np.empty(n.value)
    if lib.isca_b200_host_table(C.byref(config), table_id, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, None) != 0:
        raise IscaError("host_table: " + lib.isca_b200_last_error(None).decode())
    return out

_VERT_COORD = {"even_sigma": 0, "uneven_sigma": 1, "input": 2, "hybrid": 3}
_VERT_ADV = {"second_centered": 0}

_lib = None


def load_library() -> C.CDLL:
    """Load libisca_b200.so; raises (no fallback) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ISCA_B200_LIB", LIB_PATH)     # development only: a differently tuned build of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m isca_b200.build` "
                           "(isca_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.isca_b200_default_config.argtypes = [C.POINTER(IscaConfigStruct)]
    lib.isca_b200_default_config.restype = None
    lib.isca_b200_create.argtypes = [C.POINTER(IscaConfigStruct), C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.isca_b200_destroy.argtypes = [vp]
    lib.isca_b200_last_error.argtypes = [vp]
    lib.isca_b200_last_error.restype = C.c_char_p
    lib.isca_b200_nccl_unique_id.argtypes = [vp]
    lib.isca_b200_cold_start.argtypes = [vp]
    lib.isca_b200_set_grid_state.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.isca_b200_set_spectral_state.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.isca_b200_set_vor_div_grid.argtypes = [vp, vp, vp]
    lib.isca_b200_set_surf_geopotential.argtypes = [vp, vp]
    lib.isca_b200_set_time_pointers.argtypes = [vp, C.c_int, C.c_int]
    lib.isca_b200_get_time_pointers.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.isca_b200_step.argtypes = [vp, C.c_int]
    lib.isca_b200_step_dynamics_only.argtypes = [vp, C.c_int]
    lib.isca_b200_spectral_dynamics.argtypes = [vp] + [vp] * 10
    lib.isca_b200_spectral_dynamics_tracers.argtypes = [vp] + [vp] * 12
    for f in (lib.isca_b200_fft_r2c, lib.isca_b200_fft_c2r, lib.isca_b200_legendre_inv):
        f.argtypes = [vp, vp, vp, C.c_int]
    lib.isca_b200_legendre_fwd.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.isca_b200_implicit_correction.argtypes = [vp] + [vp] * 9 + [C.c_double]
    lib.isca_b200_diag_accumulate.argtypes = [vp, C.c_int]
    lib.isca_b200_diag_fetch.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    lib.isca_b200_get_field.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.isca_b200_get_spectral.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.isca_b200_get_scalar.argtypes = [vp, C.c_int, dp]
    lib.isca_b200_get_table.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.isca_b200_spherical_to_grid.argtypes = [vp, vp, vp, C.c_int]
    lib.isca_b200_grid_to_spherical.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.isca_b200_uv_grid_from_vor_div.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    lib.isca_b200_vor_div_from_uv_grid.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    lib.isca_b200_time_transforms.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.isca_b200_profile_step.argtypes = [vp, C.c_int, dp, C.c_int, C.c_char_p, C.c_int]
    lib.isca_b200_ipc_handles.argtypes = [vp, vp]
    lib.isca_b200_set_peer_handles.argtypes = [vp, vp]
    ip = C.POINTER(C.c_int)
    lib.isca_b200_decomposition.argtypes = [C.POINTER(IscaConfigStruct), C.c_int, C.c_int, ip, ip, ip, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("isca_b200_default_config", "isca_b200_last_error"):
            fn.restype = C.c_int
    _lib = lib
    return lib


class IscaError(RuntimeError):
    """error_mesg(..., FATAL) analogue."""


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 creates it, the host runtime broadcasts it to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.isca_b200_nccl_unique_id(buf) != 0:
        raise IscaError("nccl_unique_id: " + lib.isca_b200_last_error(None).decode())
    return buf.raw


def decomposition(config, rank, nranks):
    """The library's (rank, nranks) decomposition: dict(lat_start, lat_count, m_list, owner, pos). CPU only."""
    lib = load_library()
    M = config.num_fourier
    j0, jc, nm = C.c_int(), C.c_int(), C.c_int()
    m_list = np.zeros(M + 1, dtype=np.int32)
    owner = np.zeros(M + 1, dtype=np.int32)
    pos = np.zeros(M + 1, dtype=np.int32)
    rc = lib.isca_b200_decomposition(C.byref(config), rank, nranks, C.byref(j0), C.byref(jc), C.byref(nm),
                                     m_list.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p),
                                     pos.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise IscaError("decomposition: " + lib.isca_b200_last_error(None).decode())
    return dict(lat_start=j0.value, lat_count=jc.value, m_list=m_list[: nm.value].copy(), owner=owner, pos=pos)


def make_config(**kw) -> IscaConfigStruct:
    """IscaConfig with the reference namelist defaults, overridden by keyword (namelist) values.
    String-valued namelist options (vert_coord_option, vert_advect_uv/t) are accepted by name."""
    lib = load_library()
    c = IscaConfigStruct()
    lib.isca_b200_default_config(C.byref(c))
    names = {f[0] for f in IscaConfigStruct._fields_}
    keep = []
    for k, v in kw.items():
        if k not in names:
            raise IscaError(f"unknown namelist variable {k!r}")
        if k == "vert_coord_option" and isinstance(v, str):
            if v not in _VERT_COORD:
                raise IscaError(f'"{v}" is not a valid value for vert_coord_option')
            v = _VERT_COORD[v]
        if k in ("vert_advect_uv", "vert_advect_t") and isinstance(v, str):
            if v.lower() not in _VERT_ADV:
                raise IscaError(f'"{v}" is not a supported value for {k}')
            v = _VERT_ADV[v.lower()]
        if k == "valid_range_t":
            c.valid_range_t[0], c.valid_range_t[1] = float(v[0]), float(v[1])
            continue
        if k in ("pk", "bk"):
            arr = np.ascontiguousarray(v, dtype=np.float64)
            keep.append(arr)
            setattr(c, k, arr.ctypes.data_as(C.POINTER(C.c_double)))
            continue
        if isinstance(v, bool):
            v = int(v)
        setattr(c, k, v)
    c._keepalive = keep
    return c


def config_from_namelist_object(cfg) -> IscaConfigStruct:
    """Build an IscaConfig from any object exposing the namelist variables as attributes
    (e.g. a dataclass); attributes that are not part of IscaConfig are ignored."""
    names = {f[0] for f in IscaConfigStruct._fields_} - {"abi_version", "pk", "bk"}
    kw = {}
    for n in names:
        if hasattr(cfg, n):
            kw[n] = getattr(cfg, n)
    return make_config(**kw)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _in(a, shape=None, dtype=np.float64):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise IscaError(f"array has shape {a.shape}, expected {shape}")
    return a


class Atmosphere:
    """atmosphere_mod mirror.  State lives on the GPU; fields are fetched lazily."""

    def __init__(self, config: IscaConfigStruct, rank: int = 0, nranks: int = 1, nccl_unique_id: bytes | None = None,
                 _adopt_handle=None):
        self.lib = load_library()
        self.cfg = config
        self.h = C.c_void_p()
        self._owns = _adopt_handle is None
        if _adopt_handle is not None:              # a core owned by another object (the moist-model driver)
            self.h = C.c_void_p(_adopt_handle)
            self.I, self.J, self.K = config.lon_max, config.lat_max, config.num_levels
            self.M, self.N = config.num_fourier, config.num_spherical
            self.Jloc = self.J
            return
        uid = None
        if nccl_unique_id is not None:
            self._uid_buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            uid = C.cast(self._uid_buf, C.c_void_p)
        rc = self.lib.isca_b200_create(C.byref(config), rank, nranks, uid, C.byref(self.h))
        if rc != 0:
            msg = self.lib.isca_b200_last_error(None)
            self.h = None
            raise IscaError("atmosphere_init: " + (msg.decode() if msg else f"error {rc}"))
        self.I, self.J, self.K = config.lon_max, config.lat_max, config.num_levels
        self.M, self.N = config.num_fourier, config.num_spherical
        self.Jloc = self.J // nranks

    # ---- error handling ------------------------------------------------------------------
    def _ck(self, rc, where):
        if rc != 0:
            msg = self.lib.isca_b200_last_error(self.h)
            raise IscaError(f"{where}: " + (msg.decode() if msg else f"error {rc}"))

    # ---- peer-memory transpose (multi-rank) ------------------------------------------------
    def ipc_handles(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.isca_b200_ipc_handles(self.h, buf), "ipc_handles")
        return buf.raw

    def set_peer_handles(self, all_handles):
        """all_handles: list of the nranks 128-byte blobs returned by ipc_handles(), in rank order."""
        blob = b"".join(bytes(x) for x in all_handles)
        self._blob = C.create_string_buffer(blob, len(blob))
        self._ck(self.lib.isca_b200_set_peer_handles(self.h, self._blob), "set_peer_handles")

    # ---- atmosphere_mod ------------------------------------------------------------------
    @classmethod
    def atmosphere_init(cls, config, cold_start=True, **kw):
        a = cls(config, **kw)
        if cold_start:
            a.cold_start()
        return a

    def cold_start(self):
        self._ck(self.lib.isca_b200_cold_start(self.h), "spectral_init_cond")

    def atmosphere(self, n_steps: int = 1):
        """atmosphere(Time), n_steps times."""
        self._ck(self.lib.isca_b200_step(self.h, n_steps), "atmosphere")

    def atmosphere_dynamics_only(self, n_steps: int = 1):
        self._ck(self.lib.isca_b200_step_dynamics_only(self.h, n_steps), "atmosphere")

    def atmosphere_end(self):
        if self.h is not None:
            if self._owns:
                self.lib.isca_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.atmosphere_end()
        except Exception:
            pass

    # ---- spectral_dynamics_mod -------------------------------------------------------------
    def spectral_dynamics(self, dt_ug=None, dt_vg=None, dt_tg=None, dt_tracers=None, want=("psg", "ug", "vg", "tg")):
        """spectral_dynamics(Time, psg_final, ug_final, vg_final, tg_final, tracer_attributes, grid_tracers_final, ..., dt_psg,
        dt_ug, dt_vg, dt_tg, dt_tracers, wg_full, p_full, ...) with host arrays in and out."""
        s3 = (self.K, self.Jloc, self.I)
        a = [_in(x, s3) for x in (dt_ug, dt_vg, dt_tg, dt_tracers)]
        out = {}
        for name, shape in (("psg", (self.Jloc, self.I)), ("ug", s3), ("vg", s3), ("tg", s3), ("grid_tracers", s3), ("wg_full", s3),
                            ("p_full", s3)):
            out[name] = np.empty(shape) if name in want else None
        self._ck(self.lib.isca_b200_spectral_dynamics_tracers(self.h, None, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                                              _ptr(out["psg"]), _ptr(out["ug"]), _ptr(out["vg"]), _ptr(out["tg"]),
                                                              _ptr(out["grid_tracers"]), _ptr(out["wg_full"]), _ptr(out["p_full"])),
                 "spectral_dynamics")
        return {k: v for k, v in out.items() if v is not None}

    # ---- stage-level transforms, implicit_correction, device-side diagnostics ------------------
    def trans_grid_to_fourier(self, grid):
        g = _in(grid); nlev = g.shape[0]
        out = np.empty((nlev, self.J, self.M + 1), dtype=np.complex128)
        self._ck(self.lib.isca_b200_fft_r2c(self.h, _ptr(g), _ptr(out), nlev), "trans_grid_to_fourier")
        return out

    def trans_fourier_to_grid(self, fourier):
        f = _in(fourier, None, np.complex128); nlev = f.shape[0]
        out = np.empty((nlev, self.J, self.I))
        self._ck(self.lib.isca_b200_fft_c2r(self.h, _ptr(f), _ptr(out), nlev), "trans_fourier_to_grid")
        return out

    def trans_spherical_to_fourier(self, spec):
        sp = _in(spec, None, np.complex128); nlev = sp.shape[0]
        out = np.empty((nlev, self.J, self.M + 1), dtype=np.complex128)
        self._ck(self.lib.isca_b200_legendre_inv(self.h, _ptr(sp), _ptr(out), nlev), "trans_spherical_to_fourier")
        return out

    def trans_fourier_to_spherical(self, fourier, do_truncation=True):
        f = _in(fourier, None, np.complex128); nlev = f.shape[0]
        out = np.empty((nlev, self.N + 1, self.M + 1), dtype=np.complex128)
        self._ck(self.lib.isca_b200_legendre_fwd(self.h, _ptr(f), _ptr(out), nlev, int(do_truncation)), "trans_fourier_to_spherical")
        return out

    def implicit_correction(self, dt_divs, dt_ts, dt_ln_ps, divs, ts, ln_ps, delta_t, previous=0, current=1):
        """implicit_correction(dt_divs, dt_ts, dt_ln_ps, divs, ts, ln_ps, delta_t, previous, current): divs/ts [2, K, N+1, M+1],
        ln_ps [2, N+1, M+1] hold the two time levels; returns the corrected (dt_divs, dt_ts, dt_ln_ps)."""
        s3, s2 = (self.K, self.N + 1, self.M + 1), (self.N + 1, self.M + 1)
        o = [np.array(_in(x, s, np.complex128), copy=True) for x, s in ((dt_divs, s3), (dt_ts, s3), (dt_ln_ps, s2))]
        lv = [_in(x[i], s, np.complex128) for x, s in ((divs, s3), (ts, s3), (ln_ps, s2)) for i in (previous, current)]
        self._ck(self.lib.isca_b200_implicit_correction(self.h, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), *[_ptr(x) for x in lv],
                                                        float(delta_t)), "implicit_correction")
        return tuple(o)

    def diag_accumulate(self, field_id):
        self._ck(self.lib.isca_b200_diag_accumulate(self.h, field_id), "diag_accumulate")

    def diag_fetch(self, field_id, reset=True):
        """time mean of the accumulated samples of a grid field and their number"""
        n2 = (self.Jloc, self.I)
        shape = n2 if field_id in (F_PS, F_SLP) else ((self.K + 1,) + n2 if field_id in (F_P_HALF, F_Z_HALF) else (self.K,) + n2)
        out = np.empty(shape); cnt = C.c_int(0)
        self._ck(self.lib.isca_b200_diag_fetch(self.h, field_id, _ptr(out), int(reset), C.byref(cnt)), "diag_fetch")
        return out, cnt.value

    # ---- state I/O (restart path / diag mirrors) ---------------------------------------------
    def set_grid_state(self, slot, ug=None, vg=None, tg=None, psg=None, tracers=None):
        s3 = (self.K, self.Jloc, self.I)
        a = [_in(ug, s3), _in(vg, s3), _in(tg, s3), _in(psg, (self.Jloc, self.I)), _in(tracers, s3)]
        self._ck(self.lib.isca_b200_set_grid_state(self.h, slot, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4])),
                 "set_grid_state")

    def set_spectral_state(self, slot, vors=None, divs=None, ts=None, ln_ps=None):
        s3 = (self.K, self.N + 1, self.M + 1)
        a = [_in(vors, s3, np.complex128), _in(divs, s3, np.complex128), _in(ts, s3, np.complex128),
             _in(ln_ps, (self.N + 1, self.M + 1), np.complex128)]
        self._ck(self.lib.isca_b200_set_spectral_state(self.h, slot, *[_ptr(x) for x in a]), "set_spectral_state")

    def set_vor_div_grid(self, vorg, divg):
        s3 = (self.K, self.Jloc, self.I)
        a, b = _in(vorg, s3), _in(divg, s3)
        self._ck(self.lib.isca_b200_set_vor_div_grid(self.h, _ptr(a), _ptr(b)), "set_vor_div_grid")

    def set_surf_geopotential(self, sg):
        a = _in(sg, (self.Jloc, self.I))
        self._ck(self.lib.isca_b200_set_surf_geopotential(self.h, _ptr(a)), "set_surf_geopotential")

    def set_time_pointers(self, previous, current):
        self._ck(self.lib.isca_b200_set_time_pointers(self.h, previous, current), "set_time_pointers")

    def get_time_pointers(self):
        p, c = C.c_int(), C.c_int()
        self._ck(self.lib.isca_b200_get_time_pointers(self.h, C.byref(p), C.byref(c)), "get_time_pointers")
        return p.value, c.value

    def get_field(self, field_id, level=LEVEL_CURRENT, out=None):
        if field_id in (F_PS, F_SLP):
            shape = (self.Jloc, self.I)
        elif field_id in (F_P_HALF, F_Z_HALF):
            shape = (self.K + 1, self.Jloc, self.I)
        else:
            shape = (self.K, self.Jloc, self.I)
        if out is None:
            out = np.empty(shape)
        self._ck(self.lib.isca_b200_get_field(self.h, field_id, level, _ptr(out)), "get_field")
        return out

    def get_spectral(self, field_id, level=LEVEL_CURRENT):
        shape = (self.N + 1, self.M + 1) if field_id in (S_LNPS, S_DT_LNPS) else (self.K, self.N + 1, self.M + 1)
        out = np.empty(shape, dtype=np.complex128)
        self._ck(self.lib.isca_b200_get_spectral(self.h, field_id, level, _ptr(out)), "get_spectral")
        return out

    def get_scalar(self, scalar_id):
        v = C.c_double()
        self._ck(self.lib.isca_b200_get_scalar(self.h, scalar_id, C.byref(v)), "get_scalar")
        return v.value

    def enable_tendency_capture(self):
        self.get_scalar(100)

    def get_table(self, table_id):
        n = {TB_SIN_LAT: self.J, TB_WTS_LAT: self.J, TB_DEG_LAT: self.J, TB_DEG_LON: self.I,
             TB_PK: self.K + 1, TB_BK: self.K + 1}[table_id]
        out = np.empty(n)
        self._ck(self.lib.isca_b200_get_table(self.h, table_id, _ptr(out), n), "get_table")
        return out

    def state(self):
        """Snapshot of the prognostic state under the reference variable names."""
        return dict(
            vors=self.get_spectral(S_VOR), divs=self.get_spectral(S_DIV), ts=self.get_spectral(S_T),
            ln_ps=self.get_spectral(S_LNPS),
            vors_prev=self.get_spectral(S_VOR, LEVEL_PREVIOUS), divs_prev=self.get_spectral(S_DIV, LEVEL_PREVIOUS),
            ts_prev=self.get_spectral(S_T, LEVEL_PREVIOUS), ln_ps_prev=self.get_spectral(S_LNPS, LEVEL_PREVIOUS),
            ug=self.get_field(F_U), vg=self.get_field(F_V), tg=self.get_field(F_T), psg=self.get_field(F_PS),
            vorg=self.get_field(F_VOR), divg=self.get_field(F_DIV), wg_full=self.get_field(F_WG_FULL),
            p_full=self.get_field(F_P_FULL), z_full=self.get_field(F_Z_FULL))

    # ---- transforms_mod --------------------------------------------------------------------
    def trans_spherical_to_grid(self, spec):
        spec = np.ascontiguousarray(spec, dtype=np.complex128)
        two_d = spec.ndim == 2
        s = spec[None] if two_d else spec
        nlev = s.shape[0]
        s = _in(s, (nlev, self.N + 1, self.M + 1), np.complex128)
        grid = np.empty((nlev, self.Jloc, self.I))
        self._ck(self.lib.isca_b200_spherical_to_grid(self.h, _ptr(s), _ptr(grid), nlev), "trans_spherical_to_grid")
        return grid[0] if two_d else grid

    def trans_grid_to_spherical(self, grid, do_truncation=True):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        two_d = grid.ndim == 2
        gq = grid[None] if two_d else grid
        nlev = gq.shape[0]
        gq = _in(gq, (nlev, self.Jloc, self.I))
        spec = np.empty((nlev, self.N + 1, self.M + 1), dtype=np.complex128)
        self._ck(self.lib.isca_b200_grid_to_spherical(self.h, _ptr(gq), _ptr(spec), nlev, int(do_truncation)),
                 "trans_grid_to_spherical")
        return spec[0] if two_d else spec

    def uv_grid_from_vor_div(self, vors, divs):
        vors = np.ascontiguousarray(vors, dtype=np.complex128)
        divs = np.ascontiguousarray(divs, dtype=np.complex128)
        nlev = vors.shape[0]
        ug = np.empty((nlev, self.Jloc, self.I))
        vg = np.empty_like(ug)
        self._ck(self.lib.isca_b200_uv_grid_from_vor_div(self.h, _ptr(vors), _ptr(divs), _ptr(ug), _ptr(vg), nlev),
                 "uv_grid_from_vor_div")
        return ug, vg

    def vor_div_from_uv_grid(self, ug, vg):
        ug = np.ascontiguousarray(ug, dtype=np.float64)
        vg = np.ascontiguousarray(vg, dtype=np.float64)
        nlev = ug.shape[0]
        vors = np.empty((nlev, self.N + 1, self.M + 1), dtype=np.complex128)
        divs = np.empty_like(vors)
        self._ck(self.lib.isca_b200_vor_div_from_uv_grid(self.h, _ptr(ug), _ptr(vg), _ptr(vors), _ptr(divs), nlev),
                 "vor_div_from_uv_grid")
        return vors, divs

    def spectral_dynamics_into(self, tend, outs):
        """spectral_dynamics with caller-owned (e.g. pinned) host arrays: tend = [dt_ug, dt_vg, dt_tg(, dt_tracers)],
        outs = dict(psg=, ug=, vg=, tg=(, grid_tracers=)); no allocation."""
        self._ck(self.lib.isca_b200_spectral_dynamics_tracers(self.h, None, _ptr(tend[0]), _ptr(tend[1]), _ptr(tend[2]),
                                                              _ptr(tend[3]) if len(tend) > 3 else None,
                                                              _ptr(outs.get("psg")), _ptr(outs.get("ug")), _ptr(outs.get("vg")),
                                                              _ptr(outs.get("tg")), _ptr(outs.get("grid_tracers")),
                                                              _ptr(outs.get("wg_full")), _ptr(outs.get("p_full"))),
                 "spectral_dynamics")

    def profile_step(self, n_steps=10):
        """Average milliseconds per kernel group over n eager steps (CUDA events on the library's stream)."""
        ms = (C.c_double * 64)()
        names = C.create_string_buffer(4096)
        n = self.lib.isca_b200_profile_step(self.h, n_steps, ms, 64, names, 4096)
        if n < 0:
            self._ck(1, "profile_step")
        return dict(zip(names.value.decode().split(";"), [ms[i] for i in range(n)]))

    def time_transforms(self, nlev, reps=5):
        ms = (C.c_double * 4)()
        self._ck(self.lib.isca_b200_time_transforms(self.h, nlev, reps, ms), "time_transforms")
        return dict(legendre_inv=ms[0], fft_inv=ms[1], fft_fwd=ms[2], legendre_fwd=ms[3])
