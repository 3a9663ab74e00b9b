"""CPU restatement of the Isca time step in C++/OpenMP (TEST INFRASTRUCTURE / CPU BASELINE ONLY; see oracle/cstep/isca_cstep.cpp).

`CStep(cfg)` takes the same `Config` as `oracle.isca_oracle.SpectralCore`, builds the init-time tables with the NumPy oracle
(Gaussian grid, Legendre functions, coefficient tables, vertical coordinate, semi-implicit matrices for dt and 2 dt, damping
coefficients, finite-volume grid metrics) and hands them to the compiled step.  The state interface mirrors SpectralCore
(`cold_start`, `step`, `state`, the same array attributes), so the two can be compared step by step (tests/test_cstep.py) and
`bench.py --impl reference` can time the compiled step on all host cores.

Build: `python -m oracle.cstep` (g++ -O2 -fopenmp, as the reference's mkmf template builds with -O2) -> oracle/_build/libisca_cstep.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cstep", "isca_cstep.cpp")
OUT = os.path.join(HERE, "_build", "libisca_cstep.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(SRC) > os.path.getmtime(OUT):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC])
    return OUT


class _Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("I", "J", "K", "M", "N", "num_tracers", "do_mass", "do_energy", "do_water", "use_implicit",
                                       "no_forcing", "do_conserve_energy")] + \
               [(n, C.c_double) for n in ("dt_atmos", "robert_coeff", "raw_filter_coeff", "tracer_robert_coeff", "radius", "grav", "rdgas",
                                          "kappa", "cp_air", "water_correction_limit", "valid_t_lo", "valid_t_hi", "ref_ps", "t_zero",
                                          "t_strat", "delh", "delv", "eps", "sigma_b", "P00", "tka", "tks", "vkf", "trflux", "trsink",
                                          "fv_dx", "alpha_implicit")]


_TABLES = ("legendre", "legendre_wts", "cosm_lat", "wts_lat", "rad_lat", "coriolis", "triangle_mask", "eigen", "coef_uvm", "coef_uvc",
           "coef_uvp", "coef_alpm", "coef_alpp", "coef_dym", "coef_dx", "coef_dyp", "pk", "bk", "damping", "damping_vor", "damping_div",
           "ref_t", "ref_ln_p_half", "ref_ln_p_full", "h", "wave_dt", "wave_2dt", "fv_c", "fv_cc", "fv_dy", "fv_dyy", "fv_dy_plus",
           "fv_dy_minus")


class _Tables(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in _TABLES]


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CStep:
    def __init__(self, cfg):
        from . import isca_oracle as O
        from .fv_advection import FVGrid
        self.cfg = cfg
        lib = C.CDLL(build())
        lib.cstep_create.restype = C.c_void_p
        lib.cstep_error.restype = C.c_char_p
        self.lib = lib
        if cfg.use_virtual_temperature or cfg.eddy_sponge_coeff or cfg.zmu_sponge_coeff or cfg.zmv_sponge_coeff:
            raise ValueError("cstep: virtual temperature / sponges are not restated")
        if cfg.vert_advect_uv != "second_centered" or cfg.vert_advect_t != "second_centered" or cfg.num_tracers > 1:
            raise ValueError("cstep: only second_centered u, v, T advection and at most the sphum grid tracer")
        tb = O.Tables(cfg)
        self.tb = tb
        pk, bk = O.compute_vert_coord(cfg)
        pg = O.PressGeopot(cfg, pk, bk)
        damp = O.SpectralDamping(cfg, tb)
        hs = O.HSForcing(cfg, tb)
        fv = FVGrid(cfg, tb)
        K = cfg.num_levels
        wave = []
        if cfg.use_implicit:
            imp = O.Implicit(cfg, tb, pk, bk, pg)
            for dt in (cfg.dt_atmos, 2 * cfg.dt_atmos):
                imp.xi = dt * imp.alpha
                imp.build_wave_matrices()
                wave.append(np.ascontiguousarray(imp.wave_matrix))
            ref = dict(ref_t=imp.ref_t, ref_ln_p_half=imp.ref_ln_p_half, ref_ln_p_full=imp.ref_ln_p_full, h=imp.h)
        else:
            wave = [np.zeros((cfg.num_spherical, K, K))] * 2
            ref = dict(ref_t=np.zeros(K), ref_ln_p_half=np.zeros(K + 1), ref_ln_p_full=np.zeros(K), h=np.zeros(K))
        t = dict(legendre=tb.legendre, legendre_wts=tb.legendre_wts, cosm_lat=tb.cosm_lat, wts_lat=tb.wts_lat, rad_lat=tb.rad_lat,
                 coriolis=tb.coriolis, triangle_mask=tb.triangle_mask, eigen=tb.eigen_laplacian, coef_uvm=tb.coef_uvm, coef_uvc=tb.coef_uvc,
                 coef_uvp=tb.coef_uvp, coef_alpm=tb.coef_alpm, coef_alpp=tb.coef_alpp, coef_dym=tb.coef_dym, coef_dx=tb.coef_dx,
                 coef_dyp=tb.coef_dyp, pk=pk, bk=bk, damping=damp.damping, damping_vor=damp.damping_vor, damping_div=damp.damping_div,
                 wave_dt=wave[0], wave_2dt=wave[1], fv_c=fv.c, fv_cc=fv.cc, fv_dy=fv.dy, fv_dyy=fv.dyy, fv_dy_plus=fv.dy_plus,
                 fv_dy_minus=fv.dy_minus, **ref)
        self._keep = {k: np.ascontiguousarray(np.nan_to_num(np.asarray(v, dtype=np.float64))) for k, v in t.items()}
        T = _Tables(**{k: _p(v) for k, v in self._keep.items()})
        P = _Params(I=cfg.lon_max, J=cfg.lat_max, K=K, M=cfg.num_fourier, N=cfg.num_spherical, num_tracers=cfg.num_tracers,
                    do_mass=int(cfg.do_mass_correction), do_energy=int(cfg.do_energy_correction), do_water=int(cfg.do_water_correction),
                    use_implicit=int(cfg.use_implicit), no_forcing=int(cfg.no_forcing), do_conserve_energy=int(cfg.do_conserve_energy),
                    dt_atmos=cfg.dt_atmos, robert_coeff=cfg.robert_coeff, raw_filter_coeff=cfg.raw_filter_coeff,
                    tracer_robert_coeff=cfg.tracer_robert_coeff, radius=cfg.radius, grav=cfg.grav, rdgas=cfg.rdgas, kappa=cfg.kappa,
                    cp_air=cfg.cp_air, water_correction_limit=cfg.water_correction_limit, valid_t_lo=cfg.valid_range_t[0],
                    valid_t_hi=cfg.valid_range_t[1], ref_ps=cfg.reference_sea_level_press, t_zero=cfg.t_zero, t_strat=cfg.t_strat,
                    delh=cfg.delh, delv=cfg.delv, eps=cfg.eps, sigma_b=cfg.sigma_b, P00=cfg.P00, tka=hs.tka, tks=hs.tks, vkf=hs.vkf,
                    trflux=cfg.trflux, trsink=hs.trsink, fv_dx=fv.dx, alpha_implicit=cfg.alpha_implicit)
        self.h = lib.cstep_create(C.byref(P), C.byref(T))
        if not self.h:
            raise ValueError("cstep_create: unsupported sizes (lon_max must be a power of two)")
        J, I, M, N = cfg.lat_max, cfg.lon_max, cfg.num_fourier, cfg.num_spherical
        z3 = lambda: np.zeros((2, K, N + 1, M + 1), dtype=np.complex128)
        self.vors, self.divs, self.ts = z3(), z3(), z3()
        self.ln_ps = np.zeros((2, N + 1, M + 1), dtype=np.complex128)
        g3 = lambda: np.zeros((2, K, J, I))
        self.ug, self.vg, self.tg = g3(), g3(), g3()
        self.psg = np.zeros((2, J, I))
        self.vorg, self.divg, self.wg_full = np.zeros((K, J, I)), np.zeros((K, J, I)), np.zeros((K, J, I))
        self.grid_tracers = np.zeros((2, max(cfg.num_tracers, 1), K, J, I))
        self.previous = self.current = 0
        self.threads = lib.cstep_threads()

    def _v(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def push(self):
        """host arrays -> compiled core (both time levels)"""
        q = np.ascontiguousarray(self.grid_tracers[:, 0])
        self.lib.cstep_set_state(C.c_void_p(self.h), self._v(self.vors), self._v(self.divs), self._v(self.ts), self._v(self.ln_ps), self._v(self.ug),
                                 self._v(self.vg), self._v(self.tg), self._v(self.psg), self._v(self.vorg), self._v(self.divg), self._v(q),
                                 int(self.previous), int(self.current))

    def pull(self):
        q = np.zeros_like(self.grid_tracers[:, 0])
        pr, cu = C.c_int(0), C.c_int(0)
        self.lib.cstep_get_state(C.c_void_p(self.h), self._v(self.vors), self._v(self.divs), self._v(self.ts), self._v(self.ln_ps), self._v(self.ug),
                                 self._v(self.vg), self._v(self.tg), self._v(self.psg), self._v(self.vorg), self._v(self.divg), self._v(q),
                                 self._v(self.wg_full), C.byref(pr), C.byref(cu))
        self.grid_tracers[:, 0] = q
        self.previous, self.current = pr.value, cu.value

    def load_from(self, core):
        """copy the state of an oracle.isca_oracle.SpectralCore (or any object with the same attributes)"""
        for k in ("vors", "divs", "ts", "ln_ps", "ug", "vg", "tg", "psg", "vorg", "divg"):
            getattr(self, k)[...] = getattr(core, k)
        if self.cfg.num_tracers:
            self.grid_tracers[:, 0] = core.grid_tracers[:, 0]
        self.previous, self.current = core.previous, core.current
        self.push()

    def cold_start(self):
        from .isca_oracle import SpectralCore
        c = SpectralCore(self.cfg)
        c.cold_start()
        self.load_from(c)

    def step(self, n: int = 1):
        if self.lib.cstep_step(C.c_void_p(self.h), int(n)):
            raise FloatingPointError(self.lib.cstep_error(C.c_void_p(self.h)).decode())

    def state(self):
        self.pull()
        c, p = self.current, self.previous
        return dict(vors=self.vors[c].copy(), divs=self.divs[c].copy(), ts=self.ts[c].copy(), ln_ps=self.ln_ps[c].copy(),
                    vors_prev=self.vors[p].copy(), divs_prev=self.divs[p].copy(), ts_prev=self.ts[p].copy(), ln_ps_prev=self.ln_ps[p].copy(),
                    ug=self.ug[c].copy(), vg=self.vg[c].copy(), tg=self.tg[c].copy(), psg=self.psg[c].copy(), vorg=self.vorg.copy(),
                    divg=self.divg.copy(), wg_full=self.wg_full.copy(), q=self.grid_tracers[c, 0].copy(), q_prev=self.grid_tracers[p, 0].copy())

    def spherical_to_grid(self, spec):
        spec = np.ascontiguousarray(spec, dtype=np.complex128)
        nlev = spec.shape[0]
        g = np.zeros((nlev, self.cfg.lat_max, self.cfg.lon_max))
        self.lib.cstep_spherical_to_grid(C.c_void_p(self.h), self._v(spec), self._v(g), nlev)
        return g

    def grid_to_spherical(self, grid, do_truncation=True):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        nlev = grid.shape[0]
        s = np.zeros((nlev, self.cfg.num_spherical + 1, self.cfg.num_fourier + 1), dtype=np.complex128)
        self.lib.cstep_grid_to_spherical(C.c_void_p(self.h), self._v(grid), self._v(s), nlev, int(do_truncation))
        return s

    def close(self):
        if self.h:
            self.lib.cstep_destroy(C.c_void_p(self.h))
            self.h = None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
