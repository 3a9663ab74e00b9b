#!/usr/bin/env python
"""bench.py -- model-days/sec of the Isca time step (column physics + spectral dynamical core) on B200.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W     # CPU arm (the reference's algorithm on the host cores)

Headline workload (BASELINE.json: "model-days/sec T170L40", configs[3]): the MiMA test case (exp/test_cases/MiMA/MiMA_test_case.py)
at T170 L40, dt = 150 s -- idealized_moist_phys with RRTMG radiation every 7200 s (48 steps), SIMPLE_BETTS_MILLER convection,
large-scale condensation, Monin-Obukhov surface fluxes, K-profile diffusivity, implicit vertical diffusion, slab ocean, Rayleigh
sponge -- followed by spectral_dynamics with the sphum grid tracer.  A "step" is one `atmosphere(Time)` call.  `--workload hs` selects
the Held-Suarez dry core instead.  The other BASELINE configurations are measured as named extra keys of the same JSON line
(`extra.hs_t85l40`, `extra.hs_t170l40`, `extra.hs_t341l60`, `extra.frierson_t85l40`).

The timed region starts on a radiation step, so K steps contain ceil(K / 48) RRTMG calls (pessimistic for K < 48; `steady_state`
gives the same measurement over whole radiation cycles).  One JSON line is printed by rank 0."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = {  # lon, lat, M, dt (s) -- SURVEY.md section 8d (dt by CFL scaling from T42's 600 s)
    "T21": (64, 32, 21, 1200.0), "T42": (128, 64, 42, 600.0), "T85": (256, 128, 85, 300.0),
    "T170": (512, 256, 170, 150.0), "T341": (1024, 512, 341, 75.0),
}
MOIST_DT = {"T21": 900.0, "T42": 720.0, "T85": 360.0, "T170": 150.0, "T341": 75.0}
DT_RAD = 7200


def hs_namelist(res: str, levels: int, tracer: bool = True):
    """exp/test_cases/held_suarez/held_suarez_test_case.py namelist.  The reference's `dry` build carries the sphum grid tracer of
    src/extra/model/dry/field_table (finite-volume horizontal + PPM vertical advection, Held-Suarez tracer source/sink,
    do_water_correction at its default .true.): tracer=True reproduces that; tracer=False is the tracer-free dynamical core."""
    lon, lat, M, dt = RES[res]
    return dict(lon_max=lon, lat_max=lat, num_fourier=M, num_spherical=M + 1, num_levels=levels, dt_atmos=dt,
                damping_order=4, water_correction_limit=200.e2, reference_sea_level_press=1.0e5,
                valid_range_t=(100., 800.), initial_sphum=0.0, vert_coord_option="uneven_sigma",
                scale_heights=6.0, exponent=7.5, surf_res=0.5, do_water_correction=bool(tracer), num_tracers=1 if tracer else 0)


# ---------------------------------------------------------------------------------------------
# algorithmic work per step and rank-share
# ---------------------------------------------------------------------------------------------
def work_model(res: str, K: int, moist: bool = False, nlev_sponge: int = 0):
    I, J, M, _ = RES[res]
    T = (M + 1) * (M + 4) // 2                      # retained (m,n) pairs incl. the extra row
    lev_inv = (2 * K + 2) + (5 * K + 1)             # gradient batch + future-state batch
    lev_fwd = 4 * K + 1
    leg_flops = 2.0 * T * J                         # per level, either direction (complex x real MAC = 4 flop, hemispheric fold)
    four_bytes = 16.0 * (M + 1) * J                 # Fourier intermediate per level
    grid_bytes = 8.0 * I * J
    spec_bytes = 16.0 * T
    col = 8.0 * I * J                               # bytes of one double per column
    w = {
        "legendre_inv": dict(flops=leg_flops * lev_inv, bytes=(spec_bytes + four_bytes) * lev_inv, bound="tensor"),
        "legendre_fwd": dict(flops=leg_flops * lev_fwd, bytes=(spec_bytes + four_bytes) * lev_fwd, bound="tensor"),
        "fft_inv": dict(bytes=(four_bytes + grid_bytes) * lev_inv),
        "fft_fwd": dict(bytes=(four_bytes + grid_bytes) * lev_fwd),
        # grid column kernel: reads u,v,T (cur, prev), vor, div, dxT, dyT (10 x 3-D), writes A,B,dT,Phi,wg_full (5 x 3-D);
        # with external physics tendencies (moist model) 3 more planes are read
        "grid_step": dict(bytes=grid_bytes * K * (18 if moist else 15)),
        # spectral step: specB (4K+1 levels) in, 3x2 state levels in, 3 state out x2, work arrays, specC out (5K+1)
        "spectral": dict(bytes=spec_bytes * K * (4 + 6 + 6 + 8 + 5)),
        # corrections: colsum_energy reads 3 x 3-D, apply_energy r/w 1 x 3-D
        "corrections": dict(bytes=grid_bytes * K * 5, bound="latency"),
        # grid tracer, horizontal step (tracer_horiz_kernel): reads q_prev, u, v; writes tr_future: 4 3-D planes
        "tracer_horiz": dict(bytes=grid_bytes * K * 4),
        # grid tracer, PPM sweep (tracer_ppm_kernel): reads tr_future, wg (K+1), q_prev, q_cur; writes q_fut, q_cur: 7 planes
        "tracer_ppm": dict(bytes=grid_bytes * K * 7),
        # water fixer: column sums (2-D) + apply (reads q_fut, q_cur; writes q_cur): 3 planes
        "tracer_water": dict(bytes=grid_bytes * K * 3),
    }
    if moist:                                        # column physics, bytes per column x columns (kernel headers of physics*.cu)
        w.update({
            "phys_press_heights": dict(bytes=col * 2 * (5 * K + 4)),
            "phys_convection": dict(bytes=col * (8 * K + 9)),
            # conv_post_kernel: reads the increments, T, q and both tendencies (6K), writes T, q after convection and the tendencies (4K)
            "phys_conv_post": dict(bytes=col * (10 * K + 3)),
            "phys_lscale_cond": dict(bytes=col * (6 * K + 2)),
            # cond_post_kernel: reads the increments and both tendencies (4K), writes the tendencies (2K)
            "phys_cond_post": dict(bytes=col * (6 * K + 2)),
            "phys_surface_flux": dict(bytes=col * 45, bound="latency"),
            "phys_radiation": dict(bytes=col * 3 * K),
            "phys_damping": dict(bytes=col * 15 * max(nlev_sponge, 1)),
            # diffusivity: k_m, k_t written on every level (2K) + surface inputs; the reads of t, q, u, v, their tendencies (tau + 1
            # variables are formed in the kernel) and z_full stop at the PBL top (data dependent) and are not counted
            "phys_diffusivity": dict(bytes=col * (2 * K + 4), bound="latency"),
            "phys_vert_diff_down": dict(bytes=col * (19 * K + 14)),
            "phys_mixed_layer_vert_diff_up": dict(bytes=col * (5 * K + 20)),
        })
    return w, dict(T=T, lev_inv=lev_inv, lev_fwd=lev_fwd)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices=(0,)):
        self.ids, self.proc, self.lines, self.t_mark = ",".join(str(i) for i in indices), None, [], 0.0

    def start(self):
        """one nvidia-smi process for all the GPUs of the job (rank 0 runs it), started BEFORE the timed region and waited for until
        its first sample has arrived: NVML initialisation touches every GPU of the box and takes about a second, which would
        otherwise fall into (and, at 8 ranks in lock step, double) a timed region of a few hundred milliseconds"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.ids}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 8.0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        self.t_mark = time.time()                    # samples from here on belong to the timed region

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for t, ln in self.lines if t >= self.t_mark]
        lines = inside if inside else [ln for _, ln in self.lines[-len(self.ids.split(",")):]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json hbm_gbs)")
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


FP64_TENSOR_PEAK_TFLOPS = 37.2   # measured on this pool: tools/probe_fp64.cu, profiles/r01_probe_fp64.txt (DMMA m8n8k4)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from committed `ncu` captures at T170 L40 on 1 GPU
# (profiles/r01i_ncu_summary.txt, profiles/r02/r02a_moist_step_t170_ncu.txt, profiles/r02/r02g_ncu_summary.txt,
# profiles/r02/r02p_ncu_summary.txt); null for kernels / configurations without a capture
NCU_TRAFFIC = {"fft_inv": 448.0e6, "fft_fwd": 244.7e6, "grid_step": 698.4e6, "tracer_horiz": 151.1e6, "tracer_ppm": 290.6e6,
               "legendre_inv": 301.2e6, "legendre_fwd": 365.0e6}
NCU_TRAFFIC_MOIST = {"grid_step": 842.9e6, "phys_vert_diff_down": 1251.3e6, "phys_convection": 578.5e6, "phys_conv_post": 379.1e6,
                     "phys_lscale_cond": 222.5e6, "phys_cond_post": 212.8e6,
                     "phys_press_heights": 309.4e6, "phys_diffusivity": 169.9e6, "tracer_horiz": 195.6e6, "tracer_ppm": 343.0e6,
                     "fft_inv": 447.4e6, "fft_fwd": 245.1e6, "legendre_inv": 301.2e6, "legendre_fwd": 365.4e6}


def group_ms(groups):
    def gsum(prefix):
        return sum(v for k, v in groups.items() if k.startswith(prefix))
    g = {"legendre_inv": gsum("legendre_inv"), "legendre_fwd": gsum("legendre_fwd"), "fft_inv": gsum("fft_inv"),
         "fft_fwd": gsum("fft_fwd"), "grid_step": gsum("grid_step"), "spectral": gsum("spec"),
         "corrections": gsum("corr"), "tracer_horiz": gsum("tracer_horiz") + gsum("tracer_halo"),
         "tracer_ppm": gsum("tracer_ppm"), "tracer_water": gsum("tracer_water") + gsum("tracer_reduce")}
    for k, v in groups.items():
        if k.startswith("phys_") and k != "phys_rrtmg_call":
            g[k] = v
    return g, gsum("exchange")


def roofline_of(groups, wm, world, peaks, traffic_table):
    g_ms, exch_ms = group_ms(groups)
    tot = sum(v for k, v in groups.items() if k != "phys_rrtmg_call")
    cand = {k: v for k, v in g_ms.items() if k in wm and wm[k].get("bound", "hbm") == "hbm" and v > 0}
    dom = max(cand, key=cand.get)
    dom_bytes = wm[dom]["bytes"] / world
    achieved = dom_bytes / (g_ms[dom] * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic_table.get(dom) if traffic_table else None,
            "peak_source": peaks["source"], "share_of_step": g_ms[dom] / tot, "algorithmic_bytes_per_launch": dom_bytes,
            "ms_per_launch": g_ms[dom]}
    leg_ms = g_ms["legendre_inv"] + g_ms["legendre_fwd"]
    leg_fl = (wm["legendre_inv"]["flops"] + wm["legendre_fwd"]["flops"]) / world
    legendre = {"tflops": leg_fl / (leg_ms * 1e-3) / 1e12, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s (fp64 DMMA)",
                "frac": leg_fl / (leg_ms * 1e-3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS, "ms_per_step": leg_ms,
                "algorithmic_flops_per_step": leg_fl, "bound": "tensor (fp64 DMMA)",
                "peak_source": "measured mma.sync m8n8k4 f64 peak on this pool (profiles/r01_probe_fp64.txt)"}
    per_kernel = {}
    for k, v in g_ms.items():
        if k in wm and v > 0:
            b = wm[k]["bytes"] / world
            per_kernel[k] = {"ms": round(v, 5), "bound": wm[k].get("bound", "hbm"), "algorithmic_GBps": round(b / (v * 1e-3) / 1e9, 1),
                             "frac_of_hbm_peak": round(b / (v * 1e-3) / 1e9 / peaks["hbm_gbs"], 3)}
    return roof, legendre, per_kernel, exch_ms


# ---------------------------------------------------------------------------------------------
# CPU arm, every part compiled C++/OpenMP on all host cores: the restatement of the dynamical core + tracer (oracle/cstep) and, for the
# moist workload, the host build of the production column-physics sources and of the RRTMG column arithmetic.  Bounded samples
# (DESIGN.md section 5).
# ---------------------------------------------------------------------------------------------
def cpu_dynamics(res, K, dt, steps, moist=False, spin=2):
    from oracle.isca_oracle import held_suarez_config, frierson_config
    from oracle.cstep import CStep
    cfg = frierson_config(res, K, dt) if moist else held_suarez_config(res, K, dt, num_tracers=1)
    if moist:
        cfg.no_forcing = True                        # the physics tendencies are timed separately (cpu_physics_cpp)
    t0 = time.time()
    cs = CStep(cfg)
    cs.cold_start()
    t_init = time.time() - t0
    cs.step(spin)
    t0 = time.time()
    cs.step(steps)
    sec = (time.time() - t0) / steps
    th = cs.threads
    cs.close()
    return dict(sec_per_step=sec, threads=th, init_s=t_init)


def cpu_physics_cpp_worker(K, I, J):
    """(subprocess of cpu_physics_cpp) the per-step column physics of the MiMA configuration on I x J columns with the HOST build of the
    production column-physics sources (tests/host/build_phys_cpu.py: the same kernels, launches = OpenMP loops over the columns):
    simplified Betts-Miller convection, lscale_cond, surface_flux, Rayleigh sponge, diffusivity (tau + 1 variables), gcm_vert_diff_down,
    mixed_layer, gcm_vert_diff_up -- on a moist, convectively active state of the NumPy oracle (T42 grid) tiled to I x J.  Prints one
    JSON line with the milliseconds spent inside the kernels (the staging copies of the host-array C ABI are not counted)."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import build_phys_cpu
    lib_path = build_phys_cpu.build()
    from isca_b200 import api
    api._lib = C.CDLL(lib_path)                          # this process only: the ctypes mirrors are bound to the host build
    from isca_b200 import physics, moist
    from test_gpu_moist import build
    cfg, core, mp = build("T42", K, 720.0, "SIMPLE_BETTS_MILLER", seed=1, damping=True)
    cur, prev = core.current, core.previous
    rj, ri = J // cfg.lat_max, I // cfg.lon_max
    T3 = lambda a: np.ascontiguousarray(np.tile(a, (1, rj, ri)))
    T2 = lambda a: np.ascontiguousarray(np.tile(a, (rj, ri)))
    tg, q, u, v = T3(core.tg[prev]), T3(core.grid_tracers[prev, 0]), T3(core.ug[prev]), T3(core.vg[prev])
    pf, ph, zf, zh = T3(core.p_full[cur]), T3(core.p_half[cur]), T3(core.z_full[cur]), T3(core.z_half[cur])
    t_surf = T2(mp.t_surf)
    Kk = tg.shape[0]
    dt, delta_t = 150.0, 300.0
    cp = physics.ColumnPhysics(I, J, Kk, **moist.MIMA_PHYSICS_NML)
    ms = api._lib.isca_cpu_kernel_ms
    ms.restype, ms.argtypes = C.c_double, [C.c_int]
    z2, z3 = np.zeros((J, I)), np.zeros_like(tg)
    parts = {}

    def timed(name, f):
        f()                                              # warm-up (first touch of the staging buffers)
        best = None
        for _ in range(3):                               # best of three: other processes' / libraries' spinning threads disturb single runs
            ms(1)
            out = f()
            t = ms(1)
            best = t if best is None else min(best, t)
        parts[name] = best
        return out
    conv = timed("convection", lambda: cp.qe_moist_convection(delta_t, tg, q, pf, ph))
    t2, q2 = tg + conv["deltaT"], q + conv["deltaq"]
    rain, tdel, qdel = timed("lscale_cond", lambda: cp.lscale_cond(t2, q2, pf, ph))
    dt_t, dt_q = (conv["deltaT"] + tdel) / delta_t, (conv["deltaq"] + qdel) / delta_t
    rough = np.full((J, I), 3.21e-05)
    sf = timed("surface_flux", lambda: cp.surface_flux(np.zeros((J, I), np.int32), np.zeros((J, I)), t_atm=tg[-1], q_atm=q[-1], u_atm=u[-1],
                                                       v_atm=v[-1], p_atm=pf[-1], z_atm=zf[-1] - zh[-1], p_surf=ph[-1], t_surf=t_surf,
                                                       t_ca=t_surf, u_surf=z2, v_surf=z2, rough_mom=rough, rough_heat=rough,
                                                       rough_moist=rough, rough_scale=rough, gust=z2))
    pref = np.append(0.5 * (ph[1:, 0, 0] + ph[:-1, 0, 0]), 1.0e5) * 1.0e5 / ph[-1, 0, 0]
    udt, vdt, tdt = timed("damping", lambda: cp.rayleigh_damping(delta_t, pf, u, v, pref))
    dt_u, dt_v, dt_t = udt, vdt, dt_t + tdt
    h, km, kt = timed("diffusivity", lambda: cp.diffusivity(tg + delta_t * dt_t, q + delta_t * dt_q, u + delta_t * dt_u, v + delta_t * dt_v,
                                                            pf, ph, zf, zh, sf["u_star"], sf["b_star"]))
    timed("vert_diff_down", lambda: cp.gcm_vert_diff_down(delta_t, u, v, tg, q, km, kt, ph, pf, zf, sf["flux_u"], sf["flux_v"],
                                                          sf["dtaudu_atm"], sf["dtaudv_atm"], dt_u, dt_v, dt_t, dt_q))
    cp.mixed_layer_init(np.full((J, I), 100.0 * 1.035e3 * 3989.24495292815), z2)
    timed("mixed_layer", lambda: cp.mixed_layer(dt, t_surf, sf["flux_t"], sf["flux_q"], sf["flux_r"], 200.0 + z2, 350.0 + z2, sf["dhdt_surf"],
                                                sf["dedt_surf"], sf["dedq_surf"], sf["drdt_surf"], sf["dhdt_atm"], sf["dedq_atm"]))
    timed("vert_diff_up", lambda: cp.gcm_vert_diff_up(delta_t))
    flags = np.bincount(conv["convflag"].astype(int).ravel(), minlength=3).tolist()
    print(json.dumps(dict(ncol=I * J, K=Kk, kernel_ms=parts, total_ms=sum(parts.values()), convflag_counts=flags,
                          threads=int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)))))


def cpu_physics_cpp(K, I, J):
    """C++/OpenMP per-step column physics on the host cores: see cpu_physics_cpp_worker (run in a subprocess so that this process's
    ctypes bindings stay on the product library)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-physics-cpp", f"{K},{I},{J}"], capture_output=True, text=True, timeout=1800)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("cpu_physics_cpp failed: " + (r.stderr or r.stdout)[-2000:])
    return json.loads(lines[-1])


def cpu_rrtmg_sample(K, ncol=8192):
    """One RRTMG radiation call (LW + SW, half of the columns in daylight) on `ncol` synthetic columns with the C++/OpenMP host build of
    the column arithmetic (tests/host/rrtm_host.cpp, the build `pytest -m "not gpu"` checks against the NumPy oracle), all host cores.
    Returns seconds for the sample; the caller scales by the column count."""
    import ctypes as C
    import subprocess
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from rrtm_bench import columns
    out = os.path.join(ROOT, "oracle", "_build", "librrtm_host_omp.so")
    src = os.path.join(ROOT, "tests", "host", "rrtm_host.cpp")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    table = os.path.join(ROOT, "isca_b200", "data", "rrtmg_tables.bin").encode()
    g = columns(ncol, K)
    F = lambda a: np.asfortranarray(a, dtype=np.float64)
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    a = {k: F(v) for k, v in g.items()}
    co2 = F(np.full((ncol, K), 300e-6))
    u, d, hr = F(np.zeros((ncol, K + 1))), F(np.zeros((ncol, K + 1))), F(np.zeros((ncol, K)))
    cp = C.c_double(287.04 / (2 / 7))
    tsfc = np.ascontiguousarray(g["tsfc"])
    cz = np.ascontiguousarray(np.clip(np.cos(np.linspace(-np.pi, np.pi, ncol)), 0.0, 1.0))
    alb = np.full(ncol, 0.3)
    lib.rrtm_host_lw.restype = lib.rrtm_host_sw.restype = C.c_int

    def lw():
        return lib.rrtm_host_lw(table, cp, ncol, K, P(a["play"]), P(a["plev"]), P(a["tlay"]), P(a["tlev"]), P(tsfc), P(a["h2o"]), P(a["o3"]), P(co2),
                                None, None, None, None, None, None, None, None, P(u), P(d), P(hr), None, None)

    def sw():
        return lib.rrtm_host_sw(table, cp, ncol, K, P(a["play"]), P(a["plev"]), P(a["tlay"]), P(a["h2o"]), P(a["o3"]), P(co2), None, None, None,
                                P(alb), P(cz), C.c_double(1.0), C.c_double(1368.22), P(u), P(d), P(hr), None, None, None)
    if lw() or sw():                                   # first call: table load, page faults
        raise RuntimeError("rrtm_host failed")
    t0 = time.time(); lw(); t_lw = time.time() - t0
    t0 = time.time(); sw(); t_sw = time.time() - t0
    return dict(ncol=ncol, sec_lw=t_lw, sec_sw=t_sw, olr_mean=float(u[:, -1].mean()))


def run_cpu(workload, res, K, steps):
    I, J, M, dt_hs = RES[res]
    if workload == "hs":
        d = cpu_dynamics(res, K, dt_hs, steps)
        sec = d["sec_per_step"]
        sample = (f"{steps} model steps of {res} L{K} after cold start + 2 steps: C++/OpenMP restatement of the step (oracle/cstep, "
                  f"rectangular Legendre loops as the reference) on {d['threads']} host threads")
        return dict(sec_per_step=sec, threads=d["threads"], value=dt_hs / 86400.0 / sec, sample=sample, parts={"dynamics_s": sec})
    dt = MOIST_DT[res]
    d = cpu_dynamics(res, K, dt, steps, moist=True)
    ph = cpu_physics_cpp(K, I, J)
    rr = cpu_rrtmg_sample(K)
    rr_call = (rr["sec_lw"] + rr["sec_sw"]) * (I * J) / rr["ncol"]
    per_rad = DT_RAD / dt
    phys_s = ph["total_ms"] * 1e-3
    sec = d["sec_per_step"] + phys_s + rr_call / per_rad
    sample = (f"all parts C++/OpenMP on {d['threads']} host threads. Dynamics + tracer: {steps} steps of {res} L{K} with the restatement of the "
              f"step (oracle/cstep, rectangular Legendre loops as the reference); per-step column physics (convection, condensation, surface flux, "
              f"sponge, diffusivity, vertical diffusion, mixed layer): the host build of the production kernels' sources (tests/host/build_phys_cpu.py, "
              f"launches = OpenMP loops) on all {I * J} columns of a moist NumPy-oracle state tiled from T42, kernel time only; one RRTMG call "
              f"(LW + SW): host build of the column arithmetic on {rr['ncol']} columns, scaled to {I * J} and amortised over {per_rad:g} steps. "
              f"Not included: compute_pressures_and_heights, the radiation glue")
    return dict(sec_per_step=sec, threads=d["threads"], value=dt / 86400.0 / sec, sample=sample,
                parts={"dynamics_cpp_s": d["sec_per_step"], "physics_cpp_s": phys_s, "physics_cpp_kernels_ms": ph["kernel_ms"],
                       "rrtmg_cpp_scaled_per_call_s": rr_call})


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=480)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mima", choices=["mima", "hs"])
    ap.add_argument("--res", default="T170")
    ap.add_argument("--levels", type=int, default=40)
    ap.add_argument("--spinup", type=int, default=-1, help="on-device spin-up steps before warm-up (default: 2 model days moist, 200 steps hs)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-physics-cpp", default="", help=argparse.SUPPRESS)               # internal: "K,I,J" -> cpu_physics_cpp_worker
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (for `ncu --profile-from-start off`; numbers of such a run are not bench values)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra BASELINE configurations (hs T85/T170/T341, Frierson T85)")
    args = ap.parse_args()
    if args.cpu_physics_cpp:
        cpu_physics_cpp_worker(*[int(x) for x in args.cpu_physics_cpp.split(",")])
        return

    # The contract is ONE JSON line on stdout: keep a private handle on the real stdout for it and point fd 1 at stderr so that
    # library chatter (e.g. NCCL's version banner, printed at communicator creation) cannot land in front of it.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    res, K = args.res, args.levels
    I, J, M, dt_hs = RES[res]
    moistw = args.workload == "mima"
    dt = MOIST_DT[res] if moistw else dt_hs
    spin = args.spinup if args.spinup >= 0 else (int(2 * 86400 / dt) if moistw else 200)
    if moistw:
        workload = (f"MiMA {res} L{K} (lon {I} x lat {J}, dt={dt:g}s; exp/test_cases/MiMA: RRTMG every {DT_RAD}s with an analytic ozone layer, "
                    f"SIMPLE_BETTS_MILLER, lscale_cond, Monin-Obukhov surface flux, diffusivity, vert_diff, 100 m slab with q-flux, Rayleigh sponge; "
                    f"sphum grid tracer), fp64, synthetic cold start + {spin}-step on-device spin-up")
    else:
        workload = (f"Held-Suarez dry core {res} L{K} (lon {I} x lat {J}, dt={dt:g}s), sphum grid tracer of the reference's dry field_table "
                    f"advected (FV + PPM) with water fixer, fp64, synthetic cold start + {spin}-step spin-up")
    metric, unit = "model_days_per_sec", "model-days/s"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        if world > 1:                                    # torchrun pins OMP_NUM_THREADS=1 per rank; this arm is one process on all host cores
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        steps = max(1, min(args.steps, args.cpu_steps if args.steps > 20 else args.steps))
        r = run_cpu(args.workload, res, K, steps)
        line = {
            "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": steps, "warmup": 2, "ms_per_step": r["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload},
            "note": "CPU arm = restatement (port) of the reference algorithm, every part compiled C++/OpenMP on all host cores (cpu_baseline.sample): the Fortran/MPI reference cannot be built here (no Fortran "
                    "compiler in the image or on the GPU box)",
            "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["threads"], "kind": "port", "sample": r["sample"], "parts_s": r["parts"]},
            "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    # ------------------------------------------------------------------ this repo's CUDA arm
    import torch
    from isca_b200 import api, moist
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    J_glob = J
    Jloc = J // world                                # this rank's latitude block

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def new_uid():
        if dist is None:
            return None
        box = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def map_peers(core):
        if world > 1 and os.environ.get("ISCA_B200_NO_P2P") is None:
            handles = [None] * world
            dist.all_gather_object(handles, core.ipc_handles())
            core.set_peer_handles(handles)           # Legendre/FFT epilogues now store straight into peer memory

    use_graph = world == 1 or os.environ.get("ISCA_B200_NO_GRAPH_MULTI") is None
    peaks = measured_peaks()
    pin = lambda shape: torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()

    def analytic_ozone(Kk, Jl, Il, sigma_full):
        """stand-in for ozone_1990.nc (no netCDF input in the benchmark): a stratospheric layer peaking near 10 hPa, mass mixing ratio"""
        p = sigma_full * 1.0e5
        prof = np.where(p < 1.0e4, 1.2e-5 * np.exp(-((np.log(p) - np.log(1.0e3)) ** 2) / 2), 6e-8)
        return np.repeat(np.repeat(prof[:, None, None], Jl, 1), Il, 2)

    def hs_run(r_res, r_K, steps, spinup, profile=True):
        """device-resident Held-Suarez run: ms/step from CUDA events on the library stream, max over ranks"""
        rI, rJ, rM, rdt = RES[r_res]
        atm = api.Atmosphere(api.make_config(**hs_namelist(r_res, r_K, True)), rank=rank, nranks=world, nccl_unique_id=new_uid())
        map_peers(atm)
        atm.cold_start()
        atm.atmosphere(spinup)
        atm.atmosphere(max(args.warmup, 3))
        l0 = atm.get_scalar(api.SC_KERNEL_LAUNCHES)
        barrier()
        t0 = time.time()
        atm.atmosphere(steps)
        barrier()
        wall = max_over_ranks(time.time() - t0)
        ms = max_over_ranks(atm.get_scalar(api.SC_LAST_STEP_MS))
        launches = int(atm.get_scalar(api.SC_KERNEL_LAUNCHES) - l0) * world
        out = {"ms_per_step": ms, "value": rdt / 86400.0 / (ms * 1e-3), "unit": unit, "steps": steps, "steps_per_sec": 1e3 / ms,
               "wall_ms_per_step": wall / steps * 1e3, "n_gpus": world, "dt": rdt,
               "workload": f"Held-Suarez {r_res} L{r_K} with the sphum grid tracer, {spinup}-step spin-up"}
        groups = None
        if profile:
            groups = atm.profile_step(10)
            wm, _ = work_model(r_res, r_K)
            roof, leg, per_kernel, exch = roofline_of(groups, wm, world, peaks,
                                                      NCU_TRAFFIC if (world == 1 and r_res == "T170" and r_K == 40) else None)
            out.update(roofline=roof, legendre_gemm=leg, kernels=per_kernel, exchange_ms_per_step=exch)
        out["T_range_K"] = [atm.get_scalar(api.SC_T_MIN), atm.get_scalar(api.SC_T_MAX)]
        return out, atm, launches, groups

    line = {}
    if moistw:
        m = moist.mima_test_case(res, K, dt, rank=rank, nranks=world, nccl_unique_id=new_uid())
        map_peers(m.core)
        pk, bk = m.core.get_table(api.TB_PK), m.core.get_table(api.TB_BK)
        sig = np.maximum(0.5 * (bk[:-1] + bk[1:]) + 0.5 * (pk[:-1] + pk[1:]) / 1.0e5, 1e-5)
        o3_host = pin((K, Jloc, I))
        o3_host[...] = analytic_ozone(K, Jloc, I, sig)
        m.set_ozone(o3_host)
        per_rad = int(round(DT_RAD / dt))
        warm = max(args.warmup, 3)
        warm += (-(spin + warm)) % per_rad           # the first timed step is a radiation step (the alarm fires at steps 0, 48, 96, ...)
        m.atmosphere(spin)
        m.atmosphere(warm)
        l0 = m.core.get_scalar(api.SC_KERNEL_LAUNCHES)
        sampler = ClockSampler(range(world)) if rank == 0 else None      # one nvidia-smi for all the GPUs of the job
        if sampler:
            sampler.start()
        barrier()
        if sampler:
            sampler.mark()
        if args.profile_range:
            torch.cuda.cudart().cudaProfilerStart()
        t0 = time.time()
        m.atmosphere(args.steps)
        barrier()
        if args.profile_range:
            torch.cuda.cudart().cudaProfilerStop()
        wall = max_over_ranks(time.time() - t0)
        clocks = sampler.stop() if sampler else None
        ms_step, _ = m.timing()
        ms_per_step = max_over_ranks(ms_step)
        core_launches = int(m.core.get_scalar(api.SC_KERNEL_LAUNCHES) - l0)
        rad_calls = -(-args.steps // per_rad)
        # launches of this repo's kernels in the timed region: the dynamical core counts its own; the physics sequence issues 24 per step
        # (moist_model.cu: 2 press_heights, 2 convection, 2 condensation, 2 surface, radiation add, 2 sponge, 2+2 diffusivity, gust fill,
        # vert_diff_down, mixed_layer, vert_diff_up, 4 memsets) and a radiation call 7 (coszen, prepare, fix_top, sw, lw setcoef, lw, finish)
        launches = (core_launches + 24 * args.steps + 7 * rad_calls) * world
        value = dt / 86400.0 / (ms_per_step * 1e-3)
        # steady state over whole radiation cycles (informational; same timing method)
        n_ss = per_rad * max(1, min(10, 480 // per_rad))
        m.atmosphere((-(spin + warm + args.steps)) % per_rad)
        barrier()
        m.atmosphere(n_ss)
        barrier()
        ms_ss = max_over_ranks(m.timing()[0])
        steady = {"ms_per_step": ms_ss, "value": dt / 86400.0 / (ms_ss * 1e-3), "unit": unit, "steps": n_ss,
                  "radiation_calls": n_ss // per_rad}
        # per-kernel-group timings: eager steps with CUDA events on the library stream; the alarm is kept off these steps except one
        m.atmosphere((-(spin + warm + args.steps + n_ss)) % per_rad + 1)
        groups = m.profile_step(per_rad)              # one whole radiation cycle: 47 plain steps + 1 radiation step
        nlev_sp = 0
        wm, _ = work_model(res, K, moist=True, nlev_sponge=max(1, int(np.sum(sig * 1.0e5 < 50.0))))
        roofline, legendre, per_kernel, exch_ms = roofline_of(groups, wm, world, peaks,
                                                              NCU_TRAFFIC_MOIST if (world == 1 and res == "T170" and K == 40) else None)
        rr_ms = groups.get("phys_rrtmg_call", 0.0) * per_rad       # profile_step averages over the cycle's steps
        # ---- end to end: the call a host model makes every step with HOST buffers: the ozone field of the radiation (H2D, what
        # interpolator_mod hands over) in, one atmosphere(Time), and the fields of the MiMA diag_table (ps, precipitation, t_surf, sphum,
        # ucomp, vcomp, temp, vor, div: what send_data passes to diag_manager) out, pinned host memory, copies inside the timed region
        def out_set():
            o = [(0, fid, api.LEVEL_CURRENT, pin((K, Jloc, I))) for fid in (api.F_TRACER0, api.F_U, api.F_V, api.F_T, api.F_VOR, api.F_DIV)]
            o.append((0, api.F_PS, api.LEVEL_CURRENT, pin((Jloc, I))))
            o += [(1, "precip", 0, pin((Jloc, I))), (1, "t_surf", 0, pin((Jloc, I)))]
            return o
        sets = [out_set(), out_set()]                  # two alternating sets of pinned host arrays: set n-1 is consumed while step n runs
        consumed = []

        def e2e_loop(n):
            for i in range(n):
                m.step_io(o3_host, sets[i % 2])        # H2D ozone, atmosphere(Time), D2H of the diag_table fields: pipelined
                if i > 0:
                    m.io_wait(1)                       # the downloads of step i-1 are complete: hand them on (here: read one value each)
                    consumed.append(float(sets[(i - 1) % 2][3][3][0, 0, 0]))
            m.io_sync()
        m.atmosphere((-(spin + warm + args.steps + n_ss + 1 + per_rad)) % per_rad)
        e2e_loop(3)
        barrier()
        t0 = time.time()
        e2e_loop(args.e2e_steps)
        barrier()
        e2e_sec = max_over_ranks(time.time() - t0) / args.e2e_steps
        h2d = 8 * K * J_glob * I
        d2h = (6 * K + 3) * 8 * J_glob * I
        e2e = {"value": dt / 86400.0 / e2e_sec, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_sec * 1e3, "steps": args.e2e_steps,
               "api": "per step: isca_b200_moist_step_io = host ozone field in (H2D), atmosphere(Time), the MiMA diag_table fields (sphum, ucomp, "
                      "vcomp, temp, vor, div, ps, precipitation, t_surf) out to pinned host arrays (D2H); the copies of step n overlap step n+1 "
                      "(two alternating host sets, isca_b200_moist_io_wait), everything inside the timed region incl. the final drain",
               "note": f"{args.e2e_steps} steps after 3 untimed ones, none of them a radiation step" if args.e2e_steps + 3 < per_rad else ""}
        tr = m.core.get_field(api.F_T)
        extra_info = {"radiation_calls_in_timed_region": rad_calls, "steps_per_radiation_call": per_rad,
                      "rrtmg_call_ms": rr_ms, "olr_mean_rank0": float(m.get("olr").mean()),
                      "precip_mean_mm_per_day_rank0": float(m.get("precip").mean() * 86400.0),
                      "T_range_K_rank0": [float(tr.min()), float(tr.max())]}
        m.atmosphere_end()
        working_set_mb = 8.0 * I * J * K * 60 / 1e6
    else:
        out, atm, launches, groups = hs_run(res, K, args.steps, spin)
        sampler = ClockSampler(range(world)) if rank == 0 else None      # clocks of a second pass of the same timed region
        if sampler:
            sampler.start()
        barrier()
        if sampler:
            sampler.mark()
        atm.atmosphere(args.steps)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms_per_step = max_over_ranks(atm.get_scalar(api.SC_LAST_STEP_MS))
        value = dt / 86400.0 / (ms_per_step * 1e-3)
        wall = out["wall_ms_per_step"] * args.steps / 1e3
        roofline, legendre, per_kernel, exch_ms = out["roofline"], out["legendre_gemm"], out["kernels"], out["exchange_ms_per_step"]
        steady = None
        # end to end through spectral_dynamics with HOST tendencies in and HOST state out every step
        n3 = (K, Jloc, I)
        tend = [pin(n3) for _ in range(4)]
        outs = {"psg": pin((Jloc, I)), "ug": pin(n3), "vg": pin(n3), "tg": pin(n3), "grid_tracers": pin(n3)}
        for _ in range(3):
            atm.spectral_dynamics_into(tend, outs)
        barrier()
        t0 = time.time()
        for _ in range(args.e2e_steps):
            atm.spectral_dynamics_into(tend, outs)
        barrier()
        e2e_sec = max_over_ranks(time.time() - t0) / args.e2e_steps
        e2e = {"value": dt / 86400.0 / e2e_sec, "unit": unit, "h2d_bytes_per_step": 4 * 8 * K * J_glob * I,
               "d2h_bytes_per_step": (4 * K + 1) * 8 * J_glob * I, "ms_per_step": e2e_sec * 1e3, "steps": args.e2e_steps,
               "api": "isca_b200_spectral_dynamics_tracers: host tendencies in, host state out, every step (reference spectral_dynamics argument list)"}
        extra_info = {"T_range_K": out["T_range_K"]}
        atm.atmosphere_end()
        working_set_mb = (8.0 * I * J * K * 25 + 16.0 * (M + 1) * J * (5 * K + 1)) / 1e6

    # ------------------------------------------------------------------ the other BASELINE configurations (named extra keys)
    extra = {}
    if not args.no_extra:
        for key, (xr, xk, xsteps, xspin) in (("hs_t85l40", ("T85", 40, 400, 200)), ("hs_t170l40", ("T170", 40, 300, 200)),
                                            ("hs_t341l60", ("T341", 60, 60, 60))):
            if (not moistw) and xr == res and xk == K:
                continue
            try:
                o, a, _, _ = hs_run(xr, xk, xsteps, xspin)
                a.atmosphere_end()
                o["cuda_graph"] = use_graph
                extra[key] = o
            except Exception as e:                     # an extra configuration never takes the headline line down
                extra[key] = {"error": str(e)[:300]}
        if world == 1:
            try:                                       # BASELINE config 3: Frierson grey-radiation aquaplanet T85 L40 on one GPU
                mm = moist.frierson_test_case("T85", 40, 360.0, reference_options=True)
                mm.core.cold_start(); mm.idealized_moist_phys_init()
                mm.atmosphere(int(10 * 86400 / 360.0))
                barrier()
                mm.atmosphere(300)
                barrier()
                ms_m, ms_phys = mm.timing()
                extra["frierson_t85l40"] = {"workload": "Frierson grey-radiation aquaplanet T85 L40 (dt=360s, the scheme namelists of exp/test_cases/frierson: SIMPLE_BETTS_MILLER, slab 2.5 m, sponge, use_tau = .false.), "
                                                        "10-day on-device spin-up", "n_gpus": 1, "ms_per_step": ms_m,
                                            "ms_physics_last_step": ms_phys, "value": 360.0 / 86400.0 / (ms_m * 1e-3), "unit": unit, "steps": 300,
                                            "precip_mean_mm_per_day": float(mm.get("precip").mean() * 86400.0)}
                mm.atmosphere_end()
            except Exception as e:
                extra["frierson_t85l40"] = {"error": str(e)[:300]}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        try:
            r = run_cpu(args.workload, res, K, args.cpu_steps)
            cpu_baseline = {"value": r["value"], "unit": unit, "cores": r["threads"], "kind": "port", "ms_per_step": r["sec_per_step"] * 1e3,
                            "sample": r["sample"], "parts_s": r["parts"]}
            if moistw and not args.no_extra:         # the all-compiled part of the CPU arm on its own: the dry core of the same grid
                rh = run_cpu("hs", res, K, args.cpu_steps)
                extra.setdefault("hs_t170l40" if res == "T170" else f"hs_{res.lower()}l{K}", {})["cpu_baseline"] = {
                    "value": rh["value"], "unit": unit, "cores": rh["threads"], "kind": "port", "ms_per_step": rh["sec_per_step"] * 1e3,
                    "sample": rh["sample"]}
        except Exception as e:
            cpu_baseline = {"error": str(e)[:300]}

    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "parallelism": f"dp{world} (latitudes x zonal wavenumbers)",
                   "steps_per_sec": 1e3 / ms_per_step, "wall_ms_per_step": wall / args.steps * 1e3,
                   "l2": f"per-step working set ~{working_set_mb:.0f} MB > 126 MB L2: inputs larger than L2, no flush",
                   "cuda_graph": bool(use_graph) if not moistw else "dynamics: " + ("graph replay" if use_graph else "eager") + "; physics: eager launches",
                   **extra_info},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline, "legendre_gemm": legendre, "kernels": per_kernel,
        "kernel_groups_ms": {k: round(v, 5) for k, v in groups.items()}, "exchange_ms_per_step": exch_ms,
        "steady_state": steady,
        "cpu_baseline": cpu_baseline,
        "extra": extra,
    }
    print(json.dumps(line), file=json_out, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
