#!/usr/bin/env python
"""RRTMG kernel timing (informational arm of bench.py, also usable alone): one radiation call's LW and SW kernels on a
T170-sized batch of synthetic clear-sky columns (512 x 256 = 131072 columns, 40 layers), timed with CUDA events through
isca_b200_rrtm_time.  Prints ONE JSON line.  bench.py runs this in a subprocess with a timeout so that a fault in these new
kernels cannot take the headline line down."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def columns(nc, K, seed=7):
    rng = np.random.default_rng(seed)
    ps = rng.uniform(900, 1040, nc)
    sig = np.linspace(0, 1, K + 1) ** 2.2
    ph = sig[None, ::-1] * ps[:, None]
    ph[:, -1] = 0.02
    pl = 0.5 * (ph[:, :-1] + ph[:, 1:])
    ts = 300.0 - 45.0 * rng.uniform(0, 1, nc) ** 2
    z = 7.0 * np.log(ps[:, None] / pl)
    t = np.maximum(ts[:, None] - 6.5 * z, 205.0)
    t = np.where(pl < 30, t + (30 - pl) / 30 * 40.0, t)
    tl = np.empty((nc, K + 1))
    tl[:, 1:-1] = 0.5 * (t[:, :-1] + t[:, 1:])
    tl[:, 0] = t[:, 0]
    tl[:, -1] = t[:, -1]
    h2o = np.maximum(0.02 * np.exp(-z / 2.2) * (ts[:, None] / 300.0) ** 8, 2e-7)
    o3 = np.where(pl < 100, 7e-6 * np.exp(-((np.log(pl) - np.log(10)) ** 2) / 2), 4e-8)
    return dict(play=pl, plev=ph, tlay=t, tlev=tl, tsfc=ts, h2o=h2o, o3=o3)


def main():
    I, J, K = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (512, 256, 40)))
    reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 5
    from isca_b200 import rrtm
    nc = I * J
    g = columns(nc, K)
    r = rrtm.Rrtm(num_lon=I, num_lat=J, num_levels=K)
    t0 = time.time()
    u, d, hr = r.rrtmg_lw(g["play"], g["plev"], g["tlay"], g["tlev"], g["tsfc"], g["h2o"], g["o3"], 300e-6)
    t1 = time.time()
    cz = np.clip(np.cos(np.linspace(-np.pi, np.pi, nc)), 0.0, 1.0)         # half of the columns in daylight, as on the sphere
    su, sd, shr = r.rrtmg_sw(g["play"], g["plev"], g["tlay"], g["h2o"], g["o3"], 300e-6, albedo=0.3, coszen=cz)
    t2 = time.time()
    lw_ms, sw_ms = r.time_kernel(0, reps), r.time_kernel(1, reps)
    out = {"workload": f"RRTMG clear sky, {nc} columns x {K} layers (T170-sized batch when 131072), 140 LW + 112 SW g-points, "
                       "half of the columns in daylight", "columns": nc, "layers": K,
           "lw_kernel_ms": lw_ms, "sw_kernel_ms": sw_ms,
           "lw_columns_per_s": nc / (lw_ms * 1e-3), "sw_columns_per_s": nc / (sw_ms * 1e-3),
           "amortised_ms_per_step_dt_rad_7200_dt_150": (lw_ms + sw_ms) / 48.0,
           "c_abi_host_call_s": {"rrtmg_lw": t1 - t0, "rrtmg_sw": t2 - t1},
           "olr_mean": float(u[:, -1].mean()), "surf_lw_down_mean": float(d[:, 0].mean()),
           "sfc_sw_down_mean": float(sd[:, 0].mean()), "toa_sw_up_mean": float(su[:, -1].mean()),
           "toa_sw_down_over_s0cosz": float((sd[cz > 0, -1] / (1368.22 * cz[cz > 0])).mean()),
           "finite": bool(np.isfinite(u).all() and np.isfinite(hr).all() and np.isfinite(su).all() and np.isfinite(shr).all())}
    r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
