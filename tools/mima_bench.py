#!/usr/bin/env python
"""Device-resident timing of the MiMA configuration (exp/test_cases/MiMA/MiMA_test_case.py: RRTMG every 7200 s, sponge,
q-flux, 100 m slab) on one GPU.  usage: mima_bench.py [RES K DT SPIN_DAYS STEPS]   (default: T85 40 360 3 240).
Prints ONE JSON line.  bench.py runs it in a subprocess with a timeout (informational arm `mima_model`): the RRTMG kernels'
first GPU runs happen there and must not be able to take the headline line down."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def analytic_ozone(K, J, I, sigma_full):
    """stand-in for ozone_1990.nc (no netCDF input in the benchmark): a stratospheric layer peaking near 10 hPa, mass mixing ratio"""
    p = sigma_full * 1.0e5
    prof = np.where(p < 1.0e4, 1.2e-5 * np.exp(-((np.log(p) - np.log(1.0e3)) ** 2) / 2), 6e-8)
    return np.repeat(np.repeat(prof[:, None, None], J, 1), I, 2)


def main():
    a = sys.argv[1:]
    res, K, dt, days, steps = (a[0], int(a[1]), float(a[2]), float(a[3]), int(a[4])) if len(a) >= 5 else ("T85", 40, 360.0, 3.0, 240)
    from isca_b200 import api, moist
    I, J, M = moist.RESOLUTIONS[res]
    m = moist.mima_test_case(res, K, dt)
    pk, bk = m.core.get_table(api.TB_PK), m.core.get_table(api.TB_BK)
    sig = 0.5 * (bk[:-1] + bk[1:]) + 0.5 * (pk[:-1] + pk[1:]) / 1.0e5
    m.set_ozone(analytic_ozone(K, J, I, np.maximum(sig, 1e-5)))
    m.atmosphere(int(days * 86400 / dt))
    m.atmosphere(steps)
    ms, ms_phys = m.timing()
    t = m.core.get_field(api.F_T)
    dt_rad = 7200 if 7200 % int(dt) == 0 else int(dt) * max(1, round(7200 / dt))
    out = {"workload": f"MiMA {res} L{K} (dt={dt:g}s, RRTMG every {dt_rad}s with analytic ozone, SIMPLE_BETTS_MILLER, sponge, q-flux, "
                       f"100 m slab), {days:g}-day on-device spin-up", "ms_per_step": ms, "ms_physics_last_step": ms_phys,
           "value": dt / 86400.0 / (ms * 1e-3), "unit": "model-days/s", "steps": steps,
           "radiation_steps_in_timed_region": steps * dt / dt_rad,
           "olr_mean": float(m.get("olr").mean()), "toa_sw_mean": float(m.get("toa_sw").mean()),
           "precip_mean_mm_per_day": float(m.get("precip").mean() * 86400.0), "t_surf_range": [float(m.get("t_surf").min()), float(m.get("t_surf").max())],
           "T_range": [float(t.min()), float(t.max())]}
    m.atmosphere_end()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
