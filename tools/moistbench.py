"""Device-resident timing of the idealized moist model (BASELINE config 3: Frierson T85 L40; also T170 L40) on one GPU.
usage: moistbench.py [RES K DT SPIN_DAYS [--ncu-steps N]]   (default: T85 40 360 20 and T170 40 150 10)"""
import ctypes, json, sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from oracle.isca_oracle import frierson_config          # configuration dataclass only (namelist values); no CPU compute is timed here
from isca_b200 import api, moist
from test_gpu_moist import FRIERSON_PHYS

args = [a for a in sys.argv[1:] if not a.startswith("--")]
ncu_steps = int(sys.argv[sys.argv.index("--ncu-steps") + 1]) if "--ncu-steps" in sys.argv else 0
cases = [(args[0], int(args[1]), float(args[2]), float(args[3]))] if len(args) >= 4 else [("T85", 40, 360.0, 20.0), ("T170", 40, 150.0, 10.0)]
out = {}
for res, K, dt, days in cases:
    cfg = frierson_config(res, K, dt)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=FRIERSON_PHYS, mixed_layer_depth=2.5, albedo_value=0.31)
    m.core.cold_start()
    m.idealized_moist_phys_init()
    spin = int(days * 86400 / dt)
    hist = []
    done = 0
    for frac in (0.25, 0.5, 0.75, 1.0):
        n = int(spin * frac) - done
        m.atmosphere(n); done += n
        hist.append(dict(day=round(done * dt / 86400, 2), convflag=np.bincount(m.get("convflag").astype(int).ravel(), minlength=3).tolist(),
                         precip_mean_mm_day=float(m.get("precip").mean() * 86400), t_surf_max=float(m.get("t_surf").max())))
    if ncu_steps:
        rt = ctypes.CDLL("libcudart.so")
        rt.cudaProfilerStart()
        m.atmosphere(ncu_steps)
        rt.cudaProfilerStop()
    steps = 200
    m.atmosphere(steps)
    ms, ms_phys = m.timing()
    t, q = m.core.get_field(api.F_T), m.core.get_field(api.F_TRACER0)
    out[f"frierson_{res}L{K}"] = dict(ms_per_step=round(ms, 4), ms_physics_last_step=round(ms_phys, 4), steps_per_sec=round(1e3 / ms, 1),
                                      model_days_per_sec=round(dt / 86400 / (ms * 1e-3), 4), dt=dt, spin_up_days=days,
                                      T_range=[float(t.min()), float(t.max())], q_max=float(q.max()), history=hist)
    m.atmosphere_end()
print(json.dumps(out, indent=1))
