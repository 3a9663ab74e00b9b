"""Device-resident timing of the idealized moist model (BASELINE config 3: Frierson T85 L40; also T170 L40) on one GPU."""
import json, sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from oracle.isca_oracle import frierson_config          # configuration dataclass only (namelist values); no CPU compute is timed here
from isca_b200 import api, moist
from test_gpu_moist import FRIERSON_PHYS
out = {}
for res, K, dt, spin, steps in (("T85", 40, 360.0, 300, 200), ("T170", 40, 150.0, 300, 200)):
    cfg = frierson_config(res, K, dt)
    m = moist.MoistAtmosphere(api.config_from_namelist_object(cfg), physics_nml=FRIERSON_PHYS, mixed_layer_depth=2.5, albedo_value=0.31)
    m.core.cold_start()
    m.idealized_moist_phys_init()
    m.atmosphere(spin)
    m.atmosphere(steps)
    ms, ms_phys = m.timing()
    t, q = m.core.get_field(api.F_T), m.core.get_field(api.F_TRACER0)
    out[f"frierson_{res}L{K}"] = dict(ms_per_step=round(ms, 4), ms_physics_last_step=round(ms_phys, 4), model_days_per_sec=round(dt / 86400 / (ms * 1e-3), 4),
                                      T_range=[float(t.min()), float(t.max())], q_max=float(q.max()), precip_max=float(m.get("precip").max()),
                                      convflag_counts=np.bincount(m.get("convflag").astype(int).ravel(), minlength=3).tolist())
    m.atmosphere_end()
print(json.dumps(out, indent=1))
