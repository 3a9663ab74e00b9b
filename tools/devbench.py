"""Quick per-kernel timing on the GPU box (T170 L40 by default): transforms at the step's batch sizes and the
per-group profile of the full step.  (developer tool)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isca_b200 import api
import bench

res = sys.argv[1] if len(sys.argv) > 1 else "T170"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = api.make_config(**bench.hs_namelist(res, K))
atm = api.Atmosphere.atmosphere_init(cfg)
atm.atmosphere(30)
out = {"plan": os.environ.get("ISCA_B200_FFT_PLAN", "default"), "res": res, "K": K}
out["transforms_283"] = {k: round(v, 4) for k, v in atm.time_transforms(7 * K + 3, 10).items()}
out["transforms_161"] = {k: round(v, 4) for k, v in atm.time_transforms(4 * K + 1, 10).items()}
out["groups"] = {k: round(v, 4) for k, v in atm.profile_step(20).items()}
atm.atmosphere(200)
out["ms_per_step"] = atm.get_scalar(api.SC_LAST_STEP_MS)
print(json.dumps(out))
atm.atmosphere_end()
