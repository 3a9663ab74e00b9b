"""Device-resident timing of the column-physics kernels at the T170 L40 window (run on the GPU box)."""
import json, sys
sys.path.insert(0, ".")
from isca_b200 import physics
names = ["lscale_cond", "two_stream_gray_rad_down", "two_stream_gray_rad_up", "rayleigh_damping", "gcm_vert_diff_down", "gcm_vert_diff_up"]
out = {}
for (I, J, K) in ((512, 256, 40), (1024, 512, 60)):
    cp = physics.ColumnPhysics(I, J, K, do_evap=1, atm_abs=0.2, sponge_pbottom=2.0e5, trayfric=-0.5)
    for w, n in list(enumerate(names)) + [(9, "betts_miller")]:
        ms, by = cp.time_kernel(w, reps=50)
        out[f"{n}_{I}x{J}x{K}"] = dict(ms=round(ms, 5), algorithmic_GBps=round(by / ms / 1e6, 1))
print(json.dumps(out, indent=1))
