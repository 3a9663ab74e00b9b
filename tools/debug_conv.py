import sys, numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from isca_b200 import physics
from oracle import physics as O
from test_gpu_physics import conv_case
K, J, I, seed, nml = 25, 16, 24, 1, dict(rhbm=0.7, Tmin=160.0, Tmax=350.0)
svp, t, q, pf, ph = conv_case(O, K, J, I, seed)
cp = physics.ColumnPhysics(I, J, K, **nml)
g = cp.qe_moist_convection(720.0, t, q, pf, ph)
o = O.SBMConvection(svp, **nml)(720.0, t, q, pf, ph)
for name in ("qref", "Tref", "deltaq", "deltaT"):
    d = np.abs(g[name] - o[name])
    bad = np.argwhere(d > 1e-9 * np.abs(o[name]).max())
    print(name, "mismatches:", len(bad))
    for (k, j, i) in bad[:6]:
        print("  k,j,i", k, j, i, "gpu", g[name][k, j, i], "cpu", o[name][k, j, i], "qin", q[k, j, i], "tin", t[k, j, i], "flag", o["convflag"][j, i],
              "kLZB", o["kLZBs"][j, i], "kLCL", o["kLCLs"][j, i], "rain", o["rain"][j, i], "dq gpu/cpu", g["deltaq"][k, j, i], o["deltaq"][k, j, i])
