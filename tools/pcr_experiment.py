#!/usr/bin/env python
"""Numerical experiment for DESIGN.md section 10 item 1 (several threads per column in gcm_vert_diff): does a parallel cyclic reduction
of the implicit vertical-diffusion systems stay within the 1e-10 parity bound of the sequential elimination the reference uses
(vert_diff.F90:951-1001)?  The tridiagonal systems are built as vert_diff_down_kernel builds them (mu = g / dp, nu = rho K / dz) on a
moist NumPy-oracle state (T42 L40) with K-profile-like diffusivities; both solvers in fp64.  CPU only (oracle = test infrastructure).
Result (2026-10): state 8e-16, tendency (x - d) / dt 1e-13 ... 3e-12 relative to its maximum -> parity-safe."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_moist import build          # noqa: E402
from oracle import physics as P           # noqa: E402

cfg, core, mp = build("T42", 40, 720.0, "SIMPLE_BETTS_MILLER", seed=1, damping=True)
cur = core.current
t, ph, zf, zh = core.tg[cur], core.p_half[cur], core.z_full[cur], core.z_half[cur]
K = t.shape[0]


def thomas(a, b, c, d):
    e, f, x = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    e[0] = -a[0] / b[0]; f[0] = d[0] / b[0]
    for k in range(1, K):
        g = 1.0 / (b[k] + c[k] * e[k - 1]); e[k] = -a[k] * g; f[k] = (d[k] - c[k] * f[k - 1]) * g
    x[K - 1] = f[K - 1]
    for k in range(K - 2, -1, -1):
        x[k] = e[k] * x[k + 1] + f[k]
    return x


def pcr(a, b, c, d):
    """a couples x[k+1], c couples x[k-1]; log2(K) elimination rounds, every level independent within a round"""
    a, b, c, d = a.copy(), b.copy(), c.copy(), d.copy()
    s = 1
    while s < K:
        an, bn, cn, dn = np.zeros_like(a), b.copy(), np.zeros_like(c), d.copy()
        al, ga = np.zeros_like(a), np.zeros_like(a)
        al[s:] = -c[s:] / b[:-s]; ga[:-s] = -a[:-s] / b[s:]
        bn[s:] += al[s:] * a[:-s]; dn[s:] += al[s:] * d[:-s]; cn[s:] = al[s:] * c[:-s]
        bn[:-s] += ga[:-s] * c[s:]; dn[:-s] += ga[:-s] * d[s:]; an[:-s] = ga[:-s] * a[s:]
        a, b, c, d = an, bn, cn, dn
        s *= 2
    return d / b


def run(diff, delt, label):
    mu = P.GRAV / (ph[1:] - ph[:-1])
    nu = np.zeros_like(t)
    nu[1:] = 2.0 * ph[1:K] / (P.RDGAS * (t[1:] + t[:-1])) * diff[1:] / (zf[:-1] - zf[1:])
    a, c = np.zeros_like(t), np.zeros_like(t)
    a[:-1] = -mu[:-1] * nu[1:] * delt; c[1:] = -mu[1:] * nu[1:] * delt
    b = 1.0 - a - c
    d = t + 0.3 * np.random.default_rng(0).standard_normal(t.shape)
    x1, x2 = thomas(a, b, c, d), pcr(a, b, c, d)
    t1, t2 = (x1 - d) / delt, (x2 - d) / delt
    print(f"{label:28s} max|off-diagonal| {np.abs(a).max():.3f}  state {np.abs(x1 - x2).max() / np.abs(x1).max():.1e}  "
          f"tendency {np.abs(t1 - t2).max() / np.abs(t1).max():.1e}")


zag = zf - zh[K][None]
run(50 * np.exp(-zag / 1000.0) + 0.1, 1440.0, "K-profile, 50 m2/s")
run(300 * np.exp(-zag / 1500.0) + 1.0, 1440.0, "strong mixing, 300 m2/s")
run(50 * np.exp(-zag / 1000.0) + 0.1, 300.0, "dt = 150 s (T170)")
