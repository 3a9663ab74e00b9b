"""CPU oracle (TEST INFRASTRUCTURE ONLY): Lin-Rood A-grid van-Leer horizontal tracer advection.

NumPy restatement of atmos_spectral/model/fv_advection.F90 (single rank: the 2-row halo exchange of
:161-162,259 reduces to the polar mirror rows of :164-178,266-280).  Arrays are [lev, lat, lon]
(== Fortran (lon, lat, lev)).  Only imported by the oracle / tests."""
from __future__ import annotations
import numpy as np

PI = 3.14159265358979323846


class FVGrid:
    """fv_advection_init (:59-121) with yy = lat_boundaries_global of transforms.F90:314-321."""

    def __init__(self, cfg, tb):
        nx, ny = cfg.lon_max, cfg.lat_max
        self.nx, self.ny = nx, ny
        yy = np.zeros(ny + 1)
        yy[0] = -0.5 * PI
        sum_wts = 0.0
        for j in range(1, ny):
            sum_wts = sum_wts + tb.wts_lat[j - 1]
            yy[j] = np.arcsin(sum_wts - 1.0)
        yy[ny] = 0.5 * PI
        y = 0.5 * (yy[1:] + yy[:-1])
        self.c = np.cos(y)
        self.s = np.sin(y)
        self.cc = np.cos(yy)
        # dy(-1:ny+2) stored with offset 1: dy[j+1] = Fortran dy(j)
        dy = np.zeros(ny + 4)
        dy[2:ny + 2] = yy[1:] - yy[:-1]
        dy[0] = dy[3]            # dy(-1)   = dy(2)
        dy[1] = dy[2]            # dy(0)    = dy(1)
        dy[ny + 2] = dy[ny + 1]  # dy(ny+1) = dy(ny)
        dy[ny + 3] = dy[ny]      # dy(ny+2) = dy(ny-1)
        # dyy(1:ny+1): distance between full points (Fortran index j -> dyy[j-1])
        dyy = np.zeros(ny + 1)
        dyy[1:ny] = y[1:] - y[:-1]
        dyy[0] = 2 * (y[0] - yy[0])
        dyy[ny] = 2 * (yy[ny] - y[ny - 1])
        # dy_plus(0:ny+1), dy_minus(0:ny+1) from the un-scaled dy (:99-100): index j -> [j]
        self.dy_plus = np.array([dy[j + 1] / (dy[j + 1] + dy[j + 2]) for j in range(0, ny + 2)])
        self.dy_minus = np.array([dy[j + 1] / (dy[j] + dy[j + 1]) for j in range(0, ny + 2)])
        self.dy = dy * cfg.radius
        self.dyy = dyy * cfg.radius
        self.dx = 2.0 * PI * cfg.radius / float(nx)

    def DY(self, j):             # Fortran dy(j), j = -1..ny+2
        return self.dy[j + 1]


def _sign1(x):
    return np.where(x >= 0.0, 1.0, -1.0)


def _slope_x(q):
    """:446-479 (monotone=.true.). q [k, j, i]."""
    grad = q - np.roll(q, 1, axis=-1)                     # grad(i) = q(i) - q(i-1)
    slope = (np.roll(grad, -1, axis=-1) + grad) / 2       # (grad(i+1) + grad(i))/2
    qm, qp = np.roll(q, 1, axis=-1), np.roll(q, -1, axis=-1)
    q_min = np.minimum(np.minimum(qm, q), qp)
    q_max = np.maximum(np.maximum(qm, q), qp)
    return _sign1(slope) * np.minimum(np.minimum(np.abs(slope), 2.0 * (q - q_min)), 2.0 * (q_max - q))


def _find_cell_x(b, nx):
    """:427-442: 1-based source cell ii = i-1 - floor(b), wrapped."""
    i = np.arange(1, nx + 1)
    ii = (i - 1)[None, None, :] - np.floor(b).astype(np.int64)
    ii = np.where(ii > nx, ii - nx, ii)
    ii = np.where(ii < 1, ii + nx, ii)
    return ii


def _integer_flux_x(c, q):
    """:483-521 for rows where |b| > 1 (c = b)."""
    nx = q.shape[-1]
    ii = np.trunc(c).astype(np.int64)
    flux = np.zeros_like(q)
    K, J, _ = q.shape
    for k in range(K):
        for j in range(J):
            for i in range(1, nx + 1):
                n = ii[k, j, i - 1]
                if n >= 1:
                    if i - n >= 1:
                        flux[k, j, i - 1] = np.sum(q[k, j, i - n - 1:i - 1])
                    else:
                        flux[k, j, i - 1] = np.sum(q[k, j, 0:i - 1]) + np.sum(q[k, j, i - n + nx - 1:nx])
                elif n <= -1:
                    if i - 1 - n <= nx:
                        flux[k, j, i - 1] = -np.sum(q[k, j, i - 1:i - 1 - n])
                    else:
                        flux[k, j, i - 1] = -np.sum(q[k, j, i - 1:nx]) - np.sum(q[k, j, 0:i - 1 - n - nx])
    return flux


def _vanleer_x(g, dq_dt, uc, q, dt):
    """:330-375."""
    nx = g.nx
    b = uc * dt / (g.dx * g.c[None, :, None])
    bb = b - np.trunc(b)
    flux = np.zeros_like(q)
    rows = np.max(np.abs(b), axis=(0, 2)) > 1.0
    if np.any(rows):
        for j in np.nonzero(rows)[0]:
            flux[:, j:j + 1, :] = _integer_flux_x(b[:, j:j + 1, :], q[:, j:j + 1, :])
    s = _slope_x(q)
    ii = _find_cell_x(b, nx) - 1
    qq = np.take_along_axis(q, ii, axis=-1)
    ss = np.take_along_axis(s, ii, axis=-1)
    flux = flux + bb * (qq + 0.5 * ss * (_sign1(bb) - bb))
    fnext = np.roll(flux, -1, axis=-1)                     # flux(i+1), periodic
    return dq_dt - (fnext - flux) / dt


def _semi_x(g, ua, q, dt):
    """:379-412."""
    nx = g.nx
    b = ua * dt / (g.dx * g.c[None, :, None])
    i_left = _find_cell_x(b, nx)
    i_right = i_left + 1
    i_right = np.where(i_right > nx, 1, i_right)
    bb = b - np.floor(b)
    q_left = np.take_along_axis(q, i_left - 1, axis=-1)
    q_right = np.take_along_axis(q, i_right - 1, axis=-1)
    return bb * q_left + (1.0 - bb) * q_right - q


def _with_polar_halo(q, sign=1.0, rows=2):
    """rows south/north of the poles: value at the antipodal longitude (:164-178, 266-280)."""
    nx = q.shape[-1]
    sh = np.roll(q, -(nx // 2), axis=-1)                   # q(ii(i)) with ii = i + nx/2
    south = [sign * sh[:, r, :] for r in range(rows)]      # row 0 <- row 1(j=1), row -1 <- j=2
    north = [sign * sh[:, q.shape[1] - 1 - r, :] for r in range(rows)]
    parts = [south[r][:, None, :] for r in reversed(range(rows))] + [q] + [north[r][:, None, :] for r in range(rows)]
    return np.concatenate(parts, axis=1)                   # index j (1-based) -> j + rows - 1


def _semi_y(g, va, qx, dt):
    """:416-433. qx has a 2-row halo: Fortran j -> qx[:, j+1, :]."""
    ny = g.ny
    dq = np.zeros_like(va)
    for j in range(1, ny + 1):
        v = va[:, j - 1, :]
        up = v * dt * (qx[:, j, :] - qx[:, j + 1, :]) / g.dyy[j - 1]        # (qx(j-1) - qx(j))/dyy(j)
        dn = v * dt * (qx[:, j + 1, :] - qx[:, j + 2, :]) / g.dyy[j]       # (qx(j) - qx(j+1))/dyy(j+1)
        dq[:, j - 1, :] = np.where(v >= 0.0, up, dn)
    return dq


def _slope_sphere(g, q):
    """:525-550. q with 2-row halo; returns slope for Fortran j = 0..ny+1 -> index j."""
    ny = g.ny
    slope = np.zeros((q.shape[0], ny + 2, q.shape[2]))
    for j in range(0, ny + 2):
        qj, qp, qm = q[:, j + 1, :], q[:, j + 2, :], q[:, j, :]
        slope[:, j, :] = (qp - qj) * g.dy_plus[j] + (qj - qm) * g.dy_minus[j]
    qc = q[:, 1:ny + 3, :]
    q_min = np.minimum(np.minimum(q[:, 0:ny + 2, :], qc), q[:, 2:ny + 4, :])
    q_max = np.maximum(np.maximum(q[:, 0:ny + 2, :], qc), q[:, 2:ny + 4, :])
    return _sign1(slope) * np.minimum(np.minimum(np.abs(slope), 2.0 * (qc - q_min)), 2.0 * (q_max - qc))


def _vanleer_sphere(g, dq_dt, vc, q, dt):
    """:288-326. vc: interfaces j = 1..ny+1 -> index j-1; q with 2-row halo."""
    ny = g.ny
    s = _slope_sphere(g, q)                                                # index j = 0..ny+1
    flux = np.zeros((q.shape[0], ny + 1, q.shape[2]))
    for j in range(1, ny + 2):
        v = vc[:, j - 1, :]
        dtdy_m = dt / g.DY(j - 1)
        dtdy = dt / g.DY(j)
        fp = v * g.cc[j - 1] * (q[:, j, :] + 0.5 * s[:, j - 1, :] * (1.0 - dtdy_m * v))       # q(j-1), s(j-1)
        fm = v * g.cc[j - 1] * (q[:, j + 1, :] - 0.5 * s[:, j, :] * (1.0 + dtdy * v))         # q(j), s(j)
        flux[:, j - 1, :] = np.where(v >= 0.0, fp, fm)
    flux[:, 0, :] = 0.0
    flux[:, ny, :] = 0.0
    out = dq_dt.copy()
    for j in range(1, ny + 1):
        dyc = 1.0 / (g.DY(j) * g.c[j - 1])
        out[:, j - 1, :] = out[:, j - 1, :] - dyc * (flux[:, j, :] - flux[:, j - 1, :])
    return out


def a_grid_horiz_advection(g: FVGrid, ua, va, q, dt, dq_dt):
    """a_grid_horiz_advection_3d (:126-200) + advection_sphere_3d (:241-284), flux=.false."""
    nx, ny = g.nx, g.ny
    vx = _with_polar_halo(va, sign=-1.0, rows=1)           # vx(0) = -vx(ii,1), vx(ny+1) = -vx(ii,ny); index j -> j
    qx = _with_polar_halo(q, sign=1.0, rows=2)             # index j -> j+1
    uc = 0.5 * (np.roll(ua, 1, axis=-1) + ua)              # uc(i) = 0.5 (ua(i-1) + ua(i))
    vc = 0.5 * (vx[:, 0:ny + 1, :] + vx[:, 1:ny + 2, :])   # vc(j) = 0.5 (vx(j-1) + vx(j)), j = 1..ny+1
    div = np.zeros_like(q)
    for j in range(1, ny + 1):
        div[:, j - 1, :] = (vc[:, j, :] * g.cc[j] - vc[:, j - 1, :] * g.cc[j - 1]) / (g.c[j - 1] * g.DY(j))
    ucn = np.roll(uc, -1, axis=-1)                         # uc(i+1)
    div = div + (ucn - uc) / (g.c[None, :, None] * g.dx)
    dq_dt = dq_dt + q * div
    # advection_sphere
    q1 = q + _semi_x(g, ua, q, 0.5 * dt)
    q2 = q + _semi_y(g, va, qx, 0.5 * dt)
    q1x = _with_polar_halo(q1, sign=1.0, rows=2)
    dq_dt = _vanleer_x(g, dq_dt, uc, q2, dt)
    dq_dt = _vanleer_sphere(g, dq_dt, vc, q1x, dt)
    return dq_dt
