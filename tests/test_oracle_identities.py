"""Pins the CPU oracle (oracle/isca_oracle.py) with mathematical identities.

The reference ships no golden vectors for this path and cannot be compiled here, so these
identities -- plus the committed fixtures in tests/golden/ -- are what the oracle stands on."""
import numpy as np
import pytest
from oracle.isca_oracle import (SpectralCore, Tables, Transforms, Config, held_suarez_config, compute_gaussian,
                                invert, vert_advection, SECOND_CENTERED, FINITE_VOLUME_PARABOLIC, RESOLUTIONS)


@pytest.fixture(scope="module")
def t21():
    cfg = held_suarez_config("T21", 25, 1200.0)
    core = SpectralCore(cfg)
    return core


def rand_spec(tb, nlev, seed=0):
    rng = np.random.default_rng(seed)
    shape = (nlev,) + tb.triangle_mask.shape
    s = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * tb.triangle_mask
    s[:, :, 0] = s[:, :, 0].real
    return s


def test_gauss_weights_and_exactness():
    for nh in (16, 32, 64):
        x, w = compute_gaussian(nh)
        assert abs(w.sum() - 1.0) < 1e-14            # hemisphere weights sum to 1
        assert np.all(np.diff(x) < 0) and x[0] < 1.0  # counted from the pole
        # exact quadrature of even polynomials up to degree 4*nh-2 on [-1,1]
        for deg in (0, 2, 10, 4 * nh - 2):
            assert abs(2 * np.sum(w * x ** deg) - 2.0 / (deg + 1)) < 1e-13


def test_legendre_orthonormality(t21):
    tb = t21.tb
    P, w = tb.legendre, tb.wts_hem
    M = tb.cfg.num_fourier
    for m in (0, 1, 7, M):
        nmax = M - m + 1
        for par in (0, 1):
            Q = P[:, par:nmax + 1:2, m]
            G = 2 * np.einsum("jn,jp,j->np", Q, Q, w)
            assert np.abs(G - np.eye(G.shape[0])).max() < 1e-13


def test_round_trip_and_constant(t21):
    tb, tr = t21.tb, t21.tr
    s = rand_spec(tb, 3)
    assert np.abs(tr.grid_to_spherical(tr.spherical_to_grid(s)) - s).max() < 1e-12
    c = tr.grid_to_spherical(np.full((tb.cfg.lat_max, tb.cfg.lon_max), 3.5))
    assert abs(c[0, 0] - np.sqrt(2.0) * 3.5) < 1e-13      # S(0,0) = sqrt(2) c  (spectral_dynamics.F90:1231)
    c[0, 0] = 0
    assert np.abs(c).max() < 1e-13


def test_analytic_harmonic_and_fft_convention(t21):
    tb, tr = t21.tb, t21.tr
    I = tb.cfg.lon_max
    lon = np.arange(I) * 2 * np.pi / I
    g = np.cos(3 * lon)[None, :] * tb.cos_lat[:, None] ** 3      # ~ Y_3^3
    s = tr.grid_to_spherical(g)
    big = np.abs(s) > 1e-12
    assert big.sum() == 1 and big[0, 3]
    four = tr.grid_to_fourier(g)
    assert np.abs(four[:, 3] - 0.5 * tb.cos_lat ** 3).max() < 1e-14      # c_k = (1/N) sum x e^{-ikx}


def test_operator_identities(t21):
    tb, tr = t21.tb, t21.tr
    v = rand_spec(tb, 2, 1); d = rand_spec(tb, 2, 2)
    v[:, 0, 0] = 0; d[:, 0, 0] = 0
    u, w = tr.uv_grid_from_vor_div(v, d)
    v2, d2 = tr.vor_div_from_uv_grid(u, w)
    assert np.abs(v2 - v).max() < 1e-12 and np.abs(d2 - d).max() < 1e-12
    # solid-body rotation u = U cos(lat): vorticity = 2 U sin(lat) / a
    U = 20.0
    ug = U * tb.cos_lat[:, None] * np.ones((1, tb.cfg.lon_max))
    vor, div = tr.vor_div_from_uv_grid(ug[None], 0 * ug[None])
    vg = tr.spherical_to_grid(vor)[0]
    assert np.abs(vg - 2 * U * tb.sin_lat[:, None] / tb.cfg.radius).max() < 1e-16 + 1e-12 * U / tb.cfg.radius
    assert np.abs(div).max() < 1e-18
    # Laplacian eigenvalues
    s = rand_spec(tb, 1, 3)
    L = tb.spherical_wave
    assert np.abs(tr.laplacian(s) + s * L * (L + 1) / tb.cfg.radius ** 2).max() < 1e-24


def test_matrix_invert_and_implicit(t21):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 7)) + 5 * np.eye(7)
    assert np.abs(invert(A) @ A - np.eye(7)).max() < 1e-13
    impl = t21.impl
    impl.dt = 0.0
    impl.xi = 600.0
    impl.build_wave_matrices()
    K = t21.cfg.num_levels
    for L in (0, 5, 21):
        A = np.eye(K) + impl.xi ** 2 * L * (L + 1) / t21.cfg.radius ** 2 * impl.div_mat
        assert np.abs(impl.wave_matrix[L] @ A - np.eye(K)).max() < 1e-12


def test_implicit_identity_when_alpha_zero():
    cfg = held_suarez_config("T21", 10, 1200.0)
    cfg.alpha_implicit = 0.0
    core = SpectralCore(cfg)
    core.cold_start()
    tb = core.tb
    d, t, p = rand_spec(tb, 10, 4), rand_spec(tb, 10, 5), rand_spec(tb, 1, 6)[0]
    d2, t2, p2 = core.impl.implicit_correction(d.copy(), t.copy(), p.copy(), core.divs, core.ts, core.ln_ps, 1200.0, 0, 0)
    assert np.abs(d2 - d).max() == 0 and np.abs(t2 - t).max() == 0 and np.abs(p2 - p).max() == 0


def test_vert_advection_constant_field_and_schemes():
    rng = np.random.default_rng(0)
    K = 12
    dz = 1000.0 + 500 * rng.random((K, 3, 4))
    w = np.zeros((K + 1, 3, 4)); w[1:K] = 2.0 * rng.standard_normal((K - 1, 3, 4))      # Courant numbers < 1
    r = np.full((K, 3, 4), 7.0)
    for sch in (SECOND_CENTERED, FINITE_VOLUME_PARABOLIC):
        assert np.abs(vert_advection(100.0, w, dz, r, sch)).max() < 1e-12     # advective form: constant stays put
    r = rng.random((K, 3, 4))
    a = vert_advection(100.0, w, dz, r, SECOND_CENTERED)
    # hand formula at an interior level
    k = 5
    flux = lambda kk: w[kk] * 0.5 * (r[kk] + r[kk - 1])
    ref = -(flux(k + 1) - flux(k) - r[k] * (w[k + 1] - w[k])) / dz[k]
    assert np.abs(a[k] - ref).max() < 1e-15


def test_hs_teq_closed_form(t21):
    cfg, tb = t21.cfg, t21.tb
    p = np.full((1, cfg.lat_max, 1), 5.0e4)
    teq = t21.hs.teq(p)[0, :, 0]
    s2 = tb.sin_lat ** 2
    ref = np.maximum(200.0, (315.0 - 60.0 * s2 - 10.0 * np.log(0.5) * (1 - s2)) * 0.5 ** (2.0 / 7.0))
    assert np.abs(teq - ref).max() < 1e-10


def test_step_conserves_mass_and_is_stable(t21):
    core = SpectralCore(held_suarez_config("T21", 25, 1200.0))
    core.cold_start()
    m0 = core.tr.area_weighted_global_mean(core.psg[core.current])
    for _ in range(20):
        core.step()
    m1 = core.tr.area_weighted_global_mean(core.psg[core.current])
    assert abs(m1 - m0) / m0 < 1e-13                       # mass fixer (identity 7)
    assert np.isfinite(core.tg).all()
    # spectral and grid temperature are consistent after the energy fixer
    tg = core.tr.spherical_to_grid(core.ts[core.current])
    assert np.abs(tg - core.tg[core.current]).max() < 1e-9


def test_resolutions_satisfy_check_dynamics_nml():
    for name, r in RESOLUTIONS.items():
        assert r["num_spherical"] == r["num_fourier"] + 1
        assert r["lat_max"] >= (3 * r["num_fourier"] + 1) / 2        # alias-free quadratic terms
        assert r["lon_max"] >= 3 * r["num_fourier"] + 1


def test_hybrid_vertical_coordinate():
    """vert_coord_option = 'hybrid' (vert_coordinate.F90:141-183): pure pressure aloft, pure sigma near the surface, monotone
    half-level pressures, p_half(surface) = ps; the library's host tables agree with the oracle."""
    from oracle.isca_oracle import Config, compute_vert_coord, RESOLUTIONS
    cfg = Config(**RESOLUTIONS["T21"], num_levels=30, dt_atmos=1200.0)
    cfg.vert_coord_option = "hybrid"
    cfg.scale_heights, cfg.surf_res, cfg.exponent, cfg.p_press, cfg.p_sigma = 6.0, 0.5, 3.0, 0.1, 0.3
    cfg.reference_sea_level_press = 1.0e5
    pk, bk = compute_vert_coord(cfg)
    assert pk[-1] == 0.0 and bk[-1] == 1.0                       # surface: sigma
    top = bk == 0.0
    assert top.any() and np.all(pk[top] > 0.0)                   # aloft: pure pressure levels, non-zero top
    for ps in (9.0e4, 1.0e5, 1.05e5):                           # (very low surface pressures fold the blended zone)
        ph = pk + bk * ps
        assert np.all(np.diff(ph) > 0.0) and ph[-1] == ps
    mix = (bk > 0) & (pk > 0)
    assert mix.any()                                             # and a blended zone in between


def test_make_symmetric_keeps_the_model_zonally_symmetric():
    """spectral_dynamics_nml make_symmetric (spherical.F90:185: triangle_mask = 0 for m > 0; the `axisymmetric` test case): the
    truncation removes every zonal wavenumber but 0, so the cold-start perturbation (m = 1, 5) disappears and the run stays zonally
    symmetric; the m = 0 part evolves as in the full model started from a symmetric state."""
    import numpy as np
    from oracle.isca_oracle import SpectralCore, held_suarez_config
    cfg = held_suarez_config("T21", 8, 1200.0)
    cfg.make_symmetric = True
    core = SpectralCore(cfg)
    assert core.tb.triangle_mask[:, 1:].max() == 0.0 and core.tb.triangle_mask[:21, 0].min() == 1.0
    core.cold_start()
    assert np.abs(core.vors).max() == 0.0                   # the perturbation sits at m = 1 and 5
    for _ in range(5):
        core.step()
    st = core.state()
    for k in ("vors", "divs", "ts", "ln_ps"):
        assert np.abs(st[k][..., 1:]).max() == 0.0, k
    for k in ("ug", "tg", "psg"):
        assert np.abs(st[k] - st[k][..., :1]).max() < 1e-12 * np.abs(st[k]).max(), k
    assert np.abs(st["tg"] - 264.0).max() > 1e-3            # the Held-Suarez forcing is acting


def test_spectral_diagnostics_derived_fields():
    """spectral_diagnostics (spectral_dynamics.F90:1747-1835): sea-level pressure equals the surface pressure over a flat surface and
    exceeds it over raised ground by the hypsometric amount of the standard lapse rate; the second moments are plain products"""
    cfg = held_suarez_config("T21", 12, 1200.0, num_tracers=1)
    core = SpectralCore(cfg)
    core.cold_start()
    for _ in range(4):
        core.step()
    d = core.spectral_diagnostics()
    c = core.current
    assert np.allclose(d["slp"], core.psg[c], rtol=1e-14)                    # surf_geopotential = 0
    assert np.array_equal(d["ucomp_vcomp"], core.ug[c] * core.vg[c]) and np.array_equal(d["sphum_w"], core.grid_tracers[c, 0] * core.wg_full)
    assert np.allclose(d["wspd"] ** 2, d["ucomp_sq"] + d["vcomp_sq"], rtol=1e-13)
    core.surf_geopotential = core.surf_geopotential + cfg.grav * 500.0     # a 500 m plateau everywhere
    s2 = core.spectral_diagnostics()["slp"]
    gamma, ps = 0.006, core.psg[c]
    k = np.argmax(core.p_full[c] / ps[None] > 0.8, axis=0)
    tl = np.take_along_axis(core.tg[c], k[None], 0)[0] * (np.take_along_axis(core.p_full[c], k[None], 0)[0] / ps) ** (-cfg.rdgas * gamma / cfg.grav)
    assert np.allclose(s2, ps * (1.0 + gamma * 500.0 / tl) ** (cfg.grav / (cfg.rdgas * gamma)), rtol=1e-13)
    assert np.all(s2 > ps * 1.05) and np.all(s2 < ps * 1.08)                 # ~ exp(500 m / 8 km)
