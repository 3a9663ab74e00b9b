"""The product's host-side table code (isca_b200/csrc/host_tables.cpp: Gaussian grid, vertical coordinate, Legendre functions,
Laplacian eigenvalues, spectral damping, semi-implicit matrices) through the handle-free entry point isca_b200_host_table: runs
without a GPU, so SURVEY section 8 rows a1, a2, a8, a16 are checked against the oracle (and the committed golden tables) in the CPU
suite as well as on the B200."""
import os
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("res,K,coord", [("T21", 10, None), ("T42", 25, None), ("T85", 30, "uneven_sigma"), ("T42", 20, "hybrid")])
def test_host_tables_match_oracle(lib_built, res, K, coord):
    from isca_b200 import api
    from oracle.isca_oracle import held_suarez_config, SpectralCore
    cfg = held_suarez_config(res, K, 900.0)
    if coord:
        cfg.vert_coord_option = coord
    core = SpectralCore(cfg)
    tb = core.tb
    c = api.config_from_namelist_object(cfg)
    T = lambda i: api.host_table(c, i)
    assert rel(T(api.TB_SIN_LAT), tb.sin_lat) < 1e-14 and rel(T(api.TB_WTS_LAT), tb.wts_lat) < 1e-13
    assert rel(T(api.TB_DEG_LAT), tb.deg_lat) < 1e-14 and rel(T(api.TB_DEG_LON), tb.deg_lon) < 1e-15
    assert rel(T(api.TB_PK), core.pk) < 1e-13 or np.abs(core.pk).max() == 0
    assert rel(T(api.TB_BK), core.bk) < 1e-13
    m, n = T(api.TB_ROW_M).astype(int), T(api.TB_ROW_N).astype(int)
    M = cfg.num_fourier
    assert m.size == sum(M - mm + 2 for mm in range(M + 1))                     # triangular rows incl. the extra meridional one
    assert set(zip(m, n)) == {(mm, nn) for mm in range(M + 1) for nn in range(M - mm + 2)}
    Jh = cfg.lat_max // 2
    leg = T(api.TB_LEGENDRE).reshape(m.size, Jh)
    assert rel(leg, tb.legendre[:, n, m].T) < 1e-12                             # oracle [j, n, m]
    assert rel(T(api.TB_EIGEN_LAPLACIAN), tb.eigen_laplacian[n, m]) < 1e-14
    assert rel(T(api.TB_DAMPING), core.damp.damping[n, m]) < 1e-13
    assert rel(T(api.TB_REF_T), core.impl.ref_t * np.ones(K)) < 1e-14
    # implicit.F90 differentiates ln p_full numerically (eps = 1e-5 of the reference pressure): rounding in the last bit of ln p is
    # amplified by 1e5, so two correct implementations agree to ~1e-10 only
    assert rel(T(api.TB_IMPLICIT_H), core.impl.h) < 1e-8
    assert rel(T(api.TB_DIV_MAT).reshape(K, K), core.impl.div_mat) < 1e-8
    core.impl.dt, core.impl.xi = 2 * cfg.dt_atmos, 2 * cfg.dt_atmos * cfg.alpha_implicit
    core.impl.build_wave_matrices()
    wm = T(api.TB_WAVE_MATRIX).reshape(-1, K, K)
    assert wm.shape[0] == core.impl.wave_matrix.shape[0] or wm.shape[0] == M + 1
    nL = min(wm.shape[0], core.impl.wave_matrix.shape[0])
    assert rel(wm[:nL], core.impl.wave_matrix[:nL]) < 1e-8
    # independent check of the Gaussian grid: NumPy's Gauss-Legendre nodes and weights
    x, w = np.polynomial.legendre.leggauss(cfg.lat_max)
    assert rel(T(api.TB_SIN_LAT), x) < 1e-14 and rel(T(api.TB_WTS_LAT), w) < 2e-12      # weights ~ 1/P_n'(x)^2: conditioning


def test_host_tables_match_the_committed_golden_tables(lib_built):
    from isca_b200 import api
    from oracle.isca_oracle import held_suarez_config
    g = np.load(os.path.join(HERE, "golden", "tables_t42l25.npz"))
    cfg = held_suarez_config("T42", 25, 600.0)
    c = api.config_from_namelist_object(cfg)
    assert rel(api.host_table(c, api.TB_SIN_LAT), g["sin_lat"]) < 1e-14
    assert rel(api.host_table(c, api.TB_WTS_LAT), g["wts_lat"]) < 1e-13
    assert rel(api.host_table(c, api.TB_BK), g["bk"]) < 1e-13
    m, n = api.host_table(c, api.TB_ROW_M).astype(int), api.host_table(c, api.TB_ROW_N).astype(int)
    leg = api.host_table(c, api.TB_LEGENDRE).reshape(m.size, 32)
    for mm, key in ((0, "legendre_m0"), (21, "legendre_m21")):
        rows = np.nonzero(m == mm)[0]
        got = leg[rows[np.argsort(n[rows])]].T                                  # [j, n]
        ref = g[key][:, :got.shape[1]]                                          # the library keeps the triangular rows n <= M - m + 1
        assert got.shape[1] == 42 - mm + 2 and rel(got, ref) < 1e-12, key
    assert rel(api.host_table(c, api.TB_IMPLICIT_H), g["h_impl"]) < 1e-8
    assert rel(api.host_table(c, api.TB_DIV_MAT).reshape(25, 25), g["div_mat"]) < 1e-8


def test_host_table_errors(lib_built):
    from isca_b200 import api
    from oracle.isca_oracle import held_suarez_config
    c = api.config_from_namelist_object(held_suarez_config("T21", 10, 1200.0))
    with pytest.raises(api.IscaError):
        api.host_table(c, 99)
    c.abi_version = 0
    with pytest.raises(api.IscaError):
        api.host_table(c, api.TB_PK)
