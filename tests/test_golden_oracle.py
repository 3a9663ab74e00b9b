"""Oracle vs the committed golden fixtures (regression pin of the checker itself; CPU only)."""
import os
import numpy as np
from oracle.isca_oracle import SpectralCore, held_suarez_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tables_fixture():
    g = np.load(os.path.join(GOLDEN, "tables_t42l25.npz"))
    core = SpectralCore(held_suarez_config("T42", 25, 600.0))
    assert np.array_equal(core.tb.sin_lat, g["sin_lat"])
    assert np.array_equal(core.tb.wts_lat, g["wts_lat"])
    assert np.array_equal(core.bk, g["bk"])
    assert np.array_equal(core.tb.legendre[:, :, 21], g["legendre_m21"])
    assert np.abs(core.impl.div_mat - g["div_mat"]).max() <= 1e-12 * np.abs(g["div_mat"]).max()


def test_step_fixture():
    g = np.load(os.path.join(GOLDEN, "hs_t21l10_40steps.npz"))
    core = SpectralCore(held_suarez_config("T21", 10, 1200.0))
    core.cold_start()
    for _ in range(40):
        core.step()
    st = core.state()
    for k in ("ln_ps", "ts", "psg", "tg"):
        assert np.abs(st[k] - g[k]).max() <= 1e-9 * np.abs(g[k]).max(), k
