"""Generates tests/golden/*.npz with the NumPy oracle (the reference cannot be compiled here: no
Fortran compiler; the reference's own tests hold no golden vectors for this path -- SURVEY.md F5).

    python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.isca_oracle import SpectralCore, held_suarez_config   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cfg = held_suarez_config("T21", 10, 1200.0)
    core = SpectralCore(cfg)
    core.cold_start()
    for _ in range(40):
        core.step()
    st = core.state()
    np.savez_compressed(os.path.join(HERE, "hs_t21l10_40steps.npz"),
                        ln_ps=st["ln_ps"], ts=st["ts"], psg=st["psg"], tg=st["tg"],
                        vors=st["vors"], divs=st["divs"])
    # table fixture: Gaussian latitudes / weights / sigma levels of the T42 L25 Held-Suarez grid
    core42 = SpectralCore(held_suarez_config("T42", 25, 600.0))
    np.savez_compressed(os.path.join(HERE, "tables_t42l25.npz"), sin_lat=core42.tb.sin_lat, wts_lat=core42.tb.wts_lat,
                        bk=core42.bk, legendre_m0=core42.tb.legendre[:, :, 0], legendre_m21=core42.tb.legendre[:, :, 21],
                        h_impl=core42.impl.h, div_mat=core42.impl.div_mat)


if __name__ == "__main__":
    main()
