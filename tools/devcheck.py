"""Developer check on a GPU box: parity of every stage against the oracle + first timings.
(test infrastructure; imports oracle/)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.isca_oracle import SpectralCore, held_suarez_config
from isca_b200 import api


def rel(a, b):
    d = np.abs(a - b).max()
    s = max(np.abs(b).max(), 1e-300)
    return d / s


def check(res, K, dt, nsteps=3):
    cfg = held_suarez_config(res, K, dt)
    core = SpectralCore(cfg)
    core.cold_start()
    atm = api.Atmosphere(api.config_from_namelist_object(cfg))
    tb, tr = core.tb, core.tr
    print(f"== {res} L{K}: tables")
    print("   sin_lat", rel(atm.get_table(api.TB_SIN_LAT), tb.sin_lat), "wts", rel(atm.get_table(api.TB_WTS_LAT), tb.wts_lat),
          "bk", rel(atm.get_table(api.TB_BK), core.bk))
    rng = np.random.default_rng(1)
    N, M = cfg.num_spherical, cfg.num_fourier
    nl = 5
    s = (rng.standard_normal((nl, N + 1, M + 1)) + 1j * rng.standard_normal((nl, N + 1, M + 1))) * tb.triangle_mask
    s[:, :, 0] = s[:, :, 0].real
    g = atm.trans_spherical_to_grid(s)
    print("   S2G", rel(g, tr.spherical_to_grid(s)))
    s2 = atm.trans_grid_to_spherical(g)
    print("   G2S", rel(s2, tr.grid_to_spherical(g)), "roundtrip", rel(s2, s))
    s3 = atm.trans_grid_to_spherical(g, do_truncation=False)
    tri1 = (tb.spherical_wave <= cfg.num_fourier + 1)
    print("   G2S notrunc", rel(s3 * tri1, tr.grid_to_spherical(g, do_truncation=False) * tri1))
    v = s.copy(); d = s[::-1].copy(); v[:, 0, 0] = 0; d[:, 0, 0] = 0
    ug, vg = atm.uv_grid_from_vor_div(v, d)
    uo, vo = tr.uv_grid_from_vor_div(v, d)
    print("   uv_from_vor_div", rel(ug, uo), rel(vg, vo))
    v2, d2 = atm.vor_div_from_uv_grid(ug, vg)
    vo2, do2 = tr.vor_div_from_uv_grid(uo, vo)
    print("   vor_div_from_uv", rel(v2, vo2), rel(d2, do2))
    atm.cold_start()
    atm.enable_tendency_capture()
    st, so = atm.state(), core.state()
    print("   cold start:", {k: float(f"{rel(st[k], so[k]):.2e}") for k in ("vors", "ts", "ln_ps", "ug", "vg", "tg", "psg", "vorg", "p_full", "z_full")})
    for i in range(nsteps):
        core.step(keep=True)
        atm.atmosphere(1)
        st, so = atm.state(), core.state()
        errs = {k: float(f"{rel(st[k], so[k]):.2e}") for k in so}
        tend = {k: float(f"{rel(atm.get_spectral(i_), core.last[k]):.2e}") for k, i_ in
                (("dt_vors", api.S_DT_VOR), ("dt_divs", api.S_DT_DIV), ("dt_ts", api.S_DT_T), ("dt_ln_ps", api.S_DT_LNPS))}
        print(f"   step {i+1}: state {errs}")
        print(f"   step {i+1}: tendencies {tend}")
    atm.atmosphere_end()


def check_spunup(res, K, dt, spin=150, nsteps=3):
    """parity from a developed (well-conditioned) state: oracle spin-up, state upload, compare steps"""
    cfg = held_suarez_config(res, K, dt)
    core = SpectralCore(cfg)
    core.cold_start()
    for _ in range(spin):
        core.step()
    atm = api.Atmosphere(api.config_from_namelist_object(cfg))
    for slot in (0, 1):
        atm.set_grid_state(slot, core.ug[slot], core.vg[slot], core.tg[slot], core.psg[slot])
        atm.set_spectral_state(slot, core.vors[slot], core.divs[slot], core.ts[slot], core.ln_ps[slot])
    atm.set_vor_div_grid(core.vorg, core.divg)
    atm.set_time_pointers(core.previous, core.current)
    atm.enable_tendency_capture()
    print(f"== {res} L{K}: spun-up {spin} steps, max|u| = {np.abs(core.ug).max():.2f}")
    for i in range(nsteps):
        core.step(keep=True)
        atm.atmosphere(1)
        st, so = atm.state(), core.state()
        errs = {k: float(f"{rel(st[k], so[k]):.2e}") for k in so}
        tend = {k: float(f"{rel(atm.get_spectral(i_), core.last[k]):.2e}") for k, i_ in
                (("dt_vors", api.S_DT_VOR), ("dt_divs", api.S_DT_DIV), ("dt_ts", api.S_DT_T), ("dt_ln_ps", api.S_DT_LNPS))}
        print(f"   step {i+1}: state {errs}")
        print(f"   step {i+1}: tendencies {tend}")
    atm.atmosphere_end()


def timing(res, K, dt, n=20):
    cfg = held_suarez_config(res, K, dt)
    atm = api.Atmosphere(api.config_from_namelist_object(cfg))
    atm.cold_start()
    atm.atmosphere(5)
    t = time.time(); atm.atmosphere(n); wall = (time.time() - t) / n * 1e3
    print(f"== {res} L{K}: {atm.get_scalar(api.SC_LAST_STEP_MS):.3f} ms/step (events), {wall:.3f} ms wall; launches/step",
          atm.get_scalar(api.SC_KERNEL_LAUNCHES) / atm.get_scalar(api.SC_STEP_COUNT))
    print("   transforms (80 lev):", atm.time_transforms(80, 5))
    st = atm.get_field(api.F_T)
    print("   T range", st.min(), st.max())
    atm.atmosphere_end()


if __name__ == "__main__":
    check("T21", 25, 1200.0)
    check("T42", 25, 600.0)
    check_spunup("T21", 25, 1200.0, spin=400)
    check_spunup("T42", 25, 600.0, spin=100)
    if "--big" in sys.argv:
        check("T85", 40, 300.0, nsteps=2)
    timing("T85", 40, 300.0)
    timing("T170", 40, 150.0)
