"""TEST INFRASTRUCTURE ONLY -- NumPy fp64 restatement of the reference's RRTMG clear-sky radiation as Isca runs it.

Only `tests/`, `__graft_entry__.smoke()`, dev tools and `bench.py`'s CPU arm may import this module; the product
(`isca_b200/`) never does.

Parity unpinned: the reference holds no golden vectors for RRTMG and cannot be compiled here (no Fortran compiler), so
this restatement is pinned only by (i) the table extraction being checked against the reference's own netCDF copies
of the coefficients (`tests/test_oracle_rrtmg.py`), (ii) physical properties (flux/heating consistency, energy
conservation of the two-stream adding, known clear-sky magnitudes of a mid-latitude-summer-like column).

Follows, function by function (paths relative to /root/reference/src/atmos_param/rrtm_radiation):
  rrtmg_lw/gcm_model/src/rrtmg_lw_rad.nomcica.f90   rrtmg_lw :81-476, inatm :479-901
  rrtmg_lw/gcm_model/src/rrtmg_lw_setcoef.f90       setcoef :30-410
  rrtmg_lw/gcm_model/src/rrtmg_lw_taumol.f90        taumol, taugb1..taugb16
  rrtmg_lw/gcm_model/src/rrtmg_lw_rtrnmr.f90        rtrnmr (clear-sky branch: `icld = 0` sends the reference there)
  rrtmg_lw/gcm_model/src/rrtmg_lw_init.f90          rrtmg_lw_ini (exp/tfn tables :118-138, constants lwdatinit)
  rrtmg_sw/gcm_model/src/rrtmg_sw_rad.nomcica.f90   rrtmg_sw :73-686, inatm_sw
  rrtmg_sw/gcm_model/src/rrtmg_sw_setcoef.f90       setcoef_sw
  rrtmg_sw/gcm_model/src/rrtmg_sw_taumol.f90        taumol16..taumol29
  rrtmg_sw/gcm_model/src/rrtmg_sw_spcvrt.f90        spcvrt_sw (clear sky, no aerosol: `icld = iaer = 0`)
  rrtmg_sw/gcm_model/src/rrtmg_sw_reftra.f90        reftra_sw
  rrtmg_sw/gcm_model/src/rrtmg_sw_vrtqdr.f90        vrtqdr_sw
  rrtm_radiation.F90                                interp_temp :502-544, run_rrtmg :547-1056

Arrays are [ncol, nlay] with layer index 0 = the LOWEST model layer (RRTMG's order; the glue reverses the model's
top-down order exactly as `run_rrtmg` does).  Table indices (jp, jt, indself, ...) keep the reference's 1-based values.
All column loops are vectorised; layer, band loops are explicit.
"""
import os
import struct

import numpy as np

TABLE_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "isca_b200", "data", "rrtmg_tables.bin")

# lwdatinit / swdatinit (rrtmg_lw_init.f90:213-260, rrtmg_sw_init.f90)
GRAV = 9.8066
AVOGAD = 6.02214199e+23
SECDY = 8.6400e4
AMD, AMW = 28.9660, 18.0160            # inatm: molecular weights of dry air and water vapour
ONEMINUS = 1.0 - 1.0e-6
PI = 2.0 * np.arcsin(1.0)
FLUXFAC = PI * 2.0e4
NTBL, TBLINT, PADE = 10000, 10000.0, 0.278
BPADE = 1.0 / PADE
LW_DELWAVE = np.array([340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.])
LW_NGC = [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2]
SW_NGC = [6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12]
RRSW_SCON = 1.36822e+03                # parrrsw.f90:115
STPFAC = 296.0 / 1013.0


def load_tables(path=TABLE_FILE):
    """reader of the file written by tools/make_rrtmg_tables.py -> {name: ndarray with the Fortran dims, order='F'}"""
    b = open(path, "rb").read()
    magic, ver, n = struct.unpack_from("<8sii", b, 0)
    assert magic == b"ISCARRTM" and ver == 1
    o = 16
    ents = []
    for _ in range(n):
        name, nd, *rest = struct.unpack_from("<48si6iq", b, o)
        o += 84
        ents.append((name.rstrip(b"\0").decode(), nd, rest[:6], rest[6]))
    data = np.frombuffer(b, dtype="<f8", offset=o)
    tab = {}
    for name, nd, dims, off in ents:
        sz = int(np.prod(dims[:nd]))
        tab[name] = data[off:off + sz].reshape(dims[:nd], order="F")
    return tab


_TAB = None


def tables():
    global _TAB
    if _TAB is None:
        _TAB = load_tables()
    return _TAB


def heatfac(cp_air):
    return GRAV * SECDY / (cp_air * 1.0e2)


# ----------------------------------------------------------------------------------------------------------------
# lookup tables of rrtmg_lw_ini (:118-138) and rrtmg_sw_ini
# ----------------------------------------------------------------------------------------------------------------
_LWTBL = None


def lw_exp_tables():
    global _LWTBL
    if _LWTBL is None:
        expeps = 1.0e-20
        tau = np.zeros(NTBL + 1)
        ex = np.zeros(NTBL + 1)
        tfn = np.zeros(NTBL + 1)
        tau[NTBL] = 1.0e10
        ex[0] = 1.0
        ex[NTBL] = expeps
        tfn[NTBL] = 1.0
        itr = np.arange(1, NTBL)
        t = itr / float(NTBL)
        tau[1:NTBL] = BPADE * t / (1.0 - t)
        ex[1:NTBL] = np.maximum(np.exp(-tau[1:NTBL]), expeps)
        tt, ee = tau[1:NTBL], ex[1:NTBL]
        with np.errstate(divide="ignore", invalid="ignore"):
            tfn[1:NTBL] = np.where(tt < 0.06, tt / 6.0, 1.0 - 2.0 * ((1.0 / tt) - (ee / (1.0 - ee))))
        _LWTBL = (tau, ex, tfn)
    return _LWTBL


def sw_exp_table():
    return lw_exp_tables()[1]            # same construction (rrtmg_sw_init.f90: exp_tbl)


# ----------------------------------------------------------------------------------------------------------------
# shared helpers
# ----------------------------------------------------------------------------------------------------------------
def _inatm_common(plev, h2ovmr):
    """coldry of inatm / inatm_sw: dry-air column density (molecules / cm2) per layer"""
    amm = (1.0 - h2ovmr) * AMD + h2ovmr * AMW
    return (plev[:, :-1] - plev[:, 1:]) * 1.0e3 * AVOGAD / (1.0e2 * GRAV * amm * (1.0 + h2ovmr))


def _pt_indices(pavel, tavel, preflog, tref):
    """jp, jt, jt1 and the fac00..fac11 weights (identical in setcoef and setcoef_sw)"""
    plog = np.log(pavel)
    jp = (36.0 - 5.0 * (plog + 0.04)).astype(np.int64)          # int() truncation; argument > 0 in range
    jp = np.clip(jp, 1, 58)
    jp1 = jp + 1
    fp = 5.0 * (preflog[jp - 1] - plog)
    jt = np.trunc(3.0 + (tavel - tref[jp - 1]) / 15.0).astype(np.int64)
    jt = np.clip(jt, 1, 4)
    ft = ((tavel - tref[jp - 1]) / 15.0) - (jt - 3).astype(float)
    jt1 = np.trunc(3.0 + (tavel - tref[jp1 - 1]) / 15.0).astype(np.int64)
    jt1 = np.clip(jt1, 1, 4)
    ft1 = ((tavel - tref[jp1 - 1]) / 15.0) - (jt1 - 3).astype(float)
    compfp = 1.0 - fp
    fac10 = compfp * ft
    fac00 = compfp * (1.0 - ft)
    fac11 = fp * ft1
    fac01 = fp * (1.0 - ft1)
    return plog, jp, jt, jt1, fac00, fac01, fac10, fac11


def _itab(tbl, ind, frac):
    """tbl(ind,ig) + frac*(tbl(ind+1,ig) - tbl(ind,ig)); tbl Fortran dims (nT, ng), ind 1-based [ncol] -> [ncol, ng]"""
    i = np.clip(ind, 1, tbl.shape[0] - 1)
    return tbl[i - 1, :] + frac[:, None] * (tbl[i, :] - tbl[i - 1, :])


def _absview(k):
    """ka(...,ng) -> absa(n, ng) (the reference's equivalence of ka and absa)"""
    ng = k.shape[-1]
    return k.reshape((-1, ng), order="F")


def _g(absx, ind):
    """absx(ind, :) for a 1-based [ncol] index, clipped (out-of-range rows belong to unselected branches)"""
    return absx[np.clip(ind, 1, absx.shape[0]) - 1, :]


def _single(absx, ind0, ind1, f00, f10, f01, f11):
    return (f00[:, None] * _g(absx, ind0) + f10[:, None] * _g(absx, ind0 + 1)
            + f01[:, None] * _g(absx, ind1) + f11[:, None] * _g(absx, ind1 + 1))


def _specparm(cola, colb, rat, mult):
    speccomb = cola + rat * colb
    specparm = np.minimum(cola / speccomb, ONEMINUS)
    specmult = mult * specparm
    js = 1 + specmult.astype(np.int64)
    fs = np.mod(specmult, 1.0)
    return speccomb, specparm, js, fs


def _major3(absa, ind, specparm, fs, fa, fb, speccomb):
    """lower-atmosphere key-species term of one (jp) side: the three-branch interpolation in the species ratio of
    e.g. rrtmg_lw_taumol.f90:taugb3 (specparm < 0.125, > 0.875, otherwise); fa, fb = fac00, fac10 (or fac01, fac11)"""
    lo = specparm < 0.125
    hi = specparm > 0.875
    p = np.where(lo, fs - 1.0, -fs)
    p4 = p ** 4
    fk0 = p4
    fk1 = 1.0 - p - 2.0 * p4
    fk2 = p + p4
    c = lambda x: x[:, None]
    t_lo = (c(fk0 * fa) * _g(absa, ind) + c(fk1 * fa) * _g(absa, ind + 1) + c(fk2 * fa) * _g(absa, ind + 2)
            + c(fk0 * fb) * _g(absa, ind + 9) + c(fk1 * fb) * _g(absa, ind + 10) + c(fk2 * fb) * _g(absa, ind + 11))
    t_hi = (c(fk2 * fa) * _g(absa, ind - 1) + c(fk1 * fa) * _g(absa, ind) + c(fk0 * fa) * _g(absa, ind + 1)
            + c(fk2 * fb) * _g(absa, ind + 8) + c(fk1 * fb) * _g(absa, ind + 9) + c(fk0 * fb) * _g(absa, ind + 10))
    t_mid = (c((1.0 - fs) * fa) * _g(absa, ind) + c(fs * fa) * _g(absa, ind + 1)
             + c((1.0 - fs) * fb) * _g(absa, ind + 9) + c(fs * fb) * _g(absa, ind + 10))
    return c(speccomb) * np.where(c(lo), t_lo, np.where(c(hi), t_hi, t_mid))


def _major2(absb, ind, fs, fa, fb, speccomb, stride):
    c = lambda x: x[:, None]
    return c(speccomb) * (c((1.0 - fs) * fa) * _g(absb, ind) + c(fs * fa) * _g(absb, ind + 1)
                          + c((1.0 - fs) * fb) * _g(absb, ind + stride) + c(fs * fb) * _g(absb, ind + stride + 1))


def _minor2(k, j, f, indm, mfrac):
    """minor species with a (species-ratio, temperature) table k(nsp, 19, ng): the n2om1/n2om2/absn2o pattern"""
    j = np.clip(j, 1, k.shape[0] - 1)
    i = indm
    m1 = k[j - 1, i - 1, :] + f[:, None] * (k[j, i - 1, :] - k[j - 1, i - 1, :])
    m2 = k[j - 1, i, :] + f[:, None] * (k[j, i, :] - k[j - 1, i, :])
    return m1 + mfrac[:, None] * (m2 - m1)


def _fracs2(fr, jpl, fpl):
    """fracrefa(ig, jpl) + fpl*(fracrefa(ig, jpl+1) - fracrefa(ig, jpl)); fr Fortran dims (ng, nsp)"""
    j = np.clip(jpl, 1, fr.shape[1] - 1)
    return fr[:, j - 1].T + fpl[:, None] * (fr[:, j].T - fr[:, j - 1].T)


# ----------------------------------------------------------------------------------------------------------------
# longwave
# ----------------------------------------------------------------------------------------------------------------
def lw_inatm(plev, h2ovmr, co2vmr, o3vmr, n2ovmr, ch4vmr, o2vmr, ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr):
    """inatm (rrtmg_lw_rad.nomcica.f90:479-901): column amounts; CO (wkl 5) is zero as in the reference."""
    coldry = _inatm_common(plev, h2ovmr)
    vmr = [h2ovmr, co2vmr, o3vmr, n2ovmr, np.zeros_like(h2ovmr), ch4vmr, o2vmr]
    summol = np.zeros_like(coldry)
    for v in vmr[1:]:
        summol = summol + v
    wbrodl = coldry * (1.0 - summol)
    wkl = [coldry * v for v in vmr]
    wx = [coldry * v * 1.0e-20 for v in (ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr)]
    amttl = np.zeros(coldry.shape[0])
    wvttl = np.zeros(coldry.shape[0])
    for l in range(coldry.shape[1]):
        amttl = amttl + coldry[:, l] + wkl[0][:, l]
        wvttl = wvttl + wkl[0][:, l]
    wvsh = (AMW * wvttl) / (AMD * amttl)
    pwvcm = wvsh * (1.0e3 * plev[:, 0]) / (1.0e2 * GRAV)
    return coldry, wkl, wbrodl, wx, pwvcm


def lw_setcoef(pavel, tavel, tz, tbound, semiss, coldry, wkl, wbroad):
    """setcoef (rrtmg_lw_setcoef.f90:30-410) -> dict of per-layer coefficients, Planck terms per band"""
    T = tables()
    totplnk = T["lw_totplnk"]                      # (181, 16)
    nc, nl = pavel.shape

    def pl(t):
        ind = np.clip((t - 159.0).astype(np.int64), 1, 180)
        frac = t - 159.0 - ind.astype(float)
        return ind, frac

    indbound, tbndfrac = pl(tbound)
    indlev0, t0frac = pl(tz[:, 0])
    planklay = np.zeros((nc, nl, 16))
    planklev = np.zeros((nc, nl + 1, 16))
    plankbnd = semiss * (totplnk[indbound - 1, :] + tbndfrac[:, None] * (totplnk[indbound, :] - totplnk[indbound - 1, :]))
    planklev[:, 0, :] = totplnk[indlev0 - 1, :] + t0frac[:, None] * (totplnk[indlev0, :] - totplnk[indlev0 - 1, :])
    for lay in range(nl):
        indlay, tlayfrac = pl(tavel[:, lay])
        indlev, tlevfrac = pl(tz[:, lay + 1])
        planklay[:, lay, :] = totplnk[indlay - 1, :] + tlayfrac[:, None] * (totplnk[indlay, :] - totplnk[indlay - 1, :])
        planklev[:, lay + 1, :] = totplnk[indlev - 1, :] + tlevfrac[:, None] * (totplnk[indlev, :] - totplnk[indlev - 1, :])

    plog, jp, jt, jt1, fac00, fac01, fac10, fac11 = _pt_indices(pavel, tavel, T["lw_preflog"], T["lw_tref"])
    lower = plog > 4.56                                        # `if (plog .le. 4.56) go to 5300`
    laytrop = lower.sum(axis=1)
    water = wkl[0] / coldry
    scalefac = pavel * STPFAC / tavel
    forfac = scalefac / (1.0 + water)
    factor = (332.0 - tavel) / 36.0
    indfor_l = np.minimum(2, np.maximum(1, factor.astype(np.int64)))
    forfrac_l = factor - indfor_l
    factor_u = (tavel - 188.0) / 36.0
    indfor = np.where(lower, indfor_l, 3)
    forfrac = np.where(lower, forfrac_l, factor_u - 1.0)
    selffac = water * forfac
    factor = (tavel - 188.0) / 7.2
    indself = np.minimum(9, np.maximum(1, np.trunc(factor).astype(np.int64) - 7))
    selffrac = factor - (indself + 7)
    scaleminor = pavel / tavel
    scaleminorn2 = (pavel / tavel) * (wbroad / (coldry + wkl[0]))
    factor = (tavel - 180.8) / 7.2
    indminor = np.minimum(18, np.maximum(1, np.trunc(factor).astype(np.int64)))
    minorfrac = factor - indminor
    col = {}
    for name, i in (("h2o", 0), ("co2", 1), ("o3", 2), ("n2o", 3), ("co", 4), ("ch4", 5), ("o2", 6)):
        col[name] = 1.0e-20 * wkl[i]
    for name in ("co2", "o3", "n2o", "co", "ch4"):
        col[name] = np.where(col[name] == 0.0, 1.0e-32 * coldry, col[name])
    colbrd = 1.0e-20 * wbroad
    selffac = col["h2o"] * selffac
    forfac = col["h2o"] * forfac
    return dict(laytrop=laytrop, lower=lower, jp=jp, jt=jt, jt1=jt1, fac00=fac00, fac01=fac01, fac10=fac10, fac11=fac11,
                forfac=forfac, forfrac=forfrac, indfor=indfor, selffac=selffac, selffrac=selffrac, indself=indself,
                scaleminor=scaleminor, scaleminorn2=scaleminorn2, indminor=indminor, minorfrac=minorfrac,
                col=col, colbrd=colbrd, planklay=planklay, planklev=planklev, plankbnd=plankbnd)


def _adjcol(colx, coldry, chiref, thresh, base, expo):
    """the `ratco2 .gt. 3.0 -> adjfac = 2.0+(ratco2-2.0)**0.77` pattern of taugb3/6/7/8/9/13"""
    chi = colx / coldry
    rat = 1.0e20 * chi / chiref
    with np.errstate(invalid="ignore"):
        adjfac = base + (np.maximum(rat, base) - base) ** expo
    return np.where(rat > thresh, adjfac * chiref * coldry * 1.0e-20, colx)


def lw_taumol(pavel, wx, coldry, sc):
    """taumol (rrtmg_lw_taumol.f90): gaseous optical depth taug[ncol, nlay, 140] and Planck fractions fracs"""
    T = tables()
    chi = T["lw_chi_mls"]                                       # chi_mls(7, 59)
    nc, nl = pavel.shape
    ngpt = sum(LW_NGC)
    taug = np.zeros((nc, nl, ngpt))
    fracs = np.zeros((nc, nl, ngpt))
    ngs = np.concatenate([[0], np.cumsum(LW_NGC)])
    col = sc["col"]
    CHI = lambda sp, lev: chi[sp - 1, lev - 1]

    for lay in range(nl):
        low = sc["lower"][:, lay]
        lowc = low[:, None]
        jp, jt, jt1 = sc["jp"][:, lay], sc["jt"][:, lay], sc["jt1"][:, lay]
        f00, f01, f10, f11 = sc["fac00"][:, lay], sc["fac01"][:, lay], sc["fac10"][:, lay], sc["fac11"][:, lay]
        inds, indf, indm = sc["indself"][:, lay], sc["indfor"][:, lay], sc["indminor"][:, lay]
        selffac, selffrac = sc["selffac"][:, lay], sc["selffrac"][:, lay]
        forfac, forfrac = sc["forfac"][:, lay], sc["forfrac"][:, lay]
        mfrac = sc["minorfrac"][:, lay]
        h2o, co2, o3, n2o, co, ch4, o2 = (col[k][:, lay] for k in ("h2o", "co2", "o3", "n2o", "co", "ch4", "o2"))
        cdry = coldry[:, lay]
        colbrd = sc["colbrd"][:, lay]
        pp = pavel[:, lay]
        chi_jp1 = lambda sp: chi[sp - 1, jp]                    # chi_mls(sp, jp+1)
        rat = lambda a, b: chi[a - 1, jp - 1] / chi[b - 1, jp - 1]
        rat1 = lambda a, b: chi[a - 1, jp] / chi[b - 1, jp]
        c = lambda x: x[:, None]

        def ind_lower(nsp, js, js1):
            return ((jp - 1) * 5 + (jt - 1)) * nsp + js, (jp * 5 + (jt1 - 1)) * nsp + js1

        def ind_upper(nsp, js, js1):
            return ((jp - 13) * 5 + (jt - 1)) * nsp + js, ((jp - 12) * 5 + (jt1 - 1)) * nsp + js1

        def tself(b):
            return c(selffac) * _itab(T["lw%02d_selfref" % b], inds, selffrac)

        def tfor(b):
            return c(forfac) * _itab(T["lw%02d_forref" % b], indf, forfrac)

        def single_lower(b):
            i0, i1 = ind_lower(1, 1, 1)
            return _single(_absview(T["lw%02d_ka" % b]), i0, i1, f00, f10, f01, f11)

        def single_upper(b):
            i0, i1 = ind_upper(1, 1, 1)
            return _single(_absview(T["lw%02d_kb" % b]), i0, i1, f00, f10, f01, f11)

        def binary_lower(b, ca, cb, sa, sb):
            sc0, sp0, js, fs = _specparm(ca, cb, rat(sa, sb), 8.0)
            sc1, sp1, js1, fs1 = _specparm(ca, cb, rat1(sa, sb), 8.0)
            i0, i1 = ind_lower(9, js, js1)
            absa = _absview(T["lw%02d_ka" % b])
            return _major3(absa, i0, sp0, fs, f00, f10, sc0) + _major3(absa, i1, sp1, fs1, f01, f11, sc1)

        def binary_upper(b, ca, cb, sa, sb):
            sc0, sp0, js, fs = _specparm(ca, cb, rat(sa, sb), 4.0)
            sc1, sp1, js1, fs1 = _specparm(ca, cb, rat1(sa, sb), 4.0)
            i0, i1 = ind_upper(5, js, js1)
            absb = _absview(T["lw%02d_kb" % b])
            return _major2(absb, i0, fs, f00, f10, sc0, 5) + _major2(absb, i1, fs1, f01, f11, sc1, 5)

        def planck2(fr, ca, cb, refrat, mult):
            _, _, jpl, fpl = _specparm(ca, cb, refrat, mult)
            return _fracs2(fr, jpl, fpl)

        def minor2(k, ca, cb, refrat, mult):
            _, _, jm, fm = _specparm(ca, cb, refrat, mult)
            return _minor2(k, jm, fm, indm, mfrac)

        def minor1(k):
            return _itab(k, indm, mfrac)

        def put(b, tau_l, fr_l, tau_u, fr_u):
            s = slice(ngs[b - 1], ngs[b])
            ng = LW_NGC[b - 1]
            z = np.zeros((nc, ng))
            tl = z if tau_l is None else tau_l
            tu = z if tau_u is None else tau_u
            bl = lambda f: z if f is None else (np.broadcast_to(f, (nc, ng)) if f.ndim == 1 else f)
            taug[:, lay, s] = np.where(lowc, tl, tu)
            fracs[:, lay, s] = np.where(lowc, bl(fr_l), bl(fr_u))

        # ---- band 1: 10-350 cm-1 (low key h2o; low minor n2) (high key h2o; high minor n2)
        corr_l = np.where(pp < 250.0, 1.0 - 0.15 * (250.0 - pp) / 154.4, 1.0)
        corr_u = 1.0 - 0.15 * (pp / 95.6)
        scalen2 = colbrd * sc["scaleminorn2"][:, lay]
        tl = c(corr_l) * (c(h2o) * single_lower(1) + tself(1) + tfor(1) + c(scalen2) * minor1(T["lw01_ka_mn2"]))
        tu = c(corr_u) * (c(h2o) * single_upper(1) + tfor(1) + c(scalen2) * minor1(T["lw01_kb_mn2"]))
        put(1, tl, T["lw01_fracrefa"], tu, T["lw01_fracrefb"])
        # ---- band 2: 350-500 (h2o)
        corr_l = 1.0 - 0.05 * (pp - 100.0) / 900.0
        tl = c(corr_l) * (c(h2o) * single_lower(2) + tself(2) + tfor(2))
        tu = c(h2o) * single_upper(2) + tfor(2)
        put(2, tl, T["lw02_fracrefa"], tu, T["lw02_fracrefb"])
        # ---- band 3: 500-630 (h2o, co2; minor n2o)
        adjn2o = _adjcol(n2o, cdry, chi_jp1(4), 1.5, 0.5, 0.65)
        tl = (binary_lower(3, h2o, co2, 1, 2) + tself(3) + tfor(3)
              + c(adjn2o) * minor2(T["lw03_ka_mn2o"], h2o, co2, CHI(1, 3) / CHI(2, 3), 8.0))
        fl = planck2(T["lw03_fracrefa"], h2o, co2, CHI(1, 9) / CHI(2, 9), 8.0)
        tu = (binary_upper(3, h2o, co2, 1, 2) + tfor(3)
              + c(adjn2o) * minor2(T["lw03_kb_mn2o"], h2o, co2, CHI(1, 13) / CHI(2, 13), 4.0))
        fu = planck2(T["lw03_fracrefb"], h2o, co2, CHI(1, 13) / CHI(2, 13), 4.0)
        put(3, tl, fl, tu, fu)
        # ---- band 4: 630-700 (h2o, co2) (o3, co2)
        tl = binary_lower(4, h2o, co2, 1, 2) + tself(4) + tfor(4)
        fl = planck2(T["lw04_fracrefa"], h2o, co2, CHI(1, 11) / CHI(2, 11), 8.0)
        tu = binary_upper(4, o3, co2, 3, 2).copy()
        tu[:, 7:14] *= np.array([0.92, 0.88, 1.07, 1.1, 0.99, 0.88, 0.943])
        fu = planck2(T["lw04_fracrefb"], o3, co2, CHI(3, 13) / CHI(2, 13), 4.0)
        put(4, tl, fl, tu, fu)
        # ---- band 5: 700-820 (h2o, co2; minor o3, ccl4) (o3, co2)
        ccl4 = c(wx[0][:, lay]) * T["lw05_ccl4"][None, :]
        tl = (binary_lower(5, h2o, co2, 1, 2) + tself(5) + tfor(5)
              + minor2(T["lw05_ka_mo3"], h2o, co2, CHI(1, 7) / CHI(2, 7), 8.0) * c(o3) + ccl4)
        fl = planck2(T["lw05_fracrefa"], h2o, co2, CHI(1, 5) / CHI(2, 5), 8.0)
        tu = binary_upper(5, o3, co2, 3, 2) + ccl4
        fu = planck2(T["lw05_fracrefb"], o3, co2, CHI(3, 43) / CHI(2, 43), 4.0)
        put(5, tl, fl, tu, fu)
        # ---- band 6: 820-980 (h2o; minor co2, cfc11, cfc12) (nothing; cfc11, cfc12)
        adjco2 = _adjcol(co2, cdry, chi_jp1(2), 3.0, 2.0, 0.77)
        cfc = c(wx[1][:, lay]) * T["lw06_cfc11adj"][None, :] + c(wx[2][:, lay]) * T["lw06_cfc12"][None, :]
        tl = c(h2o) * single_lower(6) + tself(6) + tfor(6) + c(adjco2) * minor1(T["lw06_ka_mco2"]) + cfc
        tu = 0.0 + cfc
        put(6, tl, T["lw06_fracrefa"], tu, T["lw06_fracrefa"])
        # ---- band 7: 980-1080 (h2o, o3; minor co2) (o3; minor co2)
        adjco2 = _adjcol(co2, cdry, chi_jp1(2), 3.0, 3.0, 0.79)
        tl = (binary_lower(7, h2o, o3, 1, 3) + tself(7) + tfor(7)
              + c(adjco2) * minor2(T["lw07_ka_mco2"], h2o, o3, CHI(1, 3) / CHI(3, 3), 8.0))
        fl = planck2(T["lw07_fracrefa"], h2o, o3, CHI(1, 3) / CHI(3, 3), 8.0)
        adjco2u = _adjcol(co2, cdry, chi_jp1(2), 3.0, 2.0, 0.79)
        tu = (c(o3) * single_upper(7) + c(adjco2u) * minor1(T["lw07_kb_mco2"])).copy()
        tu[:, 5:11] *= np.array([0.92, 0.88, 1.07, 1.1, 0.99, 0.855])
        put(7, tl, fl, tu, T["lw07_fracrefb"])
        # ---- band 8: 1080-1180 (h2o; minor co2, o3, n2o, cfc12, cfc22) (o3; minor co2, n2o)
        adjco2 = _adjcol(co2, cdry, chi_jp1(2), 3.0, 2.0, 0.65)
        cfc = c(wx[2][:, lay]) * T["lw08_cfc12"][None, :] + c(wx[3][:, lay]) * T["lw08_cfc22adj"][None, :]
        tl = (c(h2o) * single_lower(8) + tself(8) + tfor(8) + c(adjco2) * minor1(T["lw08_ka_mco2"])
              + c(o3) * minor1(T["lw08_ka_mo3"]) + c(n2o) * minor1(T["lw08_ka_mn2o"]) + cfc)
        tu = (c(o3) * single_upper(8) + c(adjco2) * minor1(T["lw08_kb_mco2"])
              + c(n2o) * minor1(T["lw08_kb_mn2o"]) + cfc)
        put(8, tl, T["lw08_fracrefa"], tu, T["lw08_fracrefb"])
        # ---- band 9: 1180-1390 (h2o, ch4; minor n2o) (ch4; minor n2o)
        adjn2o = _adjcol(n2o, cdry, chi_jp1(4), 1.5, 0.5, 0.65)
        tl = (binary_lower(9, h2o, ch4, 1, 6) + tself(9) + tfor(9)
              + c(adjn2o) * minor2(T["lw09_ka_mn2o"], h2o, ch4, CHI(1, 3) / CHI(6, 3), 8.0))
        fl = planck2(T["lw09_fracrefa"], h2o, ch4, CHI(1, 9) / CHI(6, 9), 8.0)
        tu = c(ch4) * single_upper(9) + c(adjn2o) * minor1(T["lw09_kb_mn2o"])
        put(9, tl, fl, tu, T["lw09_fracrefb"])
        # ---- band 10: 1390-1480 (h2o)
        tl = c(h2o) * single_lower(10) + tself(10) + tfor(10)
        tu = c(h2o) * single_upper(10) + tfor(10)
        put(10, tl, T["lw10_fracrefa"], tu, T["lw10_fracrefb"])
        # ---- band 11: 1480-1800 (h2o; minor o2)
        scaleo2 = o2 * sc["scaleminor"][:, lay]
        tl = c(h2o) * single_lower(11) + tself(11) + tfor(11) + c(scaleo2) * minor1(T["lw11_ka_mo2"])
        tu = c(h2o) * single_upper(11) + tfor(11) + c(scaleo2) * minor1(T["lw11_kb_mo2"])
        put(11, tl, T["lw11_fracrefa"], tu, T["lw11_fracrefb"])
        # ---- band 12: 1800-2080 (h2o, co2) (nothing)
        tl = binary_lower(12, h2o, co2, 1, 2) + tself(12) + tfor(12)
        fl = planck2(T["lw12_fracrefa"], h2o, co2, CHI(1, 10) / CHI(2, 10), 8.0)
        put(12, tl, fl, None, None)
        # ---- band 13: 2080-2250 (h2o, n2o; minor co2, co) (nothing; minor o3)
        adjco2 = _adjcol(co2, cdry, 3.55e-4, 3.0, 2.0, 0.68)
        tl = (binary_lower(13, h2o, n2o, 1, 4) + tself(13) + tfor(13)
              + c(adjco2) * minor2(T["lw13_ka_mco2"], h2o, n2o, CHI(1, 1) / CHI(4, 1), 8.0)
              + c(co) * minor2(T["lw13_ka_mco"], h2o, n2o, CHI(1, 3) / CHI(4, 3), 8.0))
        fl = planck2(T["lw13_fracrefa"], h2o, n2o, CHI(1, 5) / CHI(4, 5), 8.0)
        tu = c(o3) * minor1(T["lw13_kb_mo3"])
        put(13, tl, fl, tu, T["lw13_fracrefb"])
        # ---- band 14: 2250-2380 (co2)
        tl = c(co2) * single_lower(14) + tself(14) + tfor(14)
        tu = c(co2) * single_upper(14)
        put(14, tl, T["lw14_fracrefa"], tu, T["lw14_fracrefb"])
        # ---- band 15: 2380-2600 (n2o, co2; minor n2) (nothing)
        scalen2 = colbrd * sc["scaleminor"][:, lay]
        tl = (binary_lower(15, n2o, co2, 4, 2) + tself(15) + tfor(15)
              + c(scalen2) * minor2(T["lw15_ka_mn2"], n2o, co2, CHI(4, 1) / CHI(2, 1), 8.0))
        fl = planck2(T["lw15_fracrefa"], n2o, co2, CHI(4, 1) / CHI(2, 1), 8.0)
        put(15, tl, fl, None, None)
        # ---- band 16: 2600-3250 (h2o, ch4) (ch4)
        tl = binary_lower(16, h2o, ch4, 1, 6) + tself(16) + tfor(16)
        fl = planck2(T["lw16_fracrefa"], h2o, ch4, CHI(1, 6) / CHI(6, 6), 8.0)
        tu = c(ch4) * single_upper(16)
        put(16, tl, fl, tu, T["lw16_fracrefb"])
    return taug, fracs


def lw_rtrnmr_clear(pz, semiss, sc, pwvcm, fracs, taut, cp_air):
    """rtrnmr with no cloudy layer (rrtmg_lw_rtrnmr.f90:390-700, `icldlyr = 0` branches)"""
    tau_tbl, exp_tbl, tfn_tbl = lw_exp_tables()
    nc, nl, ngpt = taut.shape
    a0 = np.array([1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668] + [1.66] * 7)
    a1 = np.array([0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006] + [0.0] * 7)
    a2 = np.array([0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414] + [0.0] * 7)
    secdiff = np.zeros((nc, 16))
    for ib in range(16):
        if ib == 0 or ib == 3 or ib >= 9:
            secdiff[:, ib] = 1.66
        else:
            secdiff[:, ib] = np.clip(a0[ib] + a1[ib] * np.exp(a2[ib] * pwvcm), 1.50, 1.80)
    totu = np.zeros((nc, nl + 1))
    totd = np.zeros((nc, nl + 1))
    rec_6, wtdiff = 0.166667, 0.5
    ig0 = 0
    for ib in range(16):
        ng = LW_NGC[ib]
        s = slice(ig0, ig0 + ng)
        urad = np.zeros((nc, nl + 1))
        drad = np.zeros((nc, nl + 1))
        radld = np.zeros((nc, ng))
        atrans = np.zeros((nc, nl, ng))
        bbugas = np.zeros((nc, nl, ng))
        for lev in range(nl, 0, -1):
            plfrac = fracs[:, lev - 1, s]
            blay = sc["planklay"][:, lev - 1, ib][:, None]
            dplankup = sc["planklev"][:, lev, ib][:, None] - blay
            dplankdn = sc["planklev"][:, lev - 1, ib][:, None] - blay
            odepth = np.maximum(secdiff[:, ib][:, None] * taut[:, lev - 1, s], 0.0)
            small = odepth <= 0.06
            tblind = odepth / (BPADE + odepth)
            itr = (TBLINT * tblind + 0.5).astype(np.int64)
            at = np.where(small, odepth - 0.5 * odepth * odepth, 1.0 - exp_tbl[itr])
            tfac = np.where(small, rec_6 * odepth, tfn_tbl[itr])
            bbd = plfrac * (blay + tfac * dplankdn)
            bbugas[:, lev - 1, :] = plfrac * (blay + tfac * dplankup)
            atrans[:, lev - 1, :] = at
            radld = radld + (bbd - radld) * at
            drad[:, lev - 1] += _seqsum(radld)
        rad0 = fracs[:, 0, s] * sc["plankbnd"][:, ib][:, None]
        reflect = 1.0 - semiss[:, ib][:, None]
        radlu = rad0 + reflect * radld
        urad[:, 0] += _seqsum(radlu)
        for lev in range(1, nl + 1):
            radlu = radlu + (bbugas[:, lev - 1, :] - radlu) * atrans[:, lev - 1, :]
            urad[:, lev] += _seqsum(radlu)
        totu += urad * wtdiff * LW_DELWAVE[ib]
        totd += drad * wtdiff * LW_DELWAVE[ib]
        ig0 += ng
    totu *= FLUXFAC
    totd *= FLUXFAC
    fnet = totu - totd
    htr = heatfac(cp_air) * (fnet[:, :-1] - fnet[:, 1:]) / (pz[:, :-1] - pz[:, 1:])
    return totu, totd, htr


def _seqsum(x):
    acc = np.zeros(x.shape[0])
    for i in range(x.shape[1]):
        acc = acc + x[:, i]
    return acc


def rrtmg_lw(play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr=0.0, n2ovmr=0.0, o2vmr=0.0,
             cfc11vmr=0.0, cfc12vmr=0.0, cfc22vmr=0.0, ccl4vmr=0.0, emis=1.0, cp_air=287.04 / (2.0 / 7.0),
             return_optics=False):
    """rrtmg_lw (clear sky, `icld = 0`, `idrv = 0`): -> uflx, dflx [ncol, nlay+1] (W/m2), hr [ncol, nlay] (K/day)"""
    full = lambda v: np.broadcast_to(np.asarray(v, dtype=float), play.shape).copy()
    h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr = map(full, (h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr))
    cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr = map(full, (cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr))
    nc = play.shape[0]
    semiss = np.broadcast_to(np.asarray(emis, dtype=float), (nc, 16)) if np.ndim(emis) < 2 else emis
    coldry, wkl, wbrodl, wx, pwvcm = lw_inatm(plev, h2ovmr, co2vmr, o3vmr, n2ovmr, ch4vmr, o2vmr,
                                              ccl4vmr, cfc11vmr, cfc12vmr, cfc22vmr)
    sc = lw_setcoef(play, tlay, tlev, tsfc, semiss, coldry, wkl, wbrodl)
    taug, fracs = lw_taumol(play, wx, coldry, sc)
    uflx, dflx, hr = lw_rtrnmr_clear(plev, semiss, sc, pwvcm, fracs, taug, cp_air)
    if return_optics:
        return uflx, dflx, hr, dict(taug=taug, fracs=fracs, sc=sc, pwvcm=pwvcm, coldry=coldry)
    return uflx, dflx, hr


# ----------------------------------------------------------------------------------------------------------------
# shortwave
# ----------------------------------------------------------------------------------------------------------------
def sw_setcoef(pavel, tavel, coldry, wkl):
    """setcoef_sw (rrtmg_sw_setcoef.f90); wkl = [h2o, co2, o3, n2o, co(=0), ch4, o2] column amounts"""
    T = tables()
    plog, jp, jt, jt1, fac00, fac01, fac10, fac11 = _pt_indices(pavel, tavel, T["sw_preflog"], T["sw_tref"])
    lower = plog > 4.56
    laytrop = lower.sum(axis=1)
    water = wkl[0] / coldry
    scalefac = pavel * STPFAC / tavel
    forfac = scalefac / (1.0 + water)
    factor = (332.0 - tavel) / 36.0
    indfor_l = np.minimum(2, np.maximum(1, factor.astype(np.int64)))
    forfrac_l = factor - indfor_l
    indfor = np.where(lower, indfor_l, 3)
    forfrac = np.where(lower, forfrac_l, (tavel - 188.0) / 36.0 - 1.0)
    selffac = np.where(lower, water * forfac, 0.0)
    factor = (tavel - 188.0) / 7.2
    indself = np.minimum(9, np.maximum(1, np.trunc(factor).astype(np.int64) - 7))
    selffrac = np.where(lower, factor - (indself + 7), 0.0)
    col = {}
    for name, i in (("h2o", 0), ("co2", 1), ("o3", 2), ("n2o", 3), ("ch4", 5), ("o2", 6)):
        col[name] = 1.0e-20 * wkl[i]
    colmol = 1.0e-20 * coldry + col["h2o"]
    for name in ("co2", "n2o", "ch4", "o2"):
        col[name] = np.where(col[name] == 0.0, 1.0e-32 * coldry, col[name])
    return dict(laytrop=laytrop, lower=lower, jp=jp, jt=jt, jt1=jt1, fac00=fac00, fac01=fac01, fac10=fac10,
                fac11=fac11, forfac=forfac, forfrac=forfrac, indfor=indfor, selffac=selffac, selffrac=selffrac,
                indself=indself, col=col, colmol=colmol)


# per band: (lower key species, strrat, upper key species, layreffr, laysolfr search in the lower (True) / upper part)
def sw_taumol(sc):
    """taumol_sw: taug, taur [ncol, nlay, 112] and the solar source sfluxzen [ncol, 112]"""
    T = tables()
    nc, nl = sc["jp"].shape
    ngpt = sum(SW_NGC)
    ngs = np.concatenate([[0], np.cumsum(SW_NGC)])
    taug = np.zeros((nc, nl, ngpt))
    taur = np.zeros((nc, nl, ngpt))
    sflux = np.zeros((nc, ngpt))
    col, colmol = sc["col"], sc["colmol"]
    laytrop = sc["laytrop"]
    jpall = np.concatenate([sc["jp"], np.zeros((nc, 1), dtype=np.int64)], axis=1)
    c = lambda x: x[:, None]

    def laysolfr_lower(layreffr):
        """`laysolfr = laytrop; do lay = 1, laytrop: if (jp(lay) < layreffr .and. jp(lay+1) >= layreffr)
        laysolfr = min(lay+1, laytrop)` (1-based result)"""
        ls = laytrop.copy()
        for lay in range(1, nl + 1):
            act = (lay <= laytrop) & (jpall[:, lay - 1] < layreffr) & (jpall[:, lay] >= layreffr)
            ls = np.where(act, np.minimum(lay + 1, laytrop), ls)
        return ls

    def laysolfr_upper(layreffr):
        """`laysolfr = nlayers; do lay = laytrop+1, nlayers: if (jp(lay-1) < layreffr .and. jp(lay) >= layreffr)
        laysolfr = lay`; the source is set when `lay == laysolfr` is met during the same loop"""
        ls = np.full(nc, nl, dtype=np.int64)
        for lay in range(2, nl + 1):
            act = (lay > laytrop) & (jpall[:, lay - 2] < layreffr) & (jpall[:, lay - 1] >= layreffr)
            ls = np.where(act, lay, ls)
        return ls

    solfr = {16: laysolfr_upper(18), 17: laysolfr_upper(30), 18: laysolfr_lower(6), 19: laysolfr_lower(3),
             20: laysolfr_lower(3), 21: laysolfr_lower(8), 22: laysolfr_lower(2), 23: laysolfr_lower(6),
             24: laysolfr_lower(1), 25: laysolfr_lower(2), 26: laytrop.copy(), 27: laysolfr_upper(32),
             28: laysolfr_upper(58), 29: laysolfr_upper(49)}

    for lay in range(nl):
        low = sc["lower"][:, lay]
        lowc = low[:, None]
        jp, jt, jt1 = sc["jp"][:, lay], sc["jt"][:, lay], sc["jt1"][:, lay]
        f00, f01, f10, f11 = sc["fac00"][:, lay], sc["fac01"][:, lay], sc["fac10"][:, lay], sc["fac11"][:, lay]
        inds, indf = sc["indself"][:, lay], sc["indfor"][:, lay]
        selffac, selffrac = sc["selffac"][:, lay], sc["selffrac"][:, lay]
        forfac, forfrac = sc["forfac"][:, lay], sc["forfrac"][:, lay]
        h2o, co2, o3, ch4, o2 = (col[k][:, lay] for k in ("h2o", "co2", "o3", "ch4", "o2"))
        cm = colmol[:, lay]

        def self_for(b):
            return (c(selffac) * _itab(T["sw%02d_selfref" % b], inds, selffrac)
                    + c(forfac) * _itab(T["sw%02d_forref" % b], indf, forfrac))

        def for_only(b):
            return c(forfac) * _itab(T["sw%02d_forref" % b], indf, forfrac)

        def single_lower(b):
            i0 = ((jp - 1) * 5 + (jt - 1)) + 1
            i1 = (jp * 5 + (jt1 - 1)) + 1
            return _single(_absview(T["sw%02d_ka" % b]), i0, i1, f00, f10, f01, f11)

        def single_upper(b):
            i0 = ((jp - 13) * 5 + (jt - 1)) + 1
            i1 = ((jp - 12) * 5 + (jt1 - 1)) + 1
            return _single(_absview(T["sw%02d_kb" % b]), i0, i1, f00, f10, f01, f11)

        def binary(b, ca, cb, strrat, lower_part):
            mult, nsp = (8.0, 9) if lower_part else (4.0, 5)
            speccomb, _, js, fs = _specparm(ca, cb, strrat, mult)
            if lower_part:
                i0 = ((jp - 1) * 5 + (jt - 1)) * nsp + js
                i1 = (jp * 5 + (jt1 - 1)) * nsp + js
                ab = _absview(T["sw%02d_ka" % b])
            else:
                i0 = ((jp - 13) * 5 + (jt - 1)) * nsp + js
                i1 = ((jp - 12) * 5 + (jt1 - 1)) * nsp + js
                ab = _absview(T["sw%02d_kb" % b])
            t = (_major2(ab, i0, fs, f00, f10, np.ones(nc), nsp) + _major2(ab, i1, fs, f01, f11, np.ones(nc), nsp))
            return c(speccomb) * t, js, fs

        def put(b, tl, tu, rl, ru, src_l=None, src_u=None):
            s = slice(ngs[b - 16], ngs[b - 15])
            ng = SW_NGC[b - 16]
            z = np.zeros((nc, ng))
            bl = lambda f: z if f is None else (np.broadcast_to(f, (nc, ng)) if np.ndim(f) < 2 else f)
            taug[:, lay, s] = np.where(lowc, bl(tl), bl(tu))
            taur[:, lay, s] = np.where(lowc, bl(rl), bl(ru))
            here = (solfr[b] == lay + 1)
            src = src_l if src_l is not None else src_u
            # the lower-part bands set the source inside the lower loop, the upper-part bands inside the upper loop
            ok = here & (low if src_l is not None else ~low)
            sflux[:, s] = np.where(c(ok), bl(src), sflux[:, s])

        def src2(b, js, fs):
            fr = T["sw%02d_sfluxref" % b]                       # (ng, nsp)
            j = np.clip(js, 1, fr.shape[1] - 1)
            return fr[:, j - 1].T + c(fs) * (fr[:, j].T - fr[:, j - 1].T)

        rayl = lambda b: c(cm) * T["sw%02d_rayl" % b][None, :]          # scalar (1,) or per-g (ng,) broadcasts
        # band 16: 2600-3250 (h2o, ch4) (ch4)
        tl, js, fs = binary(16, h2o, ch4, 252.131, True)
        put(16, tl + c(h2o) * self_for(16), c(ch4) * single_upper(16), rayl(16), rayl(16),
            src_u=T["sw16_sfluxref"])
        # band 17: 3250-4000 (h2o, co2) (h2o, co2)
        tl, js, fs = binary(17, h2o, co2, 0.364641, True)
        tu, jsu, fsu = binary(17, h2o, co2, 0.364641, False)
        put(17, tl + c(h2o) * self_for(17), tu + c(h2o) * for_only(17), rayl(17), rayl(17), src_u=src2(17, jsu, fsu))
        # band 18: 4000-4650 (h2o, ch4) (ch4)
        tl, js, fs = binary(18, h2o, ch4, 38.9589, True)
        put(18, tl + c(h2o) * self_for(18), c(ch4) * single_upper(18), rayl(18), rayl(18), src_l=src2(18, js, fs))
        # band 19: 4650-5150 (h2o, co2) (co2)
        tl, js, fs = binary(19, h2o, co2, 5.49281, True)
        put(19, tl + c(h2o) * self_for(19), c(co2) * single_upper(19), rayl(19), rayl(19), src_l=src2(19, js, fs))
        # band 20: 5150-6150 (h2o) (h2o) + ch4 continuum
        ach4 = c(ch4) * T["sw20_absch4"][None, :]
        put(20, c(h2o) * (single_lower(20) + self_for(20)) + ach4, c(h2o) * (single_upper(20) + for_only(20)) + ach4,
            rayl(20), rayl(20), src_l=T["sw20_sfluxref"])
        # band 21: 6150-7700 (h2o, co2) (h2o, co2)
        tl, js, fs = binary(21, h2o, co2, 0.0045321, True)
        tu, jsu, fsu = binary(21, h2o, co2, 0.0045321, False)
        put(21, tl + c(h2o) * self_for(21), tu + c(h2o) * for_only(21), rayl(21), rayl(21), src_l=src2(21, js, fs))
        # band 22: 7700-8050 (h2o, o2) (o2)
        o2adj = 1.6
        o2cont = c(4.35e-4 * o2 / (350.0 * 2.0))
        tl, js, fs = binary(22, h2o, o2, o2adj * 0.022708, True)
        put(22, tl + c(h2o) * self_for(22) + o2cont, c(o2 * o2adj) * single_upper(22) + o2cont, rayl(22), rayl(22),
            src_l=src2(22, js, fs))
        # band 23: 8050-12850 (h2o) (nothing)
        put(23, c(h2o) * (1.029 * single_lower(23) + self_for(23)), None, rayl(23), rayl(23), src_l=T["sw23_sfluxref"])
        # band 24: 12850-16000 (h2o, o2) (o2) + o3
        tl, js, fs = binary(24, h2o, o2, 0.124692, True)
        ra = T["sw24_rayla"]
        j = np.clip(js, 1, ra.shape[1] - 1)
        rl = c(cm) * (ra[:, j - 1].T + c(fs) * (ra[:, j].T - ra[:, j - 1].T))
        put(24, tl + c(o3) * T["sw24_abso3a"][None, :] + c(h2o) * self_for(24),
            c(o2) * single_upper(24) + c(o3) * T["sw24_abso3b"][None, :], rl, c(cm) * T["sw24_raylb"][None, :],
            src_l=src2(24, js, fs))
        # band 25: 16000-22650 (h2o) (nothing) + o3
        put(25, c(h2o) * single_lower(25) + c(o3) * T["sw25_abso3a"][None, :], c(o3) * T["sw25_abso3b"][None, :],
            rayl(25), rayl(25), src_l=T["sw25_sfluxref"])
        # band 26: 22650-29000 (nothing)
        put(26, None, None, rayl(26), rayl(26), src_l=T["sw26_sfluxref"])
        # band 27: 29000-38000 (o3) (o3)
        put(27, c(o3) * single_lower(27), c(o3) * single_upper(27), rayl(27), rayl(27),
            src_u=(50.15 / 48.37) * T["sw27_sfluxref"])
        # band 28: 38000-50000 (o3, o2) (o3, o2)
        tl, js, fs = binary(28, o3, o2, 6.67029e-07, True)
        tu, jsu, fsu = binary(28, o3, o2, 6.67029e-07, False)
        put(28, tl, tu, rayl(28), rayl(28), src_u=src2(28, jsu, fsu))
        # band 29: 820-2600 (h2o) (co2) + co2 / h2o continua
        put(29, c(h2o) * (single_lower(29) + self_for(29)) + c(co2) * T["sw29_absco2"][None, :],
            c(co2) * single_upper(29) + c(h2o) * T["sw29_absh2o"][None, :], rayl(29), rayl(29),
            src_u=T["sw29_sfluxref"])
    return taug, taur, sflux


def _sw_exp(ze):
    """exp(-ze) as spcvrt_sw / reftra_sw evaluate it: series below od_lo = 0.06, otherwise the Pade-indexed table"""
    exp_tbl = sw_exp_table()
    tblind = ze / (BPADE + ze)
    itind = (TBLINT * tblind + 0.5).astype(np.int64)
    return np.where(ze <= 0.06, 1.0 - ze + 0.5 * ze * ze, exp_tbl[itind])


def sw_reftra(pgg, prmuz, ptau, pw):
    """reftra_sw for layers with lrtchk true (kmodts = 2: PIFM, Zdunkowski et al.) -> pref, prefd, ptra, ptrad"""
    zwcrit, eps = 0.9999995, 1.0e-08
    zto1, zw, zg = ptau, pw, pgg
    zg3 = 3.0 * zg
    zgamma1 = (8.0 - zw * (5.0 + zg3)) * 0.25
    zgamma2 = 3.0 * (zw * (1.0 - zg)) * 0.25
    zgamma3 = (2.0 - zg3 * prmuz) * 0.25
    zgamma4 = 1.0 - zgamma3
    zwo = zw / (1.0 - (1.0 - zw) * (zg / (1.0 - zg)) ** 2)
    cons = zwo >= zwcrit
    # conservative scattering
    za = zgamma1 * prmuz
    za1 = za - zgamma3
    zgt = zgamma1 * zto1
    ze1 = np.minimum(zto1 / prmuz, 500.0)
    ze2 = _sw_exp(ze1)
    pref_c = (zgt - za1 * (1.0 - ze2)) / (1.0 + zgt)
    ptra_c = 1.0 - pref_c
    prefd_c = zgt / (1.0 + zgt)
    ptrad_c = 1.0 - prefd_c
    one = ze2 == 1.0
    pref_c = np.where(one, 0.0, pref_c)
    ptra_c = np.where(one, 1.0, ptra_c)
    prefd_c = np.where(one, 0.0, prefd_c)
    ptrad_c = np.where(one, 1.0, ptrad_c)
    # non-conservative scattering
    with np.errstate(invalid="ignore", divide="ignore"):
        za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3
        za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4
        zrk = np.sqrt(np.maximum(zgamma1 ** 2 - zgamma2 ** 2, 0.0))
        zrp = zrk * prmuz
        zrp1 = 1.0 + zrp
        zrm1 = 1.0 - zrp
        zrk2 = 2.0 * zrk
        zrpp = 1.0 - zrp * zrp
        zrkg = zrk + zgamma1
        zr1 = zrm1 * (za2 + zrk * zgamma3)
        zr2 = zrp1 * (za2 - zrk * zgamma3)
        zr3 = zrk2 * (zgamma3 - za2 * prmuz)
        zr4 = zrpp * zrkg
        zr5 = zrpp * (zrk - zgamma1)
        zt1 = zrp1 * (za1 + zrk * zgamma4)
        zt2 = zrm1 * (za1 - zrk * zgamma4)
        zt3 = zrk2 * (zgamma4 + za1 * prmuz)
        zt4, zt5 = zr4, zr5
        zbeta = (zgamma1 - zrk) / zrkg
        ze1 = np.minimum(zrk * zto1, 500.0)
        ze2 = np.minimum(zto1 / prmuz, 500.0)
        zem1 = _sw_exp(ze1)
        zep1 = 1.0 / zem1
        zem2 = _sw_exp(ze2)
        zep2 = 1.0 / zem2
        zdenr = zr4 * zep1 + zr5 * zem1
        zdent = zt4 * zep1 + zt5 * zem1
        tiny = (zdenr >= -eps) & (zdenr <= eps)
        pref_n = np.where(tiny, eps, zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr)
        ptra_n = np.where(tiny, zem2, zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent)
        zemm = zem1 * zem1
        zdend = 1.0 / ((1.0 - zbeta * zemm) * zrkg)
        prefd_n = zgamma2 * (1.0 - zemm) * zdend
        ptrad_n = zrk2 * zem1 * zdend
    return (np.where(cons, pref_c, pref_n), np.where(cons, prefd_c, prefd_n),
            np.where(cons, ptra_c, ptra_n), np.where(cons, ptrad_c, ptrad_n))


def sw_vrtqdr(pref, prefd, ptra, ptrad, pdbt, ptdbt, albp, albd):
    """vrtqdr_sw: adding method; inputs [.., klev] ordered top (0) to bottom, ptdbt [.., klev+1]; the surface
    (index klev) reflectances are the direct / diffuse albedos -> pfd, pfu [.., klev+1] (top first)"""
    klev = pref.shape[-1]
    shp = pref.shape[:-1]
    pref = np.concatenate([pref, np.broadcast_to(albp, shp)[..., None]], axis=-1)
    prefd = np.concatenate([prefd, np.broadcast_to(albd, shp)[..., None]], axis=-1)
    prup = np.zeros(shp + (klev + 1,))
    prupd = np.zeros(shp + (klev + 1,))
    prup[..., klev] = pref[..., klev]
    prupd[..., klev] = prefd[..., klev]
    k = klev - 1
    zreflect = 1.0 / (1.0 - prefd[..., klev] * prefd[..., k])
    prup[..., k] = pref[..., k] + (ptrad[..., k] * ((ptra[..., k] - pdbt[..., k]) * prefd[..., klev]
                                                    + pdbt[..., k] * pref[..., klev])) * zreflect
    prupd[..., k] = prefd[..., k] + ptrad[..., k] * ptrad[..., k] * prefd[..., klev] * zreflect
    for ikx in range(klev - 2, -1, -1):
        ikp = ikx + 1
        zreflect = 1.0 / (1.0 - prupd[..., ikp] * prefd[..., ikx])
        prup[..., ikx] = pref[..., ikx] + (ptrad[..., ikx] * ((ptra[..., ikx] - pdbt[..., ikx]) * prupd[..., ikp]
                                                              + pdbt[..., ikx] * prup[..., ikp])) * zreflect
        prupd[..., ikx] = prefd[..., ikx] + ptrad[..., ikx] * ptrad[..., ikx] * prupd[..., ikp] * zreflect
    ztdn = np.zeros(shp + (klev + 1,))
    prdnd = np.zeros(shp + (klev + 1,))
    ztdn[..., 0] = 1.0
    ztdn[..., 1] = ptra[..., 0]
    prdnd[..., 1] = prefd[..., 0]
    for jk in range(1, klev):
        zreflect = 1.0 / (1.0 - prefd[..., jk] * prdnd[..., jk])
        ztdn[..., jk + 1] = ptdbt[..., jk] * ptra[..., jk] + (ptrad[..., jk] * ((ztdn[..., jk] - ptdbt[..., jk])
                                                                                 + ptdbt[..., jk] * pref[..., jk] * prdnd[..., jk])) * zreflect
        prdnd[..., jk + 1] = prefd[..., jk] + ptrad[..., jk] * ptrad[..., jk] * prdnd[..., jk] * zreflect
    zreflect = 1.0 / (1.0 - prdnd * prupd)
    pfu = (ptdbt * prup + (ztdn - ptdbt) * prupd) * zreflect
    pfd = ptdbt + (ztdn - ptdbt + ptdbt * prup * prdnd) * zreflect
    return pfd, pfu


def rrtmg_sw(play, plev, tlay, h2ovmr, o3vmr, co2vmr, ch4vmr=0.0, n2ovmr=0.0, o2vmr=0.0, albedo=0.3, coszen=0.5,
             adjes=1.0, scon=1368.22, cp_air=287.04 / (2.0 / 7.0), return_optics=False):
    """rrtmg_sw (clear sky, no aerosol; `dyofyr = 0` so the Earth-Sun factor is `adjes`): -> swuflx, swdflx
    [ncol, nlay+1] (W/m2, index 0 = surface), swhr [ncol, nlay] (K/day); columns with coszen < 1e-10 are zero"""
    full = lambda v: np.broadcast_to(np.asarray(v, dtype=float), play.shape).copy()
    h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr = map(full, (h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr))
    nc, nl = play.shape
    albedo = np.broadcast_to(np.asarray(albedo, dtype=float), (nc,))
    coszen = np.broadcast_to(np.asarray(coszen, dtype=float), (nc,))
    zepzen = 1.0e-10
    day = coszen >= zepzen
    coldry = _inatm_common(plev, h2ovmr)
    wkl = [coldry * v for v in (h2ovmr, co2vmr, o3vmr, n2ovmr, np.zeros_like(h2ovmr), ch4vmr, o2vmr)]
    sc = sw_setcoef(play, tlay, coldry, wkl)
    taug, taur, sflux = sw_taumol(sc)
    prmu0 = np.where(day, coszen, 1.0)
    adjflux = adjes * (scon / RRSW_SCON)
    ngpt = taug.shape[2]
    # spcvrt_sw, clear sky: layers reordered top (jk = 1) to bottom
    ztauc = (taur + taug)[:, ::-1, :]
    zomcc = taur[:, ::-1, :] / ztauc
    zgcc = np.zeros_like(ztauc)
    mu = prmu0[:, None, None]
    zref, zrefd, ztra, ztrad = sw_reftra(zgcc, mu, ztauc, zomcc)
    zdbt = _sw_exp(ztauc / mu)
    ztdbt = np.ones((nc, nl + 1, ngpt))
    for jk in range(nl):
        ztdbt[:, jk + 1, :] = zdbt[:, jk, :] * ztdbt[:, jk, :]
    mv = lambda a: np.moveaxis(a, 1, -1)                        # -> [ncol, g, level]
    alb = albedo[:, None]
    pfd, pfu = sw_vrtqdr(mv(zref), mv(zrefd), mv(ztra), mv(ztrad), mv(zdbt), mv(ztdbt), alb, alb)
    zincflx = adjflux * sflux * prmu0[:, None]                  # [ncol, g]
    fu = np.zeros((nc, nl + 1))
    fd = np.zeros((nc, nl + 1))
    for ig in range(ngpt):                                      # sequential accumulation over g-points
        fu = fu + zincflx[:, ig][:, None] * pfu[:, ig, :]
        fd = fd + zincflx[:, ig][:, None] * pfd[:, ig, :]
    swuflx = fu[:, ::-1] * day[:, None]                         # index 0 = surface
    swdflx = fd[:, ::-1] * day[:, None]
    net = swdflx - swuflx
    pdp = plev[:, :-1] - plev[:, 1:]
    swhr = (net[:, 1:] - net[:, :-1]) * heatfac(cp_air) / pdp
    swhr[:, nl - 1] = 0.0                                       # rrtmg_sw_rad.nomcica.f90: swhr(iplon,nlayers) = 0
    swhr = swhr * day[:, None]
    if return_optics:
        return swuflx, swdflx, swhr, dict(taug=taug, taur=taur, sflux=sflux, sc=sc)
    return swuflx, swdflx, swhr


# ----------------------------------------------------------------------------------------------------------------
# rrtm_radiation.F90 glue
# ----------------------------------------------------------------------------------------------------------------
def interp_temp(z_full, z_half, t):
    """interp_temp (rrtm_radiation.F90:502-544); arrays [..., k] top-down, z_half has K+1 levels"""
    K = t.shape[-1]
    t_half = np.zeros(t.shape[:-1] + (K + 1,))
    for k in range(1, K):
        dzk2 = 1.0 / (z_full[..., k - 1] - z_full[..., k])
        dzk = (z_half[..., k] - z_full[..., k]) * dzk2
        dzk1 = (z_full[..., k - 1] - z_half[..., k]) * dzk2
        t_half[..., k] = t[..., k] * dzk1 + t[..., k - 1] * dzk
    t_half[..., 0] = 0.5 * (3 * t[..., 0] - t[..., 1])
    t_half[..., K] = t[..., K - 2] + (z_half[..., K] - z_full[..., K - 2]) * (t[..., K - 1] - t[..., K - 2]) \
        / (z_full[..., K - 1] - z_full[..., K - 2])
    return t_half


# shared/constants/constants.F90:84-86,166-169,237
RDGAS, RVGAS, GAS_CONSTANT = 287.04, 461.50, 8.314
WTMAIR, WTMOZONE = 2.896440E+01, 47.99820
WTMH2O = WTMAIR * (RDGAS / RVGAS)


def run_rrtmg_columns(p_full, p_half, t, t_half, q, t_surf, albedo, coszen, o3vmr=0.0, co2ppmv=300.0,
                      h2o_lower_limit=2.0e-7, temp_lower_limit=100.0, temp_upper_limit=370.0, solrad=1.0,
                      solr_cnst=1368.22, ch4=0.0, n2o=0.0, o2=0.0, cfc11=0.0, cfc12=0.0, cfc22=0.0, ccl4=0.0,
                      cp_air=287.04 / (2.0 / 7.0), gas_constant=GAS_CONSTANT, rdgas=RDGAS, wtmh2o=WTMH2O):
    """The column part of run_rrtmg (rrtm_radiation.F90:816-950): inputs [ncol, K] top-down in Pa / K / kg kg-1,
    o3vmr top-down; -> tdt_rad [ncol, K] (K/s, top-down), flux_sw = net surface SW down, flux_lw = surface LW down,
    olr, toa_sw (net down)"""
    h2o_vmr = (q / (1.0 - q)) * (1000.0 * gas_constant / rdgas) / wtmh2o
    rev = lambda a: a[:, ::-1]
    pfull = rev(p_full) * 0.01
    phalf = rev(p_half) * 0.01
    K = p_full.shape[1]
    top = phalf[:, K]
    if top.min() <= 0.0:
        phalf = phalf.copy()
        phalf[:, K] = pfull[:, K - 1] * 0.5
    tfull = np.clip(rev(t), temp_lower_limit, temp_upper_limit)
    thalf = np.clip(rev(t_half), temp_lower_limit, temp_upper_limit)
    h2o = np.maximum(rev(h2o_vmr), h2o_lower_limit)
    o3 = rev(np.broadcast_to(np.asarray(o3vmr, dtype=float), p_full.shape))
    co2 = co2ppmv * 1.0e-6
    swu, swd, swhr = rrtmg_sw(pfull, phalf, tfull, h2o, o3, co2, ch4, n2o, o2, albedo, coszen, solrad, solr_cnst, cp_air)
    lwu, lwd, lwhr = rrtmg_lw(pfull, phalf, tfull, thalf, t_surf, h2o, o3, co2, ch4, n2o, o2, cfc11, cfc12, cfc22,
                              ccl4, 1.0, cp_air)
    daypersec = 1.0 / 86400.0
    tdt = rev(swhr) * daypersec + rev(lwhr) * daypersec
    return dict(tdt_rad=tdt, tdt_sw=rev(swhr) * daypersec, tdt_lw=rev(lwhr) * daypersec,
                flux_sw=swd[:, 0] - swu[:, 0], flux_lw=lwd[:, 0], olr=lwu[:, K] - lwd[:, K],
                toa_sw=swd[:, K] - swu[:, K], swu=swu, swd=swd, lwu=lwu, lwd=lwd)


# ----------------------------------------------------------------------------------------------------------------
# astronomy_mod (shared/astronomy/astronomy.f90) as run_rrtmg uses it, and the radiation time stepping of run_rrtmg
# ----------------------------------------------------------------------------------------------------------------
class Astronomy:
    """astronomy_nml defaults of Isca (ecc = 0, obliq = 23.439, per = 102.932, num_angles = 3600) and the orbit table of
    `orbit` (astronomy.f90: 4th-order Runge-Kutta of d(angle)/dt = r_inv_squared)"""

    def __init__(self, ecc=0.0, obliq=23.439, per=102.932, num_angles=3600):
        self.ecc, self.obliq, self.per, self.num_angles = ecc, obliq, per, num_angles
        self.deg_to_rad = np.pi / 180.0
        orb = np.zeros(num_angles + 1)
        dt = 2 * np.pi / float(num_angles)
        dt = dt * np.sqrt(1.0 - ecc ** 2)
        for n in range(1, num_angles + 1):
            d1 = dt * self.r_inv_squared(orb[n - 1])
            d2 = dt * self.r_inv_squared(orb[n - 1] + 0.5 * d1)
            d3 = dt * self.r_inv_squared(orb[n - 1] + 0.5 * d2)
            d4 = dt * self.r_inv_squared(orb[n - 1] + d3)
            orb[n] = orb[n - 1] + (d1 / 6.0 + d2 / 3.0 + d3 / 3.0 + d4 / 6.0)
        self.orb_angle = orb

    def r_inv_squared(self, ang):
        r = (1.0 - self.ecc ** 2) / (1.0 + self.ecc * np.cos(ang - self.per * self.deg_to_rad))
        return r ** (-2)

    def angle(self, t):
        norm_time = t * float(self.num_angles) / (2 * np.pi)
        i = int(np.floor(norm_time)) % self.num_angles
        x = norm_time - np.floor(norm_time)
        return ((1.0 - x) * self.orb_angle[i] + x * self.orb_angle[i + 1]) % (2 * np.pi)

    def declination(self, ang):
        return np.arcsin(-np.sin(self.obliq * self.deg_to_rad) * np.sin(ang))

    def half_day(self, lat, dec):
        eps = 1.0e-05
        lat = np.where(lat == 0.5 * np.pi, lat - eps, lat)
        lat = np.where(lat == -0.5 * np.pi, lat + eps, lat)
        c = -np.tan(lat) * np.tan(dec)
        with np.errstate(invalid="ignore"):
            return np.where(c <= -1.0, np.pi, np.where(c >= 1.0, 0.0, np.arccos(np.clip(c, -1.0, 1.0))))

    def diurnal_solar(self, lat, lon, gmt, time_since_ae, dt=None):
        """diurnal_solar_2d (astronomy.f90:1123-1410), allow_negative_cosz absent -> cosz, fracday, rrsun"""
        twopi = 2 * np.pi
        ang = self.angle(time_since_ae)
        dec = self.declination(ang)
        rrsun = self.r_inv_squared(ang)
        aa = np.sin(lat) * np.sin(dec)
        bb = np.cos(lat) * np.cos(dec)
        t = gmt + lon - np.pi
        t = np.where(t >= np.pi, t - twopi, t)
        t = np.where(t < -np.pi, t + twopi, t)
        h = self.half_day(lat, dec)
        if dt is None:
            day = np.abs(t) < h
            cosz = np.where(day, aa + bb * np.cos(t), 0.0)
            return np.maximum(0.0, cosz), day.astype(float), rrsun
        tt = t + dt
        st, stt, sh = np.sin(t), np.sin(tt), np.sin(h)
        cosz = np.zeros_like(t)
        with np.errstate(divide="ignore", invalid="ignore"):
            w = lambda m, v: np.where(m, v, cosz)
            cosz = w((t < -h) & (tt < -h), 0.0)
            cosz = w((t < -h) & (np.abs(tt) <= h), (aa * (tt + h) / (tt - t)) + bb * (stt + sh) / (tt - t))
            cosz = w((t < -h) & (h != 0.0) & (h < tt), aa * (2. * h) / (tt - t) + bb * (sh + sh) / (tt - t))
            cosz = w((np.abs(t) <= h) & (np.abs(tt) <= h), aa + bb * (stt - st) / (tt - t))
            cosz = w((np.abs(t) <= h) & (h < tt), (aa * (h - t) / (tt - t)) + bb * (sh - st) / (tt - t))
            cosz = w((twopi - h < tt) & (t <= h), aa * ((tt + (2. * h) - t - twopi) / (tt - t)) + bb * (((sh - st) / (tt - t)) + ((stt + sh) / (tt - t))))
            cosz = w((h < t) & (twopi - h >= tt), 0.0)
            cosz = w((h < t) & (twopi - h < tt) & (tt < twopi + h), aa * (tt + h - twopi) / (tt - t) + bb * (stt + sh) / (tt - t))
            cosz = w((h < t) & (twopi - h < tt) & (tt > twopi + h), aa * (2. * h) / (tt - t) + bb * (sh + sh) / (tt - t))
        fracday = np.zeros_like(t)
        f = lambda m, v: np.where(m, v, fracday)
        fracday = f((t < -h) & (tt < -h), 0.0)
        fracday = f((t < -h) & (np.abs(tt) <= h), (tt + h) / dt)
        fracday = f((t < -h) & (h < tt), (h + h) / dt)
        fracday = f((np.abs(t) <= h) & (np.abs(tt) <= h), (tt - t) / dt)
        fracday = f((np.abs(t) <= h) & (h < tt), (h - t) / dt)
        fracday = f(h < t, 0.0)
        fracday = np.where(twopi - h < tt, fracday + (tt + h - twopi) / dt, fracday)
        return np.maximum(0.0, cosz), fracday, rrsun


class RrtmRadiation:
    """run_rrtmg (rrtm_radiation.F90:547-1056) as idealized_moist_phys calls it every step: radiation alarm (`dt_rad`,
    `store_intermediate_rad`), zenith angle (`do_rad_time_avg`, `solday`, `equinox_day`, `frierson_solar_rad`), the column
    computation on a radiation step.  Arrays [K, J, I]; o3 as read from the ozone file (mmr when input_o3_file_is_mmr)."""

    def __init__(self, lat2d, lon2d, dt_atmos, dt_rad=0, dt_rad_avg=-1, do_rad_time_avg=True, store_intermediate_rad=True, solday=0,
                 equinox_day=0.75, frierson_solar_rad=False, del_sol=0.95, del_sw=0.0, o3=None, input_o3_file_is_mmr=True,
                 day_in_s=86400.0, year_in_s=360 * 86400, astronomy=None, lonstep=1, **column_kw):
        self.lat, self.lon = lat2d, lon2d
        self.dt_rad = int(dt_rad) if dt_rad > 0 else int(dt_atmos)
        self.dt_rad_avg = self.dt_rad if dt_rad_avg <= 0 else dt_rad_avg
        self.do_rad_time_avg, self.store = do_rad_time_avg, store_intermediate_rad
        self.solday, self.equinox_day = solday, equinox_day
        self.frierson, self.del_sol, self.del_sw = frierson_solar_rad, del_sol, del_sw
        self.day_in_s, self.year_in_s = day_in_s, year_in_s
        self.astro = astronomy or Astronomy()
        self.o3, self.o3_mmr = o3, input_o3_file_is_mmr
        self.kw = column_kw
        self.lonstep = int(lonstep)
        self.dt_last = -float(self.dt_rad)
        self.tdt_rad = self.sw_flux = self.lw_flux = None
        self.coszen = None
        self.n_rad_calls = 0

    def zenith(self, total_seconds):
        if self.frierson:
            p2 = (1.0 - 3.0 * np.sin(self.lat) ** 2) / 4.0
            return 0.25 * (1.0 + self.del_sol * p2 + self.del_sw * np.sin(self.lat))
        if self.solday > 0:                           # Time_loc = set_time(seconds, solday)
            total_seconds = (total_seconds % 86400.0) + self.solday * 86400.0
        frac_of_day = total_seconds / self.day_in_s
        frac_of_year = (self.solday * self.day_in_s) / self.year_in_s if self.solday > 0 else total_seconds / self.year_in_s
        gmt = abs(np.fmod(frac_of_day, 1.0)) * 2.0 * np.pi
        time_since_ae = ((frac_of_year - self.equinox_day) % 1.0) * 2.0 * np.pi
        dt = (self.dt_rad_avg / self.day_in_s) * 2.0 * np.pi if self.do_rad_time_avg else None
        return self.astro.diurnal_solar(self.lat, self.lon, gmt, time_since_ae, dt)[0]

    def __call__(self, total_seconds, p_full, p_half, z_full, z_half, t, q, t_surf, albedo, tdt):
        """-> tdt + radiative heating, flux_sw, flux_lw"""
        if total_seconds - self.dt_last >= self.dt_rad:
            self.dt_last = total_seconds
        else:
            if self.store:
                return tdt + self.tdt_rad, self.sw_flux, self.lw_flux
            z2 = np.zeros_like(t_surf)
            return tdt, z2, z2
        K, J, I = t.shape
        self.coszen = self.zenith(total_seconds)
        ls = self.lonstep
        sub = lambda a: a[..., ::ls]                  # `p_full(1:si:lonstep,:,:)` (rrtm_radiation.F90:831-846)
        col = lambda a: a.reshape(a.shape[0], -1).T
        th = interp_temp(col(sub(z_full)), col(sub(z_half)), col(sub(t)))
        o3v = 0.0
        if self.o3 is not None:
            o3v = col(sub(self.o3)) * ((1000.0 * GAS_CONSTANT / RDGAS) / WTMOZONE if self.o3_mmr else 1.0)
        o = run_rrtmg_columns(col(sub(p_full)), col(sub(p_half)), col(sub(t)), th, col(sub(q)), sub(t_surf).ravel(), sub(albedo).ravel(),
                              sub(self.coszen).ravel(), o3vmr=o3v, **self.kw)
        self.n_rad_calls += 1
        Is = I // ls

        def back(a2):
            """[.., J, I/lonstep] -> [.., J, I]: `di*x(i1) + (1-di)*x(i)`, closed toroidally (rrtm_radiation.F90:918-935)"""
            if ls == 1:
                return a2
            out = np.empty(a2.shape[:-1] + (I,))
            nxt = np.roll(a2, -1, axis=-1)
            for ij in range(ls):
                di = ij * (1.0 / ls)
                out[..., ij::ls] = di * nxt + (1.0 - di) * a2
            return out
        sw3 = (o["tdt_sw"].T.reshape(K, J, Is), o["tdt_lw"].T.reshape(K, J, Is))
        self.tdt_rad = back(sw3[0] + sw3[1]) if ls > 1 else o["tdt_rad"].T.reshape(K, J, I)
        self.sw_flux = back(o["flux_sw"].reshape(J, Is))
        self.lw_flux = back(o["flux_lw"].reshape(J, Is))
        self.olr, self.toa_sw = back(o["olr"].reshape(J, Is)), back(o["toa_sw"].reshape(J, Is))
        return tdt + self.tdt_rad, self.sw_flux, self.lw_flux
