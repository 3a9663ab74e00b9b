#!/usr/bin/env python
"""Build the RRTMG coefficient-table file the radiation kernels load (`isca_b200/data/rrtmg_tables.bin`).

Run HERE (where /root/reference exists); the output is committed because the GPU box has no reference tree.

What it does (all citations relative to /root/reference/src/atmos_param/rrtm_radiation):

1. reads the *data* of the reference -- the `(/ ... /)` array constructors of
   `rrtmg_lw/gcm_model/src/rrtmg_lw_k_g.f90`, `rrtmg_sw/gcm_model/src/rrtmg_sw_k_g.f90` (absorption coefficients on the
   16 original g-points per band), and the reference profiles / Planck tables of `rrtmg_lw_setcoef.f90:lwatmref,
   lwavplank` and `rrtmg_sw_setcoef.f90:swatmref` -- with a small Fortran-constructor parser (no code is copied, the
   decimal literals are converted with Python's correctly rounded `float`, which is what a Fortran compiler does
   for `_rb` = double literals);
2. performs the g-point reduction of `rrtmg_lw_init.f90:rrtmg_lw_ini` (:170-207 relative weights `rwgt`, `cmbgb1..16`)
   and `rrtmg_sw_init.f90:rrtmg_sw_ini` (`cmbgb16s..29`): every `k`-like table is the `rwgt`-weighted sum of the
   original g-points of a group, the Planck fractions (`fracref*`) and solar source functions (`sfluxref*`) are plain
   sums, in the reference's summation order (sequential over the group);
3. writes one flat little-endian file: magic `ISCARRTM`, int32 version, int32 n_entries, then per entry
   name[48], int32 ndim, int32 dims[6] (Fortran order, first index fastest), int64 offset (in doubles), and the
   float64 payload.  Names: `lwNN_<name>` / `swNN_<name>` for band tables (reduced names of the reference modules:
   ka, kb, selfref, forref, fracrefa, ka_mn2, ...), `lw_pref`, `lw_chi_mls`, `lw_totplnk`, ... for the shared ones.

The oracle (`oracle/rrtmg.py`) and the CUDA library read the same file.
"""
import os
import re
import struct
import sys
from collections import OrderedDict

import numpy as np

REF = "/root/reference/src/atmos_param/rrtm_radiation"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "isca_b200", "data", "rrtmg_tables.bin")

# rrtmg_lw_init.f90:lwcmbdat / rrtmg_sw_init.f90:swcmbdat -- number of reduced g-points per band and the number of
# original g-points combined into each reduced one (the reference's ngc / ngn; ngm and ngs follow from them).
LW_NGC = [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2]
LW_NGN = [[1, 1, 2, 2, 2, 2, 2, 2, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2], [1] * 16,
          [1] * 13 + [3], [1] * 16, [2] * 8, [2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2], [2] * 8,
          [1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2], [2, 2, 2, 2, 4, 4], [1, 1, 2, 2, 2, 2, 3, 3],
          [1, 1, 1, 1, 2, 2, 4, 4], [3, 3, 4, 6], [8, 8], [8, 8], [4, 12]]
SW_NGC = [6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12]
SW_NGN = [[2, 2, 2, 2, 4, 4], [1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2], [1, 1, 1, 1, 2, 2, 4, 4], [1, 1, 1, 1, 2, 2, 4, 4],
          [1, 1, 1, 1, 1, 1, 1, 1, 2, 6], [1, 1, 1, 1, 1, 1, 1, 1, 2, 6], [8, 8], [2, 2, 1, 1, 1, 1, 1, 1, 2, 4],
          [2] * 8, [1, 1, 2, 2, 4, 6], [1, 1, 2, 2, 4, 6], [1, 1, 1, 1, 1, 1, 4, 6], [1, 1, 2, 2, 4, 6],
          [1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1]]
# Gaussian weights of the 16 original g-points (lwcmbdat / swcmbdat `wt`)
WT = [0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893, 0.0832767040,
      0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086, 0.0022199750, 0.0014140010,
      0.0005330000, 0.0000750000]


def strip_comments(text):
    out = []
    for line in text.splitlines():
        i = line.find("!")
        if i >= 0:
            line = line[:i]
        out.append(line.rstrip())
    return out


def join_continuations(lines):
    """Fortran free-form continuation: a trailing '&' joins the next line."""
    stmts, cur = [], ""
    for line in lines:
        s = line.strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:]
        if s.endswith("&"):
            cur += s[:-1] + " "
        else:
            stmts.append(cur + s)
            cur = ""
    if cur:
        stmts.append(cur)
    return stmts


def fnum(tok):
    tok = tok.strip().lower().replace("_rb", "").replace("d", "e")
    return float(tok)


def split_top(s):
    """split on commas that are not inside parentheses"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def parse_module_decls(path):
    """-> {name: [(lo, hi) per dim]}, g-axis index per name (dimension spelled noNN), and scalar names"""
    stmts = join_continuations(strip_comments(open(path).read()))
    params, decls, gaxis, scalars = {}, {}, {}, set()
    for s in stmts:
        m = re.match(r"integer\(kind=im\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\d+)", s, re.I)
        if m:
            params[m.group(1).lower()] = int(m.group(2))
            continue
        m = re.match(r"real\(kind=rb\)\s*,\s*dimension\((.*?)\)\s*::\s*(.*)$", s, re.I)
        if m:                                     # `real(kind=rb), dimension(no6) :: fracrefao` style
            s = "real(kind=rb) :: " + ", ".join("%s(%s)" % (n.strip(), m.group(1)) for n in m.group(2).split(","))
        m = re.match(r"real\(kind=rb\)\s*::\s*(.*)$", s, re.I)
        if not m:
            continue
        for item in split_top(m.group(1)):
            item = item.strip()
            mm = re.match(r"(\w+)\s*\((.*)\)$", item)
            if not mm:
                scalars.add(item.lower())
                continue
            name = mm.group(1).lower()
            dims = []
            if re.search(r"\bng\d+\b", mm.group(2), re.I):
                continue                      # the reduced arrays of the module: produced here, not parsed
            for ax, d in enumerate(mm.group(2).split(",")):
                d = d.strip().lower()
                if re.match(r"no\d+$", d):
                    gaxis[name] = ax
                if ":" in d:
                    lo, hi = d.split(":")
                    dims.append((int(lo), int(params.get(hi, hi)) if not hi.isdigit() else int(hi)))
                else:
                    dims.append((1, params[d] if d in params else int(d)))
            decls[name] = dims
    return decls, gaxis, scalars


def parse_assignments(stmts, decls, scalars=()):
    """Array-constructor assignments `name(sec) = (/ ... /)` -> {name: ndarray in Fortran order}."""
    arrays = {}
    for s in stmts:
        m = re.match(r"(\w+)\s*\(([^)]*)\)\s*=\s*\(/(.*)/\)\s*$", s)
        if m:
            name = m.group(1).lower()
            if name not in decls:
                continue
            dims = decls[name]
            if name not in arrays:
                arrays[name] = np.full([hi - lo + 1 for lo, hi in dims], np.nan, order="F")
            vals = np.array([fnum(t) for t in m.group(3).split(",")])
            idx = []
            for (lo, hi), sec in zip(dims, m.group(2).split(",")):
                sec = sec.strip()
                if sec == ":":
                    idx.append(slice(None))
                elif ":" in sec:
                    a, b = sec.split(":")
                    idx.append(slice(int(a) - lo, int(b) - lo + 1))
                else:
                    idx.append(int(sec) - lo)
            target = arrays[name][tuple(idx)]
            assert target.size == vals.size, (name, m.group(2), target.shape, vals.size)
            arrays[name][tuple(idx)] = vals.reshape(target.shape, order="F")
            continue
        m = re.match(r"(\w+)\s*=\s*([-+0-9.eEdD]+(_rb)?)\s*$", s)
        if m and m.group(1).lower() in scalars:
            arrays[m.group(1).lower()] = np.array(fnum(m.group(2)))
    for k, v in arrays.items():
        assert not np.isnan(v).any(), "unfilled entries in " + k
    return arrays


def split_subroutines(path):
    stmts = join_continuations(strip_comments(open(path).read()))
    subs, cur, name = OrderedDict(), None, None
    for s in stmts:
        m = re.match(r"subroutine\s+(\w+)", s, re.I)
        if m:
            name, cur = m.group(1).lower(), []
            continue
        if re.match(r"end\s+subroutine", s, re.I):
            if name:
                subs[name] = cur
            name, cur = None, None
            continue
        if cur is not None:
            cur.append(s)
    return subs


def rwgt_for_band(ngn):
    """rrtmg_lw_init.f90:170-207: rwgt(ig) = wt(ig) / sum of wt over the group ig belongs to (1 when ngc == 16)."""
    if len(ngn) == 16:
        return [1.0] * 16
    r, ig = [], 0
    for n in ngn:
        wtsum = 0.0
        for k in range(n):
            wtsum = wtsum + WT[ig + k]
        for k in range(n):
            r.append(WT[ig + k] / wtsum)
        ig += n
    return r


def reduce_g(arr, ax, ngn, rwgt, plain):
    """cmbgbNN: combine the original g-points along axis `ax`, sequential sums in the reference's order"""
    arr = np.moveaxis(arr, ax, -1)
    out = np.zeros(arr.shape[:-1] + (len(ngn),))
    ig = 0
    for igc, n in enumerate(ngn):
        acc = np.zeros(arr.shape[:-1])
        for k in range(n):
            acc = acc + (arr[..., ig + k] if plain else arr[..., ig + k] * rwgt[ig + k])
        out[..., igc] = acc
        ig += n
    assert ig == 16
    return np.asfortranarray(np.moveaxis(out, -1, ax))


def reduced_name(name):
    # kao -> ka, kbo_mn2o -> kb_mn2o, selfrefo -> selfref, fracrefao -> fracrefa, ccl4o -> ccl4, cfc11adjo -> cfc11adj
    m = re.match(r"(k[ab])o(_\w+)?$", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    assert name.endswith("o"), name
    return name[:-1]


def build_family(prefix, kg_path, mod_fmt, bands, ngn_all):
    out = OrderedDict()
    subs = split_subroutines(kg_path)
    for ib, band in enumerate(bands):
        decls, gaxis, scalars = parse_module_decls(mod_fmt % band)
        arrays = parse_assignments(subs["%s_kgb%02d" % (prefix, band)], decls, scalars)
        ngn = ngn_all[ib]
        rw = rwgt_for_band(ngn)
        for name, arr in arrays.items():
            if arr.ndim == 0:
                out["%s%02d_%s" % (prefix, band, name)] = arr.reshape(1)
                continue
            if name not in gaxis:
                raise RuntimeError("no g axis for %s band %d" % (name, band))
            plain = name.startswith("fracref") or name.startswith("sfluxref")
            red = reduce_g(arr, gaxis[name], ngn, rw, plain)
            out["%s%02d_%s" % (prefix, band, reduced_name(name))] = red
    return out


def build_shared():
    out = OrderedDict()
    subs = split_subroutines(REF + "/rrtmg_lw/gcm_model/src/rrtmg_lw_setcoef.f90")
    decls = {"pref": [(1, 59)], "preflog": [(1, 59)], "tref": [(1, 59)], "chi_mls": [(1, 7), (1, 59)]}
    a = parse_assignments(subs["lwatmref"], decls)
    for k in ("pref", "preflog", "tref", "chi_mls"):
        out["lw_" + k] = a[k]
    decls = {"totplnk": [(1, 181), (1, 16)], "totplk16": [(1, 181)]}
    a = parse_assignments(subs["lwavplank"], decls)
    out["lw_totplnk"] = a["totplnk"]
    out["lw_totplk16"] = a["totplk16"]
    subs = split_subroutines(REF + "/rrtmg_sw/gcm_model/src/rrtmg_sw_setcoef.f90")
    decls = {"pref": [(1, 59)], "preflog": [(1, 59)], "tref": [(1, 59)]}
    a = parse_assignments(subs["swatmref"], decls)
    for k in ("pref", "preflog", "tref"):
        out["sw_" + k] = a[k]
    out["lw_ngc"] = np.array(LW_NGC, dtype=float)
    out["sw_ngc"] = np.array(SW_NGC, dtype=float)
    return out


def write_tables(tables, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    n = len(tables)
    hdr = struct.pack("<8sii", b"ISCARRTM", 1, n)
    entry_sz = 48 + 4 + 6 * 4 + 8
    off = 0
    entries, payload = [], []
    for name, arr in tables.items():
        arr = np.asarray(arr, dtype="<f8")
        dims = list(arr.shape) + [1] * (6 - arr.ndim)
        entries.append(struct.pack("<48si6iq", name.encode(), arr.ndim, *dims, off))
        flat = arr.ravel(order="F")
        payload.append(flat.tobytes())
        off += flat.size
    with open(path, "wb") as f:
        f.write(hdr)
        for e in entries:
            assert len(e) == entry_sz
            f.write(e)
        for p in payload:
            f.write(p)
    return off


def main():
    tables = OrderedDict()
    tables.update(build_shared())
    tables.update(build_family("lw", REF + "/rrtmg_lw/gcm_model/src/rrtmg_lw_k_g.f90",
                               REF + "/rrtmg_lw/gcm_model/modules/rrlw_kg%02d.f90", range(1, 17), LW_NGN))
    tables.update(build_family("sw", REF + "/rrtmg_sw/gcm_model/src/rrtmg_sw_k_g.f90",
                               REF + "/rrtmg_sw/gcm_model/modules/rrsw_kg%02d.f90", range(16, 30), SW_NGN))
    out = sys.argv[1] if len(sys.argv) > 1 else OUT
    nd = write_tables(tables, out)
    print("%d tables, %d doubles (%.1f MB) -> %s" % (len(tables), nd, nd * 8 / 1e6, os.path.normpath(out)))
    for k, v in tables.items():
        print("  %-22s %s" % (k, tuple(v.shape)))


if __name__ == "__main__":
    main()
