// TEST INFRASTRUCTURE ONLY -- serial host build of the full Betts-Miller column code of isca_b200/csrc/physics_bm_column.h.
// physics_bm.cu runs bm_column from one CUDA thread per column; this file runs the very same function in a plain loop so that
// `pytest -m "not gpu"` can check it on a machine without a GPU.  Compiled by tests/test_betts_miller.py with g++ into
// tests/host/_build/; never linked into the product library.
#include "../../isca_b200/csrc/physics_bm_column.h"
#include <vector>

using namespace isca_bm;

static const double LCL[127] = {ISCA_BM_LCLTABLE_VALUES};

extern "C" {

const double* bm_host_lcltable() { return LCL; }

// svp: TABLE | DTABLE | D2TABLE (n each) and (tminl, dtinvl, tepsl, dtres); cfg: tau_bm, rhbm, buoyancy_kick, rdgas, rvgas, cp_air, hlv,
// kappa, grav, es0; flags: do_simp, do_shallower, do_changeqref, do_envsat.  Arrays [K][ncol] (phalf [K+1][ncol]).
int bm_host_run(int ncol, int K, double dt, const double* table, int n, const double* sp, const double* cfg, const int* flags,
                const double* tin, const double* qin, const double* pfull, const double* phalf, double* rain, double* tdel, double* qdel,
                double* q_ref, double* t_ref, int* bmflag, int* klzb, int* klcl, double* cape, double* cin, double* invtau_t,
                double* invtau_q) {
  BmSvp s{table, table + n, table + 2 * n, sp[0], sp[1], sp[2], sp[3], n};
  BmConst c;
  c.tau_bm = cfg[0]; c.rhbm = cfg[1]; c.buoyancy_kick = cfg[2]; c.rdgas = cfg[3]; c.rvgas = cfg[4]; c.cp_air = cfg[5]; c.hlv = cfg[6];
  c.kappa = cfg[7]; c.grav = cfg[8]; c.es0 = cfg[9];
  c.do_simp = flags[0]; c.do_shallower = flags[1]; c.do_changeqref = flags[2]; c.do_envsat = flags[3];
  c.lcltable = LCL;
  std::vector<double> tp(K), rp(K);
  int bad = 0;
  for (int col = 0; col < ncol; ++col) {
    BmOut o;
    bm_column(c, s, K, (size_t)ncol, (size_t)col, dt, tin, qin, pfull, phalf, tdel, qdel, q_ref, t_ref, tp.data(), rp.data(), o);
    rain[col] = o.rain; cape[col] = o.cape; cin[col] = o.cin; invtau_t[col] = o.invtau_t; invtau_q[col] = o.invtau_q;
    bmflag[col] = o.bmflag; klzb[col] = o.klzb; klcl[col] = o.klcl;
    bad |= o.bad;
  }
  return bad;
}

}  // extern "C"
