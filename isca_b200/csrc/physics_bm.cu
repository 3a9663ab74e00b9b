// physics_bm.cu -- the full Betts-Miller convection scheme (betts_miller_mod, atmos_param/betts_miller/betts_miller.f90) of
// convection_scheme = 'FULL_BETTS_MILLER' (idealized_moist_phys.F90:889-916): one CUDA thread per column runs bm_column
// (physics_bm_column.h).  Like the simplified scheme (physics_conv.cu) the trip counts depend on the column (LCL, LZB, top of the
// shallow convection), so warps diverge, while every global access stays coalesced across the 32 columns of a warp; the parcel
// temperature / mixing ratio live in thread-local arrays, the four 3-D outputs double as the working profiles.
// bytes/column: read tin, qin, pfull (3K) + phalf (K+1); write tdel, qdel, q_ref, t_ref (4K) + 8 scalars  ~ (8K + 9) * 8
#include "physics_common.h"
#include "physics_bm_column.h"

using namespace isca_phys;

namespace {

__constant__ double d_lcltable[127] = {ISCA_BM_LCLTABLE_VALUES};

__global__ void __launch_bounds__(128, ISCA_COL_MINB) betts_miller_kernel(isca_bm::BmSvp s, isca_bm::BmConst c, int ncol, int K, double dt,
    const double* __restrict__ tin, const double* __restrict__ qin, const double* __restrict__ pfull, const double* __restrict__ phalf,
    double* __restrict__ rain, double* __restrict__ tdel, double* __restrict__ qdel, double* __restrict__ q_ref, double* __restrict__ t_ref,
    int* __restrict__ bmflag, int* __restrict__ klzbs, int* __restrict__ klcls, double* __restrict__ cape, double* __restrict__ cin,
    double* __restrict__ invtau_t, double* __restrict__ invtau_q, int* err) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  double tp[ISCA_KMAX], rp[ISCA_KMAX];
  c.lcltable = d_lcltable;
  isca_bm::BmOut o;
  isca_bm::bm_column(c, s, K, (size_t)ncol, (size_t)col, dt, tin, qin, pfull, phalf, tdel, qdel, q_ref, t_ref, tp, rp, o);
  rain[col] = o.rain; cape[col] = o.cape; cin[col] = o.cin; invtau_t[col] = o.invtau_t; invtau_q[col] = o.invtau_q;
  bmflag[col] = o.bmflag; klzbs[col] = o.klzb; klcls[col] = o.klcl;
  if (o.bad) atomicOr(err, 1);
}

}  // namespace

namespace isca_phys {

void launch_betts_miller(IscaPhysics p, double dt, const double* tin, const double* qin, const double* p_full, const double* p_half,
                         double* rain, double* tdel, double* qdel, double* q_ref, double* t_ref, int* bmflag, int* klzbs, int* klcls,
                         double* cape, double* cin, double* invtau_t, double* invtau_q) {
  const IscaBettsMillerConfig& b = p->bm;
  isca_bm::BmConst c;
  c.tau_bm = b.tau_bm; c.rhbm = b.rhbm; c.buoyancy_kick = b.buoyancy_kick;
  c.do_simp = b.do_simp; c.do_shallower = b.do_shallower; c.do_changeqref = b.do_changeqref; c.do_envsat = b.do_envsat;
  c.rdgas = p->cfg.rdgas; c.rvgas = p->cfg.rvgas; c.cp_air = p->cfg.cp_air; c.hlv = p->cfg.hlv; c.kappa = p->cfg.rdgas / p->cfg.cp_air;
  c.grav = p->cfg.grav; c.es0 = p->cfg.es0; c.lcltable = nullptr;
  isca_bm::BmSvp s{p->svp.tab, p->svp.dtab, p->svp.d2tab, p->svp.tminl, p->svp.dtinvl, p->svp.tepsl, p->svp.dtres, p->svp.n};
  betts_miller_kernel<<<col_blocks(p, 128), 128, 0, p->st>>>(s, c, (int)p->ncol, p->K, dt, tin, qin, p_full, p_half, rain, tdel, qdel, q_ref,
                                                             t_ref, bmflag, klzbs, klcls, cape, cin, invtau_t, invtau_q, p->d_err);
}

}  // namespace isca_phys

extern "C" {

int isca_b200_betts_miller_default_config(IscaBettsMillerConfig* c) {
  if (!c) return fail(nullptr, "null argument");
  c->abi_version = 1;
  c->tau_bm = 7200.; c->rhbm = .8; c->do_simp = 1; c->do_shallower = 0; c->do_changeqref = 0; c->do_envsat = 0; c->do_taucape = 0;
  c->capetaubm = 900.; c->tau_min = 2400.; c->buoyancy_kick = 0.;
  return 0;
}

int isca_b200_betts_miller_init(IscaPhysics p, const IscaBettsMillerConfig* c) {
  if (!p) return fail(nullptr, "null handle");
  if (!c) return fail(p, "betts_miller_init: null argument");
  if (c->abi_version != 1) return fail(p, "IscaBettsMillerConfig abi_version mismatch");
  if (c->do_taucape)
    return fail(p, "betts_miller_nml: do_taucape is not built (the reference rescales the module's tau_bm inside the grid loop, "
                   "betts_miller.f90:237-240: the result depends on the order of the columns)");
  if (!(c->tau_bm > 0.0)) return fail(p, "betts_miller_nml: tau_bm must be positive");
  p->bm = *c;
  return 0;
}

int isca_b200_betts_miller(IscaPhysics p, double dt, const double* tin, const double* qin, const double* pfull, const double* phalf,
                           double* rain, double* snow, double* tdel, double* qdel, double* q_ref, int* bmflag, int* klzbs, double* cape,
                           double* cin, double* t_ref, double* invtau_bm_t, double* invtau_bm_q, double* capeflag, int* klcls) {
  if (!p) return fail(nullptr, "null handle");
  if (!snow || !bmflag || !klzbs || !klcls) return fail(p, "null output array");
  size_t nc = p->ncol, n3 = nc * p->K;
  Dev* b = p->buf;
  if (up(p, b[0], tin, n3) || up(p, b[1], qin, n3) || up(p, b[2], pfull, n3) || up(p, b[3], phalf, n3 + nc)) return 1;
  for (int i = 4; i < 8; ++i) if (!b[i].ensure(n3)) return fail(p, "cudaMalloc failed");
  for (int i = 8; i < 15; ++i) if (!b[i].ensure(nc)) return fail(p, "cudaMalloc failed");
  int* iflag = reinterpret_cast<int*>(b[13].p);                // three int planes share two double-sized buffers
  int* ilzb = iflag + nc;
  int* ilcl = reinterpret_cast<int*>(b[14].p);
  launch_betts_miller(p, dt, b[0].p, b[1].p, b[2].p, b[3].p, b[8].p, b[4].p, b[5].p, b[6].p, b[7].p, iflag, ilzb, ilcl, b[9].p, b[10].p,
                      b[11].p, b[12].p);
  if (down(p, b[8], rain, nc) || down(p, b[4], tdel, n3) || down(p, b[5], qdel, n3) || down(p, b[6], q_ref, n3) || down(p, b[7], t_ref, n3) ||
      down(p, b[9], cape, nc) || down(p, b[10], cin, nc) || down(p, b[11], invtau_bm_t, nc) || down(p, b[12], invtau_bm_q, nc)) return 1;
  PCK(cudaMemcpyAsync(bmflag, iflag, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaMemcpyAsync(klzbs, ilzb, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaMemcpyAsync(klcls, ilcl, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  std::memset(snow, 0, nc * sizeof(double));                   // snow = 0. (:436)
  if (capeflag) std::memset(capeflag, 0, nc * sizeof(double)); // capeflag1 is never assigned in the reference (:197)
  return finish(p, "betts_miller");
}

}  // extern "C"
