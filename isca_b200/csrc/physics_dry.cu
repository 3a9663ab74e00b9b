// physics_dry.cu -- dry convection scheme (convection_scheme = 'dry', idealized_moist_phys.F90:918-928) behind the physics C ABI.
#include "physics_common.h"
#include "physics_dry_kernels.h"

using namespace isca_phys;

namespace isca_phys {
void launch_dry_convection(IscaPhysics p, double tau, double gamma, const double* tg, const double* p_full, const double* p_half, double* tp,
                           double* dt_tg, double* cape, double* cin, int* lzb, int* lcl) {
  const int nb = col_blocks(p, 128);
  dryconv_k::dry_convection_kernel<<<nb, 128, 0, p->st>>>((int)p->ncol, p->K, tau, gamma, p->pc.rdgas / p->pc.cp_air, p->pc.rdgas, tg, p_full,
                                                          p_half, tp, dt_tg, cape, cin, lzb, lcl, p->d_err);
}
}  // namespace isca_phys

extern "C" int isca_b200_dry_convection(IscaPhysics p, double tau, double gamma, const double* tg, const double* p_full, const double* p_half,
                                        double* dt_tg, double* cape, double* cin, int* lzb, int* lcl) {
  if (!p) return fail(nullptr, "null handle");
  if (!(tau > 0.0)) return fail(p, "dry_convection: tau must be positive (dry_convection_nml has no default)");
  if (!dt_tg || !cape || !cin || !lzb || !lcl) return fail(p, "null output array");
  const size_t nc = p->ncol, n3 = nc * p->K;
  if (up(p, p->buf[0], tg, n3) || up(p, p->buf[1], p_full, n3) || up(p, p->buf[2], p_half, n3 + nc)) return 1;
  if (!p->buf[3].ensure(n3) || !p->buf[4].ensure(n3) || !p->buf[5].ensure(nc) || !p->buf[6].ensure(nc) || !p->buf[7].ensure(nc))
    return fail(p, "cudaMalloc failed");
  int* ib = reinterpret_cast<int*>(p->buf[7].p);                 // lzb | lcl (2 * nc ints fit in nc doubles)
  launch_dry_convection(p, tau, gamma, p->buf[0].p, p->buf[1].p, p->buf[2].p, p->buf[3].p, p->buf[4].p, p->buf[5].p, p->buf[6].p, ib, ib + nc);
  if (down(p, p->buf[4], dt_tg, n3) || down(p, p->buf[5], cape, nc) || down(p, p->buf[6], cin, nc)) return 1;
  PCK(cudaMemcpyAsync(lzb, ib, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaMemcpyAsync(lcl, ib + nc, nc * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  int e = 0;
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->st));
  PCK(cudaStreamSynchronize(p->st));
  if (e) {
    PCK(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->st));
    return fail(p, e == 65 ? "dry_convection: LCL above LZB" : (e == 64 ? "dry_convection: LCL defined, LZB not defined" : "dry_convection: device error flag set by an earlier kernel"));
  }
  return 0;
}
