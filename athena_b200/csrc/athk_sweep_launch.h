// athk_sweep_launch.h -- declarations of the per-Riemann-solver sweep launchers.  Each solver is a
// separate set of kernel instantiations in its own translation unit (athk_api.cu: HLLC + HLLD,
// athk_rs_roe.cu, athk_rs_hlle.cu, athk_rs_llf.cu); bodies in athk_sweep_launch_impl.h.
#ifndef ATHK_SWEEP_LAUNCH_H_
#define ATHK_SWEEP_LAUNCH_H_
#include "athk_ctx.h"

static inline SweepArgs make_sweep_args(athk_ctx *c, double dt, int breg, bool store_flux) {
  SweepArgs a;
  memset(&a, 0, sizeof(a));
  a.g = c->g; a.eos = c->eos; a.geo = c->geo;
  a.w = c->w; a.bcc = c->bcc; a.b = c->cfg.mhd ? c->b[breg] : nullptr;
  for (int d = 0; d < 3; ++d) a.flx[d] = store_flux ? c->flux[d] : nullptr;
  a.ef = c->ef; a.wct = c->wct;
  a.d2 = c->flux[1];   // fused path only: the x2/x3 flux arrays double as scratch terms
  a.d3 = c->flux[2];
  a.dt = dt;
  return a;
}

// operator path: CalculateFluxes for one stage (x1, x2, x3 face kernels), xorder = order
template <int RS, bool MHD>
int athk_flux3_rs(athk_ctx *c, int order, double dt, int breg);
// fused path: x2 / x3 tile sweeps then the x1 sweep that integrates
template <int RS, bool MHD, int AR>
int athk_sweeps_rs(athk_ctx *c, int order, const SweepArgs &a);

#endif  // ATHK_SWEEP_LAUNCH_H_
