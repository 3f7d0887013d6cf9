// athk_sweep_launch_impl.h -- bodies of the sweep launchers; included only by the translation
// units that instantiate them for their Riemann solver (ATHK_INSTANTIATE_RS at the bottom).
#ifndef ATHK_SWEEP_LAUNCH_IMPL_H_
#define ATHK_SWEEP_LAUNCH_IMPL_H_
#include "athk_sweep_launch.h"

template <int DIR, int ORDER, bool MHD, int RS>
static int launch_flux(athk_ctx *c, const SweepArgs &a) {
  const Grid &g = c->g;
  // loop limits of calculate_fluxes.cpp:68-76,176-182,299-302 (3-D)
  Range3 r;
  if (DIR == 0) {
    r = {g.ks, g.ke, g.js, g.je, g.is, g.ie + 1};
    if (MHD) { r.kl--; r.ku++; r.jl--; r.ju++; }
  } else if (DIR == 1) {
    r = {g.ks, g.ke, g.js, g.je + 1, g.is - 1, g.ie + 1};
    if (MHD) { r.kl--; r.ku++; }
  } else {
    r = {g.ks, g.ke + 1, g.js, g.je, g.is, g.ie};
    if (MHD) { r.jl--; r.ju++; r.il--; r.iu++; }
  }
  FacePoint<DIR, ORDER, MHD, RS> f = {a};
  LAUNCH_NAMED(c, DIR == 0 ? "calculate_fluxes_x1" : (DIR == 1 ? "calculate_fluxes_x2" : "calculate_fluxes_x3"),
               launch_points(c->stream, g.nb, r, f));
  return ATHK_OK;
}

template <int ORDER, bool MHD, int RS>
static int launch_flux3(athk_ctx *c, double dt, int breg) {
  const SweepArgs a = make_sweep_args(c, dt, breg, true);
  int rc = launch_flux<0, ORDER, MHD, RS>(c, a);
  if (!rc) rc = launch_flux<1, ORDER, MHD, RS>(c, a);
  if (!rc) rc = launch_flux<2, ORDER, MHD, RS>(c, a);
  return rc;
}


template <int ORDER, bool MHD, int RS, int AR>
static int launch_sweeps(athk_ctx *c, const SweepArgs &a) {
  const Grid &g = c->g;
#ifndef ATHK_EMULATE
  // x2 columns (i,k) / x3 columns (i,j): thread-marching kernel (athk_sweep_kernels.cuh);
  // extents of calculate_fluxes.cpp:176-182,299-302
  LAUNCH_NAMED(c, ORDER == 1 ? "sweep_x2_donor" : (ORDER == 2 ? "sweep_x2_plm" : "sweep_x2_ppm"),
               (launch_sweep_march<1, ORDER, MHD, RS, AR>(c->stream, a, g.nb, MHD ? g.ks - 1 : g.ks,
                                                 MHD ? g.ke + 1 : g.ke, g.is - 1, g.ie + 1)));
  LAUNCH_NAMED(c, ORDER == 1 ? "sweep_x3_donor" : (ORDER == 2 ? "sweep_x3_plm" : "sweep_x3_ppm"),
               (launch_sweep_march<2, ORDER, MHD, RS, AR>(c->stream, a, g.nb, MHD ? g.js - 1 : g.js,
                                                 MHD ? g.je + 1 : g.je, MHD ? g.is - 1 : g.is,
                                                 MHD ? g.ie + 1 : g.ie)));
#else
  {   // host harness: the same sweeps as per-thread marches
    Range3 r = {g.ks, g.ke, 0, 0, g.is - 1, g.ie + 1};
    if (MHD) { r.kl--; r.ku++; }
    SweepMarch<1, ORDER, MHD, RS> f = {a};
    CU(launch_points(c->stream, g.nb, r, f, 128));
  }
  {
    Range3 r = {0, 0, g.js, g.je, g.is, g.ie};
    if (MHD) { r.jl--; r.ju++; r.il--; r.iu++; }
    SweepMarch<2, ORDER, MHD, RS> f = {a};
    CU(launch_points(c->stream, g.nb, r, f, 128));
  }
#endif
  {   // x1 last (calculate_fluxes.cpp:68-76): sums d1+d2+d3 and integrates
    Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie + 1};
    if (MHD) { r.kl--; r.ku++; r.jl--; r.ju++; }
    SweepX1<ORDER, MHD, RS> f = {a};
    LAUNCH_NAMED(c, ORDER == 1 ? "sweep_x1_donor_integrate" : (ORDER == 2 ? "sweep_x1_plm_integrate" : "sweep_x1_ppm_integrate"),
                 (launch_sweep_x1<ORDER, MHD, RS, AR>(c->stream, g.nb, r, f)));
  }
  return ATHK_OK;
}


template <int RS, bool MHD>
int athk_flux3_rs(athk_ctx *c, int order, double dt, int breg) {
  if (order == 1) return launch_flux3<1, MHD, RS>(c, dt, breg);
  if (order == 2) return launch_flux3<2, MHD, RS>(c, dt, breg);
  return launch_flux3<3, MHD, RS>(c, dt, breg);
}

template <int RS, bool MHD, int AR>
int athk_sweeps_rs(athk_ctx *c, int order, const SweepArgs &a) {
  if (order == 1) return launch_sweeps<1, MHD, RS, AR>(c, a);
  if (order == 2) return launch_sweeps<2, MHD, RS, AR>(c, a);
  return launch_sweeps<3, MHD, RS, AR>(c, a);
}

#define ATHK_INSTANTIATE_RS(RS, MHD)                                                \
  template int athk_flux3_rs<RS, MHD>(athk_ctx *, int, double, int);               \
  template int athk_sweeps_rs<RS, MHD, 0>(athk_ctx *, int, const SweepArgs &);

// the fast-arithmetic set of a solver: ONLY in a translation unit compiled with --fmad=true
#define ATHK_INSTANTIATE_RS_FAST(RS, MHD) \
  template int athk_sweeps_rs<RS, MHD, 1>(athk_ctx *, int, const SweepArgs &);
#define ATHK_EXTERN_RS_FAST(RS, MHD) \
  extern template int athk_sweeps_rs<RS, MHD, 1>(athk_ctx *, int, const SweepArgs &);

// in a unit that calls, but does not own, a solver's launchers
#define ATHK_EXTERN_RS(RS, MHD)                                                     \
  extern template int athk_flux3_rs<RS, MHD>(athk_ctx *, int, double, int);        \
  extern template int athk_sweeps_rs<RS, MHD, 0>(athk_ctx *, int, const SweepArgs &);

#endif  // ATHK_SWEEP_LAUNCH_IMPL_H_
