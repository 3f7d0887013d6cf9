// athk_scalars.h -- passive scalars: PassiveScalars::CalculateFluxes + ComputeUpwindFlux
// (scalars/calculate_scalar_fluxes.cpp:41-391,399-417), PassiveScalars::AddFluxDivergence
// (scalars/add_scalar_flux_divergence.cpp:42-96) and the IntegrateScalars register update
// (task_list/time_integrator.cpp:2376-2403).  The primitive r = s/rho is recovered in
// Cons2PrimPoint (athk_ops.h; eos/eos_scalars.cpp:31-60).
//
// Reconstruction of r is the reference's generic-array variant of the hydro one (dc_simple.cpp,
// plm_simple.cpp:35-143, ppm_simple.cpp:46-299): same limiter arithmetic per variable, and for
// xorder=3 the concentration floor on both face states (calculate_scalar_fluxes.cpp:91-97).
//
// Extents: the reference also fills s_flux on the transverse ghost rows (js-1, je+1, ...); nothing
// reads them on a uniform mesh (only AddFluxDivergence over the active zone consumes s_flux), so
// the operators here cover the faces of the active zone only.
#ifndef ATHK_SCALARS_H_
#define ATHK_SCALARS_H_

#include "athk_sweep.h"

namespace athk {

// where the hydro mass flux F(rho) of a face lives: component IDN of Hydro::flux[d] (operator
// path, block stride 5*sn) or component 0 of s_flux[d], which the fused sweeps use as scratch
// for it (block stride ns*sn)
struct MassFlux {
  const Real *p[3];
  idx_t bstride;
};

// upwinded flux of one scalar through the DIR-face below cell c (index n_dir along DIR);
// rb = &r(n, cell c)
template <int DIR, int ORDER>
ATHK_HD Real scalar_face_flux(const Grid &g, const Geom &geo, Real sfloor, const Real *rb, int blk,
                              int n_dir, Real mass_flx) {
  const idx_t s = (DIR == 0) ? 1 : (DIR == 1 ? g.sj : g.sk);
  Real rl, rr;
  if (ORDER == 1) {
    rl = rb[-s];
    rr = rb[0];
  } else if (ORDER == 2) {
    const Real q0 = rb[-2 * s], q1 = rb[-s], q2 = rb[0], q3 = rb[s];
    const Real *cpv = geo.cp[DIR] + blk * geo.PG;
    const Real *cmv = geo.cm[DIR] + blk * geo.PG;
    Real dqm = vl_slope(q1 - q0, q2 - q1);
    rl = q1 + cpv[n_dir - 1] * dqm;
    dqm = vl_slope(q2 - q1, q3 - q2);
    rr = q2 - cmv[n_dir] * dqm;
  } else {
    Real up, lo;
    ppm_cell(rb[-3 * s], rb[-2 * s], rb[-s], rb[0], rb[s], &up, &lo);
    rl = up;
    ppm_cell(rb[-2 * s], rb[-s], rb[0], rb[s], rb[2 * s], &up, &lo);
    rr = lo;
    rl = (rl > sfloor) ? rl : sfloor;   // ApplyPassiveScalarFloors (eos_scalars.cpp:161-173)
    rr = (rr > sfloor) ? rr : sfloor;
  }
  return (mass_flx >= 0.0) ? mass_flx * rl : mass_flx * rr;   // ComputeUpwindFlux
}

// Both DIR-faces of cell c at once (lower face n_dir, upper face n_dir + 1) for the fused integrate:
// the cell's own reconstruction serves both faces, so a direction costs three reconstructions
// (cells n-1, n, n+1) instead of four.  Same values as two scalar_face_flux calls.
template <int DIR, int ORDER>
ATHK_HD void scalar_face_flux_pair(const Grid &g, const Geom &geo, Real sfloor, const Real *rb, int blk,
                                   int n_dir, Real m_lo, Real m_hi, Real *f_lo, Real *f_hi) {
  const idx_t s = (DIR == 0) ? 1 : (DIR == 1 ? g.sj : g.sk);
  Real up_m, lo_0, up_0, lo_p;   // states: cell n-1 upper, cell n lower/upper, cell n+1 lower
  if (ORDER == 1) {
    up_m = rb[-s];
    lo_0 = up_0 = rb[0];
    lo_p = rb[s];
  } else if (ORDER == 2) {
    const Real q0 = rb[-2 * s], q1 = rb[-s], q2 = rb[0], q3 = rb[s], q4 = rb[2 * s];
    const Real *cpv = geo.cp[DIR] + blk * geo.PG;
    const Real *cmv = geo.cm[DIR] + blk * geo.PG;
    const Real dm = vl_slope(q1 - q0, q2 - q1);   // cell n-1
    const Real d0 = vl_slope(q2 - q1, q3 - q2);   // cell n
    const Real dp = vl_slope(q3 - q2, q4 - q3);   // cell n+1
    up_m = q1 + cpv[n_dir - 1] * dm;
    lo_0 = q2 - cmv[n_dir] * d0;
    up_0 = q2 + cpv[n_dir] * d0;
    lo_p = q3 - cmv[n_dir + 1] * dp;
  } else {
    // ComputeUpwindFlux takes ONE state per face: the cell upwind of it.  Only those cells are
    // reconstructed (two of the three; cell n alone when both of its faces carry mass out of it);
    // the states that would be discarded are never formed, the ones used are the same bits
    Real t;
    up_m = lo_0 = up_0 = lo_p = 0.0;
    if (m_lo >= 0.0) {
      ppm_cell(rb[-3 * s], rb[-2 * s], rb[-s], rb[0], rb[s], &up_m, &t);
      up_m = (up_m > sfloor) ? up_m : sfloor;   // ApplyPassiveScalarFloors (eos_scalars.cpp:161-173)
    }
    if (!(m_lo >= 0.0) || m_hi >= 0.0) {
      ppm_cell(rb[-2 * s], rb[-s], rb[0], rb[s], rb[2 * s], &up_0, &lo_0);
      lo_0 = (lo_0 > sfloor) ? lo_0 : sfloor;
      up_0 = (up_0 > sfloor) ? up_0 : sfloor;
    }
    if (!(m_hi >= 0.0)) {
      ppm_cell(rb[-s], rb[0], rb[s], rb[2 * s], rb[3 * s], &t, &lo_p);
      lo_p = (lo_p > sfloor) ? lo_p : sfloor;
    }
  }
  *f_lo = (m_lo >= 0.0) ? m_lo * up_m : m_lo * lo_0;   // ComputeUpwindFlux
  *f_hi = (m_hi >= 0.0) ? m_hi * up_0 : m_hi * lo_p;
}

// PassiveScalars::CalculateFluxes as an operator: one thread per DIR-face, all scalars
template <int DIR, int ORDER>
struct ScalarFacePoint {
  Grid g;
  Geom geo;
  EosParams eos;    // for the concentration floor of block blk
  int ns;
  const Real *r;    // [nb][ns][sn]
  MassFlux mf;
  Real *sflux;      // s_flux[DIR]  [nb][ns][sn]
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const Real sfloor = eos_sfloor(eos, blk);
    const idx_t c = gidx(g, k, j, i);
    const Real m = mf.p[DIR][(idx_t)blk * mf.bstride + c];
    const int n_dir = (DIR == 0) ? i : (DIR == 1 ? j : k);
    for (int n = 0; n < ns; ++n) {
      const idx_t o = ((idx_t)blk * ns + n) * g.sn + c;
      sflux[o] = scalar_face_flux<DIR, ORDER>(g, geo, sfloor, r + o, blk, n_dir, m);
    }
  }
};

// Fused IntegrateScalars for one active cell: the six face fluxes of every scalar are formed in
// place from r and the mass fluxes the sweeps left behind (each face is evaluated by both of its
// cells -- some flops against one less pass over s_flux[3]; within a cell the two faces of a
// direction share the cell's reconstruction), then WeightedAve + flux divergence in the
// reference's order.
template <int ORDER>
struct ScalarIntegratePoint {
  Grid g;
  Geom geo;
  EosParams eos;       // for the concentration floor of block blk
  int ns;
  const Real *r;
  MassFlux mf;
  HydroIntegrate hi;   // registers s / s1 and the stage weights
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const Real sfloor = eos_sfloor(eos, blk);
    const idx_t c = gidx(g, k, j, i);
    const Real dx1 = geo.dxf[0][blk * geo.PG + i], dx2 = geo.dxf[1][blk * geo.PG + j],
               dx3 = geo.dxf[2][blk * geo.PG + k];
    const Real a1 = dx2 * dx3, a2 = dx1 * dx3, a3 = dx1 * dx2;
    const Real vol = dx1 * dx2 * dx3;
    const idx_t mb = (idx_t)blk * mf.bstride + c;
    const Real m1lo = mf.p[0][mb], m1hi = mf.p[0][mb + 1];
    const Real m2lo = mf.p[1][mb], m2hi = mf.p[1][mb + g.sj];
    const Real m3lo = mf.p[2][mb], m3hi = mf.p[2][mb + g.sk];
    for (int n = 0; n < ns; ++n) {
      const idx_t o = ((idx_t)blk * ns + n) * g.sn + c;
      const Real *rb = r + o;
      Real flo, fhi;
      scalar_face_flux_pair<0, ORDER>(g, geo, sfloor, rb, blk, i, m1lo, m1hi, &flo, &fhi);
      Real dflx = (a1 * fhi - a1 * flo);
      scalar_face_flux_pair<1, ORDER>(g, geo, sfloor, rb, blk, j, m2lo, m2hi, &flo, &fhi);
      dflx += (a2 * fhi - a2 * flo);
      scalar_face_flux_pair<2, ORDER>(g, geo, sfloor, rb, blk, k, m3lo, m3hi, &flo, &fhi);
      dflx += (a3 * fhi - a3 * flo);
      integrate_var(hi, o, dflx, vol);
    }
  }
};

}  // namespace athk
#endif  // ATHK_SCALARS_H_
