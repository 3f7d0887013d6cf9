// athk_physics.h -- point-wise FP64 physics of the hydro/MHD finite-volume update.
//
// Every function here is pure (registers in, registers out), marked __host__ __device__ so the
// identical arithmetic can be exercised by the GPU kernels and by the pre-GPU index-logic harness
// under tests/emul (test infrastructure, never a product path).
//
// PARITY CONTRACT: results must be bit-identical to the reference's double-precision CPU build
// (g++ -O3, no fast-math, no FMA contraction).  Hence:
//   * this translation unit is compiled with --fmad=false;
//   * every expression keeps the reference's association order (C++ evaluates + - * / left to
//     right; parentheses are reproduced where the reference groups terms);
//   * min/max follow std::min/std::max tie semantics (return the first argument on ties).
// Reference sources followed (file:line under /root/reference/src):
//   eos/adiabatic_hydro.cpp:38-80, eos/adiabatic_mhd.cpp:41-90,149-156 (W(U), sound/fast speed)
//   reconstruct/plm.cpp:104-112,149-154 (van Leer slope, uniform Cartesian branch)
//   hydro/rsolvers/hydro/hllc.cpp:32-179, hydro/rsolvers/mhd/hlld.cpp:36-376
//   hydro/hydro.cpp:154-158 (CT upwind weight)
#ifndef ATHK_PHYSICS_H_
#define ATHK_PHYSICS_H_

#include <math.h>

#include "athk_math.h"

namespace athk {

typedef double Real;

// variable indices, as src/athena.hpp:169-192
enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4 };
enum { IVX = 1, IVY = 2, IVZ = 3, IPR = 4, IBY = 5, IBZ = 6 };
enum { IB1 = 0, IB2 = 1, IB3 = 2 };

#define ATHK_TINY_NUMBER 1.0e-20

ATHK_HD Real sqr(Real x) { return x * x; }
// std::min(a,b) == (b < a) ? b : a ; std::max(a,b) == (a < b) ? b : a
ATHK_HD Real min_std(Real a, Real b) { return (b < a) ? b : a; }
ATHK_HD Real max_std(Real a, Real b) { return (a < b) ? b : a; }

// IEEE division a/b with a shortcut for an exactly-zero numerator: 0/b == 0*b (signed zero,
// same bits) for every finite non-zero b.  Static uniform regions (v = 0, equal pressures) make
// several HLLD/HLLC numerators exactly 0, which sends the GPU's FP64 division into its
// ~100-instruction slow path; the shortcut returns the identical result in 3 instructions.
ATHK_HD Real fdiv(Real a, Real b) { return fdiv_ieee(a, b); }

struct EosParams {
  Real gamma, gm1, igm1;   // igm1 = 1.0/(gamma-1.0)
  Real dfloor, pfloor, sfloor;
  // Every MeshBlock of the reference owns an EquationOfState with its own floors (eos/eos.hpp:
  // 161-163), and they can differ: null = the three values above for every block, else
  // [nb][3] = (dfloor, pfloor, sfloor) of each local block (athk_set_block_floors)
  const Real *blk_floors;
  // isothermal EOS (eos/isothermal_hydro.cpp, isothermal_mhd.cpp): p = iso_cs^2 * rho, no energy
  // equation.  The registers keep their five components; the fifth stays zero (zero flux, zero
  // source) and the pressure slot of w is a dummy.
  int iso;
  Real iso_cs;
  // special relativity (configure.py -s, hydro): eos/adiabatic_hydro_sr.cpp.  w holds (rho, u^1, u^2,
  // u^3, pgas) with u^i the spatial components of the four-velocity; u holds (D, M^1, M^2, M^3, E).
  int sr;
  int sr_solver;    // RS_HLLE, RS_HLLC or RS_LLF (hlle_rel.cpp, hllc_rel.cpp, llf_rel.cpp)
  Real gamma_max;   // <hydro> gamma_max (velocity ceiling, default 1000)
};

// floors of block blk (uniform branch; a no-op unless per-block floors were set)
ATHK_HD Real eos_dfloor(const EosParams &e, int blk) { return e.blk_floors ? e.blk_floors[3 * blk] : e.dfloor; }
ATHK_HD Real eos_pfloor(const EosParams &e, int blk) { return e.blk_floors ? e.blk_floors[3 * blk + 1] : e.pfloor; }
ATHK_HD Real eos_sfloor(const EosParams &e, int blk) { return e.blk_floors ? e.blk_floors[3 * blk + 2] : e.sfloor; }
ATHK_HD EosParams eos_of_block(const EosParams &e, int blk) {
  EosParams r = e;
  if (e.blk_floors) {
    r.dfloor = e.blk_floors[3 * blk];
    r.pfloor = e.blk_floors[3 * blk + 1];
    r.sfloor = e.blk_floors[3 * blk + 2];
  }
  return r;
}

ATHK_HD EosParams make_eos(Real gamma, Real dfloor, Real pfloor, Real sfloor) {
  EosParams e;
  e.gamma = gamma;
  e.gm1 = gamma - 1.0;
  e.igm1 = 1.0 / (gamma - 1.0);
  e.dfloor = dfloor;
  e.pfloor = pfloor;
  e.sfloor = sfloor;
  e.blk_floors = nullptr;
  e.iso = 0;
  e.iso_cs = 0.0;
  e.sr = 0;
  e.sr_solver = 0;
  e.gamma_max = 1000.0;
  return e;
}

//----------------------------------------------------------------------------------------
// EquationOfState::ConservedToPrimitive, one cell.  u[] is updated in place when a floor
// fires (density floor / energy write-back), exactly as the reference does.
// eos/adiabatic_hydro.cpp:58-72
ATHK_HD void cons2prim_hydro(const EosParams &eos, Real u[5], Real w[5]) {
  u[IDN] = (u[IDN] > eos.dfloor) ? u[IDN] : eos.dfloor;
  w[IDN] = u[IDN];
  Real di = 1.0 / u[IDN];
  w[IVX] = u[IM1] * di;
  w[IVY] = u[IM2] * di;
  w[IVZ] = u[IM3] * di;
  Real e_k = 0.5 * di * (sqr(u[IM1]) + sqr(u[IM2]) + sqr(u[IM3]));
  Real p = eos.gm1 * (u[IEN] - e_k);
  u[IEN] = (p > eos.pfloor) ? u[IEN] : ((eos.pfloor / eos.gm1) + e_k);
  w[IPR] = (p > eos.pfloor) ? p : eos.pfloor;
}

// eos/adiabatic_mhd.cpp:63-84 (bcc already interpolated by the caller, field/field.cpp:112-180)
ATHK_HD void cons2prim_mhd(const EosParams &eos, Real u[5], Real bc1, Real bc2, Real bc3,
                           Real w[5]) {
  u[IDN] = (u[IDN] > eos.dfloor) ? u[IDN] : eos.dfloor;
  w[IDN] = u[IDN];
  Real di = 1.0 / u[IDN];
  w[IVX] = u[IM1] * di;
  w[IVY] = u[IM2] * di;
  w[IVZ] = u[IM3] * di;
  Real pb = 0.5 * (sqr(bc1) + sqr(bc2) + sqr(bc3));
  Real e_k = 0.5 * di * (sqr(u[IM1]) + sqr(u[IM2]) + sqr(u[IM3]));
  Real p = eos.gm1 * (u[IEN] - e_k - pb);
  u[IEN] = (p > eos.pfloor) ? u[IEN] : ((eos.pfloor / eos.gm1) + e_k + pb);
  w[IPR] = (p > eos.pfloor) ? p : eos.pfloor;
}

// eos/isothermal_hydro.cpp:50-72, isothermal_mhd.cpp:60-78: density floor, velocities; nothing else
ATHK_HD void cons2prim_iso(const EosParams &eos, Real u[5], Real w[5]) {
  u[IDN] = (u[IDN] > eos.dfloor) ? u[IDN] : eos.dfloor;
  w[IDN] = u[IDN];
  Real di = 1.0 / u[IDN];
  w[IVX] = u[IM1] * di;
  w[IVY] = u[IM2] * di;
  w[IVZ] = u[IM3] * di;
  w[IPR] = 0.0;
}
// eos/isothermal_mhd.cpp:121-129
ATHK_HD Real fast_speed_iso(const EosParams &eos, Real d, Real by, Real bz, Real bx) {
  Real asq = (eos.iso_cs * eos.iso_cs) * d;
  Real vaxsq = bx * bx;
  Real ct2 = by * by + bz * bz;
  Real qsq = vaxsq + ct2 + asq;
  Real tmp = vaxsq + ct2 - asq;
  return sqrt(0.5 * (qsq + sqrt(tmp * tmp + 4.0 * asq * ct2)) / d);
}

//----------------------------------------------------------------------------------------
// Special-relativistic hydro (eos/adiabatic_hydro_sr.cpp).
// std::isfinite without <cmath> differences between the two compilation passes
ATHK_HD bool finite_ath(Real x) { return (x - x) == 0.0; }

// ConservedToPrimitiveNormal (adiabatic_hydro_sr.cpp:303-411): fixed-point iteration on the gas
// pressure with an Aitken step every third iteration (Newman & Hamlin 2014)
ATHK_HD bool cons2prim_sr_normal(Real dd, Real ee, Real mm_sq, Real mm1, Real mm2, Real mm3, Real gamma_adi,
                                 Real pgas_old, Real w[5], Real *p_gamma_lor) {
  const int max_iterations = 15;
  const Real tol = 1.0e-12;
  const Real pgas_uniform_min = 1.0e-12;
  const Real a_min = 1.0e-12;
  const Real v_sq_max = 1.0 - 1.0e-12;
  const Real rr_max = 1.0 - 1.0e-12;
  Real pgas_min = -ee;
  pgas_min = max_std(pgas_min, pgas_uniform_min);
  // pgas[n % 3] of the reference kept in three scalars (p0, p1, p2): no dynamically indexed array
  Real p0 = max_std(pgas_old, pgas_min), p1 = 0.0, p2 = 0.0;
  Real result = 0.0;
  int n;
  for (n = 0; n < max_iterations; ++n) {
    const int r = n % 3;
    if (r != 2) {
      const Real pin = (r == 0) ? p0 : p1;
      Real a = ee + pin;
      a = max_std(a, a_min);
      Real v_sq = mm_sq / sqr(a);
      v_sq = min_std(max_std(v_sq, 0.0), v_sq_max);
      Real gamma_sq = 1.0 / (1.0 - v_sq);
      Real gamma = sqrt(gamma_sq);
      Real wgas = a / gamma_sq;
      Real rho = dd / gamma;
      Real pout = (gamma_adi - 1.0) / gamma_adi * (wgas - rho);
      pout = max_std(pout, pgas_min);
      if (r == 0) p1 = pout; else p2 = pout;
      if (pout > pgas_min && fabs(pout - pin) < tol) { result = pout; break; }
    } else {
      Real rr = (p2 - p1) / (p1 - p0);
      if (!finite_ath(rr) || fabs(rr) > rr_max) continue;
      p0 = p1 + (p2 - p1) / (1.0 - rr);
      p0 = max_std(p0, pgas_min);
      if (p0 > pgas_min && fabs(p0 - p2) < tol) { result = p0; break; }
    }
  }
  if (n == max_iterations) return false;
  w[IPR] = result;   // (pgas[(n+1)%3]: the value the break left there)
  if (!finite_ath(w[IPR])) return false;
  Real a = ee + w[IPR];
  a = max_std(a, a_min);
  Real v_sq = mm_sq / sqr(a);
  v_sq = min_std(max_std(v_sq, 0.0), v_sq_max);
  Real gamma_sq = 1.0 / (1.0 - v_sq);
  Real gamma = sqrt(gamma_sq);
  w[IDN] = dd / gamma;
  if (!finite_ath(w[IDN])) return false;
  Real v1 = mm1 / a;
  Real v2 = mm2 / a;
  Real v3 = mm3 / a;
  w[IVX] = gamma * v1;
  w[IVY] = gamma * v2;
  w[IVZ] = gamma * v3;
  if (!finite_ath(w[IVX]) || !finite_ath(w[IVY]) || !finite_ath(w[IVZ])) return false;
  *p_gamma_lor = gamma;
  return true;
}

// EquationOfState::ConservedToPrimitive of one cell (adiabatic_hydro_sr.cpp:67-216).  On entry w[]
// holds what the output array held before the call (the reference's fallback state and the
// initial guess of the iteration); returns true when u[] was adjusted and has to be written back.
ATHK_HD bool cons2prim_sr(const EosParams &eos, Real u[5], Real w[5]) {
  const Real mm_sq_ee_sq_max = 1.0 - 1.0e-12;
  const Real gamma_adi = eos.gamma;
  Real dd = u[IDN], ee = u[IEN];
  Real mm0 = sqr(u[IM1]) + sqr(u[IM2]) + sqr(u[IM3]);
  Real mm1 = u[IM1], mm2 = u[IM2], mm3 = u[IM3];
  bool fixed = false;
  const Real dd_min = eos.dfloor;
  if (dd < dd_min) { dd = dd_min; fixed = true; }
  const Real ee_min = eos.dfloor + eos.pfloor / (gamma_adi - 1.0);
  if (ee < ee_min) { ee = ee_min; fixed = true; }
  const Real mm_sq_max = mm_sq_ee_sq_max * sqr(ee);
  if (mm0 > mm_sq_max) {
    const Real factor = sqrt(mm_sq_max / mm0);
    mm0 = mm_sq_max;
    mm1 *= factor;
    mm2 *= factor;
    mm3 *= factor;
    fixed = true;
  }
  const Real rho_old = max_std(w[IDN], eos.dfloor);
  const Real pgas_old = max_std(w[IPR], eos.pfloor);
  const Real u1_old = w[IVX], u2_old = w[IVY], u3_old = w[IVZ];
  Real gamma = 0.0;
  bool success = cons2prim_sr_normal(dd, ee, mm0, mm1, mm2, mm3, gamma_adi, pgas_old, w, &gamma);
  if (!success) {
    w[IDN] = rho_old; w[IPR] = pgas_old; w[IVX] = u1_old; w[IVY] = u2_old; w[IVZ] = u3_old;
    fixed = true;
  }
  const Real rho_add = max_std(eos.dfloor - w[IDN], 0.0);
  const Real pgas_add = max_std(eos.pfloor - w[IPR], 0.0);
  if (success && (rho_add > 0.0 || pgas_add > 0.0)) {
    const Real wgas_add = rho_add + gamma_adi / (gamma_adi - 1.0) * pgas_add;
    dd += rho_add * gamma;
    ee += wgas_add * sqr(gamma) + pgas_add;
    success = cons2prim_sr_normal(dd, ee, mm0, mm1, mm2, mm3, gamma_adi, pgas_old, w, &gamma);
    if (!success) {
      w[IDN] = rho_old; w[IPR] = pgas_old; w[IVX] = u1_old; w[IVY] = u2_old; w[IVZ] = u3_old;
    }
    fixed = true;
  }
  if (!success) gamma = sqrt(1.0 + sqr(w[IVX]) + sqr(w[IVY]) + sqr(w[IVZ]));
  if (gamma > eos.gamma_max) {
    const Real factor = sqrt((sqr(eos.gamma_max) - 1.0) / (sqr(gamma) - 1.0));
    w[IVX] *= factor;
    w[IVY] *= factor;
    w[IVZ] *= factor;
    fixed = true;
  }
  if (w[IDN] < eos.dfloor) { w[IDN] = eos.dfloor; fixed = true; }
  if (w[IPR] < eos.pfloor) { w[IPR] = eos.pfloor; fixed = true; }
  if (!success) {
    w[IDN] = eos.dfloor;
    w[IPR] = eos.pfloor;
    w[IVX] = w[IVY] = w[IVZ] = 0.0;
  }
  if (fixed) {
    const Real wgas = w[IDN] + gamma_adi / (gamma_adi - 1.0) * w[IPR];
    gamma = sqrt(1.0 + sqr(w[IVX]) + sqr(w[IVY]) + sqr(w[IVZ]));
    u[IDN] = w[IDN] * gamma;
    u[IEN] = wgas * sqr(gamma) - w[IPR];
    u[IM1] = wgas * gamma * w[IVX];
    u[IM2] = wgas * gamma * w[IVY];
    u[IM3] = wgas * gamma * w[IVZ];
  }
  return fixed;
}

// EquationOfState::SoundSpeedsSR (adiabatic_hydro_sr.cpp:271-281; Mignone & Bodo 2005 eq. 23)
ATHK_HD void sound_speeds_sr(const EosParams &eos, Real rho_h, Real pgas, Real vx, Real gamma_lorentz_sq,
                             Real *plambda_plus, Real *plambda_minus) {
  Real cs_sq = eos.gamma * pgas / rho_h;
  Real sigma_s = cs_sq / (gamma_lorentz_sq * (1.0 - cs_sq));
  Real relative_speed = sqrt(sigma_s * (1.0 + sigma_s - sqr(vx)));
  Real sigma_s_tmp = 1.0 / (1.0 + sigma_s);
  *plambda_plus = sigma_s_tmp * (vx + relative_speed);
  *plambda_minus = sigma_s_tmp * (vx - relative_speed);
}

// eos/adiabatic_hydro.cpp SoundSpeed ; eos/adiabatic_mhd.cpp:149-156 FastMagnetosonicSpeed
ATHK_HD Real sound_speed(const EosParams &eos, Real d, Real p) { return sqrt(eos.gamma * p / d); }

template <class M>
ATHK_HD Real fast_speed_m(M &m, const EosParams &eos, Real d, Real p, Real by, Real bz, Real bx) {
  Real asq = eos.gamma * p;
  Real vaxsq = bx * bx;
  Real ct2 = (by * by + bz * bz);
  Real qsq = vaxsq + ct2 + asq;
  Real tmp = vaxsq + ct2 - asq;
  return m.sqrt(m.div(0.5 * (qsq + m.sqrt(tmp * tmp + 4.0 * asq * ct2)), d));
}
ATHK_HD Real fast_speed(const EosParams &eos, Real d, Real p, Real by, Real bz, Real bx) {
  MathIEEE m;
  return fast_speed_m(m, eos, d, p, by, bz, bx);
}

//----------------------------------------------------------------------------------------
// PLM, uniform-Cartesian van Leer limiter (reconstruct/plm.cpp:104-112): limited slope of one
// variable from its left/right one-sided differences.
// The reference evaluates 2*dw2/(dwl+dwr) unconditionally and then overwrites it with 0 when
// dw2 <= 0; skipping the division in that case gives the same bits and keeps 0/0 (uniform
// regions) off the FP64 division slow path.  !(dw2 <= 0) keeps NaN inputs on the division.
template <class M>
ATHK_HD Real vl_slope_m(M &m, Real dwl, Real dwr) { return m.vl(dwl * dwr, dwl, dwr); }
ATHK_HD Real vl_slope(Real dwl, Real dwr) {
  MathIEEE m;
  return vl_slope_m(m, dwl, dwr);
}

//----------------------------------------------------------------------------------------
// PPM of one variable in one cell, uniform Cartesian branch of reconstruct/ppm.cpp:107-320
// (4th-order interface values with c1..c4 = 1/2, c5 = 1/6, c6 = -1/6, reconstruction.cpp:494-505;
// Colella-Dawson interface limiter :139-193; Colella-Sekora / McCorquodale-Colella
// extremum-preserving limiter :226-279, time/ext_preserving = true, C2 = 1.25).
// Returns the left state of the upper face (ql_iph) and the right state of the lower face (qr_imh).
// (qd/6 and qe/qd below go through fdiv: both numerators are exactly 0 wherever the profile is
//  locally linear or uniform, which would otherwise take the FP64 division's slow path)
ATHK_HD Real sign_ath(Real x) { return (x < 0.0) ? -1.0 : 1.0; }   // SIGN() of defs.hpp.in:150

ATHK_HD void ppm_cell(Real q_im2, Real q_im1, Real q, Real q_ip1, Real q_ip2, Real *ql_iph,
                      Real *qr_imh) {
  const Real C2 = 1.25;
  const Real c1 = 0.5, c2 = 0.5, c3 = 0.5, c4 = 0.5, c5 = 1.0 / 6.0, c6 = -1.0 / 6.0;
  Real qa = (q - q_im1);
  Real qb = (q_ip1 - q);
  const Real dd_im1 = c1 * qa + c2 * (q_im1 - q_im2);
  const Real dd = c1 * qb + c2 * qa;
  const Real dd_ip1 = c1 * (q_ip2 - q_ip1) + c2 * qb;
  Real dph = (c3 * q_im1 + c4 * q) + (c5 * dd_im1 + c6 * dd);
  Real dph_ip1 = (c3 * q + c4 * q_ip1) + (c5 * dd + c6 * dd_ip1);
  const Real d2qc_im1 = q_im2 + q - 2.0 * q_im1;
  const Real d2qc = q_im1 + q_ip1 - 2.0 * q;
  const Real d2qc_ip1 = q + q_ip2 - 2.0 * q_ip1;
  {   // i-1/2 interface
    const Real qa_tmp = dph - q_im1;
    const Real qb_tmp = q - dph;
    qa = 3.0 * (q_im1 + q - 2.0 * dph);
    qb = d2qc_im1;
    const Real qc = d2qc;
    Real qd = 0.0;
    if (sign_ath(qa) == sign_ath(qb) && sign_ath(qa) == sign_ath(qc))
      qd = sign_ath(qa) * min_std(C2 * fabs(qb), min_std(C2 * fabs(qc), fabs(qa)));
    const Real dph_tmp = 0.5 * (q_im1 + q) - fdiv(qd, 6.0);
    if (qa_tmp * qb_tmp < 0.0) dph = dph_tmp;
  }
  {   // i+1/2 interface
    const Real qa_tmp = dph_ip1 - q;
    const Real qb_tmp = q_ip1 - dph_ip1;
    qa = 3.0 * (q + q_ip1 - 2.0 * dph_ip1);
    qb = d2qc;
    const Real qc = d2qc_ip1;
    Real qd = 0.0;
    if (sign_ath(qa) == sign_ath(qb) && sign_ath(qa) == sign_ath(qc))
      qd = sign_ath(qa) * min_std(C2 * fabs(qb), min_std(C2 * fabs(qc), fabs(qa)));
    const Real dphip1_tmp = 0.5 * (q + q_ip1) - fdiv(qd, 6.0);
    if (qa_tmp * qb_tmp < 0.0) dph_ip1 = dphip1_tmp;
  }
  const Real d2qf = 6.0 * (dph + dph_ip1 - 2.0 * q);
  Real qminus = dph, qplus = dph_ip1;
  const Real dqf_minus = q - qminus;
  const Real dqf_plus = qplus - q;
  {
    const Real qa_tmp = dqf_minus * dqf_plus;
    const Real qb_tmp = (q_ip1 - q) * (q - q_im1);
    qa = d2qc_im1;
    qb = d2qc;
    const Real qc = d2qc_ip1;
    const Real qd = d2qf;
    Real qe = 0.0;
    if (sign_ath(qa) == sign_ath(qb) && sign_ath(qa) == sign_ath(qc) && sign_ath(qa) == sign_ath(qd))
      qe = sign_ath(qd) * min_std(min_std(C2 * fabs(qa), C2 * fabs(qb)), min_std(C2 * fabs(qc), fabs(qd)));
    qa = max_std(fabs(q_im1), fabs(q_im2));
    qb = max_std(max_std(fabs(q), fabs(q_ip1)), fabs(q_ip2));
    Real rho = 0.0;
    if (fabs(qd) > (1.0e-12) * max_std(qa, qb)) rho = fdiv(qe, qd);
    const Real tmp_m = q - rho * dqf_minus;
    const Real tmp_p = q + rho * dqf_plus;
    const Real tmp2_m = q - 2.0 * dqf_plus;
    const Real tmp2_p = q + 2.0 * dqf_minus;
    if ((qa_tmp <= 0.0 || qb_tmp <= 0.0)) {
      if (rho <= (1.0 - (1.0e-12))) {
        qminus = tmp_m;
        qplus = tmp_p;
      }
    } else {
      if (fabs(dqf_minus) >= 2.0 * fabs(dqf_plus)) qminus = tmp2_m;
      if (fabs(dqf_plus) >= 2.0 * fabs(dqf_minus)) qplus = tmp2_p;
    }
  }
  *ql_iph = qplus;
  *qr_imh = qminus;
}

//----------------------------------------------------------------------------------------
// Hydro::GetWeightForCT (hydro/hydro.cpp:154-158)
template <class M>
ATHK_HD Real ct_weight_m(M &m, Real dflx, Real rhol, Real rhor, Real dx, Real dt) {
  Real v_over_c = m.fdiv((1024.0) * dt * dflx, (dx * (rhol + rhor)));
  Real tmp_min = min_std(0.5, v_over_c);
  return 0.5 + max_std(-0.5, tmp_min);
}
ATHK_HD Real ct_weight(Real dflx, Real rhol, Real rhor, Real dx, Real dt) {
  MathIEEE m;
  return ct_weight_m(m, dflx, rhol, rhor, dx, dt);
}

//----------------------------------------------------------------------------------------
// HLLC (hydro/rsolvers/hydro/hllc.cpp:53-175).  wl/wr = (d, vx, vy, vz, p) with vx the
// sweep-normal velocity; f = (mass, mom-x, mom-y, mom-z, energy) in the same rotated frame.
template <class M>
ATHK_HD void hllc_flux_m(M &m, const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  const Real gamma = eos.gamma;
  const Real igm1 = 1.0 / eos.gm1;

  Real cl = m.sqrt(m.div(eos.gamma * wl[IPR], wl[IDN]));
  Real cr = m.sqrt(m.div(eos.gamma * wr[IPR], wr[IDN]));
  Real el = wl[IPR] * igm1 + 0.5 * wl[IDN] * (sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ]));
  Real er = wr[IPR] * igm1 + 0.5 * wr[IDN] * (sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ]));
  Real rhoa = .5 * (wl[IDN] + wr[IDN]);
  Real ca = .5 * (cl + cr);
  Real pmid = .5 * (wl[IPR] + wr[IPR] + (wl[IVX] - wr[IVX]) * rhoa * ca);
  // umid, rhol, rhor are computed by the reference but never reach the flux (PVRS leftovers)

  // ((gamma + 1) / (2 * gamma) is evaluated at run time by the reference too; one division, hoisted)
  const Real gfac = (gamma + 1) / (2 * gamma);
  Real ql = 1.0, qr = 1.0;
  if (!(pmid <= wl[IPR])) ql = m.sqrt(1.0 + gfac * (m.div(pmid, wl[IPR]) - 1.0));
  if (!(pmid <= wr[IPR])) qr = m.sqrt(1.0 + gfac * (m.div(pmid, wr[IPR]) - 1.0));

  Real al = wl[IVX] - cl * ql;
  Real ar = wr[IVX] + cr * qr;

  Real bp = ar > 0.0 ? ar : (ATHK_TINY_NUMBER);
  Real bm = al < 0.0 ? al : -(ATHK_TINY_NUMBER);

  Real vxl = wl[IVX] - al;
  Real vxr = wr[IVX] - ar;

  Real tl = wl[IPR] + vxl * wl[IDN] * wl[IVX];
  Real tr = wr[IPR] + vxr * wr[IDN] * wr[IVX];

  Real ml = wl[IDN] * vxl;
  Real mr = -(wr[IDN] * vxr);

  Real am = m.fdiv((tl - tr), (ml + mr));
  Real cp = m.div((ml * tr + mr * tl), (ml + mr));
  cp = cp > 0.0 ? cp : 0.0;

  vxl = wl[IVX] - bm;
  vxr = wr[IVX] - bp;

  Real fl[5], fr[5];
  fl[IDN] = wl[IDN] * vxl;
  fr[IDN] = wr[IDN] * vxr;
  fl[IVX] = wl[IDN] * wl[IVX] * vxl + wl[IPR];
  fr[IVX] = wr[IDN] * wr[IVX] * vxr + wr[IPR];
  fl[IVY] = wl[IDN] * wl[IVY] * vxl;
  fr[IVY] = wr[IDN] * wr[IVY] * vxr;
  fl[IVZ] = wl[IDN] * wl[IVZ] * vxl;
  fr[IVZ] = wr[IDN] * wr[IVZ] * vxr;
  fl[IEN] = el * vxl + wl[IPR] * wl[IVX];
  fr[IEN] = er * vxr + wr[IPR] * wr[IVX];

  Real sl, sr, sm;
  if (am >= 0.0) {
    sl = m.fdiv(am, (am - bm));
    sr = 0.0;
    sm = m.div(-bm, (am - bm));
  } else {
    sl = 0.0;
    sr = m.fdiv(-am, (bp - am));
    sm = m.div(bp, (bp - am));
  }

  f[IDN] = sl * fl[IDN] + sr * fr[IDN];
  f[IVX] = sl * fl[IVX] + sr * fr[IVX] + sm * cp;
  f[IVY] = sl * fl[IVY] + sr * fr[IVY];
  f[IVZ] = sl * fl[IVZ] + sr * fr[IVZ];
  f[IEN] = sl * fl[IEN] + sr * fr[IEN] + sm * cp * am;
}

ATHK_HD void hllc_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  MathIEEE m;
  hllc_flux_m(m, eos, wl, wr, f);
}

//----------------------------------------------------------------------------------------
// HLLD (hydro/rsolvers/mhd/hlld.cpp:56-365).  wl/wr = (d, vx, vy, vz, p, by, bz) in the sweep
// frame, bxi = normal face field.  f = (mass, mx, my, mz, E, by-flux, bz-flux).
struct Cons7 {
  Real d, mx, my, mz, e, by, bz;
};

template <class M>
ATHK_HD void hlld_flux_m(M &m, const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                         Real f[7]) {
  const Real SMALL_NUMBER = 1.0e-4;
  const Real igm1 = eos.igm1;
  Real spd[5];
  Cons7 ul, ur, ulst, uldst, urdst, urst;

  Real bxsq = bxi * bxi;
  Real pbl = 0.5 * (bxsq + (sqr(wl[IBY]) + sqr(wl[IBZ])));
  Real pbr = 0.5 * (bxsq + (sqr(wr[IBY]) + sqr(wr[IBZ])));

  ul.d = wl[IDN];
  ul.mx = wl[IVX] * ul.d;
  ul.my = wl[IVY] * ul.d;
  ul.mz = wl[IVZ] * ul.d;
  ul.by = wl[IBY];
  ul.bz = wl[IBZ];

  ur.d = wr[IDN];
  ur.mx = wr[IVX] * ur.d;
  ur.my = wr[IVY] * ur.d;
  ur.mz = wr[IVZ] * ur.d;
  ur.by = wr[IBY];
  ur.bz = wr[IBZ];

  // outer wave speeds, Miyoshi & Kusano eq. 67
  Real cfl = fast_speed_m(m, eos, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi);
  Real cfr = fast_speed_m(m, eos, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi);
  spd[0] = min_std(wl[IVX] - cfl, wr[IVX] - cfr);
  spd[4] = max_std(wl[IVX] + cfl, wr[IVX] + cfr);

  Real ptl = wl[IPR] + pbl;
  Real ptr = wr[IPR] + pbr;

  Real sdl = spd[0] - wl[IVX];
  Real sdr = spd[4] - wr[IVX];

  // S_M, eq. 38
  spd[2] = m.fdiv((sdr * ur.mx - sdl * ul.mx + (ptl - ptr)), (sdr * ur.d - sdl * ul.d));

  Real sdml = spd[0] - spd[2];
  Real sdmr = spd[4] - spd[2];
  Real sdml_inv = m.rcp(sdml);
  Real sdmr_inv = m.rcp(sdmr);
  // eq. 43
  ulst.d = ul.d * sdl * sdml_inv;
  urst.d = ur.d * sdr * sdmr_inv;
  Real ulst_d_inv = m.rcp(ulst.d);
  Real urst_d_inv = m.rcp(urst.d);
  Real sqrtdl = m.sqrt(ulst.d);
  Real sqrtdr = m.sqrt(urst.d);

  // eq. 51
  spd[1] = spd[2] - m.fdiv(fabs(bxi), sqrtdl);   // (a field-free normal direction is common: exact 0 numerator)
  spd[3] = spd[2] + m.fdiv(fabs(bxi), sqrtdr);

  // eq. 41
  Real ptstl = ptl + ul.d * sdl * (spd[2] - wl[IVX]);
  Real ptstr = ptr + ur.d * sdr * (spd[2] - wr[IVX]);
  Real ptst = 0.5 * (ptstr + ptstl);

  // left star state, eqs. 39, 44-47
  ulst.mx = ulst.d * spd[2];
  if (fabs(ul.d * sdl * sdml - bxsq) < (SMALL_NUMBER)*ptst) {
    ulst.my = ulst.d * wl[IVY];
    ulst.mz = ulst.d * wl[IVZ];
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    const Real den = (ul.d * sdl * sdml - bxsq), rden = m.div_rcp(den);   // one divisor, two quotients
    Real tmp = m.fdiv_by(bxi * (sdl - sdml), den, rden);
    ulst.my = ulst.d * (wl[IVY] - ul.by * tmp);
    ulst.mz = ulst.d * (wl[IVZ] - ul.bz * tmp);
    tmp = m.div_by((ul.d * sqr(sdl) - bxsq), den, rden);
    ulst.by = ul.by * tmp;
    ulst.bz = ul.bz * tmp;
  }
  // (v.B and the energy of the star states, eq. 48: formed below for the selected side only)

  // right star state
  urst.mx = urst.d * spd[2];
  if (fabs(ur.d * sdr * sdmr - bxsq) < (SMALL_NUMBER)*ptst) {
    urst.my = urst.d * wr[IVY];
    urst.mz = urst.d * wr[IVZ];
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    const Real den = (ur.d * sdr * sdmr - bxsq), rden = m.div_rcp(den);
    Real tmp = m.fdiv_by(bxi * (sdr - sdmr), den, rden);
    urst.my = urst.d * (wr[IVY] - ur.by * tmp);
    urst.mz = urst.d * (wr[IVZ] - ur.bz * tmp);
    tmp = m.div_by((ur.d * sqr(sdr) - bxsq), den, rden);
    urst.by = ur.by * tmp;
    urst.bz = ur.bz * tmp;
  }

  // double-star states, eqs. 59-63
  Real invsumd = m.rcp(sqrtdl + sqrtdr);
  Real bxsig = (bxi > 0.0 ? 1.0 : -1.0);

  uldst.d = ulst.d;
  urdst.d = urst.d;
  uldst.mx = ulst.mx;
  urdst.mx = urst.mx;

  Real tmp = invsumd * (sqrtdl * (ulst.my * ulst_d_inv) + sqrtdr * (urst.my * urst_d_inv) +
                        bxsig * (urst.by - ulst.by));
  uldst.my = uldst.d * tmp;
  urdst.my = urdst.d * tmp;

  tmp = invsumd * (sqrtdl * (ulst.mz * ulst_d_inv) + sqrtdr * (urst.mz * urst_d_inv) +
                   bxsig * (urst.bz - ulst.bz));
  uldst.mz = uldst.d * tmp;
  urdst.mz = urdst.d * tmp;

  tmp = invsumd * (sqrtdl * urst.by + sqrtdr * ulst.by +
                   bxsig * sqrtdl * sqrtdr * ((urst.my * urst_d_inv) - (ulst.my * ulst_d_inv)));
  uldst.by = urdst.by = tmp;

  tmp = invsumd * (sqrtdl * urst.bz + sqrtdr * ulst.bz +
                   bxsig * sqrtdl * sqrtdr * ((urst.mz * urst_d_inv) - (ulst.mz * ulst_d_inv)));
  uldst.bz = urdst.bz = tmp;

  const Real vbdst = spd[2] * bxi + m.fdiv((uldst.my * uldst.by + uldst.mz * uldst.bz), uldst.d);

  // Flux of the region the face lies in (hlld.cpp:323-365): F_L or F_R plus the jump terms
  // S*(U_a - U_b) across the waves between that side and the face.  The reference evaluates
  // F_L, F_R and all four jumps and then picks; only the picked side is evaluated here (same
  // expressions, same association, hence the same bits) -- 2 x 7 fewer values stay live across
  // the star-state algebra, which is what the 96-128 register budget of the sweeps is short of.
  bool left;
  int depth;
  if (spd[0] >= 0.0) { left = true; depth = 0; }
  else if (spd[4] <= 0.0) { left = false; depth = 0; }
  else if (spd[1] >= 0.0) { left = true; depth = 1; }
  else if (spd[3] <= 0.0) { left = false; depth = 1; }
  else if (spd[2] >= 0.0) { left = true; depth = 2; }
  else { left = false; depth = 2; }
  const Real vx = left ? wl[IVX] : wr[IVX], vy = left ? wl[IVY] : wr[IVY], vz = left ? wl[IVZ] : wr[IVZ];
  const Real pt = left ? ptl : ptr;
  const Real s_out = left ? spd[0] : spd[4], s_in = left ? spd[1] : spd[3];
  Cons7 u0, u1, u2, fx;
  u0.d = left ? ul.d : ur.d;         u1.d = left ? ulst.d : urst.d;         u2.d = left ? uldst.d : urdst.d;
  u0.mx = left ? ul.mx : ur.mx;      u1.mx = left ? ulst.mx : urst.mx;      u2.mx = left ? uldst.mx : urdst.mx;
  u0.my = left ? ul.my : ur.my;      u1.my = left ? ulst.my : urst.my;      u2.my = left ? uldst.my : urdst.my;
  u0.mz = left ? ul.mz : ur.mz;      u1.mz = left ? ulst.mz : urst.mz;      u2.mz = left ? uldst.mz : urdst.mz;
  {   // total energy of that side (hlld.cpp:84-98); the other side's is never used
    const Real dd = left ? wl[IDN] : wr[IDN], pp = left ? wl[IPR] : wr[IPR], pb = left ? pbl : pbr;
    const Real ke = 0.5 * dd * (sqr(vx) + (sqr(vy) + sqr(vz)));
    u0.e = pp * igm1 + ke + pb;
  }
  u0.by = left ? ul.by : ur.by;      u1.by = left ? ulst.by : urst.by;      u2.by = left ? uldst.by : urdst.by;
  u0.bz = left ? ul.bz : ur.bz;      u1.bz = left ? ulst.bz : urst.bz;      u2.bz = left ? uldst.bz : urdst.bz;
  {   // star / double-star energies of that side, eqs. 48 and 63 (hlld.cpp:226-230,262-266,301-303)
    const Real sd = left ? sdl : sdr, sdm_inv = left ? sdml_inv : sdmr_inv;
    const Real u1d_inv = left ? ulst_d_inv : urst_d_inv, sqrtd = left ? sqrtdl : sqrtdr;
    const Real vbst = (u1.mx * bxi + (u1.my * u1.by + u1.mz * u1.bz)) * u1d_inv;
    u1.e = (sd * u0.e - pt * vx + ptst * spd[2] + bxi * (vx * bxi + (vy * u0.by + vz * u0.bz) - vbst)) * sdm_inv;
    const Real de = sqrtd * bxsig * (vbst - vbdst);
    u2.e = left ? u1.e - de : u1.e + de;
  }
  // physical flux of that side (hlld.cpp:123-139)
  fx.d = u0.mx;
  fx.mx = u0.mx * vx + pt - bxsq;
  fx.my = u0.my * vx - bxi * u0.by;
  fx.mz = u0.mz * vx - bxi * u0.bz;
  fx.e = vx * (u0.e + pt - bxsq) - bxi * (vy * u0.by + vz * u0.bz);
  fx.by = u0.by * vx - bxi * vy;
  fx.bz = u0.bz * vx - bxi * vz;
#define ATHK_HLLD_PICK(x, out)                                      \
  {                                                                 \
    Real r_ = fx.x;                                                 \
    if (depth >= 1) r_ = fx.x + s_out * (u1.x - u0.x);              \
    if (depth == 2) r_ = r_ + s_in * (u2.x - u1.x);                 \
    out = r_;                                                       \
  }
  ATHK_HLLD_PICK(d, f[IDN])
  ATHK_HLLD_PICK(mx, f[IVX])
  ATHK_HLLD_PICK(my, f[IVY])
  ATHK_HLLD_PICK(mz, f[IVZ])
  ATHK_HLLD_PICK(e, f[IEN])
  ATHK_HLLD_PICK(by, f[IBY])
  ATHK_HLLD_PICK(bz, f[IBZ])
#undef ATHK_HLLD_PICK
}

ATHK_HD void hlld_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                       Real f[7]) {
  MathIEEE m;
  hlld_flux_m(m, eos, wl, wr, bxi, f);
}

}  // namespace athk
#endif  // ATHK_PHYSICS_H_
