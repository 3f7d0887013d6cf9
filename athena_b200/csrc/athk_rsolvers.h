// athk_rsolvers.h -- the reference's alternative Riemann solvers (one is linked per build there,
// Makefile.in:73; selected at run time here through athk_config.rsolver).  Same conventions as
// hllc_flux / hlld_flux in athk_physics.h: states in the sweep frame, adiabatic EOS, bit-exact
// association order.
//   hydro: HLLE  hydro/rsolvers/hydro/hlle.cpp:38-162
//          LLF   hydro/rsolvers/hydro/llf.cpp:34-123
//          Roe   hydro/rsolvers/hydro/roe.cpp:43-175 (+ RoeFlux :185-300; LLF fallback when an
//                intermediate density is negative)
//   MHD:   HLLE  hydro/rsolvers/mhd/hlle_mhd.cpp:25-182
//          LLF   hydro/rsolvers/mhd/llf_mhd.cpp
#ifndef ATHK_RSOLVERS_H_
#define ATHK_RSOLVERS_H_

#include "athk_physics.h"

namespace athk {

enum { RS_HLLC = 0, RS_HLLD = 1, RS_ROE = 2, RS_HLLE = 3, RS_LLF = 4,   // == enum athk_rsolver
       RS_HLLE_ISO = 5, RS_HLLD_ISO = 6,     // isothermal EOS: hlle.cpp's barotropic branch, hlld_iso.cpp
       RS_SR = 7 };                          // special relativity: *_rel.cpp, chosen by eos.sr_solver

//----------------------------------------------------------------------------------------
// Special-relativistic hydro (configure.py -s), GENERAL_RELATIVITY == 0: no frame transformation,
// interface velocity 0.  wl / wr: (rho, u^x, u^y, u^z, pgas) in the sweep's rotated order, u^i the
// four-velocity's spatial part.  eos.sr_solver selects
//   HLLE  hlle_rel.cpp:94-286 (HLLETransforming)
//   HLLC  hllc_rel.cpp:96-336 (HLLCTransforming; Mignone & Bodo 2005)
//   LLF   llf_rel.cpp:93-215  (LLFTransforming)
// which share the left / right states, their wave speeds (SoundSpeedsSR) and fluxes.
ATHK_HD void sr_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  const Real gamma_adi = eos.gamma;
  const Real gamma_prime = gamma_adi / (gamma_adi - 1.0);
  const Real rho_l = wl[IDN], pgas_l = wl[IPR];
  Real u_l[4];
  u_l[0] = sqrt(1.0 + sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ]));
  u_l[1] = wl[IVX]; u_l[2] = wl[IVY]; u_l[3] = wl[IVZ];
  const Real rho_r = wr[IDN], pgas_r = wr[IPR];
  Real u_r[4];
  u_r[0] = sqrt(1.0 + sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ]));
  u_r[1] = wr[IVX]; u_r[2] = wr[IVY]; u_r[3] = wr[IVZ];
  Real lambda_p_l, lambda_m_l, lambda_p_r, lambda_m_r;
  const Real wgas_l = rho_l + gamma_prime * pgas_l;
  sound_speeds_sr(eos, wgas_l, pgas_l, u_l[1] / u_l[0], sqr(u_l[0]), &lambda_p_l, &lambda_m_l);
  const Real wgas_r = rho_r + gamma_prime * pgas_r;
  sound_speeds_sr(eos, wgas_r, pgas_r, u_r[1] / u_r[0], sqr(u_r[0]), &lambda_p_r, &lambda_m_r);
  const Real lambda_l = min_std(lambda_m_l, lambda_m_r);
  const Real lambda_r = max_std(lambda_p_l, lambda_p_r);
  Real cons_l[5], flux_l[5], cons_r[5], flux_r[5];
  cons_l[IDN] = rho_l * u_l[0];
  cons_l[IEN] = wgas_l * u_l[0] * u_l[0] - pgas_l;
  cons_l[IVX] = wgas_l * u_l[1] * u_l[0];
  cons_l[IVY] = wgas_l * u_l[2] * u_l[0];
  cons_l[IVZ] = wgas_l * u_l[3] * u_l[0];
  flux_l[IDN] = rho_l * u_l[1];
  flux_l[IEN] = wgas_l * u_l[0] * u_l[1];
  flux_l[IVX] = wgas_l * u_l[1] * u_l[1] + pgas_l;
  flux_l[IVY] = wgas_l * u_l[2] * u_l[1];
  flux_l[IVZ] = wgas_l * u_l[3] * u_l[1];
  cons_r[IDN] = rho_r * u_r[0];
  cons_r[IEN] = wgas_r * u_r[0] * u_r[0] - pgas_r;
  cons_r[IVX] = wgas_r * u_r[1] * u_r[0];
  cons_r[IVY] = wgas_r * u_r[2] * u_r[0];
  cons_r[IVZ] = wgas_r * u_r[3] * u_r[0];
  flux_r[IDN] = rho_r * u_r[1];
  flux_r[IEN] = wgas_r * u_r[0] * u_r[1];
  flux_r[IVX] = wgas_r * u_r[1] * u_r[1] + pgas_r;
  flux_r[IVY] = wgas_r * u_r[2] * u_r[1];
  flux_r[IVZ] = wgas_r * u_r[3] * u_r[1];
  if (eos.sr_solver == RS_LLF) {
    const Real lambda = max_std(lambda_r, -lambda_l);
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = 0.5 * (flux_l[n] + flux_r[n] - lambda * (cons_r[n] - cons_l[n]));
    return;
  }
  Real lambda_diff_inv = 1.0 / (lambda_r - lambda_l);
  Real flux_hll[5];
#pragma unroll
  for (int n = 0; n < 5; ++n)
    flux_hll[n] = (lambda_r * flux_l[n] - lambda_l * flux_r[n] + lambda_l * lambda_r * (cons_r[n] - cons_l[n])) *
                  lambda_diff_inv;
  if (eos.sr_solver == RS_HLLE) {
    if (lambda_l >= 0.0) {
#pragma unroll
      for (int n = 0; n < 5; ++n) f[n] = flux_l[n];
    } else if (lambda_r <= 0.0) {
#pragma unroll
      for (int n = 0; n < 5; ++n) f[n] = flux_r[n];
    } else {
#pragma unroll
      for (int n = 0; n < 5; ++n) f[n] = flux_hll[n];
    }
    return;
  }
  // HLLC
  Real cons_hll[5];
#pragma unroll
  for (int n = 0; n < 5; ++n)
    cons_hll[n] = (lambda_r * cons_r[n] - lambda_l * cons_l[n] + flux_l[n] - flux_r[n]) * lambda_diff_inv;
  Real lambda_star;
  const Real b = -(cons_hll[IEN] + flux_hll[IVX]);
  if (fabs(flux_hll[IEN]) > ATHK_TINY_NUMBER) {
    lambda_star = -2.0 * cons_hll[IVX] / (b - sqrt(sqr(b) - 4.0 * flux_hll[IEN] * cons_hll[IVX]));
  } else {
    lambda_star = -cons_hll[IVX] / b;
  }
  const Real pgas_star = -flux_hll[IEN] * lambda_star + flux_hll[IVX];
  if (lambda_l >= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = flux_l[n];
  } else if (lambda_r <= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = flux_r[n];
  } else if (lambda_star >= 0.0) {   // aL region (MB2005 16, 14)
    Real cons_lstar[5];
    const Real vx_l_ratio = u_l[1] / u_l[0];
    lambda_diff_inv = 1.0 / (lambda_l - lambda_star);
#pragma unroll
    for (int n = 0; n < 5; ++n) cons_lstar[n] = cons_l[n] * (lambda_l - vx_l_ratio);
    cons_lstar[IEN] += pgas_star * lambda_star - pgas_l * vx_l_ratio;
    cons_lstar[IVX] += pgas_star - pgas_l;
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      cons_lstar[n] *= lambda_diff_inv;
      f[n] = flux_l[n] + lambda_l * (cons_lstar[n] - cons_l[n]);
    }
  } else {                           // c region
    Real cons_rstar[5];
    const Real vx_r_ratio = u_r[1] / u_r[0];
    lambda_diff_inv = 1.0 / (lambda_r - lambda_star);
#pragma unroll
    for (int n = 0; n < 5; ++n) cons_rstar[n] = cons_r[n] * (lambda_r - vx_r_ratio);
    cons_rstar[IEN] += pgas_star * lambda_star - pgas_r * vx_r_ratio;
    cons_rstar[IVX] += pgas_star - pgas_r;
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      cons_rstar[n] *= lambda_diff_inv;
      f[n] = flux_r[n] + lambda_r * (cons_rstar[n] - cons_r[n]);
    }
  }
}

//----------------------------------------------------------------------------------------
ATHK_HD void hlle_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  const Real gm1 = eos.gm1;
  const Real igm1 = 1.0 / gm1;
  Real wroe[5], fl[5], fr[5];
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl * sqrtdr;
  wroe[IVX] = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  wroe[IVY] = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  wroe[IVZ] = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  Real el = wl[IPR] * igm1 + 0.5 * wl[IDN] * (sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ]));
  Real er = wr[IPR] * igm1 + 0.5 * wr[IDN] * (sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ]));
  Real hroe = ((el + wl[IPR]) / sqrtdl + (er + wr[IPR]) / sqrtdr) * isdlpdr;
  Real cl = sound_speed(eos, wl[IDN], wl[IPR]);
  Real cr = sound_speed(eos, wr[IDN], wr[IPR]);
  Real q = hroe - 0.5 * (sqr(wroe[IVX]) + sqr(wroe[IVY]) + sqr(wroe[IVZ]));
  Real a = (q < 0.0) ? 0.0 : sqrt(gm1 * q);
  Real al = min_std((wroe[IVX] - a), (wl[IVX] - cl));
  Real ar = max_std((wroe[IVX] + a), (wr[IVX] + cr));
  Real bp = ar > 0.0 ? ar : 0.0;
  Real bm = al < 0.0 ? al : 0.0;
  Real vxl = wl[IVX] - bm;
  Real vxr = wr[IVX] - bp;
  fl[IDN] = wl[IDN] * vxl;
  fr[IDN] = wr[IDN] * vxr;
  fl[IVX] = wl[IDN] * wl[IVX] * vxl;
  fr[IVX] = wr[IDN] * wr[IVX] * vxr;
  fl[IVY] = wl[IDN] * wl[IVY] * vxl;
  fr[IVY] = wr[IDN] * wr[IVY] * vxr;
  fl[IVZ] = wl[IDN] * wl[IVZ] * vxl;
  fr[IVZ] = wr[IDN] * wr[IVZ] * vxr;
  fl[IVX] += wl[IPR];
  fr[IVX] += wr[IPR];
  fl[IEN] = el * vxl + wl[IPR] * wl[IVX];
  fr[IEN] = er * vxr + wr[IPR] * wr[IVX];
  Real tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
#pragma unroll
  for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp;
}

//----------------------------------------------------------------------------------------
// shared by LLF and Roe: physical fluxes, conserved jump
ATHK_HD void hydro_lr_fluxes(const EosParams &eos, const Real wl[5], const Real wr[5], Real fl[5],
                             Real fr[5], Real du[5], Real *pel, Real *per) {
  const Real gm1 = eos.gm1;
  Real mxl = wl[IDN] * wl[IVX];
  Real mxr = wr[IDN] * wr[IVX];
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl * wl[IVX];
  fr[IVX] = mxr * wr[IVX];
  fl[IVY] = mxl * wl[IVY];
  fr[IVY] = mxr * wr[IVY];
  fl[IVZ] = mxl * wl[IVZ];
  fr[IVZ] = mxr * wr[IVZ];
  Real el = 0.0, er = 0.0;
  if (eos.iso) {   // NON_BAROTROPIC_EOS == 0 (llf.cpp:82-85, roe.cpp:112-115): no energy equation
    fl[IVX] += (eos.iso_cs * eos.iso_cs) * wl[IDN];
    fr[IVX] += (eos.iso_cs * eos.iso_cs) * wr[IDN];
    fl[IEN] = fr[IEN] = 0.0;
  } else {
    el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ]));
    er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ]));
    fl[IVX] += wl[IPR];
    fr[IVX] += wr[IPR];
    fl[IEN] = (el + wl[IPR]) * wl[IVX];
    fr[IEN] = (er + wr[IPR]) * wr[IVX];
  }
  du[IDN] = wr[IDN] - wl[IDN];
  du[IVX] = wr[IDN] * wr[IVX] - wl[IDN] * wl[IVX];
  du[IVY] = wr[IDN] * wr[IVY] - wl[IDN] * wl[IVY];
  du[IVZ] = wr[IDN] * wr[IVZ] - wl[IDN] * wl[IVZ];
  du[IEN] = er - el;
  *pel = el;
  *per = er;
}

ATHK_HD void llf_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  Real fl[5], fr[5], du[5], el, er;
  Real cl = eos.iso ? eos.iso_cs : sound_speed(eos, wl[IDN], wl[IPR]);   // EquationOfState::SoundSpeed
  Real cr = eos.iso ? eos.iso_cs : sound_speed(eos, wr[IDN], wr[IPR]);
  Real a = 0.5 * max_std((fabs(wl[IVX]) + cl), (fabs(wr[IVX]) + cr));
  hydro_lr_fluxes(eos, wl, wr, fl, fr, du, &el, &er);
#pragma unroll
  for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
}

// Roe for isothermal hydrodynamics: roe.cpp:43-175 with NON_BAROTROPIC_EOS == 0, RoeFlux :296-345
// (four waves; NWAVE - 1 = 3 is the last eigenvalue)
ATHK_HD void roe_iso_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  const Real iso_cs = eos.iso_cs;
  Real fl[5], fr[5], du[5], ev[4], el, er;
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  const Real v1 = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  const Real v2 = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  const Real v3 = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  hydro_lr_fluxes(eos, wl, wr, fl, fr, du, &el, &er);
#pragma unroll
  for (int n = 0; n < 4; ++n) f[n] = 0.5 * (fl[n] + fr[n]);
  f[IEN] = 0.0;
  int llf_flag = 0;
  ev[0] = v1 - iso_cs;
  ev[1] = v1;
  ev[2] = v1;
  ev[3] = v1 + iso_cs;
  Real a[4];
  a[0] = du[0] * (0.5 + 0.5 * v1 / iso_cs);
  a[0] -= du[1] * 0.5 / iso_cs;
  a[1] = du[0] * (-v2);
  a[1] += du[2];
  a[2] = du[0] * (-v3);
  a[2] += du[3];
  a[3] = du[0] * (0.5 - 0.5 * v1 / iso_cs);
  a[3] += du[1] * 0.5 / iso_cs;
  Real coeff[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];
  Real dens = wl[IDN] + a[0];
  if (dens < 0.0) llf_flag = 1;
  dens += a[3];
  if (dens < 0.0) llf_flag = 1;
  f[0] += coeff[0];
  f[0] += coeff[3];
  f[1] += coeff[0] * (v1 - iso_cs);
  f[1] += coeff[3] * (v1 + iso_cs);
  f[2] += coeff[0] * v2;
  f[2] += coeff[1];
  f[2] += coeff[3] * v2;
  f[3] += coeff[0] * v3;
  f[3] += coeff[2];
  f[3] += coeff[3] * v3;
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) f[n] = fl[n];
  }
  if (ev[3] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) f[n] = fr[n];
  }
  if (llf_flag != 0) {
    Real am = 0.5 * max_std((fabs(wl[IVX]) + iso_cs), (fabs(wr[IVX]) + iso_cs));
#pragma unroll
    for (int n = 0; n < 4; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - am * du[n];
  }
}

ATHK_HD void roe_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  if (eos.iso) { roe_iso_hydro_flux(eos, wl, wr, f); return; }
  const Real gm1 = eos.gm1;
  Real wroe[5], fl[5], fr[5], du[5], ev[5], el, er;
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl * sqrtdr;
  wroe[IVX] = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  wroe[IVY] = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  wroe[IVZ] = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  hydro_lr_fluxes(eos, wl, wr, fl, fr, du, &el, &er);
  wroe[IPR] = ((el + wl[IPR]) / sqrtdl + (er + wr[IPR]) / sqrtdr) * isdlpdr;   // Roe enthalpy
#pragma unroll
  for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]);
  int llf_flag = 0;
  {   // RoeFlux (roe.cpp:185-270)
    const Real v1 = wroe[IVX], v2 = wroe[IVY], v3 = wroe[IVZ];
    const Real h = wroe[IPR];
    Real vsq = v1 * v1 + v2 * v2 + v3 * v3;
    Real q = h - 0.5 * vsq;
    Real cs_sq = (q < 0.0) ? (ATHK_TINY_NUMBER) : gm1 * q;
    Real cs = sqrt(cs_sq);
    ev[0] = v1 - cs;
    ev[1] = v1;
    ev[2] = v1;
    ev[3] = v1;
    ev[4] = v1 + cs;
    Real a[5];
    Real na = 0.5 / cs_sq;
    a[0] = du[0] * (0.5 * gm1 * vsq + v1 * cs);
    a[0] -= du[1] * (gm1 * v1 + cs);
    a[0] -= du[2] * gm1 * v2;
    a[0] -= du[3] * gm1 * v3;
    a[0] += du[4] * gm1;
    a[0] *= na;
    a[1] = du[0] * (-v2);
    a[1] += du[2];
    a[2] = du[0] * (-v3);
    a[2] += du[3];
    Real qa = gm1 / cs_sq;
    a[3] = du[0] * (1.0 - na * gm1 * vsq);
    a[3] += du[1] * qa * v1;
    a[3] += du[2] * qa * v2;
    a[3] += du[3] * qa * v3;
    a[3] -= du[4] * qa;
    a[4] = du[0] * (0.5 * gm1 * vsq - v1 * cs);
    a[4] -= du[1] * (gm1 * v1 - cs);
    a[4] -= du[2] * gm1 * v2;
    a[4] -= du[3] * gm1 * v3;
    a[4] += du[4] * gm1;
    a[4] *= na;
    Real coeff[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];
    Real dens = wl[IDN] + a[0];
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    f[0] += coeff[0];
    f[0] += coeff[3];
    f[0] += coeff[4];
    f[1] += coeff[0] * (v1 - cs);
    f[1] += coeff[3] * v1;
    f[1] += coeff[4] * (v1 + cs);
    f[2] += coeff[0] * v2;
    f[2] += coeff[1];
    f[2] += coeff[3] * v2;
    f[2] += coeff[4] * v2;
    f[3] += coeff[0] * v3;
    f[3] += coeff[2];
    f[3] += coeff[3] * v3;
    f[3] += coeff[4] * v3;
    f[4] += coeff[0] * (h - v1 * cs);
    f[4] += coeff[1] * v2;
    f[4] += coeff[2] * v3;
    f[4] += coeff[3] * 0.5 * vsq;
    f[4] += coeff[4] * (h + v1 * cs);
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = fl[n];
  }
  if (ev[4] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = fr[n];
  }
  if (llf_flag != 0) {
    Real cl = sound_speed(eos, wl[IDN], wl[IPR]);
    Real cr = sound_speed(eos, wr[IDN], wr[IPR]);
    Real a = 0.5 * max_std((fabs(wl[IVX]) + cl), (fabs(wr[IVX]) + cr));
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
}

//========================================================================================
// MHD.  wl/wr = (d, vx, vy, vz, p, by, bz), bxi = face-normal field; f[IBY], f[IBZ] are the
// induction fluxes the caller turns into EMFs (-f[IBY] -> ey, f[IBZ] -> ez).

// shared by LLF and Roe (llf_mhd.cpp:78-125 == roe_mhd.cpp:108-160)
ATHK_HD void mhd_lr_fluxes(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                           Real fl[7], Real fr[7], Real du[7], Real *pel, Real *per, Real *ppbl,
                           Real *ppbr) {
  const Real gm1 = eos.gm1;
  Real mxl = wl[IDN] * wl[IVX];
  Real mxr = wr[IDN] * wr[IVX];
  Real pbl = 0.5 * (bxi * bxi + sqr(wl[IBY]) + sqr(wl[IBZ]));
  Real pbr = 0.5 * (bxi * bxi + sqr(wr[IBY]) + sqr(wr[IBZ]));
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl * wl[IVX] + pbl - sqr(bxi);
  fr[IVX] = mxr * wr[IVX] + pbr - sqr(bxi);
  fl[IVY] = mxl * wl[IVY] - bxi * wl[IBY];
  fr[IVY] = mxr * wr[IVY] - bxi * wr[IBY];
  fl[IVZ] = mxl * wl[IVZ] - bxi * wl[IBZ];
  fr[IVZ] = mxr * wr[IVZ] - bxi * wr[IBZ];
  Real el = 0.0, er = 0.0;
  if (eos.iso) {   // llf_mhd.cpp:106-109, roe_mhd.cpp:132-135
    fl[IVX] += (eos.iso_cs * eos.iso_cs) * wl[IDN];
    fr[IVX] += (eos.iso_cs * eos.iso_cs) * wr[IDN];
    fl[IEN] = fr[IEN] = 0.0;
  } else {
    el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ])) + pbl;
    er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ])) + pbr;
    fl[IVX] += wl[IPR];
    fr[IVX] += wr[IPR];
    fl[IEN] = (el + wl[IPR] + pbl - bxi * bxi) * wl[IVX];
    fr[IEN] = (er + wr[IPR] + pbr - bxi * bxi) * wr[IVX];
    fl[IEN] -= bxi * (wl[IBY] * wl[IVY] + wl[IBZ] * wl[IVZ]);
    fr[IEN] -= bxi * (wr[IBY] * wr[IVY] + wr[IBZ] * wr[IVZ]);
  }
  fl[IBY] = wl[IBY] * wl[IVX] - bxi * wl[IVY];
  fr[IBY] = wr[IBY] * wr[IVX] - bxi * wr[IVY];
  fl[IBZ] = wl[IBZ] * wl[IVX] - bxi * wl[IVZ];
  fr[IBZ] = wr[IBZ] * wr[IVX] - bxi * wr[IVZ];
  du[IDN] = wr[IDN] - wl[IDN];
  du[IVX] = wr[IDN] * wr[IVX] - wl[IDN] * wl[IVX];
  du[IVY] = wr[IDN] * wr[IVY] - wl[IDN] * wl[IVY];
  du[IVZ] = wr[IDN] * wr[IVZ] - wl[IDN] * wl[IVZ];
  du[IEN] = er - el;
  du[IBY] = wr[IBY] - wl[IBY];
  du[IBZ] = wr[IBZ] - wl[IBZ];
  *pel = el;
  *per = er;
  *ppbl = pbl;
  *ppbr = pbr;
}

ATHK_HD void llf_mhd_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                          Real f[7]) {
  Real fl[7], fr[7], du[7], el, er, pbl, pbr;
  Real cfl = eos.iso ? fast_speed_iso(eos, wl[IDN], wl[IBY], wl[IBZ], bxi) : fast_speed(eos, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi);
  Real cfr = eos.iso ? fast_speed_iso(eos, wr[IDN], wr[IBY], wr[IBZ], bxi) : fast_speed(eos, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi);
  Real a = 0.5 * max_std((fabs(wl[IVX]) + cfl), (fabs(wr[IVX]) + cfr));
  mhd_lr_fluxes(eos, wl, wr, bxi, fl, fr, du, &el, &er, &pbl, &pbr);
#pragma unroll
  for (int n = 0; n < 7; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
}

ATHK_HD void hlle_mhd_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                           Real f[7]) {
  const Real gm1 = eos.gm1;
  Real wroe[7], fl[7], fr[7];
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl * sqrtdr;
  wroe[IVX] = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  wroe[IVY] = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  wroe[IVZ] = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  wroe[IBY] = (sqrtdr * wl[IBY] + sqrtdl * wr[IBY]) * isdlpdr;
  wroe[IBZ] = (sqrtdr * wl[IBZ] + sqrtdl * wr[IBZ]) * isdlpdr;
  Real x = 0.5 * (sqr(wl[IBY] - wr[IBY]) + sqr(wl[IBZ] - wr[IBZ])) / (sqr(sqrtdl + sqrtdr));
  Real y = 0.5 * (wl[IDN] + wr[IDN]) / wroe[IDN];
  Real pbl = 0.5 * (bxi * bxi + sqr(wl[IBY]) + sqr(wl[IBZ]));
  Real pbr = 0.5 * (bxi * bxi + sqr(wr[IBY]) + sqr(wr[IBZ]));
  Real el = 0.0, er = 0.0, hroe = 0.0;
  if (!eos.iso) {
    el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (sqr(wl[IVX]) + sqr(wl[IVY]) + sqr(wl[IVZ])) + pbl;
    er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (sqr(wr[IVX]) + sqr(wr[IVY]) + sqr(wr[IVZ])) + pbr;
    hroe = ((el + wl[IPR] + pbl) / sqrtdl + (er + wr[IPR] + pbr) / sqrtdr) * isdlpdr;
  }
  Real cl = eos.iso ? fast_speed_iso(eos, wl[IDN], wl[IBY], wl[IBZ], bxi) : fast_speed(eos, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi);
  Real cr = eos.iso ? fast_speed_iso(eos, wr[IDN], wr[IBY], wr[IBZ], bxi) : fast_speed(eos, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi);
  Real btsq = sqr(wroe[IBY]) + sqr(wroe[IBZ]);
  Real vaxsq = bxi * bxi / wroe[IDN];
  Real bt_starsq, twid_asq;
  if (eos.iso) {   // hlle_mhd.cpp:103-106
    bt_starsq = btsq * y;
    twid_asq = eos.iso_cs * eos.iso_cs + x;
  } else {
    bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
    Real hp = hroe - (vaxsq + btsq / wroe[IDN]);
    Real vsq = sqr(wroe[IVX]) + sqr(wroe[IVY]) + sqr(wroe[IVZ]);
    twid_asq = max_std((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), 0.0);
  }
  Real ct2 = bt_starsq / wroe[IDN];
  Real tsum = vaxsq + ct2 + twid_asq;
  Real tdif = vaxsq + ct2 - twid_asq;
  Real cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  Real cfsq = 0.5 * (tsum + cf2_cs2);
  Real a = sqrt(cfsq);
  Real al = min_std((wroe[IVX] - a), (wl[IVX] - cl));
  Real ar = max_std((wroe[IVX] + a), (wr[IVX] + cr));
  Real bp = ar > 0.0 ? ar : 0.0;
  Real bm = al < 0.0 ? al : 0.0;
  Real vxl = wl[IVX] - bm;
  Real vxr = wr[IVX] - bp;
  fl[IDN] = wl[IDN] * vxl;
  fr[IDN] = wr[IDN] * vxr;
  fl[IVX] = wl[IDN] * wl[IVX] * vxl + pbl - sqr(bxi);
  fr[IVX] = wr[IDN] * wr[IVX] * vxr + pbr - sqr(bxi);
  fl[IVY] = wl[IDN] * wl[IVY] * vxl - bxi * wl[IBY];
  fr[IVY] = wr[IDN] * wr[IVY] * vxr - bxi * wr[IBY];
  fl[IVZ] = wl[IDN] * wl[IVZ] * vxl - bxi * wl[IBZ];
  fr[IVZ] = wr[IDN] * wr[IVZ] * vxr - bxi * wr[IBZ];
  if (eos.iso) {   // hlle_mhd.cpp:147-150
    fl[IVX] += (eos.iso_cs * eos.iso_cs) * wl[IDN];
    fr[IVX] += (eos.iso_cs * eos.iso_cs) * wr[IDN];
    fl[IEN] = fr[IEN] = 0.0;
  } else {
    fl[IVX] += wl[IPR];
    fr[IVX] += wr[IPR];
    fl[IEN] = el * vxl + wl[IVX] * (wl[IPR] + pbl - bxi * bxi);
    fr[IEN] = er * vxr + wr[IVX] * (wr[IPR] + pbr - bxi * bxi);
    fl[IEN] -= bxi * (wl[IBY] * wl[IVY] + wl[IBZ] * wl[IVZ]);
    fr[IEN] -= bxi * (wr[IBY] * wr[IVY] + wr[IBZ] * wr[IVZ]);
  }
  fl[IBY] = wl[IBY] * vxl - bxi * wl[IVY];
  fr[IBY] = wr[IBY] * vxr - bxi * wr[IVY];
  fl[IBZ] = wl[IBZ] * vxl - bxi * wl[IVZ];
  fr[IBZ] = wr[IBZ] * vxr - bxi * wr[IVZ];
  Real tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
#pragma unroll
  for (int n = 0; n < 7; ++n) f[n] = 0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp;
}

// roe_mhd.cpp:43-210 with RoeFlux :245-470 (adiabatic branch).  The LLF fallback of the reference
// leaves the induction components at their Roe values (roe_mhd.cpp:193-205); so does this.
// Roe for isothermal MHD: roe_mhd.cpp:43-210 with NON_BAROTROPIC_EOS == 0 and RoeFlux's isothermal
// eigensystem (:255-350, :496-640; six waves).  The reference's wave vector has no energy slot there
// (IBY = 4, IBZ = 5); here the field components stay in slots 5 and 6 and slot 4 carries a zero flux.
// As in the adiabatic solver the LLF fallback replaces the hydro components only (:186-197).
ATHK_HD void roe_iso_mhd_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                              Real f[7]) {
  const Real iso_cs = eos.iso_cs;
  Real fl[7], fr[7], du[7], ev[6], el, er, pbl, pbr;
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  const Real d = sqrtdl * sqrtdr;
  const Real v1 = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  const Real v2 = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  const Real v3 = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  const Real b2 = (sqrtdr * wl[IBY] + sqrtdl * wr[IBY]) * isdlpdr;
  const Real b3 = (sqrtdr * wl[IBZ] + sqrtdl * wr[IBZ]) * isdlpdr;
  const Real b1 = bxi;
  const Real x = 0.5 * (sqr(wl[IBY] - wr[IBY]) + sqr(wl[IBZ] - wr[IBZ])) / (sqr(sqrtdl + sqrtdr));
  const Real y = 0.5 * (wl[IDN] + wr[IDN]) / d;
  mhd_lr_fluxes(eos, wl, wr, bxi, fl, fr, du, &el, &er, &pbl, &pbr);
#pragma unroll
  for (int n = 0; n < 7; ++n) f[n] = 0.5 * (fl[n] + fr[n]);
  f[IEN] = 0.0;
  int llf_flag = 0;
  {
    Real di = 1.0 / d;
    Real btsq = b2 * b2 + b3 * b3;
    Real vaxsq = b1 * b1 * di;
    Real bt_starsq = btsq * y;
    Real twid_csq = (iso_cs * iso_cs) + x;
    Real ct2 = bt_starsq * di;
    Real tsum = vaxsq + ct2 + twid_csq;
    Real tdif = vaxsq + ct2 - twid_csq;
    Real cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_csq * ct2);
    Real cfsq = 0.5 * (tsum + cf2_cs2);
    Real cf = sqrt(cfsq);
    Real cssq = twid_csq * vaxsq / cfsq;
    Real cs = sqrt(cssq);
    Real bt = sqrt(btsq);
    Real bt_star = sqrt(bt_starsq);
    Real bet2 = 0.0;
    Real bet3 = 0.0;
    if (bt != 0.0) {
      bet2 = b2 / bt;
      bet3 = b3 / bt;
    }
    Real bet2_star = bet2 / sqrt(y);
    Real bet3_star = bet3 / sqrt(y);
    Real bet_starsq = bet2_star * bet2_star + bet3_star * bet3_star;
    Real q2_star = 0.0;
    Real q3_star = 0.0;
    if (bet_starsq != 0.0) {
      q2_star = bet2_star / bet_starsq;
      q3_star = bet3_star / bet_starsq;
    }
    Real alpha_f, alpha_s;
    if ((cfsq - cssq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else if ((twid_csq - cssq) <= 0.0) {
      alpha_f = 0.0;
      alpha_s = 1.0;
    } else if ((cfsq - twid_csq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else {
      alpha_f = sqrt((twid_csq - cssq) / (cfsq - cssq));
      alpha_s = sqrt((cfsq - twid_csq) / (cfsq - cssq));
    }
    Real sqrtd = sqrt(d);
    Real isqrtd = 1.0 / sqrtd;
    Real s = sign_ath(b1);
    Real twid_c = sqrt(twid_csq);
    Real qf = cf * alpha_f * s;
    Real qs = cs * alpha_s * s;
    Real af_prime = twid_c * alpha_f * isqrtd;
    Real as_prime = twid_c * alpha_s * isqrtd;
    Real vqstr = (v2 * q2_star + v3 * q3_star);
    Real vax = sqrt(vaxsq);
    Real norm = 0.5 / twid_csq;
    Real cff = norm * alpha_f * cf;
    Real css = norm * alpha_s * cs;
    Real qf_hat = qf * norm;
    Real qs_hat = qs * norm;
    Real af = norm * af_prime * d;
    Real as = norm * as_prime * d;
    Real afpb = norm * af_prime * bt_star;
    Real aspb = norm * as_prime * bt_star;
    (void)isqrtd;

    ev[0] = v1 - cf;
    ev[1] = v1 - vax;
    ev[2] = v1 - cs;
    ev[3] = v1 + cs;
    ev[4] = v1 + vax;
    ev[5] = v1 + cf;

    // (du of the reference: 0..3 hydro, 4 = By, 5 = Bz)
    const Real du4 = du[IBY], du5 = du[IBZ];
    Real a[6];
    a[0] = du[0] * (cff * (cf + v1) - qs_hat * vqstr - aspb);
    a[0] -= du[1] * cff;
    a[0] += du[2] * qs_hat * q2_star;
    a[0] += du[3] * qs_hat * q3_star;
    a[0] += du4 * as * q2_star;
    a[0] += du5 * as * q3_star;

    a[1] = du[0] * (v2 * bet3 - v3 * bet2);
    a[1] -= du[2] * bet3;
    a[1] += du[3] * bet2;
    a[1] -= du4 * sqrtd * bet3 * s;
    a[1] += du5 * sqrtd * bet2 * s;
    a[1] *= 0.5;

    a[2] = du[0] * (css * (cs + v1) + qf_hat * vqstr + afpb);
    a[2] -= du[1] * css;
    a[2] -= du[2] * qf_hat * q2_star;
    a[2] -= du[3] * qf_hat * q3_star;
    a[2] -= du4 * af * q2_star;
    a[2] -= du5 * af * q3_star;

    a[3] = du[0] * (css * (cs - v1) - qf_hat * vqstr + afpb);
    a[3] += du[1] * css;
    a[3] += du[2] * qf_hat * q2_star;
    a[3] += du[3] * qf_hat * q3_star;
    a[3] -= du4 * af * q2_star;
    a[3] -= du5 * af * q3_star;

    a[4] = du[0] * (v3 * bet2 - v2 * bet3);
    a[4] += du[2] * bet3;
    a[4] -= du[3] * bet2;
    a[4] -= du4 * sqrtd * bet3 * s;
    a[4] += du5 * sqrtd * bet2 * s;
    a[4] *= 0.5;

    a[5] = du[0] * (cff * (cf - v1) + qs_hat * vqstr - aspb);
    a[5] += du[1] * cff;
    a[5] -= du[2] * qs_hat * q2_star;
    a[5] -= du[3] * qs_hat * q3_star;
    a[5] += du4 * as * q2_star;
    a[5] += du5 * as * q3_star;

    Real coeff[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];

    Real dens = wl[IDN] + a[0] * alpha_f;
    if (dens < 0.0) llf_flag = 1;
    dens += a[2] * alpha_s;
    if (dens < 0.0) llf_flag = 1;
    dens += a[3] * alpha_s;
    if (dens < 0.0) llf_flag = 1;

    f[0] += coeff[0] * alpha_f;
    f[0] += coeff[2] * alpha_s;
    f[0] += coeff[3] * alpha_s;
    f[0] += coeff[5] * alpha_f;

    f[1] += coeff[0] * alpha_f * (v1 - cf);
    f[1] += coeff[2] * alpha_s * (v1 - cs);
    f[1] += coeff[3] * alpha_s * (v1 + cs);
    f[1] += coeff[5] * alpha_f * (v1 + cf);

    f[2] += coeff[0] * (alpha_f * v2 + qs * bet2_star);
    f[2] -= coeff[1] * bet3;
    f[2] += coeff[2] * (alpha_s * v2 - qf * bet2_star);
    f[2] += coeff[3] * (alpha_s * v2 + qf * bet2_star);
    f[2] += coeff[4] * bet3;
    f[2] += coeff[5] * (alpha_f * v2 - qs * bet2_star);

    f[3] += coeff[0] * (alpha_f * v3 + qs * bet3_star);
    f[3] += coeff[1] * bet2;
    f[3] += coeff[2] * (alpha_s * v3 - qf * bet3_star);
    f[3] += coeff[3] * (alpha_s * v3 + qf * bet3_star);
    f[3] -= coeff[4] * bet2;
    f[3] += coeff[5] * (alpha_f * v3 - qs * bet3_star);

    f[IBY] += coeff[0] * as_prime * bet2_star;
    f[IBY] -= coeff[1] * bet3 * s / sqrtd;
    f[IBY] -= coeff[2] * af_prime * bet2_star;
    f[IBY] -= coeff[3] * af_prime * bet2_star;
    f[IBY] -= coeff[4] * bet3 * s / sqrtd;
    f[IBY] += coeff[5] * as_prime * bet2_star;

    f[IBZ] += coeff[0] * as_prime * bet3_star;
    f[IBZ] += coeff[1] * bet2 * s / sqrtd;
    f[IBZ] -= coeff[2] * af_prime * bet3_star;
    f[IBZ] -= coeff[3] * af_prime * bet3_star;
    f[IBZ] += coeff[4] * bet2 * s / sqrtd;
    f[IBZ] += coeff[5] * as_prime * bet3_star;
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) f[n] = fl[n];
  }
  if (ev[5] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) f[n] = fr[n];
  }
  if (llf_flag != 0) {
    Real cfl = fast_speed_iso(eos, wl[IDN], wl[IBY], wl[IBZ], bxi);
    Real cfr = fast_speed_iso(eos, wr[IDN], wr[IBY], wr[IBZ], bxi);
    Real am = 0.5 * max_std((fabs(wl[IVX]) + cfl), (fabs(wr[IVX]) + cfr));
#pragma unroll
    for (int n = 0; n < 4; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - am * du[n];
  }
  f[IEN] = 0.0;
}

ATHK_HD void roe_mhd_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                          Real f[7]) {
  if (eos.iso) { roe_iso_mhd_flux(eos, wl, wr, bxi, f); return; }
  const Real gm1 = eos.gm1;
  Real wroe[7], fl[7], fr[7], du[7], ev[7], el, er, pbl, pbr;
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl * sqrtdr;
  wroe[IVX] = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  wroe[IVY] = (sqrtdl * wl[IVY] + sqrtdr * wr[IVY]) * isdlpdr;
  wroe[IVZ] = (sqrtdl * wl[IVZ] + sqrtdr * wr[IVZ]) * isdlpdr;
  wroe[IBY] = (sqrtdr * wl[IBY] + sqrtdl * wr[IBY]) * isdlpdr;
  wroe[IBZ] = (sqrtdr * wl[IBZ] + sqrtdl * wr[IBZ]) * isdlpdr;
  const Real x = 0.5 * (sqr(wl[IBY] - wr[IBY]) + sqr(wl[IBZ] - wr[IBZ])) / (sqr(sqrtdl + sqrtdr));
  const Real y = 0.5 * (wl[IDN] + wr[IDN]) / wroe[IDN];
  mhd_lr_fluxes(eos, wl, wr, bxi, fl, fr, du, &el, &er, &pbl, &pbr);
  wroe[IPR] = ((el + wl[IPR] + pbl) / sqrtdl + (er + wr[IPR] + pbr) / sqrtdr) * isdlpdr;
#pragma unroll
  for (int n = 0; n < 7; ++n) f[n] = 0.5 * (fl[n] + fr[n]);
  int llf_flag = 0;
  {
    const Real d = wroe[IDN], v1 = wroe[IVX], v2 = wroe[IVY], v3 = wroe[IVZ];
    const Real b1 = bxi, b2 = wroe[IBY], b3 = wroe[IBZ];
    Real di = 1.0 / d;
    Real btsq = b2 * b2 + b3 * b3;
    Real vaxsq = b1 * b1 * di;
    Real vsq = v1 * v1 + v2 * v2 + v3 * v3;
    Real hp = wroe[IPR] - (vaxsq + btsq * di);
    Real bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
    Real twid_csq = max_std((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), ATHK_TINY_NUMBER);
    Real ct2 = bt_starsq * di;
    Real tsum = vaxsq + ct2 + twid_csq;
    Real tdif = vaxsq + ct2 - twid_csq;
    Real cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_csq * ct2);
    Real cfsq = 0.5 * (tsum + cf2_cs2);
    Real cf = sqrt(cfsq);
    Real cssq = twid_csq * vaxsq / cfsq;
    Real cs = sqrt(cssq);
    Real bt = sqrt(btsq);
    Real bt_star = sqrt(bt_starsq);
    Real bet2 = 0.0;
    Real bet3 = 0.0;
    if (bt != 0.0) {
      bet2 = b2 / bt;
      bet3 = b3 / bt;
    }
    Real bet2_star = bet2 / sqrt(gm1 - (gm1 - 1.0) * y);
    Real bet3_star = bet3 / sqrt(gm1 - (gm1 - 1.0) * y);
    Real bet_starsq = bet2_star * bet2_star + bet3_star * bet3_star;
    Real vbet = v2 * bet2_star + v3 * bet3_star;
    Real q2_star = 0.0;
    Real q3_star = 0.0;
    if (bet_starsq != 0.0) {
      q2_star = bet2_star / bet_starsq;
      q3_star = bet3_star / bet_starsq;
    }
    Real alpha_f, alpha_s;
    if ((cfsq - cssq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else if ((twid_csq - cssq) <= 0.0) {
      alpha_f = 0.0;
      alpha_s = 1.0;
    } else if ((cfsq - twid_csq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else {
      alpha_f = sqrt((twid_csq - cssq) / (cfsq - cssq));
      alpha_s = sqrt((cfsq - twid_csq) / (cfsq - cssq));
    }
    Real sqrtd = sqrt(d);
    Real isqrtd = 1.0 / sqrtd;
    Real s = sign_ath(b1);
    Real twid_c = sqrt(twid_csq);
    Real qf = cf * alpha_f * s;
    Real qs = cs * alpha_s * s;
    Real af_prime = twid_c * alpha_f * isqrtd;
    Real as_prime = twid_c * alpha_s * isqrtd;
    Real afpbb = af_prime * bt_star * bet_starsq;
    Real aspbb = as_prime * bt_star * bet_starsq;
    Real vqstr = (v2 * q2_star + v3 * q3_star);
    Real vax = sqrt(vaxsq);
    Real norm = 0.5 / twid_csq;
    Real cff = norm * alpha_f * cf;
    Real css = norm * alpha_s * cs;
    Real qf_hat = qf * norm;
    Real qs_hat = qs * norm;
    Real af = norm * af_prime * d;
    Real as = norm * as_prime * d;
    Real afpb = norm * af_prime * bt_star;
    Real aspb = norm * as_prime * bt_star;

    ev[0] = v1 - cf;
    ev[1] = v1 - vax;
    ev[2] = v1 - cs;
    ev[3] = v1;
    ev[4] = v1 + cs;
    ev[5] = v1 + vax;
    ev[6] = v1 + cf;

    Real a[7];
    Real alpha_f_bar = alpha_f * gm1 * norm;
    Real alpha_s_bar = alpha_s * gm1 * norm;
    Real gm1a = gm1 / twid_csq;

    a[0] = du[0] * (alpha_f_bar * (vsq - hp) + cff * (cf + v1) - qs_hat * vqstr - aspb);
    a[0] -= du[1] * (alpha_f_bar * v1 + cff);
    a[0] -= du[2] * (alpha_f_bar * v2 - qs_hat * q2_star);
    a[0] -= du[3] * (alpha_f_bar * v3 - qs_hat * q3_star);
    a[0] += du[4] * alpha_f_bar;
    a[0] += du[5] * (as * q2_star - alpha_f_bar * b2);
    a[0] += du[6] * (as * q3_star - alpha_f_bar * b3);

    a[1] = du[0] * (v2 * bet3 - v3 * bet2);
    a[1] -= du[2] * bet3;
    a[1] += du[3] * bet2;
    a[1] -= du[5] * sqrtd * bet3 * s;
    a[1] += du[6] * sqrtd * bet2 * s;
    a[1] *= 0.5;

    a[2] = du[0] * (alpha_s_bar * (vsq - hp) + css * (cs + v1) + qf_hat * vqstr + afpb);
    a[2] -= du[1] * (alpha_s_bar * v1 + css);
    a[2] -= du[2] * (alpha_s_bar * v2 + qf_hat * q2_star);
    a[2] -= du[3] * (alpha_s_bar * v3 + qf_hat * q3_star);
    a[2] += du[4] * alpha_s_bar;
    a[2] -= du[5] * (af * q2_star + alpha_s_bar * b2);
    a[2] -= du[6] * (af * q3_star + alpha_s_bar * b3);

    a[3] = du[0] * (1.0 - gm1a * (0.5 * vsq - (gm1 - 1.0) * x / gm1));
    a[3] += du[1] * gm1a * v1;
    a[3] += du[2] * gm1a * v2;
    a[3] += du[3] * gm1a * v3;
    a[3] -= du[4] * gm1a;
    a[3] += du[5] * gm1a * b2;
    a[3] += du[6] * gm1a * b3;

    a[4] = du[0] * (alpha_s_bar * (vsq - hp) + css * (cs - v1) - qf_hat * vqstr + afpb);
    a[4] -= du[1] * (alpha_s_bar * v1 - css);
    a[4] -= du[2] * (alpha_s_bar * v2 - qf_hat * q2_star);
    a[4] -= du[3] * (alpha_s_bar * v3 - qf_hat * q3_star);
    a[4] += du[4] * alpha_s_bar;
    a[4] -= du[5] * (af * q2_star + alpha_s_bar * b2);
    a[4] -= du[6] * (af * q3_star + alpha_s_bar * b3);

    a[5] = du[0] * (v3 * bet2 - v2 * bet3);
    a[5] += du[2] * bet3;
    a[5] -= du[3] * bet2;
    a[5] -= du[5] * sqrtd * bet3 * s;
    a[5] += du[6] * sqrtd * bet2 * s;
    a[5] *= 0.5;

    a[6] = du[0] * (alpha_f_bar * (vsq - hp) + cff * (cf - v1) + qs_hat * vqstr - aspb);
    a[6] -= du[1] * (alpha_f_bar * v1 - cff);
    a[6] -= du[2] * (alpha_f_bar * v2 + qs_hat * q2_star);
    a[6] -= du[3] * (alpha_f_bar * v3 + qs_hat * q3_star);
    a[6] += du[4] * alpha_f_bar;
    a[6] += du[5] * (as * q2_star - alpha_f_bar * b2);
    a[6] += du[6] * (as * q3_star - alpha_f_bar * b3);

    Real coeff[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];

    Real dens = wl[IDN] + a[0] * alpha_f;
    if (dens < 0.0) llf_flag = 1;
    dens += a[2] * alpha_s;
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    dens += a[4] * alpha_s;
    if (dens < 0.0) llf_flag = 1;

    f[0] += coeff[0] * alpha_f;
    f[0] += coeff[2] * alpha_s;
    f[0] += coeff[3];
    f[0] += coeff[4] * alpha_s;
    f[0] += coeff[6] * alpha_f;

    f[1] += coeff[0] * (alpha_f * (v1 - cf));
    f[1] += coeff[2] * (alpha_s * (v1 - cs));
    f[1] += coeff[3] * v1;
    f[1] += coeff[4] * (alpha_s * (v1 + cs));
    f[1] += coeff[6] * (alpha_f * (v1 + cf));

    f[2] += coeff[0] * (alpha_f * v2 + qs * bet2_star);
    f[2] -= coeff[1] * bet3;
    f[2] += coeff[2] * (alpha_s * v2 - qf * bet2_star);
    f[2] += coeff[3] * v2;
    f[2] += coeff[4] * (alpha_s * v2 + qf * bet2_star);
    f[2] += coeff[5] * bet3;
    f[2] += coeff[6] * (alpha_f * v2 - qs * bet2_star);

    f[3] += coeff[0] * (alpha_f * v3 + qs * bet3_star);
    f[3] += coeff[1] * bet2;
    f[3] += coeff[2] * (alpha_s * v3 - qf * bet3_star);
    f[3] += coeff[3] * v3;
    f[3] += coeff[4] * (alpha_s * v3 + qf * bet3_star);
    f[3] -= coeff[5] * bet2;
    f[3] += coeff[6] * (alpha_f * v3 - qs * bet3_star);

    f[4] += coeff[0] * (alpha_f * (hp - v1 * cf) + qs * vbet + aspbb);
    f[4] -= coeff[1] * (v2 * bet3 - v3 * bet2);
    f[4] += coeff[2] * (alpha_s * (hp - v1 * cs) - qf * vbet - afpbb);
    f[4] += coeff[3] * (0.5 * vsq + (gm1 - 1.0) * x / gm1);
    f[4] += coeff[4] * (alpha_s * (hp + v1 * cs) + qf * vbet - afpbb);
    f[4] += coeff[5] * (v1 * bet3 - v3 * bet2);
    f[4] += coeff[6] * (alpha_f * (hp + v1 * cf) - qs * vbet + aspbb);

    f[5] += coeff[0] * as_prime * bet2_star;
    f[5] -= coeff[1] * bet3 * s * isqrtd;
    f[5] -= coeff[2] * af_prime * bet2_star;
    f[5] -= coeff[4] * af_prime * bet2_star;
    f[5] -= coeff[5] * bet3 * s * isqrtd;
    f[5] += coeff[6] * as_prime * bet2_star;

    f[6] += coeff[0] * as_prime * bet3_star;
    f[6] += coeff[1] * bet2 * s * isqrtd;
    f[6] -= coeff[2] * af_prime * bet3_star;
    f[6] -= coeff[4] * af_prime * bet3_star;
    f[6] += coeff[5] * bet2 * s * isqrtd;
    f[6] += coeff[6] * as_prime * bet3_star;
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) f[n] = fl[n];
  }
  if (ev[6] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) f[n] = fr[n];
  }
  if (llf_flag != 0) {
    Real cfl = fast_speed(eos, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi);
    Real cfr = fast_speed(eos, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi);
    Real a = 0.5 * max_std((fabs(wl[IVX]) + cfl), (fabs(wr[IVX]) + cfr));
#pragma unroll
    for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
}

//----------------------------------------------------------------------------------------
// Isothermal EOS.  HLLE for hydro: hydro/rsolvers/hydro/hlle.cpp:56-150 with NON_BAROTROPIC_EOS == 0
// (Roe-averaged velocity, a = iso_cs); the energy flux is zero.
ATHK_HD void hlle_iso_hydro_flux(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  const Real iso_cs = eos.iso_cs;
  Real fl[4], fr[4];
  Real sqrtdl = sqrt(wl[IDN]);
  Real sqrtdr = sqrt(wr[IDN]);
  Real isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  Real wroe_vx = (sqrtdl * wl[IVX] + sqrtdr * wr[IVX]) * isdlpdr;
  Real cl = iso_cs, cr = iso_cs;   // EquationOfState::SoundSpeed (isothermal_hydro.cpp)
  Real a = iso_cs;
  Real al = min_std((wroe_vx - a), (wl[IVX] - cl));
  Real ar = max_std((wroe_vx + a), (wr[IVX] + cr));
  Real bp = ar > 0.0 ? ar : 0.0;
  Real bm = al < 0.0 ? al : 0.0;
  Real vxl = wl[IVX] - bm;
  Real vxr = wr[IVX] - bp;
  fl[IDN] = wl[IDN] * vxl;
  fr[IDN] = wr[IDN] * vxr;
  fl[IVX] = wl[IDN] * wl[IVX] * vxl;
  fr[IVX] = wr[IDN] * wr[IVX] * vxr;
  fl[IVY] = wl[IDN] * wl[IVY] * vxl;
  fr[IVY] = wr[IDN] * wr[IVY] * vxr;
  fl[IVZ] = wl[IDN] * wl[IVZ] * vxl;
  fr[IVZ] = wr[IDN] * wr[IVZ] * vxr;
  fl[IVX] += (iso_cs * iso_cs) * wl[IDN];
  fr[IVX] += (iso_cs * iso_cs) * wr[IDN];
  Real tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
#pragma unroll
  for (int n = 0; n < 4; ++n) f[n] = 0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp;
  f[IEN] = 0.0;
}

// HLLD for isothermal MHD (Mignone 2007): hydro/rsolvers/mhd/hlld_iso.cpp:52-245.
// wl/wr = (d, vx, vy, vz, -, by, bz); f = (mass, mx, my, mz, 0, by-flux, bz-flux)
ATHK_HD void hlld_iso_flux(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi, Real f[7]) {
  const Real SMALL_NUMBER = 1.0e-4;
  const Real dfloor = eos.dfloor, cs = eos.iso_cs;
  Real spd[5];
  struct C6 { Real d, mx, my, mz, by, bz; } ul, ur, ulst, urst, ucst, fl, fr;
  ul.d = wl[IDN]; ul.mx = wl[IVX] * ul.d; ul.my = wl[IVY] * ul.d; ul.mz = wl[IVZ] * ul.d; ul.by = wl[IBY]; ul.bz = wl[IBZ];
  ur.d = wr[IDN]; ur.mx = wr[IVX] * ur.d; ur.my = wr[IVY] * ur.d; ur.mz = wr[IVZ] * ur.d; ur.by = wr[IBY]; ur.bz = wr[IBZ];
  Real cfl = fast_speed_iso(eos, wl[IDN], wl[IBY], wl[IBZ], bxi);
  Real cfr = fast_speed_iso(eos, wr[IDN], wr[IBY], wr[IBZ], bxi);
  spd[0] = min_std(wl[IVX] - cfl, wr[IVX] - cfr);
  spd[4] = max_std(wl[IVX] + cfl, wr[IVX] + cfr);
  Real bxsq = bxi * bxi;
  Real ptl = sqr(cs) * wl[IDN] + 0.5 * (bxsq + sqr(wl[IBY]) + sqr(wl[IBZ]));
  Real ptr = sqr(cs) * wr[IDN] + 0.5 * (bxsq + sqr(wr[IBY]) + sqr(wr[IBZ]));
  fl.d = ul.mx;
  fl.mx = ul.mx * wl[IVX] + ptl - bxsq;
  fl.my = ul.my * wl[IVX] - bxi * ul.by;
  fl.mz = ul.mz * wl[IVX] - bxi * ul.bz;
  fl.by = ul.by * wl[IVX] - bxi * wl[IVY];
  fl.bz = ul.bz * wl[IVX] - bxi * wl[IVZ];
  fr.d = ur.mx;
  fr.mx = ur.mx * wr[IVX] + ptr - bxsq;
  fr.my = ur.my * wr[IVX] - bxi * ur.by;
  fr.mz = ur.mz * wr[IVX] - bxi * ur.bz;
  fr.by = ur.by * wr[IVX] - bxi * wr[IVY];
  fr.bz = ur.bz * wr[IVX] - bxi * wr[IVZ];
  Real idspd = 1.0 / (spd[4] - spd[0]);
  Real dhll = (spd[4] * ur.d - spd[0] * ul.d - fr.d + fl.d) * idspd;
  dhll = max_std(dhll, dfloor);
  Real sqrtdhll = sqrt(dhll);
  Real fdhll = (spd[4] * fl.d - spd[0] * fr.d + spd[4] * spd[0] * (ur.d - ul.d)) * idspd;
  Real fmxhll = (spd[4] * fl.mx - spd[0] * fr.mx + spd[4] * spd[0] * (ur.mx - ul.mx)) * idspd;
  Real ustar = fdhll / dhll;
  Real mxhll = (spd[4] * ur.mx - spd[0] * ul.mx - fr.mx + fl.mx) * idspd;
  spd[1] = ustar - fabs(bxi) / sqrtdhll;
  spd[3] = ustar + fabs(bxi) / sqrtdhll;
  ulst.d = dhll;
  ulst.mx = mxhll;
  Real tmp = (spd[0] - spd[1]) * (spd[0] - spd[3]);
  if (fabs(spd[0] - spd[1]) < (SMALL_NUMBER)*cs) {
    ulst.my = ul.my; ulst.mz = ul.mz; ulst.by = ul.by; ulst.bz = ul.bz;
  } else {
    Real mfact = bxi * (ustar - wl[IVX]) / tmp;
    Real bfact = (ul.d * sqr(spd[0] - wl[IVX]) - bxsq) / (dhll * tmp);
    ulst.my = dhll * wl[IVY] - ul.by * mfact;
    ulst.mz = dhll * wl[IVZ] - ul.bz * mfact;
    ulst.by = ul.by * bfact;
    ulst.bz = ul.bz * bfact;
  }
  urst.d = dhll;
  urst.mx = mxhll;
  tmp = (spd[4] - spd[1]) * (spd[4] - spd[3]);
  if (fabs(spd[4] - spd[3]) < (SMALL_NUMBER)*cs) {
    urst.my = ur.my; urst.mz = ur.mz; urst.by = ur.by; urst.bz = ur.bz;
  } else {
    Real mfact = bxi * (ustar - wr[IVX]) / tmp;
    Real bfact = (ur.d * sqr(spd[4] - wr[IVX]) - bxsq) / (dhll * tmp);
    urst.my = dhll * wr[IVY] - ur.by * mfact;
    urst.mz = dhll * wr[IVZ] - ur.bz * mfact;
    urst.by = ur.by * bfact;
    urst.bz = ur.bz * bfact;
  }
  Real x = sqrtdhll * (bxi > 0.0 ? 1.0 : -1.0);
  ucst.d = dhll;
  ucst.mx = mxhll;
  ucst.my = 0.5 * (ulst.my + urst.my + (urst.by - ulst.by) * x);
  ucst.mz = 0.5 * (ulst.mz + urst.mz + (urst.bz - ulst.bz) * x);
  ucst.by = 0.5 * (ulst.by + urst.by + (urst.my - ulst.my) / x);
  ucst.bz = 0.5 * (ulst.bz + urst.bz + (urst.mz - ulst.mz) / x);
  if (spd[0] >= 0.0) {
    f[IDN] = fl.d; f[IVX] = fl.mx; f[IVY] = fl.my; f[IVZ] = fl.mz; f[IBY] = fl.by; f[IBZ] = fl.bz;
  } else if (spd[4] <= 0.0) {
    f[IDN] = fr.d; f[IVX] = fr.mx; f[IVY] = fr.my; f[IVZ] = fr.mz; f[IBY] = fr.by; f[IBZ] = fr.bz;
  } else if (spd[1] >= 0.0) {
    f[IDN] = fl.d + spd[0] * (ulst.d - ul.d);
    f[IVX] = fl.mx + spd[0] * (ulst.mx - ul.mx);
    f[IVY] = fl.my + spd[0] * (ulst.my - ul.my);
    f[IVZ] = fl.mz + spd[0] * (ulst.mz - ul.mz);
    f[IBY] = fl.by + spd[0] * (ulst.by - ul.by);
    f[IBZ] = fl.bz + spd[0] * (ulst.bz - ul.bz);
  } else if (spd[3] <= 0.0) {
    f[IDN] = fr.d + spd[4] * (urst.d - ur.d);
    f[IVX] = fr.mx + spd[4] * (urst.mx - ur.mx);
    f[IVY] = fr.my + spd[4] * (urst.my - ur.my);
    f[IVZ] = fr.mz + spd[4] * (urst.mz - ur.mz);
    f[IBY] = fr.by + spd[4] * (urst.by - ur.by);
    f[IBZ] = fr.bz + spd[4] * (urst.bz - ur.bz);
  } else {
    f[IDN] = dhll * ustar;
    f[IVX] = fmxhll;
    f[IVY] = ucst.my * ustar - bxi * ucst.by;
    f[IVZ] = ucst.mz * ustar - bxi * ucst.bz;
    f[IBY] = ucst.by * ustar - bxi * ucst.my / ucst.d;
    f[IBZ] = ucst.bz * ustar - bxi * ucst.mz / ucst.d;
  }
  f[IEN] = 0.0;
}

//----------------------------------------------------------------------------------------
// compile-time selection: every solver is its own kernel instantiation, so the registers of the
// default HLLC / HLLD kernels do not pay for the Roe eigen-decomposition.
constexpr int default_rsolver(bool mhd) { return mhd ? RS_HLLD : RS_HLLC; }

template <int RS>
ATHK_HD void hydro_riemann(const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  if (RS == RS_HLLE) hlle_hydro_flux(eos, wl, wr, f);
  else if (RS == RS_SR) sr_hydro_flux(eos, wl, wr, f);
  else if (RS == RS_HLLE_ISO) hlle_iso_hydro_flux(eos, wl, wr, f);
  else if (RS == RS_ROE) roe_hydro_flux(eos, wl, wr, f);
  else if (RS == RS_LLF) llf_hydro_flux(eos, wl, wr, f);
  else hllc_flux(eos, wl, wr, f);
}

template <int RS>
ATHK_HD void mhd_riemann(const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                         Real f[7]) {
  if (RS == RS_HLLE) hlle_mhd_flux(eos, wl, wr, bxi, f);
  else if (RS == RS_HLLD_ISO) hlld_iso_flux(eos, wl, wr, bxi, f);
  else if (RS == RS_ROE) roe_mhd_flux(eos, wl, wr, bxi, f);
  else if (RS == RS_LLF) llf_mhd_flux(eos, wl, wr, bxi, f);
  else hlld_flux(eos, wl, wr, bxi, f);
}

// Math-policy forms (athk_math.h): the default solvers HLLC / HLLD issue their divisions and square
// roots through M; the alternative solvers keep the compiler's operations.
template <int RS, class M>
ATHK_HD void hydro_riemann_m(M &m, const EosParams &eos, const Real wl[5], const Real wr[5], Real f[5]) {
  if (RS == RS_HLLC) hllc_flux_m(m, eos, wl, wr, f);
  else hydro_riemann<RS>(eos, wl, wr, f);
}

template <int RS, class M>
ATHK_HD void mhd_riemann_m(M &m, const EosParams &eos, const Real wl[7], const Real wr[7], const Real bxi,
                           Real f[7]) {
  if (RS == RS_HLLD) hlld_flux_m(m, eos, wl, wr, bxi, f);
  else mhd_riemann<RS>(eos, wl, wr, bxi, f);
}

// Math policy of the fused sweeps (athk_math.h): HLLC / HLLD faces run the straight-line Newton
// sequences (MathSL) and fall back to the compiler's operations for a face whose operands left
// their domain; body(m) must be safe to run twice (it only computes and stores).
#ifndef ATHK_SL_MATH
#define ATHK_SL_MATH 1
#endif
// AR = 1: the fast-arithmetic kernel set (MathFast, no fallback); only instantiated by athk_fast.cu
template <int RS, int AR = 0, class F>
ATHK_HD void with_math(F &&body) {
#if defined(__CUDA_ARCH__)
  if (AR == 1) {
    MathFast m;
    body(m);
    return;
  }
#endif
#if ATHK_SL_MATH && defined(__CUDA_ARCH__)
  if (RS == RS_HLLC || RS == RS_HLLD || RS == RS_SR) {   // (RS_SR: reconstruction and flux divergence; its solvers keep the compiler's operations)
    MathSL m;
    body(m);
    if (!m.ok()) {
      MathIEEE mi;
      body(mi);
    }
    return;
  }
#endif
  MathIEEE mi;
  body(mi);
}

// Riemann solve at one face + the per-face outputs (EMFs, CT weight, optional flux store)
}  // namespace athk
#endif  // ATHK_RSOLVERS_H_
