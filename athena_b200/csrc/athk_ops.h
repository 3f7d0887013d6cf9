// athk_ops.h -- per-point operator bodies of the MeshBlock update ("one thread = one zone / face /
// edge").  Each functor is a POD of raw pointers + sizes with
//     void operator()(int blk, int k, int j, int i) const
// so that the CUDA launcher (athk_api.cu: one thread per point, i fastest => coalesced rows) and
// the index-logic harness under tests/emul (plain loops, test infrastructure only) execute the
// SAME body.  Arithmetic lives in athk_physics.h; this file is extents, stencils and layout.
//
// Device layout ("padded box"): every 3-D array of every kind (cell, face, edge) of a MeshBlock
// uses the same strides  idx(k,j,i) = k*sk + j*sj + i  with sj = P1 >= nc1+1 (rounded up to a
// multiple of 4 doubles = one 32-byte sector), sk = P1*(nc2+1), plane size sn = sk*(nc3+1).
// A multi-component array of block b is [b][n][sn].  One index computation serves all arrays.
#ifndef ATHK_OPS_H_
#define ATHK_OPS_H_

#include "athk_rsolvers.h"

namespace athk {

typedef long long idx_t;

struct Grid {
  int nx1, nx2, nx3, ng;
  int nc1, nc2, nc3;            // cells including ghosts
  int is, ie, js, je, ks, ke;   // active range (AthenaArray indexing: is = ng)
  int P1, P2, P3;               // padded extents
  int nb;                       // MeshBlocks in this context
  idx_t sj, sk, sn;             // strides (elements)
};

ATHK_HD idx_t gidx(const Grid &g, int k, int j, int i) { return (idx_t)k * g.sk + (idx_t)j * g.sj + i; }

// per-block 1-D geometry, each [nb][PG] with PG = max(P1,P2,P3)
struct Geom {
  int PG;
  const Real *dxf[3];   // dx1f(i), dx2f(j), dx3f(k)
  const Real *cp[3];    // (xf(n+1) - xv(n)) / dxf(n)   plm.cpp:152
  const Real *cm[3];    // (xv(n) - xf(n)) / dxf(n)     plm.cpp:153
};

// table entry towards a coarser or finer neighbour (== ATHK_NBR_REFINED, include/athk.h)
enum { NBR_REFINED = -1000000 };
// neighbour table entry for offsets (ox1,ox2,ox3) in {-1,0,1}
ATHK_HD int nbr_slot(int ox1, int ox2, int ox3) { return (ox3 + 1) * 9 + (ox2 + 1) * 3 + (ox1 + 1); }

//========================================================================================
// HydroSourceTerms::ConstantAcceleration (hydro/srcterms/constant_acc.cpp:25-76) for one cell:
// momentum and energy sources from the stage-input primitives, direction after direction as the
// reference's three loops add them
struct ConstAccPoint {
  Grid g;
  const Real *w;   // [nb][5][sn] primitives at the beginning of the stage
  Real *u;         // [nb][5][sn] conserved register the stage has just written
  Real dt;         // beta * dt_mesh
  Real g1, g2, g3;
  int iso;         // isothermal EOS: no energy source (constant_acc.cpp:38: if (NON_BAROTROPIC_EOS))
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real *wb = w + (idx_t)blk * 5 * g.sn + c;
    Real *ub = u + (idx_t)blk * 5 * g.sn + c;
    const Real rho = wb[0];
    Real e = ub[(idx_t)IEN * g.sn];
    if (g1 != 0.0) {
      const Real src = dt * rho * g1;
      ub[(idx_t)IM1 * g.sn] += src;
      e += src * wb[(idx_t)IVX * g.sn];
    }
    if (g2 != 0.0) {
      const Real src = dt * rho * g2;
      ub[(idx_t)IM2 * g.sn] += src;
      e += src * wb[(idx_t)IVY * g.sn];
    }
    if (g3 != 0.0) {
      const Real src = dt * rho * g3;
      ub[(idx_t)IM3 * g.sn] += src;
      e += src * wb[(idx_t)IVZ * g.sn];
    }
    if (!iso) ub[(idx_t)IEN * g.sn] = e;
  }
};

//========================================================================================
// EquationOfState::ConservedToPrimitive for one cell (+ CalculateCellCenteredField for MHD)
// (SR = true: special-relativistic hydro, its own instantiation so that the Newtonian kernels keep
//  their register budget)
template <bool MHD, bool SR = false>
struct Cons2PrimPoint {
  Grid g;
  EosParams eos;
  Real *u;          // [nb][5][sn]   in/out (floors written back)
  const Real *b;    // [nb][3][sn]   face fields (x1f,x2f,x3f)
  Real *w;          // [nb][5][sn]
  Real *bcc;        // [nb][3][sn]
  const int *nbr;   // [nb][27]
  // optional fusion of Hydro::NewBlockTimeStep (new_blockdt.cpp:72-124) into the last stage's
  // W(U): active cells also fold min(dx/(|v|+c)) into dt_out[blk] (null = off)
  Real *dt_out;
  Geom geo;
  // passive scalars (ns > 0): EquationOfState::PassiveScalarConservedToPrimitive
  // (eos/eos_scalars.cpp:31-60) over the same range, after the density floor
  Real *s;          // [nb][ns][sn]  in/out (floored to sfloor*rho)
  Real *r;          // [nb][ns][sn]
  int ns;
  ATHK_HD void fold_dt(int blk, Real dtc) const {
#if defined(__CUDA_ARCH__)
    // positive doubles order like their bit patterns; the racy pre-read only filters
    if (dtc < dt_out[blk])
      atomicMin((unsigned long long *)&dt_out[blk], (unsigned long long)__double_as_longlong(dtc));
#else
    if (dtc < dt_out[blk]) dt_out[blk] = dtc;
#endif
  }
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const Real dtc = run(blk, k, j, i);
    if (dt_out != nullptr && dtc < 1.0e300) fold_dt(blk, dtc);
  }
  // W(U) of one cell; returns the cell's time-step bound when the fold is on (else a huge number).
  // The launcher reduces it per CTA (k_points_min) or operator() folds it cell by cell.
  ATHK_HD Real run(int blk, int k, int j, int i) const {
    const Real none = 1.0e308;
    // range: interior plus ghost sides with a neighbour (time_integrator.cpp:2206-2212)
    const int *nb = nbr + blk * 27;
    if (i < g.is && nb[nbr_slot(-1, 0, 0)] == -1) return none;
    if (i > g.ie && nb[nbr_slot(1, 0, 0)] == -1) return none;
    if (j < g.js && nb[nbr_slot(0, -1, 0)] == -1) return none;
    if (j > g.je && nb[nbr_slot(0, 1, 0)] == -1) return none;
    if (k < g.ks && nb[nbr_slot(0, 0, -1)] == -1) return none;
    if (k > g.ke && nb[nbr_slot(0, 0, 1)] == -1) return none;
    const idx_t c = gidx(g, k, j, i);
    Real *ub = u + (idx_t)blk * 5 * g.sn + c;
    Real *wb = w + (idx_t)blk * 5 * g.sn + c;
    Real uc[5], wc[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) uc[n] = ub[n * g.sn];
    const Real d0 = uc[IDN], e0 = uc[IEN];
    if (MHD) {
      const Real *bb = b + (idx_t)blk * 3 * g.sn + c;
      Real *bc = bcc + (idx_t)blk * 3 * g.sn + c;
      // uniform Cartesian: unweighted average of the two faces (field.cpp:139-142)
      Real bc1 = 0.5 * bb[0] + 0.5 * bb[1];
      Real bc2 = 0.5 * bb[g.sn] + 0.5 * bb[g.sn + g.sj];
      Real bc3 = 0.5 * bb[2 * g.sn] + 0.5 * bb[2 * g.sn + g.sk];
      bc[0] = bc1;
      bc[g.sn] = bc2;
      bc[2 * g.sn] = bc3;
      if (eos.iso) cons2prim_iso(eos_of_block(eos, blk), uc, wc);
      else cons2prim_mhd(eos_of_block(eos, blk), uc, bc1, bc2, bc3, wc);
    } else {
      if (SR) {   // the output array still holds W of two calls ago: fallback state and first guess
#pragma unroll
        for (int n = 0; n < 5; ++n) wc[n] = wb[n * g.sn];
        if (cons2prim_sr(eos_of_block(eos, blk), uc, wc)) {
          ub[IM1 * g.sn] = uc[IM1];
          ub[IM2 * g.sn] = uc[IM2];
          ub[IM3 * g.sn] = uc[IM3];
          ub[IDN * g.sn] = uc[IDN];
          ub[IEN * g.sn] = uc[IEN];
        }
      } else if (eos.iso) cons2prim_iso(eos_of_block(eos, blk), uc, wc);
      else cons2prim_hydro(eos_of_block(eos, blk), uc, wc);
    }
    if (uc[IDN] != d0) ub[IDN * g.sn] = uc[IDN];
    if (uc[IEN] != e0) ub[IEN * g.sn] = uc[IEN];
#pragma unroll
    for (int n = 0; n < 5; ++n) wb[n * g.sn] = wc[n];
    for (int n = 0; n < ns; ++n) {
      const idx_t o = ((idx_t)blk * ns + n) * g.sn + c;
      const Real s0 = s[o];
      const Real sfl = eos_sfloor(eos, blk);
      const Real sn = (s0 < sfl * uc[IDN]) ? sfl * uc[IDN] : s0;
      if (sn != s0) s[o] = sn;
      r[o] = sn / uc[IDN];
    }
    Real dtcell = none;
    if (dt_out != nullptr && i >= g.is && i <= g.ie && j >= g.js && j <= g.je && k >= g.ks &&
        k <= g.ke) {
      // (three independent sqrt / sqrt / divide chains: issued straight-line they overlap, see
      //  athk_math.h; min() folding is idempotent, so a flagged cell is simply redone)
      with_math<RS_HLLD>([&](auto &m) {
      Real dt1 = geo.dxf[0][blk * geo.PG + i], dt2 = geo.dxf[1][blk * geo.PG + j],
           dt3 = geo.dxf[2][blk * geo.PG + k];
      if (MHD) {
        const Real *bb = b + (idx_t)blk * 3 * g.sn + c;
        const Real b1 = 0.5 * bb[0] + 0.5 * bb[1], b2 = 0.5 * bb[g.sn] + 0.5 * bb[g.sn + g.sj],
                   b3 = 0.5 * bb[2 * g.sn] + 0.5 * bb[2 * g.sn + g.sk];
        Real bx = b1 + fabs(bb[0] - b1);
        if (eos.iso) {   // (isothermal EOS: compiler operations, eos/isothermal_mhd.cpp:121)
          dt1 = dt1 / max_std(0.0, (fabs(wc[IVX]) + fast_speed_iso(eos, wc[IDN], b2, b3, bx)));
          bx = b2 + fabs(bb[g.sn] - b2);
          dt2 = dt2 / max_std(0.0, (fabs(wc[IVY]) + fast_speed_iso(eos, wc[IDN], b3, b1, bx)));
          bx = b3 + fabs(bb[2 * g.sn] - b3);
          dt3 = dt3 / max_std(0.0, (fabs(wc[IVZ]) + fast_speed_iso(eos, wc[IDN], b1, b2, bx)));
        } else {
        dt1 = m.div(dt1, max_std(0.0, (fabs(wc[IVX]) + fast_speed_m(m, eos, wc[IDN], wc[IPR], b2, b3, bx))));
        bx = b2 + fabs(bb[g.sn] - b2);
        dt2 = m.div(dt2, max_std(0.0, (fabs(wc[IVY]) + fast_speed_m(m, eos, wc[IDN], wc[IPR], b3, b1, bx))));
        bx = b3 + fabs(bb[2 * g.sn] - b3);
        dt3 = m.div(dt3, max_std(0.0, (fabs(wc[IVZ]) + fast_speed_m(m, eos, wc[IDN], wc[IPR], b1, b2, bx))));
        }
      } else {
        if (!SR) {   // (SR: the cell widths themselves, new_blockdt.cpp:130)
        const Real cs = eos.iso ? eos.iso_cs : m.sqrt(m.div(eos.gamma * wc[IPR], wc[IDN]));
        dt1 = m.div(dt1, max_std(0.0, (fabs(wc[IVX]) + cs)));
        dt2 = m.div(dt2, max_std(0.0, (fabs(wc[IVY]) + cs)));
        dt3 = m.div(dt3, max_std(0.0, (fabs(wc[IVZ]) + cs)));
        }
      }
      dtcell = min_std(min_std(dt1, dt2), dt3);
      });
    }
    return dtcell;
  }
};

//========================================================================================
// Field::ComputeCornerE, 3-D branch (calculate_corner_e.cpp:130-229): cell-centred -(v x B)
// is evaluated on the fly (same bits as the reference's cc_e_ scratch array).
struct CornerEPoint {
  Grid g;
  const Real *w, *bcc;
  const Real *ef;    // [nb][6][sn]: 0 e3_x1f, 1 e2_x1f, 2 e1_x2f, 3 e3_x2f, 4 e2_x3f, 5 e1_x3f
  const Real *wct;   // [nb][3][sn]
  Real *e;           // [nb][3][sn]: x1e,x2e,x3e
  ATHK_HD Real cce(const Real *wb, const Real *bc, idx_t c, int comp) const {
    // comp 0: vz*by - vy*bz ; 1: vx*bz - vz*bx ; 2: vy*bx - vx*by
    const Real v1 = wb[IVX * g.sn + c], v2 = wb[IVY * g.sn + c], v3 = wb[IVZ * g.sn + c];
    const Real b1 = bc[c], b2 = bc[g.sn + c], b3 = bc[2 * g.sn + c];
    if (comp == 0) return v3 * b2 - v2 * b3;
    if (comp == 1) return v1 * b3 - v3 * b1;
    return v2 * b1 - v1 * b2;
  }
  // k-marching form (launch_points_kcol): the edges (k0..k1, j, i).  Same expressions as
  // operator() below; what plane k shares with plane k-1 (four cell-centred EMFs, two CT weights, two
  // face EMFs) is carried in registers, and every load of a plane is issued before its three
  // stores (the arrays are not declared restrict, so a store fences the loads behind it).
  struct CellVB { Real v1, v2, v3, b1, b2, b3; };
  ATHK_HD CellVB ldcell(const Real *wb, const Real *bc, idx_t c) const {
    CellVB q;
    q.v1 = wb[IVX * g.sn + c]; q.v2 = wb[IVY * g.sn + c]; q.v3 = wb[IVZ * g.sn + c];
    q.b1 = bc[c]; q.b2 = bc[g.sn + c]; q.b3 = bc[2 * g.sn + c];
    return q;
  }
  ATHK_HD static Real cc0(const CellVB &q) { return q.v3 * q.b2 - q.v2 * q.b3; }
  ATHK_HD static Real cc1(const CellVB &q) { return q.v1 * q.b3 - q.v3 * q.b1; }
  ATHK_HD static Real cc2(const CellVB &q) { return q.v2 * q.b1 - q.v1 * q.b2; }
  ATHK_HD void col(int blk, int k0, int k1, int j, int i) const {
    const idx_t si = 1, sj = g.sj, sk = g.sk;
    const Real *wb = w + (idx_t)blk * 5 * g.sn;
    const Real *bc = bcc + (idx_t)blk * 3 * g.sn;
    const Real *efb = ef + (idx_t)blk * 6 * g.sn;
    const Real *e3_x1f = efb, *e2_x1f = efb + g.sn, *e1_x2f = efb + 2 * g.sn,
               *e3_x2f = efb + 3 * g.sn, *e2_x3f = efb + 4 * g.sn, *e1_x3f = efb + 5 * g.sn;
    const Real *w_x1f = wct + (idx_t)blk * 3 * g.sn, *w_x2f = w_x1f + g.sn,
               *w_x3f = w_x1f + 2 * g.sn;
    Real *eb = e + (idx_t)blk * 3 * g.sn;
    idx_t c = gidx(g, k0, j, i);
    Real p_cc0, p_cc0j, p_cc1, p_cc1i, p_wx2, p_wx1, p_e1x2, p_e2x1;   // plane k-1
    {
      const CellVB a = ldcell(wb, bc, c - sk), aj = ldcell(wb, bc, c - sk - sj), ai = ldcell(wb, bc, c - sk - si);
      p_cc0 = cc0(a); p_cc0j = cc0(aj); p_cc1 = cc1(a); p_cc1i = cc1(ai);
      p_wx2 = w_x2f[c - sk]; p_wx1 = w_x1f[c - sk]; p_e1x2 = e1_x2f[c - sk]; p_e2x1 = e2_x1f[c - sk];
    }
    for (int k = k0; k <= k1; ++k, c += sk) {
      const CellVB q = ldcell(wb, bc, c), qj = ldcell(wb, bc, c - sj), qi = ldcell(wb, bc, c - si);
      const Real v1ji = wb[IVX * g.sn + c - sj - si], v2ji = wb[IVY * g.sn + c - sj - si];
      const Real b1ji = bc[c - sj - si], b2ji = bc[g.sn + c - sj - si];
      const Real wx1 = w_x1f[c], wx1j = w_x1f[c - sj], wx2 = w_x2f[c], wx2i = w_x2f[c - si];
      const Real wx3 = w_x3f[c], wx3j = w_x3f[c - sj], wx3i = w_x3f[c - si];
      const Real e1x2 = e1_x2f[c], e1x3 = e1_x3f[c], e1x3j = e1_x3f[c - sj];
      const Real e2x1 = e2_x1f[c], e2x3 = e2_x3f[c], e2x3i = e2_x3f[c - si];
      const Real e3x1 = e3_x1f[c], e3x1j = e3_x1f[c - sj], e3x2 = e3_x2f[c], e3x2i = e3_x2f[c - si];
      const Real c0 = cc0(q), c0j = cc0(qj), c1 = cc1(q), c1i = cc1(qi);
      Real r1, r2, r3;
      {
        const Real cc00 = c0, cc0m = c0j, ccm0 = p_cc0, ccmm = p_cc0j;
        Real de1_l3 = (1.0 - p_wx2) * (e1x3 - ccm0) + (p_wx2) * (e1x3j - ccmm);
        Real de1_r3 = (1.0 - wx2) * (e1x3 - cc00) + (wx2) * (e1x3j - cc0m);
        Real de1_l2 = (1.0 - wx3j) * (e1x2 - cc0m) + (wx3j) * (p_e1x2 - ccmm);
        Real de1_r2 = (1.0 - wx3) * (e1x2 - cc00) + (wx3) * (p_e1x2 - ccm0);
        r1 = 0.25 * (de1_l3 + de1_r3 + de1_l2 + de1_r2 + p_e1x2 + e1x2 + e1x3j + e1x3);
      }
      {
        const Real cc00 = c1, cc0m = c1i, ccm0 = p_cc1, ccmm = p_cc1i;
        Real de2_l3 = (1.0 - p_wx1) * (e2x3 - ccm0) + (p_wx1) * (e2x3i - ccmm);
        Real de2_r3 = (1.0 - wx1) * (e2x3 - cc00) + (wx1) * (e2x3i - cc0m);
        Real de2_l1 = (1.0 - wx3i) * (e2x1 - cc0m) + (wx3i) * (p_e2x1 - ccmm);
        Real de2_r1 = (1.0 - wx3) * (e2x1 - cc00) + (wx3) * (p_e2x1 - ccm0);
        r2 = 0.25 * (de2_l3 + de2_r3 + de2_l1 + de2_r1 + e2x3i + e2x3 + p_e2x1 + e2x1);
      }
      {
        const Real cc00 = cc2(q), cc0m = cc2(qi), ccm0 = cc2(qj), ccmm = v2ji * b1ji - v1ji * b2ji;
        Real de3_l2 = (1.0 - wx1j) * (e3x2 - ccm0) + (wx1j) * (e3x2i - ccmm);
        Real de3_r2 = (1.0 - wx1) * (e3x2 - cc00) + (wx1) * (e3x2i - cc0m);
        Real de3_l1 = (1.0 - wx2i) * (e3x1 - cc0m) + (wx2i) * (e3x1j - ccmm);
        Real de3_r1 = (1.0 - wx2) * (e3x1 - cc00) + (wx2) * (e3x1j - ccm0);
        r3 = 0.25 * (de3_l1 + de3_r1 + de3_l2 + de3_r2 + e3x2i + e3x2 + e3x1j + e3x1);
      }
      eb[c] = r1;
      eb[g.sn + c] = r2;
      eb[2 * g.sn + c] = r3;
      p_cc0 = c0; p_cc0j = c0j; p_cc1 = c1; p_cc1i = c1i;
      p_wx2 = wx2; p_wx1 = wx1; p_e1x2 = e1x2; p_e2x1 = e2x1;
    }
  }
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const idx_t si = 1, sj = g.sj, sk = g.sk;
    const Real *wb = w + (idx_t)blk * 5 * g.sn;
    const Real *bc = bcc + (idx_t)blk * 3 * g.sn;
    const Real *efb = ef + (idx_t)blk * 6 * g.sn;
    const Real *e3_x1f = efb, *e2_x1f = efb + g.sn, *e1_x2f = efb + 2 * g.sn,
               *e3_x2f = efb + 3 * g.sn, *e2_x3f = efb + 4 * g.sn, *e1_x3f = efb + 5 * g.sn;
    const Real *w_x1f = wct + (idx_t)blk * 3 * g.sn, *w_x2f = w_x1f + g.sn,
               *w_x3f = w_x1f + 2 * g.sn;
    Real *eb = e + (idx_t)blk * 3 * g.sn;
    {
      const Real cc00 = cce(wb, bc, c, 0), cc0m = cce(wb, bc, c - sj, 0),
                 ccm0 = cce(wb, bc, c - sk, 0), ccmm = cce(wb, bc, c - sk - sj, 0);
      Real de1_l3 = (1.0 - w_x2f[c - sk]) * (e1_x3f[c] - ccm0) +
                    (w_x2f[c - sk]) * (e1_x3f[c - sj] - ccmm);
      Real de1_r3 = (1.0 - w_x2f[c]) * (e1_x3f[c] - cc00) + (w_x2f[c]) * (e1_x3f[c - sj] - cc0m);
      Real de1_l2 = (1.0 - w_x3f[c - sj]) * (e1_x2f[c] - cc0m) +
                    (w_x3f[c - sj]) * (e1_x2f[c - sk] - ccmm);
      Real de1_r2 = (1.0 - w_x3f[c]) * (e1_x2f[c] - cc00) + (w_x3f[c]) * (e1_x2f[c - sk] - ccm0);
      eb[c] = 0.25 * (de1_l3 + de1_r3 + de1_l2 + de1_r2 + e1_x2f[c - sk] + e1_x2f[c] +
                      e1_x3f[c - sj] + e1_x3f[c]);
    }
    {
      const Real cc00 = cce(wb, bc, c, 1), cc0m = cce(wb, bc, c - si, 1),
                 ccm0 = cce(wb, bc, c - sk, 1), ccmm = cce(wb, bc, c - sk - si, 1);
      Real de2_l3 = (1.0 - w_x1f[c - sk]) * (e2_x3f[c] - ccm0) +
                    (w_x1f[c - sk]) * (e2_x3f[c - si] - ccmm);
      Real de2_r3 = (1.0 - w_x1f[c]) * (e2_x3f[c] - cc00) + (w_x1f[c]) * (e2_x3f[c - si] - cc0m);
      Real de2_l1 = (1.0 - w_x3f[c - si]) * (e2_x1f[c] - cc0m) +
                    (w_x3f[c - si]) * (e2_x1f[c - sk] - ccmm);
      Real de2_r1 = (1.0 - w_x3f[c]) * (e2_x1f[c] - cc00) + (w_x3f[c]) * (e2_x1f[c - sk] - ccm0);
      eb[g.sn + c] = 0.25 * (de2_l3 + de2_r3 + de2_l1 + de2_r1 + e2_x3f[c - si] + e2_x3f[c] +
                             e2_x1f[c - sk] + e2_x1f[c]);
    }
    {
      const Real cc00 = cce(wb, bc, c, 2), cc0m = cce(wb, bc, c - si, 2),
                 ccm0 = cce(wb, bc, c - sj, 2), ccmm = cce(wb, bc, c - sj - si, 2);
      Real de3_l2 = (1.0 - w_x1f[c - sj]) * (e3_x2f[c] - ccm0) +
                    (w_x1f[c - sj]) * (e3_x2f[c - si] - ccmm);
      Real de3_r2 = (1.0 - w_x1f[c]) * (e3_x2f[c] - cc00) + (w_x1f[c]) * (e3_x2f[c - si] - cc0m);
      Real de3_l1 = (1.0 - w_x2f[c - si]) * (e3_x1f[c] - cc0m) +
                    (w_x2f[c - si]) * (e3_x1f[c - sj] - ccmm);
      Real de3_r1 = (1.0 - w_x2f[c]) * (e3_x1f[c] - cc00) + (w_x2f[c]) * (e3_x1f[c - sj] - ccm0);
      eb[2 * g.sn + c] = 0.25 * (de3_l1 + de3_r1 + de3_l2 + de3_r2 + e3_x2f[c - si] + e3_x2f[c] +
                                 e3_x1f[c - sj] + e3_x1f[c]);
    }
  }
};

//========================================================================================
// EMF correction, same-level neighbours (flux_correction_fc.cpp:51-300, 689-940, 1355-1541).
// Per-block surface buffer: the six boundary planes of (e1,e2,e3).
struct EmfSurf {
  int nx1, nx2, nx3;
  // section sizes
  ATHK_HD idx_t n_x1_e2() const { return (idx_t)(nx3 + 1) * nx2; }
  ATHK_HD idx_t n_x1_e3() const { return (idx_t)nx3 * (nx2 + 1); }
  ATHK_HD idx_t n_x2_e1() const { return (idx_t)(nx3 + 1) * nx1; }
  ATHK_HD idx_t n_x2_e3() const { return (idx_t)nx3 * (nx1 + 1); }
  ATHK_HD idx_t n_x3_e1() const { return (idx_t)(nx2 + 1) * nx1; }
  ATHK_HD idx_t n_x3_e2() const { return (idx_t)nx2 * (nx1 + 1); }
  ATHK_HD idx_t n_x1() const { return n_x1_e2() + n_x1_e3(); }
  ATHK_HD idx_t n_x2() const { return n_x2_e1() + n_x2_e3(); }
  ATHK_HD idx_t n_x3() const { return n_x3_e1() + n_x3_e2(); }
  ATHK_HD idx_t per_block() const { return 2 * (n_x1() + n_x2() + n_x3()); }
  // offsets of a face section: dir 0..2, side 0 (inner) / 1 (outer)
  ATHK_HD idx_t face_off(int dir, int side) const {
    idx_t o = 0;
    if (dir >= 1) o += 2 * n_x1();
    if (dir >= 2) o += 2 * n_x2();
    const idx_t n = (dir == 0) ? n_x1() : (dir == 1 ? n_x2() : n_x3());
    return o + side * n;
  }
  // element offsets inside a face section (kk,jj,ii are zero-based from ks,js,is)
  ATHK_HD idx_t x1_e2(int kk, int jj) const { return (idx_t)kk * nx2 + jj; }
  ATHK_HD idx_t x1_e3(int kk, int jj) const { return n_x1_e2() + (idx_t)kk * (nx2 + 1) + jj; }
  ATHK_HD idx_t x2_e1(int kk, int ii) const { return (idx_t)kk * nx1 + ii; }
  ATHK_HD idx_t x2_e3(int kk, int ii) const { return n_x2_e1() + (idx_t)kk * (nx1 + 1) + ii; }
  ATHK_HD idx_t x3_e1(int jj, int ii) const { return (idx_t)jj * nx1 + ii; }
  ATHK_HD idx_t x3_e2(int jj, int ii) const { return n_x3_e1() + (idx_t)jj * (nx1 + 1) + ii; }
};

// LoadFluxBoundaryBufferSameLevel for all six faces of a block at once
struct EmfPackPoint {
  Grid g;
  EmfSurf sf;
  const Real *e;   // [nb][3][sn]
  Real *buf;       // [nb][per_block]
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real *eb = e + (idx_t)blk * 3 * g.sn;
    Real *bb = buf + (idx_t)blk * sf.per_block();
    const int kk = k - g.ks, jj = j - g.js, ii = i - g.is;
    const bool i_ok = (i <= g.ie), j_ok = (j <= g.je), k_ok = (k <= g.ke);
    if (i == g.is || i == g.ie + 1) {
      Real *s = bb + sf.face_off(0, i == g.is ? 0 : 1);
      if (j_ok) s[sf.x1_e2(kk, jj)] = eb[g.sn + c];
      if (k_ok) s[sf.x1_e3(kk, jj)] = eb[2 * g.sn + c];
    }
    if (j == g.js || j == g.je + 1) {
      Real *s = bb + sf.face_off(1, j == g.js ? 0 : 1);
      if (i_ok) s[sf.x2_e1(kk, ii)] = eb[c];
      if (k_ok) s[sf.x2_e3(kk, ii)] = eb[2 * g.sn + c];
    }
    if (k == g.ks || k == g.ke + 1) {
      Real *s = bb + sf.face_off(2, k == g.ks ? 0 : 1);
      if (i_ok) s[sf.x3_e1(jj, ii)] = eb[c];
      if (j_ok) s[sf.x3_e2(jj, ii)] = eb[g.sn + c];
    }
  }
};

// SetFluxBoundarySameLevel (+=) followed by AverageFluxBoundary, per boundary edge point.
// Accumulation order = neighbour enumeration order of bvals_base.cpp:326-372: faces (x1,x2,x3),
// then edges.  Remote neighbours (<= -2) read the received message of their link; it holds the
// one face section this kernel indexes for that neighbour (see emf_link_section).
struct EmfApplyPoint {
  Grid g;
  EmfSurf sf;
  Real *e;
  const Real *buf;       // local blocks' packed surfaces
  const int *nbr;        // [nb][27]
  const Real *rbuf;      // received EMF messages of remote links (may be null)
  const idx_t *rbase;    // [nlinks] message offset minus the face-section offset it stands for
  ATHK_HD const Real *surf_of(int id) const {
    if (id >= 0) return buf + (idx_t)id * sf.per_block();
    return rbuf + rbase[-id - 2];
  }
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const int a = (i == g.is) ? -1 : (i == g.ie + 1 ? 1 : 0);
    const int bq = (j == g.js) ? -1 : (j == g.je + 1 ? 1 : 0);
    const int cq = (k == g.ks) ? -1 : (k == g.ke + 1 ? 1 : 0);
    if (a == 0 && bq == 0 && cq == 0) return;
    const idx_t c = gidx(g, k, j, i);
    Real *eb = e + (idx_t)blk * 3 * g.sn;
    const int *nb = nbr + blk * 27;
    const int kk = k - g.ks, jj = j - g.js, ii = i - g.is;
    // ---- e1(k,j,i): i in [is,ie]; boundaries in x2 (bq) and x3 (cq)
    if (i <= g.ie && (bq != 0 || cq != 0)) {
      Real v = eb[c];
      int cnt = 1;
      if (bq != 0) {
        const int id = nb[nbr_slot(0, bq, 0)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(1, bq < 0 ? 1 : 0) + sf.x2_e1(kk, ii)]; ++cnt; }
      }
      if (cq != 0) {
        const int id = nb[nbr_slot(0, 0, cq)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(2, cq < 0 ? 1 : 0) + sf.x3_e1(jj, ii)]; ++cnt; }
      }
      if (bq != 0 && cq != 0) {
        const int id = nb[nbr_slot(0, bq, cq)];
        if (id != -1 && id != NBR_REFINED) {
          v += surf_of(id)[sf.face_off(1, bq < 0 ? 1 : 0) + sf.x2_e1(kk - cq * g.nx3, ii)];
          ++cnt;
        }
        if (cnt > 1) v *= (1.0 / (Real)cnt);
      } else if (cnt > 1) {
        v *= 0.5;
      }
      eb[c] = v;
    }
    // ---- e2(k,j,i): j in [js,je]; boundaries in x1 (a) and x3 (cq)
    if (j <= g.je && (a != 0 || cq != 0)) {
      Real v = eb[g.sn + c];
      int cnt = 1;
      if (a != 0) {
        const int id = nb[nbr_slot(a, 0, 0)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(0, a < 0 ? 1 : 0) + sf.x1_e2(kk, jj)]; ++cnt; }
      }
      if (cq != 0) {
        const int id = nb[nbr_slot(0, 0, cq)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(2, cq < 0 ? 1 : 0) + sf.x3_e2(jj, ii)]; ++cnt; }
      }
      if (a != 0 && cq != 0) {
        const int id = nb[nbr_slot(a, 0, cq)];
        if (id != -1 && id != NBR_REFINED) {
          v += surf_of(id)[sf.face_off(0, a < 0 ? 1 : 0) + sf.x1_e2(kk - cq * g.nx3, jj)];
          ++cnt;
        }
        if (cnt > 1) v *= (1.0 / (Real)cnt);
      } else if (cnt > 1) {
        v *= 0.5;
      }
      eb[g.sn + c] = v;
    }
    // ---- e3(k,j,i): k in [ks,ke]; boundaries in x1 (a) and x2 (bq)
    if (k <= g.ke && (a != 0 || bq != 0)) {
      Real v = eb[2 * g.sn + c];
      int cnt = 1;
      if (a != 0) {
        const int id = nb[nbr_slot(a, 0, 0)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(0, a < 0 ? 1 : 0) + sf.x1_e3(kk, jj)]; ++cnt; }
      }
      if (bq != 0) {
        const int id = nb[nbr_slot(0, bq, 0)];
        if (id != -1 && id != NBR_REFINED) { v += surf_of(id)[sf.face_off(1, bq < 0 ? 1 : 0) + sf.x2_e3(kk, ii)]; ++cnt; }
      }
      if (a != 0 && bq != 0) {
        const int id = nb[nbr_slot(a, bq, 0)];
        if (id != -1 && id != NBR_REFINED) {
          v += surf_of(id)[sf.face_off(0, a < 0 ? 1 : 0) + sf.x1_e3(kk, jj - bq * g.nx2)];
          ++cnt;
        }
        if (cnt > 1) v *= (1.0 / (Real)cnt);
      } else if (cnt > 1) {
        v *= 0.5;
      }
      eb[2 * g.sn + c] = v;
    }
  }
};

//========================================================================================
// MeshBlock::WeightedAve (mesh/weighted_ave.cpp:35-230 and :366-520), 2-input form used by
// vl2/rk2/rk3 (wght[3]=wght[4]=0).  The branch on zero weights is resolved on the host; MODE:
//   0: out += w1*in1 + w2*in2     1: out += w1*in1
//   2: out  = w1*in1 + w2*in2     3: out  = w1*in1
//   4: out  = w0*out + w1*in1 + w2*in2    5: out = w0*out + w1*in1    6: out *= w0
struct WeightedAvePoint {
  Grid g;
  int nvar;     // components handled per point
  int fc;       // 1: face-centred extents (x1f to ie+1, x2f to je+1, x3f to ke+1)
  int mode;
  Real w0, w1, w2;
  Real *out;
  const Real *in1, *in2;
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    for (int n = 0; n < nvar; ++n) {
      if (fc) {
        const bool ok = (n == 0) ? (j <= g.je && k <= g.ke)
                                 : (n == 1 ? (i <= g.ie && k <= g.ke) : (i <= g.ie && j <= g.je));
        if (!ok) continue;
      }
      const idx_t o = ((idx_t)blk * nvar + n) * g.sn + c;
      Real r;
      switch (mode) {
        case 0: r = out[o]; r += w1 * in1[o] + w2 * in2[o]; break;
        case 1: r = out[o]; r += w1 * in1[o]; break;
        case 2: r = w1 * in1[o] + w2 * in2[o]; break;
        case 3: r = w1 * in1[o]; break;
        case 4: r = w0 * out[o] + w1 * in1[o] + w2 * in2[o]; break;
        case 5: r = w0 * out[o] + w1 * in1[o]; break;
        default: r = out[o]; r *= w0; break;
      }
      out[o] = r;
    }
  }
};

//========================================================================================
// Hydro::AddFluxDivergence (add_flux_divergence.cpp:39-96), Cartesian areas/volume
// (coordinates.cpp:448-545): A1 = dx2f*dx3f, A2 = dx1f*dx3f, A3 = dx1f*dx2f, V = dx1f*dx2f*dx3f
struct FluxDivPoint {
  Grid g;
  Geom geo;
  int nvar;
  Real wght;
  const Real *f1, *f2, *f3;   // [nb][nvar][sn]
  Real *u;                    // [nb][nvar][sn]
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real dx1 = geo.dxf[0][blk * geo.PG + i], dx2 = geo.dxf[1][blk * geo.PG + j],
               dx3 = geo.dxf[2][blk * geo.PG + k];
    const Real a1 = dx2 * dx3, a2 = dx1 * dx3, a3 = dx1 * dx2;
    const Real vol = dx1 * dx2 * dx3;
    for (int n = 0; n < nvar; ++n) {
      const idx_t o = ((idx_t)blk * nvar + n) * g.sn + c;
      Real dflx = (a1 * f1[o + 1] - a1 * f1[o]);
      dflx += (a2 * f2[o + g.sj] - a2 * f2[o]);
      dflx += (a3 * f3[o + g.sk] - a3 * f3[o]);
      u[o] -= wght * dflx / vol;
    }
  }
};

//========================================================================================
// Field::CT (field/ct.cpp:31-116): launched over k in [ks,ke+1], j in [js,je+1], i in [is,ie+1]
struct CTPoint {
  Grid g;
  Geom geo;
  Real wght;
  const Real *e;   // [nb][3][sn]
  Real *b;         // [nb][3][sn]
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real *e1 = e + (idx_t)blk * 3 * g.sn, *e2 = e1 + g.sn, *e3 = e1 + 2 * g.sn;
    Real *bb = b + (idx_t)blk * 3 * g.sn;
    const Real dx1 = geo.dxf[0][blk * geo.PG + i], dx2 = geo.dxf[1][blk * geo.PG + j],
               dx3 = geo.dxf[2][blk * geo.PG + k];
    if (k <= g.ke && j <= g.je) {   // B1 at x1-face i
      const Real area = dx2 * dx3;
      Real v = bb[c];
      v -= (wght / area) * (dx3 * e3[c + g.sj] - dx3 * e3[c]);
      v += (wght / area) * (dx2 * e2[c + g.sk] - dx2 * e2[c]);
      bb[c] = v;
    }
    if (k <= g.ke && i <= g.ie) {   // B2 at x2-face j
      const Real area = dx1 * dx3;
      Real v = bb[g.sn + c];
      v += (wght / area) * (dx3 * e3[c + 1] - dx3 * e3[c]);
      v -= (wght / area) * (dx1 * e1[c + g.sk] - dx1 * e1[c]);
      bb[g.sn + c] = v;
    }
    if (j <= g.je && i <= g.ie) {   // B3 at x3-face k
      const Real area = dx1 * dx2;
      Real v = bb[2 * g.sn + c];
      v -= (wght / area) * (dx2 * e2[c + 1] - dx2 * e2[c]);
      v += (wght / area) * (dx1 * e1[c + g.sj] - dx1 * e1[c]);
      bb[2 * g.sn + c] = v;
    }
  }
};

//========================================================================================
// Slab geometry of one same-level boundary message (bvals_cc.cpp:287-302,386-420 for cell-
// centred data, bvals_fc.cpp:344-397,583-680 for face fields).  For neighbour offset o along a
// direction: dim = extent, ss = first index the sender packs, rs = first ghost index the
// receiver fills.  A face field has nx+1 entries along its own direction and the shared face
// is neither sent nor overwritten (bvals_fc.cpp:591-593).
struct Slab {
  int d1, d2, d3;      // extents (i,j,k)
  int s1, s2, s3;      // sender start
  int r1, r2, r3;      // receiver start
  ATHK_HD idx_t count() const { return (idx_t)d1 * d2 * d3; }
};

ATHK_HD void slab_axis(int o, int s, int e, int nx, int ng, bool own_dir, int *d, int *ss, int *rs) {
  if (o == 0) { *d = nx + (own_dir ? 1 : 0); *ss = s; *rs = s; }
  else if (o > 0) { *d = ng; *ss = e - ng + 1; *rs = e + (own_dir ? 2 : 1); }
  else { *d = ng; *ss = s + (own_dir ? 1 : 0); *rs = s - ng; }
}

// comp < 0: cell-centred; comp = 0,1,2: face field x1f,x2f,x3f
ATHK_HD Slab make_slab(const Grid &g, int comp, int o1, int o2, int o3) {
  Slab sl;
  slab_axis(o1, g.is, g.ie, g.nx1, g.ng, comp == 0, &sl.d1, &sl.s1, &sl.r1);
  slab_axis(o2, g.js, g.je, g.nx2, g.ng, comp == 1, &sl.d2, &sl.s2, &sl.r2);
  slab_axis(o3, g.ks, g.ke, g.nx3, g.ng, comp == 2, &sl.d3, &sl.s3, &sl.r3);
  return sl;
}

// one link = (local block, neighbour offset) whose neighbour lives on another GPU.  Messages of
// all links are concatenated in per-peer send / receive buffers; the element layout of one ghost
// message is [u: nvar_u][x1f][x2f][x3f][s: nvar_s], each section dense (k,j,i).
// slab geometry depends on the neighbour offset only: 27 offsets x {cc, x1f, x2f, x3f}, plus the
// section offsets of the ghost message; built once on the host (athk_create)
struct SlabTable {
  Slab slab[27][4];     // [nbr_slot][comp+1]
  idx_t sec[27][6];     // [nbr_slot][section]: 0 u, 1..3 faces, 4 scalars, 5 total
};

struct RemoteLinks {
  int n;
  const SlabTable *tab;
  const int *lb, *o1, *o2, *o3;          // [n]
  const idx_t *gsend, *grecv;            // [n] element offsets of the ghost message
  const idx_t *esend, *erecv;            // [n] element offsets of the EMF message
  const idx_t *ebase;                    // [n] erecv - face_off(dir,side) the apply kernel expects
  Real *gsbuf;                           // ghost send buffer
  const Real *grbuf;                     // ghost receive buffer
  Real *esbuf;
  const Real *erbuf;
  // peer-to-peer transport: where each link's message goes in the RECEIVER's memory (null = pack
  // into gsbuf / esbuf for the host framework to move)
  Real *const *gdst;                     // [n]
  Real *const *edst;                     // [n]
};

ATHK_HD idx_t ghost_section_off(const Grid &g, bool mhd, int nscal, int o1, int o2, int o3,
                                int section /*0 u, 1..3 faces, 4 scalars, 5 = total*/) {
  idx_t off = 0;
  if (section == 0) return off;
  off += 5 * make_slab(g, -1, o1, o2, o3).count();
  for (int d = 0; d < 3; ++d) {
    if (section == 1 + d) return off;
    if (mhd) off += make_slab(g, d, o1, o2, o3).count();
  }
  if (section == 4) return off;
  off += (idx_t)nscal * make_slab(g, -1, o1, o2, o3).count();
  return off;
}

// LoadBoundaryBufferSameLevel for every remote link: one thread per message element
struct RemoteGhostPack {
  Grid g;
  int mhd, nscal;
  RemoteLinks rl;
  const Real *u, *b, *s;
  int part;   // 0 the whole message; 1 its cell-centred sections (u, s); 2 its face-field sections
  // (link, e): e = element index inside the link's message
  ATHK_HD void operator()(int link, idx_t e) const {
    const int o1 = rl.o1[link], o2 = rl.o2[link], o3 = rl.o3[link], blk = rl.lb[link];
    Real *dst = (rl.gdst ? rl.gdst[link] : rl.gsbuf + rl.gsend[link]) + e;
    const int os = nbr_slot(o1, o2, o3);
    if (e >= rl.tab->sec[os][5]) return;
    const Slab cc = rl.tab->slab[os][0];
    idx_t base = 0;
    const idx_t ncc = cc.count();
    if (part != 0) {   // is e in a section this launch packs?
      const bool in_fc = (e >= rl.tab->sec[os][1] && e < rl.tab->sec[os][4]);
      if ((part == 2) != in_fc) return;
    }
    if (e < 5 * ncc) {
      const int n = (int)(e / ncc);
      idx_t q = e - n * ncc;
      const int li = (int)(q % cc.d1); q /= cc.d1;
      const int lj = (int)(q % cc.d2);
      const int lk = (int)(q / cc.d2);
      *dst = u[((idx_t)blk * 5 + n) * g.sn + gidx(g, cc.s3 + lk, cc.s2 + lj, cc.s1 + li)];
      return;
    }
    base = 5 * ncc;
    if (mhd) {
      for (int d = 0; d < 3; ++d) {
        const Slab f = rl.tab->slab[os][1 + d];
        const idx_t nf = f.count();
        if (e < base + nf) {
          idx_t q = e - base;
          const int li = (int)(q % f.d1); q /= f.d1;
          const int lj = (int)(q % f.d2);
          const int lk = (int)(q / f.d2);
          *dst = b[((idx_t)blk * 3 + d) * g.sn + gidx(g, f.s3 + lk, f.s2 + lj, f.s1 + li)];
          return;
        }
        base += nf;
      }
    }
    if (nscal > 0 && e < base + nscal * ncc) {
      idx_t q = e - base;
      const int n = (int)(q / ncc);
      q -= n * ncc;
      const int li = (int)(q % cc.d1); q /= cc.d1;
      const int lj = (int)(q % cc.d2);
      const int lk = (int)(q / cc.d2);
      *dst = s[((idx_t)blk * nscal + n) * g.sn + gidx(g, cc.s3 + lk, cc.s2 + lj, cc.s1 + li)];
    }
  }
};

// EMF message of a remote link: the packed face section the receiver's EmfApplyPoint indexes
// (face links: that face; x1x2 / x1x3 edge links: the X1 face section; x2x3 edges: X2 section)
ATHK_HD bool emf_link_section(int o1, int o2, int o3, int *dir, int *o_dir) {
  const int nz = (o1 != 0) + (o2 != 0) + (o3 != 0);
  if (nz == 3 || nz == 0) return false;   // corners exchange no EMF
  if (o1 != 0) { *dir = 0; *o_dir = o1; }
  else if (o2 != 0) { *dir = 1; *o_dir = o2; }
  else { *dir = 2; *o_dir = o3; }
  return true;
}

struct RemoteEmfPack {
  EmfSurf sf;
  RemoteLinks rl;
  const Real *buf;   // local packed surfaces [nb][per_block]
  ATHK_HD void operator()(int link, idx_t e) const {
    int dir, od;
    if (!emf_link_section(rl.o1[link], rl.o2[link], rl.o3[link], &dir, &od)) return;
    const idx_t n = (dir == 0) ? sf.n_x1() : (dir == 1 ? sf.n_x2() : sf.n_x3());
    if (e >= n) return;
    // my boundary facing the neighbour: neighbour at +1 -> my outer side
    Real *dst = rl.edst ? rl.edst[link] : rl.esbuf + rl.esend[link];
    dst[e] = buf[(idx_t)rl.lb[link] * sf.per_block() + sf.face_off(dir, od > 0 ? 1 : 0) + e];
  }
};

//========================================================================================
// Same-level ghost fill.  Neighbour on the same GPU: the ghost zone reads straight from the
// neighbour block's active zone (what LoadBoundaryBufferSameLevel + CopyVariableBufferSame
// Process + SetBoundarySameLevel move through a staging buffer; bvals_cc.cpp:287-302,386-420,
// bvals_var.cpp:108-119).  Neighbour on another GPU (table entry <= -2): read the received
// message.  Launched over the full box k in [0,nc3-1] etc.
template <bool REMOTE>
struct GhostFillCC {
  Grid g;
  int nvar;
  Real *u;          // [nb][nvar][sn]
  const int *nbr;
  int mhd, nscal, section;   // section 0 = hydro u, 4 = scalars (remote message layout)
  RemoteLinks rl;
  int mode;                  // 0 all neighbours, 1 same-GPU neighbours only, 2 remote links only
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const int o1 = (i < g.is) ? -1 : (i > g.ie ? 1 : 0);
    const int o2 = (j < g.js) ? -1 : (j > g.je ? 1 : 0);
    const int o3 = (k < g.ks) ? -1 : (k > g.ke ? 1 : 0);
    if (o1 == 0 && o2 == 0 && o3 == 0) return;
    const int id = nbr[blk * 27 + nbr_slot(o1, o2, o3)];
    if (id == -1 || id == NBR_REFINED) return;   // physical boundary; coarser / finer neighbour (athk_refine.h)
    const idx_t dst = gidx(g, k, j, i);
    // (source and destination are the same array: a store would fence the loads behind it, so the
    //  components are loaded five at a time before any is stored)
    const Real *src;
    idx_t sstride;
    if (id >= 0) {
      if (mode == 2) return;
      src = u + (idx_t)id * nvar * g.sn + gidx(g, k - o3 * g.nx3, j - o2 * g.nx2, i - o1 * g.nx1);
      sstride = g.sn;
    } else if (REMOTE && mode != 1) {
      const int link = -id - 2;
      const int os = nbr_slot(o1, o2, o3);
      const Slab cc = rl.tab->slab[os][0];
      const idx_t e = ((idx_t)(k - cc.r3) * cc.d2 + (j - cc.r2)) * cc.d1 + (i - cc.r1);
      src = rl.grbuf + rl.grecv[link] + rl.tab->sec[os][section] + e;
      sstride = cc.count();
    } else {
      return;
    }
    Real *d = u + (idx_t)blk * nvar * g.sn + dst;
    for (int n0 = 0; n0 < nvar; n0 += 5) {
      Real v[5];
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (n0 + q < nvar) v[q] = src[(idx_t)(n0 + q) * sstride];
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (n0 + q < nvar) d[(idx_t)(n0 + q) * g.sn] = v[q];
    }
  }
};

// Face-centred version (bvals_fc.cpp:344-397 load, 583-680 set).
// Launched over k in [0,nc3], j in [0,nc2], i in [0,nc1].
template <bool REMOTE>
struct GhostFillFC {
  Grid g;
  Real *b;          // [nb][3][sn]
  const int *nbr;
  int nscal;
  RemoteLinks rl;
  int mode;         // as GhostFillCC
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    // (the three components are loaded before any is stored: same array on both sides)
    const Real *src[3] = {nullptr, nullptr, nullptr};
    const idx_t dst = gidx(g, k, j, i);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // validity of (k,j,i) for component d
      if (i > g.nc1 - (d == 0 ? 0 : 1)) continue;
      if (j > g.nc2 - (d == 1 ? 0 : 1)) continue;
      if (k > g.nc3 - (d == 2 ? 0 : 1)) continue;
      const int o1 = (i < g.is) ? -1 : (i > g.ie + (d == 0 ? 1 : 0) ? 1 : 0);
      const int o2 = (j < g.js) ? -1 : (j > g.je + (d == 1 ? 1 : 0) ? 1 : 0);
      const int o3 = (k < g.ks) ? -1 : (k > g.ke + (d == 2 ? 1 : 0) ? 1 : 0);
      if (o1 == 0 && o2 == 0 && o3 == 0) continue;
      const int id = nbr[blk * 27 + nbr_slot(o1, o2, o3)];
      if (id == -1) continue;
      if (id >= 0) {
        if (mode == 2) continue;
        src[d] = b + ((idx_t)id * 3 + d) * g.sn + gidx(g, k - o3 * g.nx3, j - o2 * g.nx2, i - o1 * g.nx1);
      } else if (REMOTE && mode != 1) {
        const int link = -id - 2;
        const int os = nbr_slot(o1, o2, o3);
        const Slab f = rl.tab->slab[os][1 + d];
        const idx_t e = ((idx_t)(k - f.r3) * f.d2 + (j - f.r2)) * f.d1 + (i - f.r1);
        src[d] = rl.grbuf + rl.grecv[link] + rl.tab->sec[os][1 + d] + e;
      }
    }
    Real v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (src[d]) v[d] = *src[d];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (src[d]) b[((idx_t)blk * 3 + d) * g.sn + dst] = v[d];
  }
};

//========================================================================================
// Hydro::NewBlockTimeStep, one cell (new_blockdt.cpp:72-124): returns min(dt1,dt2,dt3)
template <bool MHD>
struct CellDt {
  Grid g;
  EosParams eos;
  Geom geo;
  const Real *w, *bcc, *b;
  ATHK_HD Real operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real *wb = w + (idx_t)blk * 5 * g.sn + c;
    const Real d = wb[0], vx = wb[IVX * g.sn], vy = wb[IVY * g.sn], vz = wb[IVZ * g.sn],
               p = wb[IPR * g.sn];
    Real dt1 = geo.dxf[0][blk * geo.PG + i], dt2 = geo.dxf[1][blk * geo.PG + j],
         dt3 = geo.dxf[2][blk * geo.PG + k];
    if (MHD) {
      const Real *bc = bcc + (idx_t)blk * 3 * g.sn + c;
      const Real *bf = b + (idx_t)blk * 3 * g.sn + c;
      const Real b1 = bc[0], b2 = bc[g.sn], b3 = bc[2 * g.sn];
      Real bx = b1 + fabs(bf[0] - b1);
      Real cf = eos.iso ? fast_speed_iso(eos, d, b2, b3, bx) : fast_speed(eos, d, p, b2, b3, bx);
      dt1 /= max_std(0.0, (fabs(vx) + cf));
      bx = b2 + fabs(bf[g.sn] - b2);
      cf = eos.iso ? fast_speed_iso(eos, d, b3, b1, bx) : fast_speed(eos, d, p, b3, b1, bx);
      dt2 /= max_std(0.0, (fabs(vy) + cf));
      bx = b3 + fabs(bf[2 * g.sn] - b3);
      cf = eos.iso ? fast_speed_iso(eos, d, b1, b2, bx) : fast_speed(eos, d, p, b1, b2, bx);
      dt3 /= max_std(0.0, (fabs(vz) + cf));
    } else if (!eos.sr) {   // (SR: the cell widths themselves, maximum characteristic c = 1; new_blockdt.cpp:130)
      const Real cs = eos.iso ? eos.iso_cs : sound_speed(eos, d, p);
      dt1 /= max_std(0.0, (fabs(vx) + cs));
      dt2 /= max_std(0.0, (fabs(vy) + cs));
      dt3 /= max_std(0.0, (fabs(vz) + cs));
    }
    return min_std(min_std(dt1, dt2), dt3);
  }
};

}  // namespace athk
#endif  // ATHK_OPS_H_
