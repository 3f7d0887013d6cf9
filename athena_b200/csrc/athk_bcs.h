// athk_bcs.h -- physical boundary conditions on the device: BoundaryValues::ApplyPhysicalBoundaries
// (bvals/bvals.cpp:453-660) with the built-in reflecting / outflow functions for cell-centred
// primitives (bvals/cc/hydro/reflect_hydro.cpp, bvals/cc/outflow_cc.cpp; passive scalars
// bvals/cc/reflect_cc.cpp) and the face-centred field (bvals/fc/reflect_fc.cpp, outflow_fc.cpp),
// followed on the same ghost slab by Field::CalculateCellCenteredField, EquationOfState::
// PrimitiveToConserved (eos/adiabatic_hydro.cpp:86-120, adiabatic_mhd.cpp:99-135) and
// PassiveScalarPrimitiveToConserved (eos/eos_scalars.cpp:133-152).
//
// Order matters and is the reference's: x1 faces, then x2 (over the x1 ghosts too), then x3 (over
// both); a transverse direction is extended into its ghost zones only where that side is NOT a
// physical boundary (it was filled by the neighbour exchange).  One launch per (face, step), all
// MeshBlocks at once; blocks whose face is not a physical boundary return immediately.
#ifndef ATHK_BCS_H_
#define ATHK_BCS_H_

#include "athk_ops.h"

namespace athk {

enum { BC_NONE = 0, BC_REFLECT = 1, BC_OUTFLOW = 2, BC_USER = 3 };   // == enum athk_bc

// transverse limits of the pass over faces of direction DIR for one block (bvals.cpp:457-470,560)
template <int DIR>
ATHK_HD void bc_limits(const Grid &g, const int *fl, int *al, int *au, int *bl, int *bu) {
  // (a, b) = the two transverse directions in the order (x1,x2,x3) minus DIR
  if (DIR == 0) {
    *al = g.js - (fl[2] == BC_NONE ? g.ng : 0);
    *au = g.je + (fl[3] == BC_NONE ? g.ng : 0);
    *bl = g.ks - (fl[4] == BC_NONE ? g.ng : 0);
    *bu = g.ke + (fl[5] == BC_NONE ? g.ng : 0);
  } else if (DIR == 1) {
    *al = g.is - g.ng;
    *au = g.ie + g.ng;
    *bl = g.ks - (fl[4] == BC_NONE ? g.ng : 0);
    *bu = g.ke + (fl[5] == BC_NONE ? g.ng : 0);
  } else {
    *al = g.is - g.ng;
    *au = g.ie + g.ng;
    *bl = g.js - g.ng;
    *bu = g.je + g.ng;
  }
}

// Step 1: fill w, r and b in the ghost slab of face (DIR, side).  Launched with the DIR coordinate
// running over the layer m = 1..ng and the transverse coordinates over the whole padded box.
template <int DIR>
struct PhysBCFillPoint {
  Grid g;
  int side;            // 0 inner, 1 outer
  const int *flags;    // [nb][6]
  int mhd, ns;
  Real *w, *b, *r;
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const int *fl = flags + blk * 6;
    const int bc = fl[2 * DIR + side];
    if (bc == BC_NONE || bc == BC_USER) return;   // (user faces are filled by the enrolled function)
    const int m = (DIR == 0) ? i : (DIR == 1 ? j : k);
    const int ta = (DIR == 0) ? j : i;                  // first transverse coordinate
    const int tb = (DIR == 2) ? j : k;                  // second
    int al, au, bl, bu;
    bc_limits<DIR>(g, fl, &al, &au, &bl, &bu);
    const int n_s = (DIR == 0) ? g.is : (DIR == 1 ? g.js : g.ks);
    const int n_e = (DIR == 0) ? g.ie : (DIR == 1 ? g.je : g.ke);
    const bool refl = (bc == BC_REFLECT);
    // cell layers: ghost gn <- source sn ; normal faces: ghost gf <- source sf
    const int gn = side ? n_e + m : n_s - m;
    const int sn = refl ? (side ? n_e - m + 1 : n_s + m - 1) : (side ? n_e : n_s);
    const int gf = side ? n_e + m + 1 : n_s - m;
    const int sf = refl ? (side ? n_e - m + 1 : n_s + m) : (side ? n_e + 1 : n_s);
    const idx_t sdir = (DIR == 0) ? 1 : (DIR == 1 ? g.sj : g.sk);
    // index of (n = 0, ta, tb)
    const idx_t base = (DIR == 0) ? gidx(g, tb, ta, 0) : (DIR == 1 ? gidx(g, tb, 0, ta) : gidx(g, 0, tb, ta));
    const bool in_a = (ta >= al && ta <= au), in_b = (tb >= bl && tb <= bu);
    if (in_a && in_b) {
      Real *wb = w + (idx_t)blk * 5 * g.sn + base;
#pragma unroll
      for (int n = 0; n < 5; ++n) {
        const Real v = wb[(idx_t)n * g.sn + sn * sdir];
        wb[(idx_t)n * g.sn + gn * sdir] = (refl && n == IVX + DIR) ? -v : v;   // reflect the normal velocity
      }
      for (int n = 0; n < ns; ++n) {
        Real *rb = r + ((idx_t)blk * ns + n) * g.sn + base;
        rb[gn * sdir] = rb[sn * sdir];
      }
      if (mhd) {   // normal field component
        Real *bn = b + ((idx_t)blk * 3 + DIR) * g.sn + base;
        const Real v = bn[sf * sdir];
        bn[gf * sdir] = refl ? -v : v;
      }
    }
    if (mhd) {     // transverse components carry one more face in their own direction
      const int da = (DIR == 0) ? 1 : 0, db = (DIR == 2) ? 1 : 2;   // their component indices
      if (ta >= al && ta <= au + 1 && in_b) {
        Real *bt = b + ((idx_t)blk * 3 + da) * g.sn + base;
        bt[gn * sdir] = bt[sn * sdir];
      }
      if (in_a && tb >= bl && tb <= bu + 1) {
        Real *bt = b + ((idx_t)blk * 3 + db) * g.sn + base;
        bt[gn * sdir] = bt[sn * sdir];
      }
    }
  }
};

// Step 2: bcc, U(W) and s = r*rho on the cells of the same slab
template <int DIR>
struct PhysBCConsPoint {
  Grid g;
  int side;
  const int *flags;
  EosParams eos;
  int mhd, ns;
  const Real *w, *b, *r;
  Real *bcc, *u, *s;
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const int *fl = flags + blk * 6;
    if (fl[2 * DIR + side] == BC_NONE) return;
    const int m = (DIR == 0) ? i : (DIR == 1 ? j : k);
    const int ta = (DIR == 0) ? j : i;
    const int tb = (DIR == 2) ? j : k;
    int al, au, bl, bu;
    bc_limits<DIR>(g, fl, &al, &au, &bl, &bu);
    if (ta < al || ta > au || tb < bl || tb > bu) return;
    const int n_s = (DIR == 0) ? g.is : (DIR == 1 ? g.js : g.ks);
    const int n_e = (DIR == 0) ? g.ie : (DIR == 1 ? g.je : g.ke);
    const int gn = side ? n_e + m : n_s - m;
    const idx_t c = (DIR == 0) ? gidx(g, tb, ta, gn) : (DIR == 1 ? gidx(g, tb, gn, ta) : gidx(g, gn, tb, ta));
    const Real *wb = w + (idx_t)blk * 5 * g.sn + c;
    const Real w_d = wb[0], w_vx = wb[g.sn], w_vy = wb[2 * g.sn], w_vz = wb[3 * g.sn], w_p = wb[4 * g.sn];
    const Real igm1 = 1.0 / eos.gm1;
    Real *ub = u + (idx_t)blk * 5 * g.sn + c;
    if (eos.sr) {   // EquationOfState::PrimitiveToConserved, adiabatic_hydro_sr.cpp:229-262
      const Real gamma_prime = eos.gamma / (eos.gamma - 1.0);
      const Real u0 = sqrt(1.0 + sqr(w_vx) + sqr(w_vy) + sqr(w_vz));
      const Real wgas_u0 = (w_d + gamma_prime * w_p) * u0;
      ub[0] = w_d * u0;
      ub[4 * g.sn] = wgas_u0 * u0 - w_p;
      ub[g.sn] = wgas_u0 * w_vx;
      ub[2 * g.sn] = wgas_u0 * w_vy;
      ub[3 * g.sn] = wgas_u0 * w_vz;
      for (int n = 0; n < ns; ++n) {   // PassiveScalarPrimitiveToConserved: s = r * u(IDN), eos_scalars.cpp:96-112
        const idx_t o = ((idx_t)blk * ns + n) * g.sn + c;
        s[o] = r[o] * (w_d * u0);
      }
      return;
    }
    ub[0] = w_d;
    ub[g.sn] = w_vx * w_d;
    ub[2 * g.sn] = w_vy * w_d;
    ub[3 * g.sn] = w_vz * w_d;
    if (mhd) {
      const Real *bb = b + (idx_t)blk * 3 * g.sn + c;
      Real *bc = bcc + (idx_t)blk * 3 * g.sn + c;
      const Real bcc1 = 0.5 * bb[0] + 0.5 * bb[1];                         // field.cpp:139-142
      const Real bcc2 = 0.5 * bb[g.sn] + 0.5 * bb[g.sn + g.sj];
      const Real bcc3 = 0.5 * bb[2 * g.sn] + 0.5 * bb[2 * g.sn + g.sk];
      bc[0] = bcc1;
      bc[g.sn] = bcc2;
      bc[2 * g.sn] = bcc3;
      if (!eos.iso)   // (isothermal EOS: no energy, eos/isothermal_mhd.cpp:88-112)
        ub[4 * g.sn] = w_p * igm1 + 0.5 * (w_d * (sqr(w_vx) + sqr(w_vy) + sqr(w_vz)) +
                                           (sqr(bcc1) + sqr(bcc2) + sqr(bcc3)));
    } else {
      if (!eos.iso) ub[4 * g.sn] = w_p * igm1 + 0.5 * w_d * (sqr(w_vx) + sqr(w_vy) + sqr(w_vz));
    }
    for (int n = 0; n < ns; ++n) {
      const idx_t o = ((idx_t)blk * ns + n) * g.sn + c;
      s[o] = r[o] * w_d;
    }
  }
};

}  // namespace athk
#endif  // ATHK_BCS_H_
