// athk_refine.h -- static mesh refinement, cell-centred (hydro) variables: the arithmetic of
//   MeshRefinement::RestrictCellCenteredValues      mesh/mesh_refinement.cpp:116-186
//   MeshRefinement::ProlongateCellCenteredValues    mesh/mesh_refinement.cpp:464-545
//   CellCenteredBoundaryVariable::LoadFluxBoundaryBufferToCoarser / SetFluxBoundaryFromFiner
//                                                   bvals/cc/flux_correction_cc.cpp:77-160,222-275
//   ConservedToPrimitive / PrimitiveToConserved on the coarse buffer and on the prolongated ghost
//   zones (bvals/bvals_refine.cpp:437-443,646-660; eos/adiabatic_hydro.cpp:38-120)
// as "box operations".  The host restates the reference's index logic (which box of which block a
// neighbour of which level fills: bvals/cc/bvals_cc.cpp:309-560, bvals/bvals_refine.cpp:95-330) into a
// list of boxes per phase; one launch executes a whole phase, one thread per destination cell.
// Boxes of one phase may overlap only where they write identical values (the reference's loop over
// neighbours rewrites such cells too).
//
// Coarse buffers (Hydro::coarse_cons_, coarse_prim_): dense [block][5][cn3][cn2][cn1] with
// cn = nx/2 + 2*NGHOST and first active index cis = NGHOST (mesh/meshblock.cpp:66-100); of those
// ghost layers cnghost = (NGHOST+1)/2 + 1 are used.
#ifndef ATHK_REFINE_H_
#define ATHK_REFINE_H_

#include "athk_ops.h"

namespace athk {

enum {
  BOX_RESTRICT = 1,       // own u -> own coarse_cons over the coarse box [lo,hi]
  BOX_FROM_FINER = 2,     // u[dst] box [lo,hi]  <- coarse_cons[src] from coarse origin so
  BOX_FROM_COARSER = 3,   // coarse_cons[dst] box [lo,hi] <- u[src] from fine origin so
  BOX_C2P_COARSE = 4,     // coarse_cons -> coarse_prim over the coarse box (floors written back)
  BOX_PROLONG = 5,        // coarse_prim box [lo,hi] -> w, eight fine cells per coarse cell
  BOX_P2C_FINE = 6,       // w -> u over the fine box [lo,hi]
  BOX_FLUX_FROM_FINER = 7, // flux[dir][dst] face box [lo,hi] <- area-weighted flux[dir][src], fine origin so
  BOX_COARSE_BC = 8       // reflecting / outflow function on coarse_prim, one ghost layer (bvals_refine.cpp:445-530):
                          // box [lo,hi] = the ghost layer, dir = its direction, src = 1 reflect / 2 outflow,
                          // so[dir] = the active coarse cell it is copied from
};

struct BoxOp {   // == athk_box_op (include/athk.h)
  int kind, dst, src, dir;
  int lo[3], hi[3];   // destination box, inclusive, (i, j, k)
  int so[3];          // source origin (i, j, k) matching lo
  int pad;
};

struct CoarseGrid {
  int cng;                       // cnghost
  int cn1, cn2, cn3;             // cells including coarse ghosts
  int cis, cjs, cks;             // first active coarse index (= NGHOST)
  idx_t csj, csk, csn;
  int CPG;                       // stride of the per-block coarse coordinate arrays
};

ATHK_HD idx_t cidx(const CoarseGrid &cg, int k, int j, int i) { return (idx_t)k * cg.csk + (idx_t)j * cg.csj + i; }

struct BoxOpPoint {
  Grid g;
  CoarseGrid cg;
  Geom geo;
  EosParams eos;
  const BoxOp *ops;
  Real *u;                 // [nb][5][sn]   conserved register the exchange acts on
  Real *w;                 // [nb][5][sn]
  Real *cu, *cw;           // [nb][5][csn]  coarse_cons_, coarse_prim_
  Real *flx[3];            // [nb][5][sn]
  const Real *xv;          // [nb][3][PG]   fine x1v, x2v, x3v
  const Real *cxv;         // [nb][3][CPG]  coarse (pcoarsec) x1v, x2v, x3v
  ATHK_HD void operator()(int opi, idx_t e) const {
    const BoxOp op = ops[opi];
    const int n1 = op.hi[0] - op.lo[0] + 1, n2 = op.hi[1] - op.lo[1] + 1, n3 = op.hi[2] - op.lo[2] + 1;
    if (n1 <= 0 || n2 <= 0 || n3 <= 0 || e >= (idx_t)n1 * n2 * n3) return;
    const int di = (int)(e % n1);
    const int dj = (int)((e / n1) % n2);
    const int dk = (int)(e / ((idx_t)n1 * n2));
    const int i = op.lo[0] + di, j = op.lo[1] + dj, k = op.lo[2] + dk;
    switch (op.kind) {
      case BOX_RESTRICT: restrict_cell(op.dst, k, j, i); break;
      case BOX_FROM_FINER: {
        const idx_t s = cidx(cg, op.so[2] + dk, op.so[1] + dj, op.so[0] + di), d = gidx(g, k, j, i);
        Real v[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) v[n] = cu[((idx_t)op.src * 5 + n) * cg.csn + s];
#pragma unroll
        for (int n = 0; n < 5; ++n) u[((idx_t)op.dst * 5 + n) * g.sn + d] = v[n];
        break;
      }
      case BOX_FROM_COARSER: {
        const idx_t s = gidx(g, op.so[2] + dk, op.so[1] + dj, op.so[0] + di), d = cidx(cg, k, j, i);
        Real v[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) v[n] = u[((idx_t)op.src * 5 + n) * g.sn + s];
#pragma unroll
        for (int n = 0; n < 5; ++n) cu[((idx_t)op.dst * 5 + n) * cg.csn + d] = v[n];
        break;
      }
      case BOX_C2P_COARSE: {
        const idx_t c = cidx(cg, k, j, i);
        Real *ub = cu + (idx_t)op.dst * 5 * cg.csn + c, *wb = cw + (idx_t)op.dst * 5 * cg.csn + c;
        Real uc[5], wc[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) uc[n] = ub[n * cg.csn];
        const Real d0 = uc[IDN], e0 = uc[IEN];
        cons2prim_hydro(eos_of_block(eos, op.dst), uc, wc);
        if (uc[IDN] != d0) ub[IDN * cg.csn] = uc[IDN];
        if (uc[IEN] != e0) ub[IEN * cg.csn] = uc[IEN];
#pragma unroll
        for (int n = 0; n < 5; ++n) wb[n * cg.csn] = wc[n];
        break;
      }
      case BOX_PROLONG: prolong_cell(op.dst, k, j, i); break;
      case BOX_P2C_FINE: {
        const idx_t c = gidx(g, k, j, i);
        const Real *wb = w + (idx_t)op.dst * 5 * g.sn + c;
        Real *ub = u + (idx_t)op.dst * 5 * g.sn + c;
        const Real w_d = wb[0], w_vx = wb[g.sn], w_vy = wb[2 * g.sn], w_vz = wb[3 * g.sn], w_p = wb[4 * g.sn];
        const Real igm1 = 1.0 / eos.gm1;   // eos/adiabatic_hydro.cpp:86-120
        ub[0] = w_d;
        ub[g.sn] = w_vx * w_d;
        ub[2 * g.sn] = w_vy * w_d;
        ub[3 * g.sn] = w_vz * w_d;
        ub[4 * g.sn] = w_p * igm1 + 0.5 * w_d * (sqr(w_vx) + sqr(w_vy) + sqr(w_vz));
        break;
      }
      case BOX_FLUX_FROM_FINER: flux_cell(op, k, j, i, dk, dj, di); break;
      case BOX_COARSE_BC: {   // bvals/cc/outflow_cc.cpp, bvals/cc/hydro/reflect_hydro.cpp with ngh = 1
        const int si = (op.dir == 0) ? op.so[0] : i, sj = (op.dir == 1) ? op.so[1] : j, sk = (op.dir == 2) ? op.so[2] : k;
        const idx_t s = cidx(cg, sk, sj, si), d = cidx(cg, k, j, i);
        Real *wb = cw + (idx_t)op.dst * 5 * cg.csn;
        Real v[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) v[n] = wb[n * cg.csn + s];
        if (op.src == 1) v[IVX + op.dir] = -v[IVX + op.dir];   // reflect the normal velocity
#pragma unroll
        for (int n = 0; n < 5; ++n) wb[n * cg.csn + d] = v[n];
        break;
      }
      default: break;
    }
  }
  // mesh_refinement.cpp:123-151 (3-D, Cartesian volumes dx1f*dx2f*dx3f, coordinates/cartesian.cpp)
  ATHK_HD void restrict_cell(int blk, int ck, int cj, int ci) const {
    const int k = (ck - cg.cks) * 2 + g.ks, j = (cj - cg.cjs) * 2 + g.js, i = (ci - cg.cis) * 2 + g.is;
    const Real *dx1 = geo.dxf[0] + blk * geo.PG, *dx2 = geo.dxf[1] + blk * geo.PG, *dx3 = geo.dxf[2] + blk * geo.PG;
    // fvol_[a][b](i) = CellVolume(k+a, j+b, i) = dx1f(i)*dx2f(j+b)*dx3f(k+a)
    const Real v00i = dx1[i] * dx2[j] * dx3[k], v00p = dx1[i + 1] * dx2[j] * dx3[k];
    const Real v01i = dx1[i] * dx2[j + 1] * dx3[k], v01p = dx1[i + 1] * dx2[j + 1] * dx3[k];
    const Real v10i = dx1[i] * dx2[j] * dx3[k + 1], v10p = dx1[i + 1] * dx2[j] * dx3[k + 1];
    const Real v11i = dx1[i] * dx2[j + 1] * dx3[k + 1], v11p = dx1[i + 1] * dx2[j + 1] * dx3[k + 1];
    const Real tvol = ((v00i + v01i) + (v00p + v01p)) + ((v10i + v11i) + (v10p + v11p));
    const idx_t c = gidx(g, k, j, i), cc = cidx(cg, ck, cj, ci);
    Real r[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      const Real *f = u + ((idx_t)blk * 5 + n) * g.sn + c;
      r[n] = (((f[0] * v00i + f[g.sj] * v01i) + (f[1] * v00p + f[g.sj + 1] * v01p)) +
              ((f[g.sk] * v10i + f[g.sk + g.sj] * v11i) + (f[g.sk + 1] * v10p + f[g.sk + g.sj + 1] * v11p))) / tvol;
    }
#pragma unroll
    for (int n = 0; n < 5; ++n) cu[((idx_t)blk * 5 + n) * cg.csn + cc] = r[n];
  }
  ATHK_HD static Real sign_of(Real x) { return (x < 0.0) ? -1.0 : 1.0; }   // SIGN(), defs.hpp.in:150
  // mesh_refinement.cpp:470-530 (3-D): minmod-limited gradients of the coarse primitives
  ATHK_HD void prolong_cell(int blk, int k, int j, int i) const {
    const int fk = (k - cg.cks) * 2 + g.ks, fj = (j - cg.cjs) * 2 + g.js, fi = (i - cg.cis) * 2 + g.is;
    const Real *cx1 = cxv + ((idx_t)blk * 3 + 0) * cg.CPG, *cx2 = cxv + ((idx_t)blk * 3 + 1) * cg.CPG,
               *cx3 = cxv + ((idx_t)blk * 3 + 2) * cg.CPG;
    const Real *fx1 = xv + ((idx_t)blk * 3 + 0) * geo.PG, *fx2 = xv + ((idx_t)blk * 3 + 1) * geo.PG,
               *fx3 = xv + ((idx_t)blk * 3 + 2) * geo.PG;
    const Real x3c = cx3[k], dx3m = x3c - cx3[k - 1], dx3p = cx3[k + 1] - x3c;
    const Real dx3fm = x3c - fx3[fk], dx3fp = fx3[fk + 1] - x3c;
    const Real x2c = cx2[j], dx2m = x2c - cx2[j - 1], dx2p = cx2[j + 1] - x2c;
    const Real dx2fm = x2c - fx2[fj], dx2fp = fx2[fj + 1] - x2c;
    const Real x1c = cx1[i], dx1m = x1c - cx1[i - 1], dx1p = cx1[i + 1] - x1c;
    const Real dx1fm = x1c - fx1[fi], dx1fp = fx1[fi + 1] - x1c;
    const idx_t cc = cidx(cg, k, j, i), fc = gidx(g, fk, fj, fi);
    for (int n = 0; n < 5; ++n) {
      const Real *cb = cw + ((idx_t)blk * 5 + n) * cg.csn + cc;
      Real *fb = w + ((idx_t)blk * 5 + n) * g.sn + fc;
      const Real ccval = cb[0];
      const Real gx1m = (ccval - cb[-1]) / dx1m;
      const Real gx1p = (cb[1] - ccval) / dx1p;
      const Real gx1c = 0.5 * (sign_of(gx1m) + sign_of(gx1p)) * min_std(fabs(gx1m), fabs(gx1p));
      const Real gx2m = (ccval - cb[-cg.csj]) / dx2m;
      const Real gx2p = (cb[cg.csj] - ccval) / dx2p;
      const Real gx2c = 0.5 * (sign_of(gx2m) + sign_of(gx2p)) * min_std(fabs(gx2m), fabs(gx2p));
      const Real gx3m = (ccval - cb[-cg.csk]) / dx3m;
      const Real gx3p = (cb[cg.csk] - ccval) / dx3p;
      const Real gx3c = 0.5 * (sign_of(gx3m) + sign_of(gx3p)) * min_std(fabs(gx3m), fabs(gx3p));
      fb[0] = ccval - (gx1c * dx1fm + gx2c * dx2fm + gx3c * dx3fm);
      fb[1] = ccval + (gx1c * dx1fp - gx2c * dx2fm - gx3c * dx3fm);
      fb[g.sj] = ccval - (gx1c * dx1fm - gx2c * dx2fp + gx3c * dx3fm);
      fb[g.sj + 1] = ccval + (gx1c * dx1fp + gx2c * dx2fp - gx3c * dx3fm);
      fb[g.sk] = ccval - (gx1c * dx1fm + gx2c * dx2fm - gx3c * dx3fp);
      fb[g.sk + 1] = ccval + (gx1c * dx1fp - gx2c * dx2fm + gx3c * dx3fp);
      fb[g.sk + g.sj] = ccval - (gx1c * dx1fm - gx2c * dx2fp - gx3c * dx3fp);
      fb[g.sk + g.sj + 1] = ccval + (gx1c * dx1fp + gx2c * dx2fp + gx3c * dx3fp);
    }
  }
  // flux_correction_cc.cpp:93-160 (restriction on the fine block) + :222-275 (set on the coarse one)
  ATHK_HD void flux_cell(const BoxOp &op, int k, int j, int i, int dk, int dj, int di) const {
    const int sblk = op.src;
    const Real *dx1 = geo.dxf[0] + sblk * geo.PG, *dx2 = geo.dxf[1] + sblk * geo.PG, *dx3 = geo.dxf[2] + sblk * geo.PG;
    // fine cell of the source block under destination face (di,dj,dk)
    const int fi = op.so[0] + (op.dir == 0 ? 0 : 2 * di), fj = op.so[1] + (op.dir == 1 ? 0 : 2 * dj),
              fk = op.so[2] + (op.dir == 2 ? 0 : 2 * dk);
    Real a0, a1, a2, a3;
    idx_t o1, o2, o3;   // offsets of the 2nd, 3rd, 4th fine face in the reference's order
    if (op.dir == 0) {          // (k,j), (k,j+1), (k+1,j), (k+1,j+1): Face1Area = dx2f*dx3f
      a0 = dx2[fj] * dx3[fk]; a1 = dx2[fj + 1] * dx3[fk]; a2 = dx2[fj] * dx3[fk + 1]; a3 = dx2[fj + 1] * dx3[fk + 1];
      o1 = g.sj; o2 = g.sk; o3 = g.sk + g.sj;
    } else if (op.dir == 1) {   // (k,i), (k,i+1), (k+1,i), (k+1,i+1): Face2Area = dx1f*dx3f
      a0 = dx1[fi] * dx3[fk]; a1 = dx1[fi + 1] * dx3[fk]; a2 = dx1[fi] * dx3[fk + 1]; a3 = dx1[fi + 1] * dx3[fk + 1];
      o1 = 1; o2 = g.sk; o3 = g.sk + 1;
    } else {                    // (j,i), (j,i+1), (j+1,i), (j+1,i+1): Face3Area = dx1f*dx2f
      a0 = dx1[fi] * dx2[fj]; a1 = dx1[fi + 1] * dx2[fj]; a2 = dx1[fi] * dx2[fj + 1]; a3 = dx1[fi + 1] * dx2[fj + 1];
      o1 = 1; o2 = g.sj; o3 = g.sj + 1;
    }
    const Real tarea = a0 + a1 + a2 + a3;
    const idx_t s = gidx(g, fk, fj, fi), d = gidx(g, k, j, i);
    Real r[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      const Real *f = flx[op.dir] + ((idx_t)sblk * 5 + n) * g.sn + s;
      r[n] = (f[0] * a0 + f[o1] * a1 + f[o2] * a2 + f[o3] * a3) / tarea;
    }
#pragma unroll
    for (int n = 0; n < 5; ++n) flx[op.dir][((idx_t)op.dst * 5 + n) * g.sn + d] = r[n];
  }
};

}  // namespace athk
#endif  // ATHK_REFINE_H_
