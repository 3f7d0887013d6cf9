// athk_refine.h -- static mesh refinement: the arithmetic of
//   MeshRefinement::RestrictCellCenteredValues      mesh/mesh_refinement.cpp:116-186
//   MeshRefinement::ProlongateCellCenteredValues    mesh/mesh_refinement.cpp:464-545
//   MeshRefinement::RestrictFieldX1/X2/X3           mesh/mesh_refinement.cpp:272-460
//   MeshRefinement::ProlongateSharedFieldX1/X2/X3   mesh/mesh_refinement.cpp:776-1220 (Cartesian branch)
//   MeshRefinement::ProlongateInternalField         mesh/mesh_refinement.cpp:1223-1340 (Toth & Roe)
//   FaceCenteredBoundaryVariable::LoadFluxBoundaryBufferToCoarser / SetFluxBoundaryFromFiner /
//   ClearCoarseFluxBoundary / AverageFluxBoundary   bvals/fc/flux_correction_fc.cpp:315-570,944-1190,1214-1541
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
  BOX_COARSE_BC = 8,      // reflecting / outflow function on coarse_prim, one ghost layer (bvals_refine.cpp:445-530):
                          // box [lo,hi] = the ghost layer, dir = its direction, src = 1 reflect / 2 outflow,
                          // so[dir] = the active coarse cell it is copied from
  // ---- face-centred field (dir = component: 0 x1f, 1 x2f, 2 x3f)
  BOX_FC_COPY = 9,          // b[dst].dir box [lo,hi] <- b[src].dir from origin so (SetBoundarySameLevel on a refined mesh)
  BOX_FC_RESTRICT = 10,     // own b.dir -> own coarse_b.dir over the coarse face box [lo,hi]
  BOX_FC_FROM_FINER = 11,   // b[dst].dir box <- coarse_b[src].dir from coarse origin so
  BOX_FC_FROM_COARSER = 12, // coarse_b[dst].dir box <- b[src].dir from fine origin so
  BOX_FC_PROLONG_SHARED = 13,   // coarse_b.dir coarse face box -> the four fine faces on each coarse face
  BOX_FC_PROLONG_INTERNAL = 14, // coarse cell box: the twelve internal fine faces of each coarse cell
  // ---- EMF correction across levels (dir = component: 0 e1, 1 e2, 2 e3)
  BOX_EMF_CLEAR = 15,       // e[dst].dir box = 0
  BOX_EMF_FROM_FINER = 16,  // e[dst].dir box += length-weighted pair of the finer src's packed surface EMFs,
                            // fine origin so (in the pair's direction and in the others: stride 2)
  BOX_EMF_SCALE = 17,       // e[dst].dir box *= 1/src
  BOX_COARSE_BC_FC = 18     // reflecting / outflow function on coarse_b_.dir, one ghost layer of boundary direction
                            // pad (bvals/fc/reflect_fc.cpp, outflow_fc.cpp with ngh = 1): box [lo,hi] = the ghost
                            // layer, copied from index so[pad]; src = 1 negates (the reflected normal component)
};

struct BoxOp {   // == athk_box_op (include/athk.h)
  int kind, dst, src, dir;
  int lo[3], hi[3];   // destination box, inclusive, (i, j, k)
  int so[3];          // source origin (i, j, k) matching lo
  int pad;
  int peer;           // 0: block src lives in this context; r + 1: in the context of GPU r (PeerView)
};

// What a box reads of its SOURCE block.  Entry 0 is the context itself; entry r + 1 the context of GPU r
// of the same process, whose arrays the kernel reads over NVLink (peer access; athk_refine_set_peers).
struct PeerView {
  const Real *u, *s, *b;          // registers the exchange acts on
  const Real *cu, *cs, *cb;       // coarse buffers
  const Real *emfbuf;             // packed raw surface EMFs
  const Real *flx[3], *sflx[3];
  const Real *dxf[3];             // [nb][PG] cell widths of that context's blocks
};
enum { ATHK_MAX_PEERS = 8 };

struct CoarseGrid {
  int cng;                       // cnghost
  int cn1, cn2, cn3;             // cells including coarse ghosts
  int cis, cjs, cks;             // first active coarse index (= NGHOST)
  idx_t csj, csk, csn;
  int CPG;                       // stride of the per-block coarse coordinate arrays
  idx_t cfj, cfk, cfn;           // coarse face arrays (coarse_b_): one box of (cn1+1, cn2+1, cn3+1) per component
};

ATHK_HD idx_t cidx(const CoarseGrid &cg, int k, int j, int i) { return (idx_t)k * cg.csk + (idx_t)j * cg.csj + i; }
ATHK_HD idx_t cfidx(const CoarseGrid &cg, int k, int j, int i) { return (idx_t)k * cg.cfk + (idx_t)j * cg.cfj + i; }

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
  // passive scalars (ns = 0: none): boxes with pad = 1 act on them instead of the hydro variables
  int ns;
  Real *s;                 // [nb][ns][sn]  the scalar register the exchange acts on
  Real *r;                 // [nb][ns][sn]
  Real *cs, *cr;           // [nb][ns][csn] coarse_s_, coarse_r_
  Real *sflx[3];           // [nb][ns][sn]
  // MHD (null / 0 for hydro)
  int mhd;
  Real *b;                 // [nb][3][sn]   the field register the exchange acts on
  Real *bcc;               // [nb][3][sn]
  Real *cb;                // [nb][3][cfn]  coarse_b_
  Real *cbcc;              // [nb][3][csn]  coarse_bcc_
  Real *emf;               // [nb][3][sn]   edge EMFs
  const Real *emfbuf;      // [nb][per_block] packed raw surface EMFs (EmfPackPoint)
  EmfSurf sf;
  PeerView pv[1 + ATHK_MAX_PEERS];
  ATHK_HD void operator()(int opi, idx_t e) const {
    const BoxOp op = ops[opi];
    const int n1 = op.hi[0] - op.lo[0] + 1, n2 = op.hi[1] - op.lo[1] + 1, n3 = op.hi[2] - op.lo[2] + 1;
    if (n1 <= 0 || n2 <= 0 || n3 <= 0 || e >= (idx_t)n1 * n2 * n3) return;
    const int di = (int)(e % n1);
    const int dj = (int)((e / n1) % n2);
    const int dk = (int)(e / ((idx_t)n1 * n2));
    const int i = op.lo[0] + di, j = op.lo[1] + dj, k = op.lo[2] + dk;
    switch (op.kind) {
      case BOX_RESTRICT:
        if (op.pad == 1) restrict_cell(s, cs, ns, op.dst, k, j, i);
        else restrict_cell(u, cu, 5, op.dst, k, j, i);
        break;
      case BOX_FROM_FINER: {
        const idx_t so = cidx(cg, op.so[2] + dk, op.so[1] + dj, op.so[0] + di), d = gidx(g, k, j, i);
        const int nv = (op.pad == 1) ? ns : 5;
        const Real *src = (op.pad == 1) ? pv[op.peer].cs : pv[op.peer].cu;
        Real *dst = (op.pad == 1) ? s : u;
        for (int n = 0; n < nv; ++n) dst[((idx_t)op.dst * nv + n) * g.sn + d] = src[((idx_t)op.src * nv + n) * cg.csn + so];
        break;
      }
      case BOX_FROM_COARSER: {
        const idx_t so = gidx(g, op.so[2] + dk, op.so[1] + dj, op.so[0] + di), d = cidx(cg, k, j, i);
        const int nv = (op.pad == 1) ? ns : 5;
        const Real *src = (op.pad == 1) ? pv[op.peer].s : pv[op.peer].u;
        Real *dst = (op.pad == 1) ? cs : cu;
        for (int n = 0; n < nv; ++n) dst[((idx_t)op.dst * nv + n) * cg.csn + d] = src[((idx_t)op.src * nv + n) * g.sn + so];
        break;
      }
      case BOX_C2P_COARSE: {
        const idx_t c = cidx(cg, k, j, i);
        Real *ub = cu + (idx_t)op.dst * 5 * cg.csn + c, *wb = cw + (idx_t)op.dst * 5 * cg.csn + c;
        Real uc[5], wc[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) uc[n] = ub[n * cg.csn];
        const Real d0 = uc[IDN], e0 = uc[IEN];
        if (mhd) {   // CalculateCellCenteredField on coarse_b_ (field.cpp:139-142, uniform average), then W(U)
          const idx_t fc = cfidx(cg, k, j, i);
          const Real *bb = cb + (idx_t)op.dst * 3 * cg.cfn + fc;
          const Real bc1 = 0.5 * bb[0] + 0.5 * bb[1];
          const Real bc2 = 0.5 * bb[cg.cfn] + 0.5 * bb[cg.cfn + cg.cfj];
          const Real bc3 = 0.5 * bb[2 * cg.cfn] + 0.5 * bb[2 * cg.cfn + cg.cfk];
          Real *bc = cbcc + (idx_t)op.dst * 3 * cg.csn + c;
          bc[0] = bc1; bc[cg.csn] = bc2; bc[2 * cg.csn] = bc3;
          if (eos.iso) cons2prim_iso(eos_of_block(eos, op.dst), uc, wc);
          else cons2prim_mhd(eos_of_block(eos, op.dst), uc, bc1, bc2, bc3, wc);
        } else {
          if (eos.iso) cons2prim_iso(eos_of_block(eos, op.dst), uc, wc);
          else cons2prim_hydro(eos_of_block(eos, op.dst), uc, wc);
        }
        if (uc[IDN] != d0) ub[IDN * cg.csn] = uc[IDN];
        if (uc[IEN] != e0) ub[IEN * cg.csn] = uc[IEN];
#pragma unroll
        for (int n = 0; n < 5; ++n) wb[n * cg.csn] = wc[n];
        // PassiveScalarConservedToPrimitive on the coarse buffer (bvals_refine.cpp:443-449, eos_scalars.cpp:31-60)
        for (int n = 0; n < ns; ++n) {
          const idx_t o = ((idx_t)op.dst * ns + n) * cg.csn + c;
          const Real s0 = cs[o];
          const Real sfl = eos_sfloor(eos, op.dst);
          const Real sn = (s0 < sfl * uc[IDN]) ? sfl * uc[IDN] : s0;
          if (sn != s0) cs[o] = sn;
          cr[o] = sn / uc[IDN];
        }
        break;
      }
      case BOX_PROLONG:
        if (op.pad == 1) prolong_cell(cr, r, ns, op.dst, k, j, i);
        else prolong_cell(cw, w, 5, op.dst, k, j, i);
        break;
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
        if (mhd) {   // bvals_refine.cpp:640-657: bcc of the prolongated field, then U(W, bcc) (adiabatic_mhd.cpp:99-135)
          const Real *bb = b + (idx_t)op.dst * 3 * g.sn + c;
          Real *bc = bcc + (idx_t)op.dst * 3 * g.sn + c;
          const Real bcc1 = 0.5 * bb[0] + 0.5 * bb[1];
          const Real bcc2 = 0.5 * bb[g.sn] + 0.5 * bb[g.sn + g.sj];
          const Real bcc3 = 0.5 * bb[2 * g.sn] + 0.5 * bb[2 * g.sn + g.sk];
          bc[0] = bcc1; bc[g.sn] = bcc2; bc[2 * g.sn] = bcc3;
          if (!eos.iso)   // (isothermal: no energy, eos/isothermal_mhd.cpp:88-112)
            ub[4 * g.sn] = w_p * igm1 + 0.5 * (w_d * (sqr(w_vx) + sqr(w_vy) + sqr(w_vz)) +
                                               (sqr(bcc1) + sqr(bcc2) + sqr(bcc3)));
        } else if (!eos.iso) {
          ub[4 * g.sn] = w_p * igm1 + 0.5 * w_d * (sqr(w_vx) + sqr(w_vy) + sqr(w_vz));
        }
        // PassiveScalarPrimitiveToConserved (bvals_refine.cpp:658-662, eos_scalars.cpp:133-152)
        for (int n = 0; n < ns; ++n) {
          const idx_t o = ((idx_t)op.dst * ns + n) * g.sn + c;
          s[o] = r[o] * w_d;
        }
        break;
      }
      case BOX_FLUX_FROM_FINER: flux_cell(op, k, j, i, dk, dj, di); break;
      case BOX_COARSE_BC: {   // bvals/cc/outflow_cc.cpp, bvals/cc/hydro/reflect_hydro.cpp with ngh = 1
        const int si = (op.dir == 0) ? op.so[0] : i, sj = (op.dir == 1) ? op.so[1] : j, sk = (op.dir == 2) ? op.so[2] : k;
        const idx_t so = cidx(cg, sk, sj, si), d = cidx(cg, k, j, i);
        if (op.pad == 1) {   // coarse_r_: bvals/cc/reflect_cc.cpp, outflow_cc.cpp (plain copies)
          for (int n = 0; n < ns; ++n) cr[((idx_t)op.dst * ns + n) * cg.csn + d] = cr[((idx_t)op.dst * ns + n) * cg.csn + so];
          break;
        }
        Real *wb = cw + (idx_t)op.dst * 5 * cg.csn;
        Real v[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) v[n] = wb[n * cg.csn + so];
        if (op.src == 1) v[IVX + op.dir] = -v[IVX + op.dir];   // reflect the normal velocity
#pragma unroll
        for (int n = 0; n < 5; ++n) wb[n * cg.csn + d] = v[n];
        break;
      }
      case BOX_FC_COPY: {
        Real *bd = b + ((idx_t)op.dst * 3 + op.dir) * g.sn;
        const Real *bs = pv[op.peer].b + ((idx_t)op.src * 3 + op.dir) * g.sn;
        bd[gidx(g, k, j, i)] = bs[gidx(g, op.so[2] + dk, op.so[1] + dj, op.so[0] + di)];
        break;
      }
      case BOX_FC_RESTRICT: restrict_face(op.dst, op.dir, k, j, i); break;
      case BOX_FC_FROM_FINER:
        b[((idx_t)op.dst * 3 + op.dir) * g.sn + gidx(g, k, j, i)] =
            pv[op.peer].cb[((idx_t)op.src * 3 + op.dir) * cg.cfn + cfidx(cg, op.so[2] + dk, op.so[1] + dj, op.so[0] + di)];
        break;
      case BOX_FC_FROM_COARSER:
        cb[((idx_t)op.dst * 3 + op.dir) * cg.cfn + cfidx(cg, k, j, i)] =
            pv[op.peer].b[((idx_t)op.src * 3 + op.dir) * g.sn + gidx(g, op.so[2] + dk, op.so[1] + dj, op.so[0] + di)];
        break;
      case BOX_FC_PROLONG_SHARED: prolong_shared_face(op.dst, op.dir, k, j, i); break;
      case BOX_FC_PROLONG_INTERNAL: prolong_internal(op.dst, k, j, i); break;
      case BOX_COARSE_BC_FC: {
        const int d = op.pad;
        const int si = (d == 0) ? op.so[0] : i, sj = (d == 1) ? op.so[1] : j, sk = (d == 2) ? op.so[2] : k;
        Real *cbd = cb + ((idx_t)op.dst * 3 + op.dir) * cg.cfn;
        const Real v = cbd[cfidx(cg, sk, sj, si)];
        cbd[cfidx(cg, k, j, i)] = (op.src == 1) ? -v : v;
        break;
      }
      case BOX_EMF_CLEAR: emf[((idx_t)op.dst * 3 + op.dir) * g.sn + gidx(g, k, j, i)] = 0.0; break;
      case BOX_EMF_FROM_FINER: emf_from_finer(op, k, j, i, dk, dj, di); break;
      case BOX_EMF_SCALE: emf[((idx_t)op.dst * 3 + op.dir) * g.sn + gidx(g, k, j, i)] *= (1.0 / (Real)op.src); break;
      default: break;
    }
  }
  // MeshRefinement::RestrictFieldX1/X2/X3, 3-D Cartesian (mesh_refinement.cpp:280-299, 338-362, 411-426):
  // area-weighted average of the four fine faces on a coarse face, summed in the reference's order
  ATHK_HD void restrict_face(int blk, int dir, int ck, int cj, int ci) const {
    const int k = (ck - cg.cks) * 2 + g.ks, j = (cj - cg.cjs) * 2 + g.js, i = (ci - cg.cis) * 2 + g.is;
    const Real *dx1 = geo.dxf[0] + blk * geo.PG, *dx2 = geo.dxf[1] + blk * geo.PG, *dx3 = geo.dxf[2] + blk * geo.PG;
    const Real *f = b + ((idx_t)blk * 3 + dir) * g.sn + gidx(g, k, j, i);
    Real a0, a1, a2, a3;
    idx_t o1, o2, o3;
    if (dir == 0) {          // (k,j), (k,j+1), (k+1,j), (k+1,j+1): Face1Area = dx2f(j)*dx3f(k)
      a0 = dx2[j] * dx3[k]; a1 = dx2[j + 1] * dx3[k]; a2 = dx2[j] * dx3[k + 1]; a3 = dx2[j + 1] * dx3[k + 1];
      o1 = g.sj; o2 = g.sk; o3 = g.sk + g.sj;
    } else if (dir == 1) {   // (k,i), (k,i+1), (k+1,i), (k+1,i+1): Face2Area = dx1f(i)*dx3f(k)
      a0 = dx1[i] * dx3[k]; a1 = dx1[i + 1] * dx3[k]; a2 = dx1[i] * dx3[k + 1]; a3 = dx1[i + 1] * dx3[k + 1];
      o1 = 1; o2 = g.sk; o3 = g.sk + 1;
    } else {                 // (j,i), (j,i+1), (j+1,i), (j+1,i+1): Face3Area = dx1f(i)*dx2f(j)
      a0 = dx1[i] * dx2[j]; a1 = dx1[i + 1] * dx2[j]; a2 = dx1[i] * dx2[j + 1]; a3 = dx1[i + 1] * dx2[j + 1];
      o1 = 1; o2 = g.sj; o3 = g.sj + 1;
    }
    const Real tarea = a0 + a1 + a2 + a3;
    cb[((idx_t)blk * 3 + dir) * cg.cfn + cfidx(cg, ck, cj, ci)] = (f[0] * a0 + f[o1] * a1 + f[o2] * a2 + f[o3] * a3) / tarea;
  }
  ATHK_HD static Real lim_grad(Real gm, Real gp) {   // 0.5*(SIGN(gm)+SIGN(gp))*min(|gm|,|gp|)
    return 0.5 * (sign_of(gm) + sign_of(gp)) * min_std(fabs(gm), fabs(gp));
  }
  // MeshRefinement::ProlongateSharedFieldX1/X2/X3, 3-D Cartesian (fluxinterp_ = false;
  // mesh_refinement.cpp:838-875, 968-1008, 1112-1153): x?s? = x?v (coordinates/cartesian.cpp)
  ATHK_HD void prolong_shared_face(int blk, int dir, int k, int j, int i) const {
    const int fk = (k - cg.cks) * 2 + g.ks, fj = (j - cg.cjs) * 2 + g.js, fi = (i - cg.cis) * 2 + g.is;
    const Real *cx1 = cxv + ((idx_t)blk * 3 + 0) * cg.CPG, *cx2 = cxv + ((idx_t)blk * 3 + 1) * cg.CPG,
               *cx3 = cxv + ((idx_t)blk * 3 + 2) * cg.CPG;
    const Real *fx1 = xv + ((idx_t)blk * 3 + 0) * geo.PG, *fx2 = xv + ((idx_t)blk * 3 + 1) * geo.PG,
               *fx3 = xv + ((idx_t)blk * 3 + 2) * geo.PG;
    const Real *c = cb + ((idx_t)blk * 3 + dir) * cg.cfn + cfidx(cg, k, j, i);
    Real *f = b + ((idx_t)blk * 3 + dir) * g.sn + gidx(g, fk, fj, fi);
    const Real ccval = c[0];
    if (dir == 0) {
      const Real x3c = cx3[k], dx3m = x3c - cx3[k - 1], dx3p = cx3[k + 1] - x3c;
      const Real fx3m = fx3[fk], fx3p = fx3[fk + 1];
      const Real x2c = cx2[j], dx2m = x2c - cx2[j - 1], dx2p = cx2[j + 1] - x2c;
      const Real fx2m = fx2[fj], fx2p = fx2[fj + 1];
      const Real gx2c = lim_grad((ccval - c[-cg.cfj]) / dx2m, (c[cg.cfj] - ccval) / dx2p);
      const Real gx3c = lim_grad((ccval - c[-cg.cfk]) / dx3m, (c[cg.cfk] - ccval) / dx3p);
      f[0] = ccval - gx2c * (x2c - fx2m) - gx3c * (x3c - fx3m);
      f[g.sj] = ccval + gx2c * (fx2p - x2c) - gx3c * (x3c - fx3m);
      f[g.sk] = ccval - gx2c * (x2c - fx2m) + gx3c * (fx3p - x3c);
      f[g.sk + g.sj] = ccval + gx2c * (fx2p - x2c) + gx3c * (fx3p - x3c);
    } else if (dir == 1) {
      const Real x3c = cx3[k], dx3m = x3c - cx3[k - 1], dx3p = cx3[k + 1] - x3c;
      const Real fx3m = fx3[fk], fx3p = fx3[fk + 1];
      const Real x1c = cx1[i], dx1m = x1c - cx1[i - 1], dx1p = cx1[i + 1] - x1c;
      const Real fx1m = fx1[fi], fx1p = fx1[fi + 1];
      const Real gx1c = lim_grad((ccval - c[-1]) / dx1m, (c[1] - ccval) / dx1p);
      const Real gx3c = lim_grad((ccval - c[-cg.cfk]) / dx3m, (c[cg.cfk] - ccval) / dx3p);
      f[0] = ccval - gx1c * (x1c - fx1m) - gx3c * (x3c - fx3m);
      f[1] = ccval + gx1c * (fx1p - x1c) - gx3c * (x3c - fx3m);
      f[g.sk] = ccval - gx1c * (x1c - fx1m) + gx3c * (fx3p - x3c);
      f[g.sk + 1] = ccval + gx1c * (fx1p - x1c) + gx3c * (fx3p - x3c);
    } else {
      const Real x2c = cx2[j], dx2m = x2c - cx2[j - 1], dx2p = cx2[j + 1] - x2c;
      const Real fx2m = fx2[fj], fx2p = fx2[fj + 1];
      const Real x1c = cx1[i], dx1m = x1c - cx1[i - 1], dx1p = cx1[i + 1] - x1c;
      const Real fx1m = fx1[fi], fx1p = fx1[fi + 1];
      const Real gx1c = lim_grad((ccval - c[-1]) / dx1m, (c[1] - ccval) / dx1p);
      const Real gx2c = lim_grad((ccval - c[-cg.cfj]) / dx2m, (c[cg.cfj] - ccval) / dx2p);
      f[0] = ccval - gx1c * (x1c - fx1m) - gx2c * (x2c - fx2m);
      f[1] = ccval + gx1c * (fx1p - x1c) - gx2c * (x2c - fx2m);
      f[g.sj] = ccval - gx1c * (x1c - fx1m) + gx2c * (fx2p - x2c);
      f[g.sj + 1] = ccval + gx1c * (fx1p - x1c) + gx2c * (fx2p - x2c);
    }
  }
  // MeshRefinement::ProlongateInternalField, 3-D (mesh_refinement.cpp:1228-1338): the divergence-
  // preserving interpolation of Toth & Roe for the twelve fine faces inside one coarse cell; reads
  // only the fine faces ON the coarse cell's surface.  Accumulation order as the reference's jj, ii loops.
  ATHK_HD void prolong_internal(int blk, int k, int j, int i) const {
    const int fk = (k - cg.cks) * 2 + g.ks, fj = (j - cg.cjs) * 2 + g.js, fi = (i - cg.cis) * 2 + g.is;
    const Real *dx1 = geo.dxf[0] + blk * geo.PG, *dx2 = geo.dxf[1] + blk * geo.PG, *dx3 = geo.dxf[2] + blk * geo.PG;
    Real *b1 = b + ((idx_t)blk * 3 + 0) * g.sn, *b2 = b + ((idx_t)blk * 3 + 1) * g.sn, *b3 = b + ((idx_t)blk * 3 + 2) * g.sn;
    auto X1 = [&](int kk, int jj, int ii) -> Real & { return b1[gidx(g, kk, jj, ii)]; };
    auto X2 = [&](int kk, int jj, int ii) -> Real & { return b2[gidx(g, kk, jj, ii)]; };
    auto X3 = [&](int kk, int jj, int ii) -> Real & { return b3[gidx(g, kk, jj, ii)]; };
    auto A1 = [&](int kk, int jj) { return dx2[jj] * dx3[kk]; };   // Face1Area(k,j,.)
    auto A2 = [&](int kk, int ii) { return dx1[ii] * dx3[kk]; };   // Face2Area(k,.,i)
    auto A3 = [&](int jj, int ii) { return dx1[ii] * dx2[jj]; };   // Face3Area(.,j,i)
    Real Uxx = 0.0, Vyy = 0.0, Wzz = 0.0, Uxyz = 0.0, Vxyz = 0.0, Wxyz = 0.0;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const int js = 2 * jj - 1, fjj = fj + jj, fjp = fj + 2 * jj;
#pragma unroll
      for (int ii = 0; ii < 2; ii++) {
        const int is = 2 * ii - 1, fii = fi + ii, fip = fi + 2 * ii;
        const Real sis = (Real)is, sjs = (Real)js, sij = (Real)(is * js);
        const Real x2lo = X2(fk, fjp, fii) * A2(fk, fii), x2hi = X2(fk + 1, fjp, fii) * A2(fk + 1, fii);
        const Real x3hi = X3(fk + 2, fjj, fii) * A3(fjj, fii), x3lo = X3(fk, fjj, fii) * A3(fjj, fii);
        const Real x1lo = X1(fk, fjj, fip) * A1(fk, fjj), x1hi = X1(fk + 1, fjj, fip) * A1(fk + 1, fjj);
        Uxx += sis * (sjs * (x2lo + x2hi) + (x3hi - x3lo));
        Vyy += sjs * ((x3hi - x3lo) + sis * (x1lo + x1hi));
        Wzz += sis * (x1hi - x1lo) + sjs * (x2hi - x2lo);
        Uxyz += sij * (x1hi - x1lo);
        Vxyz += sij * (x2hi - x2lo);
        Wxyz += sij * (x3hi - x3lo);
      }
    }
    const Real Sdx1 = sqr(dx1[fi] + dx1[fi + 1]);
    const Real Sdx2 = sqr(dx2[fj] + dx2[fj + 1]);     // GetEdge2Length = dx2f(j)
    const Real Sdx3 = sqr(dx3[fk] + dx3[fk + 1]);     // GetEdge3Length = dx3f(k)
    Uxx *= 0.125; Vyy *= 0.125; Wzz *= 0.125;
    Uxyz *= 0.125 / (Sdx2 + Sdx3);
    Vxyz *= 0.125 / (Sdx1 + Sdx3);
    Wxyz *= 0.125 / (Sdx1 + Sdx2);
    // x1 faces at fi+1
    {
      const Real a00 = A1(fk, fj), a01 = A1(fk, fj + 1), a10 = A1(fk + 1, fj), a11 = A1(fk + 1, fj + 1);
      const Real v00 = (0.5 * (X1(fk, fj, fi) * a00 + X1(fk, fj, fi + 2) * a00) + Uxx - Sdx3 * Vxyz - Sdx2 * Wxyz) / a00;
      const Real v01 = (0.5 * (X1(fk, fj + 1, fi) * a01 + X1(fk, fj + 1, fi + 2) * a01) + Uxx - Sdx3 * Vxyz + Sdx2 * Wxyz) / a01;
      const Real v10 = (0.5 * (X1(fk + 1, fj, fi) * a10 + X1(fk + 1, fj, fi + 2) * a10) + Uxx + Sdx3 * Vxyz - Sdx2 * Wxyz) / a10;
      const Real v11 = (0.5 * (X1(fk + 1, fj + 1, fi) * a11 + X1(fk + 1, fj + 1, fi + 2) * a11) + Uxx + Sdx3 * Vxyz + Sdx2 * Wxyz) / a11;
      X1(fk, fj, fi + 1) = v00; X1(fk, fj + 1, fi + 1) = v01; X1(fk + 1, fj, fi + 1) = v10; X1(fk + 1, fj + 1, fi + 1) = v11;
    }
    // x2 faces at fj+1 (sarea_x2_[.][0], [.][2], [.][1] = Face2Area at fj, fj+2, fj+1: dx1f(i)*dx3f(k), no j)
    {
      const Real a0i = A2(fk, fi), a0p = A2(fk, fi + 1), a1i = A2(fk + 1, fi), a1p = A2(fk + 1, fi + 1);
      const Real v00 = (0.5 * (X2(fk, fj, fi) * a0i + X2(fk, fj + 2, fi) * a0i) + Vyy - Sdx3 * Uxyz - Sdx1 * Wxyz) / a0i;
      const Real v01 = (0.5 * (X2(fk, fj, fi + 1) * a0p + X2(fk, fj + 2, fi + 1) * a0p) + Vyy - Sdx3 * Uxyz + Sdx1 * Wxyz) / a0p;
      const Real v10 = (0.5 * (X2(fk + 1, fj, fi) * a1i + X2(fk + 1, fj + 2, fi) * a1i) + Vyy + Sdx3 * Uxyz - Sdx1 * Wxyz) / a1i;
      const Real v11 = (0.5 * (X2(fk + 1, fj, fi + 1) * a1p + X2(fk + 1, fj + 2, fi + 1) * a1p) + Vyy + Sdx3 * Uxyz + Sdx1 * Wxyz) / a1p;
      X2(fk, fj + 1, fi) = v00; X2(fk, fj + 1, fi + 1) = v01; X2(fk + 1, fj + 1, fi) = v10; X2(fk + 1, fj + 1, fi + 1) = v11;
    }
    // x3 faces at fk+1 (note the reference's operand order: fk+2 first)
    {
      const Real a0i = A3(fj, fi), a0p = A3(fj, fi + 1), a1i = A3(fj + 1, fi), a1p = A3(fj + 1, fi + 1);
      const Real v00 = (0.5 * (X3(fk + 2, fj, fi) * a0i + X3(fk, fj, fi) * a0i) + Wzz - Sdx2 * Uxyz - Sdx1 * Vxyz) / a0i;
      const Real v01 = (0.5 * (X3(fk + 2, fj, fi + 1) * a0p + X3(fk, fj, fi + 1) * a0p) + Wzz - Sdx2 * Uxyz + Sdx1 * Vxyz) / a0p;
      const Real v10 = (0.5 * (X3(fk + 2, fj + 1, fi) * a1i + X3(fk, fj + 1, fi) * a1i) + Wzz + Sdx2 * Uxyz - Sdx1 * Vxyz) / a1i;
      const Real v11 = (0.5 * (X3(fk + 2, fj + 1, fi + 1) * a1p + X3(fk, fj + 1, fi + 1) * a1p) + Wzz + Sdx2 * Uxyz + Sdx1 * Vxyz) / a1p;
      X3(fk + 1, fj, fi) = v00; X3(fk + 1, fj, fi + 1) = v01; X3(fk + 1, fj + 1, fi) = v10; X3(fk + 1, fj + 1, fi + 1) = v11;
    }
  }
  // raw (as computed by ComputeCornerE) surface EMF of a block from its packed surface buffer
  ATHK_HD Real raw_emf(const Real *surf, int blk, int comp, int k, int j, int i) const {
    const Real *s = surf + (idx_t)blk * sf.per_block();
    const int kk = k - g.ks, jj = j - g.js, ii = i - g.is;
    const bool ib = (i == g.is || i == g.ie + 1), jb = (j == g.js || j == g.je + 1);
    if (comp == 0) {
      if (jb) return s[sf.face_off(1, j == g.js ? 0 : 1) + sf.x2_e1(kk, ii)];
      return s[sf.face_off(2, k == g.ks ? 0 : 1) + sf.x3_e1(jj, ii)];
    } else if (comp == 1) {
      if (ib) return s[sf.face_off(0, i == g.is ? 0 : 1) + sf.x1_e2(kk, jj)];
      return s[sf.face_off(2, k == g.ks ? 0 : 1) + sf.x3_e2(jj, ii)];
    }
    if (ib) return s[sf.face_off(0, i == g.is ? 0 : 1) + sf.x1_e3(kk, jj)];
    return s[sf.face_off(1, j == g.js ? 0 : 1) + sf.x2_e3(kk, ii)];
  }
  // LoadFluxBoundaryBufferToCoarser (flux_correction_fc.cpp:333-570) + SetFluxBoundaryFromFiner (:944-1165):
  // the two fine edges under a coarse edge, weighted by their lengths (Cartesian: dx?f of the fine block)
  ATHK_HD void emf_from_finer(const BoxOp &op, int k, int j, int i, int dk, int dj, int di) const {
    const int sblk = op.src, c = op.dir;
    const int fi = op.so[0] + 2 * di, fj = op.so[1] + 2 * dj, fk = op.so[2] + 2 * dk;
    const Real *dx = pv[op.peer].dxf[c] + sblk * geo.PG;
    const int n = (c == 0) ? fi : (c == 1 ? fj : fk);
    const Real el1 = dx[n], el2 = dx[n + 1];
    const Real *surf = pv[op.peer].emfbuf;
    const Real v1 = raw_emf(surf, sblk, c, fk, fj, fi);
    const Real v2 = raw_emf(surf, sblk, c, fk + (c == 2), fj + (c == 1), fi + (c == 0));
    emf[((idx_t)op.dst * 3 + c) * g.sn + gidx(g, k, j, i)] += (v1 * el1 + v2 * el2) / (el1 + el2);
  }
  // mesh_refinement.cpp:123-151 (3-D, Cartesian volumes dx1f*dx2f*dx3f, coordinates/cartesian.cpp)
  ATHK_HD void restrict_cell(const Real *fine, Real *coarse, int nv, int blk, int ck, int cj, int ci) const {
    const int k = (ck - cg.cks) * 2 + g.ks, j = (cj - cg.cjs) * 2 + g.js, i = (ci - cg.cis) * 2 + g.is;
    const Real *dx1 = geo.dxf[0] + blk * geo.PG, *dx2 = geo.dxf[1] + blk * geo.PG, *dx3 = geo.dxf[2] + blk * geo.PG;
    // fvol_[a][b](i) = CellVolume(k+a, j+b, i) = dx1f(i)*dx2f(j+b)*dx3f(k+a)
    const Real v00i = dx1[i] * dx2[j] * dx3[k], v00p = dx1[i + 1] * dx2[j] * dx3[k];
    const Real v01i = dx1[i] * dx2[j + 1] * dx3[k], v01p = dx1[i + 1] * dx2[j + 1] * dx3[k];
    const Real v10i = dx1[i] * dx2[j] * dx3[k + 1], v10p = dx1[i + 1] * dx2[j] * dx3[k + 1];
    const Real v11i = dx1[i] * dx2[j + 1] * dx3[k + 1], v11p = dx1[i + 1] * dx2[j + 1] * dx3[k + 1];
    const Real tvol = ((v00i + v01i) + (v00p + v01p)) + ((v10i + v11i) + (v10p + v11p));
    const idx_t c = gidx(g, k, j, i), cc = cidx(cg, ck, cj, ci);
    for (int n = 0; n < nv; ++n) {
      const Real *f = fine + ((idx_t)blk * nv + n) * g.sn + c;
      coarse[((idx_t)blk * nv + n) * cg.csn + cc] =
          (((f[0] * v00i + f[g.sj] * v01i) + (f[1] * v00p + f[g.sj + 1] * v01p)) +
           ((f[g.sk] * v10i + f[g.sk + g.sj] * v11i) + (f[g.sk + 1] * v10p + f[g.sk + g.sj + 1] * v11p))) / tvol;
    }
  }
  ATHK_HD static Real sign_of(Real x) { return (x < 0.0) ? -1.0 : 1.0; }   // SIGN(), defs.hpp.in:150
  // mesh_refinement.cpp:470-530 (3-D): minmod-limited gradients of the coarse primitives
  ATHK_HD void prolong_cell(const Real *cprim, Real *fprim, int nv, int blk, int k, int j, int i) const {
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
    for (int n = 0; n < nv; ++n) {
      const Real *cb = cprim + ((idx_t)blk * nv + n) * cg.csn + cc;
      Real *fb = fprim + ((idx_t)blk * nv + n) * g.sn + fc;
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
    const PeerView &sv = pv[op.peer];
    const Real *dx1 = sv.dxf[0] + sblk * geo.PG, *dx2 = sv.dxf[1] + sblk * geo.PG, *dx3 = sv.dxf[2] + sblk * geo.PG;
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
    const idx_t so = gidx(g, fk, fj, fi), d = gidx(g, k, j, i);
    const int nv = (op.pad == 1) ? ns : 5;
    Real *fl = (op.pad == 1) ? sflx[op.dir] : flx[op.dir];
    const Real *fsrc = (op.pad == 1) ? sv.sflx[op.dir] : sv.flx[op.dir];
    // (source and destination are faces of different blocks: no element is both read and written)
    for (int n = 0; n < nv; ++n) {
      const Real *f = fsrc + ((idx_t)sblk * nv + n) * g.sn + so;
      fl[((idx_t)op.dst * nv + n) * g.sn + d] = (f[0] * a0 + f[o1] * a1 + f[o2] * a2 + f[o3] * a3) / tarea;
    }
  }
};

}  // namespace athk
#endif  // ATHK_REFINE_H_
