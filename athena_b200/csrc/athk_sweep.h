// athk_sweep.h -- fused directional sweeps of Hydro::CalculateFluxes + AddFluxDivergence +
// the integrator's WeightedAve (the "stage" fast path behind athk_stage()).
//
// Same arithmetic as the per-operator kernels (athk_ops.h) -- bit-identical results -- but laid
// out for the B200: FP64 issue is the binding resource of HLLD+PLM, HBM the second, so
//   * every cell's PLM slopes are computed once per direction (x2/x3 sweeps march along the
//     sweep direction with the previous cell's left state held in registers; the CPU code's
//     wl_/wlb_ pencil swap, calculate_fluxes.cpp:236,355, becomes a register rotation);
//   * each primitive is loaded once per sweep as a coalesced row (threads along i);
//   * fluxes never go to HBM: each sweep folds (A F)_+ - (A F)_- into a 5-component scratch
//     "dflx" (add_flux_divergence.cpp:57-83 accumulation order x1, x2, x3), and the x3 sweep
//     applies  u_out = WeightedAve(...) - wght*dflx/vol  (time_integrator.cpp:1797-1824).
// Only the face EMFs + CT weights (needed one cell outside the block by ComputeCornerE) are
// written out.
#ifndef ATHK_SWEEP_H_
#define ATHK_SWEEP_H_

#include "athk_ops.h"

namespace athk {

// how the x3 sweep finishes the hydro update for an interior cell
struct HydroIntegrate {
  // registers after the host-side pointer bookkeeping: see athk_stage() in athk_api.cu
  const Real *u;      // "u"  on entry  [nb][5][sn]
  const Real *u1;     // "u1" on entry (may be null when u1_zero)
  Real *u_out;        // array that holds "u" after the stage
  Real *u1_out;       // array that holds "u1" after the stage (written only if delta != 0 and no swap)
  Real delta, g1, g2; // stage weights (gamma_3 = 0 for vl2/rk2/rk3)
  Real wght;          // beta*dt
  int swap;           // gamma = (0,1,0): SwapAthenaArray instead of averaging
  int u1_zero;        // stage 1: u1 was ZeroClear'ed (value known, no need to read it)
  int wa_mode;        // WeightedAvePoint mode for the (g1,g2) average when !swap
};

struct SweepArgs {
  Grid g;
  EosParams eos;
  Geom geo;
  const Real *w, *bcc, *b;   // primitives, cell-centred field, face field of the stage input
  Real *flx[3];              // optional flux arrays (null = do not store)
  Real *ef;                  // [nb][6][sn]
  Real *wct;                 // [nb][3][sn]
  Real *d2, *d3;             // [nb][5][sn] scratch: x2 / x3 terms of the flux divergence
  Real dt;                   // pmy_mesh->dt
  Real *mf[3];               // passive scalars only: where to leave the mass flux of each face
  idx_t mf_bstride;          //   (block stride; null = no scalars)  see athk_scalars.h
  HydroIntegrate hi;
};

// rotated primitive vector of one cell: (d, v_n, v_t1, v_t2, p, b_t1, b_t2)
template <int DIR, bool MHD>
ATHK_HD void load_q(const SweepArgs &a, int blk, idx_t c, Real q[7]) {
  const Grid &g = a.g;
  const Real *wb = a.w + (idx_t)blk * 5 * g.sn + c;
  q[0] = wb[0];
  q[1] = wb[(idx_t)(IVX + DIR) * g.sn];
  q[2] = wb[(idx_t)(IVX + (DIR + 1) % 3) * g.sn];
  q[3] = wb[(idx_t)(IVX + (DIR + 2) % 3) * g.sn];
  q[4] = wb[(idx_t)IPR * g.sn];
  if (MHD) {
    const Real *bc = a.bcc + (idx_t)blk * 3 * g.sn + c;
    q[5] = bc[(idx_t)((DIR + 1) % 3) * g.sn];
    q[6] = bc[(idx_t)((DIR + 2) % 3) * g.sn];
  }
}

// (dxw = dxf[DIR] of the face's upper cell, loaded by the caller: the marching kernels fetch it
//  before they wait for their asynchronous copies)
template <int DIR, bool MHD, int RS, class M>
ATHK_HD void solve_face_dx(M &m, const SweepArgs &a, int blk, idx_t c, Real dxw, const Real wl[7],
                           const Real wr[7], const Real bxi, Real f[7], bool store) {
  const Grid &g = a.g;
  if (MHD) {
    if (RS == RS_HLLD_ISO) mhd_riemann_m<RS>(m, eos_of_block(a.eos, blk), wl, wr, bxi, f);   // (uses the block's density floor)
    else mhd_riemann_m<RS>(m, a.eos, wl, wr, bxi, f);
    if (store) {
      Real *eb = a.ef + ((idx_t)blk * 6 + 2 * DIR) * g.sn + c;
      eb[0] = -f[IBY];
      eb[g.sn] = f[IBZ];
      a.wct[((idx_t)blk * 3 + DIR) * g.sn + c] = ct_weight_m(m, f[IDN], wl[IDN], wr[IDN], dxw, a.dt);
    }
  } else {
    hydro_riemann_m<RS>(m, a.eos, wl, wr, f);
  }
  if (store && a.mf[DIR]) a.mf[DIR][(idx_t)blk * a.mf_bstride + c] = f[IDN];
  if (store && a.flx[DIR]) {
    Real *fb = a.flx[DIR] + (idx_t)blk * 5 * g.sn + c;
    fb[0] = f[IDN];
    fb[(idx_t)(IVX + DIR) * g.sn] = f[IVX];
    fb[(idx_t)(IVX + (DIR + 1) % 3) * g.sn] = f[IVY];
    fb[(idx_t)(IVX + (DIR + 2) % 3) * g.sn] = f[IVZ];
    fb[(idx_t)IEN * g.sn] = f[IEN];
  }
}
template <int DIR, bool MHD, int RS, class M>
ATHK_HD void solve_face_m(M &m, const SweepArgs &a, int blk, idx_t c, int n_dir, const Real wl[7],
                          const Real wr[7], const Real bxi, Real f[7], bool store) {
  const Real dxw = (MHD && store) ? a.geo.dxf[DIR][blk * a.geo.PG + n_dir] : 0.0;
  solve_face_dx<DIR, MHD, RS>(m, a, blk, c, dxw, wl, wr, bxi, f, store);
}
template <int DIR, bool MHD, int RS>
ATHK_HD void solve_face(const SweepArgs &a, int blk, idx_t c, int n_dir, const Real wl[7],
                        const Real wr[7], const Real bxi, Real f[7], bool store) {
  with_math<RS>([&](auto &m) { solve_face_m<DIR, MHD, RS>(m, a, blk, c, n_dir, wl, wr, bxi, f, store); });
}

// PLM (or donor cell) of one cell from its two neighbours: left state of the upper face and
// right state of the lower face (plm.cpp:104-112,149-154 / dc.cpp)
template <int ORDER, bool MHD, class M>
ATHK_HD void recon_cell_m(M &m, const Real qm[7], const Real q0[7], const Real qp[7], Real cp, Real cm,
                          Real wl_up[7], Real wr_lo[7]) {
  const int NW = MHD ? 7 : 5;
#pragma unroll
  for (int n = 0; n < NW; ++n) {
    if (ORDER == 1) {
      wl_up[n] = q0[n];
      wr_lo[n] = q0[n];
    } else {
      const Real dwm = vl_slope_m(m, q0[n] - qm[n], qp[n] - q0[n]);
      wl_up[n] = q0[n] + cp * dwm;
      wr_lo[n] = q0[n] - cm * dwm;
    }
  }
}
template <int ORDER, bool MHD>
ATHK_HD void recon_cell(const Real qm[7], const Real q0[7], const Real qp[7], Real cp, Real cm,
                        Real wl_up[7], Real wr_lo[7]) {
  MathIEEE m;
  recon_cell_m<ORDER, MHD>(m, qm, q0, qp, cp, cm, wl_up, wr_lo);
}

// WeightedAve + AddFluxDivergence of ONE variable of one active cell (time_integrator.cpp:
// 1806-1821 + add_flux_divergence.cpp:86-90; the scalar twin :2386-2403): o = its offset in the
// registers, dsum = its summed flux divergence.
ATHK_HD void integrate_var(const HydroIntegrate &h, idx_t o, Real dsum, Real vol) {
  // load only what the stage's weights use (vl2 stage 2 never reads u, stage 1 never reads u1)
  const Real un = (h.delta != 0.0 || !h.swap) ? h.u[o] : 0.0;
  Real u1n = h.u1_zero ? 0.0 : h.u1[o];
  if (h.delta != 0.0) u1n += h.delta * un;          // WeightedAve(u1,u,..,{1,delta,0,0,0})
  Real r;
  if (h.swap) {
    r = u1n;                                        // u <-> u1
  } else {
    switch (h.wa_mode) {                            // WeightedAve(u,u1,..,{g1,g2,0,0,0})
      case 1: r = un; r += h.g2 * u1n; break;
      case 3: r = h.g2 * u1n; break;
      case 5: r = h.g1 * un + h.g2 * u1n; break;
      case 6: r = un; r *= h.g1; break;
      default: r = un; break;
    }
    if (h.delta != 0.0) h.u1_out[o] = u1n;
  }
  r -= fdiv(h.wght * dsum, vol);                    // AddFluxDivergence (0 / vol == 0 * vol)
  h.u_out[o] = r;
}

// the same with the register values already loaded (un = h.u[o] where the stage reads it, u1n =
// h.u1[o] likewise) and the division issued through a math policy
// (rvol = m.div_rcp(vol): the five variables of a cell divide by the same volume)
template <class M>
ATHK_HD void integrate_var_m(M &m, const HydroIntegrate &h, idx_t o, Real dsum, Real vol, Real rvol, Real un, Real u1n) {
  if (h.delta != 0.0) u1n += h.delta * un;
  Real r;
  if (h.swap) {
    r = u1n;
  } else {
    switch (h.wa_mode) {
      case 1: r = un; r += h.g2 * u1n; break;
      case 3: r = h.g2 * u1n; break;
      case 5: r = h.g1 * un + h.g2 * u1n; break;
      case 6: r = un; r *= h.g1; break;
      default: r = un; break;
    }
    if (h.delta != 0.0) h.u1_out[o] = u1n;
  }
  r -= m.fdiv_by(h.wght * dsum, vol, rvol);   // (static regions: the divergence is exactly 0)
  h.u_out[o] = r;
}

// finish the hydro update of one interior cell (all 5 variables)
ATHK_HD void integrate_cell(const SweepArgs &a, int blk, idx_t c, const Real dsum[5], Real vol) {
  const Grid &g = a.g;
#pragma unroll
  for (int n = 0; n < 5; ++n) integrate_var(a.hi, ((idx_t)blk * 5 + n) * g.sn + c, dsum[n], vol);
}

//========================================================================================
// One face of a DIR-sweep, self-contained: reconstruct both adjacent cells (donor cell, PLM or
// PPM), then Riemann solve.  c = index of the cell above the face, n_dir = its index along the
// sweep.
// Where a face's stencil comes from.  cell(off, q): rotated primitives of the cell `off` cells
// above the face's upper cell (off = -ORDER .. ORDER-1); bx(): the face-normal field.
// GlobalStencil reads the arrays in HBM/L2; the tile kernels (athk_sweep_kernels.cuh) read rows
// they staged into shared memory.
template <int DIR, bool MHD>
struct GlobalStencil {
  const SweepArgs &a;
  int blk;
  idx_t c, s;
  ATHK_HD void cell(int off, Real q[7]) const { load_q<DIR, MHD>(a, blk, c + off * s, q); }
  ATHK_HD Real bx() const { return MHD ? a.b[((idx_t)blk * 3 + DIR) * a.g.sn + c] : 0.0; }
};

template <int DIR, int ORDER, bool MHD, int RS, class ST, class M>
ATHK_HD void face_compute_from_m(M &m, const SweepArgs &a, int blk, idx_t c, int n_dir, const ST &st, Real f[7],
                                 bool store) {
  const int NW = MHD ? 7 : 5;
  Real wl[7], wr[7];
  if (ORDER == 3) {
    // PPM: L state from cell n-1 (stencil n-3..n+1), R state from cell n (stencil n-2..n+2)
    Real qm3[7], qm2[7], qm1[7], q0[7], qp1[7], qp2[7];
    st.cell(-3, qm3);
    st.cell(-2, qm2);
    st.cell(-1, qm1);
    st.cell(0, q0);
    st.cell(1, qp1);
    st.cell(2, qp2);
#pragma unroll
    for (int v = 0; v < NW; ++v) {
      Real up, lo;
      ppm_cell(qm3[v], qm2[v], qm1[v], q0[v], qp1[v], &up, &lo);
      wl[v] = up;
      ppm_cell(qm2[v], qm1[v], q0[v], qp1[v], qp2[v], &up, &lo);
      wr[v] = lo;
    }
    // ApplyPrimitiveFloors on both states (ppm.cpp:326-332)
    const Real dfl = eos_dfloor(a.eos, blk), pfl = eos_pfloor(a.eos, blk);
    wl[IDN] = (wl[IDN] > dfl) ? wl[IDN] : dfl;
    wl[IPR] = (wl[IPR] > pfl) ? wl[IPR] : pfl;
    wr[IDN] = (wr[IDN] > dfl) ? wr[IDN] : dfl;
    wr[IPR] = (wr[IPR] > pfl) ? wr[IPR] : pfl;
  } else if (ORDER == 2) {
    Real q0[7], q1[7], q2[7], q3[7];
    st.cell(-1, q1);
    st.cell(0, q2);
    st.cell(-2, q0);
    st.cell(1, q3);
    const Real *cpv = a.geo.cp[DIR] + blk * a.geo.PG;
    const Real *cmv = a.geo.cm[DIR] + blk * a.geo.PG;
    Real t0[7], t1[7];
    recon_cell_m<ORDER, MHD>(m, q0, q1, q2, cpv[n_dir - 1], cmv[n_dir - 1], wl, t0);   // cell below
    recon_cell_m<ORDER, MHD>(m, q1, q2, q3, cpv[n_dir], cmv[n_dir], t1, wr);           // cell above
  } else {
    st.cell(-1, wl);
    st.cell(0, wr);
  }
  solve_face_m<DIR, MHD, RS>(m, a, blk, c, n_dir, wl, wr, st.bx(), f, store);
}
template <int DIR, int ORDER, bool MHD, int RS, class ST>
ATHK_HD void face_compute_from(const SweepArgs &a, int blk, idx_t c, int n_dir, const ST &st, Real f[7],
                               bool store) {
  with_math<RS>([&](auto &m) { face_compute_from_m<DIR, ORDER, MHD, RS>(m, a, blk, c, n_dir, st, f, store); });
}

template <int DIR, int ORDER, bool MHD, int RS>
ATHK_HD void face_compute(const SweepArgs &a, int blk, idx_t c, int n_dir, Real f[7], bool store) {
  const Grid &g = a.g;
  const GlobalStencil<DIR, MHD> st = {a, blk, c, (DIR == 0) ? (idx_t)1 : (DIR == 1 ? g.sj : g.sk)};
  face_compute_from<DIR, ORDER, MHD, RS>(a, blk, c, n_dir, st, f, store);
}

// Hydro::CalculateFluxes as an operator: one thread per face, fluxes stored (a.flx[DIR] set)
template <int DIR, int ORDER, bool MHD, int RS = default_rsolver(MHD)>
struct FacePoint {
  SweepArgs a;
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    Real f[7];
    face_compute<DIR, ORDER, MHD, RS>(a, blk, gidx(a.g, k, j, i), (DIR == 0) ? i : (DIR == 1 ? j : k), f, true);
  }
};

//========================================================================================
// x2 / x3 sweep as a per-column loop: the host harness (tests/emul) form of the CTA-marching
// tile kernels in athk_sweep_kernels.cuh -- same faces, same stores.
//   DIR=1: column (i,k), marches j ; DIR=2: column (i,j), marches k.
// Column extents follow calculate_fluxes.cpp:176-182,299-302; faces s..e+1 along the sweep.
template <int DIR, int ORDER, bool MHD, int RS = default_rsolver(MHD)>
struct SweepMarch {
  SweepArgs a;
  // (blk, k, j, i): for DIR=1 the j argument is ignored, for DIR=2 the k argument is ignored
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const Grid &g = a.g;
    const idx_t s = (DIR == 1) ? g.sj : g.sk;
    const int n_s = (DIR == 1) ? g.js : g.ks;        // first active index along the sweep
    const int n_e = (DIR == 1) ? g.je : g.ke;
    // is this column part of the active zone (gets a flux divergence)?
    const bool interior = (i >= g.is && i <= g.ie) &&
                          ((DIR == 1) ? (k >= g.ks && k <= g.ke) : (j >= g.js && j <= g.je));
    Real area = 0.0;
    if (interior) {
      const Real dx1 = a.geo.dxf[0][blk * a.geo.PG + i];
      if (DIR == 1) area = dx1 * a.geo.dxf[2][blk * a.geo.PG + k];   // Face2Area
      else area = dx1 * a.geo.dxf[1][blk * a.geo.PG + j];            // Face3Area
    }
    Real f[7], fprev[5];
    for (int n = n_s; n <= n_e + 1; ++n) {
      const idx_t c = (DIR == 1) ? gidx(g, k, n, i) : gidx(g, n, j, i);
      face_compute<DIR, ORDER, MHD, RS>(a, blk, c, n, f, true);
      if (interior && n > n_s) {
        // cell n-1: (A F)(n) - (A F)(n-1)
        const idx_t cc = c - s;
        const int vmap[5] = {IDN, IVX + DIR, IVX + (DIR + 1) % 3, IVX + (DIR + 2) % 3, IEN};
        const Real fcur[5] = {f[IDN], f[IVX], f[IVY], f[IVZ], f[IEN]};
        Real *dst = (DIR == 1) ? a.d2 : a.d3;
#pragma unroll
        for (int v = 0; v < 5; ++v)
          dst[((idx_t)blk * 5 + vmap[v]) * g.sn + cc] = (area * fcur[v] - area * fprev[v]);
      }
      fprev[0] = f[IDN]; fprev[1] = f[IVX]; fprev[2] = f[IVY]; fprev[3] = f[IVZ]; fprev[4] = f[IEN];
    }
  }
};

//========================================================================================
// x1 sweep: one thread per face (coalesced along i).  compute() is the face solve; the GPU
// kernel passes the flux of face i+1 to the owner of cell i with a warp shuffle, the host
// harness (tests/emul) simply solves face i+1 again.
template <int ORDER, bool MHD, int RS = default_rsolver(MHD)>
struct SweepX1 {
  SweepArgs a;
  ATHK_HD void compute(int blk, int k, int j, int i, Real f[7], bool store) const {
    face_compute<0, ORDER, MHD, RS>(a, blk, gidx(a.g, k, j, i), i, f, store);
  }
  // cell (k,j,i) of the active zone.  The x1 sweep runs LAST: it forms
  //   dflx = (A1 F(i+1) - A1 F(i)); dflx += d2; dflx += d3     (add_flux_divergence.cpp:57-83)
  // from the x2/x3 sweeps' scratch terms and finishes the update (integrate_cell).
  ATHK_HD void store_d1(int blk, int k, int j, int i, const Real f[7], const Real fnext[7]) const {
    const Grid &g = a.g;
    const idx_t c = gidx(g, k, j, i);
    const Real dx1 = a.geo.dxf[0][blk * a.geo.PG + i], dx2 = a.geo.dxf[1][blk * a.geo.PG + j],
               dx3 = a.geo.dxf[2][blk * a.geo.PG + k];
    const Real area = dx2 * dx3;   // Face1Area
    const int src[5] = {IDN, IVX, IVY, IVZ, IEN};
    Real dsum[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {   // DIR=0: rotated index == variable index
      const idx_t o = ((idx_t)blk * 5 + src[v]) * g.sn + c;
      Real d = (area * fnext[src[v]] - area * f[src[v]]);
      d += a.d2[o];
      d += a.d3[o];
      dsum[src[v]] = d;
    }
    integrate_cell(a, blk, c, dsum, dx1 * dx2 * dx3);
  }
  // L2 prefetch of what store_d1 will read (d2, d3 and the register it averages), issued before
  // the Riemann solve: no registers held, and the final loads hit L2 instead of HBM.
  ATHK_HD void prefetch_cell(int blk, int k, int j, int i) const {
#if defined(__CUDA_ARCH__)
    const Grid &g = a.g;
    const idx_t c = gidx(g, k, j, i);
    const HydroIntegrate &h = a.hi;
    const Real *ureg = (h.delta != 0.0 || !h.swap) ? h.u : h.u1;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const idx_t o = ((idx_t)blk * 5 + v) * g.sn + c;
#ifndef ATHK_X1_PF
#define ATHK_X1_PF "prefetch.global.L2 [%0];"
#endif
      asm volatile(ATHK_X1_PF ::"l"(a.d2 + o));
      asm volatile(ATHK_X1_PF ::"l"(a.d3 + o));
      asm volatile(ATHK_X1_PF ::"l"(ureg + o));
    }
#endif
  }
  ATHK_HD bool cell_active(int k, int j, int i) const {
    const Grid &g = a.g;
    return i <= g.ie && j >= g.js && j <= g.je && k >= g.ks && k <= g.ke;
  }
  // host harness path
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    Real f[7], fn[7];
    compute(blk, k, j, i, f, true);
    if (cell_active(k, j, i)) {
      compute(blk, k, j, i + 1, fn, false);
      store_d1(blk, k, j, i, f, fn);
    }
  }
};

//========================================================================================
// Field::CT fused with the integrator's face-centred WeightedAve (time_integrator.cpp:
// 1853-1885, field/ct.cpp:31-116, mesh/weighted_ave.cpp:366-520).  Same launch domain as CTPoint.
struct FieldIntegrate {
  const Real *b, *b1;
  Real *b_out, *b1_out;
  Real delta, g1, g2, wght;
  int swap, b1_zero, wa_mode;
};

struct CTIntegratePoint {
  Grid g;
  Geom geo;
  const Real *e;
  FieldIntegrate fi;
  ATHK_HD Real averaged(idx_t o) const {
    const Real bn = fi.b[o];
    Real b1n = fi.b1_zero ? 0.0 : fi.b1[o];
    if (fi.delta != 0.0) b1n += fi.delta * bn;
    Real r;
    if (fi.swap) {
      r = b1n;
    } else {
      switch (fi.wa_mode) {
        case 1: r = bn; r += fi.g2 * b1n; break;
        case 3: r = fi.g2 * b1n; break;
        case 5: r = fi.g1 * bn + fi.g2 * b1n; break;
        case 6: r = bn; r *= fi.g1; break;
        default: r = bn; break;
      }
      if (fi.delta != 0.0) fi.b1_out[o] = b1n;
    }
    return r;
  }
  // the two register values averaged() reads, and the average from them (same expressions)
  ATHK_HD void ld_regs(idx_t o, Real *bn, Real *b1n) const {
    *bn = (fi.delta != 0.0 || !fi.swap) ? fi.b[o] : 0.0;
    *b1n = fi.b1_zero ? 0.0 : fi.b1[o];
  }
  ATHK_HD Real averaged_from(idx_t o, Real bn, Real b1n) const {
    if (fi.delta != 0.0) b1n += fi.delta * bn;
    Real r;
    if (fi.swap) {
      r = b1n;
    } else {
      switch (fi.wa_mode) {
        case 1: r = bn; r += fi.g2 * b1n; break;
        case 3: r = fi.g2 * b1n; break;
        case 5: r = fi.g1 * bn + fi.g2 * b1n; break;
        case 6: r = bn; r *= fi.g1; break;
        default: r = bn; break;
      }
      if (fi.delta != 0.0) fi.b1_out[o] = b1n;
    }
    return r;
  }
  // k-marching form (launch_points_kcol): faces (k0..k1, j, i); e1, e2 of plane k+1 become those of
  // plane k at the next step, and every load of a plane is issued before its stores
  ATHK_HD void col(int blk, int k0, int k1, int j, int i) const {
    idx_t c = gidx(g, k0, j, i);
    const Real *e1 = e + (idx_t)blk * 3 * g.sn, *e2 = e1 + g.sn, *e3 = e1 + 2 * g.sn;
    const Real dx1 = geo.dxf[0][blk * geo.PG + i], dx2 = geo.dxf[1][blk * geo.PG + j];
    const Real wght = fi.wght;
    const bool jok = j <= g.je, iok = i <= g.ie;
    Real e1c = e1[c], e2c = e2[c];
    for (int k = k0; k <= k1; ++k, c += g.sk) {
      const idx_t o = (idx_t)blk * 3 * g.sn + c;
      const bool kok = k <= g.ke;
      const bool f1 = kok && jok, f2 = kok && iok, f3 = jok && iok;
      const Real dx3 = geo.dxf[2][blk * geo.PG + k];
      // (all inside the padded box: k+1 <= nc3, j+1 <= nc2, i+1 <= nc1 for every launched point)
      const Real e3c = e3[c], e3j = e3[c + g.sj], e3i = e3[c + 1];
      const Real e1k = e1[c + g.sk], e1j = e1[c + g.sj], e2k = e2[c + g.sk], e2i = e2[c + 1];
      Real bn1 = 0.0, b1n1 = 0.0, bn2 = 0.0, b1n2 = 0.0, bn3 = 0.0, b1n3 = 0.0;
      if (f1) ld_regs(o, &bn1, &b1n1);
      if (f2) ld_regs(o + g.sn, &bn2, &b1n2);
      if (f3) ld_regs(o + 2 * g.sn, &bn3, &b1n3);
      if (f1) {
        const Real area = dx2 * dx3;
        Real v = averaged_from(o, bn1, b1n1);
        v -= (wght / area) * (dx3 * e3j - dx3 * e3c);
        v += (wght / area) * (dx2 * e2k - dx2 * e2c);
        fi.b_out[o] = v;
      }
      if (f2) {
        const Real area = dx1 * dx3;
        Real v = averaged_from(o + g.sn, bn2, b1n2);
        v += (wght / area) * (dx3 * e3i - dx3 * e3c);
        v -= (wght / area) * (dx1 * e1k - dx1 * e1c);
        fi.b_out[o + g.sn] = v;
      }
      if (f3) {
        const Real area = dx1 * dx2;
        Real v = averaged_from(o + 2 * g.sn, bn3, b1n3);
        v -= (wght / area) * (dx2 * e2i - dx2 * e2c);
        v += (wght / area) * (dx1 * e1j - dx1 * e1c);
        fi.b_out[o + 2 * g.sn] = v;
      }
      e1c = e1k;
      e2c = e2k;
    }
  }
  ATHK_HD void operator()(int blk, int k, int j, int i) const {
    const idx_t c = gidx(g, k, j, i);
    const Real *e1 = e + (idx_t)blk * 3 * g.sn, *e2 = e1 + g.sn, *e3 = e1 + 2 * g.sn;
    const idx_t o = (idx_t)blk * 3 * g.sn + c;
    const Real dx1 = geo.dxf[0][blk * geo.PG + i], dx2 = geo.dxf[1][blk * geo.PG + j],
               dx3 = geo.dxf[2][blk * geo.PG + k];
    const Real wght = fi.wght;
    if (k <= g.ke && j <= g.je) {
      const Real area = dx2 * dx3;
      Real v = averaged(o);
      v -= (wght / area) * (dx3 * e3[c + g.sj] - dx3 * e3[c]);
      v += (wght / area) * (dx2 * e2[c + g.sk] - dx2 * e2[c]);
      fi.b_out[o] = v;
    }
    if (k <= g.ke && i <= g.ie) {
      const Real area = dx1 * dx3;
      Real v = averaged(o + g.sn);
      v += (wght / area) * (dx3 * e3[c + 1] - dx3 * e3[c]);
      v -= (wght / area) * (dx1 * e1[c + g.sk] - dx1 * e1[c]);
      fi.b_out[o + g.sn] = v;
    }
    if (j <= g.je && i <= g.ie) {
      const Real area = dx1 * dx2;
      Real v = averaged(o + 2 * g.sn);
      v -= (wght / area) * (dx2 * e2[c + 1] - dx2 * e2[c]);
      v += (wght / area) * (dx1 * e1[c + g.sj] - dx1 * e1[c]);
      fi.b_out[o + 2 * g.sn] = v;
    }
  }
};

}  // namespace athk
#endif  // ATHK_SWEEP_H_
