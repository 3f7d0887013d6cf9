// athk_sweep_kernels.cuh -- CUDA-only launch shapes of the fused x2/x3 sweeps (athk_sweep.h holds
// the arithmetic, shared with the host harness).
//
// HLLD in FP64 needs ~100-150 registers, so resident warps -- not HBM -- bound these kernels
// (ncu, profiles/): the per-thread "march" of SweepMarch carries its rolling stencil in
// registers and drops to 8 warps/SM.  Here the CTA marches instead of the thread:
//   CTA = 32 columns (lanes, consecutive i => coalesced rows) x R faces along the sweep.
//   Each step every thread solves ONE face (no loop-carried registers), parks its 5 hydro fluxes
//   in shared memory, and after one __syncthreads the thread owning cell n-1 stores
//   (A F)(n) - (A F)(n-1) from its own flux and the row below (the previous step's top row for
//   r == 0: triple-buffered, so one barrier per step suffices and no face is solved twice)
//   into the sweep's scratch term d2 / d3.  The x1 sweep (warp shuffles, no barrier) runs last,
//   sums d1 + d2 + d3 in the reference's order and applies the integrator update.
//
// Stencil staging (ATHK_TILE_STAGE=1, default): with ~5 warps per scheduler a face solve that
// starts with 28 dependent global loads spends 40 % of its cycles on the long scoreboard (ncu
// source page, profiles/).  Instead the CTA keeps a ring of cell ROWS in shared memory -- the
// pencil the sweep marches along, with its ghost halo: every step each thread cp.async's the one
// cell it will first need at the NEXT step (8 doubles: w[5], the two transverse bcc, the face
// field), the copies land during the Riemann solves, and stencils are read from shared memory.
// Each cell is fetched from L2/HBM once per CTA instead of 2*ORDER times.
#ifndef ATHK_SWEEP_KERNELS_CUH_
#define ATHK_SWEEP_KERNELS_CUH_

#include "athk_sweep.h"

namespace athk {

#ifndef ATHK_TILE_R
#define ATHK_TILE_R 5
#endif
#ifndef ATHK_TILE_MINB
#define ATHK_TILE_MINB 4
#endif
#ifndef ATHK_TILE_PREFETCH
#define ATHK_TILE_PREFETCH 1   // (unstaged kernel) 1: prefetch.global.L2 of the next step's cell, 2: .L1, 0: off
#endif
#ifndef ATHK_TILE_STAGE
#define ATHK_TILE_STAGE 1
#endif

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows of the staged pencil: [slot][variable][lane]
template <bool MHD>
struct SmemStencil {
  static constexpr int NQ = MHD ? 8 : 5;
  const Real *ring;   // + lane
  int slot0;          // ring slot of the face's upper cell
  int nr;             // ring size
  __device__ __forceinline__ const Real *row(int off) const {
    int sl = slot0 + off;
    sl = (sl < 0) ? sl + nr : (sl >= nr ? sl - nr : sl);
    return ring + sl * (NQ * 32);
  }
  __device__ __forceinline__ void cell(int off, Real q[7]) const {
    const Real *p = row(off);
#pragma unroll
    for (int v = 0; v < (MHD ? 7 : 5); ++v) q[v] = p[v * 32];
  }
  __device__ __forceinline__ Real bx() const { return MHD ? row(0)[7 * 32] : 0.0; }
};

// PPM of every variable of the cell `off0` cells above the stencil's reference cell: left state of
// its upper face (up) and right state of its lower face (lo), with the primitive floors the
// reference applies to both face states (ppm.cpp:326-332)
// One reconstruction per cell (instead of two per face).  PPM: always.  PLM: measured on the MHD
// blast at 256^3 -- the x1 sweep gains 3 % (shuffles are free), the x2/x3 tiles LOSE 9 % (the
// hand-over costs a second barrier per step and the slopes of uniform regions are free anyway),
// so the tiles keep the face-centric PLM.
#ifndef ATHK_PLM_SHARE_TILES
#define ATHK_PLM_SHARE_TILES 0
#endif
#ifndef ATHK_PLM_SHARE_X1
#define ATHK_PLM_SHARE_X1 1
#endif
#define ATHK_SHARED_RECON(ORDER) ((ORDER) == 3 || ((ORDER) == 2 && ATHK_PLM_SHARE_TILES))
#ifndef ATHK_DONOR_CELLS_X1
#define ATHK_DONOR_CELLS_X1 0   // donor cell through the cell-owner x1 kernel: measured slower (1.56 vs 1.42 ms)
#endif
#define ATHK_SHARED_RECON_X1(ORDER) \
  ((ORDER) == 3 || ((ORDER) == 2 && ATHK_PLM_SHARE_X1) || ((ORDER) == 1 && ATHK_DONOR_CELLS_X1))

// PLM of every variable of the cell `off0` cells above the stencil's reference cell (index n_cell
// along DIR): the two states recon_cell gives
template <int DIR, bool MHD, class ST>
__device__ __forceinline__ void plm_recon_cell(const SweepArgs &a, int blk, const ST &st, int off0, int n_cell,
                                               Real up[7], Real lo[7]) {
  Real qm[7], q0[7], qp[7];
  st.cell(off0 - 1, qm);
  st.cell(off0, q0);
  st.cell(off0 + 1, qp);
  const Real cp = a.geo.cp[DIR][blk * a.geo.PG + n_cell], cm = a.geo.cm[DIR][blk * a.geo.PG + n_cell];
  with_math<MHD ? RS_HLLD : RS_HLLC>([&](auto &m) { recon_cell_m<2, MHD>(m, qm, q0, qp, cp, cm, up, lo); });
}

template <bool MHD, class ST>
__device__ __forceinline__ void ppm_recon_cell(const EosParams &eos, int blk, const ST &st, int off0,
                                               Real up[7], Real lo[7]) {
  Real qm2[7], qm1[7], q0[7], qp1[7], qp2[7];
  st.cell(off0 - 2, qm2);
  st.cell(off0 - 1, qm1);
  st.cell(off0, q0);
  st.cell(off0 + 1, qp1);
  st.cell(off0 + 2, qp2);
#pragma unroll
  for (int v = 0; v < (MHD ? 7 : 5); ++v) ppm_cell(qm2[v], qm1[v], q0[v], qp1[v], qp2[v], &up[v], &lo[v]);
  const Real dfl = eos_dfloor(eos, blk), pfl = eos_pfloor(eos, blk);
  up[IDN] = (up[IDN] > dfl) ? up[IDN] : dfl;
  up[IPR] = (up[IPR] > pfl) ? up[IPR] : pfl;
  lo[IDN] = (lo[IDN] > dfl) ? lo[IDN] : dfl;
  lo[IPR] = (lo[IPR] > pfl) ? lo[IPR] : pfl;
}

template <int DIR, int ORDER, bool MHD, int RS>
__global__ void __launch_bounds__(32 * ATHK_TILE_R, ATHK_TILE_MINB)
k_sweep_tile_staged(const SweepArgs a, const int nb, const int o_lo, const int n_o, const int il,
                    const int ni) {
  constexpr int R = ATHK_TILE_R;
  constexpr int LO = ORDER, HI = ORDER - 1;   // cells n-LO .. n+HI around face n
  constexpr int NR = 2 * R + LO + HI;         // rows in use by a step + rows in flight
  constexpr int NQ = MHD ? 8 : 5;
  // fluxes of the step: double-buffered (a fast warp writes step t+1 while a slow one still
  // reads step t); the top row also feeds r == 0 of the NEXT step, hence its own 3-deep ring
  extern __shared__ __align__(16) Real smem_tile[];
  Real(*sF)[R][5][32] = reinterpret_cast<Real(*)[R][5][32]>(smem_tile);                 // [2]
  Real(*sTop)[5][32] = reinterpret_cast<Real(*)[5][32]>(smem_tile + 2 * R * 5 * 32);     // [3]
  Real(*sQ)[NQ][32] = reinterpret_cast<Real(*)[NQ][32]>(smem_tile + (2 * R + 3) * 5 * 32);   // [NR]
  // PPM only: left states handed from cell n to face n+1 ([R] rows + the top row double-buffered)
  constexpr int NWS = MHD ? 7 : 5;
  Real(*sW)[NWS][32] = reinterpret_cast<Real(*)[NWS][32]>(smem_tile + ((2 * R + 3) * 5 + NR * NQ) * 32);   // [R]
  Real(*sWtop)[NWS][32] = sW + R;                                                                            // [2]
  const Grid &g = a.g;
  const int lane = threadIdx.x, r = threadIdx.y;
  // Per-column invariants (MeshBlock, first cell, face area, flags) live in shared memory and are
  // re-read where needed: kept in registers they would be spilled around the Riemann solve, and
  // local-memory reloads miss the 56 KB of L1 this kernel leaves (profiles/).
  __shared__ idx_t sC0[32];
  __shared__ Real sArea[32];
  __shared__ int sBlk[32];   // blk | interior << 29 | colvalid << 30
  const idx_t s = (DIR == 1) ? g.sj : g.sk;
  const int n_s = (DIR == 1) ? g.js : g.ks;
  const int n_e = (DIR == 1) ? g.je : g.ke;
  if (r == 0) {
    const long long col = (long long)blockIdx.x * 32 + lane;
    const long long ncol = (long long)nb * n_o * ni;
    const bool colvalid = col < ncol;
    int blk = 0, o = o_lo, i = il;
    if (colvalid) {
      const int row = (int)(col / ni);
      i = il + (int)(col - (long long)row * ni);
      blk = row / n_o;
      o = o_lo + (row - blk * n_o);
    }
    const int k0 = (DIR == 1) ? o : 0, j0 = (DIR == 1) ? 0 : o;
    const bool interior = colvalid && (i >= g.is && i <= g.ie) &&
                          ((DIR == 1) ? (o >= g.ks && o <= g.ke) : (o >= g.js && o <= g.je));
    Real area = 0.0;
    if (interior) {
      const Real dx1 = a.geo.dxf[0][blk * a.geo.PG + i];
      area = dx1 * a.geo.dxf[(DIR == 1) ? 2 : 1][blk * a.geo.PG + o];   // Face2Area / Face3Area
    }
    sC0[lane] = (DIR == 1) ? gidx(g, k0, 0, i) : gidx(g, 0, j0, i);   // the column's cell 0 along the sweep
    sArea[lane] = area;
    sBlk[lane] = blk | (interior ? (1 << 29) : 0) | (colvalid ? (1 << 30) : 0);
  }
  __syncthreads();
  const volatile idx_t *vC0 = sC0;
  const volatile Real *vArea = sArea;
  const volatile int *vBlk = sBlk;
  Real *dst = (DIR == 1) ? a.d2 : a.d3;
  const int nf = n_e + 1 - n_s + 1;            // faces n_s .. n_e+1
  const int nsteps = (nf + R - 1) / R;
  const int vmap[5] = {IDN, IVX + DIR, IVX + (DIR + 1) % 3, IVX + (DIR + 2) % 3, IEN};
  const int row_first = n_s - LO;            // ring slot 0
  const int row_last = n_e + 1 + HI;         // last cell any face needs
  // asynchronous copy of cell row m of this thread's column into its ring slot
  auto stage_row = [&](int m) {
    const int bf = vBlk[lane];
    if ((bf >> 30) && m <= row_last) {
      const int blk = bf & 0x1fffffff;
      Real *d = &sQ[(m - row_first) % NR][0][lane];
      const idx_t cm = vC0[lane] + (idx_t)m * s;
      const Real *wsrc = a.w + (idx_t)blk * 5 * g.sn + cm;
      cp_async8(d, wsrc);
      cp_async8(d + 32, wsrc + (idx_t)(IVX + DIR) * g.sn);
      cp_async8(d + 2 * 32, wsrc + (idx_t)(IVX + (DIR + 1) % 3) * g.sn);
      cp_async8(d + 3 * 32, wsrc + (idx_t)(IVX + (DIR + 2) % 3) * g.sn);
      cp_async8(d + 4 * 32, wsrc + (idx_t)IPR * g.sn);
      if (MHD) {
        const Real *bc = a.bcc + (idx_t)blk * 3 * g.sn + cm;
        cp_async8(d + 5 * 32, bc + (idx_t)((DIR + 1) % 3) * g.sn);
        cp_async8(d + 6 * 32, bc + (idx_t)((DIR + 2) % 3) * g.sn);
        cp_async8(d + 7 * 32, a.b + ((idx_t)blk * 3 + DIR) * g.sn + cm);
      }
    }
  };
  // prologue: rows n_s-LO .. n_s+R-1+HI
  for (int m = r; m < R + LO + HI; m += R) stage_row(row_first + m);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (int step = 0; step < nsteps; ++step) {
    // n is re-derived from the (warp-uniform) step counter every iteration: left to itself the
    // compiler strength-reduces it into several per-thread induction variables, which it then has
    // to spill around the Riemann solve and reload -- through a 56 KB L1 -- at the top of the loop
    int sv = step;
    asm volatile("" : "+r"(sv));
    const int n = n_s + sv * R + r;
    const bool fvalid = (vBlk[lane] >> 30) && n <= n_e + 1;
    const int buf = step & 1;
    // rows first needed by the next step: n0+R+HI .. n0+2R-1+HI, one per thread row
    stage_row(n + R + HI);
    cp_async_commit();
    Real f[7];
    if constexpr (ATHK_SHARED_RECON(ORDER)) {
      // PPM: every thread reconstructs ITS cell n once (parabola -> left state of face n+1,
      // right state of face n) and hands the left state to the thread above through shared
      // memory, instead of every face rebuilding the parabolas of both adjacent cells.
      constexpr int NW = MHD ? 7 : 5;
      Real wr[7];
      const SmemStencil<MHD> st = {&sQ[0][0][lane], (n - row_first) % NR, NR};
      if (fvalid) {
        Real up[7];
        if constexpr (ORDER == 3) ppm_recon_cell<MHD>(a.eos, vBlk[lane] & 0x1fffffff, st, 0, up, wr);
        else plm_recon_cell<DIR, MHD>(a, vBlk[lane] & 0x1fffffff, st, 0, n, up, wr);
#pragma unroll
        for (int v = 0; v < NW; ++v) sW[r][v][lane] = up[v];
        if (r == R - 1) {
#pragma unroll
          for (int v = 0; v < NW; ++v) sWtop[step & 1][v][lane] = up[v];
        }
        if (step == 0 && r == 0) {   // nobody owns cell n_s-1: the first face's left state
          Real lo_unused[7];
          if constexpr (ORDER == 3) ppm_recon_cell<MHD>(a.eos, vBlk[lane] & 0x1fffffff, st, -1, up, lo_unused);
          else plm_recon_cell<DIR, MHD>(a, vBlk[lane] & 0x1fffffff, st, -1, n - 1, up, lo_unused);
#pragma unroll
          for (int v = 0; v < NW; ++v) sWtop[1][v][lane] = up[v];
        }
      }
      __syncthreads();
      if (fvalid) {
        const int blk = vBlk[lane] & 0x1fffffff;
        const idx_t c = vC0[lane] + (idx_t)n * s;
        Real wl[7];
        const Real(*src)[32] = (r > 0) ? sW[r - 1] : sWtop[(step + 1) & 1];
#pragma unroll
        for (int v = 0; v < NW; ++v) wl[v] = src[v][lane];
        solve_face<DIR, MHD, RS>(a, blk, c, n, wl, wr, st.bx(), f, true);
        Real(*mine)[32] = (r == R - 1) ? sTop[step % 3] : sF[buf][r];
        mine[0][lane] = f[IDN];
        mine[1][lane] = f[IVX];
        mine[2][lane] = f[IVY];
        mine[3][lane] = f[IVZ];
        mine[4][lane] = f[IEN];
      }
    } else if (fvalid) {
      const int blk = vBlk[lane] & 0x1fffffff;
      const idx_t c = vC0[lane] + (idx_t)n * s;
      const SmemStencil<MHD> st = {&sQ[0][0][lane], (n - row_first) % NR, NR};
      face_compute_from<DIR, ORDER, MHD, RS>(a, blk, c, n, st, f, true);
      Real(*mine)[32] = (r == R - 1) ? sTop[step % 3] : sF[buf][r];
      mine[0][lane] = f[IDN];
      mine[1][lane] = f[IVX];
      mine[2][lane] = f[IVY];
      mine[3][lane] = f[IVZ];
      mine[4][lane] = f[IEN];
    }
    cp_async_wait_all();
    __syncthreads();
    const int bf = vBlk[lane];
    if (fvalid && ((bf >> 29) & 1) && n > n_s) {   // this thread owns cell n-1: pure store, no RMW
      const Real(*below)[32] = (r > 0) ? sF[buf][r - 1] : sTop[(step + 2) % 3];
      const Real fcur[5] = {f[IDN], f[IVX], f[IVY], f[IVZ], f[IEN]};
      const idx_t cc = vC0[lane] + (idx_t)(n - 1) * s;
      const Real area = vArea[lane];
      const int blk = bf & 0x1fffffff;
#pragma unroll
      for (int v = 0; v < 5; ++v)
        dst[((idx_t)blk * 5 + vmap[v]) * g.sn + cc] = (area * fcur[v] - area * below[v][lane]);
    }
  }
}

template <int DIR, int ORDER, bool MHD, int RS>
__global__ void __launch_bounds__(32 * ATHK_TILE_R, ATHK_TILE_MINB)
k_sweep_tile(const SweepArgs a, const int nb, const int o_lo, const int n_o, const int il,
             const int ni) {
  constexpr int R = ATHK_TILE_R;
  __shared__ Real sF[3][R][5][32];
  const Grid &g = a.g;
  const int lane = threadIdx.x, r = threadIdx.y;
  const long long col = (long long)blockIdx.x * 32 + lane;
  const long long ncol = (long long)nb * n_o * ni;
  const bool colvalid = col < ncol;
  int blk = 0, o = o_lo, i = il;
  if (colvalid) {
    const int row = (int)(col / ni);
    i = il + (int)(col - (long long)row * ni);
    blk = row / n_o;
    o = o_lo + (row - blk * n_o);
  }
  const idx_t s = (DIR == 1) ? g.sj : g.sk;
  const int n_s = (DIR == 1) ? g.js : g.ks;
  const int n_e = (DIR == 1) ? g.je : g.ke;
  const int k0 = (DIR == 1) ? o : 0, j0 = (DIR == 1) ? 0 : o;
  const bool interior = colvalid && (i >= g.is && i <= g.ie) &&
                        ((DIR == 1) ? (o >= g.ks && o <= g.ke) : (o >= g.js && o <= g.je));
  Real area = 0.0;
  if (interior) {
    const Real dx1 = a.geo.dxf[0][blk * a.geo.PG + i];
    area = dx1 * a.geo.dxf[(DIR == 1) ? 2 : 1][blk * a.geo.PG + o];   // Face2Area / Face3Area
  }
  Real *dst = (DIR == 1) ? a.d2 : a.d3;
  const int nf = n_e + 1 - n_s + 1;            // faces n_s .. n_e+1
  const int nsteps = (nf + R - 1) / R;
  const int vmap[5] = {IDN, IVX + DIR, IVX + (DIR + 1) % 3, IVX + (DIR + 2) % 3, IEN};
  for (int step = 0; step < nsteps; ++step) {
    const int n = n_s + step * R + r;
    const bool fvalid = colvalid && n <= n_e + 1;
    const int buf = step % 3;
    Real f[7];
    idx_t c = 0;
    if (fvalid) {
      c = (DIR == 1) ? gidx(g, k0, n, i) : gidx(g, n, j0, i);
#if defined(ATHK_TILE_PREFETCH) && ATHK_TILE_PREFETCH > 0
#if ATHK_TILE_PREFETCH == 2
#define ATHK_PF "prefetch.global.L1 [%0];"
#else
#define ATHK_PF "prefetch.global.L2 [%0];"
#endif
      if (n + R <= n_e + 2) {   // next step's cell of this thread: pull its lines towards L2/L1
        const idx_t cn = c + (idx_t)R * s;
        const Real *wb = a.w + (idx_t)blk * 5 * g.sn + cn;
#pragma unroll
        for (int v = 0; v < 5; ++v) asm volatile(ATHK_PF ::"l"(wb + (idx_t)v * g.sn));
        if (MHD) {
          const Real *bc = a.bcc + (idx_t)blk * 3 * g.sn + cn;
          asm volatile(ATHK_PF ::"l"(bc + (idx_t)((DIR + 1) % 3) * g.sn));
          asm volatile(ATHK_PF ::"l"(bc + (idx_t)((DIR + 2) % 3) * g.sn));
          asm volatile(ATHK_PF ::"l"(a.b + ((idx_t)blk * 3 + DIR) * g.sn + cn));
        }
      }
#endif
      face_compute<DIR, ORDER, MHD, RS>(a, blk, c, n, f, true);
      sF[buf][r][0][lane] = f[IDN];
      sF[buf][r][1][lane] = f[IVX];
      sF[buf][r][2][lane] = f[IVY];
      sF[buf][r][3][lane] = f[IVZ];
      sF[buf][r][4][lane] = f[IEN];
    }
    __syncthreads();
    if (fvalid && interior && n > n_s) {   // this thread owns cell n-1: pure store, no RMW
      const Real(*below)[32] = (r > 0) ? sF[buf][r - 1] : sF[(step + 2) % 3][R - 1];
      const Real fcur[5] = {f[IDN], f[IVX], f[IVY], f[IVZ], f[IEN]};
      const idx_t cc = c - s;
#pragma unroll
      for (int v = 0; v < 5; ++v)
        dst[((idx_t)blk * 5 + vmap[v]) * g.sn + cc] = (area * fcur[v] - area * below[v][lane]);
    }
  }
}

//========================================================================================
// x2 / x3 sweep, thread-marching form (round 2; default).  The straight-line Newton sequences of
// athk_math.h want ~160 registers and no spills, i.e. 12 warps per SM -- too few to hide anything
// behind other warps, and CTA-wide barriers (tile kernel above) then cost more than they share.
// So every THREAD owns one column (lanes along i => every row access is a coalesced 256-byte
// segment) and marches along the sweep on its own:
//   * a private ring of cell rows in shared memory, [slot][variable][thread]: the cell it first
//     needs ATHK_MARCH_PD steps from now is copied by cp.async at the top of each step and waited
//     for with cp.async.wait_group -- a thread only ever reads what it copied itself, so the
//     kernel has NO barrier of any kind and warps drift freely;
//   * PLM / PPM reconstruct each cell ONCE: the left state of face n+1 is parked in the thread's
//     shared-memory slot (double-buffered, so a face can be recomputed with the compiler's
//     operations when the straight-line sequences flag it) until the next step picks it up, and so
//     is the previous face's flux for the divergence term -- nothing is carried in registers
//     across the Riemann solve.
// > 48 KB of dynamic shared memory is opt-in per kernel AND per device (a process may drive
// several GPUs): remember the largest size granted on each device
template <class K>
static inline cudaError_t ensure_dynamic_smem(K kernel, size_t smem, size_t *granted /*[64], zero-initialised*/) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (smem > granted[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    granted[dev] = smem;
  }
  return cudaSuccess;
}

template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef ATHK_MARCH_THREADS
#define ATHK_MARCH_THREADS 128
#endif
#ifndef ATHK_MARCH_MINB
#define ATHK_MARCH_MINB 3
#endif
#ifndef ATHK_MARCH_PD
#define ATHK_MARCH_PD 2
#endif

// the thread's rows: cell(off) = the cell `off` above the face's upper cell
template <bool MHD, int NR, int NT>
struct RingStencil {
  static constexpr int NQ = MHD ? 8 : 5;
  const Real *ring;   // + thread
  int slot0;          // ring slot of the face's upper cell
  __device__ __forceinline__ const Real *row(int off) const {
    int sl = slot0 + off;
    sl = (sl < 0) ? sl + NR : (sl >= NR ? sl - NR : sl);
    return ring + sl * (NQ * NT);
  }
  __device__ __forceinline__ void cell(int off, Real q[7]) const {
    const Real *p = row(off);
#pragma unroll
    for (int v = 0; v < (MHD ? 7 : 5); ++v) q[v] = p[v * NT];
  }
  __device__ __forceinline__ Real bx() const { return MHD ? row(0)[7 * NT] : 0.0; }
};

template <int DIR, int ORDER, bool MHD, int RS>
__global__ void __launch_bounds__(ATHK_MARCH_THREADS, ATHK_MARCH_MINB)
k_sweep_march(const SweepArgs a, const int nb, const int o_lo, const int n_o, const int il, const int ni) {
  constexpr int NT = ATHK_MARCH_THREADS, PD = ATHK_MARCH_PD;
  constexpr int LO = ORDER, HI = ORDER - 1;   // cells n-LO .. n+HI around face n
  constexpr int NR = LO + HI + 1 + PD;        // rows in use by a step + rows in flight
  constexpr int NQ = MHD ? 8 : 5, NW = MHD ? 7 : 5;
  extern __shared__ __align__(16) Real smem_march[];
  Real *sQ = smem_march + threadIdx.x;                       // [NR][NQ][NT]
  Real *sF = smem_march + NR * NQ * NT + threadIdx.x;        // [5][NT]   flux of the previous face
  Real *sW = sF + 5 * NT;                                    // [2][NW][NT] left state for the next face
  const Grid &g = a.g;
  const long long col = (long long)blockIdx.x * NT + threadIdx.x;
  if (col >= (long long)nb * n_o * ni) return;
  const int rowid = (int)(col / ni);
  const int i = il + (int)(col - (long long)rowid * ni);
  const int blk = rowid / n_o;
  const int o = o_lo + (rowid - blk * n_o);
  const idx_t s = (DIR == 1) ? g.sj : g.sk;
  const int n_s = (DIR == 1) ? g.js : g.ks;
  const int n_e = (DIR == 1) ? g.je : g.ke;
  const bool interior = (i >= g.is && i <= g.ie) && ((DIR == 1) ? (o >= g.ks && o <= g.ke) : (o >= g.js && o <= g.je));
  Real area = 0.0;
  if (interior) area = a.geo.dxf[0][blk * a.geo.PG + i] * a.geo.dxf[(DIR == 1) ? 2 : 1][blk * a.geo.PG + o];   // Face2Area / Face3Area
  const idx_t c0 = (DIR == 1) ? gidx(g, o, 0, i) : gidx(g, 0, o, i);   // the column's cell 0 along the sweep
  const Real *wsrc = a.w + (idx_t)blk * 5 * g.sn + c0;
  const Real *bcsrc = MHD ? a.bcc + (idx_t)blk * 3 * g.sn + c0 : nullptr;
  const Real *bsrc = MHD ? a.b + ((idx_t)blk * 3 + DIR) * g.sn + c0 : nullptr;
  Real *dst = ((DIR == 1) ? a.d2 : a.d3) + (idx_t)blk * 5 * g.sn + c0;
  const int row_first = n_s - LO;            // ring slot 0
  const int row_last = n_e + 1 + HI;         // last cell any face needs
  auto stage_row = [&](int m, int slot) {
    if (m <= row_last) {
      Real *d = sQ + slot * (NQ * NT);
      const idx_t off = (idx_t)m * s;
      cp_async8(d, wsrc + off);
      cp_async8(d + NT, wsrc + (idx_t)(IVX + DIR) * g.sn + off);
      cp_async8(d + 2 * NT, wsrc + (idx_t)(IVX + (DIR + 1) % 3) * g.sn + off);
      cp_async8(d + 3 * NT, wsrc + (idx_t)(IVX + (DIR + 2) % 3) * g.sn + off);
      cp_async8(d + 4 * NT, wsrc + (idx_t)IPR * g.sn + off);
      if (MHD) {
        cp_async8(d + 5 * NT, bcsrc + (idx_t)((DIR + 1) % 3) * g.sn + off);
        cp_async8(d + 6 * NT, bcsrc + (idx_t)((DIR + 2) % 3) * g.sn + off);
        cp_async8(d + 7 * NT, bsrc + off);
      }
    }
  };
  // prologue: rows n_s-LO .. n_s+HI+PD-1, one commit group per row beyond the first window
  for (int m = 0; m < LO + HI + 1; ++m) stage_row(row_first + m, m);
  cp_async_commit();
#pragma unroll
  for (int m = 0; m < PD - 1; ++m) {
    stage_row(row_first + LO + HI + 1 + m, LO + HI + 1 + m);
    cp_async_commit();
  }
  int slot_up = LO;                 // ring slot of cell n
  int slot_new = (NR - 1);          // ring slot the next staged row goes to (row n+HI+PD)
  const Real dfl = eos_dfloor(a.eos, blk), pfl = eos_pfloor(a.eos, blk);
  (void)dfl; (void)pfl;
  if constexpr (ORDER >= 2) {
    // nobody before the first face owns cell n_s-1: its upper-face state opens the hand-over chain
    cp_async_wait_group<PD - 1>();
    const RingStencil<MHD, NR, NT> st = {sQ, slot_up};
    with_math<RS>([&](auto &m) {
      Real up[7], lo[7];
      if constexpr (ORDER == 3) {
        Real qm2[7], qm1[7], q0[7], qp1[7], qp2[7];
        st.cell(-3, qm2); st.cell(-2, qm1); st.cell(-1, q0); st.cell(0, qp1); st.cell(1, qp2);
#pragma unroll
        for (int v = 0; v < NW; ++v) ppm_cell(qm2[v], qm1[v], q0[v], qp1[v], qp2[v], &up[v], &lo[v]);
        up[IDN] = (up[IDN] > dfl) ? up[IDN] : dfl;
        up[IPR] = (up[IPR] > pfl) ? up[IPR] : pfl;
      } else {
        Real qm[7], q0[7], qp[7];
        st.cell(-2, qm); st.cell(-1, q0); st.cell(0, qp);
        recon_cell_m<2, MHD>(m, qm, q0, qp, a.geo.cp[DIR][blk * a.geo.PG + n_s - 1], a.geo.cm[DIR][blk * a.geo.PG + n_s - 1], up, lo);
      }
#pragma unroll
      for (int v = 0; v < NW; ++v) sW[v * NT] = up[v];
    });
  }
  int par = 0;   // which half of sW holds the left state of the face about to be solved
  const Real *gdx = a.geo.dxf[DIR] + blk * a.geo.PG, *gcp = a.geo.cp[DIR] + blk * a.geo.PG,
             *gcm = a.geo.cm[DIR] + blk * a.geo.PG;
  for (int n = n_s; n <= n_e + 1; ++n) {
    // 1-D geometry of cell n, requested before the copies are waited for
    const Real dxn = MHD ? gdx[n] : 0.0, cpn = (ORDER == 2) ? gcp[n] : 0.0, cmn = (ORDER == 2) ? gcm[n] : 0.0;
    stage_row(n + HI + PD, slot_new);
    cp_async_commit();
    cp_async_wait_group<PD>();
    const idx_t c = c0 + (idx_t)n * s;   // (offset of cell n inside the block's component)
    const RingStencil<MHD, NR, NT> st = {sQ, slot_up};
    Real f[7];
    with_math<RS>([&](auto &m) {
      Real wl[7], wr[7];
      if constexpr (ORDER == 1) {
        st.cell(-1, wl);
        st.cell(0, wr);
      } else {
        Real up[7];
        if constexpr (ORDER == 3) {
          Real qm2[7], qm1[7], q0[7], qp1[7], qp2[7];
          st.cell(-2, qm2); st.cell(-1, qm1); st.cell(0, q0); st.cell(1, qp1); st.cell(2, qp2);
#pragma unroll
          for (int v = 0; v < NW; ++v) ppm_cell(qm2[v], qm1[v], q0[v], qp1[v], qp2[v], &up[v], &wr[v]);
          up[IDN] = (up[IDN] > dfl) ? up[IDN] : dfl;
          up[IPR] = (up[IPR] > pfl) ? up[IPR] : pfl;
          wr[IDN] = (wr[IDN] > dfl) ? wr[IDN] : dfl;
          wr[IPR] = (wr[IPR] > pfl) ? wr[IPR] : pfl;
        } else {
          Real qm[7], q0[7], qp[7];
          st.cell(-1, qm); st.cell(0, q0); st.cell(1, qp);
          recon_cell_m<2, MHD>(m, qm, q0, qp, cpn, cmn, up, wr);
        }
#pragma unroll
        for (int v = 0; v < NW; ++v) {
          wl[v] = sW[(par * NW + v) * NT];
          sW[((par ^ 1) * NW + v) * NT] = up[v];
        }
      }
      solve_face_dx<DIR, MHD, RS>(m, a, blk, c, dxn, wl, wr, st.bx(), f, true);
    });
    par ^= 1;
    if (interior && n > n_s) {   // cell n-1: (A F)(n) - (A F)(n-1)
      Real *dc = dst + (idx_t)(n - 1) * s;
      dc[0] = (area * f[IDN] - area * sF[0]);
      dc[(idx_t)(IVX + DIR) * g.sn] = (area * f[IVX] - area * sF[NT]);
      dc[(idx_t)(IVX + (DIR + 1) % 3) * g.sn] = (area * f[IVY] - area * sF[2 * NT]);
      dc[(idx_t)(IVX + (DIR + 2) % 3) * g.sn] = (area * f[IVZ] - area * sF[3 * NT]);
      dc[(idx_t)IEN * g.sn] = (area * f[IEN] - area * sF[4 * NT]);
    }
    sF[0] = f[IDN]; sF[NT] = f[IVX]; sF[2 * NT] = f[IVY]; sF[3 * NT] = f[IVZ]; sF[4 * NT] = f[IEN];
    slot_up = (slot_up + 1 == NR) ? 0 : slot_up + 1;
    slot_new = (slot_new + 1 == NR) ? 0 : slot_new + 1;
  }
  cp_async_wait_all();
}

#ifndef ATHK_MARCH
#define ATHK_MARCH 1
#endif

template <int DIR, int ORDER, bool MHD, int RS>
static inline cudaError_t launch_sweep_tile(cudaStream_t st, const SweepArgs &a, int nb, int o_lo,
                                            int o_hi, int il, int iu) {
  const int n_o = o_hi - o_lo + 1, ni = iu - il + 1;
  const long long ncol = (long long)nb * n_o * ni;
  if (ncol <= 0) return cudaSuccess;
#if ATHK_MARCH
  {
    constexpr int NT = ATHK_MARCH_THREADS, NR = 2 * ORDER + ATHK_MARCH_PD, NQ = MHD ? 8 : 5, NW = MHD ? 7 : 5;
    constexpr size_t smem = sizeof(Real) * NT * (NR * NQ + 5 + 2 * NW);
    static size_t granted[64];
    cudaError_t e = ensure_dynamic_smem(k_sweep_march<DIR, ORDER, MHD, RS>, smem, granted);
    if (e != cudaSuccess) return e;
    k_sweep_march<DIR, ORDER, MHD, RS><<<(unsigned)((ncol + NT - 1) / NT), NT, smem, st>>>(a, nb, o_lo, n_o, il, ni);
    return cudaGetLastError();
  }
#endif
  dim3 block(32, ATHK_TILE_R);
  const unsigned grid = (unsigned)((ncol + 31) / 32);
#if ATHK_TILE_STAGE
  constexpr int R = ATHK_TILE_R, NR = 2 * R + 2 * ORDER - 1, NQ = MHD ? 8 : 5;
  constexpr size_t smem = sizeof(Real) * 32 * ((2 * R + 3) * 5 + NR * NQ + (ATHK_SHARED_RECON(ORDER) ? (R + 2) * (MHD ? 7 : 5) : 0));
  static bool attr_set = false;   // > 48 KB of dynamic shared memory is opt-in, once per kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep_tile_staged<DIR, ORDER, MHD, RS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_sweep_tile_staged<DIR, ORDER, MHD, RS><<<grid, block, smem, st>>>(a, nb, o_lo, n_o, il, ni);
#else
  k_sweep_tile<DIR, ORDER, MHD, RS><<<grid, block, 0, st>>>(a, nb, o_lo, n_o, il, ni);
#endif
  return cudaGetLastError();
}

//========================================================================================
// x1 sweep launcher (kernel k_sweep_x1 in athk_backend.h).  A variant that staged each warp's next
// 32 faces in shared memory with cp.async while solving the current ones was measured SLOWER
// (2.04 vs 1.89 ms per PLM sweep at 256^3): the x1 sweep's long-scoreboard stalls are register-
// spill reloads, not the global loads, so it was dropped.
//========================================================================================
// x1 sweep with one reconstruction per cell (PPM, and PLM by ATHK_PLM_SHARE_X1): the cells
// i = is-1 .. ie+1 of all rows of a (k, MeshBlock) plane form one stream, lane L of a warp owns cell t = 30*warp + L,
// reconstructs it once (left state of face i+1, right state of face i) and passes the left state
// to lane L+1 with a shuffle; lanes 1..31 solve the face below their cell, lanes 1..30 finish
// their cell with the flux of lane L+1.  Consecutive warps overlap by two cells.  The face-centric
// kernel (k_sweep_x1, athk_backend.h; still used for donor cell) rebuilds the slopes / parabolas
// of both cells at every face (PPM: 10-14 ppm_cell per face instead of 5-7).
#ifndef ATHK_X1_PPM_SHARE
#define ATHK_X1_PPM_SHARE 1
#endif

template <bool MHD>
struct GlobalRowStencil {
  const SweepArgs &a;
  int blk;
  idx_t c;
  __device__ __forceinline__ void cell(int off, Real q[7]) const { load_q<0, MHD>(a, blk, c + off, q); }
};

template <int ORDER, bool MHD, int RS>
__global__ void __launch_bounds__(128, ATHK_X1_MINB)
k_sweep_x1_cells(const SweepX1<ORDER, MHD, RS> f, const int kl, const int jl, const int nj, const int icl,
               const int ncr, const unsigned ncr_magic) {
  constexpr int NW = MHD ? 7 : 5;
  const SweepArgs &a = f.a;
  const Grid &g = a.g;
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned total = (unsigned)(nj * ncr);          // cells in the stream
  const unsigned t = warp * 30 + lane;
  const bool valid = t < total;
  const int blk = (int)blockIdx.z, k = kl + (int)blockIdx.y;
  int j = jl, ci = 0;
  if (valid) {
    const unsigned row = __umulhi(t, ncr_magic);
    ci = (int)(t - row * (unsigned)ncr);
    j = jl + (int)row;
  }
  const int i = icl + ci;                               // this lane's cell
  const idx_t c = gidx(g, k, j, i);
  Real up[7], lo[7];
#pragma unroll
  for (int v = 0; v < 7; ++v) up[v] = lo[v] = 0.0;
  const bool cell_upd = valid && lane >= 1 && lane <= 30 && ci >= 1 && ci <= ncr - 2 && f.cell_active(k, j, i);
  if (valid) {
    if (cell_upd) f.prefetch_cell(blk, k, j, i);
    const GlobalRowStencil<MHD> st = {a, blk, c};
    if constexpr (ORDER == 3) {
      ppm_recon_cell<MHD>(a.eos, blk, st, 0, up, lo);
    } else if constexpr (ORDER == 2) {
      plm_recon_cell<0, MHD>(a, blk, st, 0, i, up, lo);
    } else {   // donor cell: both face states are the cell average (dc.cpp)
      st.cell(0, up);
#pragma unroll
      for (int v = 0; v < NW; ++v) lo[v] = up[v];
    }
  }
  Real wl[7];
#pragma unroll
  for (int v = 0; v < NW; ++v) wl[v] = __shfl_up_sync(0xffffffffu, up[v], 1);
  Real fl[7];
#pragma unroll
  for (int v = 0; v < 7; ++v) fl[v] = 0.0;
  // face i (between cells i-1 and i): needs the lane below in the same row
  if (valid && lane >= 1 && ci >= 1) {
    const Real bx = MHD ? a.b[(idx_t)blk * 3 * g.sn + c] : 0.0;
    // (the face of lane 31 is solved again by lane 1 of the next warp: same values, same stores)
    solve_face<0, MHD, RS>(a, blk, c, i, wl, lo, bx, fl, true);
  }
  Real fn[7];
#pragma unroll
  for (int v = 0; v < 5; ++v) fn[v] = __shfl_down_sync(0xffffffffu, fl[v], 1);
  if (cell_upd) f.store_d1(blk, k, j, i, fl, fn);
}

//========================================================================================
// x1 sweep, row-stream form (round 2; default).  Lane-per-cell stream and warp shuffles like
// k_sweep_x1_cells, but built for 12 resident warps per SM (160 registers, no spills):
//   * every WARP owns ATHK_X1R_NJW consecutive rows of one (k, MeshBlock) plane and walks its
//     stream of cells in passes of 32.  The rows it needs are fetched into a warp-private ring in
//     shared memory by TMA bulk copies (cp.async.bulk, one per variable and row, issued by lane 0
//     one pass ahead; completion on one mbarrier per ring slot), so stencils are shared-memory
//     reads, no lane spends issue slots on row loads, and the kernel has no CTA-wide barrier
//     (warps drift freely, as in k_sweep_march);
//   * lane L reconstructs cell t, solves the face below it with the left state of lane L-1 and
//     finishes cell t-1 with that face's flux and the flux of lane L-1; lane 0 picks both up from
//     where lane 31 of the previous pass left them -- 32 faces and 32 cells per pass, no face
//     solved twice;
//   * what the update reads at the END of the pass -- the x2/x3 flux-divergence terms and the
//     register it averages -- is requested with cp.async at the START of the pass into the lane's
//     own shared-memory slot: no registers held across the Riemann solve, no exposed latency;
//   * divisions of reconstruction, Riemann solve and flux divergence go through the math policy.
#ifndef ATHK_X1R_NJW
#define ATHK_X1R_NJW 11
#endif
#ifndef ATHK_X1R_MINB
#define ATHK_X1R_MINB 3
#endif
#ifndef ATHK_X1_ROWS
#define ATHK_X1_ROWS 1
#endif
#ifndef ATHK_X1R_TMA
#define ATHK_X1R_TMA 1
#endif

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (bytes: multiple of 16)
__device__ __forceinline__ void tma_bulk_g2s(void *smem, const void *gmem, unsigned bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
               "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

struct X1RowsGeom {
  int kl, nk, jl, nj, icl, ncr;   // planes, rows, first stream cell (il - 1), cells per row
  unsigned ncr_magic;
  int njw, ngroups;               // rows per warp, warps per plane
  int nrs_mask, nrs_shift;        // ring rows: a power of two
  int nwarps;                     // warps in the launch
  int wstride;                    // doubles of shared memory per warp
};

template <int ORDER, bool MHD, int RS>
__global__ void __launch_bounds__(128, ATHK_X1R_MINB)
k_sweep_x1_rows(const SweepX1<ORDER, MHD, RS> f, const X1RowsGeom q) {
  constexpr int NW = MHD ? 7 : 5, NQ = MHD ? 8 : 5;
  const SweepArgs &a = f.a;
  const Grid &g = a.g;
  const HydroIntegrate &h = a.hi;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (gw >= q.nwarps) return;
  const int grp = gw % q.ngroups;
  const int t2 = gw / q.ngroups;
  const int k = q.kl + t2 % q.nk;
  const int blk = t2 / q.nk;
  const int j0 = q.jl + grp * q.njw;
  const int nrows = min(q.njw, q.jl + q.nj - j0);
  const int P1 = g.P1;
  const int nrs = q.nrs_mask + 1;
  extern __shared__ __align__(16) Real smem_x1r[];
  Real *wbase = smem_x1r + (size_t)(threadIdx.x >> 5) * q.wstride;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(wbase);   // [nrs] (nrs even)
  Real *ring = wbase + nrs;                                          // [nrs][NQ][P1]
  Real *geo1 = ring + nrs * NQ * P1;                                 // [3][P1]: dx1f, cp, cm
  Real *carry = geo1 + 3 * P1;                                       // [12]: up[7], flux[5] of lane 31
  Real *slot = carry + 12 + lane;                                    // [15][32]
  // source rows: w(0..4) (DIR = 0: no rotation), bcc(1,2), b.x1f
  const idx_t plane = (idx_t)k * g.sk;
  const Real *wsrc = a.w + (idx_t)blk * 5 * g.sn + plane;
  const Real *bcsrc = MHD ? a.bcc + ((idx_t)blk * 3 + 1) * g.sn + plane : nullptr;
  const Real *bsrc = MHD ? a.b + (idx_t)blk * 3 * g.sn + plane : nullptr;
#if ATHK_X1R_TMA
  if (lane == 0) {
    for (int r = 0; r < nrs; ++r) mbar_init(&bars[r], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto load_row = [&](int r) {   // lane 0 only
    Real *d = ring + (r & q.nrs_mask) * (NQ * P1);
    void *bar = &bars[r & q.nrs_mask];
    const idx_t off = (idx_t)(j0 + r) * g.sj;
    const unsigned rb = (unsigned)P1 * 8u;
    mbar_expect_tx(bar, rb * NQ);
#pragma unroll
    for (int v = 0; v < 5; ++v) tma_bulk_g2s(d + v * P1, wsrc + (idx_t)v * g.sn + off, rb, bar);
    if (MHD) {
      tma_bulk_g2s(d + 5 * P1, bcsrc + off, rb, bar);
      tma_bulk_g2s(d + 6 * P1, bcsrc + g.sn + off, rb, bar);
      tma_bulk_g2s(d + 7 * P1, bsrc + off, rb, bar);
    }
  };
#else
  const int nch = P1 >> 1;   // 16-byte chunks per row
  auto load_row = [&](int r) {
    Real *d = ring + (r & q.nrs_mask) * (NQ * P1);
    const idx_t off = (idx_t)(j0 + r) * g.sj;
#pragma unroll
    for (int v = 0; v < 5; ++v)
      for (int ch = lane; ch < nch; ch += 32) cp_async16(d + v * P1 + 2 * ch, wsrc + (idx_t)v * g.sn + off + 2 * ch);
    if (MHD) {
      for (int ch = lane; ch < nch; ch += 32) {
        cp_async16(d + 5 * P1 + 2 * ch, bcsrc + off + 2 * ch);
        cp_async16(d + 6 * P1 + 2 * ch, bcsrc + g.sn + off + 2 * ch);
        cp_async16(d + 7 * P1 + 2 * ch, bsrc + off + 2 * ch);
      }
    }
  };
#endif
  for (int ii = lane; ii < P1; ii += 32) {
    const bool in = ii < g.nc1;
    geo1[ii] = in ? a.geo.dxf[0][blk * a.geo.PG + ii] : 1.0;
    geo1[P1 + ii] = in ? a.geo.cp[0][blk * a.geo.PG + ii] : 0.0;
    geo1[2 * P1 + ii] = in ? a.geo.cm[0][blk * a.geo.PG + ii] : 0.0;
  }
  const int L = nrows * q.ncr;              // cells in this warp's stream
  const int npass = (L + 31) >> 5;
  auto row_hi = [&](int p) {                // last row pass p touches
    int t = 32 * p + 31;
    if (t > L - 1) t = L - 1;
    return (int)__umulhi((unsigned)t, q.ncr_magic);
  };
  int rows_issued = 0;
  {
    const int rh = row_hi(0);
#if ATHK_X1R_TMA
    if (lane == 0) for (int r = 0; r <= rh; ++r) load_row(r);
    rows_issued = rh + 1;
#else
    while (rows_issued <= rh) load_row(rows_issued++);
    cp_async_commit();
#endif
  }
  const Real dx3 = a.geo.dxf[2][blk * a.geo.PG + k];
  const bool needs_u = (h.delta != 0.0 || !h.swap), needs_u1 = !h.u1_zero;
  const Real *ureg = needs_u ? h.u : h.u1;
  const idx_t blk5 = (idx_t)blk * 5 * g.sn;
  for (int p = 0; p < npass; ++p) {
    __syncwarp();   // every lane is done with the rows (and the hand-over) of pass p-1
    if (p + 1 < npass) {
      const int rh = row_hi(p + 1);
#if ATHK_X1R_TMA
      if (lane == 0) for (int r = rows_issued; r <= rh; ++r) load_row(r);
      if (rh >= rows_issued) rows_issued = rh + 1;
#else
      while (rows_issued <= rh) load_row(rows_issued++);
#endif
    }
    const int t = 32 * p + lane;
    const bool valid = t < L;
    int row = 0, ci = 0;
    if (valid) {
      row = (int)__umulhi((unsigned)t, q.ncr_magic);
      ci = t - row * q.ncr;
    }
    const int j = j0 + row, i = q.icl + ci;
    const idx_t c = plane + (idx_t)j * g.sj + i;
    const bool face = valid && ci >= 1;                                     // face i: below this cell, same row
    const bool cell_upd = face && ci >= 2 && f.cell_active(k, j, i - 1);    // this lane finishes cell i-1
    if (cell_upd) {   // what the end of the pass reads, requested now
      const idx_t o = blk5 + c - 1;
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        cp_async8(slot + v * 32, a.d2 + o + (idx_t)v * g.sn);
        cp_async8(slot + (5 + v) * 32, a.d3 + o + (idx_t)v * g.sn);
        cp_async8(slot + (10 + v) * 32, ureg + o + (idx_t)v * g.sn);
      }
    }
    cp_async_commit();
#if ATHK_X1R_TMA
    if (valid) mbar_wait(&bars[row & q.nrs_mask], (unsigned)(row >> q.nrs_shift) & 1u);
#else
    cp_async_wait_group<1>();
    __syncwarp();
#endif
    const Real *rp = ring + (row & q.nrs_mask) * (NQ * P1) + i;
    Real up[7], lo[7];
#pragma unroll
    for (int v = 0; v < 7; ++v) up[v] = lo[v] = 0.0;
    if (valid) {
      if constexpr (ORDER == 1) {
#pragma unroll
        for (int v = 0; v < NW; ++v) up[v] = lo[v] = rp[v * P1];
      } else if constexpr (ORDER == 2) {
        const Real cp = geo1[P1 + i], cm = geo1[2 * P1 + i];
        with_math<RS>([&](auto &m) {
          Real qm[7], q0[7], qp[7];
#pragma unroll
          for (int v = 0; v < NW; ++v) { qm[v] = rp[v * P1 - 1]; q0[v] = rp[v * P1]; qp[v] = rp[v * P1 + 1]; }
          recon_cell_m<2, MHD>(m, qm, q0, qp, cp, cm, up, lo);
        });
      } else {
#pragma unroll
        for (int v = 0; v < NW; ++v)
          ppm_cell(rp[v * P1 - 2], rp[v * P1 - 1], rp[v * P1], rp[v * P1 + 1], rp[v * P1 + 2], &up[v], &lo[v]);
        const Real dfl = eos_dfloor(a.eos, blk), pfl = eos_pfloor(a.eos, blk);
        up[IDN] = (up[IDN] > dfl) ? up[IDN] : dfl;
        up[IPR] = (up[IPR] > pfl) ? up[IPR] : pfl;
        lo[IDN] = (lo[IDN] > dfl) ? lo[IDN] : dfl;
        lo[IPR] = (lo[IPR] > pfl) ? lo[IPR] : pfl;
      }
    }
    // left state of the face: the cell below, i.e. lane L-1 (lane 0: lane 31 of the previous pass)
    Real wl[7];
#pragma unroll
    for (int v = 0; v < NW; ++v) {
      wl[v] = __shfl_up_sync(0xffffffffu, up[v], 1);
      if (lane == 0) wl[v] = carry[v];
    }
    Real fl[7];
#pragma unroll
    for (int v = 0; v < 7; ++v) fl[v] = 0.0;
    if (face) {
      const Real bx = MHD ? rp[7 * P1] : 0.0;
      const Real dx1 = geo1[i];
      with_math<RS>([&](auto &m) { solve_face_dx<0, MHD, RS>(m, a, blk, c, dx1, wl, lo, bx, fl, true); });
    }
    // flux of the face below cell i-1: lane L-1 (lane 0: lane 31 of the previous pass)
    Real fb[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      fb[v] = __shfl_up_sync(0xffffffffu, fl[v], 1);
      if (lane == 0) fb[v] = carry[7 + v];
    }
    __syncwarp();   // lane 0 has read the hand-over of the previous pass
    if (lane == 31) {
#pragma unroll
      for (int v = 0; v < NW; ++v) carry[v] = up[v];
#pragma unroll
      for (int v = 0; v < 5; ++v) carry[7 + v] = fl[v];
    }
    cp_async_wait_all();
    if (cell_upd) {
      const Real dx1 = geo1[i - 1], dx2 = a.geo.dxf[1][blk * a.geo.PG + j];
      const Real area = dx2 * dx3, vol = dx1 * dx2 * dx3;   // Face1Area, cell volume
      const idx_t o = blk5 + c - 1;
      with_math<RS>([&](auto &m) {
#pragma unroll
        for (int v = 0; v < 5; ++v) {   // DIR = 0: rotated index == variable index
          Real d = (area * fl[v] - area * fb[v]);
          d += slot[v * 32];
          d += slot[(5 + v) * 32];
          const idx_t ov = o + (idx_t)v * g.sn;
          const Real ur = slot[(10 + v) * 32];
          const Real un = needs_u ? ur : 0.0;
          const Real u1n = needs_u1 ? (needs_u ? h.u1[ov] : ur) : 0.0;
          integrate_var_m(m, h, ov, d, vol, un, u1n);
        }
      });
    }
  }
}

template <int ORDER, bool MHD, int RS>
static inline cudaError_t launch_sweep_x1(cudaStream_t st, int nb, const Range3 &r,
                                          const SweepX1<ORDER, MHD, RS> &f) {
  const int nif = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (nif <= 1 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
#if ATHK_X1_ROWS
  {
    constexpr int NQ = MHD ? 8 : 5, NI = 15;
    const Grid &g = f.a.g;
    X1RowsGeom q;
    q.kl = r.kl; q.nk = nk; q.jl = r.jl; q.nj = nj; q.icl = r.il - 1; q.ncr = nif + 1;
    q.ncr_magic = magic_u32(q.ncr);
    q.njw = ATHK_X1R_NJW < nj ? ATHK_X1R_NJW : nj;
    q.ngroups = (nj + q.njw - 1) / q.njw;
    int nrs = 63 / q.ncr + 2, p2 = 2, sh = 1;   // rows two consecutive passes can span
    while (p2 < nrs) { p2 *= 2; ++sh; }
    q.nrs_mask = p2 - 1;
    q.nrs_shift = sh;
    q.nwarps = nb * nk * q.ngroups;
    q.wstride = p2 + p2 * NQ * g.P1 + 3 * g.P1 + 12 + NI * 32;   // mbarriers, ring, geometry, hand-over, lane slots
    q.wstride = (q.wstride + 1) & ~1;   // keep every warp's region 16-byte aligned
    const size_t smem = sizeof(Real) * 4 * (size_t)q.wstride;
    static size_t granted[64];
    cudaError_t e = ensure_dynamic_smem(k_sweep_x1_rows<ORDER, MHD, RS>, smem, granted);
    if (e != cudaSuccess) return e;
    k_sweep_x1_rows<ORDER, MHD, RS><<<(unsigned)((q.nwarps + 3) / 4), 128, smem, st>>>(f, q);
    return cudaGetLastError();
  }
#endif
#if ATHK_X1_PPM_SHARE
  if constexpr (ATHK_SHARED_RECON_X1(ORDER)) {
    const int ncr = nif + 1;                       // cells il-1 .. iu of a row
    const int total = nj * ncr;
    const int nwarps = (total - 2 + 29) / 30;      // stride 30, two cells of overlap
    dim3 grid((unsigned)((nwarps + 3) / 4), (unsigned)nk, (unsigned)nb);
    k_sweep_x1_cells<ORDER, MHD, RS><<<grid, 128, 0, st>>>(f, r.kl, r.jl, nj, r.il - 1, ncr, magic_u32(ncr));
    return cudaGetLastError();
  }
#endif
  const int total = nj * nif;
  const int nchunks = (total - 1 + 30) / 31;
  dim3 grid((unsigned)((nchunks + 3) / 4), (unsigned)nk, (unsigned)nb);
  k_sweep_x1<SweepX1<ORDER, MHD, RS>><<<grid, 128, 0, st>>>(f, r.kl, r.jl, nj, r.il, nif, magic_u32(nif));
  return cudaGetLastError();
}

}  // namespace athk
#endif  // ATHK_SWEEP_KERNELS_CUH_
