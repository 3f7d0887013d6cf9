// athk_sweep_kernels.cuh -- CUDA-only launch shapes of the fused sweeps (athk_sweep.h holds the
// arithmetic, shared with the host harness).
//
// HLLD in FP64 with the straight-line Newton sequences of athk_math.h wants ~160 registers and no
// spills, i.e. 12 resident warps per SM.  Both kernels are therefore built so that a warp never
// waits for another one: no CTA-wide barrier, operands staged in shared memory by asynchronous
// copies (cp.async per thread in the marching kernel, TMA bulk copies in the row kernel), nothing
// carried in registers across a Riemann solve.
//   k_sweep_march   x2 / x3: one thread per column, marching along the sweep
//   k_sweep_x1_rows x1:      one lane per cell of a row stream, warp shuffles; runs last, sums the
//                            three flux-divergence terms in the reference's order and applies the
//                            integrator update
// (Round 1's CTA-marching tile kernels and the face-per-lane x1 kernel were retired with the move
//  to 12 warps per SM: their barriers / direct global loads need 20+ warps to hide.)
#ifndef ATHK_SWEEP_KERNELS_CUH_
#define ATHK_SWEEP_KERNELS_CUH_

#include "athk_sweep.h"

namespace athk {

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

//========================================================================================
// x2 / x3 sweep, thread-marching form.  The straight-line Newton sequences of
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

#ifndef ATHK_MARCH_MINB_DONOR
#define ATHK_MARCH_MINB_DONOR ATHK_MARCH_MINB
#endif
// (nseg, seglen: on small meshes a column is cut into nseg segments of seglen cells, each marched by
//  its own thread, so that one 64^3 block still fills the machine; a segment solves seglen + 1
//  faces -- the face it shares with the next segment is solved by both, same bits -- and finishes
//  its own seglen cells)
// (SEG = false is the whole-column kernel of the large meshes, compiled without the segment's limits:
//  they cost the x3 PLM instantiation 24 bytes of spills and 13 % of its time)
template <int DIR, int ORDER, bool MHD, int RS, int AR, bool SEG>
__global__ void __launch_bounds__(ATHK_MARCH_THREADS, (ORDER == 1) ? ATHK_MARCH_MINB_DONOR : ATHK_MARCH_MINB)
k_sweep_march(const SweepArgs a, const int nb, const int o_lo, const int n_o, const int il, const int ni,
              const int nseg, const int seglen) {
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
  const int seg = SEG ? blockIdx.y : 0;
  const int rowid = (int)(col / ni);
  const int i = il + (int)(col - (long long)rowid * ni);
  const int blk = rowid / n_o;
  const int o = o_lo + (rowid - blk * n_o);
  const idx_t s = (DIR == 1) ? g.sj : g.sk;
  const int n_s0 = (DIR == 1) ? g.js : g.ks;
  const int n_e0 = (DIR == 1) ? g.je : g.ke;
  // this thread's cells n_s .. n_e (faces n_s .. n_e + 1)
  const int n_s = SEG ? n_s0 + seg * seglen : n_s0;
  const int n_e = SEG ? ((n_s + seglen - 1 < n_e0) ? n_s + seglen - 1 : n_e0) : n_e0;
  if (SEG && n_s > n_e0) return;
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
    with_math<RS, AR>([&](auto &m) {
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
    with_math<RS, AR>([&](auto &m) {
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

template <int DIR, int ORDER, bool MHD, int RS, int AR>
static inline cudaError_t launch_sweep_march(cudaStream_t st, const SweepArgs &a, int nb, int o_lo,
                                             int o_hi, int il, int iu) {
  const int n_o = o_hi - o_lo + 1, ni = iu - il + 1;
  const long long ncol = (long long)nb * n_o * ni;
  if (ncol <= 0) return cudaSuccess;
  // segments per column: enough threads for ~2 waves of a B200 (148 SMs x 3 CTAs x 128 threads),
  // segments no shorter than 8 cells
  const int nx = (DIR == 1) ? a.g.nx2 : a.g.nx3;
  int nseg = 1;
  {
    const long long want = 2LL * 148 * 3 * ATHK_MARCH_THREADS;
    static const int forced = getenv("ATHK_MARCH_NSEG") ? atoi(getenv("ATHK_MARCH_NSEG")) : 0;
    if (forced > 0) nseg = forced;
    else while (ncol * nseg < want && nx / (2 * nseg) >= 8) nseg *= 2;
    if (nseg > nx) nseg = nx;
  }
  const int seglen = (nx + nseg - 1) / nseg;
  constexpr int NT = ATHK_MARCH_THREADS, NR = 2 * ORDER + ATHK_MARCH_PD, NQ = MHD ? 8 : 5, NW = MHD ? 7 : 5;
  constexpr size_t smem = sizeof(Real) * NT * (NR * NQ + 5 + 2 * NW);
  static size_t granted[2][64];
  if (nseg == 1) {
    cudaError_t e = ensure_dynamic_smem(k_sweep_march<DIR, ORDER, MHD, RS, AR, false>, smem, granted[0]);
    if (e != cudaSuccess) return e;
    k_sweep_march<DIR, ORDER, MHD, RS, AR, false><<<(unsigned)((ncol + NT - 1) / NT), NT, smem, st>>>(a, nb, o_lo, n_o, il, ni, 1, nx);
  } else {
    cudaError_t e = ensure_dynamic_smem(k_sweep_march<DIR, ORDER, MHD, RS, AR, true>, smem, granted[1]);
    if (e != cudaSuccess) return e;
    k_sweep_march<DIR, ORDER, MHD, RS, AR, true><<<dim3((unsigned)((ncol + NT - 1) / NT), (unsigned)nseg), NT, smem, st>>>(a, nb, o_lo, n_o, il, ni, nseg, seglen);
  }
  return cudaGetLastError();
}

//========================================================================================
// x1 sweep, row-stream form.  One lane per cell of a stream of rows, neighbours by warp shuffle,
// built for 12 resident warps per SM (160 registers, no spills):
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

template <int ORDER, bool MHD, int RS, int AR>
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
        with_math<RS, AR>([&](auto &m) {
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
      with_math<RS, AR>([&](auto &m) { solve_face_dx<0, MHD, RS>(m, a, blk, c, dx1, wl, lo, bx, fl, true); });
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
      with_math<RS, AR>([&](auto &m) {
        const Real rvol = m.div_rcp(vol);
#pragma unroll
        for (int v = 0; v < 5; ++v) {   // DIR = 0: rotated index == variable index
          Real d = (area * fl[v] - area * fb[v]);
          d += slot[v * 32];
          d += slot[(5 + v) * 32];
          const idx_t ov = o + (idx_t)v * g.sn;
          const Real ur = slot[(10 + v) * 32];
          const Real un = needs_u ? ur : 0.0;
          const Real u1n = needs_u1 ? (needs_u ? h.u1[ov] : ur) : 0.0;
          integrate_var_m(m, h, ov, d, vol, rvol, un, u1n);
        }
      });
    }
  }
}

template <int ORDER, bool MHD, int RS, int AR>
static inline cudaError_t launch_sweep_x1(cudaStream_t st, int nb, const Range3 &r,
                                          const SweepX1<ORDER, MHD, RS> &f) {
  const int nif = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (nif <= 1 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  constexpr int NQ = MHD ? 8 : 5, NI = 15;
  const Grid &g = f.a.g;
  X1RowsGeom q;
  q.kl = r.kl; q.nk = nk; q.jl = r.jl; q.nj = nj; q.icl = r.il - 1; q.ncr = nif + 1;
  q.ncr_magic = magic_u32(q.ncr);
  q.njw = ATHK_X1R_NJW < nj ? ATHK_X1R_NJW : nj;
  // small meshes: fewer rows per warp, so that there are ~2 waves of warps (148 SMs x 12)
  while (q.njw > 1 && (long long)nb * nk * ((nj + q.njw - 1) / q.njw) < 2LL * 148 * 12) q.njw = (q.njw + 1) / 2;
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
  cudaError_t e = ensure_dynamic_smem(k_sweep_x1_rows<ORDER, MHD, RS, AR>, smem, granted);
  if (e != cudaSuccess) return e;
  k_sweep_x1_rows<ORDER, MHD, RS, AR><<<(unsigned)((q.nwarps + 3) / 4), 128, smem, st>>>(f, q);
  return cudaGetLastError();
}

}  // namespace athk
#endif  // ATHK_SWEEP_KERNELS_CUH_
