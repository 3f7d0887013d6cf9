// athk_backend.h -- thin device-runtime layer used by athk_api.cu: CUDA runtime + the generic
// one-thread-per-point launchers (nvcc, sm_100a).
// (The test harness under tests/emul/ compiles the operator bodies for the host with
//  -DATHK_EMULATE and substitutes its own two headers for the CUDA parts of this file; that
//  library is never built into the package nor loaded by it.)
#ifndef ATHK_BACKEND_H_
#define ATHK_BACKEND_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef ATHK_EMULATE
#include <cuda_runtime.h>
#else
// test-harness build only: stand-ins for the CUDA runtime calls and plain-loop launchers live under
// tests/emul/ (on the include path of that build alone)
#include "athk_emul_runtime.h"
#endif

namespace athk {

struct Range3 {
  int kl, ku, jl, ju, il, iu;   // inclusive
};

#ifndef ATHK_EMULATE
// One thread per point of the (blk,k,j,i) box.  grid.x covers one (j,i) plane flattened with i
// fastest (consecutive lanes -> consecutive addresses; rows of 64-70 doubles would leave a 2-D
// thread block half idle), grid.y = k, grid.z = MeshBlock: the plane decode is block-uniform and
// the single per-thread division t / ni is a multiply-high by a host-computed reciprocal
// (exact for t*ni < 2^32).
#ifndef ATHK_KP_THREADS
#define ATHK_KP_THREADS 128
#endif
#ifndef ATHK_KP_MINB
#define ATHK_KP_MINB 5
#endif
static inline unsigned magic_u32(int d) { return (unsigned)((0x100000000ULL + (unsigned)d - 1) / (unsigned)d); }

template <class F, int MINB>
__global__ void __launch_bounds__(ATHK_KP_THREADS, MINB)
k_points(const F f, const int kl, const int jl, const int nj, const int il, const int ni,
         const unsigned ni_magic) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)(nj * ni)) return;
  const unsigned j = (ni == 1) ? t : __umulhi(t, ni_magic);
  const unsigned i = t - j * (unsigned)ni;
  f((int)blockIdx.z, kl + (int)blockIdx.y, jl + (int)j, il + (int)i);
}

// MINB = resident CTAs per SM the kernel is compiled for (register budget 65536 / (128 * MINB)):
// 5 for the FP64-heavy face kernels, more for streaming kernels that need loads in flight
template <int MINB, class F>
static inline cudaError_t launch_points_occ(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  const int threads = ATHK_KP_THREADS;
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  dim3 grid((unsigned)((nj * ni + threads - 1) / threads), (unsigned)nk, (unsigned)nb);
  k_points<F, MINB><<<grid, threads, 0, st>>>(f, r.kl, r.jl, nj, r.il, ni, magic_u32(ni));
  return cudaGetLastError();
}

// The same (j,i) plane decode, but every thread walks KT consecutive k planes (f.col): stencil
// kernels whose points share operands with the plane below (corner E, CT) carry those in registers
// instead of loading them again, and issue all loads of a plane before its first store.
template <class F, int MINB>
__global__ void __launch_bounds__(ATHK_KP_THREADS, MINB)
k_points_kcol(const F f, const int kl, const int ku, const int kt, const int jl, const int nj, const int il,
              const int ni, const unsigned ni_magic) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)(nj * ni)) return;
  const unsigned j = (ni == 1) ? t : __umulhi(t, ni_magic);
  const unsigned i = t - j * (unsigned)ni;
  const int k0 = kl + (int)blockIdx.y * kt;
  const int k1 = (k0 + kt - 1 < ku) ? k0 + kt - 1 : ku;
  f.col((int)blockIdx.z, k0, k1, jl + (int)j, il + (int)i);
}

template <int MINB, class F>
static inline cudaError_t launch_points_kcol(cudaStream_t st, int nb, const Range3 &r, int kt, const F &f) {
  const int threads = ATHK_KP_THREADS;
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  if (kt < 1) kt = 1;
  dim3 grid((unsigned)((nj * ni + threads - 1) / threads), (unsigned)((nk + kt - 1) / kt), (unsigned)nb);
  k_points_kcol<F, MINB><<<grid, threads, 0, st>>>(f, r.kl, r.ku, kt, r.jl, nj, r.il, ni, magic_u32(ni));
  return cudaGetLastError();
}

template <class F>
static inline cudaError_t launch_points(cudaStream_t st, int nb, const Range3 &r, const F &f,
                                        int = ATHK_KP_THREADS) {
  return launch_points_occ<ATHK_KP_MINB>(st, nb, r, f);
}

// The same launch shape for a functor whose run() returns a per-point value to be min-reduced per
// MeshBlock (Hydro::NewBlockTimeStep folded into W(U)): warp shuffles + shared memory, then ONE
// atomicMin per CTA on out[blk] (positive doubles order like their bit patterns; min is exact and
// order independent).  No thread leaves before the reduction.
template <class F, int MINB>
__global__ void __launch_bounds__(ATHK_KP_THREADS, MINB)
k_points_min(const F f, double *out, const int kl, const int jl, const int nj, const int il, const int ni,
             const unsigned ni_magic) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 1.0e308;
  if (t < (unsigned)(nj * ni)) {
    const unsigned j = (ni == 1) ? t : __umulhi(t, ni_magic);
    const unsigned i = t - j * (unsigned)ni;
    m = f.run((int)blockIdx.z, kl + (int)blockIdx.y, jl + (int)j, il + (int)i);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, m, o);
    m = (v < m) ? v : m;
  }
  __shared__ double sm[ATHK_KP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < ATHK_KP_THREADS / 32; ++w) m = (sm[w] < m) ? sm[w] : m;
    if (m < 1.0e300 && m < out[blockIdx.z])
      atomicMin((unsigned long long *)&out[blockIdx.z], (unsigned long long)__double_as_longlong(m));
  }
}

template <int MINB, class F>
static inline cudaError_t launch_points_min(cudaStream_t st, int nb, const Range3 &r, const F &f, double *out) {
  const int threads = ATHK_KP_THREADS;
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  dim3 grid((unsigned)((nj * ni + threads - 1) / threads), (unsigned)nk, (unsigned)nb);
  k_points_min<F, MINB><<<grid, threads, 0, st>>>(f, out, r.kl, r.jl, nj, r.il, ni, magic_u32(ni));
  return cudaGetLastError();
}

// Points of an outer box that are NOT in an inner box (ghost zones around the active zone, the
// boundary surface of the edge box): six disjoint slabs, flattened into one index so that only
// the shell is launched -- a 64^3 block has 17 % of its cells in the ghost shell and 9 % of its
// edges on the surface.
struct Shell {
  int k0[6], j0[6], i0[6], nj[6], ni[6];
  unsigned end[6];   // cumulative point counts
};
static inline Shell make_shell(const Range3 &o, const Range3 &in) {
  Shell s;
  const Range3 b[6] = {{o.kl, in.kl - 1, o.jl, o.ju, o.il, o.iu},   {in.ku + 1, o.ku, o.jl, o.ju, o.il, o.iu},
                       {in.kl, in.ku, o.jl, in.jl - 1, o.il, o.iu}, {in.kl, in.ku, in.ju + 1, o.ju, o.il, o.iu},
                       {in.kl, in.ku, in.jl, in.ju, o.il, in.il - 1}, {in.kl, in.ku, in.jl, in.ju, in.iu + 1, o.iu}};
  unsigned acc = 0;
  for (int q = 0; q < 6; ++q) {
    const int nk = b[q].ku - b[q].kl + 1, nj = b[q].ju - b[q].jl + 1, ni = b[q].iu - b[q].il + 1;
    s.k0[q] = b[q].kl; s.j0[q] = b[q].jl; s.i0[q] = b[q].il;
    s.nj[q] = nj > 0 ? nj : 1; s.ni[q] = ni > 0 ? ni : 1;
    if (nk > 0 && nj > 0 && ni > 0) acc += (unsigned)nk * nj * ni;
    s.end[q] = acc;
  }
  return s;
}

template <class F>
__global__ void __launch_bounds__(256) k_shell(const F f, const Shell s) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= s.end[5]) return;
  int q = 0;
#pragma unroll
  for (int m = 0; m < 5; ++m) q += (t >= s.end[m]) ? 1 : 0;
  if (q > 0) t -= s.end[q - 1];
  const unsigned ni = (unsigned)s.ni[q], nj = (unsigned)s.nj[q];
  const unsigned row = t / ni, i = t - row * ni;
  const unsigned k = row / nj, j = row - k * nj;
  f((int)blockIdx.y, s.k0[q] + (int)k, s.j0[q] + (int)j, s.i0[q] + (int)i);
}

template <class F>
static inline cudaError_t launch_shell(cudaStream_t st, int nb, const Range3 &outer, const Range3 &inner,
                                       const F &f) {
  const Shell s = make_shell(outer, inner);
  if (s.end[5] == 0 || nb <= 0) return cudaSuccess;
  dim3 grid((s.end[5] + 255) / 256, (unsigned)nb);
  k_shell<F><<<grid, 256, 0, st>>>(f, s);
  return cudaGetLastError();
}

// (launch_sweep_x1: athk_sweep_kernels.cuh)
// (link, element) launcher for the variable-length remote messages
template <class F>
__global__ void __launch_bounds__(256) k_links(const F f, const long long maxcount) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < maxcount) f((int)blockIdx.y, e);
}
template <class F>
static inline cudaError_t launch_links(cudaStream_t st, int nlinks, long long maxcount, const F &f) {
  if (nlinks <= 0 || maxcount <= 0) return cudaSuccess;
  dim3 grid((unsigned)((maxcount + 255) / 256), (unsigned)nlinks);
  k_links<F><<<grid, 256, 0, st>>>(f, maxcount);
  return cudaGetLastError();
}
#else
#include "athk_emul_launch.h"
#endif

}  // namespace athk
#endif  // ATHK_BACKEND_H_
