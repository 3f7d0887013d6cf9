// athk_backend.h -- thin device-runtime layer used by athk_api.cu.
//
// Product build (nvcc, sm_100a): CUDA runtime + one-thread-per-point kernels.
// Test-harness build (g++ -DATHK_EMULATE, tests/emul/ only): the same operator bodies are run by
// plain host loops so that index logic can be debugged in a container without a GPU.  The
// harness library has a different soname (libathk_emul.so), is never built by
// __graft_entry__.build() into the package and is never loaded by athena_b200; it is not a
// fallback.
#ifndef ATHK_BACKEND_H_
#define ATHK_BACKEND_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef ATHK_EMULATE
#include <cuda_runtime.h>
#else
// ---- minimal stand-ins for the CUDA runtime calls the API layer uses ----------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
#define cudaSuccess 0
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
#endif

namespace athk {

struct Range3 {
  int kl, ku, jl, ju, il, iu;   // inclusive
};

#ifndef ATHK_EMULATE
// One thread per point of the (blk,k,j,i) box, flattened with i fastest so that consecutive
// lanes touch consecutive addresses (rows of 64-70 doubles would leave half of a 2-D thread
// block idle).  256-thread CTAs; the grid is sized to the exact point count.
#ifndef ATHK_KP_THREADS
#define ATHK_KP_THREADS 128
#endif
#ifndef ATHK_KP_MINB
#define ATHK_KP_MINB 5
#endif
template <class F>
__global__ void __launch_bounds__(ATHK_KP_THREADS, ATHK_KP_MINB) k_points(const F f, const int nb, const int kl, const int nk,
                                                const int jl, const int nj, const int il,
                                                const int ni) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nb * nk * nj * ni;
  if (t >= total) return;
  const int row = (int)(t / ni);
  const int i = (int)(t - (long long)row * ni);
  const int pl = row / nj;
  const int j = row - pl * nj;
  const int blk = pl / nk;
  const int k = pl - blk * nk;
  f(blk, kl + k, jl + j, il + i);
}

template <class F>
static inline cudaError_t launch_points(cudaStream_t st, int nb, const Range3 &r, const F &f,
                                        int threads = ATHK_KP_THREADS) {
  if (threads > ATHK_KP_THREADS) threads = ATHK_KP_THREADS;
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  const long long total = (long long)nb * nk * nj * ni;
  const long long grid = (total + threads - 1) / threads;
  k_points<F><<<(unsigned)grid, threads, 0, st>>>(f, nb, r.kl, nk, r.jl, nj, r.il, ni);
  return cudaGetLastError();
}

// x1 sweep: one lane per face, warps advance by 31 faces so that lane L+1 always holds the
// flux of the next face of lane L's cell (one redundant face per warp, 3%).
#ifndef ATHK_X1_MINB
#define ATHK_X1_MINB 5
#endif
template <class F>
__global__ void __launch_bounds__(128, ATHK_X1_MINB) k_sweep_x1(const F f, const int nb, const int kl,
                                                  const int nk, const int jl, const int nj,
                                                  const int il, const int nif) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)nb * nk * nj * nif;
  const long long s = warp * 31 + lane;
  const bool valid = s < total;
  int blk = 0, k = 0, j = 0, fi = 0;
  double fl[7];
#pragma unroll
  for (int v = 0; v < 7; ++v) fl[v] = 0.0;
  if (valid) {
    const long long row = s / nif;
    fi = (int)(s - row * nif);
    const int pl = (int)(row / nj);
    j = (int)(row - (long long)pl * nj);
    blk = pl / nk;
    k = pl - blk * nk;
    f.compute(blk, kl + k, jl + j, il + fi, fl, lane < 31 || s == total - 1);
  }
  double fn[7];
#pragma unroll
  for (int v = 0; v < 5; ++v) fn[v] = __shfl_down_sync(0xffffffffu, fl[v], 1);
  if (valid && lane < 31 && fi + 1 < nif && f.cell_active(kl + k, jl + j, il + fi))
    f.store_d1(blk, kl + k, jl + j, il + fi, fl, fn);
}

template <class F>
static inline cudaError_t launch_sweep_x1(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  const int nif = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  const long long total = (long long)nb * nk * nj * nif;
  if (total <= 0) return cudaSuccess;
  long long nwarps = (total - 1 + 30) / 31;
  if (nwarps < 1) nwarps = 1;
  const long long grid = (nwarps + 3) / 4;
  k_sweep_x1<F><<<(unsigned)grid, 128, 0, st>>>(f, nb, r.kl, nk, r.jl, nj, r.il, nif);
  return cudaGetLastError();
}
#else
template <class F>
static inline cudaError_t launch_points(cudaStream_t, int nb, const Range3 &r, const F &f,
                                        int = 256) {
  for (int b = 0; b < nb; ++b)
    for (int k = r.kl; k <= r.ku; ++k)
      for (int j = r.jl; j <= r.ju; ++j)
        for (int i = r.il; i <= r.iu; ++i) f(b, k, j, i);
  return 0;
}
template <class F>
static inline cudaError_t launch_sweep_x1(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  return launch_points(st, nb, r, f);
}
#endif

}  // namespace athk
#endif  // ATHK_BACKEND_H_
