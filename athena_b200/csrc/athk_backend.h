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
// One thread per point; x covers i (unit stride => coalesced 128-byte rows), y covers j,
// z covers (block, k).  Grid sized to the exact point count; blocks of 256 threads.
template <class F>
__global__ void __launch_bounds__(256) k_points(const F f, const int nb, const int kl, const int nk,
                                                const int jl, const int nj, const int il,
                                                const int ni) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int kz = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= ni || j >= nj || kz >= nk * nb) return;
  const int blk = kz / nk;
  const int k = kz - blk * nk;
  f(blk, kl + k, jl + j, il + i);
}

template <class F>
static inline cudaError_t launch_points(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0 || nb <= 0) return cudaSuccess;
  int bx = 32;
  while (bx < ni && bx < 128) bx <<= 1;   // 32, 64 or 128 threads along i
  int by = 256 / bx;
  if (by > nj) by = nj;
  int bz = 256 / (bx * by);
  if (bz < 1) bz = 1;
  dim3 block(bx, by, bz);
  dim3 grid((ni + bx - 1) / bx, (nj + by - 1) / by, (nk * nb + bz - 1) / bz);
  k_points<F><<<grid, block, 0, st>>>(f, nb, r.kl, nk, r.jl, nj, r.il, ni);
  return cudaGetLastError();
}
#else
template <class F>
static inline cudaError_t launch_points(cudaStream_t, int nb, const Range3 &r, const F &f) {
  for (int b = 0; b < nb; ++b)
    for (int k = r.kl; k <= r.ku; ++k)
      for (int j = r.jl; j <= r.ju; ++j)
        for (int i = r.il; i <= r.iu; ++i) f(b, k, j, i);
  return 0;
}
#endif

}  // namespace athk
#endif  // ATHK_BACKEND_H_
