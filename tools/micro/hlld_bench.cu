// Development micro-benchmark (round 2): the HLLD face solve in isolation -- FP64 pipe / issue
// efficiency as a function of the register budget and of how divisions / square roots are issued.
//   Math policies:  IEEE  = the compiler's a/b, sqrt (fast path + BSSY/CALL slow path per operation)
//                   SL    = the same Newton sequences written straight-line (no branch per
//                           operation; one "needs the slow path" flag per face)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --fmad=false -I../../athena_b200/csrc
//        hlld_bench.cu -o hlld_bench   (add -DFMAD for the contracted build: compile with --fmad=true)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "athk_physics.h"

using namespace athk;

template <class M, int THREADS, int MINB, bool PLM>
__global__ void __launch_bounds__(THREADS, MINB)
k_faces(const Real *__restrict__ st, int nstate, int nface, Real *out, EosParams eos) {
  const int tid = blockIdx.x * THREADS + threadIdx.x;
  const int stride = gridDim.x * THREADS;
  Real acc = 0.0;
  for (int f = tid; f < nface; f += stride) {
    const int s = f & (nstate - 1);
    Real wl[7], wr[7];
    M m;
    if (PLM) {
      Real q0[7], q1[7], q2[7], q3[7], t0[7], t1[7];
#pragma unroll
      for (int v = 0; v < 7; ++v) {
        q0[v] = st[(v)*nstate + ((s + nstate - 2) & (nstate - 1))];
        q1[v] = st[(v)*nstate + ((s + nstate - 1) & (nstate - 1))];
        q2[v] = st[(v)*nstate + s];
        q3[v] = st[(v)*nstate + ((s + 1) & (nstate - 1))];
      }
#pragma unroll
      for (int v = 0; v < 7; ++v) {
        Real d1 = vl_slope_m(m, q1[v] - q0[v], q2[v] - q1[v]);
        Real d2 = vl_slope_m(m, q2[v] - q1[v], q3[v] - q2[v]);
        wl[v] = q1[v] + 0.5 * d1; t0[v] = q1[v] - 0.5 * d1;
        wr[v] = q2[v] - 0.5 * d2; t1[v] = q2[v] + 0.5 * d2;
      }
      (void)t0; (void)t1;
    } else {
#pragma unroll
      for (int v = 0; v < 7; ++v) {
        wl[v] = st[v * nstate + s];
        wr[v] = st[(7 + v) * nstate + s];
      }
    }
    const Real bx = st[14 * nstate + s];
    Real fl[7];
    hlld_flux_m(m, eos, wl, wr, bx, fl);
    if (!m.ok()) {   // an operand left the straight-line sequences' domain: redo with the compiler's operations
      MathIEEE mi;
      hlld_flux_m(mi, eos, wl, wr, bx, fl);
    }
#pragma unroll
    for (int v = 0; v < 7; ++v) acc += fl[v];
  }
  out[tid] = acc;
}

static std::vector<Real> make_states(int n) {
  std::vector<Real> h((size_t)15 * n);
  srand(12345);
  auto rnd = []() { return rand() / (double)RAND_MAX - 0.5; };
  for (int s = 0; s < n; ++s) {
    const double ph = 6.283185307179586 * s / n;
    double base[7] = {1.0 + 0.3 * sin(ph), 0.4 * cos(ph), 0.3 * sin(2 * ph), -0.2 * cos(3 * ph),
                      0.6 + 0.2 * cos(ph), 0.8 * sin(ph + 1.0), 0.5 * cos(2 * ph)};
    for (int v = 0; v < 7; ++v) {
      h[(size_t)v * n + s] = base[v] + 0.01 * rnd() * (v == 0 || v == 4 ? 0.1 : 1.0);
      h[(size_t)(7 + v) * n + s] = base[v] + 0.01 * rnd() * (v == 0 || v == 4 ? 0.1 : 1.0);
    }
    h[(size_t)14 * n + s] = 0.7 + 0.2 * sin(ph);
  }
  return h;
}

template <class M, int THREADS, int MINB, bool PLM>
static double run(const char *name, const Real *d_st, int nstate, int nface, Real *d_out, std::vector<Real> *res) {
  EosParams eos = make_eos(5.0 / 3.0, 1e-20, 1e-20, 1e-20);
  int nblk = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, k_faces<M, THREADS, MINB, PLM>, THREADS, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k_faces<M, THREADS, MINB, PLM>);
  const int grid = 148 * MINB * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_faces<M, THREADS, MINB, PLM><<<grid, THREADS>>>(d_st, nstate, nface, d_out, eos);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k_faces<M, THREADS, MINB, PLM><<<grid, THREADS>>>(d_st, nstate, nface, d_out, eos);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  cudaError_t e = cudaGetLastError();
  res->resize((size_t)grid * THREADS);
  cudaMemcpy(res->data(), d_out, res->size() * sizeof(Real), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (Real x : *res) sum += x;
  printf("%-28s regs=%3d local=%4zu CTAs/SM=%d warps/SM=%2d  %7.3f ms  %6.2f Gface/s  sum=%.17g %s\n", name, fa.numRegs,
         (size_t)fa.localSizeBytes, nblk, nblk * THREADS / 32, ms, nface / ms * 1e-6, sum,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  return sum;
}

// random-operand check of the straight-line sequences against the compiler's IEEE operations
__global__ void k_check(unsigned long long seed, int iters, unsigned long long *mism) {
  unsigned long long x = seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL;
  unsigned long long bad_div = 0, bad_sqrt = 0, bad_rcp = 0, flagged = 0;
  for (int it = 0; it < iters; ++it) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    unsigned long long y = x * 0xD1342543DE82EF95ULL + 1;
    // exponents within +-40 of 1.0 (the range physics lives in), random mantissas and signs
    const int ea = 1023 + (int)((x >> 52) % 81) - 40, eb = 1023 + (int)((y >> 52) % 81) - 40;
    const double a = __longlong_as_double((long long)((x & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)ea << 52)));
    const double b = __longlong_as_double((long long)((y & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)eb << 52)));
    MathSL m;
    const double q = m.div(a, b), r = m.rcp(b), s = m.sqrt(fabs(a));
    if (!m.ok()) ++flagged;
    if (__double_as_longlong(q) != __double_as_longlong(a / b)) ++bad_div;
    if (__double_as_longlong(r) != __double_as_longlong(1.0 / b)) ++bad_rcp;
    if (__double_as_longlong(s) != __double_as_longlong(sqrt(fabs(a)))) ++bad_sqrt;
  }
  if (bad_div) atomicAdd(&mism[0], bad_div);
  if (bad_rcp) atomicAdd(&mism[1], bad_rcp);
  if (bad_sqrt) atomicAdd(&mism[2], bad_sqrt);
  if (flagged) atomicAdd(&mism[3], flagged);
}

int main(int argc, char **argv) {
  const int nstate = 1 << 16;
  const int nface = argc > 1 ? atoi(argv[1]) : (1 << 24);
  std::vector<Real> h = make_states(nstate);
  Real *d_st, *d_out;
  cudaMalloc(&d_st, h.size() * sizeof(Real));
  cudaMemcpy(d_st, h.data(), h.size() * sizeof(Real), cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, sizeof(Real) * 148 * 8 * 2 * 128);
  {
    unsigned long long *d_m, hm[4] = {0, 0, 0, 0};
    cudaMalloc(&d_m, sizeof(hm));
    cudaMemset(d_m, 0, sizeof(hm));
    k_check<<<148 * 8, 256>>>(0x1234567ULL, 20000, d_m);
    cudaMemcpy(hm, d_m, sizeof(hm), cudaMemcpyDeviceToHost);
    printf("straight-line vs IEEE over %.2e random operand pairs: div mismatches %llu, rcp %llu, sqrt %llu, flagged %llu\n",
           148.0 * 8 * 256 * 20000, hm[0], hm[1], hm[2], hm[3]);
  }
  std::vector<Real> r0, r1;
#define BOTH(T, B, P)                                                                        \
  {                                                                                          \
    double s0 = run<MathIEEE, T, B, P>(P ? "IEEE " #T "x" #B " plm" : "IEEE " #T "x" #B " donor", d_st, nstate, nface, d_out, &r0); \
    double s1 = run<MathSL, T, B, P>(P ? "SL   " #T "x" #B " plm" : "SL   " #T "x" #B " donor", d_st, nstate, nface, d_out, &r1);   \
    size_t nd = 0;                                                                           \
    if (r0.size() == r1.size()) { for (size_t q = 0; q < r0.size(); ++q) nd += (r0[q] != r1[q]); } else nd = (size_t)-1; \
    printf("   -> bitwise differences between IEEE and SL partial sums: %zu %s\n", nd, s0 == s1 ? "" : "(sums differ)"); \
  }
  BOTH(128, 8, false)
  BOTH(128, 6, false)
  BOTH(128, 5, false)
  BOTH(128, 4, false)
  BOTH(128, 3, false)
  BOTH(128, 2, false)
  BOTH(64, 4, false)
  BOTH(128, 6, true)
  BOTH(128, 5, true)
  BOTH(128, 4, true)
  BOTH(128, 3, true)
  BOTH(128, 2, true)
  return 0;
}
