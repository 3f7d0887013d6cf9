// Development micro-benchmark: FP64 instruction rate of one B200 as a function of resident warps
// and per-thread ILP (independent DFMA chains), i.e. the latency/throughput surface the FP64-bound
// sweep kernels live on.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double g_best = 0.0;
static int g_best_ilp = 0, g_best_warps = 0;

template <int ILP>
void run(int warps_per_sm) {
  const int iters = 8192 / ILP * 4;
  int threads = 32 * warps_per_sm, blocks = 148;
  if (threads > 1024) { threads /= 2; blocks *= 2; }
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<ILP><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e0);
  k<ILP><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)blocks * threads * iters * ILP;
  const double rate = inst / ms * 1e-9;
  printf("ILP=%d warps/SM=%2d  %6.2f T thread-inst/s\n", ILP, warps_per_sm, rate);
  if (rate > g_best) { g_best = rate; g_best_ilp = ILP; g_best_warps = warps_per_sm; }
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 12, 16, 20, 32, 64}) {
    run<1>(w);
    run<2>(w);
    run<4>(w);
    run<8>(w);
  }
  // one JSON line for profiles/r2_fp64_peak.json (bench.py reads best_t_thread_inst_per_s)
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("JSON {\"best_t_thread_inst_per_s\": %.3f, \"ilp\": %d, \"warps_per_sm\": %d, \"device\": \"%s\", "
         "\"sms\": %d, \"what\": \"DFMA thread-instructions per second, independent chains (tools/micro/fp64_peak.cu)\"}\n",
         g_best, g_best_ilp, g_best_warps, pr.name, pr.multiProcessorCount);
  return 0;
}
