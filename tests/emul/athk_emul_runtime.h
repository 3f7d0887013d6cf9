// TEST INFRASTRUCTURE (tests/emul): minimal stand-ins for the CUDA runtime calls the API layer of
// athena_b200/csrc uses, so that the operator bodies can be compiled for the host (g++
// -DATHK_EMULATE) and their index logic checked against the reference without a GPU.
#ifndef ATHK_EMUL_RUNTIME_H_
#define ATHK_EMUL_RUNTIME_H_
// ---- minimal stand-ins for the CUDA runtime calls the API layer uses ----------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
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
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return 0; }   // (every "device" is the host)
#endif  // ATHK_EMUL_RUNTIME_H_
