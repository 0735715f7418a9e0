// Micro-benchmark: does packed FFMA2 (fma.rn.f32x2, sm_100) save issue slots in a mixed FP32 / ALU instruction stream?
// Four kernels with the same number of scalar FMAs per thread: (a) FFMA only, (b) FFMA2 only, (c) FFMA + one ALU op per
// FMA, (d) FFMA2 + the same ALU ops.  If (d) is faster than (c), packing two environments per thread pays for an
// issue-bound kernel like k_env_step32.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[4];
  for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 4);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = __ffma2_rn(x[i], A, B);
  }
  float s = 0;
  for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(float* out, float a, float b, float lo) {
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = fmaf(x[i], a, b); if (i < 4) x[i] = fmaxf(x[i], lo); }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix2(float* out, float a, float b, float lo) {
  float2 x[4];
  for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 4);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = __ffma2_rn(x[i], A, B); if (i < 2) { x[i].x = fmaxf(x[i].x, lo); x[i].y = fmaxf(x[i].y, lo); } }
  }
  float s = 0;
  for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mixi(float* out, float a, float b, int m) {
  float x[8]; int y[8];
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = fmaf(x[i], a, b); y[i] = (y[i] ^ m) + it; }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixi2(float* out, float a, float b, int m) {
  float2 x[4]; int y[8];
  for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 4);
  for (int i = 0; i < 8; i++) y[i] = threadIdx.x + i;
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = __ffma2_rn(x[i], A, B);
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = (y[i] ^ m) + it;
  }
  float s = 0;
  for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
  for (int i = 0; i < 8; i++) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int G = 148 * 8, B = 256;
  const double fmas = (double)G * B * ITERS * 8;
  float t;
  t = timeit([&] { k_ffma<<<G, B>>>(out, 0.999f, 0.001f); });  printf("FFMA only      : %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  t = timeit([&] { k_ffma2<<<G, B>>>(out, 0.999f, 0.001f); }); printf("FFMA2 only     : %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  t = timeit([&] { k_mix<<<G, B>>>(out, 0.999f, 0.001f, -1.f); });  printf("8 FFMA + 4 FMNMX : %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  t = timeit([&] { k_mix2<<<G, B>>>(out, 0.999f, 0.001f, -1.f); }); printf("4 FFMA2 + 4 FMNMX: %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  t = timeit([&] { k_mixi<<<G, B>>>(out, 0.999f, 0.001f, 0x55); });  printf("FFMA + 2 int   : %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  t = timeit([&] { k_mixi2<<<G, B>>>(out, 0.999f, 0.001f, 0x55); }); printf("FFMA2 + 2 int  : %.3f ms  %.1f TFMA/s\n", t, fmas / t * 1e-9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
