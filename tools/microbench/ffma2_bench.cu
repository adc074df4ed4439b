// Microbenchmark: scalar FFMA vs packed FFMA2 issue rate on sm_100a (is FFMA2 2 FMAs per issue slot?).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_scalar(float* out, float a, float b, int iters) {
  float v[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_packed(float* out, float a, float b, int iters) {
  float2 v[CH];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = __ffma2_rn(v[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps_per_sm : {8, 16, 32}) {
    const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
    constexpr int CH = 8;
    float ms_s = time_it([&] { k_scalar<CH><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    float ms_p = time_it([&] { k_packed<CH><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    double fma_s = (double)blocks * threads * iters * 2 * CH, fma_p = fma_s;
    printf("warps/SM %2d: scalar %.3f ms %.2f TFMA/s | packed %.3f ms %.2f TFMA/s | speedup %.2fx\n", warps_per_sm,
           ms_s, fma_s / ms_s / 1e9, ms_p, fma_p / ms_p / 1e9, ms_s / ms_p);
  }
  printf("peak scalar FMA/s at %d kHz: %.2f T\n", clk, 148.0 * 128 * clk * 1e3 / 1e12);
  return 0;
}
