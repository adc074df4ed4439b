// How fast do the register-resident butterfly stages run with no memory traffic at all?
// Each warp repeats stages 2-5 (compile-time twiddles) and 5 general-twiddle stages on 32 packed complex
// points.  Reports FP32-pipe cycles per iteration per scheduler against the packed-op count.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../spectrogram_b200/csrc -o core_bench core_bench.cu
#include <cstdio>
#include "kernel_w32x2p.cuh"

using namespace sg;

template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k(float* out, const float2* tw, int iters) {
  C2 a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    a[i].re = P2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    a[i].im = P2(threadIdx.x * 3e-3f + i, threadIdx.x * 4e-3f - i);
  }
  float2 b[5];
#pragma unroll
  for (int u = 0; u < 5; ++u) b[u] = tw[u * 32 + (threadIdx.x & 31)];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      dit2_stage_const<2>(a); dit2_stage_const<3>(a); dit2_stage_const<4>(a); dit2_stage_const<5>(a);
    }
    if (MODE == 1 || MODE == 2) {
      dit2_stage_gen<1>(a, b[0]); dit2_stage_gen<2>(a, b[1]); dit2_stage_gen<3>(a, b[2]);
      dit2_stage_gen<4>(a, b[3]); dit2_stage_gen<5>(a, b[4]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i].re.v.x + a[i].re.v.y + a[i].im.v.x + a[i].im.v.y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NW>
void run(const char* name, float* out, const float2* tw, double packed_ops, double scalar_ops) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  k<MODE, NW><<<148, NW * 32>>>(out, tw, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE, NW><<<148, NW * 32>>>(out, tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3 / iters;        // cycles per iteration (all warps concurrently)
  const double per_sched = cyc / (NW / 4.0);               // cycles per warp-iteration per scheduler
  const double ideal = 2 * packed_ops + scalar_ops;
  printf("%-22s warps/SM %2d: %.0f cyc/iter, %.0f per warp-iter per scheduler, ideal %.0f -> pipe %.1f%%\n", name, NW,
         cyc, per_sched, ideal, 100 * ideal / per_sched);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  float2* tw; cudaMalloc(&tw, 5 * 32 * 8);
  float2 h[160];
  for (int i = 0; i < 160; ++i) h[i] = make_float2(0.8f, -0.6f);
  cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
  // stages 2-5: FADD2 120 + FFMA2 204 ; stages 6-10: FFMA2 480 + twiddle generation 44 scalar
  run<0, 4>("const stages 2-5", out, tw, 324, 0);
  run<0, 8>("const stages 2-5", out, tw, 324, 0);
  run<0, 12>("const stages 2-5", out, tw, 324, 0);
  run<1, 4>("general stages 6-10", out, tw, 480, 44);
  run<1, 8>("general stages 6-10", out, tw, 480, 44);
  run<1, 12>("general stages 6-10", out, tw, 480, 44);
  run<2, 8>("both", out, tw, 804, 44);
  run<2, 12>("both", out, tw, 804, 44);
  return 0;
}
