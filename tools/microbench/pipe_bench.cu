// Microbenchmarks of the sm_100a FP32 pipe at the occupancy the frame-pair kernel runs at (1-4 warps per
// scheduler): FFMA2 latency / issue cadence, scalar-broadcast operand form, co-issue with ALU and LDS.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int ILP, int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  extern __shared__ float4 sm[];
  float2 v[ILP];
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  int acc = threadIdx.x;
  float4 l = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) v[i] = __ffma2_rn(v[i], aa, bb);                       // packed, pair operands
        if (MODE == 1) v[i] = __ffma2_rn(v[i], make_float2(a, a), bb);        // packed, scalar broadcast
        if (MODE == 2) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }   // scalar
        if (MODE == 3) {                                                      // packed + 1 ALU op each
          v[i] = __ffma2_rn(v[i], aa, bb);
          acc = (acc ^ (acc >> 3)) + i;
        }
        if (MODE == 4) {                                                      // packed + LDS.128 every 4th
          v[i] = __ffma2_rn(v[i], aa, bb);
          if ((i & 3) == 0) { const float4 t = sm[(threadIdx.x + i * 32 + r * 7) & 1023]; l.x += t.x; }
        }
        if (MODE == 5) {                                                      // packed + 2 ALU ops each
          v[i] = __ffma2_rn(v[i], aa, bb);
          acc = (acc ^ (acc >> 3)) + i;
          acc = (acc & 0x7fffff) | (i << 24);
        }
      }
    }
  }
  float s = l.x + acc;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, int MODE>
void run(const char* name, float* out, int warps_per_sm) {
  const int threads = warps_per_sm * 32, blocks = 148;
  cudaFuncSetAttribute(k<ILP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, MODE><<<blocks, threads, 16384>>>(out, 1.0001f, 0.5f, 64);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<ILP, MODE><<<blocks, threads, 16384>>>(out, 1.0001f, 0.5f, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  const double packed_per_warp = (double)ITERS * 8 * ILP;
  // cycles per packed op (or per scalar pair) per scheduler: warps_per_sm/4 warps share a scheduler
  const double per_sched = cycles / (packed_per_warp * warps_per_sm / 4.0);
  printf("%-34s ILP %d warps/sched %d: %.2f cyc per pair-op per scheduler (%.2f per warp)\n", name, ILP,
         warps_per_sm / 4, per_sched, cycles / packed_per_warp);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  for (int w : {4, 8, 16}) {
    run<1, 0>("FFMA2 pair operands", out, w);
    run<2, 0>("FFMA2 pair operands", out, w);
    run<4, 0>("FFMA2 pair operands", out, w);
    run<8, 0>("FFMA2 pair operands", out, w);
    run<8, 1>("FFMA2 scalar-broadcast operand", out, w);
    run<1, 2>("2 x FFMA", out, w);
    run<8, 2>("2 x FFMA", out, w);
    run<8, 3>("FFMA2 + 1 ALU", out, w);
    run<8, 5>("FFMA2 + 2 ALU", out, w);
    run<8, 4>("FFMA2 + LDS.128 per 4", out, w);
  }
  return 0;
}
