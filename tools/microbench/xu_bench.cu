// Which pipe do the epilogue / untangle helpers of the frame-pair kernel run on, and how fast?
// Streams of MUFU.LG2, F2IP (cvt.rzi.u8.f32), SHFL, PRMT, FMNMX, FSEL -- alone and interleaved 1:1 with FFMA2 or with
// each other -- at 1-3 warps per scheduler.  Two streams on different pipes cost max(a, b) per pair of instructions,
// on the same pipe a + b.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o xu_bench xu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
constexpr int ILP = 8;

__device__ __forceinline__ float lg2a(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned b8(float v) { unsigned b; asm volatile("cvt.rzi.u8.f32 %0, %1;" : "=r"(b) : "f"(v)); return b; }

enum { M_LG2, M_F2IP, M_SHFL, M_PRMT, M_FMNMX, M_FSEL, M_FFMA2, M_LG2_F2IP, M_LG2_FFMA2, M_F2IP_FFMA2, M_SHFL_FFMA2,
       M_PRMT_FFMA2, M_FSEL_FFMA2, M_LG2_SHFL, M_F2IP_PRMT, M_COUNT };
const char* kNames[M_COUNT] = {"MUFU.LG2", "F2IP(+LOP)", "SHFL", "PRMT", "FMNMX", "FSEL", "FFMA2", "LG2+F2IP(+LOP)", "LG2+FFMA2",
                               "F2IP(+LOP)+FFMA2", "SHFL+FFMA2", "PRMT+FFMA2", "FSEL+FFMA2", "LG2+SHFL", "F2IP(+LOP)+PRMT"};

template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float x[ILP];
  float2 v[ILP];
  unsigned u[ILP];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = 1.5f + threadIdx.x * 0.01f + i;
    v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    u[i] = threadIdx.x * 2654435761u + i;
  }
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        constexpr bool lg = MODE == M_LG2 || MODE == M_LG2_F2IP || MODE == M_LG2_FFMA2 || MODE == M_LG2_SHFL;
        constexpr bool f2 = MODE == M_F2IP || MODE == M_LG2_F2IP || MODE == M_F2IP_FFMA2 || MODE == M_F2IP_PRMT;
        constexpr bool sh = MODE == M_SHFL || MODE == M_SHFL_FFMA2 || MODE == M_LG2_SHFL;
        constexpr bool pr = MODE == M_PRMT || MODE == M_PRMT_FFMA2 || MODE == M_F2IP_PRMT;
        constexpr bool fs = MODE == M_FSEL || MODE == M_FSEL_FFMA2;
        constexpr bool fm = MODE == M_FFMA2 || MODE == M_LG2_FFMA2 || MODE == M_F2IP_FFMA2 || MODE == M_SHFL_FFMA2 ||
                            MODE == M_PRMT_FFMA2 || MODE == M_FSEL_FFMA2;
        if (lg) x[i] = lg2a(x[i]);
        if (f2) u[i] = b8(__uint_as_float(u[i] | 0x42000000u));       // F2IP + one LOP3 per link of the chain
        if (sh) v[i].x = __shfl_sync(0xffffffffu, v[i].x, (32 - lane) & 31);
        if (pr) u[i] = __byte_perm(u[i], u[(i + 1) % ILP], 0x2140);
        if (MODE == M_FMNMX) x[i] = fminf(v[i].y, x[(i + 1) % ILP]);
        if (fs) v[i].y = (u[i] & 1) ? v[(i + 1) % ILP].y : x[i];                   // predicate is loop invariant: FSEL only
        if (fm) v[i] = __ffma2_rn(v[i], aa, bb);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + v[i].x + v[i].y + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(float* out) {
  for (int warps_per_sm : {4, 8, 12}) {
    const int threads = warps_per_sm * 32, blocks = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, 16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double links = (double)ITERS * 4 * ILP;   // chain links per warp (each link = one of every stream's instructions)
    printf("%-20s warps/sched %d: %.2f cyc per link per scheduler\n", kNames[MODE], warps_per_sm / 4,
           cycles / (links * warps_per_sm / 4.0));
  }
}

template <int M>
void run_all(float* out) {
  if constexpr (M < M_COUNT) { run<M>(out); run_all<M + 1>(out); }
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  run_all<0>(out);
  return 0;
}
