// Headline kernel: n_fft = 2048 (the reference's fftSize, UI/player.js:10), one warp per frame.
//
// The 2048 real samples of a frame are packed as 1024 complex points (z[n] = xw[2n] + i xw[2n+1]);
// the 1024-point FFT is a pure radix-2 decimation-in-time network split 5 + 5:
//   pass 1  lane b holds z[b + 32 j], j < 32, and runs stages 1-5 (a 32-point FFT) entirely in
//           registers; every twiddle is a compile-time immediate
//   xchg    32x32 transpose through the warp's shared-memory tile (row stride 34 float2:
//           128-bit stores and 64-bit loads are both bank-conflict free)
//   pass 2  lane a holds element (q, a), q < 32, and runs stages 6-10; the 31 lane-dependent
//           twiddles W_{64..1024} come from a shared-memory table
//   untangle real-FFT split X[k] = E[k] + W_2048^k O[k]: each lane keeps Z[k] for k < 512 in
//           registers and fetches Z[1024-k] through shared memory (only the upper half moves)
//   epilogue |X|^2 -> lg2 -> affine -> clamp -> byte (or dB / RGBA / linear magnitude), bytes
//           staged in shared memory so a frame's 1024-byte row leaves as 2 x 512-byte stores
// Frames are read straight from HBM/L2 as coalesced 8-byte loads (256 B per warp instruction);
// the 4x overlap between consecutive frames is absorbed by L1/L2 because consecutive frames are
// processed by neighbouring warps at the same time.
//
// Roofline: algorithmic HBM bytes per frame = 4*hop + elem*1024 (3072 B for u8 at
// hop 512); the kernel is co-bound by the FP32 pipe (~1.25 K lane-FMA-pipe instructions per lane
// per frame), so HBM fraction is reported together with the issue-slot budget.
#pragma once
#include "common.cuh"
#include "plans.cuh"
#include "ct_math.cuh"

namespace sg {

constexpr int kW32Stride = 34;                       // float2 per exchange-tile row
constexpr int kW32TileBytes = 32 * kW32Stride * 8;   // 8704 B per warp
constexpr int kW32Warps = 8;
constexpr int kW32TableBytes = kW32N * 4 + 31 * 32 * 8 + 16 * 32 * 8;  // window + tw2 + ut
constexpr int kW32SmemBytes = kW32TableBytes + kW32Warps * kW32TileBytes;

// x' = x + w*y ; y' = x - w*y = 2x - x'   (6 FMA-pipe instructions)
__device__ __forceinline__ void bfly(float2& x, float2& y, float wr, float wi) {
  const float xr = fmaf(wr, y.x, fmaf(-wi, y.y, x.x));
  const float xi = fmaf(wr, y.y, fmaf(wi, y.x, x.y));
  y.x = fmaf(2.f, x.x, -xr);
  y.y = fmaf(2.f, x.y, -xi);
  x.x = xr;
  x.y = xi;
}

template <int P, int N>
__device__ __forceinline__ void bfly_const(float2& x, float2& y) {
  if constexpr (P == 0) {                    // w = 1
    const float2 t = y;
    y = make_float2(x.x - t.x, x.y - t.y);
    x = make_float2(x.x + t.x, x.y + t.y);
  } else if constexpr (4 * P == N) {         // w = -i : w*y = (y.y, -y.x)
    const float2 t = y;
    y = make_float2(x.x - t.y, x.y + t.x);
    x = make_float2(x.x + t.y, x.y - t.x);
  } else {
    bfly(x, y, Twiddle<P, N>::re, Twiddle<P, N>::im);
  }
}

// (wx, wy) = W_N^P * base, reusing W^{P + N/4} = -i W^P
template <int P, int N>
__device__ __forceinline__ float2 twiddle_times(float2 b) {
  if constexpr (P == 0) {
    return b;
  } else if constexpr (4 * P >= N) {
    const float2 t = twiddle_times<P - N / 4, N>(b);
    return make_float2(t.y, -t.x);
  } else {
    constexpr float cx = Twiddle<P, N>::re, cy = Twiddle<P, N>::im;
    return make_float2(fmaf(cx, b.x, -cy * b.y), fmaf(cx, b.y, cy * b.x));
  }
}

// stages 1..5 of a radix-2 DIT network on 32 register-resident points (input bit-reversed)
template <int S>
__device__ __forceinline__ void dit_stage_const(float2 (&a)[32]) {
  constexpr int half = 1 << (S - 1);
  static_for<0, 16>([&](auto idx) {
    constexpr int i = decltype(idx)::value;
    constexpr int blk = i / half, p = i % half, i0 = blk * 2 * half + p;
    bfly_const<p, 2 * half>(a[i0], a[i0 + half]);
  });
}

template <int S>
__device__ __forceinline__ void dit_stage_table(float2 (&a)[32], const float2* __restrict__ tw_lane) {
  constexpr int half = 1 << (S - 1);
  static_for<0, half>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    const float2 w = tw_lane[(half - 1 + p) * 32];
    static_for<0, 16 / half>([&](auto bb) {
      constexpr int i0 = decltype(bb)::value * 2 * half + p;
      bfly(a[i0], a[i0 + half], w.x, w.y);
    });
  });
}

template <int OUT>
__global__ void __launch_bounds__(kW32Warps * 32, 2)
stft_w32_kernel(FrameGeom g, W32Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);                 // [1024] (pairs)
  float2* s_tw2 = s_win + kW32M;                                       // [31*32]
  float2* s_ut = s_tw2 + 31 * 32;                                      // [16*32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* xb = s_ut + 16 * 32 + warp * (32 * kW32Stride);              // this warp's tile

  for (int i = threadIdx.x; i < kW32M; i += blockDim.x)
    s_win[i] = __ldg(reinterpret_cast<const float2*>(pl.win) + i);
  for (int i = threadIdx.x; i < 31 * 32; i += blockDim.x) s_tw2[i] = __ldg(pl.tw2 + i);
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) s_ut[i] = __ldg(pl.ut + i);
  __syncthreads();

  const long long warps_total = (long long)gridDim.x * kW32Warps;
  for (long long f = (long long)blockIdx.x * kW32Warps + warp; f < g.total_frames; f += warps_total) {
    const long long clip = f / g.frames_per_clip, t = f - clip * g.frames_per_clip;
    const long long start = g.start0 + t * g.hop;
    const float* __restrict__ x = g.pcm + clip * g.clip_stride;

    // ---- steps 1-2: load the time block (coalesced 8-byte loads), window, bit-reverse into regs
    float2 a[32];
    const bool interior = start >= 0 && start + kW32N <= g.clip_len &&
                          ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
    if (interior) {
      const float2* __restrict__ src = reinterpret_cast<const float2*>(x + start) + lane;
      static_for<0, 32>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 v = __ldg(src + 32 * j);
        const float2 w = s_win[lane + 32 * j];
        a[bitrev(j, 5)] = make_float2(v.x * w.x, v.y * w.y);
      });
    } else if (start >= 0 && start + kW32N <= g.clip_len) {
      // inside the clip, only misaligned (odd hops): 4-byte loads, no bounds checks
      const float* __restrict__ src = x + start + 2 * lane;
      static_for<0, 32>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float v0 = __ldg(src + 64 * j), v1 = __ldg(src + 64 * j + 1);
        const float2 w = s_win[lane + 32 * j];
        a[bitrev(j, 5)] = make_float2(v0 * w.x, v1 * w.y);
      });
    } else {
      static_for<0, 32>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const long long s0 = start + 2 * (lane + 32 * j), s1 = s0 + 1;
        const float v0 = (s0 >= 0 && s0 < g.clip_len) ? __ldg(x + s0) : 0.f;
        const float v1 = (s1 >= 0 && s1 < g.clip_len) ? __ldg(x + s1) : 0.f;
        const float2 w = s_win[lane + 32 * j];
        a[bitrev(j, 5)] = make_float2(v0 * w.x, v1 * w.y);
      });
    }

    // ---- pass 1: stages 1-5 in registers
    dit_stage_const<1>(a);
    dit_stage_const<2>(a);
    dit_stage_const<3>(a);
    dit_stage_const<4>(a);
    dit_stage_const<5>(a);

    // ---- exchange: tile[b = lane][k_a]  ->  lane a reads tile[bitrev(q)][a]
    static_for<0, 16>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      *reinterpret_cast<float4*>(xb + lane * kW32Stride + 2 * q) =
          make_float4(a[2 * q].x, a[2 * q].y, a[2 * q + 1].x, a[2 * q + 1].y);
    });
    __syncwarp();
    static_for<0, 32>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      a[q] = xb[bitrev(q, 5) * kW32Stride + lane];
    });
    __syncwarp();

    // ---- pass 2: stages 6-10, twiddles W_{32*2^u}^{32 p + lane}
    const float2* tw_lane = s_tw2 + lane;
    dit_stage_table<1>(a, tw_lane);
    dit_stage_table<2>(a, tw_lane);
    dit_stage_table<3>(a, tw_lane);
    dit_stage_table<4>(a, tw_lane);
    dit_stage_table<5>(a, tw_lane);
    // now a[i] = Z[lane + 32 i]

    // ---- untangle exchange: the upper half (k >= 512) goes to smem at index k - 512;
    //      Z[1024] == Z[0] is stored at index 512 by lane 0
    static_for<16, 32>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      xb[lane + 32 * (i - 16)] = a[i];
    });
    if (lane == 0) xb[512] = a[0];
    __syncwarp();

    T* __restrict__ row = out + f * (long long)kW32M;
    unsigned char* sb = reinterpret_cast<unsigned char*>(xb + 520);   // byte staging, past the 513 entries
    static_for<0, 16>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int k = lane + 32 * i;
      int mk = kW32M - k;
      const float2 zk = a[i];
      const float2 zm = xb[512 - k];
      const float2 w = s_ut[i * 32 + lane];
      const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
      const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
      const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
      const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
      const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[1024-k]
      const float yi = fmaf(2.f, ey, -xi);
      const float pk = fmaf(xr, xr, xi * xi);
      float pm = fmaf(yr, yr, yi * yi);
      if constexpr (i == 0) {
        // lane 0: the mirror of k = 0 is the Nyquist bin (dropped); use the slot for bin 512,
        // X[512] = conj Z[512], held by lane 0 in a[16]
        if (lane == 0) {
          mk = 512;
          pm = 4.f * fmaf(a[16].x, a[16].x, a[16].y * a[16].y);
        }
      }
      if constexpr (OUT == kOutU8) {
        sb[k] = emit_power<OUT>(pk, ep);
        sb[mk] = emit_power<OUT>(pm, ep);
      } else {
        row[k] = emit_power<OUT>(pk, ep);
        row[mk] = emit_power<OUT>(pm, ep);
      }
    });
    if constexpr (OUT == kOutU8) {
      __syncwarp();
      const uint4 v0 = reinterpret_cast<const uint4*>(sb)[lane];
      const uint4 v1 = reinterpret_cast<const uint4*>(sb)[32 + lane];
      uint4* row16 = reinterpret_cast<uint4*>(row);
      row16[lane] = v0;
      row16[32 + lane] = v1;
    }
    __syncwarp();
  }
}

}  // namespace sg
