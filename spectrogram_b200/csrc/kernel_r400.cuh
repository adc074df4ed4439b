// n_fft = 400 (BASELINE config 4: 16 kHz speech front end, hop 160).  AnalyserNode itself rejects this size
// (fftSize must be a power of two -- Web Audio IndexSizeError); the batched path follows the same formulas
//  and this kernel gives it a register-resident FFT instead of the generic shared-memory one.
//
// 400 real samples pack into M = 200 complex points, M = 40 x 5:
//   5 threads per frame, 6 frames per warp (lanes 30, 31 shadow lanes 28, 29 and store nothing)
//   pass 1   thread b holds z[b + 5 j], j < 40, and runs a 40-point FFT (8 x 5, every twiddle a compile-time
//            immediate) entirely in registers, then multiplies by W_200^{b k1}
//   xchg     a [frame][b][k1] tile in shared memory (row stride 41 float2: conflict free for the half-warps)
//   pass 2   thread c takes 8 of the 40 columns k1 and runs the 5-point DFT over b: Z[k1 + 40 k2].  Columns are
//            dealt in mirror pairs (k1, 40 - k1), so Z[k] and Z[200 - k] always meet in the SAME thread and the
//            real-input untangle needs no second exchange (thread 4 also owns the self-mirrored columns 0, 20)
//   epilogue |X|^2 -> dB / byte / colour, staged per frame in the idle tile, stored as 16-byte coalesced rows
// Algorithmic bytes per frame: 4*hop + elem*200 (1440 B for float dB at hop 160).
#pragma once
#include "common.cuh"
#include "ct_math.cuh"
#include "kernel_smem.cuh"   // dft_small<4>, dft_small<5>
#include "plans.cuh"

namespace sg {

constexpr int kR4N = 400, kR4M = 200;
constexpr int kR4Warps = 8;
constexpr int kR4TileStride = 49;                         // float2 per (frame, b) row: 98 words = 2 mod 32, so the 16 lanes
                                                          // of a half-warp hit 32 distinct banks writing AND reading
constexpr int kR4TwStride = 41;                           // W_200^{b k1} rows
constexpr int kR4TileF2 = 30 * kR4TileStride;             // per warp: 11760 B
constexpr int kR4TableF2 = (kR4M + 5 * kR4TwStride + kR4M + 1) & ~1;   // window pairs + W_200^{b k1} + W_400^k, 16-byte padded
constexpr int kR4OutStrideW = 228;                        // 32-bit outputs: words per staged frame row (16-byte aligned)
constexpr int kR4OutStrideB = 208;                        // u8 outputs: bytes per staged frame row
constexpr int kR4SmemBytes = (kR4TableF2 + kR4Warps * kR4TileF2) * 8;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void fft8(float2 (&x)[8]) {
  constexpr float h = 0.70710678118654752440f;
  float2 a[4], b[4];
  static_for<0, 4>([&](auto ii) {
    constexpr int i = decltype(ii)::value;
    a[i] = cadd(x[i], x[i + 4]);
    const float2 d = csub(x[i], x[i + 4]);
    if constexpr (i == 0) b[0] = d;
    else if constexpr (i == 1) b[1] = make_float2((d.x + d.y) * h, (d.y - d.x) * h);      // * W_8^1
    else if constexpr (i == 2) b[2] = make_float2(d.y, -d.x);                            // * -i
    else b[3] = make_float2((d.y - d.x) * h, -(d.x + d.y) * h);                           // * W_8^3
  });
  dft_small<4>(a);
  dft_small<4>(b);
  static_for<0, 4>([&](auto rr) {
    constexpr int r = decltype(rr)::value;
    x[2 * r] = a[r];
    x[2 * r + 1] = b[r];
  });
}

// x *= W_N^P with the twiddle folded to immediates
template <int P, int N>
__device__ __forceinline__ float2 mul_tw(float2 x) {
  constexpr int p = ((P % N) + N) % N;
  if constexpr (p == 0) return x;
  else if constexpr (4 * p == N) return make_float2(x.y, -x.x);
  else if constexpr (2 * p == N) return make_float2(-x.x, -x.y);
  else if constexpr (4 * p == 3 * N) return make_float2(-x.y, x.x);
  else {
    constexpr float c = Twiddle<p, N>::re, s = Twiddle<p, N>::im;
    return make_float2(fmaf(c, x.x, -s * x.y), fmaf(c, x.y, s * x.x));
  }
}

// forward 40-point DFT in registers.  In: v[j] (j = 5 j1 + j2).  Out: Y[q1 + 8 q2] at v[5 q1 + q2].
__device__ __forceinline__ void fft40(float2 (&v)[40]) {
  static_for<0, 5>([&](auto jj) {
    constexpr int j2 = decltype(jj)::value;
    float2 x[8];
    static_for<0, 8>([&](auto j1) { x[decltype(j1)::value] = v[5 * decltype(j1)::value + j2]; });
    fft8(x);
    static_for<0, 8>([&](auto qq) {
      constexpr int q1 = decltype(qq)::value;
      v[5 * q1 + j2] = mul_tw<j2 * q1, 40>(x[q1]);
    });
  });
  static_for<0, 8>([&](auto qq) {
    constexpr int q1 = decltype(qq)::value;
    float2 t[5];
    static_for<0, 5>([&](auto jj) { t[decltype(jj)::value] = v[5 * q1 + decltype(jj)::value]; });
    dft_small<5>(t);
    static_for<0, 5>([&](auto q2) { v[5 * q1 + decltype(q2)::value] = t[decltype(q2)::value]; });
  });
}

// real-input untangle of one mirror pair: zk = Z[k], zm = Z[200 - k], w = W_400^k -> |2 X[k]|^2, |2 X[200-k]|^2
__device__ __forceinline__ void untangle_pair(float2 zk, float2 zm, float2 w, float& pk, float& pm) {
  const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
  const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
  const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
  const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
  const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[200-k]
  const float yi = fmaf(2.f, ey, -xi);
  pk = fmaf(xr, xr, xi * xi);
  pm = fmaf(yr, yr, yi * yi);
}

template <int OUT>
__global__ void __launch_bounds__(kR4Warps * 32, 2)
stft_r400_kernel(FrameGeom g, R400Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);      // [200] (w[2m], w[2m+1])
  float2* s_tw = s_win + kR4M;                              // [5][41] W_200^{b k1}
  float2* s_ut = s_tw + 5 * kR4TwStride;                  // [200]   W_400^k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* tile = s_win + kR4TableF2 + warp * kR4TileF2;   // 16-byte aligned (staged rows leave as uint4)

  for (int i = threadIdx.x; i < kR4M; i += blockDim.x) {
    s_win[i] = __ldg(reinterpret_cast<const float2*>(pl.win) + i);
    s_ut[i] = __ldg(pl.ut + i);
  }
  for (int i = threadIdx.x; i < 5 * kR4TwStride; i += blockDim.x) s_tw[i] = __ldg(pl.tw + i);
  __syncthreads();

  const bool active = lane < 30;
  const int le = active ? lane : lane - 2;      // lanes 30, 31 shadow lanes 28, 29
  const int g6 = le / 5, b = le - 5 * g6;       // frame slot and thread index within the frame
  float2* my_row = tile + (g6 * 5 + b) * kR4TileStride;
  const float2* frame_rows = tile + (g6 * 5) * kR4TileStride;
  const bool special = b == 4;                  // owns the self-mirrored columns 0 and 20 as its 4th pair

  // frame index of this lane's slot, advanced incrementally (no 64-bit division in the loop)
  const long long groups = (g.total_frames + 5) / 6;
  const long long wstep = (long long)gridDim.x * kR4Warps;          // groups between a warp's iterations
  const long long fstep = 6 * wstep;                                // frames
  const long long step_clip = fstep / g.frames_per_clip, step_t = fstep - step_clip * g.frames_per_clip;
  long long grp = (long long)blockIdx.x * kR4Warps + warp;
  if (grp >= groups) return;
  long long f = grp * 6 + g6;
  long long clip = f / g.frames_per_clip, t = f - clip * g.frames_per_clip;
  float* scratch = reinterpret_cast<float*>(tile) + g6 * kR4N;      // edge frames are assembled here (tile is idle)
  for (; grp < groups; grp += wstep) {
    const bool live = active && f < g.total_frames;
    long long fc = f, cc = clip, tc = t;
    if (f >= g.total_frames) {                                      // idle slots recompute the last frame
      fc = g.total_frames - 1;
      cc = fc / g.frames_per_clip;
      tc = fc - cc * g.frames_per_clip;
    }
    const long long start = g.start0 + tc * g.hop;
    const float* __restrict__ x = g.pcm + cc * g.clip_stride;

    // ---- steps 1-2: time block, window; thread b takes z[b + 5 j].
    //   span path   hop 160 and the warp's six frames are consecutive frames of one clip: they cover ONE contiguous
    //               span of 1200 samples, which the warp reads as 16-byte coalesced loads (ten per lane instead of
    //               forty 8-byte loads touching six L1 lines each) into the idle tile, padded by 12 floats per hop so
    //               the six frames' reads fall on different banks
    //   frame path  interior, 8-byte aligned frames are read straight from global memory
    //   edge path   clip edges / zero history / odd hops are assembled zero-filled in the idle tile
    // All three feed the same unrolled loader below.
    const long long f0 = grp * 6;
    const long long clip0 = __shfl_sync(0xffffffffu, clip, 0), t00 = __shfl_sync(0xffffffffu, t, 0);
    const long long start00 = g.start0 + t00 * 160;
    const float* span_src = g.pcm + clip0 * g.clip_stride + start00;
    const bool span = g.hop == 160 && f0 + 5 < g.total_frames && t00 + 5 < g.frames_per_clip && start00 >= 0 &&
                      start00 + 5 * 160 + kR4N <= g.clip_len && ((reinterpret_cast<uintptr_t>(span_src) & 15) == 0);
    const bool interior = start >= 0 && start + kR4N <= g.clip_len && ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
    const float2* src = reinterpret_cast<const float2*>(x + start) + b;
    if (span) {
      const float4* gs = reinterpret_cast<const float4*>(span_src);
      float4* sd = reinterpret_cast<float4*>(tile);
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        const int q = lane + 32 * r;                         // float4 index in the span, 300 in all
        if (q < 300) sd[q + 3 * ((q * 1639) >> 16)] = __ldg(gs + q);   // + 3 float4 per 40 (one hop)
      }
      src = reinterpret_cast<const float2*>(tile) + 86 * g6 + b;       // (160 + 12) / 2 float2 per hop
    } else if (!interior) {
      if (start >= 0 && start + kR4N <= g.clip_len) {
        // inside the clip, only misaligned (odd hops): plain copy, eight loads in flight
        const float* __restrict__ xs = x + start;
#pragma unroll 8
        for (int i = b; i < kR4N; i += 5) scratch[i] = __ldg(xs + i);
      } else {
#pragma unroll 1
        for (int i = b; i < kR4N; i += 5) {
          const long long q = start + i;
          scratch[i] = (q >= 0 && q < g.clip_len) ? __ldg(x + q) : 0.f;
        }
      }
      src = reinterpret_cast<const float2*>(scratch) + b;
    }
    const bool staged = span || __any_sync(0xffffffffu, !interior);
    if (staged) __syncwarp();
    // the address space is spelled out where it is warp uniform (a pointer that may be shared or global compiles to
    // generic loads: three L1 wavefronts each instead of two, and the long scoreboard instead of the short one)
    float2 v[40];
    if (span) {
      const unsigned sbase = (unsigned)__cvta_generic_to_shared(src);
      static_for<0, 40>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        float2 sv;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sv.x), "=f"(sv.y) : "r"(sbase + 8u * (5 * j + (j / 16) * 6)));
        const float2 w = s_win[b + 5 * j];
        v[j] = make_float2(sv.x * w.x, sv.y * w.y);
      });
    } else if (!staged) {
      static_for<0, 40>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 sv = __ldg(src + 5 * j), w = s_win[b + 5 * j];
        v[j] = make_float2(sv.x * w.x, sv.y * w.y);
      });
    } else {
      static_for<0, 40>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 sv = src[5 * j], w = s_win[b + 5 * j];
        v[j] = make_float2(sv.x * w.x, sv.y * w.y);
      });
    }
    if (staged) __syncwarp();   // the staged samples are consumed before the tile is written

    // ---- pass 1: 40-point FFT over j, then W_200^{b k1}
    fft40(v);
    {
      const float2* twb = s_tw + b * kR4TwStride;
      static_for<0, 8>([&](auto qq) {
        constexpr int q1 = decltype(qq)::value;
        static_for<0, 5>([&](auto rr) {
          constexpr int q2 = decltype(rr)::value, k1 = q1 + 8 * q2;
          my_row[k1] = cmul(v[5 * q1 + q2], twb[k1]);
        });
      });
    }
    __syncwarp();

    // ---- pass 2: 5-point DFT over b for this thread's 8 columns (4 mirror pairs)
    float2 za[4][5], zb[4][5];
    static_for<0, 4>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      int ka = b + 1 + 5 * i, kb = 40 - ka;
      if constexpr (i == 3) { if (special) { ka = 0; kb = 20; } }
      static_for<0, 5>([&](auto bb) {
        constexpr int r = decltype(bb)::value;
        za[i][r] = frame_rows[r * kR4TileStride + ka];
        zb[i][r] = frame_rows[r * kR4TileStride + kb];
      });
      dft_small<5>(za[i]);
      dft_small<5>(zb[i]);
    });
    __syncwarp();   // the tile is free for the next group

    // ---- untangle + epilogue.  Column pair (ka, 40 - ka): bin k = ka + 40 k2 meets 200 - k = kb + 40 (4 - k2).
    // [SPEC] "non-finite -> 0", decided once per frame (a non-finite sample makes every Z of its frame non-finite)
    const bool bad = !(fabsf(za[0][0].x) <= 3.4028235e38f) || !(fabsf(za[0][0].y) <= 3.4028235e38f);
    // results are staged per frame in the (now idle) tile and leave as 16-byte coalesced stores: the six frames of
    // a warp are consecutive rows of `out`
    T* stage = reinterpret_cast<T*>(tile) + g6 * (sizeof(T) == 1 ? kR4OutStrideB : kR4OutStrideW);
    auto put = [&](int k, float p) { stage[k] = emit_power_finite<OUT>(bad ? 0.f : p, ep); };
    static_for<0, 4>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int ka = b + 1 + 5 * i;
      auto general = [&] {
        static_for<0, 5>([&](auto kk) {
          constexpr int k2 = decltype(kk)::value;
          const int k = ka + 40 * k2;
          float pk, pm;
          untangle_pair(za[i][k2], zb[i][4 - k2], s_ut[k], pk, pm);
          put(k, pk);
          put(kR4M - k, pm);
        });
      };
      if constexpr (i < 3) {
        general();
      } else {
        if (!special) {
          general();
        } else {
          // column 0 (bins 0, 40, 80, 120, 160) and column 20 (bins 20, 60, 100, 140, 180) mirror into themselves
          float pk, pm;
          untangle_pair(za[3][0], za[3][0], s_ut[0], pk, pm);   put(0, pk);               // DC; the mirror is the dropped Nyquist bin
          untangle_pair(za[3][1], za[3][4], s_ut[40], pk, pm);  put(40, pk);  put(160, pm);
          untangle_pair(za[3][2], za[3][3], s_ut[80], pk, pm);  put(80, pk);  put(120, pm);
          untangle_pair(zb[3][0], zb[3][4], s_ut[20], pk, pm);  put(20, pk);  put(180, pm);
          untangle_pair(zb[3][1], zb[3][3], s_ut[60], pk, pm);  put(60, pk);  put(140, pm);
          untangle_pair(zb[3][2], zb[3][2], s_ut[100], pk, pm); put(100, pk);
        }
      }
    });
    __syncwarp();
    {
      const long long f0 = grp * 6;                                     // first frame of this warp's group
      const int live_frames = (int)min(6LL, g.total_frames - f0);
      constexpr int kVecPerFrame = kR4M * (int)sizeof(T) / 16;           // 50 (32-bit) or 12.5 -> handled below
      if constexpr (sizeof(T) == 4) {
        const uint4* sv = reinterpret_cast<const uint4*>(tile);
        uint4* dst = reinterpret_cast<uint4*>(out + f0 * (long long)kR4M);
#pragma unroll
        for (int r = 0; r < (6 * kVecPerFrame + 31) / 32; ++r) {
          const int idx = lane + 32 * r;
          const int fr = (idx * 1311) >> 16;                            // idx / 50
          if (idx < 6 * kVecPerFrame && fr < live_frames)
            dst[idx] = sv[fr * (kR4OutStrideW / 4) + (idx - fr * kVecPerFrame)];
        }
      } else {
        // u8: 200 bytes per frame = 25 8-byte words; the group's rows are contiguous and 8-byte aligned
        const uint2* sv = reinterpret_cast<const uint2*>(tile);
        uint2* dst = reinterpret_cast<uint2*>(out + f0 * (long long)kR4M);
#pragma unroll
        for (int r = 0; r < (6 * 25 + 31) / 32; ++r) {
          const int idx = lane + 32 * r;
          const int fr = (idx * 2622) >> 16;                            // idx / 25
          if (idx < 6 * 25 && fr < live_frames) dst[idx] = sv[fr * (kR4OutStrideB / 8) + (idx - fr * 25)];
        }
      }
    }
    __syncwarp();   // the stage is read before the next group's edge frames / tile rows overwrite it
    f += fstep; clip += step_clip; t += step_t;
    if (t >= g.frames_per_clip) { t -= g.frames_per_clip; ++clip; }
  }
}

}  // namespace sg
