// n_fft = 512 (BASELINE config 2: 1 h of 16 kHz speech, hop 160, float dB).  M = 256 complex points = 16 x 16:
//   16 threads per frame (2 frames per warp, 16 per CTA), 16 complex points per thread, ~64 registers, so
//   four CTAs (32 warps) share an SM and the loads need no software pipelining
//   loader   thread t holds z[t + 16 j]: the 16 lanes of a frame read 128 contiguous bytes per instruction
//            (one L1 line per frame; the 32-points-per-thread family reads 64-byte pieces of four frames)
//   pass 1   16-point radix-2 DIT in registers, compile-time twiddles
//   xchg     16 x 16 tile per frame, row stride 17 float2: conflict free both ways, __syncwarp only
//   pass 2   thread t takes column k1 = t: 16-point DIT with the 15 twiddles W_{32 h}^{16 p + t} from a
//            lane-major table -> Z[t + 16 q]
//   untangle Z[256 - k] lives in thread (16 - t) of the same frame: 16 warp shuffles fetch the mirrors in place
//   epilogue |X|^2 -> dB / byte / colour; float rows leave as 64-byte pieces, byte rows are staged in the tile
//            and leave as one 16-byte store per thread
// About 90 L1/shared wavefronts per frame against ~180 for the 32-points-per-thread family.
// Algorithmic bytes per frame: 4*hop + elem*256 (1664 B for float dB at hop 160).
#pragma once
#include "common.cuh"
#include "ct_math.cuh"
#include "kernel_w32.cuh"   // bfly, bfly_const
#include "plans.cuh"

namespace sg {

constexpr int kW16N = 512, kW16M = 256;
constexpr int kW16Threads = 256, kW16FPC = 16;             // frames per CTA
constexpr int kW16Stride = 17;                             // float2 per tile row
constexpr int kW16TileF2 = 16 * kW16Stride;                // 272 float2 = 2176 B per frame
constexpr int kW16TableF2 = kW16M + 15 * 16 + (kW16M / 2 + 2);   // window pairs + pass-2 twiddles + W_512^k (k <= 128)
constexpr int kW16SmemBytes = (kW16TableF2 + kW16FPC * kW16TileF2) * 8;

template <int S, int NPTS>
__device__ __forceinline__ void dit_stage_const_n(float2 (&a)[NPTS]) {
  constexpr int half = 1 << (S - 1);
  static_for<0, NPTS / 2>([&](auto idx) {
    constexpr int i = decltype(idx)::value;
    constexpr int blk = i / half, p = i % half, i0 = blk * 2 * half + p;
    bfly_const<p, 2 * half>(a[i0], a[i0 + half]);
  });
}

template <int OUT>
__global__ void __launch_bounds__(kW16Threads, 4)
stft_w16_kernel(FrameGeom g, W16Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using TO = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);      // [256] (w[2m], w[2m+1])
  float2* s_tw = s_win + kW16M;                             // [15][16]  row (h - 1 + p): W_{32 h}^{16 p + col}
  float2* s_ut = s_tw + 15 * 16;                            // [130]     W_512^k
  const int tid = threadIdx.x, lane = tid & 31, fs = tid >> 4, t = tid & 15;
  float2* A = s_win + kW16TableF2 + fs * kW16TileF2;

  for (int i = tid; i < kW16M; i += kW16Threads) s_win[i] = __ldg(reinterpret_cast<const float2*>(pl.win) + i);
  for (int i = tid; i < 15 * 16; i += kW16Threads) s_tw[i] = __ldg(pl.tw + i);
  for (int i = tid; i <= kW16M / 2; i += kW16Threads) s_ut[i] = __ldg(pl.ut + i);
  __syncthreads();

  const int partner = (lane & 16) | ((16 - t) & 15);
  const bool t0 = t == 0;

  // this slot's frame, advanced incrementally (no 64-bit division in the loop)
  const long long groups = (g.total_frames + kW16FPC - 1) / kW16FPC;
  const long long fstep = (long long)gridDim.x * kW16FPC;
  const long long step_clip = fstep / g.frames_per_clip, step_t = fstep - step_clip * g.frames_per_clip;
  long long f = (long long)blockIdx.x * kW16FPC + fs;
  long long fclip = f / g.frames_per_clip, ft = f - fclip * g.frames_per_clip;
  for (long long gi = blockIdx.x; gi < groups; gi += gridDim.x, f += fstep, fclip += step_clip, ft += step_t) {
    if (ft >= g.frames_per_clip) { ft -= g.frames_per_clip; ++fclip; }
    const bool live = f < g.total_frames;
    long long fc = f, clip = fclip, tt = ft;
    if (!live) {                                              // idle slots recompute the last frame, store nothing
      fc = g.total_frames - 1;
      clip = fc / g.frames_per_clip;
      tt = fc - clip * g.frames_per_clip;
    }
    const long long start = g.start0 + tt * g.hop;
    const float* __restrict__ x = g.pcm + clip * g.clip_stride;

    // ---- steps 1-2: time block, window, bit-reversed into registers
    float2 v[16];
    const bool interior = start >= 0 && start + kW16N <= g.clip_len && ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
    if (interior) {
      const float2* __restrict__ src = reinterpret_cast<const float2*>(x + start) + t;
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 s = __ldg(src + 16 * j), w = s_win[t + 16 * j];
        v[bitrev(j, 4)] = make_float2(s.x * w.x, s.y * w.y);
      });
    } else if (start >= 0 && start + kW16N <= g.clip_len) {
      // inside the clip, only misaligned (odd hops): 4-byte loads, no bounds checks
      const float* __restrict__ src = x + start + 2 * t;
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float a0 = __ldg(src + 32 * j), a1 = __ldg(src + 32 * j + 1);
        const float2 w = s_win[t + 16 * j];
        v[bitrev(j, 4)] = make_float2(a0 * w.x, a1 * w.y);
      });
    } else {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const long long s0 = start + 2 * (t + 16 * j), s1 = s0 + 1;
        const float a0 = (s0 >= 0 && s0 < g.clip_len) ? __ldg(x + s0) : 0.f;
        const float a1 = (s1 >= 0 && s1 < g.clip_len) ? __ldg(x + s1) : 0.f;
        const float2 w = s_win[t + 16 * j];
        v[bitrev(j, 4)] = make_float2(a0 * w.x, a1 * w.y);
      });
    }

    // ---- pass 1: stages 1-4 in registers
    dit_stage_const_n<1, 16>(v);
    dit_stage_const_n<2, 16>(v);
    dit_stage_const_n<3, 16>(v);
    dit_stage_const_n<4, 16>(v);

    // ---- exchange: tile[row t][k1] -> thread t reads column t, rows in bit-reversed order
    static_for<0, 16>([&](auto kk) { constexpr int k = decltype(kk)::value; A[t * kW16Stride + k] = v[k]; });
    __syncwarp();
    static_for<0, 16>([&](auto qq) { constexpr int q = decltype(qq)::value; v[q] = A[bitrev(q, 4) * kW16Stride + t]; });
    __syncwarp();

    // ---- pass 2: stages 5-8, twiddle row (h - 1 + p) of the lane-major table
    static_for<1, 5>([&](auto uu) {
      constexpr int u = decltype(uu)::value, half = 1 << (u - 1);
      static_for<0, half>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        const float2 w = s_tw[(half - 1 + p) * 16 + t];
        static_for<0, 16 / (2 * half)>([&](auto bb) {
          constexpr int i0 = decltype(bb)::value * 2 * half + p;
          bfly(v[i0], v[i0 + half], w.x, w.y);
        });
      });
    });
    // now v[q] = Z[t + 16 q]

    // [SPEC] "non-finite -> 0", decided once per frame (a non-finite sample makes every Z of its frame non-finite)
    const bool bad = !(fabsf(v[0].x) <= 3.4028235e38f) || !(fabsf(v[0].y) <= 3.4028235e38f);
    // bin 128 = conj Z[128] is thread 0's v[8]; taken before v[8] is replaced by a mirror
    const float p128 = 4.f * fmaf(v[8].x, v[8].x, v[8].y * v[8].y);

    // ---- untangle: Z[256 - k], k = t + 16 i, is thread (16 - t)'s v[15 - i] (thread 0: its own v[16 - i]).
    //      Mirrors are fetched in place, descending i, so thread 0's own sources are still intact when read.
    static_for<0, 8>([&](auto ii) {
      constexpr int i = 7 - decltype(ii)::value;
      constexpr int src = 15 - i, own = (16 - i) & 15;
      const float mx = __shfl_sync(0xffffffffu, v[src].x, partner);
      const float my = __shfl_sync(0xffffffffu, v[src].y, partner);
      v[src] = make_float2(t0 ? v[own].x : mx, t0 ? v[own].y : my);
    });

    TO* __restrict__ row = out + fc * (long long)kW16M;
    unsigned char* sb = reinterpret_cast<unsigned char*>(A);   // byte staging (the tile is idle now)
    static_for<0, 8>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int k = t + 16 * i;
      int mk = kW16M - k;
      const float2 zk = v[i], zm = v[15 - i];
      const float2 w = s_ut[k];
      const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
      const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
      const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
      const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
      const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[256-k]
      const float yi = fmaf(2.f, ey, -xi);
      float pk = fmaf(xr, xr, xi * xi);
      float pm = fmaf(yr, yr, yi * yi);
      if constexpr (i == 0) {
        // thread 0: the mirror of k = 0 is the dropped Nyquist bin; its slot carries bin 128
        pm = t0 ? p128 : pm;
        mk = t0 ? kW16M / 2 : mk;
      }
      pk = bad ? 0.f : pk;
      pm = bad ? 0.f : pm;
      if constexpr (OUT == kOutU8) {
        sb[k] = emit_power_finite<OUT>(pk, ep);
        sb[mk] = emit_power_finite<OUT>(pm, ep);
      } else {
        // 4-byte rows are staged in the frame's idle tile too: a thread's bins are 16 apart, so direct stores would
        // fill an eighth of each 32-byte sector per instruction
        reinterpret_cast<TO*>(A)[k] = emit_power_finite<OUT>(pk, ep);
        reinterpret_cast<TO*>(A)[mk] = emit_power_finite<OUT>(pm, ep);
      }
    });
    __syncwarp();
    if constexpr (OUT == kOutU8) {
      if (live) reinterpret_cast<uint4*>(row)[t] = reinterpret_cast<const uint4*>(sb)[t];
    } else if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c)   // 256 elements = 64 16-byte words, 16 threads
        reinterpret_cast<uint4*>(row)[c * 16 + t] = reinterpret_cast<const uint4*>(A)[c * 16 + t];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------
// n_fft = 256: M = 128 = 16 x 8.  Eight threads per frame (4 frames per warp, 32 per CTA), 16 points per thread.
// Pass 1 as above; pass 2 is 16 columns of 8-point FFTs and thread t takes the mirror pair (t, 16 - t) -- thread 0
// the self-mirrored columns 0 and 8 -- so Z[k] and Z[128 - k] meet in the same thread: no shuffle, no second
// exchange.  The lower halves of both columns lead (k < 64, inside the W_256^k table).
constexpr int kW8N = 256, kW8M = 128;
constexpr int kW8FPC = 32;                                  // frames per CTA (256 threads)
constexpr int kW8TileF2 = 8 * kW16Stride;                   // 136 float2 = 1088 B per frame
constexpr int kW8TableF2 = kW8M + 7 * 16 + (kW8M / 2 + 2);  // window pairs + pass-2 twiddles + W_256^k (k <= 64)
constexpr int kW8SmemBytes = (kW8TableF2 + kW8FPC * kW8TileF2) * 8;

template <int OUT>
__global__ void __launch_bounds__(kW16Threads, 4)
stft_w16x8_kernel(FrameGeom g, W16Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using TO = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);      // [128]
  float2* s_tw = s_win + kW8M;                              // [7][16]  row (h - 1 + p): W_{32 h}^{16 p + col}
  float2* s_ut = s_tw + 7 * 16;                             // [66]     W_256^k
  const int tid = threadIdx.x, fs = tid >> 3, t = tid & 7;
  float2* A = s_win + kW8TableF2 + fs * kW8TileF2;

  for (int i = tid; i < kW8M; i += kW16Threads) s_win[i] = __ldg(reinterpret_cast<const float2*>(pl.win) + i);
  for (int i = tid; i < 7 * 16; i += kW16Threads) s_tw[i] = __ldg(pl.tw + i);
  for (int i = tid; i <= kW8M / 2; i += kW16Threads) s_ut[i] = __ldg(pl.ut + i);
  __syncthreads();

  const bool t0 = t == 0;
  const int ka = t, kb = t0 ? 8 : 16 - t;

  const long long groups = (g.total_frames + kW8FPC - 1) / kW8FPC;
  const long long fstep = (long long)gridDim.x * kW8FPC;
  const long long step_clip = fstep / g.frames_per_clip, step_t = fstep - step_clip * g.frames_per_clip;
  long long f = (long long)blockIdx.x * kW8FPC + fs;
  long long fclip = f / g.frames_per_clip, ft = f - fclip * g.frames_per_clip;
  for (long long gi = blockIdx.x; gi < groups; gi += gridDim.x, f += fstep, fclip += step_clip, ft += step_t) {
    if (ft >= g.frames_per_clip) { ft -= g.frames_per_clip; ++fclip; }
    const bool live = f < g.total_frames;
    long long fc = f, clip = fclip, tt = ft;
    if (!live) {
      fc = g.total_frames - 1;
      clip = fc / g.frames_per_clip;
      tt = fc - clip * g.frames_per_clip;
    }
    const long long start = g.start0 + tt * g.hop;
    const float* __restrict__ x = g.pcm + clip * g.clip_stride;

    float2 v[16];
    const bool interior = start >= 0 && start + kW8N <= g.clip_len && ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
    if (interior) {
      const float2* __restrict__ src = reinterpret_cast<const float2*>(x + start) + t;
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 s = __ldg(src + 8 * j), w = s_win[t + 8 * j];
        v[bitrev(j, 4)] = make_float2(s.x * w.x, s.y * w.y);
      });
    } else if (start >= 0 && start + kW8N <= g.clip_len) {
      const float* __restrict__ src = x + start + 2 * t;     // inside the clip, only misaligned
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float a0 = __ldg(src + 16 * j), a1 = __ldg(src + 16 * j + 1);
        const float2 w = s_win[t + 8 * j];
        v[bitrev(j, 4)] = make_float2(a0 * w.x, a1 * w.y);
      });
    } else {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const long long s0 = start + 2 * (t + 8 * j), s1 = s0 + 1;
        const float a0 = (s0 >= 0 && s0 < g.clip_len) ? __ldg(x + s0) : 0.f;
        const float a1 = (s1 >= 0 && s1 < g.clip_len) ? __ldg(x + s1) : 0.f;
        const float2 w = s_win[t + 8 * j];
        v[bitrev(j, 4)] = make_float2(a0 * w.x, a1 * w.y);
      });
    }
    dit_stage_const_n<1, 16>(v);
    dit_stage_const_n<2, 16>(v);
    dit_stage_const_n<3, 16>(v);
    dit_stage_const_n<4, 16>(v);

    // ---- exchange: tile[row t][k1]; this thread reads columns ka and kb, rows in bit-reversed order
    static_for<0, 16>([&](auto kk) { constexpr int k = decltype(kk)::value; A[t * kW16Stride + k] = v[k]; });
    __syncwarp();
    static_for<0, 8>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      v[q] = A[bitrev(q, 3) * kW16Stride + ka];
      v[8 + q] = A[bitrev(q, 3) * kW16Stride + kb];
    });
    __syncwarp();

    // ---- pass 2: stages 5-7 on both columns
    static_for<1, 4>([&](auto uu) {
      constexpr int u = decltype(uu)::value, half = 1 << (u - 1);
      static_for<0, half>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        const float2 wa = s_tw[(half - 1 + p) * 16 + ka], wb = s_tw[(half - 1 + p) * 16 + kb];
        static_for<0, 8 / (2 * half)>([&](auto bb) {
          constexpr int i0 = decltype(bb)::value * 2 * half + p;
          bfly(v[i0], v[i0 + half], wa.x, wa.y);
          bfly(v[8 + i0], v[8 + i0 + half], wb.x, wb.y);
        });
      });
    });
    // now v[q] = Z[ka + 16 q], v[8 + q] = Z[kb + 16 q]

    const bool bad = !(fabsf(v[0].x) <= 3.4028235e38f) || !(fabsf(v[0].y) <= 3.4028235e38f);
    TO* __restrict__ row = out + fc * (long long)kW8M;
    unsigned char* sb = reinterpret_cast<unsigned char*>(A);
    auto emit2 = [&](int k, int mk, float pk, float pm) {
      pk = bad ? 0.f : pk;
      pm = bad ? 0.f : pm;
      if constexpr (OUT == kOutU8) {
        sb[k] = emit_power_finite<OUT>(pk, ep);
        sb[mk] = emit_power_finite<OUT>(pm, ep);
      } else {   // 4-byte rows are staged in the frame's idle tile and leave as 16-byte stores
        reinterpret_cast<TO*>(A)[k] = emit_power_finite<OUT>(pk, ep);
        reinterpret_cast<TO*>(A)[mk] = emit_power_finite<OUT>(pm, ep);
      }
    };
    auto untangle = [&](float2 zk, float2 zm, int k, float& pk, float& pm) {
      const float2 w = s_ut[k];
      const float ex = zk.x + zm.x, ey = zk.y - zm.y;
      const float ox = zk.y + zm.y, oy = zm.x - zk.x;
      const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));
      const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
      const float yr = fmaf(2.f, ex, -xr);
      const float yi = fmaf(2.f, ey, -xi);
      pk = fmaf(xr, xr, xi * xi);
      pm = fmaf(yr, yr, yi * yi);
    };
    static_for<0, 4>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      // general pair: Z[128 - (ka + 16 q)] = column kb, element 7 - q, and vice versa;
      // thread 0: column 0 mirrors into itself (element 8 - q), column 8 into itself (element 7 - q)
      const float2 sa = v[(8 - q) % 8], sbv = v[8 + 7 - q];
      const float2 zma = t0 ? sa : v[8 + 7 - q];
      const float2 zmb = t0 ? sbv : v[7 - q];
      float pk, pm;
      const int k1 = ka + 16 * q;
      untangle(v[q], zma, k1, pk, pm);
      int mk1 = kW8M - k1;
      if constexpr (q == 0) {
        // thread 0: the mirror of k = 0 is the dropped Nyquist bin; the slot carries bin 64 = conj Z[64] (column 0, element 4)
        pm = t0 ? 4.f * fmaf(v[4].x, v[4].x, v[4].y * v[4].y) : pm;
        mk1 = t0 ? kW8M / 2 : mk1;
      }
      emit2(k1, mk1, pk, pm);
      const int k2 = kb + 16 * q;
      untangle(v[8 + q], zmb, k2, pk, pm);
      emit2(k2, kW8M - k2, pk, pm);
    });
    __syncwarp();
    if constexpr (OUT == kOutU8) {
      if (live) reinterpret_cast<uint4*>(row)[t] = reinterpret_cast<const uint4*>(sb)[t];
    } else if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c)   // 128 elements = 32 16-byte words, 8 threads
        reinterpret_cast<uint4*>(row)[c * 8 + t] = reinterpret_cast<const uint4*>(A)[c * 8 + t];
    }
    __syncwarp();
  }
}

}  // namespace sg
