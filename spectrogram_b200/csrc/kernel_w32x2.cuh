// Headline kernel, packed form: n_fft = 2048 (the reference's fftSize, UI/player.js:10), one warp per
// PAIR of consecutive frames, every butterfly a Blackwell packed-FP32 op (FFMA2 / FADD2 / FMUL2).
//
// Same network as kernel_w32.cuh (1024-point complex radix-2 DIT split 5 + 5 around a 32x32
// shared-memory transpose, then the real-FFT untangle and the dB/byte epilogue), but each lane holds
// the same element of TWO frames in the two halves of a 64-bit register pair (re_A, re_B), (im_A, im_B)
// so one issue slot does the butterfly arithmetic of both frames.  Twiddles and window values are
// shared by the two frames and enter as FFMA2's scalar-broadcast operand, which also halves their
// shared-memory traffic per frame.  The FP32 pipe does the same number of lane-FMAs either way
// (FFMA2 occupies the pipe for two passes -- tools/microbench/ffma2_bench.cu); what packing buys is
// issue slots: the un-packed kernel is issue-bound (69 % issue utilisation, 49 % FMA pipe).
//
// Loader (north_star subsystem 1): the 2048 + hop samples the two overlapping frames cover are
// brought into the warp's shared-memory stage by ONE TMA bulk copy (cp.async.bulk, completion on
// an mbarrier), issued a whole iteration ahead so HBM latency hides behind the previous pair's FFT;
// lanes then read hop-strided float2 pairs from the stage.  Frames the bulk copy cannot express
// (clip edges / zero history, hop not a multiple of 4, pairs that straddle clips) take guarded loads.
#pragma once
#include "common.cuh"
#include "ct_math.cuh"
#include "kernel_w32.cuh"

namespace sg {

constexpr int kX2Stride = 33;                         // float2 per plane row (odd: 64-bit accesses conflict free)
constexpr int kX2PlaneF2 = 32 * kX2Stride;            // 1056 float2 = 8448 B: exchange plane (re, then im)
constexpr int kX2StageFloats = kW32N + 1024 + 8;      // both frames' samples when hop <= 1024 (+ the slack of an
                                                      // unaligned span copied from its 16-byte-aligned-down address)
constexpr int kX2BytesStage = 2 * kW32M;              // u8 staging: 1024 (A,B) byte pairs
constexpr int kX2WarpBytes = kX2PlaneF2 * 8 + kX2StageFloats * 4 + kX2BytesStage + 16;   // 22800 B
constexpr int kX2Warps = 8;
constexpr int kX2SmemBytes = kW32TableBytes + kX2Warps * kX2WarpBytes;                   // ~203 KB

// ---------------------------------------------------------------- packed-pair arithmetic
struct P2 {  // (frame A, frame B)
  float2 v;
  __device__ __forceinline__ P2() {}
  __device__ __forceinline__ P2(float2 a) : v(a) {}
  __device__ __forceinline__ P2(float a, float b) : v(make_float2(a, b)) {}
};
__device__ __forceinline__ P2 bc(float s) { return P2(s, s); }   // scalar broadcast operand
__device__ __forceinline__ P2 neg(P2 a) { return P2(-a.v.x, -a.v.y); }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { return P2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { return P2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { return P2(__fmul2_rn(a.v, b.v)); }

struct C2 {  // one complex element of both frames
  P2 re, im;
};

// x' = x + w*y ; y' = 2x - x'
__device__ __forceinline__ void bfly2(C2& x, C2& y, float wr, float wi) {
  const P2 xr = fma2(y.re, bc(wr), fma2(y.im, bc(-wi), x.re));
  const P2 xi = fma2(y.im, bc(wr), fma2(y.re, bc(wi), x.im));
  y.re = fma2(x.re, bc(2.f), neg(xr));
  y.im = fma2(x.im, bc(2.f), neg(xi));
  x.re = xr;
  x.im = xi;
}

template <int P, int N>
__device__ __forceinline__ void bfly2_const(C2& x, C2& y) {
  if constexpr (P == 0) {                    // w = 1
    const C2 t = y;
    y.re = add2(x.re, neg(t.re)); y.im = add2(x.im, neg(t.im));
    x.re = add2(x.re, t.re);      x.im = add2(x.im, t.im);
  } else if constexpr (4 * P == N) {         // w = -i : w*y = (y.im, -y.re)
    const C2 t = y;
    y.re = add2(x.re, neg(t.im)); y.im = add2(x.im, t.re);
    x.re = add2(x.re, t.im);      x.im = add2(x.im, neg(t.re));
  } else {
    bfly2(x, y, Twiddle<P, N>::re, Twiddle<P, N>::im);
  }
}

template <int S>
__device__ __forceinline__ void dit2_stage_const(C2 (&a)[32]) {
  constexpr int half = 1 << (S - 1);
  static_for<0, 16>([&](auto idx) {
    constexpr int i = decltype(idx)::value;
    constexpr int blk = i / half, p = i % half, i0 = blk * 2 * half + p;
    bfly2_const<p, 2 * half>(a[i0], a[i0 + half]);
  });
}

template <int S>
__device__ __forceinline__ void dit2_stage_table(C2 (&a)[32], const float2* __restrict__ tw_lane) {
  constexpr int half = 1 << (S - 1);
  static_for<0, half>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    const float2 w = tw_lane[(half - 1 + p) * 32];
    static_for<0, 16 / half>([&](auto bb) {
      constexpr int i0 = decltype(bb)::value * 2 * half + p;
      bfly2(a[i0], a[i0 + half], w.x, w.y);
    });
  });
}

// ---------------------------------------------------------------- epilogue helpers
// float dB of one power pair: db_scale * lg2(p) + db_off, packed (lg2.approx.ftz: powers below 2^-126 read as 0)
__device__ __forceinline__ P2 db_of_power(P2 p, const Epilogue& ep) {
  return fma2(P2(lg2_ftz(p.v.x), lg2_ftz(p.v.y)), bc(ep.db_scale), bc(ep.db_off));
}
__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }
// float outputs of one power pair: dB as above, or the linear magnitude sqrt(p) * mag_scale; and what a frame with a
// non-finite sample reads in that output ([SPEC] magnitude 0: -inf dB)
template <int OUT>
__device__ __forceinline__ P2 float_of_power(P2 p, const Epilogue& ep) {
  if constexpr (OUT == kOutF32Db) return db_of_power(p, ep);
  else return mul2(P2(sqrt_ftz(p.v.x), sqrt_ftz(p.v.y)), bc(ep.mag_scale));
}
template <int OUT>
__device__ __forceinline__ float float_of_poisoned() { return OUT == kOutF32Db ? neg_inf() : 0.f; }

// byte path for one power pair (tau == 0):
//   q = p*0 + p        finite p -> p ; Inf/NaN -> NaN      ([SPEC] non-finite -> 0, via cvt(NaN) = 0)
//   v = a*lg2(q) + b   ; byte = cvt.rzi.u8.f32(v)
__device__ __forceinline__ void bytes_of_power(P2 p, const Epilogue& ep, unsigned& ba, unsigned& bb) {
  const P2 q = fma2(p, bc(0.f), p);
  const P2 v = fma2(P2(lg2_ftz(q.v.x), lg2_ftz(q.v.y)), bc(ep.byte_a), bc(ep.byte_b));
  ba = byte_of_scaled(v.v.x);
  bb = byte_of_scaled(v.v.y);
}

// ---------------------------------------------------------------- TMA bulk copy + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- frame-pair geometry
struct PairGeom {
  long long fa;            // global index of frame A (frame B = fa + 1 when has_b)
  long long clip_a, ta;    // clip and in-clip index of A
  long long clip_b, tb;
  bool has_b;
  bool tma;                // both frames inside one clip, hop <= 1024 and (aligned mode) a 16-byte aligned span, hop % 4 == 0
  int mis;                 // unaligned mode: floats between the aligned-down copy source and the span's first sample
};

// ANY_ALIGN: spans that start at any 4-byte offset (odd hops such as 441 samples = 10 ms at 44.1 kHz): the bulk copy
// starts at the aligned-down address and the lanes read the stage at an offset
template <bool ANY_ALIGN = false>
__device__ __forceinline__ PairGeom make_pair(const FrameGeom& g, long long fa, long long clip_a, long long ta) {
  PairGeom p;
  p.fa = fa; p.clip_a = clip_a; p.ta = ta;
  p.has_b = fa + 1 < g.total_frames;
  p.clip_b = clip_a; p.tb = ta;
  if (p.has_b) {
    if (ta + 1 == g.frames_per_clip) { p.clip_b = clip_a + 1; p.tb = 0; }
    else p.tb = ta + 1;
  }
  const long long start_a = g.start0 + ta * g.hop;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(g.pcm + clip_a * g.clip_stride + start_a);
  p.mis = (int)((addr & 15) >> 2);
  if constexpr (ANY_ALIGN) {
    // the copy is rounded out to 16-byte boundaries: up to 3 floats before the span (inside this clip or the
    // previous one's tail) and up to 3 after it, which must still lie inside this clip
    p.tma = p.has_b && p.clip_b == clip_a && g.hop <= 1024 && start_a >= 0 && start_a + g.hop + kW32N + 3 <= g.clip_len &&
            (clip_a > 0 || start_a >= p.mis);
  } else {
    p.tma = p.has_b && p.clip_b == clip_a && g.hop <= 1024 && (g.hop & 3) == 0 && start_a >= 0 &&
            start_a + g.hop + kW32N <= g.clip_len && (addr & 15) == 0;
  }
  return p;
}

// Stage 1 of the DIT network pairs z[b + 32 j] with z[b + 32 (j + 16)] (w = 1); the window multiply is
// folded into it: x' = s0*w0 + s1*w1, y' = s0*w0 - s1*w1  (1 FMUL + 2 FFMA instead of 2 FMUL + 2 FADD).
__device__ __forceinline__ void window_stage1(C2& lo, C2& hi, float2 a0, float2 a1, float2 b0, float2 b1, float2 w0,
                                              float2 w1) {
  const float ta = a0.x * w0.x, tb = b0.x * w0.x, ua = a0.y * w0.y, ub = b0.y * w0.y;
  lo.re = P2(fmaf(a1.x, w1.x, ta), fmaf(b1.x, w1.x, tb));
  hi.re = P2(fmaf(a1.x, -w1.x, ta), fmaf(b1.x, -w1.x, tb));
  lo.im = P2(fmaf(a1.y, w1.y, ua), fmaf(b1.y, w1.y, ub));
  hi.im = P2(fmaf(a1.y, -w1.y, ua), fmaf(b1.y, -w1.y, ub));
}

// HOPQ > 0: hop == 64*HOPQ, so frame B's element j is the stage element j + HOPQ of the same lane and the
// two frames share their loads (32 + HOPQ instead of 64 per lane).  HOPQ == 0: any hop that is a multiple of 4
// samples.  HOPQ < 0: any hop at all (north_star subsystem 1 for hops such as 441 samples): the bulk copy starts at
// the span's 16-byte-aligned-down address and the lanes read the stage at the resulting offset.
template <int OUT, int HOPQ>
__global__ void __launch_bounds__(kX2Warps * 32, 1)
stft_w32x2_kernel(FrameGeom g, W32Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);                 // [1024] (w[2n], w[2n+1])
  float2* s_tw2 = s_win + kW32M;                                       // [31*32]
  float2* s_ut = s_tw2 + 31 * 32;                                      // [16*32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_ut + 16 * 32) + warp * kX2WarpBytes;
  float2* xp = reinterpret_cast<float2*>(wbase);                                    // exchange plane
  float* stage = reinterpret_cast<float*>(wbase + kX2PlaneF2 * 8);                  // TMA destination
  uint16_t* sb16 = reinterpret_cast<uint16_t*>(wbase + kX2PlaneF2 * 8 + kX2StageFloats * 4);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + kX2PlaneF2 * 8 + kX2StageFloats * 4 + kX2BytesStage);

  for (int i = threadIdx.x; i < kW32M; i += blockDim.x)
    s_win[i] = __ldg(reinterpret_cast<const float2*>(pl.win) + i);
  for (int i = threadIdx.x; i < 31 * 32; i += blockDim.x) s_tw2[i] = __ldg(pl.tw2 + i);
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) s_ut[i] = __ldg(pl.ut + i);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncthreads();

  // frame pair (fa, fa+1); (clip, t) advanced incrementally (no 64-bit division in the loop)
  constexpr bool ANY = HOPQ < 0;
  const long long step = 2LL * gridDim.x * kX2Warps;
  const long long step_clip = step / g.frames_per_clip, step_t = step - step_clip * g.frames_per_clip;
  // bytes of the bulk copy: the span itself, or (unaligned mode) the span rounded out to 16-byte boundaries
  auto span_bytes_of = [&](const PairGeom& q) { return ANY ? (unsigned)(((kW32N + g.hop + q.mis) * 4 + 15) & ~15) : (unsigned)(kW32N + g.hop) * 4u; };
  auto span_src_of = [&](const PairGeom& q) { return g.pcm + q.clip_a * g.clip_stride + g.start0 + q.ta * g.hop - (ANY ? q.mis : 0); };
  long long fa0 = 2 * ((long long)blockIdx.x * kX2Warps + warp);
  if (fa0 >= g.total_frames) return;
  PairGeom cur = make_pair<ANY>(g, fa0, fa0 / g.frames_per_clip, fa0 % g.frames_per_clip);
  if (cur.tma && lane == 0) {
    mbar_expect_tx(bar, span_bytes_of(cur));
    tma_bulk_g2s(stage, span_src_of(cur), span_bytes_of(cur), bar);
  }
  unsigned phase = 0;

  while (true) {
    // geometry of the next pair (prefetched below)
    long long nfa = cur.fa + step, nclip = cur.clip_a + step_clip, nt = cur.ta + step_t;
    if (nt >= g.frames_per_clip) { nt -= g.frames_per_clip; ++nclip; }
    const bool has_next = nfa < g.total_frames;
    PairGeom nxt = cur;
    if (has_next) nxt = make_pair<ANY>(g, nfa, nclip, nt);

    // ---- steps 1-2 (+ FFT stage 1): time blocks of both frames, window, bit-reversed into registers
    C2 a[32];
    if (cur.tma) {
      while (!mbar_try_wait(bar, phase)) {}
      phase ^= 1;
      const float2* sa = reinterpret_cast<const float2*>(stage) + lane;
      if constexpr (HOPQ > 0) {
        float2 s[32 + HOPQ];
        static_for<0, 32 + HOPQ>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = sa[32 * m]; });
        static_for<0, 16>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);   // r1 == r0 + 1
          window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + HOPQ], s[j + 16 + HOPQ], s_win[lane + 32 * j],
                        s_win[lane + 32 * (j + 16)]);
        });
      } else if constexpr (HOPQ == 0) {
        const float2* sb = reinterpret_cast<const float2*>(stage + g.hop) + lane;
        static_for<0, 16>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
          window_stage1(a[r0], a[r1], sa[32 * j], sa[32 * (j + 16)], sb[32 * j], sb[32 * (j + 16)],
                        s_win[lane + 32 * j], s_win[lane + 32 * (j + 16)]);
        });
      } else {
        // unaligned span: the samples sit `mis` floats into the stage and frame B another `hop` further: 4-byte reads
        const float* fa_ = stage + cur.mis + 2 * lane;
        const float* fb_ = fa_ + g.hop;
        static_for<0, 16>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
          window_stage1(a[r0], a[r1], make_float2(fa_[64 * j], fa_[64 * j + 1]), make_float2(fa_[64 * (j + 16)], fa_[64 * (j + 16) + 1]),
                        make_float2(fb_[64 * j], fb_[64 * j + 1]), make_float2(fb_[64 * (j + 16)], fb_[64 * (j + 16) + 1]),
                        s_win[lane + 32 * j], s_win[lane + 32 * (j + 16)]);
        });
      }
    } else {
      const float* __restrict__ xa = g.pcm + cur.clip_a * g.clip_stride;
      const float* __restrict__ xb = g.pcm + cur.clip_b * g.clip_stride;
      const long long start_a = g.start0 + cur.ta * g.hop, start_b = g.start0 + cur.tb * g.hop;
      auto ld = [&](const float* __restrict__ x, long long s) { return (s >= 0 && s < g.clip_len) ? __ldg(x + s) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = 2 * (lane + 32 * j), o1 = 2 * (lane + 32 * (j + 16));
        window_stage1(a[r0], a[r1], make_float2(ld(xa, start_a + o0), ld(xa, start_a + o0 + 1)),
                      make_float2(ld(xa, start_a + o1), ld(xa, start_a + o1 + 1)),
                      make_float2(ld(xb, start_b + o0), ld(xb, start_b + o0 + 1)),
                      make_float2(ld(xb, start_b + o1), ld(xb, start_b + o1 + 1)), s_win[lane + 32 * j],
                      s_win[lane + 32 * (j + 16)]);
      });
    }
    __syncwarp();   // every lane has consumed the stage
    if (has_next && nxt.tma && lane == 0) {
      fence_proxy_async();   // order the generic-proxy reads above before the async-proxy overwrite
      mbar_expect_tx(bar, span_bytes_of(nxt));
      tma_bulk_g2s(stage, span_src_of(nxt), span_bytes_of(nxt), bar);
    }

    // ---- pass 1: stages 2-5 in registers (stage 1 was fused with the window), compile-time twiddles
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);

    // ---- exchange: plane[b = lane][k_a] -> lane a reads plane[bitrev(q)][a]; re pairs, then im pairs
    static_for<0, 32>([&](auto qq) { constexpr int q = decltype(qq)::value; xp[lane * kX2Stride + q] = a[q].re.v; });
    __syncwarp();
    static_for<0, 32>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      a[q].re = P2(xp[bitrev(q, 5) * kX2Stride + lane]);
    });
    __syncwarp();
    static_for<0, 32>([&](auto qq) { constexpr int q = decltype(qq)::value; xp[lane * kX2Stride + q] = a[q].im.v; });
    __syncwarp();
    static_for<0, 32>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      a[q].im = P2(xp[bitrev(q, 5) * kX2Stride + lane]);
    });
    __syncwarp();

    // ---- pass 2: stages 6-10, lane-dependent twiddles (shared by both frames)
    const float2* tw_lane = s_tw2 + lane;
    dit2_stage_table<1>(a, tw_lane);
    dit2_stage_table<2>(a, tw_lane);
    dit2_stage_table<3>(a, tw_lane);
    dit2_stage_table<4>(a, tw_lane);
    dit2_stage_table<5>(a, tw_lane);
    // now a[i] = Z[lane + 32 i] of both frames
    // a non-finite sample makes every Z of its frame non-finite: one flag per frame for the float dB path
    const P2 poison = fma2(a[0].re, bc(0.f), mul2(a[0].im, bc(0.f)));   // 0 or NaN per frame
    const bool bad_a = !(poison.v.x == 0.f), bad_b = !(poison.v.y == 0.f);
    const P2 byte_scale = add2(bc(ep.byte_a), poison);

    // ---- untangle exchange: upper half (k >= 512) to smem at index k - 512; Z[1024] == Z[0] at 512.
    //      re pairs at xp[0..513), im pairs at xp[520..1033)
    float2* xre = xp;
    float2* xim = xp + 520;
    static_for<16, 32>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      xre[lane + 32 * (i - 16)] = a[i].re.v;
      xim[lane + 32 * (i - 16)] = a[i].im.v;
    });
    if (lane == 0) { xre[512] = a[0].re.v; xim[512] = a[0].im.v; }
    __syncwarp();

    P2 pk[16], pm[16];
    static_for<0, 16>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int k = lane + 32 * i;
      const P2 zmr(xre[512 - k]), zmi(xim[512 - k]);
      const float2 w = s_ut[i * 32 + lane];
      const C2 zk = a[i];
      const P2 ex = add2(zk.re, zmr), ey = add2(zk.im, neg(zmi));          // 2E
      const P2 ox = add2(zk.im, zmi), oy = add2(zmr, neg(zk.re));          // 2O
      const P2 xr = fma2(ox, bc(w.x), fma2(oy, bc(-w.y), ex));             // 2X[k]
      const P2 xi = fma2(ox, bc(w.y), fma2(oy, bc(w.x), ey));
      const P2 yr = fma2(ex, bc(2.f), neg(xr));                            // 2 conj X[1024-k]
      const P2 yi = fma2(ey, bc(2.f), neg(xi));
      pk[i] = fma2(xr, xr, mul2(xi, xi));
      pm[i] = fma2(yr, yr, mul2(yi, yi));
    });
    // lane 0: the mirror of k = 0 is the Nyquist bin (dropped); its slot carries bin 512 = conj Z[512],
    // which lane 0 holds in a[16]
    const bool special = lane == 0;
    if (special) pm[0] = mul2(bc(4.f), fma2(a[16].re, a[16].re, mul2(a[16].im, a[16].im)));

    T* __restrict__ row_a = out + cur.fa * (long long)kW32M;
    T* __restrict__ row_b = row_a + kW32M;   // frame B is the next global frame
    static_for<0, 16>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int k = lane + 32 * i;
      int mk = kW32M - k;
      if constexpr (i == 0) { if (special) mk = 512; }
      if constexpr (OUT == kOutU8) {
        // the per-frame non-finite flag rides in the byte scale: NaN scale -> NaN -> byte 0
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), byte_scale, bc(ep.byte_b));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), byte_scale, bc(ep.byte_b));
        const unsigned ka = byte_of_scaled(vk.v.x), kb = byte_of_scaled(vk.v.y);
        const unsigned ma = byte_of_scaled(vm.v.x), mb = byte_of_scaled(vm.v.y);
        sb16[k] = (uint16_t)(ka | (kb << 8));
        sb16[mk] = (uint16_t)(ma | (mb << 8));
      } else if constexpr (OUT == kOutRgba8) {
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), byte_scale, bc(ep.byte_b));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), byte_scale, bc(ep.byte_b));
        const unsigned ka = byte_of_scaled(vk.v.x), kb = byte_of_scaled(vk.v.y);
        const unsigned ma = byte_of_scaled(vm.v.x), mb = byte_of_scaled(vm.v.y);
        row_a[k] = __ldg(ep.lut + ka); row_a[mk] = __ldg(ep.lut + ma);
        if (cur.has_b) { row_b[k] = __ldg(ep.lut + kb); row_b[mk] = __ldg(ep.lut + mb); }
      } else {
        // float dB / magnitude, packed; the non-finite rule from the per-frame flag (magnitude 0: -inf dB)
        const P2 vk = float_of_power<OUT>(pk[i], ep), vm = float_of_power<OUT>(pm[i], ep);
        const float z = float_of_poisoned<OUT>();
        row_a[k] = bad_a ? z : vk.v.x; row_a[mk] = bad_a ? z : vm.v.x;
        if (cur.has_b) { row_b[k] = bad_b ? z : vk.v.y; row_b[mk] = bad_b ? z : vm.v.y; }
      }
    });
    if constexpr (OUT == kOutU8) {
      __syncwarp();
      // de-interleave the (A,B) byte pairs: 8 bins per lane per round, 8-byte coalesced row stores
      const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
      uint2* ra = reinterpret_cast<uint2*>(row_a);
      uint2* rb = reinterpret_cast<uint2*>(row_b);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 w = s16[c * 32 + lane];
        ra[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
        if (cur.has_b) rb[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
      }
    }
    __syncwarp();
    if (!has_next) break;
    cur = nxt;
  }
}

}  // namespace sg
