// n_fft = 1024 (BASELINE config 5: 48 kHz streams at hop 128; config 3 sweep at hop 256): the frame-pair kernel
// of kernel_w32x2p.cuh folded onto HALF a warp.
//
// M = 512 complex points = 32 x 16.  A pair of consecutive frames (A, B) is owned by 16 lanes -- two pairs per
// warp -- and every lane keeps 32 complex points of both frames in 64-bit register pairs, so every butterfly is
// one packed FFMA2 / FADD2 exactly as in the n_fft 2048 kernel:
//   loader    lane t holds z[t + 16 j]: 128 contiguous bytes per half-warp per load; the two frames share all but
//             HOPJ = hop/32 of their 32 loads; the NEXT pair's loads are issued during the untangle
//   pass 1    32-point DIT in registers (stage 1 fused with the window, stages 2-5 compile-time twiddles)
//   xchg      16 x 32 tile per pair in two row-paired planes (STS.64 / LDS.128, conflict free)
//   pass 2    32 columns of 16-point FFTs; lane c takes the mirror pair (c, 32 - c) -- lane 0 the self-mirrored
//             columns 0 and 16 -- so Z[k] and Z[512 - k] meet in the same lane: no shuffles, no second exchange;
//             twiddles W_{2^u}^p * W_{32*2^u}^col are built from four per-column bases
//   epilogue  as in the 2048 kernel (per-frame non-finite decision, MUFU.LG2, FFMA2, cvt.sat, staged byte rows)
// Algorithmic bytes per frame: 4*hop + elem*512.
#pragma once
#include "common.cuh"
#include "ct_math.cuh"
#include "kernel_w32.cuh"
#include "kernel_w32x2.cuh"
#include "kernel_w32x2p.cuh"

namespace sg {

constexpr int kP16N = 1024, kP16M = 512;
constexpr int kP16Warps = 12;
constexpr int kP16PairBytes = 2 * 8 * kXpStride * 16;      // re + im planes of one pair: 8448 B
constexpr int kP16WarpBytes = 2 * kP16PairBytes;           // two pairs per warp
constexpr int kP16TableBytes = 16 * 16 * 16 + 4 * 32 * 8 + 258 * 8;   // window quads + 4 bases x 32 columns + W_1024^k
constexpr int kP16SmemBytes = kP16TableBytes + kP16Warps * kP16WarpBytes;

// stage U (1..4) of a 16-point pass 2 on a[OFF .. OFF+16): butterflies (i0, i0 + half), twiddle W_{2 half}^p * base
template <int U, int OFF>
__device__ __forceinline__ void dit2_stage_gen16(C2 (&a)[32], float2 base) {
  constexpr int half = 1 << (U - 1);
  static_for<0, half>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    const float2 w = twiddle_times<p, 2 * half>(base);
    static_for<0, 8 / half>([&](auto bb) {
      constexpr int i0 = OFF + decltype(bb)::value * 2 * half + p;
      bfly2(a[i0], a[i0 + half], w.x, w.y);
    });
  });
}

template <int OUT, int HOPJ>   // hop = 32 * HOPJ samples: frame B's element j is element j + HOPJ of the same lane
__global__ void __launch_bounds__(kP16Warps * 32, 1)
stft_p16_kernel(FrameGeom g, P16Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  constexpr int HOP = 32 * HOPJ, NLOAD = 32 + HOPJ;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [16][16] (w2[t+16j], w2[t+16(j+16)])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 16 * 16);         // [4][32]  W_{32*2^u}^col
  float2* s_ut = s_twb + 4 * 32;                                       // [258]    W_1024^k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, t = lane & 15;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_ut + 258) + warp * kP16WarpBytes + h * kP16PairBytes;
  float4* xp = reinterpret_cast<float4*>(wbase);                       // this pair's planes (re, then im)
  uint16_t* sb16 = reinterpret_cast<uint16_t*>(wbase);                 // byte stage aliases the planes

  {
    const float2* w2 = reinterpret_cast<const float2*>(pl.win);
    for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
      const int j = i >> 4, l = i & 15;
      const float2 lo = __ldg(w2 + l + 16 * j), hi = __ldg(w2 + l + 16 * (j + 16));
      s_win4[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    for (int i = threadIdx.x; i < 4 * 32; i += blockDim.x) s_twb[i] = __ldg(pl.twb + i);
    for (int i = threadIdx.x; i <= kP16M / 2; i += blockDim.x) s_ut[i] = __ldg(pl.ut + i);
  }
  __syncthreads();

  // pair geometry of THIS half-warp, advanced incrementally
  const int fpc = (int)g.frames_per_clip;
  const int step = 4 * gridDim.x * kP16Warps;           // frames between a half-warp's consecutive pairs
  const int step_clip = step / fpc, step_t = step - step_clip * fpc;
  const long long d_off = (long long)step_clip * g.clip_stride + (long long)step_t * HOP;
  const long long wrap_off = g.clip_stride - (long long)fpc * HOP;
  const unsigned pcm_lo = (unsigned)reinterpret_cast<uintptr_t>(g.pcm);
  int t_lo, t_hi;
  {
    const long long lo = g.start0 >= 0 ? 0 : (-g.start0 + HOP - 1) / HOP;
    const long long room = g.clip_len - (HOP + kP16N) - g.start0;
    const long long hi = room < 0 ? -1 : min((long long)fpc - 2, room / HOP);
    t_lo = (int)lo;
    t_hi = (int)hi;
  }
  long long fa = 4 * ((long long)blockIdx.x * kP16Warps + warp) + 2 * h;
  if (fa - 2 * h >= g.total_frames) return;              // warp-uniform: pair 0 of this warp has no frame
  int clip = (int)(min(fa, g.total_frames - 1) / fpc);
  int tt = (int)(min(fa, g.total_frames - 1) - (long long)clip * fpc);
  long long off = clip * g.clip_stride + g.start0 + (long long)tt * HOP;
  auto is_fast = [&](long long f, int tq, long long o) {
    return f + 1 < g.total_frames && tq >= t_lo && tq <= t_hi && ((pcm_lo + ((unsigned)o << 2)) & 7) == 0;
  };
  bool cur_fast = is_fast(fa, tt, off);
  const bool c0 = t == 0;
  const int ka = t, kb = c0 ? 16 : 32 - t;

  float2 s[NLOAD];
  const float2* idle_src = reinterpret_cast<const float2*>(pl.win) + t;   // >= 2048 readable floats (build_plan)
  {
    const float2* src = cur_fast ? reinterpret_cast<const float2*>(g.pcm + off) + t : idle_src;
    static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(src + 16 * m); });
  }

  while (true) {
    const bool alive = fa < g.total_frames;              // a half-warp past the end keeps marching (barriers) but stores nothing
    // ---- steps 1-2 (+ FFT stage 1)
    C2 a[32];
    if (cur_fast) {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const float4 w = s_win4[j * 16 + t];
        window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + HOPJ], s[j + 16 + HOPJ], make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    } else {
      const long long fq = alive ? fa : g.total_frames - 1;
      const int cq = (int)(fq / fpc), tq = (int)(fq - (long long)cq * fpc);
      const bool has_b = fq + 1 < g.total_frames;
      int clip_b = cq, tb = tq;
      if (has_b) { if (tq + 1 == fpc) { ++clip_b; tb = 0; } else ++tb; }
      const float* __restrict__ xa = g.pcm + cq * g.clip_stride;
      const float* __restrict__ xb = g.pcm + clip_b * g.clip_stride;
      const long long start_a = g.start0 + (long long)tq * HOP, start_b = g.start0 + (long long)tb * HOP;
      auto ld = [&](const float* __restrict__ x, long long q) { return (q >= 0 && q < g.clip_len) ? __ldg(x + q) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = 2 * (t + 16 * j), o1 = 2 * (t + 16 * (j + 16));
        const float4 w = s_win4[j * 16 + t];
        window_stage1(a[r0], a[r1], make_float2(ld(xa, start_a + o0), ld(xa, start_a + o0 + 1)),
                      make_float2(ld(xa, start_a + o1), ld(xa, start_a + o1 + 1)),
                      make_float2(ld(xb, start_b + o0), ld(xb, start_b + o0 + 1)),
                      make_float2(ld(xb, start_b + o1), ld(xb, start_b + o1 + 1)), make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    }

    // ---- pass 1: stages 2-5
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);

    // ---- exchange: rows = lanes of the pair (16), columns = k1 (32); row pairs interleaved in 16-byte units
    {
      float2* wre = reinterpret_cast<float2*>(xp) + ((t >> 1) * kXpStride) * 2 + (t & 1);
      float2* wim = wre + 8 * kXpStride * 2;
      static_for<0, 32>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        wre[2 * q] = a[q].re.v;
        wim[2 * q] = a[q].im.v;
      });
      asm volatile("bar.sync %0, 32;" ::"r"(warp + 1) : "memory");
      const float4* rre = xp;
      const float4* rim = rre + 8 * kXpStride;
      static_for<0, 8>([&](auto qq) {   // rows 2j, 2j+1 hold q' = bitrev3(j), bitrev3(j) + 8
        constexpr int q0 = decltype(qq)::value;
        constexpr int j = bitrev(q0, 3);
        const float4 ar = rre[j * kXpStride + ka], ai = rim[j * kXpStride + ka];
        const float4 br = rre[j * kXpStride + kb], bi = rim[j * kXpStride + kb];
        a[q0].re = P2(ar.x, ar.y); a[q0 + 8].re = P2(ar.z, ar.w);
        a[q0].im = P2(ai.x, ai.y); a[q0 + 8].im = P2(ai.z, ai.w);
        a[16 + q0].re = P2(br.x, br.y); a[24 + q0].re = P2(br.z, br.w);
        a[16 + q0].im = P2(bi.x, bi.y); a[24 + q0].im = P2(bi.z, bi.w);
      });
      __syncwarp();
    }

    // ---- pass 2: 16-point DIT on column ka (a[0..16)) and column kb (a[16..32))
    dit2_stage_gen16<1, 0>(a, s_twb[0 * 32 + ka]);  dit2_stage_gen16<1, 16>(a, s_twb[0 * 32 + kb]);
    dit2_stage_gen16<2, 0>(a, s_twb[1 * 32 + ka]);  dit2_stage_gen16<2, 16>(a, s_twb[1 * 32 + kb]);
    dit2_stage_gen16<3, 0>(a, s_twb[2 * 32 + ka]);  dit2_stage_gen16<3, 16>(a, s_twb[2 * 32 + kb]);
    dit2_stage_gen16<4, 0>(a, s_twb[3 * 32 + ka]);  dit2_stage_gen16<4, 16>(a, s_twb[3 * 32 + kb]);
    // now a[q] = Z[ka + 32 q], a[16 + q] = Z[kb + 32 q] of both frames

    const P2 poison = fma2(a[0].re, bc(0.f), mul2(a[0].im, bc(0.f)));   // 0 or NaN per frame
    const P2 p256 = mul2(bc(4.f), fma2(a[8].re, a[8].re, mul2(a[8].im, a[8].im)));   // lane 0: bin 256 = conj Z[256]

    // ---- next pair of this half-warp
    long long nfa = fa + step, noff = off + d_off;
    int nclip = clip + step_clip, ntt = tt + step_t;
    if (ntt >= fpc) { ntt -= fpc; ++nclip; noff += wrap_off; }
    const bool has_next = nfa < g.total_frames;
    const bool nxt_fast = has_next && is_fast(nfa, ntt, noff);
    const float2* nsrc = nxt_fast ? reinterpret_cast<const float2*>(g.pcm + noff) + t : idle_src;

    // ---- untangle, in-lane: the lower halves of both columns lead (k < 256); their mirrors are the upper halves.
    //      General lane: Z[512 - (ka + 32 q)] = column kb, element 15 - q, and vice versa.
    //      Lane 0: column 0 mirrors into itself (element 16 - q), column 16 into itself (element 15 - q).
    //      Lane 0's mirrors are first moved to where the general rule looks (in place, descending q keeps every
    //      source intact until it is read), so the steps below carry no per-step selects or temporaries.
    static_for<0, 8>([&](auto qq) {
      constexpr int q = 7 - decltype(qq)::value;
      const C2 na = a[q ? 16 - q : 0], nb = a[31 - q];
      a[31 - q].re = P2(c0 ? na.re.v.x : a[31 - q].re.v.x, c0 ? na.re.v.y : a[31 - q].re.v.y);
      a[31 - q].im = P2(c0 ? na.im.v.x : a[31 - q].im.v.x, c0 ? na.im.v.y : a[31 - q].im.v.y);
      a[15 - q].re = P2(c0 ? nb.re.v.x : a[15 - q].re.v.x, c0 ? nb.re.v.y : a[15 - q].re.v.y);
      a[15 - q].im = P2(c0 ? nb.im.v.x : a[15 - q].im.v.x, c0 ? nb.im.v.y : a[15 - q].im.v.y);
    });
    P2 pk[16], pm[16];
    static_for<0, 8>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const C2 zma = a[31 - q], zmb = a[15 - q];
      auto pair = [&](const C2& zk, const C2& zm, int k, P2& opk, P2& opm) {
        const float2 w = s_ut[k];
        const P2 ex = add2(zk.re, zm.re), ey = add2(zk.im, neg(zm.im));      // 2E
        const P2 ox = add2(zk.im, zm.im), oy = add2(zm.re, neg(zk.re));      // 2O
        const P2 xr = fma2(ox, bc(w.x), fma2(oy, bc(-w.y), ex));             // 2X[k]
        const P2 xi = fma2(ox, bc(w.y), fma2(oy, bc(w.x), ey));
        const P2 yr = fma2(ex, bc(2.f), neg(xr));                            // 2 conj X[512-k]
        const P2 yi = fma2(ey, bc(2.f), neg(xi));
        opk = fma2(xr, xr, mul2(xi, xi));
        opm = fma2(yr, yr, mul2(yi, yi));
      };
      pair(a[q], zma, ka + 32 * q, pk[2 * q], pm[2 * q]);
      pair(a[16 + q], zmb, kb + 32 * q, pk[2 * q + 1], pm[2 * q + 1]);
      if constexpr (q == 0) pm[0] = P2(c0 ? p256.v.x : pm[0].v.x, c0 ? p256.v.y : pm[0].v.y);
      // this step's share of the next pair's loads
      // (skewed towards the late steps: early on the FFT registers are still live)
      static_for<(NLOAD * q * (q + 1)) / 72, (NLOAD * (q + 1) * (q + 2)) / 72>([&](auto mm) {
        constexpr int m = decltype(mm)::value;
        s[m] = ldg_nc_f2(nsrc + 16 * m);
      });
    });

    // ---- epilogue
    const bool has_b_out = fa + 1 < g.total_frames;
    T* __restrict__ row_a = out + fa * (long long)kP16M;
    T* __restrict__ row_b = row_a + kP16M;
    auto bins_of = [&](int i, int& k, int& mk) {   // slot i = 2q (column ka) or 2q+1 (column kb)
      k = ((i & 1) ? kb : ka) + 32 * (i >> 1);
      mk = kP16M - k;
      if (i == 0 && c0) mk = kP16M / 2;
    };
    if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
      const P2 scale = add2(bc(ep.byte_a), poison);
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(i, k, mk);
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), scale, bc(ep.byte_b));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), scale, bc(ep.byte_b));
        const unsigned kA = byte_of_scaled(vk.v.x), kB = byte_of_scaled(vk.v.y);
        const unsigned mA = byte_of_scaled(vm.v.x), mB = byte_of_scaled(vm.v.y);
        if constexpr (OUT == kOutU8) {
          sb16[k] = (uint16_t)(kA | (kB << 8));
          sb16[mk] = (uint16_t)(mA | (mB << 8));
        } else if (alive) {
          row_a[k] = __ldg(ep.lut + kA); row_a[mk] = __ldg(ep.lut + mA);
          if (has_b_out) { row_b[k] = __ldg(ep.lut + kB); row_b[mk] = __ldg(ep.lut + mB); }
        }
      });
      if constexpr (OUT == kOutU8) {
        __syncwarp();
        // de-interleave the (A,B) byte pairs of this pair: 8 bins per lane per round, 8-byte row stores
        const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
        uint2* ra = reinterpret_cast<uint2*>(row_a);
        uint2* rb = reinterpret_cast<uint2*>(row_b);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 w = s16[c * 16 + t];
          if (alive) ra[c * 16 + t] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
          if (has_b_out) rb[c * 16 + t] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
        }
      }
    } else {
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(i, k, mk);
        if (alive) {
          row_a[k] = emit_power<OUT>(pk[i].v.x, ep); row_a[mk] = emit_power<OUT>(pm[i].v.x, ep);
          if (has_b_out) { row_b[k] = emit_power<OUT>(pk[i].v.y, ep); row_b[mk] = emit_power<OUT>(pm[i].v.y, ep); }
        }
      });
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, has_next)) break;     // both half-warps leave together (the exchange barrier is per warp)
    fa = nfa; off = noff; clip = nclip; tt = ntt;
    cur_fast = nxt_fast;
  }
}

}  // namespace sg
